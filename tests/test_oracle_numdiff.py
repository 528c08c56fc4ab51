"""Oracle checks for the numerical diffusion (fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_numdiff.F90).
The reference holds no golden vectors for it; the restatement is pinned by the analytic action of (-1)^(n+1) nu Laplacian^n on
a Fourier mode (spectral convergence), by conservation in a periodic box and by the boundary rules (nothing leaks through
adiabatic slip walls)."""
import numpy as np
import pytest

from cases import DensityCurrentCase


def _box(p, nex, periodic=(True, True, True), dom=(0.0, 4.0e3, 0.0, 1.0e3, 0.0, 1.0e3)):
    return DensityCurrentCase(p=p, NeX=nex, NeY=1, NeZ=1, dom=dom, periodic=periodic, perturb=0.0, modalfilter=False, dt=0.1,
                              intrp_order=p + 2)


@pytest.mark.parametrize("lap,nu,cases", [(1, 50.0, ((5, 4, 1e-2), (7, 4, 1e-4))), (2, 5.0e6, ((7, 8, 2e-2), (7, 16, 2e-3)))])
def test_fourier_mode_decay_rate(lap, nu, cases):
    k = 2 * np.pi / 4.0e3
    for p, nex, tol in cases:
        c = _box(p, nex)
        o = c.make_oracle()
        n = c.mesh.Ne * c.elem.Np
        x = c.mesh.pos_en[0].reshape(-1)
        o.arr("DDENS")[:n] = 1e-3 * np.sin(k * x)
        o.set_numdiff(True, lap, nu, nu)
        before = o.arr("DDENS")[:n].copy()
        o.numdiff_apply()
        rate = (o.arr("DDENS")[:n] - before) / c.dt
        expected = -(nu * k ** (2 * lap)) * before
        assert np.abs(rate - expected).max() <= tol * np.abs(expected).max(), (p, nex)


def test_density_weighted_variables_diffuse_the_specific_quantity():
    """MOMX = rho * u with u a Fourier mode and rho non-uniform: d(MOMX)/dt = div(nu rho grad u)."""
    c = _box(7, 8)
    o = c.make_oracle()
    n = c.mesh.Ne * c.elem.Np
    x = c.mesh.pos_en[0].reshape(-1)
    k = 2 * np.pi / 4.0e3
    rho = (o.arr("DENS_hyd")[:n] + o.arr("DDENS")[:n]).copy()
    o.arr("DDENS")[:n] += 0.1 * np.cos(k * x)
    rho = o.arr("DENS_hyd")[:n] + o.arr("DDENS")[:n]
    o.arr("MOMX")[:n] = rho * 3.0 * np.sin(k * x)
    nu = 80.0
    o.set_numdiff(True, 1, nu, nu)
    before = o.arr("MOMX")[:n].copy()
    rho0 = rho.copy()
    o.numdiff_apply()
    rate = (o.arr("MOMX")[:n] - before) / c.dt
    # d/dx (nu rho du/dx), rho = rho_h(z) + 0.1 cos(kx): z-dependence of rho_h does not matter for an x-only u
    drho = -0.1 * k * np.sin(k * x)
    expected = nu * (drho * 3.0 * k * np.cos(k * x) - rho0 * 3.0 * k * k * np.sin(k * x))
    assert np.abs(rate - expected).max() <= 2e-3 * np.abs(expected).max()


def test_conservation_and_walls():
    c = DensityCurrentCase(p=4, NeX=3, NeY=2, NeZ=2, perturb=2.0, modalfilter=False, dt=0.1, periodic=(False, True, False), intrp_order=6)
    o = c.make_oracle()
    n = c.mesh.Ne * c.elem.Np
    w = np.tile(c.elem.IntWeight_lgl, c.mesh.Ne) * c.mesh.J.reshape(-1)
    o.set_numdiff(True, 1, 75.0, 75.0, therm_bc=(1, 1, 1, 1, 1, 1))
    m0 = {k: np.sum(w * o.arr(k)[:n]) for k in ("DDENS", "DRHOT")}
    o.numdiff_apply()
    for k in ("DDENS", "DRHOT"):      # adiabatic slip walls in x and z, periodic y: nothing leaves the box
        assert abs(np.sum(w * o.arr(k)[:n]) - m0[k]) <= 1e-10 * w.sum() * max(1.0, abs(m0[k]) / w.sum())


def test_step_with_numdiff_differs_and_stays_finite():
    a = DensityCurrentCase(p=4, NeX=3, NeY=2, NeZ=2, perturb=2.0, dt=0.05, intrp_order=6)
    o1, o2 = a.make_oracle(), a.make_oracle()
    o2.set_numdiff(True, 1, 75.0, 75.0, therm_bc=(1, 0, 1, 0, 1, 1))
    o1.update(3); o2.update(3)
    n = a.mesh.Ne * a.elem.Np
    assert np.isfinite(o2.arr("MOMX")[:n]).all()
    assert np.abs(o1.arr("MOMX")[:n] - o2.arr("MOMX")[:n]).max() > 1e-8
