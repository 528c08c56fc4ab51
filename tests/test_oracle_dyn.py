"""Oracle checks for the rows the reference has NO golden vectors for (numerical flux, tendency,
pressure, boundary condition, full step are 'parity unpinned'): physical invariants that any correct restatement of
the reference's equations must satisfy."""
import numpy as np
import pytest

from cases import DensityCurrentCase, C0
from fe_project_b200 import initcond
from fe_project_b200.setup_aux import calc_phyd_hgrad


def _rest_case(**kw):
    case = DensityCurrentCase(p=7, NeX=3, NeY=2, NeZ=3, **kw)
    for k in ("DDENS", "DRHOT", "MOMX", "MOMY", "MOMZ"):
        case.fields[k][:] = 0.0
    return case


def test_hydrostatic_rest_state_has_tiny_tendency():
    """A resting, hydrostatically balanced atmosphere: DENS/RHOT/MOMX/MOMY tendencies vanish identically; the MOMZ
    tendency is -g*VFilterPM1(0) - d(DPRES)/dz = 0 because DPRES == 0 (the balance lives in the hyd fields)."""
    case = _rest_case()
    o = case.make_oracle()
    for w in ("exchange", "pressure", "bc", "tend_ex"):
        o.piece(w)
    n = case.mesh.Ne * case.elem.Np
    te = o.arr("tend_ex").reshape(5, -1)[:, :n]
    dpres = o.arr("DPRES")[:n]
    assert np.abs(dpres).max() < 1e-9 * 1e5          # pow round-off on 1e5 Pa
    assert np.abs(te[0]).max() == 0.0 and np.abs(te[1]).max() == 0.0
    assert np.abs(te[2]).max() < 1e-10               # momz: derivative of round-off pressure noise


def test_mass_and_theta_conservation_periodic():
    """Fully periodic box, no gravity-induced boundary fluxes: sum(w J q) of DDENS and DRHOT is conserved to round-off
    by the DG scheme with a single-valued numerical flux."""
    case = DensityCurrentCase(p=3, NeX=4, NeY=4, NeZ=4, dom=(0, 8e3, 0, 8e3, 0, 8e3), periodic=(True, True, True),
                              perturb=3.0, modalfilter=False, dt=0.5, intrp_order=5)
    # uniform background so that the vertical periodic wrap is consistent
    f = case.fields
    Ne = case.mesh.Ne
    f["DENS_hyd"][:Ne] = 1.2; f["PRES_hyd"][:Ne] = 1.0e5
    f["DDENS"][:Ne] = 0.01 * np.sin(2 * np.pi * case.mesh.pos_en[0] / 8e3)
    f["DRHOT"][:Ne] = 3.0 * np.cos(2 * np.pi * case.mesh.pos_en[2] / 8e3)
    case.vel_bc = {}
    o = case.make_oracle()
    o.set_consts(dict(C0, GRAV=0.0))
    w = (np.tile(case.elem.IntWeight_lgl, Ne) * case.mesh.J.reshape(-1))
    n = Ne * case.elem.Np
    m0 = [np.sum(w * o.arr(k)[:n]) for k in ("DDENS", "DRHOT", "MOMX", "MOMY", "MOMZ")]
    o.update(10)
    m1 = [np.sum(w * o.arr(k)[:n]) for k in ("DDENS", "DRHOT", "MOMX", "MOMY", "MOMZ")]
    vol = w.sum()
    for a, b, scale in zip(m0, m1, (1.2, 360.0, 3.0, 3.0, 3.0)):
        assert abs(a - b) < 1e-12 * vol * scale
    assert np.abs(o.arr("MOMX")[:n]).max() > 0.1     # something actually happened


def test_slip_wall_reflects_normal_momentum():
    case = DensityCurrentCase(p=3, NeX=2, NeY=2, NeZ=2, perturb=1.0, periodic=(False, False, False), intrp_order=5)
    o = case.make_oracle()
    o.piece("exchange"); o.piece("bc")
    m, e = case.mesh, case.elem
    nint = m.Ne * e.Np
    for f, (nm, sgn) in enumerate((("MOMY", 1), ("MOMX", 1), ("MOMY", 1), ("MOMX", 1), ("MOMZ", 1), ("MOMZ", 1))):
        sl = slice(nint + m.halo_face_off[f], nint + m.halo_face_off[f] + m.halo_face_size[f])
        src = m.VMapB[m.halo_face_off[f]: m.halo_face_off[f] + m.halo_face_size[f]]
        for var in ("MOMX", "MOMY", "MOMZ"):
            a = o.arr(var)
            if var == nm:
                assert np.array_equal(a[sl], a[src] - 2.0 * a[src])     # normal component reflected
            else:
                assert np.array_equal(a[sl], a[src])
        assert np.array_equal(o.arr("DDENS")[sl], o.arr("DDENS")[src])


def test_phyd_hgrad_two_restatements():
    """DPhydDx/DPhydDy (common.F90:624-777): C++ oracle vs NumPy set-up code, on a background with a horizontal
    pressure gradient so that the result is not trivially zero."""
    case = DensityCurrentCase(p=4, NeX=3, NeY=3, NeZ=2, intrp_order=5, periodic=(False, False, False))
    x, y = case.mesh.pos_en[0], case.mesh.pos_en[1]
    Ne = case.mesh.Ne
    case.fields["PRES_hyd"][:Ne] *= (1.0 + 1e-3 * np.sin(x / 5e3) * np.cos(y / 4e3))
    o = case.make_oracle()
    gx, gy = calc_phyd_hgrad(case.elem, case.mesh, case.fields["PRES_hyd"])
    n = Ne * case.elem.Np
    ax, ay = o.arr("DPhydDx")[:n], o.arr("DPhydDy")[:n]
    assert np.abs(ax).max() > 1e-3
    assert np.linalg.norm(gx.reshape(-1)[:n] - ax) <= 1e-11 * np.linalg.norm(ax)  # cancellation on a 1e5 Pa field
    assert np.linalg.norm(gy.reshape(-1)[:n] - ay) <= 1e-11 * np.linalg.norm(ay)
    # analytic check away from element-boundary jumps: d/dx of the smooth field
    P = case.fields["PRES_hyd"][:Ne]
    ana = None  # the DG gradient of a smooth field converges to the analytic one; 4th order elements on 8.5 km cells
    hyd = initcond.hydrostatic_const_pt(case.mesh.pos_en[2], 300.0, C0["PRES00"])[1]
    ana = hyd * 1e-3 * np.cos(x / 5e3) / 5e3 * np.cos(y / 4e3)
    assert np.abs(ax - ana.reshape(-1)).max() < 2e-2 * np.abs(ana).max()


def test_full_step_energy_budget_is_sane():
    """Density current, closed box with slip walls: mass is conserved to round-off, total energy drifts only by the
    (dissipative) Rusanov flux + modal filter."""
    case = DensityCurrentCase(p=7, NeX=4, NeY=1, NeZ=2, dom=(0.0, 12.8e3, 0.0, 3.2e3, 0.0, 6.4e3), dt=0.05)
    o = case.make_oracle()
    m0 = o.monitor()
    o.update(20)
    m1 = o.monitor()
    vol = 12.8e3 * 3.2e3 * 6.4e3
    assert abs(m1[0] - m0[0]) < 1e-12 * vol          # DDENS integral
    assert abs(m1[1] - m0[1]) < 1e-9 * abs(m0[1])    # ENGT
    assert m1[2] > m0[2]                             # kinetic energy grows as the cold bubble sinks
