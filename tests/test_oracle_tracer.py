"""Oracle checks of the tracer-advection restatement (oracle/tracer.cpp; no device kernel yet).
The reference has no golden vectors for this path (model test case `tracer_advection` only writes history files), so the
restatement is held to properties: a uniform mixing ratio stays uniform, tracer mass is conserved to round-off with and
without the limiters, the FCT + TMAR limiters keep a non-negative field non-negative where the unlimited scheme undershoots,
and -- limiter off, uniform density -- the result equals the independently restated sample/advect3d kernel (row a18)."""
import numpy as np
import pytest

from fe_project_b200.advect3d import gaussian_hill
from fe_project_b200.element import HexElement
from fe_project_b200.mesh import LocalMeshCube
from oracle_api import Oracle, OracleAdvect3D

DOM = (0, 1, 0, 1, 0, 1)
PER = (True, True, True)


def _state(p, ne, dens=None, vel=(0.5, 0.3, -0.2)):
    o = Oracle(p, ne, ne, ne, DOM, periodic=PER)
    mesh = LocalMeshCube(HexElement(p), ne, ne, ne, *DOM, periodic=PER)
    n = o.Np * o.Ne
    rho = np.ones(n) if dens is None else dens(mesh).reshape(-1)
    o.arr("DENS_hyd")[:n] = 1.0
    o.arr("DDENS")[:n] = rho - 1.0
    for nm, v in zip(("MOMX", "MOMY", "MOMZ"), vel):
        o.arr(nm)[:n] = v                      # constant mass flux: divergence-free for any density
    w = np.tile(HexElement(p).IntWeight_lgl, mesh.Ne) * mesh.J.reshape(-1)
    return o, mesh, n, rho, w


def _q(o, n, vals):
    q = np.zeros(o.Np * o.NeA)
    q[:n] = vals
    return q


@pytest.mark.parametrize("limiter_off", [True, False])
def test_uniform_mixing_ratio_is_preserved(limiter_off):
    o, mesh, n, rho, w = _state(3, 3, dens=lambda m: 1.0 + 0.2 * np.sin(2 * np.pi * m.pos_en[0]) * np.cos(2 * np.pi * m.pos_en[2]))
    q = _q(o, n, 0.7)
    o.trcadv_update(q, "ERK_SSP_3s3o", 0.01, nsteps=5, disable_limiter=limiter_off)
    assert np.abs(q[:n] - 0.7).max() <= 1e-13


@pytest.mark.parametrize("limiter_off,mf", [(True, None), (False, None), (False, (0.0, 1.0, 16, 0.0, 1.0, 16))])
def test_tracer_mass_is_conserved(limiter_off, mf):
    o, mesh, n, rho, w = _state(3, 4, dens=lambda m: 1.0 + 0.3 * np.sin(2 * np.pi * m.pos_en[1]))
    q = _q(o, n, gaussian_hill(mesh, 0.4, 0.5, 0.5, width=0.12)[:mesh.Ne].reshape(-1))
    m0 = np.sum(w * rho * q[:n])
    o.trcadv_update(q, "ERK_SSP_3s3o", 0.01, nsteps=10, modalfilter=mf, disable_limiter=limiter_off)
    assert abs(np.sum(w * rho * q[:n]) - m0) <= 2e-14 * abs(m0)
    assert np.isfinite(q[:n]).all()


def test_limiters_keep_the_tracer_non_negative():
    """A sharp box profile: the unlimited high-order scheme undershoots, FCT (net outward flux limited by the available mass) +
    TMAR (truncation and mass-aware rescaling) do not."""
    o, mesh, n, rho, w = _state(3, 5, vel=(1.0, 0.0, 0.0))
    x, y, z = (mesh.pos_en[k].reshape(-1) for k in range(3))
    box = ((np.abs(x - 0.3) < 0.15) & (np.abs(y - 0.5) < 0.2) & (np.abs(z - 0.5) < 0.2)).astype(float)
    q_lim, q_raw = _q(o, n, box), _q(o, n, box)
    o.trcadv_update(q_raw, "ERK_SSP_3s3o", 0.004, nsteps=40, disable_limiter=True)
    o.trcadv_update(q_lim, "ERK_SSP_3s3o", 0.004, nsteps=40, disable_limiter=False)
    assert q_raw[:n].min() < -1e-3                    # Gibbs undershoot of the unlimited scheme
    assert q_lim[:n].min() >= 0.0                     # exactly non-negative after TMAR
    m0 = np.sum(w * box)
    assert abs(np.sum(w * q_lim[:n]) - m0) <= 1e-13 * m0
    # the limited solution has moved with the flow: centre of mass advanced by u t = 0.16
    xc = np.sum(w * q_lim[:n] * x) / np.sum(w * q_lim[:n])
    assert abs(xc - (0.3 + 0.16)) < 0.02


@pytest.mark.parametrize("p,ne", [(3, 4), (7, 2)])
def test_equals_the_advect3d_restatement_for_uniform_density(p, ne):
    """rho = 1, constant velocity, limiter and filter off: the tracer equation is the scalar advection equation of sample/advect3d
    with the same upwind flux, the same Div operator and the same low-storage integrator -- two restatements of different
    reference files (trcadvect3d_heve.F90 vs mod_advect3d_kernel.f90) must agree to round-off."""
    vel = (0.5, 0.5, 0.5)
    o, mesh, n, rho, w = _state(p, ne, vel=vel)
    q0 = gaussian_hill(mesh, 0.3, 0.4, 0.5, width=0.15)[:mesh.Ne].reshape(-1)
    q = _q(o, n, q0)
    dt, nsteps = 0.004, 25
    o.trcadv_update(q, "ERK_SSP_3s3o", dt, nsteps=nsteps, disable_limiter=True)
    o2 = Oracle(p, ne, ne, ne, DOM, periodic=PER)
    a = OracleAdvect3D(o2, "ERK_SSP_3s3o", dt)
    a.arr("q")[:n] = q0
    for nm, v in zip("uvw", vel):
        a.arr(nm)[:n] = v
    a.update(nsteps)
    ref = a.arr("q")[:n]
    assert np.abs(q[:n] - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.abs(q[:n] - q0).max() > 1e-2            # and it did move


def test_fct_coefficient_is_one_where_nothing_is_at_risk():
    """With plenty of tracer everywhere (q >= 1) the admissible outflow exceeds the actual one: the limited and the unlimited
    runs coincide (fct_coef = 1, TMAR is the identity on a positive field)."""
    o, mesh, n, rho, w = _state(3, 3)
    q0 = 2.0 + gaussian_hill(mesh, 0.5, 0.5, 0.5, width=0.2)[:mesh.Ne].reshape(-1)
    qa, qb = _q(o, n, q0), _q(o, n, q0)
    o.trcadv_update(qa, "ERK_SSP_3s3o", 0.005, nsteps=5, disable_limiter=True)
    o.trcadv_update(qb, "ERK_SSP_3s3o", 0.005, nsteps=5, disable_limiter=False)
    assert np.abs(qa[:n] - qb[:n]).max() <= 1e-13


# ------------------------------------------------------------------------------ coupled to the dynamics (save_massflux)
def _coupled(eqs, tinteg, dt):
    from cases import DensityCurrentCase
    case = DensityCurrentCase(p=3, NeX=3, NeY=2, NeZ=3, perturb=2.0, dt=dt, tinteg=tinteg, eqs=eqs, modalfilter=False, intrp_order=7)
    o = case.make_oracle()
    o.set_tracer_coupling(True)
    n = case.mesh.Ne * case.elem.Np
    w = np.tile(case.elem.IntWeight_lgl, case.mesh.Ne) * case.mesh.J.reshape(-1)
    return case, o, n, w


@pytest.mark.parametrize("tinteg", ["ERK_SSP_3s3o", "ERK_SSP_4s3o"])
def test_tracer_mass_consistency_with_the_heve_dynamics(tinteg):
    """The sharpest check of the coupling: a uniform mixing ratio advected with the stage-averaged mass flux and dissipation
    coefficient that the dynamics stages saved (save_massflux / cal_alphdens_dyn with the Butcher weights b_ex) must follow the
    density the dynamics produced EXACTLY -- it holds only if the Rusanov mass flux of the dynamics, the weights, the
    density-weighted low-storage integrator and the boundary condition on the mass flux are all restated consistently."""
    case, o, n, w = _coupled("NONHYDRO3D_HEVE", tinteg, 0.2)
    q = np.zeros(o.Np * o.NeA); q[:n] = 0.7
    for _ in range(3):
        o.update(1)
        o.trcadv_update_coupled(q, "ERK_SSP_3s3o", 0.2, disable_limiter=True)
    assert np.abs(q[:n] - 0.7).max() <= 1e-14
    assert np.abs(o.arr("DDENS")[:n]).max() > 1e-3                     # over a density field that really moved


@pytest.mark.parametrize("eqs,tinteg,dt", [("NONHYDRO3D_HEVE", "ERK_SSP_4s3o", 0.2), ("NONHYDRO3D_HEVI", "IMEX_ARK232", 0.5)])
def test_coupled_tracer_mass_is_conserved(eqs, tinteg, dt):
    """sum w J rho q over the closed domain (slip walls, periodic y) is the same before and after, limiter on, for both equation
    sets.  (With HEVI a uniform q is NOT preserved to round-off, by construction of the reference: cal_alphdens_dyn drops the
    vertical faces' acoustic dissipation -- Gnn has no |nz| term -- while the implicit solver applies it frozen at var0.)"""
    case, o, n, w = _coupled(eqs, tinteg, dt)
    x, z = case.mesh.pos_en[0].reshape(-1), case.mesh.pos_en[2].reshape(-1)
    q = np.zeros(o.Np * o.NeA); q[:n] = 1.0 + 0.5 * np.sin(x / 3e3) * np.cos(z / 1e3)
    rho0 = o.arr("DENS_hyd")[:n] + o.arr("DDENS")[:n]
    m0 = np.sum(w * rho0 * q[:n])
    for _ in range(3):
        o.update(1)
        o.trcadv_update_coupled(q, "ERK_SSP_3s3o", dt, disable_limiter=False)
    rho1 = o.arr("DENS_hyd")[:n] + o.arr("DDENS")[:n]
    assert abs(np.sum(w * rho1 * q[:n]) - m0) <= 1e-14 * abs(m0)
    assert 0.4 < q[:n].min() and q[:n].max() < 1.6
