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
