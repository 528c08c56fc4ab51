"""Oracle checks of the HEVI rows (a6, a8-a12).  The reference holds no golden vectors for them (SURVEY.md 8c:
parity unpinned), so the restatement is pinned by identities that tie it to the HEVE rows and to itself:
 * operator splitting: explicit HEVI tendency + vertical operator (cal_vi with impl_fac = 0) == HEVE tendency;
 * the Newton step of cal_vi reduces the nonlinear residual by orders of magnitude, the linear solve (block
   Thomas + partial-pivot LU) solves its own Jacobian system;
 * an IMEX run converges to the explicit HEVE run as dt -> 0 and stays stable at vertical acoustic CFL > 1."""
import numpy as np
import pytest

from cases import DensityCurrentCase, C0, rel_l2

ORD = ("DDENS", "DRHOT", "MOMZ", "MOMX", "MOMY")   # oracle variable order


def _state(o, n):
    return np.stack([o.arr(k).copy() for k in ORD])


def test_splitting_identity():
    case = DensityCurrentCase(p=7, NeX=3, NeY=2, NeZ=3, perturb=2.0)
    oe = case.make_oracle()
    for w in ("exchange", "pressure", "bc", "tend_ex"):
        oe.piece(w)
    n = case.mesh.Ne * case.elem.Np
    te = oe.arr("tend_ex").reshape(5, -1)[:, :n].copy()
    case_i = DensityCurrentCase(p=7, NeX=3, NeY=2, NeZ=3, perturb=2.0, eqs="NONHYDRO3D_HEVI", tinteg="IMEX_ARK232")
    oi = case_i.make_oracle()
    for w in ("exchange", "pressure", "bc", "tend_ex"):
        oi.piece(w)
    N = case.mesh.NeA * case.elem.Np
    ti = oi.arr("tend_ex")[:5 * N].reshape(5, -1)[:, :n].copy()   # IMEX keeps one buffer per stage: take the first
    var0 = _state(oi, n)
    tv = oi.cal_vi(0.0, case.dt, var0)[:, :n]
    for v, nm in enumerate(ORD):
        assert rel_l2(ti[v] + tv[v], te[v]) < 1e-11, nm
    assert np.abs(tv[2]).max() > 1e-4 and np.abs(ti[3]).max() > 1e-4


@pytest.mark.parametrize("impl_fac", [0.02, 0.5])
def test_newton_step_reduces_residual(impl_fac):
    """G(q) = q - q_cur + impl_fac * A_v(q).  cal_vi returns (q* - q_cur)/impl_fac after one Newton step from var0."""
    case = DensityCurrentCase(p=7, NeX=2, NeY=2, NeZ=4, perturb=0.5, eqs="NONHYDRO3D_HEVI", tinteg="IMEX_ARK232")
    o = case.make_oracle()
    n = case.mesh.Ne * case.elem.Np
    qcur = _state(o, n)
    var0 = qcur.copy()
    rng = np.random.default_rng(0)
    var0[:, :n] += 1e-3 * rng.standard_normal((5, n)) * np.abs(qcur[:, :n]).max(axis=1, keepdims=True)
    Av0 = -o.cal_vi(0.0, 1.0, var0)[:, :n]                       # A_v(var0) evaluated about var0 (alph frozen at var0)
    G0 = var0[:, :n] - qcur[:, :n] + impl_fac * Av0
    t = o.cal_vi(impl_fac, 1.0, var0)
    qstar = qcur.copy(); qstar[:, :n] += impl_fac * t[:, :n]
    # residual at q*: evaluate A_v at q* with the dissipation coefficient still frozen at var0 -> use the oracle with
    # current state = q*, var0 = var0: tendencies for impl_fac = 0 are evaluated on PROG_VARS = var0, so instead
    # check the linear system directly: G0 + J (q* - var0) = 0  <=>  one more Newton step from q* changes little
    for k, nm in enumerate(ORD):
        o.arr(nm)[:] = qcur[k]
    t2 = o.cal_vi(impl_fac, 1.0, qstar)
    q2 = qcur.copy(); q2[:, :n] += impl_fac * t2[:, :n]
    d1 = np.linalg.norm(qstar[:, :n] - var0[:, :n]); d2 = np.linalg.norm(q2[:, :n] - qstar[:, :n])
    assert np.linalg.norm(G0) > 0 and d2 < 1e-3 * d1, (d1, d2)


def test_imex_converges_to_explicit():
    """Same initial state, T = 0.8 s: HEVI/IMEX_ARK324 with dt -> dt/2 approaches HEVE/ERK_SSP_3s3o (tiny dt)."""
    kw = dict(p=3, NeX=4, NeY=2, NeZ=4, dom=(0.0, 3.2e3, 0.0, 1.6e3, 0.0, 3.2e3), perturb=1.0, modalfilter=False, intrp_order=5)
    ref = DensityCurrentCase(dt=0.0125, tinteg="ERK_SSP_3s3o", **kw).make_oracle()
    ref.update(64)
    n = 4 * 2 * 4 * 64
    errs = []
    for dt, ns in ((0.1, 8), (0.05, 16)):
        o = DensityCurrentCase(dt=dt, tinteg="IMEX_ARK324", eqs="NONHYDRO3D_HEVI", **kw).make_oracle()
        o.update(ns)
        errs.append(max(rel_l2(o.arr(k)[:n], ref.arr(k)[:n]) for k in ("MOMX", "MOMZ", "DRHOT")))
    assert errs[1] < 0.3 * errs[0] and errs[1] < 1e-4, errs


def test_vertical_acoustic_cfl_above_one_is_stable():
    """Config 2 analogue (sample/euler3d_hevi: column, GRAV = 0, acoustic pulse): dz_node ~ 3 m, c_s ~ 347 m/s,
    dt = 0.1 s -> vertical acoustic CFL ~ 10; the explicit scheme would blow up, HEVI must stay bounded and
    conserve mass."""
    case = DensityCurrentCase(p=7, NeX=1, NeY=1, NeZ=8, dom=(0.0, 40.0e3, 0.0, 40.0e3, 0.0, 400.0), dt=0.1, modalfilter=False,
                              eqs="NONHYDRO3D_HEVI", tinteg="IMEX_ARK232", periodic=(True, True, False), intrp_order=9)
    f = case.fields
    Ne = case.mesh.Ne
    z = case.mesh.pos_en[2]
    f["DENS_hyd"][:Ne] = 1.2; f["PRES_hyd"][:Ne] = 1.0e5
    r = (z - 200.0) / 40.0
    f["DDENS"][:Ne] = 0.0; f["DRHOT"][:Ne] = np.where(np.abs(r) <= 1, 1e-2 * np.cos(0.5 * np.pi * r) ** 2, 0.0)
    o = case.make_oracle()
    o.set_consts(dict(C0, GRAV=0.0))
    o.prepare()
    w = np.tile(case.elem.IntWeight_lgl, Ne) * case.mesh.J.reshape(-1)
    n = Ne * case.elem.Np
    m0 = np.sum(w * o.arr("DRHOT")[:n])
    o.update(40)
    assert np.isfinite(o.arr("MOMZ")).all()
    assert np.abs(o.arr("DRHOT")[:n]).max() < 2e-2 and np.abs(o.arr("MOMZ")[:n]).max() < 5.0
    assert abs(np.sum(w * o.arr("DRHOT")[:n]) - m0) < 1e-12 * w.sum() * 360.0
    assert np.abs(o.arr("MOMX")[:n]).max() < 1e-12


def test_sound_wave_config2_speed_and_split():
    """BASELINE config 2 (sample/euler3d_hevi sound wave, p = 7 column) through the library HEVI path: the DRHOT pulse
    splits into two halves of amplitude A/2 travelling at c_s = sqrt(gamma p / rho) (linear acoustics), momentum stays
    vertical, and the DRHOT integral is conserved."""
    from cases import SoundWaveCase
    A = 1.0e-3
    case = SoundWaveCase(p=7, NeZ=40, dt=0.25, amplitude=A)   # horizontal explicit part: dt < ~0.3 s at p = 7 on a 10 km element
    o = case.make_oracle()
    Ne, Np = case.mesh.Ne, case.elem.Np
    n = Ne * Np
    w = np.tile(case.elem.IntWeight_lgl, Ne) * case.mesh.J.reshape(-1)
    m0 = np.sum(w * o.arr("DRHOT")[:n])
    nstep = 24
    o.update(nstep)
    z = case.mesh.pos_en[2].reshape(-1)
    dr = o.arr("DRHOT")[:n]
    cs = np.sqrt(C0["CPdry"] / C0["CVdry"] * 1.0e5 / 1.0)
    up = z > 5.0e3
    zpk = z[up][np.argmax(dr[up])]
    assert abs(zpk - (5.0e3 + cs * nstep * case.dt)) < 60.0, zpk
    assert abs(dr[up].max() - 0.5 * A) < 0.02 * A
    assert abs(dr[~up].max() - 0.5 * A) < 0.02 * A
    assert abs(np.sum(w * dr) - m0) < 1e-12 * w.sum() * 350.0
    assert np.abs(o.arr("MOMX")[:n]).max() < 1e-8 * np.abs(o.arr("MOMZ")[:n]).max()
