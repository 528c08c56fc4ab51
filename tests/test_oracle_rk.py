"""Oracle pins, timeint_rk (a13): FElib/test/common/timeint_rk/test_timeint_rk.f90 restated --
convergence orders of the 7 ERK + 2 IMEX schemes on the (damped) harmonic oscillator (:33-43, 108-280) and the
Shu-Osher -> Butcher conversion against the hand-written tableaux (:284-357, tolerance 1e-14*nstage^2)."""
import numpy as np
import pytest

from oracle_api import OracleRK, rk_tables

END, DT1, DT2 = 2.0, 0.02, 0.0025


def _run_ex(name, dt, dstep):
    t = OracleRK(name, dt, 2, 1)
    omg = 2.0 * np.pi
    u, v = np.ones(1), np.zeros(1)
    err = []
    nstep = int(END / dt + 1e-9)
    for n in range(1, nstep + 1):
        for s in range(t.info["nstage"]):
            t.tend(0, 0, s)[0] = omg * v[0]
            t.tend(0, 1, s)[0] = -omg * u[0]
            t.advance(s, u, 0)
            t.advance(s, v, 1)
        if n % dstep == 0:
            err.append(u[0] - np.cos(omg * n * dt))
    return np.array(err)


def _run_imex(name, dt, dstep):
    t = OracleRK(name, dt, 2, 1)
    omg, r = 2.0 * np.pi, 0.1
    omgg = np.sqrt(omg ** 2 - r ** 2 / 4.0)
    u, v = np.ones(1), np.zeros(1)
    err = []
    nstep = int(END / dt + 1e-9)
    for n in range(1, nstep + 1):
        for s in range(t.info["nstage"]):
            fac = t.implicit_fac(s)
            coef = fac * omg
            if abs(fac) > 0.0:
                ui = (u[0] + coef * v[0]) / (1.0 + coef ** 2)
                vi = (v[0] - coef * u[0]) / (1.0 + coef ** 2)
                t.tend(1, 0, s)[0] = (ui - u[0]) / fac
                t.tend(1, 1, s)[0] = (vi - v[0]) / fac
            else:
                t.tend(1, 0, s)[0] = omg * v[0]
                t.tend(1, 1, s)[0] = -omg * u[0]
            t.store_implicit(s, u, 0)
            t.store_implicit(s, v, 1)
            t.tend(0, 0, s)[0] = -r * u[0]
            t.tend(0, 1, s)[0] = 0.0
            t.advance(s, u, 0)
            t.advance(s, v, 1)
        if n % dstep == 0:
            tt = n * dt
            err.append(u[0] - (np.cos(omgg * tt) - r / (2.0 * omgg) * np.sin(omgg * tt)) * np.exp(-0.5 * r * tt))
    return np.array(err)


@pytest.mark.parametrize("name,order", [("ERK_1s1o", 0.98), ("ERK_4s4o", 3.98), ("ERK_SSP_2s2o", 1.98),
                                        ("ERK_SSP_3s3o", 2.98), ("ERK_SSP_4s3o", 2.98), ("ERK_SSP_5s3o_2N2*", 2.98),
                                        ("ERK_SSP_10s4o_2N", 3.98)])
def test_erk_convergence_order(name, order):
    e1, e2 = _run_ex(name, DT1, 1), _run_ex(name, DT2, 8)
    rate = np.log(np.linalg.norm(e1) / np.linalg.norm(e2)) / np.log(DT1 / DT2)
    assert rate >= order, (name, rate)


@pytest.mark.parametrize("name,order", [("IMEX_ARK232", 1.98), ("IMEX_ARK324", 2.98)])
def test_imex_convergence_order(name, order):
    e1, e2 = _run_imex(name, DT1, 1), _run_imex(name, DT2, 8)
    rate = np.log(np.linalg.norm(e1) / np.linalg.norm(e2)) / np.log(DT1 / DT2)
    assert rate >= order, (name, rate)


def test_shu_osher_to_butcher():
    t = rk_tables("ERK_SSP_3s3o")
    A = np.zeros((3, 3)); A[1, 0] = 1.0; A[2, :2] = 0.25
    assert np.sqrt(np.sum((A - t["a_ex"]) ** 2)) <= 1e-14 * 9
    assert np.sqrt(np.sum((np.array([1, 1, 4]) / 6.0 - t["b_ex"]) ** 2)) <= 1e-14 * 3
    t = rk_tables("ERK_SSP_4s3o")
    A = np.zeros((4, 4)); A[1, 0] = 0.5; A[2, :2] = 0.5; A[3, :3] = 1.0 / 6.0
    assert np.sqrt(np.sum((A - t["a_ex"]) ** 2)) <= 1e-14 * 16
    assert np.sqrt(np.sum((np.array([1, 1, 1, 3]) / 6.0 - t["b_ex"]) ** 2)) <= 1e-14 * 4


@pytest.mark.parametrize("name", ["IMEX_ARK232", "IMEX_ARK324"])
def test_imex_tableau_consistency(name):
    """Row sums of the explicit and implicit tableaux agree (c_ex == c_im), weights sum to one, stiffly accurate."""
    t = rk_tables(name)
    assert np.allclose(t["a_ex"].sum(1), t["a_im"].sum(1), atol=1e-14)
    assert np.isclose(t["b_ex"].sum(), 1.0, atol=1e-14) and np.isclose(t["b_im"].sum(), 1.0, atol=1e-14)
    assert np.allclose(t["a_im"][-1], t["b_im"], atol=1e-15)
