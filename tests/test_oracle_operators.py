"""Oracle pins, element operators: the reference's own known-answer test restated
(FElib/test/FE/element_operation_hexahedral/test_element_operation_hexahedral.f90:72-144, 236-259):
Dx, Dy, Dz of 4x^p+3y^p+2z^p against the analytic derivative, Lift against matmul(elem%Lift, f), Div with
Escale=(1,2,0.2), Gsqrt=100-x^2, for p = 1..7; tolerance sum((val-ans)^2) <= 1e-15 as in the reference.
Plus the cross-check between the two independent restatements (C++ oracle, NumPy set-up code)."""
import numpy as np
import pytest

from fe_project_b200.element import HexElement
from oracle_api import Oracle

EPS_REF = 1.0e-15  # test_element_operation_hexahedral.f90:243


def _oracle(p, lumped=False):
    return Oracle(p, 1, 1, 1, (-1, 1, -1, 1, -1, 1), lumped=lumped)


def _gen(e, p, fac):
    return (4.0 * e.x1 ** p + 3.0 * e.x2 ** p + 2.0 * e.x3 ** p) * fac


@pytest.mark.parametrize("p", range(1, 8))
def test_reference_known_answers(p):
    e = HexElement(p)
    o = _oracle(p)
    dat = _gen(e, p, 1.0)
    ans = [4.0 * e.x1 ** (p - 1) * p, 3.0 * e.x2 ** (p - 1) * p, 2.0 * e.x3 ** (p - 1) * p]
    for name, a in zip(("Dx", "Dy", "Dz"), ans):
        out = o.elem_op(name, dat)
        assert np.sum((out - a) ** 2) <= EPS_REF, name
    dat_f = np.concatenate([dat[e.Fmask[f]] for f in range(6)])
    lift_ans = o.lift_dense() @ dat_f
    lift = o.elem_op("Lift", dat_f)
    assert np.sum((lift - lift_ans) ** 2) <= EPS_REF
    Escale = (1.0, 2.0, 0.2)
    Gsqrt = 100.0 - e.x1 ** 2
    vec3 = np.concatenate([_gen(e, p, 1.0), _gen(e, p, 2.0), _gen(e, p, 3.0)])
    d4 = o.elem_op("Div", vec3, dat_f, nout=4 * e.Np).reshape(4, e.Np)
    div = (Escale[0] * d4[0] + Escale[1] * d4[1] + Escale[2] * d4[2] + d4[3]) / Gsqrt
    div_ans = (Escale[0] * ans[0] * 1.0 + Escale[1] * ans[1] * 2.0 + Escale[2] * ans[2] * 3.0 + lift_ans) / Gsqrt
    assert np.sum((div - div_ans) ** 2) <= EPS_REF


@pytest.mark.parametrize("p", [1, 2, 3, 5, 7])
@pytest.mark.parametrize("lumped", [False, True])
def test_two_restatements_agree(p, lumped):
    """C++ oracle vs NumPy set-up code (LGL nodes by Newton vs Golub-Welsch; Gauss-Jordan vs LAPACK)."""
    e = HexElement(p, lumped)
    o = _oracle(p, lumped)
    n = p + 1
    assert np.allclose(o.arr("x1d"), e.x1d, rtol=0, atol=2e-15)
    assert np.allclose(o.arr("w1d"), e.w1d, rtol=1e-14, atol=0)
    assert np.allclose(o.arr("D1D").reshape(n, n), e.D1D, rtol=0, atol=5e-13)
    assert np.allclose(o.arr("lift1d").reshape(n, 2), e.lift1d, rtol=0, atol=1e-12 * np.abs(e.lift1d).max())
    assert np.allclose(o.arr("VPOrdM1").reshape(n, n), e.VPOrdM1, rtol=0, atol=1e-13)
    assert np.allclose(o.arr("IntWeight"), e.IntWeight_lgl, rtol=1e-13)
    assert np.array_equal(o.iarr("Fmask").reshape(6, -1), e.Fmask)
    # the tensor-product lift equals the literal dense construction invM * Emat of hexahedral.F90:331-400
    dense = e.dense_reference_matrices()
    scale = np.abs(dense["Lift"]).max()
    assert np.abs(o.lift_dense() - dense["Lift"]).max() <= 1e-11 * scale
    assert np.abs(e.lift_dense() - dense["Lift"]).max() <= 1e-11 * scale
    for d, key in enumerate(("Dx1", "Dx2", "Dx3")):
        assert np.abs(o.dmat_dense(d) - dense[key]).max() <= 1e-12


def test_modal_filter_matrix_and_apply():
    """get_exp_filter (scale_element_modalfilter.F90:204-236): modes below etac untouched, highest mode damped by
    exp(-alpha); the 3-pass application equals the dense Kronecker product."""
    p = 7
    e = HexElement(p)
    o = _oracle(p)
    o.setup_dyn("NONHYDRO3D_HEVE", "ERK_SSP_3s3o", 1.0, True, (2.0 / 3.0, 1.0, 16, 0.0, 2.0, 8))
    Fh, Fv = o.arr("filt_h").reshape(8, 8), o.arr("filt_v").reshape(8, 8)
    assert np.abs(Fh - e.filter1d(2.0 / 3.0, 1.0, 16)).max() < 1e-13
    assert np.abs(Fv - e.filter1d(0.0, 2.0, 8)).max() < 1e-13
    V, invV = e.line.V, e.line.invV
    modal = invV @ Fh @ V
    assert np.allclose(np.diag(modal)[:5], 1.0, atol=1e-13)          # eta <= 2/3 -> p <= 4
    assert np.isclose(modal[7, 7], np.exp(-1.0), atol=1e-13)
    rng = np.random.default_rng(3)
    q = rng.standard_normal(e.Np)
    out = o.elem_op("ModalFilter", q)
    ref = np.einsum("kc,jb,ia,cba->kji", Fv, Fh, Fh, q.reshape(8, 8, 8)).reshape(-1)
    assert np.abs(out - ref).max() < 1e-13


def test_vfilter_pm1_removes_top_mode():
    p = 7
    e = HexElement(p)
    o = _oracle(p)
    from fe_project_b200.element import legendre_poly
    P = legendre_poly(p, e.x1d)
    top = np.tile(P[:, p][:, None, None], (1, 8, 8)).reshape(-1)      # P_7(z)
    low = np.tile(P[:, 3][:, None, None], (1, 8, 8)).reshape(-1)
    assert np.abs(o.elem_op("VFilterPM1", top)).max() < 1e-13
    assert np.abs(o.elem_op("VFilterPM1", low) - low).max() < 1e-13
