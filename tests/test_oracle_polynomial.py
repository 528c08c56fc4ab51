"""Known-answer pin from the reference: FElib/test/FE/polynomial/test_polynomial.f90:10-58, 75-121 holds the closed-form
Gauss-Lobatto / Gauss-Legendre points and weights for orders 1..4 and checks Polynomial_GenGaussLobattoPt /
_GenGaussLobattoPtIntWeight / _GenGaussLegendrePt / _GenGaussLegendrePtIntWeight against them to 5e-15.  Both restatements of
the set-up code (NumPy: fe_project_b200/element.py, C++: oracle/element.cpp) are held to the same vectors and the same
threshold; the higher orders the path uses (p = 7, and 11 for the initial-state projection) are checked by exactness of the
quadrature and by the two restatements agreeing."""
import numpy as np
import pytest

from fe_project_b200.element import (gauss_legendre_pts, gauss_legendre_weights, gauss_lobatto_pts, gauss_lobatto_weights,
                                     legendre_poly)
from oracle_api import Oracle

CHECK_EPS = 5.0e-15
s = np.sqrt
GOLD = {
    1: dict(lgl=[-1.0, 1.0], lglw=[1.0, 1.0], gl=[0.0], glw=[2.0]),
    2: dict(lgl=[-1.0, 0.0, 1.0], lglw=[1 / 3, 4 / 3, 1 / 3], gl=[-s(1 / 3), s(1 / 3)], glw=[1.0, 1.0]),
    3: dict(lgl=[-1.0, -s(1 / 5), s(1 / 5), 1.0], lglw=[1 / 6, 5 / 6, 5 / 6, 1 / 6], gl=[-s(3 / 5), 0.0, s(3 / 5)], glw=[5 / 9, 8 / 9, 5 / 9]),
    4: dict(lgl=[-1.0, -s(21.0) / 7, 0.0, s(21.0) / 7, 1.0], lglw=[1 / 10, 49 / 90, 32 / 45, 49 / 90, 1 / 10],
            gl=[-s(3 + 2 * s(6 / 5)) / s(7.0), -s(3 - 2 * s(6 / 5)) / s(7.0), s(3 - 2 * s(6 / 5)) / s(7.0), s(3 + 2 * s(6 / 5)) / s(7.0)],
            glw=[(18 - s(30.0)) / 36, (18 + s(30.0)) / 36, (18 + s(30.0)) / 36, (18 - s(30.0)) / 36]),
}


@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_reference_points_and_weights_numpy(n):
    g = GOLD[n]
    assert np.abs(gauss_lobatto_pts(n) - g["lgl"]).max() <= CHECK_EPS
    assert np.abs(gauss_lobatto_weights(n) - g["lglw"]).max() <= CHECK_EPS
    assert np.abs(gauss_legendre_pts(n) - g["gl"]).max() <= CHECK_EPS
    assert np.abs(gauss_legendre_weights(n) - g["glw"]).max() <= CHECK_EPS


@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_reference_points_and_weights_oracle(n):
    o = Oracle(n, 1, 1, 1, (0, 1, 0, 1, 0, 1))
    g = GOLD[n]
    assert np.abs(o.arr("x1d") - g["lgl"]).max() <= CHECK_EPS
    assert np.abs(o.arr("w1d") - g["lglw"]).max() <= CHECK_EPS


@pytest.mark.parametrize("n", [5, 6, 7, 11])
def test_higher_orders_quadrature_exactness_and_agreement(n):
    x, w = gauss_lobatto_pts(n), gauss_lobatto_weights(n)
    # LGL with n + 1 points integrates polynomials up to degree 2n - 1 exactly; Legendre P_k, k >= 1, integrate to zero
    P = legendre_poly(n, x)                              # (n + 1 points, n + 1 orders)
    assert abs(w.sum() - 2.0) <= 1e-14
    for k in range(1, n):
        assert abs(np.sum(w * P[:, k])) <= 1e-14, k
        assert abs(np.sum(w * P[:, k] * P[:, k]) - 2.0 / (2 * k + 1)) <= 1e-14, k
    if n <= 7:
        o = Oracle(n, 1, 1, 1, (0, 1, 0, 1, 0, 1))
        assert np.abs(o.arr("x1d") - x).max() <= CHECK_EPS
        assert np.abs(o.arr("w1d") - w).max() <= CHECK_EPS
