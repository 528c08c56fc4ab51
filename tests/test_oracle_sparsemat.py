"""Oracle pin, sparse matrix (a1): FElib/test/common/sparsemat/test_sparsemat.f90:28-93, matrix at :122-126."""
import numpy as np
import pytest

from oracle_api import sparsemat_matmul, Oracle

A = np.array([[1, 3, 0, 0, 0], [1, 2, 5, 0, 0], [4, 1, 3, 0, 0], [0, 3, 7, 4, 0], [1, 0, 0, 0, 5]], dtype=np.float64)
EPS = 1.0e-16


@pytest.mark.parametrize("ell", [False, True])
def test_reference_5x5(ell):
    x = np.ones(5)
    b, g = sparsemat_matmul(A, x, EPS, ell)
    assert np.abs(g - A).max() <= EPS          # GetVal(i,j) == A(i,j)
    assert np.abs(b - A @ x).max() <= EPS
    assert np.array_equal(b, [4, 8, 8, 14, 6])


@pytest.mark.parametrize("ell", [False, True])
def test_element_matrices_spmv(ell):
    """Dx/Dy/Dz/Lift stored sparse with the reference's drop tolerance 500*EPS (scale_sparsemat.F90:127-131)
    reproduce the tensor-product operators (General vs TensorProd3D agreement of the reference test)."""
    p = 4
    o = Oracle(p, 1, 1, 1, (-1, 1, -1, 1, -1, 1))
    rng = np.random.default_rng(0)
    q = rng.standard_normal(o.Np)
    for d, name in enumerate(("Dx", "Dy", "Dz")):
        c, _ = sparsemat_matmul(o.dmat_dense(d), q, 500 * 2.220446e-16, ell)
        assert np.abs(c - o.elem_op(name, q)).max() < 1e-12
    f = rng.standard_normal(o.NfpTot)
    c, _ = sparsemat_matmul(o.lift_dense(), f, 500 * 2.220446e-16, ell)
    assert np.abs(c - o.elem_op("Lift", f)).max() < 1e-11
