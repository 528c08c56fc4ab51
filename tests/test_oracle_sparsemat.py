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


@pytest.mark.parametrize("ell", [False, True])
def test_matmul1_2_and_matmul2(ell):
    """sparsemat_matmul1_2 (c = A (b1 .* b2), scale_sparsemat.F90:386-408) and sparsemat_matmul2 (b(NQ,N) -> c(NQ,M), :411-431) in both
    storage formats against dense products; on the reference's 5x5 matrix with b1 = b2 = ones the first is the test's (4, 8, 8, 14, 6)."""
    from oracle_api import sparsemat_matmul_ex
    c, _ = sparsemat_matmul_ex(A, np.ones(5), EPS, ell, mode=1, b2=np.ones(5))
    assert np.array_equal(c, [4, 8, 8, 14, 6])
    rng = np.random.default_rng(5)
    b1, b2 = rng.standard_normal(5), rng.standard_normal(5)
    c, _ = sparsemat_matmul_ex(A, b1, EPS, ell, mode=1, b2=b2)
    assert np.abs(c - A @ (b1 * b2)).max() <= 1e-14
    B = rng.standard_normal((5, 3))                       # (N, NQ) in C order = b(NQ,N) in Fortran order
    c, _ = sparsemat_matmul_ex(A, B, EPS, ell, mode=2, NQ=3)
    assert np.abs(c - A @ B).max() <= 1e-14


def test_python_csr_arrays_are_the_oracles():
    """The host mirror's CSR arrays (1-based val / colIdx / rowPtr, sparsemat_Init :137-195) equal the oracle's, on the reference's 5x5
    matrix and on an element matrix with the reference's drop tolerance."""
    from oracle_api import sparsemat_matmul_ex
    from fe_project_b200.advect3d import SparseMat
    o = Oracle(3, 1, 1, 1, (-1, 1, -1, 1, -1, 1))
    for mat, eps in ((A, EPS), (o.dmat_dense(0), 500 * 2.220446e-16), (o.lift_dense(), 500 * 2.220446e-16)):
        s = SparseMat(mat, eps=eps, storage_format="CSR")
        _, (val, col, rp) = sparsemat_matmul_ex(mat, np.ones(mat.shape[1]), eps, False)
        assert np.array_equal(s.val, val) and np.array_equal(s.colIdx, col) and np.array_equal(s.rowPtr, rp)
        assert s.rowPtr[0] == 1 and s.rowPtr[-1] == s.nnz + 1
