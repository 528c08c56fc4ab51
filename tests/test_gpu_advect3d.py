"""BASELINE config 1 (sample/advect3d; SURVEY.md rows a1 + a18) on the GPU through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

from cases import rel_l2
from fe_project_b200.advect3d import Advect3D, SparseMat, element_sparsemats, gaussian_hill
from fe_project_b200.element import HexElement
from fe_project_b200.mesh import LocalMeshCube
from oracle_api import Oracle, OracleAdvect3D, sparsemat_matmul

pytestmark = pytest.mark.gpu
TOL = 1.0e-10


def test_sparsemat_matmul_reference_5x5():
    """FElib/test/common/sparsemat/test_sparsemat.f90:122-129: A * ones = (4, 8, 8, 14, 6)."""
    A = np.array([[1, 3, 0, 0, 0], [1, 2, 5, 0, 0], [4, 1, 3, 0, 0], [0, 3, 7, 4, 0], [1, 0, 0, 0, 5]], dtype=np.float64)
    s = SparseMat(A, eps=1e-16)
    assert np.array_equal(s.matmul(np.ones(5)), [4, 8, 8, 14, 6])
    rng = np.random.default_rng(3)
    b = rng.standard_normal((7, 5))
    assert np.abs(s.matmul(b) - b @ A.T).max() <= 1e-14


@pytest.mark.parametrize("fmt", ["CSR", "ELL"])
def test_sparsemat_csr_ell_matmul1_matmul1_2_matmul2(fmt):
    """The three generic products of the reference type in both storage formats (scale_sparsemat.F90:355-431) on the device against the
    oracle: the reference's 5x5 known answer (test_sparsemat.f90:69-93) and the p = 3 element matrices (Dx: 64 x 64, Lift: 64 x 96)."""
    from oracle_api import sparsemat_matmul_ex
    A = np.array([[1, 3, 0, 0, 0], [1, 2, 5, 0, 0], [4, 1, 3, 0, 0], [0, 3, 7, 4, 0], [1, 0, 0, 0, 5]], dtype=np.float64)
    s = SparseMat(A, eps=1e-16, storage_format=fmt)
    assert np.array_equal(s.matmul1(np.ones(5)), [4, 8, 8, 14, 6])
    assert np.array_equal(s.matmul1_2(np.ones(5), np.ones(5)), [4, 8, 8, 14, 6])
    o = Oracle(3, 1, 1, 1, (-1, 1, -1, 1, -1, 1))
    rng = np.random.default_rng(11)
    eps = 500 * 2.220446e-16
    for dense in (A, o.dmat_dense(0), o.dmat_dense(2), o.lift_dense()):
        M, N = dense.shape
        s = SparseMat(dense, eps=1e-16 if dense is A else eps, storage_format=fmt)
        e = 1e-16 if dense is A else eps
        b1, b2, B = rng.standard_normal(N), rng.standard_normal(N), rng.standard_normal((N, 6))
        ell = fmt == "ELL"
        scale = np.abs(dense).sum(axis=1).max()
        assert np.abs(s.matmul1(b1) - sparsemat_matmul_ex(dense, b1, e, ell)[0]).max() <= 1e-14 * scale
        assert np.abs(s.matmul1_2(b1, b2) - sparsemat_matmul_ex(dense, b1, e, ell, mode=1, b2=b2)[0]).max() <= 1e-14 * scale
        assert np.abs(s.matmul2(B) - sparsemat_matmul_ex(dense, B, e, ell, mode=2, NQ=6)[0]).max() <= 1e-14 * scale


@pytest.mark.parametrize("p", [3, 7])
def test_sparsemat_matmul_element_matrices(p):
    """Dx/Dy/Dz/Lift as ELL sparsemat objects: GPU product == oracle product (same slot order), many right-hand sides."""
    e = HexElement(p)
    o = Oracle(p, 1, 1, 1, (-1, 1, -1, 1, -1, 1))
    rng = np.random.default_rng(p)
    for k, s in enumerate(element_sparsemats(e)):
        b = rng.standard_normal((5, s.N))
        got = s.matmul(b)
        dense = o.dmat_dense(k) if k < 3 else o.lift_dense()
        for j in range(5):
            ref, _ = sparsemat_matmul(dense, b[j], 500 * 2.220446e-16, True)
            assert np.abs(got[j] - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


def _case(p, ne, dt, scheme="ERK_4s4o", vel=(0.5, 0.5, 0.5), periodic=(True, True, True), width=0.05):
    e = HexElement(p)
    mesh = LocalMeshCube(e, *ne, 0, 1, 0, 1, 0, 1, periodic=periodic)
    o = Oracle(p, *ne, (0, 1, 0, 1, 0, 1), periodic=periodic)
    a = OracleAdvect3D(o, scheme, dt)
    q = gaussian_hill(mesh, width=width)
    x, y, z = (mesh.pos_en[d] for d in range(3))
    u = np.zeros_like(q); v = np.zeros_like(q); w = np.zeros_like(q)
    if vel == "swirl":   # non-constant, divergence-free-ish flow so that alpha and the jumps vary from node to node
        u[:mesh.Ne] = 0.5 + 0.3 * np.sin(2 * np.pi * y) * np.cos(2 * np.pi * z)
        v[:mesh.Ne] = -0.4 + 0.3 * np.sin(2 * np.pi * x)
        w[:mesh.Ne] = 0.2 * np.cos(2 * np.pi * x) * np.sin(2 * np.pi * y)
    else:
        u[:mesh.Ne], v[:mesh.Ne], w[:mesh.Ne] = vel
    for nm, f in zip("quvw", (q, u, v, w)):
        a.arr(nm)[:] = f.reshape(-1)
    g = Advect3D(e, mesh, scheme, dt)
    g.set(q, u, v, w)
    return mesh, a, g


@pytest.mark.parametrize("p,ne,vel", [(3, (8, 8, 8), (0.5, 0.5, 0.5)), (3, (5, 3, 4), "swirl"), (7, (2, 3, 2), "swirl")])
def test_cal_tend(p, ne, vel):
    mesh, a, g = _case(p, ne, 0.008, vel=vel)
    ref = a.cal_tend()
    got = g.cal_tend()
    assert rel_l2(got, ref) <= 1e-13


def test_shipped_config_100_and_1000_steps():
    """sample/advect3d/test.conf: 8x8x8, p = 3, ERK_4s4o, dt = 0.008, gaussian hill, u = v = w = 0.5; N = 100 and 1000."""
    mesh, a, g = _case(3, (8, 8, 8), 0.008)
    n = mesh.Ne * mesh.elem.Np
    a.update(100); g.update(100)
    assert rel_l2(g.get()[:n], a.arr("q")[:n]) <= TOL
    a.update(900); g.update(900)
    assert rel_l2(g.get()[:n], a.arr("q")[:n]) <= TOL
    w = np.tile(mesh.elem.IntWeight_lgl, mesh.Ne) * mesh.J.reshape(-1)
    assert abs(np.sum(w * g.get()[:n]) - np.sum(w * gaussian_hill(mesh)[:mesh.Ne].reshape(-1))) <= 1e-13


@pytest.mark.parametrize("scheme,p,ne", [("ERK_SSP_4s3o", 3, (4, 6, 5)), ("ERK_SSP_3s3o", 7, (2, 2, 3)), ("ERK_1s1o", 3, (3, 3, 3)),
                                         ("ERK_SSP_10s4o_2N", 3, (4, 4, 4))])
def test_steps_other_schemes_and_orders(scheme, p, ne):
    mesh, a, g = _case(p, ne, 0.002, scheme=scheme, vel="swirl", width=0.15)
    n = mesh.Ne * mesh.elem.Np
    a.update(20); g.update(20)
    assert rel_l2(g.get()[:n], a.arr("q")[:n]) <= TOL


def test_errors():
    from fe_project_b200 import _lib
    e = HexElement(3)
    mesh = LocalMeshCube(e, 2, 2, 2, 0, 1, 0, 1, 0, 1, periodic=(True, True, True))
    with pytest.raises(_lib.FedgError):
        Advect3D(e, mesh, "IMEX_ARK232", 0.01)
    with pytest.raises(_lib.FedgError):
        Advect3D(e, mesh, "ERK_4s4o", -1.0)
