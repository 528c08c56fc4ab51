"""Host-side mirror of sample/advect3d (BASELINE config 1) over the C ABI.

`SparseMat` keeps the storage of the reference type `sparsemat` (FElib/src/common/scale_sparsemat.F90:33-55,
`sparsemat_Init` :100-250): entries with |a| <= EPS are dropped (EPS = CONST_EPS*500, :130), ELL storage is
slot-major `l = i + (k-1)*M` (:172) with 1-based `colIdx`; padding slots hold value 0 and the row's own column.
`Advect3D` mirrors the program's stage loop (sample/advect3d/test_advect3d.f90:81-126) and its call
`advect3d_kernel_cal_tend(dqdt, q, u, v, w, Dx, Dy, Dz, Lift, lmesh, elem)` (mod_advect3d_kernel.f90:34-44).
All arithmetic runs in libfedg.so on the GPU; this module only builds and marshals arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .dyncore import AtmDynDGMDriver_nonhydro3d, _f64, _ptr
from .element import HexElement
from .initcond import SCALE_CONST
from .mesh import LocalMeshCube


class SparseMat:
    """`sparsemat` in ELL (default, what the kernels consume) or CSR storage.  mat: dense (M, N)."""

    def __init__(self, mat: np.ndarray, eps: float | None = None, storage_format: str = "ELL"):
        mat = np.asarray(mat, dtype=np.float64)
        self.M, self.N = mat.shape
        eps = SCALE_CONST["EPS"] * 500.0 if eps is None else eps
        keep = np.abs(mat) > eps
        self.storage_format = storage_format.upper()
        assert self.storage_format in ("ELL", "CSR")
        if self.storage_format == "CSR":            # sparsemat_Init, CSR branch (scale_sparsemat.F90:137-150, 176-195): 1-based arrays
            rows, cols = np.nonzero(keep)
            self.nnz = int(rows.size)
            self.val = np.ascontiguousarray(mat[rows, cols])
            self.colIdx = (cols + 1).astype(np.int32)
            self.rowPtr = (np.concatenate([[0], np.cumsum(keep.sum(axis=1))]) + 1).astype(np.int32)
            self.col_size = 0
            return
        self.col_size = int(keep.sum(axis=1).max())
        self.nnz = int(keep.sum())
        self.val = np.zeros(self.M * self.col_size)
        self.colIdx = np.zeros(self.M * self.col_size, dtype=np.int32)
        for i in range(self.M):
            cols = np.nonzero(keep[i])[0]
            for k in range(self.col_size):
                l = i + k * self.M
                if k < cols.size:
                    self.val[l] = mat[i, cols[k]]
                    self.colIdx[l] = cols[k] + 1
                else:
                    self.colIdx[l] = (i if i < self.N else 0) + 1

    def abi(self) -> "_lib.SparseMatDesc":
        d = _lib.SparseMatDesc()
        d.M, d.N, d.col_size = self.M, self.N, self.col_size
        d.val = self.val.ctypes.data_as(C.c_void_p)
        d.colIdx = self.colIdx.ctypes.data_as(C.c_void_p)
        return d

    def abi_any(self) -> "_lib.SparseMatAnyDesc":
        d = _lib.SparseMatAnyDesc()
        d.storage_format_id = 1 if self.storage_format == "CSR" else 2     # SPARSEMAT_STORAGE_TYPEID_* (scale_sparsemat.F90:68-69)
        d.M, d.N, d.nnz, d.col_size = self.M, self.N, self.nnz, self.col_size
        d.val = self.val.ctypes.data_as(C.c_void_p)
        d.colIdx = self.colIdx.ctypes.data_as(C.c_void_p)
        if self.storage_format == "CSR":
            d.rowPtrSize = self.M + 1
            d.rowPtr = self.rowPtr.ctypes.data_as(C.c_void_p)
        return d

    def matmul1(self, b: np.ndarray) -> np.ndarray:
        """sparsemat_matmul1, either storage: b (N,) -> (M,)."""
        b = _f64(b); c = np.zeros(self.M); d = self.abi_any()
        _lib.check(_lib.load().fedg_sparsemat_matmul1(C.byref(d), _ptr(b), _ptr(c)))
        return c

    def matmul1_2(self, b1: np.ndarray, b2: np.ndarray) -> np.ndarray:
        """sparsemat_matmul1_2: A (b1 .* b2)."""
        b1, b2 = _f64(b1), _f64(b2); c = np.zeros(self.M); d = self.abi_any()
        _lib.check(_lib.load().fedg_sparsemat_matmul1_2(C.byref(d), _ptr(b1), _ptr(b2), _ptr(c)))
        return c

    def matmul2(self, b: np.ndarray) -> np.ndarray:
        """sparsemat_matmul2: b given as (N, NQ) in C order (= the reference's b(NQ,N)) -> (M, NQ)."""
        b = _f64(b); assert b.ndim == 2 and b.shape[0] == self.N
        c = np.zeros((self.M, b.shape[1])); d = self.abi_any()
        _lib.check(_lib.load().fedg_sparsemat_matmul2(C.byref(d), _ptr(b), _ptr(c), int(b.shape[1])))
        return c

    def matmul(self, b: np.ndarray) -> np.ndarray:
        """sparsemat_matmul on the GPU: b (nvec, N) or (N,) -> (nvec, M) or (M,)."""
        assert self.storage_format == "ELL"
        b2 = _f64(np.atleast_2d(b))
        assert b2.shape[1] == self.N
        c = np.zeros((b2.shape[0], self.M))
        d = self.abi()
        _lib.check(_lib.load().fedg_sparsemat_matmul(C.byref(d), _ptr(b2), _ptr(c), int(b2.shape[0])))
        return c if np.ndim(b) == 2 else c[0]


def element_sparsemats(elem: HexElement):
    """Dx, Dy, Dz, Lift of the program's init(): `Dx%Init(refElem%Dx1, storage_format='ELL')` etc."""
    n = elem.np1
    I, D = np.eye(n), elem.D1D
    Dx1 = np.einsum("kc,jb,ia->kjicba", I, I, D).reshape(elem.Np, elem.Np)
    Dx2 = np.einsum("kc,jb,ia->kjicba", I, D, I).reshape(elem.Np, elem.Np)
    Dx3 = np.einsum("kc,jb,ia->kjicba", D, I, I).reshape(elem.Np, elem.Np)
    return SparseMat(Dx1), SparseMat(Dx2), SparseMat(Dx3), SparseMat(elem.lift_dense())


class Advect3D:
    """sample/advect3d on one GPU: q, u, v, w are (NeA, Np) arrays (node index fastest)."""

    def __init__(self, elem: HexElement, mesh: LocalMeshCube, TINTEG_SCHEME_TYPE="ERK_4s4o", TIME_DT=0.008):
        self.elem, self.mesh = elem, mesh
        self._drv = AtmDynDGMDriver_nonhydro3d(elem, mesh, SCALE_CONST)    # mesh registration (fedg_create)
        self.L, self.h = self._drv.L, self._drv.h
        self.mats = element_sparsemats(elem)
        descs = [m.abi() for m in self.mats]
        _lib.check(self.L.fedg_advect3d_init(self.h, TINTEG_SCHEME_TYPE.encode(), float(TIME_DT), *[C.byref(d) for d in descs]))
        self.n_field, self.n_int = self._drv.n_field, self._drv.n_int

    def set(self, q, u, v, w):
        a = [self._drv._chk_field(x) for x in (q, u, v, w)]
        _lib.check(self.L.fedg_advect3d_set(self.h, *[_ptr(x) for x in a]))

    def get(self) -> np.ndarray:
        q = np.zeros(self.n_field)
        _lib.check(self.L.fedg_advect3d_get(self.h, _ptr(q)))
        return q

    def cal_tend(self) -> np.ndarray:
        out = np.zeros(self.n_int)
        _lib.check(self.L.fedg_advect3d_cal_tend(self.h, _ptr(out)))
        return out

    def update(self, nsteps=1):
        _lib.check(self.L.fedg_advect3d_update(self.h, int(nsteps)))

    def last_timing(self):
        return self._drv.last_timing()


def gaussian_hill(mesh: LocalMeshCube, xc=0.25, yc=0.25, zc=0.5, width=0.05, intrp_order=7):
    """Initial q of the shipped test.conf (`InitShapeName='gaussian-hill'`, `InitGPMatPolyOrder=7`): the profile of
    sample/auxiliary/mod_fieldutil.f90:409-411 sampled on an order-7 element and projected by modal truncation
    (set_initcond, test_advect3d.f90:176-225)."""
    e = mesh.elem
    T1, src = e.l2proj_from(intrp_order)
    xq = src.x
    vx = (mesh.xmax - mesh.xmin) * np.arange(mesh.NeX + 1) / mesh.NeX + mesh.xmin
    vy = (mesh.ymax - mesh.ymin) * np.arange(mesh.NeY + 1) / mesh.NeY + mesh.ymin
    vz = mesh.FZ
    q = np.zeros((mesh.NeA, e.Np))
    for ke in range(mesh.Ne):
        ex, ey, ez = mesh.ex[ke], mesh.ey[ke], mesh.ez[ke]
        x = vx[ex] + 0.5 * (xq + 1.0) * (vx[ex + 1] - vx[ex])
        y = vy[ey] + 0.5 * (xq + 1.0) * (vy[ey + 1] - vy[ey])
        z = vz[ez] + 0.5 * (xq + 1.0) * (vz[ez + 1] - vz[ez])
        dist = ((x[None, None, :] - xc) ** 2 + (y[None, :, None] - yc) ** 2 + (z[:, None, None] - zc) ** 2) / width ** 2
        q[ke] = np.einsum("kc,jb,ia,cba->kji", T1, T1, T1, np.exp(-0.5 * dist)).reshape(-1)
    return q
