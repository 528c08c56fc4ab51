"""Host-side reference element (hexahedral, LGL tensor-product) for the DG dynamics path.

Mirrors what the reference builds once at start-up and hands to the dynamics
kernels: `HexahedralElement%Init`
(FElib/src/element/scale_element_hexahedral.F90:52-404), the 1D `LineElement`
(FElib/src/element/scale_element_line.F90), the tensor-product operator tables of
`setup_elem_operator` (FElib/src/element/scale_element_operation_tensorprod3D.F90.erb:508-565)
and the exponential modal filter of `get_exp_filter`
(FElib/src/element/scale_element_modalfilter.F90:204-236).

This is set-up code (runs once, on the host); the arrays it produces are what a
Fortran caller would pass through `fedg_create()` (include/fedg.h).  Indices are
0-based here; everything handed to the C ABI that is an index map is converted
to the reference's 1-based convention in `mesh.py`.
"""
from __future__ import annotations

import dataclasses
import numpy as np
from scipy.linalg import eigh_tridiagonal


def legendre_poly(nord: int, x: np.ndarray) -> np.ndarray:
    """P[i, n] = P_n(x_i), n = 0..nord (scale_polynomial.F90 `Polynomial_GenLegendrePoly_sub`)."""
    x = np.asarray(x, dtype=np.float64)
    P = np.zeros((x.size, nord + 1))
    P[:, 0] = 1.0
    if nord == 0:
        return P
    P[:, 1] = x
    for n in range(2, nord + 1):
        P[:, n] = ((2 * n - 1) * x * P[:, n - 1] - (n - 1) * P[:, n - 2]) / n
    return P


def dlegendre_poly(nord: int, x: np.ndarray, P: np.ndarray) -> np.ndarray:
    """dP_n/dx at x (scale_polynomial.F90 `Polynomial_GenDLegendrePoly`)."""
    x = np.asarray(x, dtype=np.float64)
    G = np.zeros((x.size, nord + 1))
    if nord == 0:
        return G
    G[:, 1] = 1.0
    for n in range(2, nord + 1):
        G[:, n] = 2.0 * x * G[:, n - 1] - G[:, n - 2] + P[:, n - 1]
    return G


def _jacobi_gauss_pts(alpha: int, beta: int, N: int) -> np.ndarray:
    """Golub-Welsch nodes (scale_polynomial.F90 `gen_JacobiGaussQuadraturePts`, LAPACK dstev there)."""
    if N == 0:
        return np.array([-(alpha - beta) / (alpha + beta + 2.0)])
    i = np.arange(N + 1, dtype=np.float64)
    h1 = 2.0 * i + alpha + beta
    with np.errstate(invalid="ignore", divide="ignore"):      # alpha = beta = 0: 0 / 0 in the first entry, set below
        d = -(alpha ** 2 - beta ** 2) / (h1 * (h1 + 2.0))
    k = np.arange(1, N + 1, dtype=np.float64)
    e = 2.0 / (h1[:-1] + 2.0) * np.sqrt(
        k * (k + alpha + beta) * (k + alpha) * (k + beta) / ((h1[:-1] + 1.0) * (h1[:-1] + 3.0)))
    if alpha + beta < 1e-16:
        d[0] = 0.0
    return eigh_tridiagonal(d, e, eigvals_only=True)


def gauss_lobatto_pts(nord: int) -> np.ndarray:
    pts = np.empty(nord + 1)
    pts[0], pts[-1] = -1.0, 1.0
    if nord > 1:
        pts[1:-1] = _jacobi_gauss_pts(1, 1, nord - 2)
    return pts


def gauss_lobatto_weights(nord: int) -> np.ndarray:
    x = gauss_lobatto_pts(nord)
    P = legendre_poly(nord, x)
    return 2.0 / (nord * (nord + 1) * P[:, nord] ** 2)


def gauss_legendre_pts(n: int) -> np.ndarray:
    return _jacobi_gauss_pts(0, 0, n - 1)


def gauss_legendre_weights(n: int) -> np.ndarray:
    x = gauss_legendre_pts(n)
    P = legendre_poly(n, x)
    dP = dlegendre_poly(n, x, P)
    return 2.0 / ((1.0 - x ** 2) * dP[:, n] ** 2)


def dlagrange_lgl(nord: int, x: np.ndarray) -> np.ndarray:
    """lr[k, n] = d l_k / dx at x_n (scale_polynomial.F90 `Polynomial_GenDLagrangePoly_lglpt`)."""
    P = legendre_poly(nord, x)
    N1 = nord + 1
    lr = np.zeros((N1, N1))
    for n in range(N1):
        s = 0.0
        for k in range(N1):
            if k == 0 and n == 0:
                lr[k, n] = -0.25 * nord * (nord + 1)
            elif k == nord and n == nord:
                lr[k, n] = 0.25 * nord * (nord + 1)
            elif k == n:
                lr[k, n] = 0.0
            else:
                lr[k, n] = P[n, nord] / (P[k, nord] * (x[n] - x[k]))
            if k != n:
                s += lr[k, n]
        lr[n, n] = -s
    return lr


def exp_filter_coefs(etac: float, alpha: float, order: int, nord: int) -> np.ndarray:
    """Modal damping factors (scale_element_modalfilter.F90 `get_exp_filter`, tend_flag=.false.)."""
    f = np.ones(nord + 1)
    for p in range(nord + 1):
        eta = p / nord
        if eta > etac and p != 0:
            f[p] = np.exp(-alpha * ((eta - etac) / (1.0 - etac)) ** order)
    return f


@dataclasses.dataclass
class LineElement:
    """1D LGL element (scale_element_line.F90 `construct_Element`)."""
    order: int
    lumped: bool = False

    def __post_init__(self):
        N = self.order
        self.Np = N + 1
        self.x = gauss_lobatto_pts(N)
        P = legendre_poly(N, self.x)
        self.V = P * np.sqrt(np.arange(N + 1) + 0.5)[None, :]
        self.invV = np.linalg.inv(self.V)
        self.Dx = dlagrange_lgl(N, self.x).T.copy()      # Dx[n, l] = d l_l/dx (x_n)
        self.w = gauss_lobatto_weights(N)
        if self.lumped:
            self.M = np.diag(self.w)
            self.invM = np.diag(1.0 / self.w)
        else:
            self.invM = self.V @ self.V.T
            self.M = np.linalg.inv(self.invM)

    def filter_mat(self, etac, alpha, order):
        return self.V @ np.diag(exp_filter_coefs(etac, alpha, order, self.order)) @ self.invV

    def trunc_mat_pm1(self):
        """Nodal matrix that removes the highest Legendre mode (IntrpMat_VPOrdM1, tensorprod3D.F90.erb:556-559)."""
        invV = self.invV.copy()
        invV[-1, :] = 0.0
        return self.V @ invV


class HexElement:
    """Tensor-product hexahedral element of order p (horizontal == vertical order).

    Attributes follow the reference names: Np, Nfp, NfpTot, Fmask (0-based node ids
    of the 6 faces, order y-, x+, y+, x-, z-, z+), D1D, lift1d, VPOrdM1, IntWeight_lgl.
    """

    def __init__(self, order: int, lumped: bool = False):
        self.order = order
        self.lumped = lumped
        self.np1 = n = order + 1
        self.Np = n ** 3
        self.Nfp = n * n
        self.Nfaces = 6
        self.NfpTot = 6 * self.Nfp
        self.line = LineElement(order, lumped)
        self.x1d = self.line.x
        self.w1d = self.line.w
        self.D1D = self.line.Dx                       # D1D[i, l]
        ids = np.arange(self.Np).reshape(n, n, n)     # ids[k, j, i]
        self.Fmask = np.stack([
            ids[:, 0, :].reshape(-1),       # y-: (i, k) -> i + k*n
            ids[:, :, n - 1].reshape(-1),   # x+: (j, k)
            ids[:, n - 1, :].reshape(-1),   # y+
            ids[:, :, 0].reshape(-1),       # x-
            ids[0, :, :].reshape(-1),       # z-: (i, j)
            ids[n - 1, :, :].reshape(-1),   # z+
        ])
        # lifting weights: Lift = invM * Emat collapses, for a tensor-product element, to
        # I (x) I (x) invM1D[:, end-node]  (hexahedral.F90:331-400 + tensorprod3D.F90.erb:537-551)
        if lumped:
            lw = np.zeros((n, 2))
            lw[0, 0] = 1.0 / self.w1d[0]
            lw[n - 1, 1] = 1.0 / self.w1d[n - 1]
        else:
            lw = np.stack([self.line.invM[:, 0], self.line.invM[:, n - 1]], axis=1)
        self.lift1d = lw                               # lift1d[m, side]; side 0 = minus face
        self.VPOrdM1 = self.line.trunc_mat_pm1()
        k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
        self.x1 = self.x1d[i].reshape(-1)
        self.x2 = self.x1d[j].reshape(-1)
        self.x3 = self.x1d[k].reshape(-1)
        self.IntWeight_lgl = (self.w1d[i] * self.w1d[j] * self.w1d[k]).reshape(-1)
        self.IndexH2Dto3D = (i + j * n).reshape(-1)

    # ---- dense forms used by conformance tests and by the ABI (`elem%Lift`, `Lift_mat`) ----
    def lift_mat(self) -> np.ndarray:
        """Lift_mat[f, k, j, i] (Fortran Lift_mat(i,j,k,f))."""
        n = self.np1
        L = np.zeros((6, n, n, n))
        lw = self.lift1d
        L[0] = lw[:, 0][None, :, None]
        L[2] = lw[:, 1][None, :, None]
        L[1] = lw[:, 1][None, None, :]
        L[3] = lw[:, 0][None, None, :]
        L[4] = lw[:, 0][:, None, None]
        L[5] = lw[:, 1][:, None, None]
        return L

    def lift_dense(self) -> np.ndarray:
        """elem%Lift as the reference stores it: (Np, NfpTot)."""
        n = self.np1
        Lm = self.lift_mat()
        out = np.zeros((self.Np, self.NfpTot))
        for f in range(6):
            for k in range(n):
                for j in range(n):
                    for i in range(n):
                        p = i + j * n + k * n * n
                        if f in (0, 2):
                            fp = i + k * n
                        elif f in (1, 3):
                            fp = j + k * n
                        else:
                            fp = i + j * n
                        out[p, f * self.Nfp + fp] = Lm[f, k, j, i]
        return out

    def filter1d(self, etac, alpha, order):
        return self.line.filter_mat(etac, alpha, order)

    def l2proj_from(self, order_in: int) -> np.ndarray:
        """1D factor of `Generate_L2ProjMat` (scale_element_base.F90 NodalTransferMat, pmax = self.order)."""
        src = LineElement(order_in, False)
        m = min(self.order, order_in) + 1
        invV_in = np.zeros((self.np1, src.Np))
        invV_in[:m, :] = src.invV[:m, :]
        return self.line.V @ invV_in, src

    # dense 3D construction following hexahedral.F90 literally (validation of the tensor shortcuts)
    def dense_reference_matrices(self):
        n = self.np1
        x = self.x1d
        P = legendre_poly(self.order, x)
        nrm = np.sqrt(np.arange(n) + 0.5)
        V1 = P * nrm[None, :]
        V = np.einsum("kc,jb,ia->kjicba", V1, V1, V1).reshape(self.Np, self.Np)
        invV = np.linalg.inv(V)
        D = self.D1D
        I = np.eye(n)
        Dx1 = np.einsum("kc,jb,ia->kjicba", I, I, D).reshape(self.Np, self.Np)
        Dx2 = np.einsum("kc,jb,ia->kjicba", I, D, I).reshape(self.Np, self.Np)
        Dx3 = np.einsum("kc,jb,ia->kjicba", D, I, I).reshape(self.Np, self.Np)
        if self.lumped:
            invM = np.diag(1.0 / self.IntWeight_lgl)
        else:
            invM = V @ V.T
        V2 = np.einsum("kc,ia->kica", V1, V1).reshape(self.Nfp, self.Nfp)
        if self.lumped:
            Medge = np.diag(np.outer(self.w1d, self.w1d).reshape(-1))
        else:
            Medge = np.linalg.inv(V2 @ V2.T)
        Emat = np.zeros((self.Np, self.NfpTot))
        for f in range(6):
            Emat[np.ix_(self.Fmask[f], np.arange(f * self.Nfp, (f + 1) * self.Nfp))] = Medge
        Lift = invM @ Emat
        return dict(V=V, invV=invV, Dx1=Dx1, Dx2=Dx2, Dx3=Dx3, Lift=Lift, invM=invM)
