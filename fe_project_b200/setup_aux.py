"""Set-up products the reference driver computes once before the first step (host side, NumPy).

`calc_phyd_hgrad` restates atm_dyn_dgm_nonhydro3d_common_calc_phyd_hgrad_lc
(FElib/src/fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_common.F90:624-777), which the driver calls
from `update_phyd_hgrad` after a restart is read; the resulting DPhydDx/DPhydDy are inputs of every
explicit tendency evaluation (rhot_heve.F90:461-464).  Vectorised over elements.
"""
from __future__ import annotations

import numpy as np

from .element import HexElement
from .mesh import LocalMeshCube


def calc_phyd_hgrad(elem: HexElement, mesh: LocalMeshCube, PRES_hyd: np.ndarray, PRES_hyd_ref: np.ndarray | None = None):
    """PRES_hyd: (NeA, Np) with interior filled.  Returns DPhydDx, DPhydDy as (NeA, Np) (halo zero)."""
    n, Np, Nfp, Ne = elem.np1, elem.Np, elem.Nfp, mesh.Ne
    P = np.array(PRES_hyd, dtype=np.float64).reshape(mesh.NeA, Np).copy()
    if PRES_hyd_ref is not None:
        P -= np.asarray(PRES_hyd_ref).reshape(mesh.NeA, Np)
    mesh.exchange_halo_numpy(P)
    flat = P.reshape(-1)
    G = mesh.Gsqrt.reshape(-1)
    G13, G23 = mesh.GI3[0].reshape(-1), mesh.GI3[1].reshape(-1)
    h2d = np.empty(elem.NfpTot, dtype=np.int64)          # IndexH2Dto3D_bnd
    a = np.arange(Nfp) % n
    h2d[0 * Nfp:1 * Nfp] = a
    h2d[1 * Nfp:2 * Nfp] = (n - 1) + a * n
    h2d[2 * Nfp:3 * Nfp] = a + (n - 1) * n
    h2d[3 * Nfp:4 * Nfp] = a * n
    h2d[4 * Nfp:5 * Nfp] = np.arange(Nfp)
    h2d[5 * Nfp:6 * Nfp] = np.arange(Nfp)
    gH_f = mesh.GsqrtH[mesh.EMap3Dto2D][:, h2d]          # (Ne, NfpTot)
    iM, iP = mesh.VMapM, mesh.VMapP
    GvM, GvP = G[iM] / gH_f, G[iP] / gH_f
    nx, ny, nz = mesh.normal_fn
    t1 = mesh.Fscale * 0.5 * GvP * flat[iP]
    t2 = mesh.Fscale * 0.5 * GvM * flat[iM]
    delx = (nx + G13[iP] * nz) * t1 - (nx + G13[iM] * nz) * t2
    dely = (ny + G23[iP] * nz) * t1 - (ny + G23[iM] * nz) * t2
    gH = mesh.GsqrtH[mesh.EMap3Dto2D][:, elem.IndexH2Dto3D]
    Gi = mesh.Gsqrt[:Ne]
    Gv = Gi / gH
    F1 = (Gv * P[:Ne]).reshape(Ne, n, n, n)               # [ke, k, j, i]
    F3 = (mesh.GI3[0, :Ne] * Gv * P[:Ne]).reshape(Ne, n, n, n)
    Fz = (mesh.GI3[1, :Ne] * Gv * P[:Ne]).reshape(Ne, n, n, n)
    D = elem.D1D
    dx = np.einsum("il,ekjl->ekji", D, F1).reshape(Ne, Np)
    dy = np.einsum("jl,ekli->ekji", D, F1).reshape(Ne, Np)
    dz1 = np.einsum("kl,elji->ekji", D, F3).reshape(Ne, Np)
    dz2 = np.einsum("kl,elji->ekji", D, Fz).reshape(Ne, Np)

    def lift(df):
        d6 = df.reshape(Ne, 6, n, n)                      # [ke, f, b, a]
        lw = elem.lift1d
        out = (lw[None, None, :, None, 0] * d6[:, 0][:, :, None, :]      # y-: (i,k) -> [k, :, i]
               + lw[None, None, None, :, 1] * d6[:, 1][:, :, :, None]    # x+: (j,k) -> [k, j, :]
               + lw[None, None, :, None, 1] * d6[:, 2][:, :, None, :]
               + lw[None, None, None, :, 0] * d6[:, 3][:, :, :, None]
               + lw[None, :, None, None, 0] * d6[:, 4][:, None, :, :]    # z-: (i,j) -> [:, j, i]
               + lw[None, :, None, None, 1] * d6[:, 5][:, None, :, :])
        return out.reshape(Ne, Np)

    E11, E22, E33 = mesh.Escale[0, 0], mesh.Escale[1, 1], mesh.Escale[2, 2]
    gx = E11 * dx + E33 * dz1 + lift(delx)
    gy = E22 * dy + E33 * dz2 + lift(dely)
    outx, outy = np.zeros((mesh.NeA, Np)), np.zeros((mesh.NeA, Np))
    outx[:Ne] = gx / Gv
    outy[:Ne] = gy / Gv
    return outx, outy
