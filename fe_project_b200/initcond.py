"""Synthetic initial states of the BASELINE configs (host side, NumPy).

Each function restates the analytic initial condition of a reference test case so that
the CUDA path and the CPU oracle start from identical arrays:

* sound wave (sample/euler3d_hevi/test_euler3d_hevi.f90:1106-1149): uniform background, cosine pulse in DRHOT.
* density current (Straka et al. 1993):
  model/atm_nonhydro3d/test/case/density_current/mod_user.F90:154-254,
  model/atm_nonhydro3d/src/preprocess/mod_mkinit_util.F90:54-165 (cosine bell + L2 projection),
  FElib/src/fluid_dyn_solver/scale_atm_dyn_dgm_hydrostatic.F90:129-170 (constant-PT base state).

Constants are SCALE's `scale_const` values (external library); they
are data handed to the C ABI, never baked into kernels.
"""
from __future__ import annotations

import numpy as np

from .element import HexElement, LineElement
from .mesh import LocalMeshCube

SCALE_CONST = dict(GRAV=9.80665, Rdry=287.04, CPdry=1004.64, CVdry=1004.64 - 287.04,
                   PRES00=1.0e5, OHM=7.2920e-5, RPlanet=6.37122e6, EPS=2.220446e-16)


def hydrostatic_const_pt(z: np.ndarray, pot_temp0: float, pres_sfc: float, c=SCALE_CONST):
    RovCP = c["Rdry"] / c["CPdry"]
    CPovR = c["CPdry"] / c["Rdry"]
    exner_sfc = (pres_sfc / c["PRES00"]) ** RovCP
    exner = exner_sfc - c["GRAV"] / (c["CPdry"] * pot_temp0) * z
    pres = c["PRES00"] * exner ** CPovR
    dens = pres / (c["Rdry"] * exner * pot_temp0)
    return dens, pres


def cosine_bell_projected(mesh: LocalMeshCube, qmax, rx, ry, rz, xc, yc, zc, intrp_order: int):
    """q(Ne, Np): cosine bell sampled on an order-`intrp_order` LGL element, L2-projected to the mesh element."""
    e = mesh.elem
    T1, src = e.l2proj_from(intrp_order)            # (np1, nq)
    xq = src.x
    nq = src.Np
    NeX, NeY, NeZ = mesh.NeX, mesh.NeY, mesh.NeZ
    vx = (mesh.xmax - mesh.xmin) * np.arange(NeX + 1) / NeX + mesh.xmin
    vy = (mesh.ymax - mesh.ymin) * np.arange(NeY + 1) / NeY + mesh.ymin
    vz = mesh.FZ
    out = np.empty((mesh.Ne, e.Np))
    h = 0.5 * (xq + 1.0)
    CH = 512                                        # elements per batch: three batched 1D contractions instead of one einsum per element
    for k0 in range(0, mesh.Ne, CH):
        sl = slice(k0, min(k0 + CH, mesh.Ne))
        ex, ey, ez = mesh.ex[sl], mesh.ey[sl], mesh.ez[sl]
        x = vx[ex][:, None] + h[None, :] * (vx[ex + 1] - vx[ex])[:, None]        # (B, nq)
        y = vy[ey][:, None] + h[None, :] * (vy[ey + 1] - vy[ey])[:, None]
        z = vz[ez][:, None] + h[None, :] * (vz[ez + 1] - vz[ez])[:, None]
        r = np.sqrt(((x[:, None, None, :] - xc) / rx) ** 2 + ((y[:, None, :, None] - yc) / ry) ** 2
                    + ((z[:, :, None, None] - zc) / rz) ** 2)
        q = np.where(r <= 1.0, qmax * (0.5 * (1.0 + np.cos(np.pi * r))), 0.0)      # [B, kq, jq, iq]
        q = q @ T1.T                                                               # [B, kq, jq, i]
        q = np.einsum("jb,Bcbi->Bcji", T1, q, optimize=True)
        q = np.einsum("kc,Bcji->Bkji", T1, q, optimize=True)
        out[sl] = q.reshape(q.shape[0], -1)
    return out


def density_current(mesh: LocalMeshCube, theta0=300.0, dtheta=-15.0, xc=0.0, yc=0.0, zc=3.0e3,
                    rx=4.0e3, ry=1.0e13, rz=2.0e3, intrp_order=11, c=SCALE_CONST):
    """Returns dict of (NeA, Np) arrays: DDENS, MOMX, MOMY, MOMZ, DRHOT, DENS_hyd, PRES_hyd (halo part zero)."""
    e = mesh.elem
    Np, Ne, NeA = e.Np, mesh.Ne, mesh.NeA
    z = mesh.pos_en[2]
    dens_hyd, pres_hyd = hydrostatic_const_pt(z, theta0, c["PRES00"], c)
    dth = cosine_bell_projected(mesh, dtheta, rx, ry, rz, xc, yc, zc, intrp_order)
    PT = theta0 + dth
    RovCp = c["Rdry"] / c["CPdry"]
    DENS = pres_hyd / (c["Rdry"] * PT * (pres_hyd / c["PRES00"]) ** RovCp)
    f = {k: np.zeros((NeA, Np)) for k in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT", "DENS_hyd", "PRES_hyd")}
    f["DENS_hyd"][:Ne] = dens_hyd
    f["PRES_hyd"][:Ne] = pres_hyd
    f["DDENS"][:Ne] = DENS - dens_hyd
    f["DRHOT"][:Ne] = DENS * PT - dens_hyd * theta0
    return f


def sound_wave(mesh: LocalMeshCube, amplitude=1.0e-12, dens0=1.0, pres0=1.0e5):
    """Initial state of sample/euler3d_hevi (set_initcond_lc, test_euler3d_hevi.f90:1106-1149): DENS_hyd = 1,
    PRES_hyd = 1e5, zero momentum, DRHOT = A cos(pi r / 2) for |r| <= 1 with r = (z - z_mid) / (0.1 Lz).
    The shipped amplitude is 1e-12."""
    e = mesh.elem
    Np, Ne, NeA = e.Np, mesh.Ne, mesh.NeA
    z = mesh.pos_en[2]
    r = (z - 0.5 * (mesh.zmax + mesh.zmin)) / ((mesh.zmax - mesh.zmin) * 0.1)
    f = {k: np.zeros((NeA, Np)) for k in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT", "DENS_hyd", "PRES_hyd")}
    f["DENS_hyd"][:Ne] = dens0
    f["PRES_hyd"][:Ne] = pres0
    f["DRHOT"][:Ne] = np.where(np.abs(r) <= 1.0, amplitude * np.cos(0.5 * np.pi * r), 0.0)
    return f
