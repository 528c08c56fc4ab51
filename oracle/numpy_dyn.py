"""Second, independent restatement of the regional HEVE dynamics rows in NumPy: test infrastructure, written from
the Fortran, not from oracle/dyn_heve.cpp, vectorised over elements with dense tensor contractions where the C++ restatement loops.

  numflux_heve   atm_dyn_dgm_nonhydro3d_rhot_heve_numflux_get_generalvc   fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_rhot_heve_numflux.F90:946-1138
  cal_tend_heve  atm_dyn_dgm_nonhydro3d_rhot_heve_cal_tend                fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_rhot_heve.F90:292-489
  drhot2pres     atm_dyn_dgm_nonhydro3d_common_DRHOT2PRES                 fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_common.F90:428-479
  apply_bc       AtmDynBnd%ApplyBC_PROGVARS_lc (SLIP / NOSLIP, flat)      fluid_dyn_solver/scale_atm_dyn_dgm_bnd.F90:270-367
  numflux_hevi   atm_dyn_dgm_nonhydro3d_rhot_hevi_numflux_get_generalvc   fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_rhot_hevi_numflux.F90:232-416
  cal_tend_hevi  atm_dyn_dgm_nonhydro3d_rhot_hevi_cal_tend                fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_rhot_hevi.F90:289-482
  numflux_hevi_global    ..._rhot_hevi_numflux_get_generalhvc             fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_rhot_hevi_numflux.F90:606-834
  cal_tend_hevi_global   atm_dyn_dgm_globalnonhydro3d_rhot_hevi_cal_tend  fluid_dyn_solver/scale_atm_dyn_dgm_globalnonhydro3d_rhot_hevi.F90:337-583

Inputs are the host-side mesh / element objects of fe_project_b200 (themselves an independent restatement of the set-up code) and flat
(NeA * Np) field arrays with the halo part filled.  tests/test_oracle_numpy_dyn.py asserts agreement with the C++ oracle to 1e-13."""
from __future__ import annotations

import numpy as np


def drhot2pres(c, DRHOT, PRES_hyd, THERM_hyd, Rtot, CVtot, CPtot):
    pres = c["PRES00"] * (Rtot / c["PRES00"] * (THERM_hyd + DRHOT)) ** (CPtot / CVtot)
    return pres, pres - PRES_hyd


def numflux_heve(elem, mesh, c, q, aux, DPRES):
    """q: dict DDENS, MOMX, MOMY, MOMZ, DRHOT; aux: dict DENS_hyd, PRES_hyd, THERM_hyd -- flat arrays incl. halo.
    Returns del_flux[var] (Ne, NfpTot) in the order DENS, RHOT, MOMZ, MOMX, MOMY."""
    iM, iP = mesh.VMapM, mesh.VMapP                      # (Ne, NfpTot)
    nx, ny, nz = mesh.normal_fn
    G = mesh.Gsqrt.reshape(-1)
    G13, G23 = mesh.GI3[0].reshape(-1), mesh.GI3[1].reshape(-1)
    gamm = c["CPdry"] / c["CVdry"]
    side = {}
    for tag, idx in (("IN", iM), ("EX", iP)):
        Gs = G[idx]
        s = dict(Gs=Gs, RGv=1.0 / Gs, G13=G13[idx], G23=G23[idx])
        for nm in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT"):
            s[nm] = Gs * q[nm][idx]
        s["Phyd"] = aux["PRES_hyd"][idx]
        s["dp"] = DPRES[idx]
        s["Dens"] = s["DDENS"] + Gs * aux["DENS_hyd"][idx]
        s["Rhot"] = Gs * aux["THERM_hyd"][idx] + s["DRHOT"]
        s["Vel"] = (s["MOMX"] * nx + s["MOMY"] * ny + ((s["MOMZ"] * s["RGv"] + s["G13"] * s["MOMX"] + s["G23"] * s["MOMY"]) * nz)) / s["Dens"]
        side[tag] = s
    I, E = side["IN"], side["EX"]
    t1 = np.abs(nx) + np.abs(ny)
    Gnn_M = t1 + (1.0 * I["RGv"] ** 2 + I["G13"] ** 2 + I["G23"] ** 2) * np.abs(nz)
    Gnn_P = t1 + (1.0 * E["RGv"] ** 2 + E["G13"] ** 2 + E["G23"] ** 2) * np.abs(nz)
    alpha = np.maximum(np.sqrt(Gnn_M * gamm * (I["Phyd"] + I["dp"]) * I["Gs"] / I["Dens"]) + np.abs(I["Vel"]),
                       np.sqrt(Gnn_P * gamm * (E["Phyd"] + E["dp"]) * E["Gs"] / E["Dens"]) + np.abs(E["Vel"]))
    hf = mesh.Fscale * 0.5
    out = {}
    out["DENS"] = hf * (E["Dens"] * E["Vel"] - I["Dens"] * I["Vel"] - alpha * (E["DDENS"] - I["DDENS"]))
    out["RHOT"] = hf * (E["Rhot"] * E["Vel"] - I["Rhot"] * I["Vel"] - alpha * (E["DRHOT"] - I["DRHOT"]))
    t3, t4 = E["Gs"] * E["dp"], I["Gs"] * I["dp"]
    mom1 = (t3 * E["RGv"] - t4 * I["RGv"]) * nz
    mom2 = (nx + E["G13"] * nz) * t3 - (nx + I["G13"] * nz) * t4
    mom3 = (ny + E["G23"] * nz) * t3 - (ny + I["G23"] * nz) * t4
    out["MOMZ"] = hf * (E["MOMZ"] * E["Vel"] - I["MOMZ"] * I["Vel"] + mom1 - alpha * (E["MOMZ"] - I["MOMZ"]))
    out["MOMX"] = hf * (E["MOMX"] * E["Vel"] - I["MOMX"] * I["Vel"] + mom2 - alpha * (E["MOMX"] - I["MOMX"]))
    out["MOMY"] = hf * (E["MOMY"] * E["Vel"] - I["MOMY"] * I["Vel"] + mom3 - alpha * (E["MOMY"] - I["MOMY"]))
    return out


def _div_lift(elem, F1, F2, F3, dflux, Ne):
    """Div_var5 of one variable: (Dx F1, Dy F2, Dz F3, Lift del_flux), each (Ne, Np)."""
    n = elem.np1
    D = elem.D1D
    a = F1.reshape(Ne, n, n, n); b = F2.reshape(Ne, n, n, n); cc = F3.reshape(Ne, n, n, n)     # [ke, k, j, i]
    dx = np.einsum("il,ekjl->ekji", D, a).reshape(Ne, -1)
    dy = np.einsum("jl,ekli->ekji", D, b).reshape(Ne, -1)
    dz = np.einsum("kl,elji->ekji", D, cc).reshape(Ne, -1)
    lift = dflux @ elem.lift_dense().T                                                     # dense Lift (Np, NfpTot)
    return dx, dy, dz, lift


def cal_tend_heve(elem, mesh, c, q, aux, DPRES, DPhydDx=None, DPhydDy=None, coriolis=None):
    """Explicit HEVE tendency of the interior elements: dict DENS_dt, RHOT_dt, MOMZ_dt, MOMX_dt, MOMY_dt, each (Ne, Np)."""
    Ne, Np = mesh.Ne, elem.Np
    ni = Ne * Np
    dfl = numflux_heve(elem, mesh, c, q, aux, DPRES)
    sh = lambda a: np.asarray(a).reshape(-1)[:ni].reshape(Ne, Np)
    G = sh(mesh.Gsqrt)
    gH = mesh.GsqrtH[mesh.EMap3Dto2D][:, elem.IndexH2Dto3D]
    RGv = 1.0 / (G / gH)
    RG = 1.0 / G
    GI1, GI2 = sh(mesh.GI3[0]), sh(mesh.GI3[1])
    dd, mx, my, mz, dr = (sh(q[k]) for k in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT"))
    RD = 1.0 / (dd + sh(aux["DENS_hyd"]))
    dp = sh(DPRES)
    F = {}
    F["DENS"] = (G * mx, G * my, G * (mz * RGv + GI1 * mx + GI2 * my))
    pt = (sh(aux["THERM_hyd"]) + dr) * RD
    F["RHOT"] = tuple(f * pt for f in F["DENS"])
    w = mz * RD
    F["MOMZ"] = (F["DENS"][0] * w, F["DENS"][1] * w, F["DENS"][2] * w + G * dp * RGv)
    gdp = G * dp
    u = mx * RD
    F["MOMX"] = (F["DENS"][0] * u + gdp, F["DENS"][1] * u, F["DENS"][2] * u + gdp * GI1)
    v = my * RD
    F["MOMY"] = (F["DENS"][0] * v, F["DENS"][1] * v + gdp, F["DENS"][2] * v + gdp * GI2)
    E11, E22, E33 = mesh.Escale[0, 0], mesh.Escale[1, 1], mesh.Escale[2, 2]
    out = {}
    for nm in ("DENS", "RHOT", "MOMZ", "MOMX", "MOMY"):
        dx, dy, dz, lift = _div_lift(elem, *F[nm], dfl[nm], Ne)
        out[nm + "_dt"] = -(E11 * dx + E22 * dy + E33 * dz + lift) * RG
    n = elem.np1
    drho = np.einsum("kl,elji->ekji", elem.VPOrdM1, dd.reshape(Ne, n, n, n)).reshape(Ne, Np)
    out["MOMZ_dt"] = out["MOMZ_dt"] - c["GRAV"] * drho
    cor = 0.0 if coriolis is None else np.asarray(coriolis).reshape(mesh.Ne2D, -1)[mesh.EMap3Dto2D][:, elem.IndexH2Dto3D]
    gx = 0.0 if DPhydDx is None else sh(DPhydDx)
    gy = 0.0 if DPhydDy is None else sh(DPhydDy)
    out["MOMX_dt"] = (-gx + cor * my) + out["MOMX_dt"]
    out["MOMY_dt"] = (-gy - cor * mx) + out["MOMY_dt"]
    return out


def numflux_hevi(elem, mesh, c, q, aux, DPRES):
    """Horizontally explicit Rusanov flux of the HEVI equation set (rhot_hevi_numflux.F90:304-414): the dissipation coefficient carries
    swV = 1 - nz^2 (none on the vertical faces), mass and theta are advected with the HORIZONTAL velocity only, MOMZ has no vertical
    pressure term, GsqrtV_ = Gsqrt_ inside the flux (:320).  Same argument conventions as numflux_heve."""
    iM, iP = mesh.VMapM, mesh.VMapP
    nx, ny, nz = mesh.normal_fn
    G = mesh.Gsqrt.reshape(-1)
    G13, G23 = mesh.GI3[0].reshape(-1), mesh.GI3[1].reshape(-1)
    gamm = c["CPdry"] / c["CVdry"]
    side = {}
    for tag, idx in (("IN", iM), ("EX", iP)):
        Gs = G[idx]
        s = dict(Gs=Gs, RGv=1.0 / Gs, G13=G13[idx], G23=G23[idx])
        for nm in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT"):
            s[nm] = Gs * q[nm][idx]
        s["Phyd"] = aux["PRES_hyd"][idx]
        s["dp"] = DPRES[idx]
        s["Dens"] = s["DDENS"] + Gs * aux["DENS_hyd"][idx]
        s["Rhot"] = Gs * aux["THERM_hyd"][idx] + s["DRHOT"]
        s["Velh"] = (s["MOMX"] * nx + s["MOMY"] * ny) / s["Dens"]
        s["Vel"] = s["Velh"] + ((s["MOMZ"] * s["RGv"] + s["G13"] * s["MOMX"] + s["G23"] * s["MOMY"]) * nz) / s["Dens"]
        side[tag] = s
    I, E = side["IN"], side["EX"]
    swV = 1.0 - nz ** 2
    alpha = swV * np.maximum(np.sqrt(gamm * (I["Phyd"] + I["dp"]) * I["Gs"] / I["Dens"]) + np.abs(I["Vel"]),
                             np.sqrt(gamm * (E["Phyd"] + E["dp"]) * E["Gs"] / E["Dens"]) + np.abs(E["Vel"]))
    hf = mesh.Fscale * 0.5
    out = {}
    out["DENS"] = hf * (E["Dens"] * E["Velh"] - I["Dens"] * I["Velh"] + (-alpha * (E["DDENS"] - I["DDENS"])))
    out["RHOT"] = hf * (E["Rhot"] * E["Velh"] - I["Rhot"] * I["Velh"] + (-alpha * (E["DRHOT"] - I["DRHOT"])))
    out["MOMZ"] = hf * (E["MOMZ"] * E["Vel"] - I["MOMZ"] * I["Vel"] + (-alpha * (E["MOMZ"] - I["MOMZ"])))
    t3, t4 = E["Gs"] * E["dp"], I["Gs"] * I["dp"]
    mom1 = (nx + E["G13"] * nz) * t3 - (nx + I["G13"] * nz) * t4
    mom2 = (ny + E["G23"] * nz) * t3 - (ny + I["G23"] * nz) * t4
    out["MOMX"] = hf * (E["MOMX"] * E["Vel"] - I["MOMX"] * I["Vel"] + mom1 + (-alpha * (E["MOMX"] - I["MOMX"])))
    out["MOMY"] = hf * (E["MOMY"] * E["Vel"] - I["MOMY"] * I["Vel"] + mom2 + (-alpha * (E["MOMY"] - I["MOMY"])))
    return out


def cal_tend_hevi(elem, mesh, c, q, aux, DPRES, DPhydDx=None, DPhydDy=None, coriolis=None):
    """Horizontally explicit tendency of the HEVI equation set (rhot_hevi.F90:372-478): DENS and RHOT take the horizontal derivatives
    and the lift only, MOMZ has no pressure term and no buoyancy here (both are in the vertical-implicit part), MOMX / MOMY as in HEVE."""
    Ne, Np = mesh.Ne, elem.Np
    ni = Ne * Np
    dfl = numflux_hevi(elem, mesh, c, q, aux, DPRES)
    sh = lambda a: np.asarray(a).reshape(-1)[:ni].reshape(Ne, Np)
    G = sh(mesh.Gsqrt)
    gH = mesh.GsqrtH[mesh.EMap3Dto2D][:, elem.IndexH2Dto3D]
    RGv = 1.0 / (G / gH)
    RG = 1.0 / G
    GI1, GI2 = sh(mesh.GI3[0]), sh(mesh.GI3[1])
    dd, mx, my, mz, dr = (sh(q[k]) for k in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT"))
    RD = 1.0 / (dd + sh(aux["DENS_hyd"]))
    gdp = G * sh(DPRES)
    Fd = (G * mx, G * my, G * (mz * RGv + GI1 * mx + GI2 * my))
    pt = (sh(aux["THERM_hyd"]) + dr) * RD
    w, u, v = mz * RD, mx * RD, my * RD
    zero = np.zeros_like(G)
    F = {"DENS": (Fd[0], Fd[1], zero), "RHOT": (Fd[0] * pt, Fd[1] * pt, zero),
         "MOMZ": (Fd[0] * w, Fd[1] * w, Fd[2] * w),
         "MOMX": (Fd[0] * u + gdp, Fd[1] * u, Fd[2] * u + gdp * GI1),
         "MOMY": (Fd[0] * v, Fd[1] * v + gdp, Fd[2] * v + gdp * GI2)}
    E11, E22, E33 = mesh.Escale[0, 0], mesh.Escale[1, 1], mesh.Escale[2, 2]
    out = {}
    for nm in ("DENS", "RHOT", "MOMZ", "MOMX", "MOMY"):
        dx, dy, dz, lift = _div_lift(elem, *F[nm], dfl[nm], Ne)
        vert = 0.0 if nm in ("DENS", "RHOT") else E33 * dz      # the vertical mass / theta fluxes belong to the implicit part
        out[nm + "_dt"] = -(E11 * dx + E22 * dy + vert + lift) * RG
    cor = 0.0 if coriolis is None else np.asarray(coriolis).reshape(mesh.Ne2D, -1)[mesh.EMap3Dto2D][:, elem.IndexH2Dto3D]
    gx = 0.0 if DPhydDx is None else sh(DPhydDx)
    gy = 0.0 if DPhydDy is None else sh(DPhydDy)
    out["MOMX_dt"] = (-gx + cor * my) + out["MOMX_dt"]
    out["MOMY_dt"] = (-gy - cor * mx) + out["MOMY_dt"]
    return out


def _face_h2d(elem, mesh):
    """elem%IndexH2Dto3D_bnd: horizontal (2D) node of every face node, and the 2D element of every element, broadcastable to (Ne, NfpTot)."""
    return mesh.EMap3Dto2D[:, None], (mesh.VMapM % elem.Np) % elem.Nfp


def numflux_hevi_global(elem, mesh, c, q, aux, DPRES):
    """numflux_get_generalhvc (rhot_hevi_numflux.F90:702-830) on a cubed-sphere panel: the horizontal metric G11 / G12 / G22 and GsqrtH
    of the OWN element's 2D node on both sides (iM2Dto3D), rgam2 = 1 / gam^2, Gnn with the metric of the face direction."""
    iM, iP = mesh.VMapM, mesh.VMapP
    nx, ny, nz = mesh.normal_fn
    G = mesh.Gsqrt.reshape(-1)
    gam = mesh.gam.reshape(-1)
    G13, G23 = mesh.GI3[0].reshape(-1), mesh.GI3[1].reshape(-1)
    k2, h2 = _face_h2d(elem, mesh)
    GH = mesh.GsqrtH[k2, h2]
    G11, G12, G22 = mesh.GIJ[0, 0][k2, h2], mesh.GIJ[0, 1][k2, h2], mesh.GIJ[1, 1][k2, h2]
    gamm = c["CPdry"] / c["CVdry"]
    side = {}
    for tag, idx in (("IN", iM), ("EX", iP)):
        Gs = G[idx]
        rgam2 = 1.0 / gam[idx] ** 2
        s = dict(Gs=Gs, rgam2=rgam2, RGv=1.0 / (Gs * rgam2 / GH), G13=G13[idx], G23=G23[idx])
        for nm in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT"):
            s[nm] = Gs * q[nm][idx]
        s["Phyd"] = aux["PRES_hyd"][idx]
        s["dp"] = DPRES[idx]
        s["Dens"] = s["DDENS"] + Gs * aux["DENS_hyd"][idx]
        s["Rhot"] = Gs * aux["THERM_hyd"][idx] + s["DRHOT"]
        s["Velh"] = (s["MOMX"] * nx + s["MOMY"] * ny) / s["Dens"]
        s["Vel"] = s["Velh"] + ((s["MOMZ"] * s["RGv"] + s["G13"] * s["MOMX"] + s["G23"] * s["MOMY"]) * nz) / s["Dens"]
        s["Gxz"] = rgam2 * (G11 * s["G13"] + G12 * s["G23"])
        s["Gyz"] = rgam2 * (G12 * s["G13"] + G22 * s["G23"])
        s["G1n"] = rgam2 * (G11 * nx + G12 * ny)
        s["G2n"] = rgam2 * (G12 * nx + G22 * ny)
        s["Gnn"] = rgam2 * (np.abs(G11 * nx) + np.abs(G22 * ny)) + (1.0 * s["RGv"] ** 2 + s["G13"] * s["Gxz"] + s["G23"] * s["Gyz"]) * np.abs(nz)
        side[tag] = s
    I, E = side["IN"], side["EX"]
    swV = 1.0 - nz ** 2
    alpha = swV * np.maximum(np.sqrt(I["Gnn"] * gamm * (I["Phyd"] + I["dp"]) * I["Gs"] / I["Dens"]) + np.abs(I["Vel"]),
                             np.sqrt(E["Gnn"] * gamm * (E["Phyd"] + E["dp"]) * E["Gs"] / E["Dens"]) + np.abs(E["Vel"]))
    hf = mesh.Fscale * 0.5
    out = {}
    out["DENS"] = hf * (E["Dens"] * E["Velh"] - I["Dens"] * I["Velh"] + (-alpha * (E["DDENS"] - I["DDENS"])))
    out["RHOT"] = hf * (E["Rhot"] * E["Velh"] - I["Rhot"] * I["Velh"] + (-alpha * (E["DRHOT"] - I["DRHOT"])))
    out["MOMZ"] = hf * (E["MOMZ"] * E["Vel"] - I["MOMZ"] * I["Vel"] + (-alpha * (E["MOMZ"] - I["MOMZ"])))
    t3, t4 = E["Gs"] * E["dp"], I["Gs"] * I["dp"]
    mom1 = (E["G1n"] + E["Gxz"] * nz) * t3 - (I["G1n"] + I["Gxz"] * nz) * t4
    mom2 = (E["G2n"] + E["Gyz"] * nz) * t3 - (I["G2n"] + I["Gyz"] * nz) * t4
    out["MOMX"] = hf * (E["MOMX"] * E["Vel"] - I["MOMX"] * I["Vel"] + mom1 + (-alpha * (E["MOMX"] - I["MOMX"])))
    out["MOMY"] = hf * (E["MOMY"] * E["Vel"] - I["MOMY"] * I["Vel"] + mom2 + (-alpha * (E["MOMY"] - I["MOMY"])))
    return out


def cal_tend_hevi_global(elem, mesh, c, q, aux, DPRES, DPhydDx, DPhydDy):
    """Horizontally explicit tendency of GLOBALNONHYDRO3D_HEVI on one panel (globalnonhydro3d_rhot_hevi.F90:421-578): contravariant
    pressure-gradient terms G11 / G12 / G22, the Christoffel terms of the equiangular gnomonic map and the Coriolis term (sign s = -1 on
    panel 6, the factor s Y on the equatorial panels)."""
    Ne, Np = mesh.Ne, elem.Np
    ni = Ne * Np
    dfl = numflux_hevi_global(elem, mesh, c, q, aux, DPRES)
    sh = lambda a: np.asarray(a).reshape(-1)[:ni].reshape(Ne, Np)
    to3 = lambda a2: a2[mesh.EMap3Dto2D][:, elem.IndexH2Dto3D]              # (Ne2D, Nfp) -> (Ne, Np)
    G = sh(mesh.Gsqrt)
    rgam2 = 1.0 / sh(mesh.gam) ** 2
    G11, G12, G22 = to3(mesh.GIJ[0, 0]) * rgam2, to3(mesh.GIJ[0, 1]) * rgam2, to3(mesh.GIJ[1, 1]) * rgam2
    RGv = 1.0 / (G * rgam2 / to3(mesh.GsqrtH))
    RG = 1.0 / G
    GI1, GI2 = sh(mesh.GI3[0]), sh(mesh.GI3[1])
    dd, mx, my, mz, dr = (sh(q[k]) for k in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT"))
    RD = 1.0 / (dd + sh(aux["DENS_hyd"]))
    gdp = G * sh(DPRES)
    Fd = (G * mx, G * my, G * (mz * RGv + GI1 * mx + GI2 * my))
    pt = (sh(aux["THERM_hyd"]) + dr) * RD
    w, u, v = mz * RD, mx * RD, my * RD
    zero = np.zeros_like(G)
    F = {"DENS": (Fd[0], Fd[1], zero), "RHOT": (Fd[0] * pt, Fd[1] * pt, zero),
         "MOMZ": (Fd[0] * w, Fd[1] * w, Fd[2] * w),
         "MOMX": (Fd[0] * u + G11 * gdp, Fd[1] * u + G12 * gdp, Fd[2] * u + gdp * (G11 * GI1 + G12 * GI2)),
         "MOMY": (Fd[0] * v + G12 * gdp, Fd[1] * v + G22 * gdp, Fd[2] * v + gdp * (G12 * GI1 + G22 * GI2))}
    E11, E22, E33 = mesh.Escale[0, 0], mesh.Escale[1, 1], mesh.Escale[2, 2]
    out = {}
    for nm in ("DENS", "RHOT", "MOMZ", "MOMX", "MOMY"):
        dx, dy, dz, lift = _div_lift(elem, *F[nm], dfl[nm], Ne)
        vert = 0.0 if nm in ("DENS", "RHOT") else E33 * dz
        out[nm + "_dt"] = -(E11 * dx + E22 * dy + vert + lift) * RG
    X, Y = to3(np.tan(mesh.pos2D[0])), to3(np.tan(mesh.pos2D[1]))
    two = 2.0 / (1.0 + X ** 2 + Y ** 2)
    sgn = -1.0 if mesh.panelID == 6 else 1.0
    cori1 = sgn * c["OHM"] * two * (-X * Y * mx + (1.0 + Y ** 2) * my)
    cori2 = sgn * c["OHM"] * two * (-(1.0 + X ** 2) * mx + X * Y * my)
    if mesh.panelID <= 4:
        cori1, cori2 = sgn * Y * cori1, sgn * Y * cori2
    gx, gy = sh(DPhydDx), sh(DPhydDy)
    out["MOMX_dt"] = (-(G11 * gx + G12 * gy) - two * Y * (X * Y * u - (1.0 + Y ** 2) * v) * mx + cori1) + out["MOMX_dt"]
    out["MOMY_dt"] = (-(G12 * gx + G22 * gy) - two * X * (-(1.0 + X ** 2) * u + X * Y * v) * my + cori2) + out["MOMY_dt"]
    return out
