"""Second, independent restatement of the regional HEVE dynamics rows in NumPy: test infrastructure, written from
the Fortran, not from oracle/dyn_heve.cpp, vectorised over elements with dense tensor contractions where the C++ restatement loops.

  numflux_heve   atm_dyn_dgm_nonhydro3d_rhot_heve_numflux_get_generalvc   fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_rhot_heve_numflux.F90:946-1138
  cal_tend_heve  atm_dyn_dgm_nonhydro3d_rhot_heve_cal_tend                fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_rhot_heve.F90:292-489
  drhot2pres     atm_dyn_dgm_nonhydro3d_common_DRHOT2PRES                 fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_common.F90:428-479
  numflux_hevi   atm_dyn_dgm_nonhydro3d_rhot_hevi_numflux_get_generalvc   fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_rhot_hevi_numflux.F90:232-416
  cal_tend_hevi  atm_dyn_dgm_nonhydro3d_rhot_hevi_cal_tend                fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_rhot_hevi.F90:289-482
  numflux_hevi_global    ..._rhot_hevi_numflux_get_generalhvc             fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_rhot_hevi_numflux.F90:606-834
  cal_tend_hevi_global   atm_dyn_dgm_globalnonhydro3d_rhot_hevi_cal_tend  fluid_dyn_solver/scale_atm_dyn_dgm_globalnonhydro3d_rhot_hevi.F90:337-583
  apply_bc       AtmDynBnd%ApplyBC_PROGVARS_lc (slip / no-slip walls, regional mesh incl. the terrain-following metric)
                                                                          fluid_dyn_solver/scale_atm_dyn_dgm_bnd.F90:270-367
  modal_filter   atm_dyn_dgm_modalfilter_apply                            fluid_dyn_solver/scale_atm_dyn_dgm_modalfilter.F90:49-130
  update         AtmDynDGMDriver_nonhydro3d%Update, the stage loop        fluid_dyn_solver/scale_atm_dyn_dgm_driver_nonhydro3d.F90:614-963
                 with the Runge-Kutta stages in their Butcher form (the reference and the C++ restatement use the low-storage / one-buffer
                 forms of scale_timeint_rk.F90: the same numbers up to round-off)
  cal_vi         atm_dyn_dgm_nonhydro3d_rhot_hevi_cal_vi (:772-965) with eval_Ax, eval_Ax_uv, vi_cal_del_flux_dyn(_uv), construct_matbnd(_uv)
                 of scale_atm_dyn_dgm_nonhydro3d_rhot_hevi_common_2.F90:111-1328 -- the column systems assembled as ONE dense matrix per column
                 and solved with numpy.linalg.solve (the reference and the C++ restatement run a block-Thomas sweep with a partial-pivot LU
                 per element: a different algorithm for the same linear system), flat and terrain-following regional meshes

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


def numflux_hevi_global(elem, mesh, c, q, aux, DPRES, hevi=True):
    """hevi = False: numflux_get_generalhvc of the HEVE set (rhot_heve_numflux.F90:1620-1770): full normal velocity for mass and theta, no
    swV factor, vertical pressure jump in MOMZ.
    numflux_get_generalhvc (rhot_hevi_numflux.F90:702-830) on a cubed-sphere panel: the horizontal metric G11 / G12 / G22 and GsqrtH
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
    swV = (1.0 - nz ** 2) if hevi else 1.0
    alpha = swV * np.maximum(np.sqrt(I["Gnn"] * gamm * (I["Phyd"] + I["dp"]) * I["Gs"] / I["Dens"]) + np.abs(I["Vel"]),
                             np.sqrt(E["Gnn"] * gamm * (E["Phyd"] + E["dp"]) * E["Gs"] / E["Dens"]) + np.abs(E["Vel"]))
    hf = mesh.Fscale * 0.5
    vm = "Velh" if hevi else "Vel"
    t3, t4 = E["Gs"] * E["dp"], I["Gs"] * I["dp"]
    pz = 0.0 if hevi else (t3 * E["RGv"] - t4 * I["RGv"]) * nz
    out = {}
    out["DENS"] = hf * (E["Dens"] * E[vm] - I["Dens"] * I[vm] + (-alpha * (E["DDENS"] - I["DDENS"])))
    out["RHOT"] = hf * (E["Rhot"] * E[vm] - I["Rhot"] * I[vm] + (-alpha * (E["DRHOT"] - I["DRHOT"])))
    out["MOMZ"] = hf * (E["MOMZ"] * E["Vel"] - I["MOMZ"] * I["Vel"] + pz + (-alpha * (E["MOMZ"] - I["MOMZ"])))
    mom1 = (E["G1n"] + E["Gxz"] * nz) * t3 - (I["G1n"] + I["Gxz"] * nz) * t4
    mom2 = (E["G2n"] + E["Gyz"] * nz) * t3 - (I["G2n"] + I["Gyz"] * nz) * t4
    out["MOMX"] = hf * (E["MOMX"] * E["Vel"] - I["MOMX"] * I["Vel"] + mom1 + (-alpha * (E["MOMX"] - I["MOMX"])))
    out["MOMY"] = hf * (E["MOMY"] * E["Vel"] - I["MOMY"] * I["Vel"] + mom2 + (-alpha * (E["MOMY"] - I["MOMY"])))
    return out


def cal_tend_hevi_global(elem, mesh, c, q, aux, DPRES, DPhydDx, DPhydDy, hevi=True):
    """hevi = False: the full tendency of GLOBALNONHYDRO3D_HEVE in the shallow-atmosphere approximation
    (globalnonhydro3d_rhot_heve.F90:338-600, cal_tend_shallow_atm): vertical mass / theta fluxes, vertical pressure gradient and buoyancy
    included.
    Horizontally explicit tendency of GLOBALNONHYDRO3D_HEVI on one panel (globalnonhydro3d_rhot_hevi.F90:421-578): contravariant
    pressure-gradient terms G11 / G12 / G22, the Christoffel terms of the equiangular gnomonic map and the Coriolis term (sign s = -1 on
    panel 6, the factor s Y on the equatorial panels)."""
    Ne, Np = mesh.Ne, elem.Np
    ni = Ne * Np
    dfl = numflux_hevi_global(elem, mesh, c, q, aux, DPRES, hevi)
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
    F = {"DENS": (Fd[0], Fd[1], zero if hevi else Fd[2]), "RHOT": (Fd[0] * pt, Fd[1] * pt, zero if hevi else Fd[2] * pt),
         "MOMZ": (Fd[0] * w, Fd[1] * w, Fd[2] * w if hevi else Fd[2] * w + G * RGv * sh(DPRES)),
         "MOMX": (Fd[0] * u + G11 * gdp, Fd[1] * u + G12 * gdp, Fd[2] * u + gdp * (G11 * GI1 + G12 * GI2)),
         "MOMY": (Fd[0] * v + G12 * gdp, Fd[1] * v + G22 * gdp, Fd[2] * v + gdp * (G12 * GI1 + G22 * GI2))}
    E11, E22, E33 = mesh.Escale[0, 0], mesh.Escale[1, 1], mesh.Escale[2, 2]
    out = {}
    for nm in ("DENS", "RHOT", "MOMZ", "MOMX", "MOMY"):
        dx, dy, dz, lift = _div_lift(elem, *F[nm], dfl[nm], Ne)
        vert = 0.0 if (hevi and nm in ("DENS", "RHOT")) else E33 * dz
        out[nm + "_dt"] = -(E11 * dx + E22 * dy + vert + lift) * RG
    if not hevi:
        n = elem.np1
        out["MOMZ_dt"] = out["MOMZ_dt"] - c["GRAV"] * np.einsum("kl,elji->ekji", elem.VPOrdM1, dd.reshape(Ne, n, n, n)).reshape(Ne, Np)
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


def cal_vi(elem, mesh, c, aux, cur, var0, impl_fac):
    """One Newton iteration of the vertical-implicit step about var0 (regional mesh, dry).  cur / var0: dicts DDENS, MOMX, MOMY, MOMZ, DRHOT
    of flat arrays (at least the Np * Ne interior values); aux: DENS_hyd, PRES_hyd.  Returns the implicit tendencies (dict, (Ne * Np,)):
    (PROG_VARS - cur) / impl_fac, or the vertical operator itself for impl_fac = 0."""
    n1, Nfp, Np, NeZ, Ne2D = elem.np1, elem.Nfp, elem.Np, mesh.NeZ, mesh.Ne2D
    ni = mesh.Ne * Np
    R4 = lambda a: np.array(np.asarray(a).reshape(-1)[:ni].reshape(NeZ, Ne2D, n1, Nfp), dtype=np.float64)    # [kz, ke2D, pv, ij]
    names = ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT")
    q0 = {k: R4(var0[k]) for k in names}
    qc = {k: R4(cur[k]) for k in names}
    pv_ = {k: q0[k].copy() for k in names}                       # PROG_VARS, starts at var0
    dh, ph = R4(aux["DENS_hyd"]), R4(aux["PRES_hyd"])
    Gv = R4(mesh.Gsqrt) / mesh.GsqrtH[None, :, None, :]           # GsqrtV = Gsqrt / GsqrtH (rhot_hevi.F90:864)
    G13, G23 = R4(mesh.GI3[0]), R4(mesh.GI3[1])
    E33 = R4(mesh.Escale[2, 2])
    Fs = np.asarray(mesh.Fscale).reshape(NeZ, Ne2D, 6, Nfp)[:, :, 4:6, :]     # vertical faces: [kz, ke2D, side, ij]
    D, VP, lw = elem.D1D, elem.VPOrdM1, elem.lift1d             # lw[pv, side]
    gamm, P00, Rd, grav = c["CPdry"] / c["CVdry"], c["PRES00"], c["Rdry"], c["GRAV"]
    rgamm = c["CVdry"] / c["CPdry"]
    rhot_hyd = P00 / Rd * (ph / P00) ** rgamm
    pres_of = lambda rhot: P00 * (Rd / P00 * rhot) ** gamm

    def faces(field):
        """interior / exterior values at the bottom and top face of every element: [kz, ke2D, side, ij]; at the column ends exterior = interior."""
        inn = np.stack([field[:, :, 0, :], field[:, :, n1 - 1, :]], axis=2)
        ext = inn.copy()
        ext[1:, :, 0, :] = field[:-1, :, n1 - 1, :]
        ext[:-1, :, 1, :] = field[1:, :, 0, :]
        return inn, ext
    nz = np.array([-1.0, 1.0])[None, None, :, None]
    wall = np.zeros((NeZ, 1, 2, 1), dtype=bool)
    wall[0, 0, 0, 0] = True; wall[NeZ - 1, 0, 1, 0] = True
    lift = lambda df: lw[None, None, :, 0, None] * df[:, :, None, 0, :] + lw[None, None, :, 1, None] * df[:, :, None, 1, :]
    Dz = lambda f: np.einsum("pl,zelj->zepj", D, f)

    # ---- dissipation coefficient of the vertical faces from var0 (vi_cal_del_flux_dyn_uv :1085-1150)
    rdens0 = 1.0 / (dh + q0["DDENS"])
    wt0 = (q0["MOMZ"] / Gv + G13 * q0["MOMX"] + G23 * q0["MOMY"]) * rdens0
    a_node = np.abs(wt0) + np.sqrt((1.0 / Gv ** 2 + G13 * G13 + G23 * G23) * gamm * pres_of(rhot_hyd + q0["DRHOT"]) * rdens0)
    aI, aE = faces(a_node)
    alph = np.maximum(aI, aE)                                     # nz^2 = 1 on the vertical faces; the same value seen from both elements

    # ---- (MOMX, MOMY): eval_Ax_uv (:546-553) and, when implicit, their own column systems (construct_matbnd_uv :940-1003)
    t_uv = {}
    for k in ("MOMX", "MOMY"):
        I, E = faces(pv_[k])
        t_uv[k] = -lift(-0.5 * Fs * alph * (E - I)) / Gv
    if impl_fac != 0.0:
        for k in ("MOMX", "MOMY"):
            b = impl_fac * t_uv[k] - pv_[k] + qc[k]
            for e2 in range(Ne2D):
                for ij in range(Nfp):
                    A = np.eye(NeZ * n1)
                    for kz in range(NeZ):
                        for side, (nb, pv1, pvn) in enumerate(((kz - 1, 0, n1 - 1), (kz + 1, n1 - 1, 0))):
                            if nb < 0 or nb >= NeZ:
                                continue
                            t1 = 0.5 * impl_fac / Gv[kz, e2, :, ij] * lw[:, side] * Fs[kz, e2, side, ij] * alph[kz, e2, side, ij]
                            A[kz * n1:(kz + 1) * n1, kz * n1 + pv1] += t1
                            A[kz * n1:(kz + 1) * n1, nb * n1 + pvn] += -t1
                    pv_[k][:, e2, :, ij] += np.linalg.solve(A, b[:, e2, :, ij].reshape(-1)).reshape(NeZ, n1)

    # ---- eval_Ax (:224-262) + vi_cal_del_flux_dyn (:1262-1322) on PROG_VARS (the horizontal momenta already updated)
    mw = pv_["MOMZ"] + Gv * G13 * pv_["MOMX"] + Gv * G23 * pv_["MOMY"]
    dens = pv_["DDENS"] + dh
    rhot = rhot_hyd + pv_["DRHOT"]
    pot = rhot / dens
    dpres_face = pres_of(dens * pot) - ph
    mwI, mwE = faces(mw); mzI, mzE = faces(pv_["MOMZ"]); ddI, ddE = faces(pv_["DDENS"]); drI, drE = faces(pv_["DRHOT"])
    ptI, ptE = faces(pot); dpI, dpE = faces(dpres_face)
    gvI, _ = faces(Gv); g13I, _ = faces(G13); g23I, _ = faces(G23); mxI, _ = faces(pv_["MOMX"]); myI, _ = faces(pv_["MOMY"])
    mzE = np.where(wall, -mzI - 2.0 * gvI * (g13I * mxI + g23I * myI), mzE)       # slip wall (:1292-1302)
    mwE = np.where(wall, -mwI, mwE)
    hf = 0.5 * Fs
    df_d = hf * ((mwE - mwI) * nz - alph * (ddE - ddI))
    df_w = hf * ((dpE - dpI) * nz - alph * (mzE - mzI))
    df_t = hf * ((ptE * mwE - ptI * mwI) * nz - alph * (drE - drI))
    t_d = -(E33 * Dz(mw) + lift(df_d)) / Gv
    t_t = -(E33 * Dz(pot * mw) + lift(df_t)) / Gv
    t_w = -(E33 * Dz(pres_of(rhot) - ph) + lift(df_w)) / Gv - grav * np.einsum("pl,zelj->zepj", VP, pv_["DDENS"])
    out = {}
    if impl_fac == 0.0:
        res = {"DDENS": t_d, "MOMZ": t_w, "DRHOT": t_t, "MOMX": t_uv["MOMX"], "MOMY": t_uv["MOMY"]}
        return {k: v.reshape(-1) for k, v in res.items()}

    # ---- the Newton system of a column as ONE dense matrix: unknown index (kz, pv, var), var = DDENS, MOMZ, DRHOT (construct_matbnd :750-871)
    wt = mw / dens
    dpd = gamm * pres_of(rhot) / rhot
    b3 = {"DDENS": impl_fac * t_d - pv_["DDENS"] + qc["DDENS"], "MOMZ": impl_fac * t_w - pv_["MOMZ"] + qc["MOMZ"],
          "DRHOT": impl_fac * t_t - pv_["DRHOT"] + qc["DRHOT"]}
    nb3 = 3 * n1
    for e2 in range(Ne2D):
        for ij in range(Nfp):
            A = np.zeros((NeZ * nb3, NeZ * nb3))
            rhs = np.zeros(NeZ * nb3)
            ix = lambda kz, pv, v: kz * nb3 + 3 * pv + v
            for kz in range(NeZ):
                gv, po, wT, dp_ = Gv[kz, e2, :, ij], pot[kz, e2, :, ij], wt[kz, e2, :, ij], dpd[kz, e2, :, ij]
                fdz = (E33[kz, e2, :, ij] / gv)[:, None] * (impl_fac * D)                     # [pv, pv2]
                for pv in range(n1):
                    for v, nm in enumerate(("DDENS", "MOMZ", "DRHOT")):
                        rhs[ix(kz, pv, v)] = b3[nm][kz, e2, pv, ij]
                    for p2 in range(n1):
                        idn = 1.0 if pv == p2 else 0.0
                        A[ix(kz, pv, 0), ix(kz, p2, 0)] += idn
                        A[ix(kz, pv, 0), ix(kz, p2, 1)] += fdz[pv, p2]
                        A[ix(kz, pv, 1), ix(kz, p2, 1)] += idn
                        A[ix(kz, pv, 1), ix(kz, p2, 0)] += impl_fac * grav * VP[pv, p2]
                        A[ix(kz, pv, 1), ix(kz, p2, 2)] += fdz[pv, p2] * dp_[p2]
                        A[ix(kz, pv, 2), ix(kz, p2, 0)] += -fdz[pv, p2] * po[p2] * wT[p2]
                        A[ix(kz, pv, 2), ix(kz, p2, 1)] += fdz[pv, p2] * po[p2]
                        A[ix(kz, pv, 2), ix(kz, p2, 2)] += idn + fdz[pv, p2] * wT[p2]
                for side, (nbz, pv1, pvn) in enumerate(((kz - 1, 0, n1 - 1), (kz + 1, n1 - 1, 0))):
                    fac = 0.5 * impl_fac / gv * lw[:, side] * Fs[kz, e2, side, ij]             # [pv]
                    t1 = fac * alph[kz, e2, side, ij]
                    t2 = fac * (-1.0 if side == 0 else 1.0)
                    for pv in range(n1):
                        r0, r1, r2 = ix(kz, pv, 0), ix(kz, pv, 1), ix(kz, pv, 2)
                        c0, c1, c2 = ix(kz, pv1, 0), ix(kz, pv1, 1), ix(kz, pv1, 2)
                        if nbz < 0 or nbz >= NeZ:                                             # slip wall
                            A[r2, c0] += 2.0 * t2[pv] * po[pv1] * wT[pv1]
                            A[r0, c1] += -2.0 * t2[pv]
                            A[r1, c1] += 2.0 * t1[pv]
                            A[r2, c1] += -2.0 * t2[pv] * po[pv1]
                            A[r2, c2] += -2.0 * t2[pv] * wT[pv1]
                        else:
                            A[r0, c0] += t1[pv]
                            A[r2, c0] += t2[pv] * po[pv1] * wT[pv1]
                            A[r0, c1] += -t2[pv]
                            A[r1, c1] += t1[pv]
                            A[r2, c1] += -t2[pv] * po[pv1]
                            A[r1, c2] += -t2[pv] * dp_[pv1]
                            A[r2, c2] += t1[pv] - t2[pv] * wT[pv1]
                            pn, wn, dn = pot[nbz, e2, pvn, ij], wt[nbz, e2, pvn, ij], dpd[nbz, e2, pvn, ij]
                            n0_, n1_, n2_ = ix(nbz, pvn, 0), ix(nbz, pvn, 1), ix(nbz, pvn, 2)
                            A[r0, n0_] += -t1[pv]
                            A[r2, n0_] += -t2[pv] * pn * wn
                            A[r0, n1_] += t2[pv]
                            A[r1, n1_] += -t1[pv]
                            A[r2, n1_] += t2[pv] * pn
                            A[r1, n2_] += t2[pv] * dn
                            A[r2, n2_] += -t1[pv] + t2[pv] * wn
            x = np.linalg.solve(A, rhs).reshape(NeZ, n1, 3)
            pv_["DDENS"][:, e2, :, ij] += x[:, :, 0]
            pv_["MOMZ"][:, e2, :, ij] += x[:, :, 1]
            pv_["DRHOT"][:, e2, :, ij] += x[:, :, 2]
    return {k: ((pv_[k] - qc[k]) / impl_fac).reshape(-1) for k in names}


PROG = ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT")


def apply_bc(elem, mesh, q, bc6):
    """ApplyBC_PROGVARS_lc (bnd.F90:326-362): the halo slots of a physical boundary get the mirrored (SLIP = 2) or negated (NOSLIP = 3)
    momentum of the interior face node; regional mesh (G11 = G22 = 1, G12 = 0), terrain-following metric included.  bc6: boundary id per
    tile face (mesh.halo_bc_types); q: dict of flat arrays incl. halo, modified in place."""
    iM, iP = mesh.VMapM.reshape(-1), mesh.VMapP.reshape(-1)
    nx, ny, nz = (a.reshape(-1) for a in mesh.normal_fn)
    nint = mesh.Ne * elem.Np
    slot = iP - nint
    on = slot >= 0
    face_of_slot = np.searchsorted(np.cumsum(mesh.halo_face_size), np.maximum(slot, 0), side="right")
    bc = np.where(on, np.asarray(bc6)[np.minimum(face_of_slot, 5)], 0)
    G = mesh.Gsqrt.reshape(-1)
    G13, G23 = mesh.GI3[0].reshape(-1), mesh.GI3[1].reshape(-1)
    ke2d = np.repeat(mesh.EMap3Dto2D, elem.NfpTot)
    h2 = (iM % elem.Np) % elem.Nfp
    sl = bc == 2
    m_, p_ = iM[sl], iP[sl]
    Gv = G[m_] / mesh.GsqrtH[ke2d[sl], h2[sl]]
    momw = q["MOMZ"][m_] / Gv + G13[m_] * q["MOMX"][m_] + G23[m_] * q["MOMY"][m_]
    fac = nz[sl] * Gv ** 2 / (1.0 + (Gv * G13[m_]) ** 2 + (Gv * G23[m_]) ** 2)
    mn = q["MOMX"][m_] * nx[sl] + q["MOMY"][m_] * ny[sl] + momw * nz[sl]
    q["MOMX"][p_] = q["MOMX"][m_] - 2.0 * mn * (nx[sl] + fac * G13[m_])
    q["MOMY"][p_] = q["MOMY"][m_] - 2.0 * mn * (ny[sl] + fac * G23[m_])
    q["MOMZ"][p_] = q["MOMZ"][m_] - 2.0 * mn * fac / Gv
    ns = bc == 3
    for k in ("MOMX", "MOMY", "MOMZ"):
        q[k][iP[ns]] = -q[k][iM[ns]]


def modal_filter(elem, mesh, q, Fh, Fv):
    """q <- F3D(Gsqrt q) / Gsqrt with the 1D filter matrices Fh (x, y) and Fv (z) (modalfilter.F90:49-130), interior elements."""
    Ne, n = mesh.Ne, elem.np1
    ni = Ne * elem.Np
    G = mesh.Gsqrt.reshape(-1)[:ni].reshape(Ne, n, n, n)
    for k in PROG:
        a = q[k][:ni].reshape(Ne, n, n, n) * G                     # [ke, k, j, i]
        a = np.einsum("il,ekjl->ekji", Fh, a)
        a = np.einsum("jl,ekli->ekji", Fh, a)
        a = np.einsum("kl,elji->ekji", Fv, a)
        q[k][:ni] = (a / G).reshape(-1)


def update(elem, mesh, c, q, aux, rk, dt, bc6, hevi=False, filt=None, nsteps=1, DPhydDx=None, DPhydDy=None):
    """nsteps dynamics steps of the regional dry model (driver_nonhydro3d.F90:703-951), q: dict of the five prognostic variables, flat
    arrays incl. halo, advanced in place.  rk: Butcher tables (oracle_api.rk_tables); filt: (Fh, Fv) or None; DPhydDx / DPhydDy: the
    horizontal gradient of the background pressure (set-up product; non-zero over topography).  Per stage: [HEVI: cal_vi about
    the state at the start of the step, StoreImplicit], halo exchange, pressure, boundary condition, explicit tendency, then the stage
    combination  q = q0 + dt sum_j (a_ex(s+1, j) k_ex_j + a_im(s+1, j) k_im_j)  (the weights b for the last stage)."""
    ns = rk["nstage"]
    ni = mesh.Ne * elem.Np
    R, cv, cp = (np.full(mesh.NeA * elem.Np, c[k]) for k in ("Rdry", "CVdry", "CPdry"))
    for _ in range(nsteps):
        q0 = {k: q[k][:ni].copy() for k in PROG}
        kex, kim = [], []
        for st in range(ns):
            if hevi:
                t = cal_vi(elem, mesh, c, aux, q, q0, rk["a_im"][st, st] * dt)
                kim.append(t)
                for k in PROG:
                    q[k][:ni] = q[k][:ni] + rk["a_im"][st, st] * dt * t[k]
            for k in PROG:
                mesh.exchange_halo_numpy(q[k])
            _, dpres = drhot2pres(c, q["DRHOT"], aux["PRES_hyd"], aux["THERM_hyd"], R, cv, cp)
            mesh.exchange_halo_numpy(dpres)
            apply_bc(elem, mesh, q, bc6)
            t = (cal_tend_hevi if hevi else cal_tend_heve)(elem, mesh, c, q, aux, dpres, DPhydDx, DPhydDy)
            kex.append({"DDENS": t["DENS_dt"].reshape(-1), "MOMX": t["MOMX_dt"].reshape(-1), "MOMY": t["MOMY_dt"].reshape(-1),
                        "MOMZ": t["MOMZ_dt"].reshape(-1), "DRHOT": t["RHOT_dt"].reshape(-1)})
            last = st == ns - 1
            for k in PROG:
                acc = q0[k].copy()
                for j in range(st + 1):
                    acc = acc + dt * (rk["b_ex"][j] if last else rk["a_ex"][st + 1, j]) * kex[j][k]
                    if hevi:
                        acc = acc + dt * (rk["b_im"][j] if last else rk["a_im"][st + 1, j]) * kim[j][k]
                q[k][:ni] = acc
        if filt is not None:
            modal_filter(elem, mesh, q, *filt)


def update_sphere(elem, cs, c, qs, auxs, rk, dt, dphyd, filt=None, nsteps=1):
    """nsteps steps of GLOBALNONHYDRO3D_HEVI on the whole cubed sphere (the six local meshes of one process, driver_nonhydro3d.F90:703-951 with
    LOCAL_MESH_NUM = 6): per stage the vertical-implicit Newton step of every panel, the panel-edge exchange (cs.exchange_numpy: index reversal
    and the change of basis of (MOMX, MOMY) across the edges) + the own top / bottom faces, pressure and its exchange, slip walls at the
    bottom and the top, the explicit tendency of the global equations, the stage combination; then the modal filter.
    qs / auxs: one dict per panel of flat arrays incl. halo; dphyd: per panel (DPhydDx, DPhydDy)."""
    ns = rk["nstage"]
    Np = elem.Np
    nis = [m.Ne * Np for m in cs.panels]
    bc6s = [m.halo_bc_types({"btm": 2, "top": 2}) for m in cs.panels]
    for _ in range(nsteps):
        q0 = [{k: q[k][:ni].copy() for k in PROG} for q, ni in zip(qs, nis)]
        kex, kim = [[] for _ in qs], [[] for _ in qs]
        for st in range(ns):
            for P, (q, m, ni) in enumerate(zip(qs, cs.panels, nis)):
                t = cal_vi(elem, m, c, auxs[P], q, q0[P], rk["a_im"][st, st] * dt)
                kim[P].append(t)
                for k in PROG:
                    q[k][:ni] = q[k][:ni] + rk["a_im"][st, st] * dt * t[k]
            dps = []
            for P, (q, m) in enumerate(zip(qs, cs.panels)):
                R, cv, cp = (np.full(m.NeA * Np, c[k]) for k in ("Rdry", "CVdry", "CPdry"))
                _, dpres = drhot2pres(c, q["DRHOT"], auxs[P]["PRES_hyd"], auxs[P]["THERM_hyd"], R, cv, cp)
                for k in PROG:
                    m.exchange_halo_numpy(q[k])            # top / bottom faces (the lateral slots are overwritten by the panel-edge exchange)
                m.exchange_halo_numpy(dpres)
                dps.append(dpres)
            cs.exchange_numpy([dict(q, DPRES=dp) for q, dp in zip(qs, dps)])
            for P, (q, m, ni) in enumerate(zip(qs, cs.panels, nis)):
                apply_bc(elem, m, q, bc6s[P])
                t = cal_tend_hevi_global(elem, m, c, q, auxs[P], dps[P], dphyd[P][0], dphyd[P][1])
                kex[P].append({"DDENS": t["DENS_dt"].reshape(-1), "MOMX": t["MOMX_dt"].reshape(-1), "MOMY": t["MOMY_dt"].reshape(-1),
                               "MOMZ": t["MOMZ_dt"].reshape(-1), "DRHOT": t["RHOT_dt"].reshape(-1)})
            last = st == ns - 1
            for P, (q, ni) in enumerate(zip(qs, nis)):
                for k in PROG:
                    acc = q0[P][k].copy()
                    for j in range(st + 1):
                        acc = acc + dt * (rk["b_ex"][j] if last else rk["a_ex"][st + 1, j]) * kex[P][j][k]
                        acc = acc + dt * (rk["b_im"][j] if last else rk["a_im"][st + 1, j]) * kim[P][j][k]
                    q[k][:ni] = acc
        if filt is not None:
            for q, m in zip(qs, cs.panels):
                modal_filter(elem, m, q, *filt)
