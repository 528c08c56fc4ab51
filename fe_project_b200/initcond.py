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


# ---- Jablonowski-Williamson (2006) baroclinic wave on a cubed-sphere panel tile -------------------------------------
JW_DEFAULT = dict(REF_TEMP=288.0, REF_PRES=1.0e5, LAPSE_RATE=5.0e-3, U0=35.0, Up=1.0, ETA0=0.252, ETAt=0.2, DELTAT=4.8e5,
                  Lon_c=np.pi / 9.0, Lat_c=2.0 / 9.0 * np.pi)


def jw_balanced_point(lat, z, c=SCALE_CONST, prm=JW_DEFAULT):
    """get_thermal_wind_balance_1point_itr + _geopot_hvari (model/atm_nonhydro3d/test/case/baroclinic_wave_global/mod_user.F90:405-515),
    vectorised: Newton iteration for eta at every point with the reference's per-point stopping rule (|del_eta| <= 5e-15;
    pres = eta of the LAST EVALUATION times REF_PRES, temp of the last evaluation).  Returns pres, temp, vel_lon."""
    lat = np.asarray(lat, dtype=np.float64); z = np.asarray(z, dtype=np.float64)
    G, Rd, R, OHM = c["GRAV"], c["Rdry"], c["RPlanet"], c["OHM"]
    U0, T0, LR, ETA0, ETAt, DT = prm["U0"], prm["REF_TEMP"], prm["LAPSE_RATE"], prm["ETA0"], prm["ETAt"], prm["DELTAT"]
    sl, cl = np.sin(lat), np.cos(lat)
    h1 = U0 * (-2.0 * sl ** 6 * (cl ** 2 + 1.0 / 3.0) + 10.0 / 63.0)
    h2 = R * OHM * (8.0 / 5.0 * cl ** 3 * (sl ** 2 + 2.0 / 3.0) - 0.25 * np.pi)
    shape = np.broadcast(lat, z).shape
    h1 = np.broadcast_to(h1, shape).reshape(-1); h2 = np.broadcast_to(h2, shape).reshape(-1)
    zf = np.broadcast_to(z, shape).reshape(-1)
    n = zf.size
    eta = np.full(n, 1.0e-8)
    eta_save = eta.copy(); temp = np.zeros(n); c32 = np.zeros(n)
    act = np.arange(n)
    for itr in range(1001):
        if act.size == 0:
            break
        e = eta[act]
        etav = 0.5 * np.pi * (e - ETA0)
        ce = np.cos(etav)
        c12 = np.sqrt(ce)
        c3 = c12 ** 3
        t = T0 * e ** (Rd * LR / G)
        gp = G / LR * (T0 - t)
        strat = ETAt > e
        es = e[strat]
        t[strat] = t[strat] + DT * (ETAt - es) ** 5
        gp[strat] = gp[strat] - Rd * DT * ((np.log(es / ETAt) + 137.0 / 60.0) * ETAt ** 5 - 5.0 * ETAt ** 4 * es
                                           + 5.0 * ETAt ** 3 * es ** 2 - 10.0 / 3.0 * ETAt ** 2 * es ** 3
                                           + 1.25 * ETAt * es ** 4 - 0.2 * es ** 5)
        t = t + 0.75 * e * np.pi * U0 / Rd * np.sin(etav) * c12 * (h1[act] * 2.0 * c3 + h2[act])
        gp = gp + U0 * c3 * (h1[act] * c3 + h2[act])
        de = -(-G * zf[act] + gp) * (-e / (Rd * t))
        eta_save[act] = e; temp[act] = t; c32[act] = c3
        eta[act] = e + de
        act = act[np.abs(de) > 5e-15]
    else:
        raise RuntimeError("JW balance iteration did not converge")
    pres = (eta_save * prm["REF_PRES"]).reshape(shape)
    vel_lon = (U0 * c32).reshape(shape) * np.broadcast_to(np.sin(2.0 * lat) ** 2, shape)
    return pres, temp.reshape(shape), vel_lon


def baroclinic_wave_global(mesh, intrp_order=8, c=SCALE_CONST, prm=JW_DEFAULT, ch=256):
    """exp_SetInitCond_baroclinicwave with skip_topo = .true. (mod_user.F90:166-262, 264-402): balanced pressure, temperature and zonal
    wind + the Gaussian zonal-wind perturbation, sampled on an order-`intrp_order` LGL element and L2-projected; DDENS = DRHOT =
    MOMZ = 0, DENS_hyd / PRES_hyd carry the balanced state.  mesh: LocalMeshCubedSpherePanel (any tile).  Dict of (NeA, Np) arrays."""
    from .cubedsphere import cs2lonlat, lonlat2cs_vec
    e = mesh.elem
    T1, src = e.l2proj_from(intrp_order)             # (np1, nq)
    nq = src.Np
    h = 0.5 * (src.x + 1.0)
    NeX, NeY = mesh.NeX, mesh.NeY
    vx = (mesh.xmax - mesh.xmin) * np.arange(NeX + 1) / NeX + mesh.xmin
    vy = (mesh.ymax - mesh.ymin) * np.arange(NeY + 1) / NeY + mesh.ymin
    vz = mesh.FZ
    R = c["RPlanet"]
    Lp = R / 10.0
    Np, Ne, NeA = e.Np, mesh.Ne, mesh.NeA
    f = {k: np.zeros((NeA, Np)) for k in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT", "DENS_hyd", "PRES_hyd")}

    def proj(q):                                     # q [B, kq, jq, iq] -> [B, Np]
        q = q @ T1.T
        q = np.einsum("jb,Bcbi->Bcji", T1, q, optimize=True)
        q = np.einsum("kc,Bcji->Bkji", T1, q, optimize=True)
        return q.reshape(q.shape[0], -1)

    def chunk(k0):
        sl = slice(k0, min(k0 + ch, Ne))
        ex, ey, ez = mesh.ex[sl], mesh.ey[sl], mesh.ez[sl]
        B = ex.size
        a1 = vx[ex][:, None] + h[None, :] * (vx[ex + 1] - vx[ex])[:, None]
        b1 = vy[ey][:, None] + h[None, :] * (vy[ey + 1] - vy[ey])[:, None]
        z1 = vz[ez][:, None] + h[None, :] * (vz[ez + 1] - vz[ez])[:, None]
        a = np.broadcast_to(a1[:, None, None, :], (B, nq, nq, nq))
        b = np.broadcast_to(b1[:, None, :, None], (B, nq, nq, nq))
        z = np.broadcast_to(z1[:, :, None, None], (B, nq, nq, nq))
        lon, lat = cs2lonlat(mesh.panelID, a, b)
        pres, temp, ulon = jw_balanced_point(lat, z, c, prm)
        # "Replace VelLon with zero" at the poles, bottom layer only (mod_user.F90:376-381)
        ulon = np.where((ez == 0)[:, None, None, None] & (np.cos(lat) < c["EPS"]), 0.0, ulon)
        r = R / Lp * np.arccos(np.clip(np.sin(prm["Lat_c"]) * np.sin(lat) + np.cos(prm["Lat_c"]) * np.cos(lat) * np.cos(lon - prm["Lon_c"]), -1.0, 1.0))
        udash = prm["Up"] * np.exp(-r ** 2)
        zero = np.zeros_like(ulon)
        U, V = lonlat2cs_vec(mesh.panelID, a, b, ulon, zero, R)
        Ud, Vd = lonlat2cs_vec(mesh.panelID, a, b, udash, zero, R)
        dens = pres / (c["Rdry"] * temp)
        f["PRES_hyd"][sl] = proj(pres)
        f["DENS_hyd"][sl] = proj(dens)
        f["MOMX"][sl] = proj(dens * (U + Ud))
        f["MOMY"][sl] = proj(dens * (V + Vd))

    # the chunks are independent and NumPy releases the GIL inside its loops: a thread per core (set-up time of the 6x32x32x12 case)
    import os
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, min(32, os.cpu_count() or 1))) as ex_:
        list(ex_.map(chunk, range(0, Ne, ch)))
    return f
