"""Builders of the synthetic BASELINE configurations (mesh + initial state + driver set-up) used by bench.py, smoke() and the
parity tests.  The CPU checker of a case is built by the test infrastructure (tests/cases.py), not here."""
from __future__ import annotations

import numpy as np

from . import initcond
from .element import HexElement
from .mesh import LocalMeshCube, LocalMeshCubedSpherePanel

C0 = initcond.SCALE_CONST
SLIP6 = dict(south="SLIP", east="SLIP", north="SLIP", west="SLIP", btm="SLIP", top="SLIP")
MF = dict(MF_ETAC_h=2.0 / 3.0, MF_ALPHA_h=1.0, MF_ORDER_h=16, MF_ETAC_v=2.0 / 3.0, MF_ALPHA_v=1.0, MF_ORDER_v=16)


class DensityCurrentCase:
    """Straka density current on NeX x NeY x NeZ elements of order p (config 3 of BASELINE.json, any size)."""

    def __init__(self, p=7, NeX=8, NeY=2, NeZ=4, dom=(0.0, 25.6e3, 0.0, 6.4e3, 0.0, 6.4e3), dt=0.08,
                 tinteg="ERK_SSP_4s3o", modalfilter=True, perturb=0.0, periodic=(False, True, False), intrp_order=11,
                 eqs="NONHYDRO3D_HEVE", NprcX=1, NprcY=1, pi=0, pj=0, mf=None):
        """NeX, NeY, NeZ are per tile; dom is the whole domain; (pi, pj) selects the tile of an NprcX x NprcY decomposition."""
        self.p, self.dom, self.dt, self.tinteg, self.modalfilter = p, dom, dt, tinteg, modalfilter
        self.mf = dict(MF, **(mf or {}))
        self.eqs = eqs
        self.periodic = periodic
        self.NprcX, self.NprcY, self.pi, self.pj = NprcX, NprcY, pi, pj
        self.elem = HexElement(p)
        self.mesh = LocalMeshCube(self.elem, NeX, NeY, NeZ, *dom, periodic=periodic, NprcX=NprcX, NprcY=NprcY, pi=pi, pj=pj)
        self.fields = initcond.density_current(self.mesh, intrp_order=intrp_order)
        if perturb:
            # deterministic smooth 3D momentum perturbation so that every term of the tendency is exercised
            x, y, z = (self.mesh.pos_en[d] for d in range(3))
            Ne = self.mesh.Ne
            kx, ky, kz = 2 * np.pi / (dom[1] - dom[0]), 2 * np.pi / (dom[3] - dom[2]), np.pi / (dom[5] - dom[4])
            self.fields["MOMX"][:Ne] = perturb * np.sin(kx * x) * np.cos(ky * y) * np.cos(kz * z)
            self.fields["MOMY"][:Ne] = perturb * 0.7 * np.sin(kx * x + 0.3) * np.sin(ky * y + 0.1) * np.cos(kz * z)
            self.fields["MOMZ"][:Ne] = perturb * 0.5 * np.sin(kx * x) * np.cos(ky * y) * np.sin(kz * z)
        self.vel_bc = SLIP6
        self.consts = C0

    def make_driver(self, hgrad_src=None):
        """hgrad_src: anything with .arr("DPhydDx") / .arr("DPhydDy") (set-up products of the model, driver_nonhydro3d.F90:1060-1095)."""
        from .dyncore import AtmDynDGMDriver_nonhydro3d
        rank = self.pi + self.pj * self.NprcX
        d = AtmDynDGMDriver_nonhydro3d(self.elem, self.mesh, self.consts, vel_bc=self.vel_bc, my_rank=rank,
                                       tile_rank=lambda qi, qj: qi + qj * self.NprcX)
        d.Init(self.eqs, self.tinteg, self.dt, MODALFILTER_FLAG=self.modalfilter, **self.mf)
        f = self.fields
        d.set_aux(f["DENS_hyd"], f["PRES_hyd"])
        if hgrad_src is not None:
            d.set_phyd_hgrad(hgrad_src.arr("DPhydDx"), hgrad_src.arr("DPhydDy"))
        d.set_prog(*(f[k] for k in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT")))
        return d


class SoundWaveCase(DensityCurrentCase):
    """Vertical sound-wave pulse of sample/euler3d_hevi (config 2 of BASELINE.json): 10 km cube, uniform background
    DENS_hyd = 1, PRES_hyd = 1e5, GRAV = 0 (PARAM_CONST of its test.conf), DRHOT = A cos(pi r / 2) for |r| <= 1 around
    mid-height (test_euler3d_hevi.f90:1106-1149), run through the library HEVI path (rows a8-a12) with IMEX_ARK232.
    The shipped amplitude 1e-12 sits at round-off of the background rho*theta, so parity runs
    use a raised amplitude; horizontally periodic, slip walls at bottom and top."""

    def __init__(self, p=7, NeX=1, NeY=1, NeZ=80, dt=10.0, tinteg="IMEX_ARK232", amplitude=1.0e-3, modalfilter=False,
                 NprcX=1, NprcY=1, pi=0, pj=0):
        self.p, self.dt, self.tinteg, self.modalfilter = p, dt, tinteg, modalfilter
        self.mf = dict(MF)
        self.dom = dom = (0.0, 10.0e3, 0.0, 10.0e3, 0.0, 10.0e3)
        self.eqs = "NONHYDRO3D_HEVI"
        self.periodic = (True, True, False)
        self.NprcX, self.NprcY, self.pi, self.pj = NprcX, NprcY, pi, pj
        self.elem = HexElement(p)
        self.mesh = LocalMeshCube(self.elem, NeX, NeY, NeZ, *dom, periodic=self.periodic, NprcX=NprcX, NprcY=NprcY, pi=pi, pj=pj)
        self.consts = dict(C0, GRAV=0.0)
        self.fields = initcond.sound_wave(self.mesh, amplitude=amplitude)
        self.vel_bc = SLIP6


class GlobalPanelCase(DensityCurrentCase):
    """One cubed-sphere panel (GLOBALNONHYDRO3D_HEVI, shallow atmosphere, no topography; BASELINE config 4 in small):
    isothermal hydrostatic background, solid-body zonal flow u = u0 cos(lat) in gradient-wind balance on an equatorial
    panel plus a smooth 3D perturbation, so that every metric, Coriolis and pressure-gradient term is exercised.  The
    lateral halo of the tile holds its own face values (no panel-edge exchange in this scope)."""

    def __init__(self, p=7, panelID=1, NeX=2, NeY=2, NeZ=3, ztop=30.0e3, dt=20.0, tinteg="IMEX_ARK324", modalfilter=True,
                 u0=30.0, T0=300.0, perturb=1.0, OHM=None, balanced=True, eqs="GLOBALNONHYDRO3D_HEVI"):
        self.p, self.dt, self.tinteg, self.modalfilter = p, dt, tinteg, modalfilter
        self.mf = dict(MF)
        self.eqs = eqs
        self.periodic = (False, False, False)
        self.NprcX = self.NprcY = 1
        self.pi = self.pj = 0
        self.panelID, self.ztop = panelID, ztop
        self.elem = HexElement(p)
        self.consts = dict(C0) if OHM is None else dict(C0, OHM=OHM)
        c = self.consts
        self.mesh = m = LocalMeshCubedSpherePanel(self.elem, panelID, NeX, NeY, NeZ, ztop, c["RPlanet"])
        Ne, Np, NeA = m.Ne, self.elem.Np, m.NeA
        z = m.pos_en[2]
        h2 = self.elem.IndexH2Dto3D
        lat = m.lat2D[m.EMap3Dto2D][:, h2]
        # background: isothermal, function of z only
        H = c["Rdry"] * T0 / c["GRAV"]
        pres_hyd = c["PRES00"] * np.exp(-z / H)
        dens_hyd = pres_hyd / (c["Rdry"] * T0)
        # balanced state: ln p = ln p_hyd(z) - (u0^2 + 2 a Omega u0) sin^2(lat) / (2 R T0), T = T0
        amp = (u0 ** 2 + 2.0 * c["RPlanet"] * c["OHM"] * u0) / (2.0 * c["Rdry"] * T0) if balanced else 0.0
        pres = pres_hyd * np.exp(-amp * np.sin(lat) ** 2)
        dens = pres / (c["Rdry"] * T0)
        theta = T0 * (c["PRES00"] / pres) ** (c["Rdry"] / c["CPdry"])
        rhot_hyd = c["PRES00"] / c["Rdry"] * (pres_hyd / c["PRES00"]) ** (c["CVdry"] / c["CPdry"])
        ua, ub = m.lonlat_to_cs_vec(u0 * np.cos(m.lat2D), np.zeros_like(m.lat2D))
        f = {k: np.zeros((NeA, Np)) for k in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT", "DENS_hyd", "PRES_hyd")}
        f["DENS_hyd"][:Ne] = dens_hyd
        f["PRES_hyd"][:Ne] = pres_hyd
        f["DDENS"][:Ne] = dens - dens_hyd
        f["DRHOT"][:Ne] = dens * theta - rhot_hyd
        f["MOMX"][:Ne] = dens * ua[m.EMap3Dto2D][:, h2]
        f["MOMY"][:Ne] = dens * ub[m.EMap3Dto2D][:, h2]
        if perturb:
            a, b = m.pos_en[0], m.pos_en[1]
            R = c["RPlanet"]
            f["MOMX"][:Ne] += perturb * dens * (3.0 / R) * np.sin(3 * a) * np.cos(2 * b) * np.cos(np.pi * z / ztop)
            f["MOMY"][:Ne] += perturb * dens * (2.0 / R) * np.cos(2 * a + 0.3) * np.sin(3 * b + 0.1) * np.cos(np.pi * z / ztop)
            f["MOMZ"][:Ne] += perturb * dens * 0.05 * np.sin(2 * a) * np.cos(3 * b) * np.sin(np.pi * z / ztop)
            f["DRHOT"][:Ne] += perturb * 0.5 * dens * np.cos(3 * a) * np.cos(2 * b) * np.sin(2 * np.pi * z / ztop)
        self.fields = f
        self.vel_bc = dict(btm="SLIP", top="SLIP")

class GlobalSphereCase:
    """The whole cubed sphere, six panel tiles of Ne x Ne x NeZ elements (BASELINE config 4).
    init = "solid_body": isothermal atmosphere in solid-body rotation (gradient-wind balance) plus, when perturb != 0, a second
    solid-body rotation about a tilted axis (smooth across every panel edge and both poles), a vertical-velocity pattern and a warm blob.
    init = "jw": the Jablonowski-Williamson baroclinic wave of test/case/baroclinic_wave_global (initcond.baroclinic_wave_global;
    no topography); `GlobalSphereCase.config4(...)` is the shipped run.conf: lumped mass matrix, stretched FZ, modal filter with
    eta_c = 0, sponge layer above 20 km, IMEX_ARK324 with dt = 75 s at NeGX = 8 (scaled with the element size)."""

    FZ_SHIPPED = (0.0, 3000.0, 8000.0, 15000.0, 30000.0)      # run.conf:41 (NeZ = 4)

    @classmethod
    def config4(cls, Ne=8, NeZ=4, ntile=1, fields_for=None, p=7, init="jw"):
        """BASELINE configs[3] at NeGX = NeGY = Ne * ntile, NeZ levels: every shipped layer is cut into NeZ / 4 equal parts.
        init = "solid_body": the same run.conf on the cheap analytic state (configs[4] sizes: the Jablonowski-Williamson set-up costs a
        Newton iteration per quadrature point, minutes of host time at 2e8 nodes)."""
        assert NeZ % 4 == 0
        sub = NeZ // 4
        fz0 = np.array(cls.FZ_SHIPPED)
        FZ = np.concatenate([fz0[:1]] + [fz0[i] + (fz0[i + 1] - fz0[i]) * np.arange(1, sub + 1) / sub for i in range(4)])
        return cls(p=p, Ne=Ne, NeZ=NeZ, ztop=30.0e3, dt=75.0 * 8.0 / (Ne * ntile), tinteg="IMEX_ARK324", modalfilter=True, ntile=ntile,
                   fields_for=fields_for, init=init, lumped=True, FZ=FZ,
                   mf=dict(MF_ETAC_h=0.0, MF_ALPHA_h=1.0, MF_ORDER_h=16, MF_ETAC_v=0.0, MF_ALPHA_v=1.0, MF_ORDER_v=16),
                   sponge=dict(SL_WDAMP_TAU=86400.0, SL_WDAMP_HEIGHT=20.0e3))

    def __init__(self, p=7, Ne=2, NeZ=2, ztop=30.0e3, dt=20.0, tinteg="IMEX_ARK324", modalfilter=True, u0=30.0, T0=300.0, perturb=1.0,
                 eqs="GLOBALNONHYDRO3D_HEVI", ntile=1, fields_for=None, init="solid_body", lumped=False, FZ=None, mf=None, sponge=None):
        """ntile = k: k x k tiles per panel, Ne elements per TILE edge (24 local meshes for k = 2); the CPU checker of the tests steps
        whole panels, so a test builds a second case with Ne * k and ntile = 1 for it.  fields_for: the local meshes whose initial state is
        evaluated (default all; a rank of a multi-GPU run passes the ones it owns)."""
        from .cubedsphere import CubedSphere, cs2cart, cs2lonlat, lonlat2cs_vec
        self.p, self.dt, self.tinteg, self.modalfilter, self.ztop = p, dt, tinteg, modalfilter, ztop
        self.eqs = eqs
        self.init = init
        self.mf = dict(MF, **(mf or {}))
        self.sponge = sponge
        self.FZ = None if FZ is None else np.asarray(FZ, dtype=np.float64)
        self.elem = HexElement(p, lumped=lumped)
        self.consts = c = dict(C0)
        self.cs = CubedSphere(self.elem, Ne, NeZ, ztop, c["RPlanet"], FZ=self.FZ, ntile=ntile, build=fields_for)
        self.vel_bc = dict(btm="SLIP", top="SLIP")
        if init == "jw":
            self.fields = [initcond.baroclinic_wave_global(m, c=c) if (fields_for is None or t in fields_for) else None
                           for t, m in enumerate(self.cs.panels)]
            return
        H = c["Rdry"] * T0 / c["GRAV"]
        amp = (u0 ** 2 + 2.0 * c["RPlanet"] * c["OHM"] * u0) / (2.0 * c["Rdry"] * T0)
        w2 = perturb * 8.0 / c["RPlanet"] * np.array([0.6, -0.3, 0.74])          # tilted rotation vector [1/s]
        def one_tile(tm):
            t, m = tm
            if fields_for is not None and t not in fields_for:
                return None
            P = m.panelID - 1
            Np, NeA, Nel = self.elem.Np, m.NeA, m.Ne
            a, b, z = m.pos_en[0], m.pos_en[1], m.pos_en[2]
            lon, lat = cs2lonlat(P + 1, a, b)
            pres_hyd = c["PRES00"] * np.exp(-z / H)
            dens_hyd = pres_hyd / (c["Rdry"] * T0)
            pres = pres_hyd * np.exp(-amp * np.sin(lat) ** 2)
            x = cs2cart(P + 1, a, b)                                               # unit sphere
            blob = perturb * 2.0 * np.exp(-((x[0] - 0.5) ** 2 + (x[1] - 0.6) ** 2 + (x[2] - 0.62) ** 2) / 0.3 ** 2) * np.sin(np.pi * z / ztop)
            temp = T0 + blob
            dens = pres / (c["Rdry"] * temp)
            theta = temp * (c["PRES00"] / pres) ** (c["Rdry"] / c["CPdry"])
            rhot_hyd = c["PRES00"] / c["Rdry"] * (pres_hyd / c["PRES00"]) ** (c["CVdry"] / c["CPdry"])
            # wind: u0 cos(lat) zonal + w2 x r
            v3 = np.stack([w2[1] * x[2] - w2[2] * x[1], w2[2] * x[0] - w2[0] * x[2], w2[0] * x[1] - w2[1] * x[0]]) * c["RPlanet"]
            elon = np.stack([-np.sin(lon), np.cos(lon), np.zeros_like(lon)])
            elat = np.stack([-np.sin(lat) * np.cos(lon), -np.sin(lat) * np.sin(lon), np.cos(lat)])
            ul = u0 * np.cos(lat) + np.sum(v3 * elon, axis=0)
            vl = np.sum(v3 * elat, axis=0)
            ua, ub = lonlat2cs_vec(P + 1, a, b, ul, vl, c["RPlanet"])
            f = {k: np.zeros((NeA, Np)) for k in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT", "DENS_hyd", "PRES_hyd")}
            f["DENS_hyd"][:Nel] = dens_hyd; f["PRES_hyd"][:Nel] = pres_hyd
            f["DDENS"][:Nel] = dens - dens_hyd
            f["DRHOT"][:Nel] = dens * theta - rhot_hyd
            f["MOMX"][:Nel] = dens * ua; f["MOMY"][:Nel] = dens * ub
            f["MOMZ"][:Nel] = perturb * dens * 0.05 * x[0] * x[1] * np.sin(np.pi * z / ztop)
            return f

        # the tiles are independent and NumPy releases the GIL in its loops: one thread per own tile
        from concurrent.futures import ThreadPoolExecutor
        import os
        with ThreadPoolExecutor(max_workers=max(1, int(os.environ.get("FEDG_INIT_THREADS", "4")))) as ex_:   # ~250 B per node of temporaries per tile in flight
            self.fields = list(ex_.map(one_tile, enumerate(self.cs.panels)))

    def make_driver(self, rank=0, nranks=1, bcast=None):
        """rank / nranks > 1: only the panels this rank owns get a device context (`g.panel_ids`)."""
        from .cubedsphere import GlobalSphereDriver
        g = GlobalSphereDriver(self.cs, self.consts, vel_bc=self.vel_bc, rank=rank, nranks=nranks, bcast=bcast)
        g.Init(self.eqs, self.tinteg, self.dt, MODALFILTER_FLAG=self.modalfilter, **self.mf)
        for d, f in zip(g.panels, [self.fields[P] for P in g.panel_ids]):
            d.set_aux(f["DENS_hyd"], f["PRES_hyd"])
            if self.sponge:
                d.sponge_init(**self.sponge)
            d.set_prog(*(f[k] for k in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT")))
        g.exchange_aux()
        if self.init == "jw":          # PRES_hyd varies horizontally: update_phyd_hgrad after the background fields were exchanged
            for d in g.panels:
                d.update_phyd_hgrad()
        return g


def rel_l2(a, b):
    a = np.asarray(a).reshape(-1); b = np.asarray(b).reshape(-1)
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)
