"""Host-side mirror of the reference's dynamics-driver interface over the C ABI.

`AtmDynDGMDriver_nonhydro3d` keeps the names and argument meaning of the reference type of the same
name (FElib/src/fluid_dyn_solver/scale_atm_dyn_dgm_driver_nonhydro3d.F90:64-150: `Init`, `Update`,
`Final`; namelist PARAM_ATMOS_DYN of model/atm_nonhydro3d/src/atmos/mod_atmos_dyn.F90:121-147:
EQS_TYPE, TINTEG_TYPE, TIME_DT, MODALFILTER_FLAG).  All arithmetic happens in libfedg.so on the GPU;
this module only marshals arrays (NumPy, node index fastest == the reference's column-major layout).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .element import HexElement
from .mesh import LocalMeshCube, BND_NOSPEC, BND_SLIP, BND_NOSLIP, BND_PERIODIC

PROG_NAMES = ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT")
_BC_NAMES = {"NOSPEC": BND_NOSPEC, "PERIODIC": BND_PERIODIC, "SLIP": BND_SLIP, "NOSLIP": BND_NOSLIP}


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class AtmDynDGMDriver_nonhydro3d:
    """One local mesh (tile) of the nonhydrostatic DG dynamical core on one GPU."""

    def __init__(self, elem: HexElement, mesh: LocalMeshCube, consts: dict, vel_bc: dict | None = None,
                 my_rank: int = 0, tile_rank=None):
        """vel_bc: {'south','east','north','west','btm','top'} -> 'SLIP' | 'NOSLIP' | id
        (PARAM_ATMOS_DYN_BND, scale_atm_dyn_dgm_bnd.F90:128-131).
        tile_rank: callable (pi, pj) -> rank owning that tile (default: everything on my_rank)."""
        self.L = _lib.load()
        self.elem, self.mesh = elem, mesh
        bc = {k: (_BC_NAMES[v.upper()] if isinstance(v, str) else int(v)) for k, v in (vel_bc or {}).items()}
        bc6 = mesh.halo_bc_types(bc)
        d = _lib.MeshDesc()
        d.polyorder = elem.order
        d.Ne, d.NeA, d.NeX, d.NeY, d.NeZ, d.Ne2D, d.Nhalo = (mesh.Ne, mesh.NeA, mesh.NeX, mesh.NeY, mesh.NeZ,
                                                               mesh.Ne2D, mesh.Nhalo)
        keep = self._keep = {}
        keep["D1D"] = _f64(elem.D1D.T)                      # column-major (i,l)
        keep["Lift"] = _f64(elem.lift_dense().T)            # column-major (Np,NfpTot)
        keep["VPOrdM1"] = _f64(elem.VPOrdM1.T)
        keep["IntWeight_lgl"] = _f64(elem.IntWeight_lgl)
        keep["Escale"] = _f64(mesh.Escale.transpose(1, 0, 2, 3))   # Fortran (Np,Ne,3,3): memory [b][a][ke][p]
        keep["Fscale"] = _f64(mesh.Fscale)
        keep["normal_fn"] = _f64(mesh.normal_fn)
        keep["J"] = _f64(mesh.J)
        keep["Gsqrt"] = _f64(mesh.Gsqrt)
        keep["GI3"] = _f64(mesh.GI3)
        keep["GsqrtH"] = _f64(mesh.GsqrtH)
        keep["zlev"] = _f64(mesh.zlev)
        keep["VMapM"], keep["VMapP"] = mesh.abi_vmapM(), mesh.abi_vmapP()
        keep["VMapB"], keep["EMap3Dto2D"] = mesh.abi_vmapB(), mesh.abi_emap3dto2d()
        if getattr(mesh, "panelID", 0):      # cubed-sphere panel tile: Fortran shapes (Nfp,Ne2D,2,2), (Np,NeA), (Nfp,Ne2D,2)
            keep["GIJ"] = _f64(mesh.GIJ.transpose(1, 0, 2, 3))
            keep["gam"] = _f64(mesh.gam)
            keep["pos2D"] = _f64(mesh.pos2D)
            d.panelID = int(mesh.panelID)
        for k, a in keep.items():
            setattr(d, k, a.ctypes.data_as(C.c_void_p))
        tile_rank = tile_rank or (lambda pi, pj: my_rank)
        for f in range(6):
            (qi, qj), fo = mesh.tile_neighbors[f]
            d.nbr_rank[f] = tile_rank(qi, qj)
            d.nbr_face[f] = fo + 1
            d.vel_bc[f] = int(bc6[f])
        d.my_rank = my_rank
        for k in ("GRAV", "Rdry", "CPdry", "CVdry", "PRES00", "OHM"):
            setattr(d, k, float(consts[k]))
        self.consts = dict(consts)
        h = C.c_void_p()
        _lib.check(self.L.fedg_create(C.byref(d), C.byref(h)))
        keep.clear()          # fedg_create copied what it needs to the device: the expanded Fortran-shaped arrays are not kept on the host
        self.h = h
        self.n_field = mesh.NeA * elem.Np
        self.n_int = mesh.Ne * elem.Np

    # ---- multi-GPU: NCCL communicator over the ranks of the tile graph ------------------------
    def init_comm(self, rank: int, nranks: int, bcast):
        """bcast(bytes_or_None) -> bytes: broadcasts rank 0's 128-byte NCCL unique id with the caller's transport
        (MPI_Bcast in the Fortran driver, torch.distributed here)."""
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            _lib.check(self.L.fedg_comm_unique_id(buf))
        raw = bcast(bytes(buf) if rank == 0 else None)
        buf2 = (C.c_ubyte * 128).from_buffer_copy(raw)
        _lib.check(self.L.fedg_comm_init(self.h, buf2, int(rank), int(nranks)))

    # ---- reference-style lifecycle -------------------------------------------------------
    def Init(self, EQS_TYPE: str, TINTEG_TYPE: str, TIME_DT: float, MODALFILTER_FLAG: bool = False,
             MF_ETAC_h=2.0 / 3.0, MF_ALPHA_h=1.0, MF_ORDER_h=16, MF_ETAC_v=2.0 / 3.0, MF_ALPHA_v=1.0, MF_ORDER_v=16):
        fh = fv = None
        if MODALFILTER_FLAG:
            # Setup_ModalFilter (tensorprod3D.F90.erb:160-181): 1D matrices, column-major
            fh = _f64(self.elem.filter1d(MF_ETAC_h, MF_ALPHA_h, MF_ORDER_h).T)
            fv = _f64(self.elem.filter1d(MF_ETAC_v, MF_ALPHA_v, MF_ORDER_v).T)
        _lib.check(self.L.fedg_dyn_init(self.h, EQS_TYPE.encode(), TINTEG_TYPE.encode(), float(TIME_DT),
                                        int(bool(MODALFILTER_FLAG)), _ptr(fh), _ptr(fv)))
        self.dt = float(TIME_DT)

    def Final(self):
        if getattr(self, "h", None):
            self.L.fedg_destroy(self.h)
            self.h = None

    __del__ = Final

    # ---- state ---------------------------------------------------------------------------
    def _chk_field(self, a, n=None):
        a = _f64(a).reshape(-1)
        if a.size != (n or self.n_field):
            raise ValueError(f"field must have {n or self.n_field} values (Np*NeA), got {a.size}")
        return a

    def set_prog(self, DDENS, MOMX, MOMY, MOMZ, DRHOT):
        arrs = [self._chk_field(a) for a in (DDENS, MOMX, MOMY, MOMZ, DRHOT)]
        _lib.check(self.L.fedg_set_prog(self.h, *[_ptr(a) for a in arrs]))

    def get_prog(self):
        out = [np.zeros(self.n_field) for _ in range(5)]
        _lib.check(self.L.fedg_get_prog(self.h, *[_ptr(a) for a in out]))
        return dict(zip(PROG_NAMES, out))

    def set_aux(self, DENS_hyd, PRES_hyd, Rtot=None, CVtot=None, CPtot=None, THERM_hyd=None):
        c = self.consts
        full = lambda v: np.full(self.n_field, float(v))
        a = [self._chk_field(DENS_hyd), self._chk_field(PRES_hyd),
             None if THERM_hyd is None else self._chk_field(THERM_hyd),
             self._chk_field(Rtot) if Rtot is not None else full(c["Rdry"]),
             self._chk_field(CVtot) if CVtot is not None else full(c["CVdry"]),
             self._chk_field(CPtot) if CPtot is not None else full(c["CPdry"])]
        _lib.check(self.L.fedg_set_aux(self.h, *[_ptr(x) for x in a]))

    def set_phyd_hgrad(self, DPhydDx, DPhydDy):
        a = None if DPhydDx is None else self._chk_field(DPhydDx)
        b = None if DPhydDy is None else self._chk_field(DPhydDy)
        _lib.check(self.L.fedg_set_phyd_hgrad(self.h, _ptr(a), _ptr(b)))

    def update_phyd_hgrad(self, PRES_hyd_ref=None):
        """update_phyd_hgrad (driver_nonhydro3d.F90:1060-1095) on the device, from the registered PRES_hyd and its exchanged halo."""
        a = None if PRES_hyd_ref is None else self._chk_field(PRES_hyd_ref)
        _lib.check(self.L.fedg_update_phyd_hgrad(self.h, _ptr(a)))

    def set_phy_tend(self, DENS_tp, MOMX_tp, MOMY_tp, MOMZ_tp, RHOT_tp, RHOH_p):
        """Physics tendencies of add_phy_tend (driver_nonhydro3d.F90:1098-1178); None switches them off."""
        a = [None if x is None else self._chk_field(x) for x in (DENS_tp, MOMX_tp, MOMY_tp, MOMZ_tp, RHOT_tp, RHOH_p)]
        _lib.check(self.L.fedg_set_phy_tend(self.h, *[_ptr(x) for x in a]))

    def numdiff_init(self, ND_LAPLACIAN_NUM=1, ND_COEF_h=0.0, ND_COEF_v=0.0, therm_bc: dict | None = None, apply_in_update=True):
        """PARAM_ATMOS_DYN_NUMDIFF (scale_atm_dyn_dgm_nonhydro3d_numdiff.F90:119-212).  therm_bc: {'south',...,'top'} -> 'ADIABAT'."""
        tb = np.zeros(6, dtype=np.int32)
        names = ("south", "east", "north", "west", "btm", "top")
        for k, v in (therm_bc or {}).items():
            tb[names.index(k)] = 1 if str(v).upper() == "ADIABAT" else 0
        self._nd_tb = tb
        _lib.check(self.L.fedg_numdiff_init(self.h, int(ND_LAPLACIAN_NUM), float(ND_COEF_h), float(ND_COEF_v), _ptr(tb), int(apply_in_update)))

    # ---- tracer advection with a prescribed mass flux (row f4; see csrc/tracer.cu for its validation status) ----
    def trcadv_init(self, TINTEG_TYPE: str, TIME_DT: float, MODALFILTER_FLAG: bool = False, MF_ETAC_h=0.0, MF_ALPHA_h=1.0, MF_ORDER_h=16,
                    MF_ETAC_v=0.0, MF_ALPHA_v=1.0, MF_ORDER_v=16, disable_limiter: bool = False):
        """AtmDynDGMDriver_trcadv3d with ONLY_TRACERADV_FLAG: the mass flux is the momentum of the registered state."""
        fh = fv = None
        if MODALFILTER_FLAG:
            fh = _f64(self.elem.filter1d(MF_ETAC_h, MF_ALPHA_h, MF_ORDER_h).T)
            fv = _f64(self.elem.filter1d(MF_ETAC_v, MF_ALPHA_v, MF_ORDER_v).T)
        _lib.check(self.L.fedg_trcadv_init(self.h, TINTEG_TYPE.encode(), float(TIME_DT), int(bool(MODALFILTER_FLAG)), _ptr(fh), _ptr(fv),
                                           int(bool(disable_limiter))))

    def trcadv_update(self, q, nsteps: int = 1, rhoq_tp=None):
        """q: (Np*NeA,) host array, interior advanced in place."""
        q = self._chk_field(q)
        tp = None if rhoq_tp is None else self._chk_field(rhoq_tp)
        _lib.check(self.L.fedg_trcadv_update(self.h, _ptr(q), _ptr(tp), int(nsteps)))
        return q

    def trcadv_couple(self, on: bool = True):
        """Tracer coupling: the dynamics stages save the averaged mass flux for fedg_trcadv_update_coupled."""
        _lib.check(self.L.fedg_trcadv_couple(self.h, int(bool(on))))

    def trcadv_update_coupled(self, q, rhoq_tp=None):
        """One tracer step with the stage-averaged mass flux of the last dynamics step; q (Np*NeA,) advanced in place."""
        q = self._chk_field(q)
        tp = None if rhoq_tp is None else self._chk_field(rhoq_tp)
        _lib.check(self.L.fedg_trcadv_update_coupled(self.h, _ptr(q), _ptr(tp)))
        return q

    def sponge_init(self, SL_WDAMP_TAU=-1.0, SL_WDAMP_HEIGHT=-1.0, SL_WDAMP_LAYER=-1, SL_HORIVELDAMP_FLAG=False):
        """PARAM_ATMOS_DYN_SPONGELAYER (scale_atm_dyn_dgm_spongelayer.F90:55-118)."""
        m = self.mesh
        if np.any(m.Gsqrt[:m.Ne] != 1.0) and not getattr(m, "panelID", 0):      # topography: the profile lives in the computational height
            z = _f64(m.pos_en[2])
            _lib.check(self.L.fedg_sponge_init_pos(self.h, float(SL_WDAMP_TAU), float(SL_WDAMP_HEIGHT), int(SL_WDAMP_LAYER),
                                                   int(SL_HORIVELDAMP_FLAG), _ptr(z)))
            return
        _lib.check(self.L.fedg_sponge_init(self.h, float(SL_WDAMP_TAU), float(SL_WDAMP_HEIGHT), int(SL_WDAMP_LAYER), int(SL_HORIVELDAMP_FLAG)))

    def numdiff_apply(self):
        _lib.check(self.L.fedg_numdiff_apply(self.h))

    def set_coriolis(self, cor):
        a = None if cor is None else _f64(cor).reshape(-1)
        _lib.check(self.L.fedg_set_coriolis(self.h, _ptr(a)))

    # ---- the step ------------------------------------------------------------------------
    def Update(self, nsteps: int = 1):
        """AtmDynDGMDriver_nonhydro3d%Update, state resident on the GPU."""
        _lib.check(self.L.fedg_dyn_update(self.h, int(nsteps)))

    def Update_host(self, fields: dict, nsteps: int = 1):
        """Update called with host arrays (in/out), as the reference driver is: H2D, steps, D2H."""
        arrs = [fields[k] for k in PROG_NAMES]
        for a in arrs:
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.size == self.n_field
        _lib.check(self.L.fedg_dyn_update_host(self.h, *[_ptr(a) for a in arrs], int(nsteps)))

    def Update_host_async(self, fields_in: dict, fields_out: dict, nsteps: int = 1, slot: int = 0):
        """Pipelined Update_host: returns once the work is queued; `Update_host_wait(slot)` completes it.  Two slots."""
        a = [fields_in[k] for k in PROG_NAMES] + [fields_out[k] for k in PROG_NAMES]
        for x in a:
            assert x.dtype == np.float64 and x.flags.c_contiguous and x.size == self.n_field
        _lib.check(self.L.fedg_dyn_update_host_async(self.h, *[_ptr(x) for x in a], int(nsteps), int(slot)))

    def Update_host_wait(self, slot: int = 0):
        _lib.check(self.L.fedg_dyn_update_host_wait(self.h, int(slot)))

    # ---- stage-level seams (timeint_rk%StoreVar0 / StoreImplicit / Advance, cal_tend_ex, cal_vi, MeshFieldComm_Exchange / _Get) ----
    def rk_store_var0(self): _lib.check(self.L.fedg_rk_store_var0(self.h))
    def rk_store_implicit(self, stage): _lib.check(self.L.fedg_rk_store_implicit(self.h, int(stage)))
    def rk_advance(self, stage): _lib.check(self.L.fedg_rk_advance(self.h, int(stage)))
    def cal_tend_ex_dev(self, stage): _lib.check(self.L.fedg_cal_tend_ex_dev(self.h, int(stage)))
    def cal_vi_dev(self, stage): _lib.check(self.L.fedg_cal_vi_dev(self.h, int(stage)))
    def halo_start(self): _lib.check(self.L.fedg_halo_start(self.h))
    def halo_wait(self): _lib.check(self.L.fedg_halo_wait(self.h))
    def modalfilter_apply(self): _lib.check(self.L.fedg_modalfilter_apply(self.h))

    def rk_get_tend(self, stage, implicit=False):
        out = [np.zeros(self.n_int) for _ in range(5)]
        _lib.check(self.L.fedg_rk_get_tend(self.h, int(bool(implicit)), int(stage), *[_ptr(a) for a in out]))
        return dict(zip(("DENS_dt", "MOMX_dt", "MOMY_dt", "MOMZ_dt", "RHOT_dt"), out))

    def Update_by_stages(self, nstage: int, hevi: bool, overlap_halo: bool = False):
        """One step through the stage-level seams, in the order of the reference's driver loop (driver_nonhydro3d.F90:703-951)."""
        self.rk_store_var0()
        for s in range(1, nstage + 1):
            if hevi:
                self.cal_vi_dev(s); self.rk_store_implicit(s)
            self.halo_start()
            if not overlap_halo:
                self.halo_wait()
            self.cal_tend_ex_dev(s)
            self.rk_advance(s)
        self.modalfilter_apply()

    def elem_div(self, vec_in: np.ndarray, vec_in_lift: np.ndarray, nelem: int):
        """ElementOperationBase3D%Div: vec_in (nelem,3,Np), vec_in_lift (nelem,NfpTot) -> (nelem,4,Np)."""
        a, b = _f64(vec_in).reshape(-1), _f64(vec_in_lift).reshape(-1)
        out = np.zeros(nelem * 4 * self.elem.Np)
        _lib.check(self.L.fedg_elem_div(self.h, _ptr(a), _ptr(b), _ptr(out), int(nelem)))
        return out.reshape(nelem, 4, self.elem.Np)

    def cal_tend_ex(self):
        out = [np.zeros(self.n_int) for _ in range(5)]
        _lib.check(self.L.fedg_cal_tend_ex(self.h, *[_ptr(a) for a in out]))
        return dict(zip(("DENS_dt", "MOMX_dt", "MOMY_dt", "MOMZ_dt", "RHOT_dt"), out))

    def cal_vi(self, impl_fac: float, var0: dict):
        """atm_dyn_nonhydro3d_cal_vi: implicit tendency of the device state about var0 ({name: (Np*NeA,) or (Np*Ne,)})."""
        a = [_f64(var0[k]).reshape(-1)[: self.n_int].copy() for k in PROG_NAMES]
        out = [np.zeros(self.n_int) for _ in range(5)]
        _lib.check(self.L.fedg_cal_vi(self.h, float(impl_fac), *[_ptr(x) for x in a], *[_ptr(x) for x in out]))
        return dict(zip(("DENS_dt", "MOMX_dt", "MOMY_dt", "MOMZ_dt", "RHOT_dt"), out))

    def get_pres(self):
        P, D = np.zeros(self.n_int), np.zeros(self.n_int)
        _lib.check(self.L.fedg_get_pres(self.h, _ptr(P), _ptr(D)))
        return P, D

    def exchange_halo(self, apply_bc=True):
        _lib.check(self.L.fedg_exchange_halo(self.h, int(apply_bc)))

    def monitor(self):
        out = np.zeros(5)
        _lib.check(self.L.fedg_monitor(self.h, _ptr(out)))
        return out

    def elem_op(self, name: str, a: np.ndarray, nelem: int):
        a = _f64(a).reshape(-1)
        out = np.zeros(self.elem.Np * nelem)
        _lib.check(self.L.fedg_elem_op(self.h, name.encode(), _ptr(a), _ptr(out), int(nelem)))
        return out

    def last_timing(self):
        t, k, n = C.c_double(), C.c_double(), C.c_long()
        _lib.check(self.L.fedg_last_timing(self.h, C.byref(t), C.byref(k), C.byref(n)))
        return dict(ms_total=t.value, ms_stage_kernels=k.value, launches=n.value)


def rk_tables(scheme: str) -> dict:
    """timeint_rk coefficient tables from the library (scale_timeint_rk_butcher_tab.F90)."""
    L = _lib.load()
    n = [C.c_int() for _ in range(4)]
    _lib.check(L.fedg_rk_info(scheme.encode(), *[C.byref(x) for x in n]))
    s = n[0].value
    a_ex, a_im = np.zeros((s, s)), np.zeros((s, s))
    b_ex, b_im = np.zeros(s), np.zeros(s)
    sig, gam = np.zeros((s + 1, s)), np.zeros((s + 1, s))
    _lib.check(L.fedg_rk_coef(scheme.encode(), *[_ptr(x) for x in (a_ex, b_ex, a_im, b_im, sig, gam)]))
    return dict(nstage=s, tend_buf_size=n[1].value, low_storage=bool(n[2].value), imex=bool(n[3].value),
                a_ex=a_ex, b_ex=b_ex, a_im=a_im, b_im=b_im, sig=sig, gam=gam)
