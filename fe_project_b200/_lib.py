"""ctypes loader of the in-tree CUDA library (fe_project_b200/libfedg.so).

There is no Python or CPU fallback: if the library is missing or does not export every
symbol of include/fedg.h, importing fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfedg.so")

# every extern "C" symbol declared in include/fedg.h
ABI_SYMBOLS = (
    "fedg_last_error", "fedg_version", "fedg_create", "fedg_destroy", "fedg_dyn_init",
    "fedg_set_prog", "fedg_get_prog", "fedg_set_aux", "fedg_set_phyd_hgrad", "fedg_set_coriolis",
    "fedg_dyn_update", "fedg_dyn_update_host", "fedg_cal_tend_ex", "fedg_cal_vi", "fedg_get_pres",
    "fedg_exchange_halo", "fedg_monitor", "fedg_rk_info", "fedg_rk_coef", "fedg_elem_op",
    "fedg_last_timing", "fedg_comm_unique_id", "fedg_comm_init",
    "fedg_set_phy_tend", "fedg_numdiff_init", "fedg_numdiff_apply", "fedg_sponge_init", "fedg_sponge_init_pos", "fedg_link_halo", "fedg_link_halo_recv", "fedg_link_halo_send", "fedg_group_exchange_halo", "fedg_group_update", "fedg_sparsemat_matmul", "fedg_sparsemat_matmul1", "fedg_sparsemat_matmul1_2", "fedg_sparsemat_matmul2", "fedg_advect3d_init", "fedg_advect3d_set", "fedg_advect3d_get",
    "fedg_advect3d_cal_tend", "fedg_advect3d_update", "fedg_trcadv_init", "fedg_trcadv_update", "fedg_trcadv_couple", "fedg_trcadv_update_coupled",
    "fedg_dyn_update_host_async", "fedg_dyn_update_host_wait", "fedg_rk_store_var0", "fedg_rk_store_implicit", "fedg_rk_advance",
    "fedg_cal_tend_ex_dev", "fedg_cal_vi_dev", "fedg_halo_start", "fedg_halo_wait", "fedg_modalfilter_apply", "fedg_rk_get_tend",
    "fedg_elem_div", "fedg_group_exchange_aux", "fedg_update_phyd_hgrad",
)


class MeshDesc(C.Structure):
    _fields_ = (
        [("polyorder", C.c_int)] + [(n, C.c_int) for n in ("Ne", "NeA", "NeX", "NeY", "NeZ", "Ne2D", "Nhalo")]
        + [(n, C.c_void_p) for n in ("D1D", "Lift", "VPOrdM1", "IntWeight_lgl", "Escale", "Fscale", "normal_fn",
                                     "J", "Gsqrt", "GI3", "GsqrtH", "zlev", "VMapM", "VMapP", "VMapB", "EMap3Dto2D")]
        + [("nbr_rank", C.c_int * 6), ("nbr_face", C.c_int * 6), ("my_rank", C.c_int), ("vel_bc", C.c_int * 6)]
        + [(n, C.c_double) for n in ("GRAV", "Rdry", "CPdry", "CVdry", "PRES00", "OHM")]
        + [(n, C.c_void_p) for n in ("GIJ", "gam", "pos2D")] + [("panelID", C.c_int)]
    )


class SparseMatDesc(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("col_size", C.c_int), ("val", C.c_void_p), ("colIdx", C.c_void_p)]


class SparseMatAnyDesc(C.Structure):
    _fields_ = [("storage_format_id", C.c_int), ("M", C.c_int), ("N", C.c_int), ("nnz", C.c_int), ("col_size", C.c_int), ("rowPtrSize", C.c_int),
                ("val", C.c_void_p), ("colIdx", C.c_void_p), ("rowPtr", C.c_void_p)]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). fe_project_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    missing = [s for s in ABI_SYMBOLS if not hasattr(L, s)]
    if missing:
        raise ImportError(f"{LIB_PATH} lacks ABI symbols: {missing}")
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    L.fedg_last_error.restype = C.c_char_p
    L.fedg_create.argtypes = [C.POINTER(MeshDesc), C.POINTER(vp)]
    L.fedg_destroy.argtypes = [vp]
    L.fedg_destroy.restype = None
    L.fedg_dyn_init.argtypes = [vp, C.c_char_p, C.c_char_p, cd, ci, vp, vp]
    L.fedg_set_prog.argtypes = [vp] * 6
    L.fedg_get_prog.argtypes = [vp] * 6
    L.fedg_set_aux.argtypes = [vp] * 7
    L.fedg_set_phyd_hgrad.argtypes = [vp] * 3
    L.fedg_set_coriolis.argtypes = [vp] * 2
    L.fedg_set_phy_tend.argtypes = [vp] * 7
    L.fedg_numdiff_init.argtypes = [vp, ci, cd, cd, vp, ci]
    L.fedg_numdiff_apply.argtypes = [vp]
    L.fedg_sponge_init.argtypes = [vp, cd, cd, ci, ci]
    L.fedg_sponge_init_pos.argtypes = [vp, cd, cd, ci, ci, vp]
    L.fedg_link_halo.argtypes = [vp, ci, vp, vp, vp]
    L.fedg_group_update.argtypes = [vp, ci, ci]
    L.fedg_link_halo_recv.argtypes = [vp, ci, ci, ci, vp]
    L.fedg_link_halo_send.argtypes = [vp, ci, ci, vp, ci]
    L.fedg_group_exchange_halo.argtypes = [vp, ci, ci]
    L.fedg_group_exchange_aux.argtypes = [vp, ci]
    L.fedg_update_phyd_hgrad.argtypes = [vp, vp]
    L.fedg_dyn_update.argtypes = [vp, ci]
    L.fedg_dyn_update_host.argtypes = [vp] * 6 + [ci]
    L.fedg_cal_tend_ex.argtypes = [vp] * 6
    L.fedg_cal_vi.argtypes = [vp, cd] + [vp] * 10
    L.fedg_get_pres.argtypes = [vp] * 3
    L.fedg_exchange_halo.argtypes = [vp, ci]
    L.fedg_monitor.argtypes = [vp, vp]
    L.fedg_rk_info.argtypes = [C.c_char_p] + [vp] * 4
    L.fedg_rk_coef.argtypes = [C.c_char_p] + [vp] * 6
    L.fedg_elem_op.argtypes = [vp, C.c_char_p, vp, vp, ci]
    L.fedg_last_timing.argtypes = [vp, vp, vp, vp]
    L.fedg_comm_unique_id.argtypes = [vp]
    L.fedg_comm_init.argtypes = [vp, vp, ci, ci]
    sp = C.POINTER(SparseMatDesc)
    L.fedg_sparsemat_matmul.argtypes = [sp, vp, vp, ci]
    spa = C.POINTER(SparseMatAnyDesc)
    L.fedg_sparsemat_matmul1.argtypes = [spa, vp, vp]
    L.fedg_sparsemat_matmul1_2.argtypes = [spa, vp, vp, vp]
    L.fedg_sparsemat_matmul2.argtypes = [spa, vp, vp, ci]
    L.fedg_advect3d_init.argtypes = [vp, C.c_char_p, cd, sp, sp, sp, sp]
    L.fedg_advect3d_set.argtypes = [vp] * 5
    L.fedg_advect3d_get.argtypes = [vp, vp]
    L.fedg_advect3d_cal_tend.argtypes = [vp, vp]
    L.fedg_advect3d_update.argtypes = [vp, ci]
    L.fedg_trcadv_init.argtypes = [vp, C.c_char_p, cd, ci, vp, vp, ci]
    L.fedg_trcadv_update.argtypes = [vp, vp, vp, ci]
    L.fedg_trcadv_couple.argtypes = [vp, ci]
    L.fedg_trcadv_update_coupled.argtypes = [vp, vp, vp]
    L.fedg_dyn_update_host_async.argtypes = [vp] * 11 + [ci, ci]
    L.fedg_dyn_update_host_wait.argtypes = [vp, ci]
    for name in ("fedg_rk_store_var0", "fedg_halo_start", "fedg_halo_wait", "fedg_modalfilter_apply"):
        getattr(L, name).argtypes = [vp]
    for name in ("fedg_rk_store_implicit", "fedg_rk_advance", "fedg_cal_tend_ex_dev", "fedg_cal_vi_dev"):
        getattr(L, name).argtypes = [vp, ci]
    L.fedg_rk_get_tend.argtypes = [vp, ci, ci] + [vp] * 5
    L.fedg_elem_div.argtypes = [vp, vp, vp, vp, ci]
    _lib = L
    return L


class FedgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fedg error {code}: {msg}")
        self.code = code


def check(rc: int):
    if rc != 0:
        raise FedgError(rc, load().fedg_last_error().decode(errors="replace"))
