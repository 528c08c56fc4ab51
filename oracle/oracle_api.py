"""ctypes binding of the CPU oracle (oracle/libfeoracle.so).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _variant_so():
    """FEO_VARIANT=perf (set by the timing legs of bench.py before the first call): the -O3 / FMA build for the widest vector
    ISA of this host; default: the parity build (-ffp-contract=off), the one every test compares against."""
    if os.environ.get("FEO_VARIANT", "") == "perf":
        try:
            flags = open("/proc/cpuinfo").read()
        except OSError:
            flags = ""
        return "libfeoracle_perf_v4.so" if ("avx512f" in flags and "avx512vl" in flags and "avx512dq" in flags) else "libfeoracle_perf_v3.so"
    return "libfeoracle.so"


def build_oracle():
    so = os.path.join(_ORACLE_DIR, _variant_so())
    srcs = [os.path.join(_ORACLE_DIR, f) for f in os.listdir(_ORACLE_DIR) if f.endswith((".cpp", ".hpp"))]
    if (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-j8", "-C", _ORACLE_DIR])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build_oracle())
        L.feo_create.restype = C.c_void_p
        L.feo_create.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_void_p]
        L.feo_create_panel.restype = C.c_void_p
        L.feo_create_panel.argtypes = [C.c_int] * 6 + [C.c_double, C.c_void_p, C.c_double, C.c_int]
        L.feo_destroy.argtypes = [C.c_void_p]
        L.feo_dims.argtypes = [C.c_void_p, C.c_void_p]
        L.feo_array.restype = C.POINTER(C.c_double)
        L.feo_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_long)]
        L.feo_iarray.restype = C.POINTER(C.c_int)
        L.feo_iarray.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_long)]
        L.feo_set_consts.argtypes = [C.c_void_p, C.c_void_p]
        L.feo_setup_dyn.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.feo_prepare.argtypes = [C.c_void_p]
        L.feo_update.argtypes = [C.c_void_p, C.c_int]
        L.feo_set_phytend.argtypes = [C.c_void_p, C.c_int]
        L.feo_set_sponge.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
        L.feo_set_numdiff.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p]
        L.feo_numdiff_apply.argtypes = [C.c_void_p]
        L.feo_sphere_exchange.argtypes = [C.c_void_p, C.c_int]
        L.feo_sphere_exchange_aux.argtypes = [C.c_void_p]
        L.feo_sphere_update.argtypes = [C.c_void_p, C.c_int]
        L.feo_monitor.argtypes = [C.c_void_p, C.c_void_p]
        L.feo_set_tracer_coupling.argtypes = [C.c_void_p, C.c_int]
        L.feo_trcadv_update_coupled.argtypes = [C.c_void_p, C.c_char_p, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.feo_trcadv_update.argtypes = [C.c_void_p, C.c_char_p, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.feo_cal_vi.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.feo_stage_piece.argtypes = [C.c_void_p, C.c_char_p]
        L.feo_elem_op.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.feo_lift_dense.argtypes = [C.c_void_p, C.c_void_p]
        L.feo_dmat_dense.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.feo_sparsemat_matmul.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.feo_rk_info.argtypes = [C.c_char_p] + [C.c_void_p] * 4
        L.feo_rk_coef.argtypes = [C.c_char_p] + [C.c_void_p] * 6
        L.feo_rk_create.restype = C.c_void_p
        L.feo_rk_create.argtypes = [C.c_char_p, C.c_double, C.c_int, C.c_long]
        L.feo_rk_destroy.argtypes = [C.c_void_p]
        L.feo_rk_tend.restype = C.POINTER(C.c_double)
        L.feo_rk_tend.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.feo_rk_implicit_fac.restype = C.c_double
        L.feo_rk_implicit_fac.argtypes = [C.c_void_p, C.c_int]
        L.feo_rk_store_implicit.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.feo_rk_advance.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.feo_last_error.restype = C.c_char_p
        L.feo_set_num_threads.argtypes = [C.c_int]
        L.feo_advect3d_create.restype = C.c_void_p
        L.feo_advect3d_create.argtypes = [C.c_void_p, C.c_char_p, C.c_double, C.c_int]
        L.feo_advect3d_destroy.argtypes = [C.c_void_p]
        L.feo_advect3d_array.restype = C.POINTER(C.c_double)
        L.feo_advect3d_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_long)]
        L.feo_advect3d_sparsemat.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.feo_advect3d_cal_tend.argtypes = [C.c_void_p, C.c_void_p]
        L.feo_advect3d_update.argtypes = [C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """One single-tile regional run of the CPU restatement."""

    def __init__(self, p, NeX, NeY, NeZ, dom=None, periodic=(False, False, False), lumped=False, FZ=None, panel=None):
        """panel = dict(panelID, ztop, RPlanet): one cubed-sphere panel tile instead of a regional cube."""
        L = lib()
        fz = None if FZ is None else np.ascontiguousarray(FZ, dtype=np.float64)
        if panel is not None:
            self.h = L.feo_create_panel(p, int(lumped), int(panel["panelID"]), NeX, NeY, NeZ, float(panel["ztop"]), _p(fz),
                                        float(panel["RPlanet"]), 1)
        else:
            dom = np.asarray(dom, dtype=np.float64)
            per = np.asarray(periodic, dtype=np.int32)
            self.h = L.feo_create(p, int(lumped), NeX, NeY, NeZ, _p(dom), _p(fz), _p(per))
        if not self.h:
            raise RuntimeError(L.feo_last_error().decode())
        d = np.zeros(8, dtype=np.int32)
        L.feo_dims(self.h, _p(d))
        self.Np, self.NfpTot, self.Ne, self.NeA, self.Nhalo, self.np1, self.Ne2D = (int(x) for x in d[:7])

    def __del__(self):
        if getattr(self, "h", None):
            lib().feo_destroy(self.h)
            self.h = None

    def _chk(self, rc):
        if rc:
            raise RuntimeError(lib().feo_last_error().decode())

    def arr(self, name) -> np.ndarray:
        n = C.c_long()
        ptr = lib().feo_array(self.h, name.encode(), C.byref(n))
        if not ptr:
            raise KeyError(name)
        return np.ctypeslib.as_array(ptr, shape=(n.value,))

    def iarr(self, name) -> np.ndarray:
        n = C.c_long()
        ptr = lib().feo_iarray(self.h, name.encode(), C.byref(n))
        if not ptr:
            raise KeyError(name)
        return np.ctypeslib.as_array(ptr, shape=(n.value,))

    def set_consts(self, c):
        a = np.array([c["GRAV"], c["Rdry"], c["CPdry"], c["CVdry"], c["PRES00"], c.get("OHM", 7.292e-5)], dtype=np.float64)
        lib().feo_set_consts(self.h, _p(a))

    def setup_dyn(self, eqs, tinteg, dt, modalfilter=False, mf=(0, 0, 0, 0, 0, 0), vel_bc=(0,) * 6):
        mf = np.asarray(mf, dtype=np.float64)
        bc = np.asarray(vel_bc, dtype=np.int32)
        self._chk(lib().feo_setup_dyn(self.h, eqs.encode(), tinteg.encode(), dt, int(modalfilter), _p(mf), _p(bc)))

    def set_numdiff(self, on=True, laplacian_num=1, coef_h=0.0, coef_v=0.0, therm_bc=(0,) * 6):
        """PARAM_ATMOS_DYN_NUMDIFF; call after setup_dyn (takes dt and the velocity BCs from it)."""
        tb = np.asarray(therm_bc, dtype=np.int32)
        lib().feo_set_numdiff(self.h, int(on), int(laplacian_num), float(coef_h), float(coef_v), _p(tb))

    def numdiff_apply(self):
        self._chk(lib().feo_numdiff_apply(self.h))

    def set_sponge(self, on=True, SL_WDAMP_TAU=-1.0, SL_WDAMP_HEIGHT=-1.0, SL_WDAMP_LAYER=-1, SL_HORIVELDAMP_FLAG=False):
        lib().feo_set_sponge(self.h, int(on), float(SL_WDAMP_TAU), float(SL_WDAMP_HEIGHT), int(SL_WDAMP_LAYER), int(SL_HORIVELDAMP_FLAG))

    def set_phytend(self, on=True):
        lib().feo_set_phytend(self.h, int(on))

    def trcadv_update(self, q, tinteg, dt, nsteps=1, modalfilter=None, disable_limiter=False, rhoq_tp=None):
        """Tracer advection with the mass flux of the handle's state (ONLY_TRACERADV_FLAG); q (Np*NeA) is advanced in place."""
        assert q.dtype == np.float64 and q.flags.c_contiguous
        mf = np.asarray(modalfilter if modalfilter else (0, 0, 0, 0, 0, 0), dtype=np.float64)
        tp = None if rhoq_tp is None else np.ascontiguousarray(rhoq_tp, dtype=np.float64)
        self._chk(lib().feo_trcadv_update(self.h, tinteg.encode(), float(dt), int(nsteps), int(modalfilter is not None), _p(mf),
                                          int(disable_limiter), _p(q), None if tp is None else _p(tp)))

    def set_tracer_coupling(self, on=True):
        lib().feo_set_tracer_coupling(self.h, int(on))

    def trcadv_update_coupled(self, q, tinteg, dt, modalfilter=None, disable_limiter=False, rhoq_tp=None):
        """One tracer step with the stage-averaged mass flux of the LAST dynamics step (set_tracer_coupling before update)."""
        assert q.dtype == np.float64 and q.flags.c_contiguous
        mf = np.asarray(modalfilter if modalfilter else (0, 0, 0, 0, 0, 0), dtype=np.float64)
        tp = None if rhoq_tp is None else np.ascontiguousarray(rhoq_tp, dtype=np.float64)
        self._chk(lib().feo_trcadv_update_coupled(self.h, tinteg.encode(), float(dt), int(modalfilter is not None), _p(mf),
                                                  int(disable_limiter), _p(q), None if tp is None else _p(tp)))

    def prepare(self):
        self._chk(lib().feo_prepare(self.h))

    def update(self, nsteps=1):
        self._chk(lib().feo_update(self.h, nsteps))

    def piece(self, what):
        self._chk(lib().feo_stage_piece(self.h, what.encode()))

    def cal_vi(self, impl_fac, dt, var0):
        """var0: (5, Np*NeA) in the order DENS, RHOT, MOMZ, MOMX, MOMY; returns the implicit tendencies, same shape."""
        var0 = np.ascontiguousarray(var0, dtype=np.float64)
        out = np.zeros_like(var0)
        self._chk(lib().feo_cal_vi(self.h, float(impl_fac), float(dt), _p(var0), _p(out)))
        return out

    def monitor(self):
        out = np.zeros(5)
        lib().feo_monitor(self.h, _p(out))
        return out

    def elem_op(self, name, a, b=None, nout=None):
        out = np.zeros(nout if nout is not None else self.Np)
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
        self._chk(lib().feo_elem_op(self.h, name.encode(), _p(a), _p(b), _p(out)))
        return out

    def lift_dense(self):
        out = np.zeros((self.Np, self.NfpTot))
        lib().feo_lift_dense(self.h, _p(out))
        return out

    def dmat_dense(self, d):
        out = np.zeros((self.Np, self.Np))
        lib().feo_dmat_dense(self.h, d, _p(out))
        return out


class OracleSphere:
    """Six panel oracles (GLOBALNONHYDRO3D_HEVI) + the panel-edge exchange and the six-panel step."""

    def __init__(self, panels):
        assert len(panels) == 6
        self.panels = list(panels)
        self._h = (C.c_void_p * 6)(*[p.h for p in self.panels])

    def _chk(self, rc):
        if rc:
            raise RuntimeError(lib().feo_last_error().decode())

    def exchange(self, with_dpres=True):
        self._chk(lib().feo_sphere_exchange(self._h, int(with_dpres)))

    def exchange_aux(self):
        self._chk(lib().feo_sphere_exchange_aux(self._h))

    def update(self, nsteps=1):
        self._chk(lib().feo_sphere_update(self._h, int(nsteps)))


class OracleAdvect3D:
    """sample/advect3d on the mesh of an Oracle (config 1): q, u, v, w (Np*NeA) + ERK integrator."""

    def __init__(self, oracle: Oracle, scheme="ERK_4s4o", dt=0.008, ell=True):
        self.o = oracle
        self.h = lib().feo_advect3d_create(oracle.h, scheme.encode(), float(dt), int(ell))
        if not self.h:
            raise RuntimeError("feo_advect3d_create failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().feo_advect3d_destroy(self.h)
            self.h = None

    def arr(self, name):
        n = C.c_long()
        ptr = lib().feo_advect3d_array(self.h, name.encode(), C.byref(n))
        if not ptr:
            raise KeyError(name)
        return np.ctypeslib.as_array(ptr, shape=(n.value,))

    def sparsemat(self, which):
        """ELL arrays of Dx/Dy/Dz/Lift (which = 0..3): (M, N, col_size, val[M*col_size], colIdx 0-based)."""
        M, N, cs = C.c_int(), C.c_int(), C.c_int()
        val, col = C.POINTER(C.c_double)(), C.POINTER(C.c_int)()
        lib().feo_advect3d_sparsemat(self.h, which, C.byref(M), C.byref(N), C.byref(cs), C.byref(val), C.byref(col))
        n = M.value * cs.value
        return M.value, N.value, cs.value, np.ctypeslib.as_array(val, shape=(n,)).copy(), np.ctypeslib.as_array(col, shape=(n,)).copy()

    def cal_tend(self):
        out = np.zeros(self.o.Np * self.o.Ne)
        lib().feo_advect3d_cal_tend(self.h, _p(out))
        return out

    def update(self, nsteps=1):
        lib().feo_advect3d_update(self.h, int(nsteps))


def sparsemat_matmul(A, b, eps, ell):
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    M, N = A.shape
    c = np.zeros(M)
    g = np.zeros((M, N))
    rc = lib().feo_sparsemat_matmul(_p(A), M, N, eps, int(ell), _p(b), _p(c), _p(g))
    assert rc == 0
    return c, g


def sparsemat_matmul_ex(A, b1, eps, ell, mode=0, b2=None, NQ=1):
    """mode 0: c = A b1; mode 1: c = A (b1 .* b2) (sparsemat_matmul1_2); mode 2: b1 is (N, NQ) C-order = Fortran b(NQ,N), returns (M, NQ)
    (sparsemat_matmul2).  Also returns the CSR arrays (1-based) of the matrix when ell is False."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    b1 = np.ascontiguousarray(b1, dtype=np.float64)
    b2 = None if b2 is None else np.ascontiguousarray(b2, dtype=np.float64)
    M, N = A.shape
    c = np.zeros((M, NQ) if mode == 2 else M)
    nnz = C.c_int()
    val, col, rp = np.zeros(M * N), np.zeros(M * N, dtype=np.int32), np.zeros(M + 1, dtype=np.int32)
    L = lib()
    L.feo_sparsemat_matmul_ex.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 7
    rc = L.feo_sparsemat_matmul_ex(_p(A), M, N, eps, int(ell), int(mode), int(NQ), _p(b1), None if b2 is None else _p(b2), _p(c),
                                   C.byref(nnz), _p(val), _p(col), _p(rp))
    assert rc == 0
    n = nnz.value
    return c, (val[:n].copy(), col[:n].copy(), rp)


def rk_tables(name):
    L = lib()
    n = [C.c_int() for _ in range(4)]
    rc = L.feo_rk_info(name.encode(), *[C.byref(x) for x in n])
    if rc:
        raise KeyError(name)
    s = n[0].value
    a_ex, a_im = np.zeros((s, s)), np.zeros((s, s))
    b_ex, b_im = np.zeros(s), np.zeros(s)
    sig, gam = np.zeros((s + 1, s)), np.zeros((s + 1, s))
    L.feo_rk_coef(name.encode(), _p(a_ex), _p(b_ex), _p(a_im), _p(b_im), _p(sig), _p(gam))
    return dict(nstage=s, tend_buf_size=n[1].value, low_storage=bool(n[2].value), imex=bool(n[3].value),
                a_ex=a_ex, b_ex=b_ex, a_im=a_im, b_im=b_im, sig=sig, gam=gam)


class OracleRK:
    def __init__(self, name, dt, nvar, n=1):
        self.h = lib().feo_rk_create(name.encode(), dt, nvar, n)
        if not self.h:
            raise RuntimeError(lib().feo_last_error().decode())
        self.n = n
        self.info = rk_tables(name)

    def __del__(self):
        if getattr(self, "h", None):
            lib().feo_rk_destroy(self.h)
            self.h = None

    def tend(self, im, var, stage):
        return np.ctypeslib.as_array(lib().feo_rk_tend(self.h, int(im), var, stage), shape=(self.n,))

    def implicit_fac(self, stage):
        return lib().feo_rk_implicit_fac(self.h, stage)

    def store_implicit(self, stage, q, var):
        lib().feo_rk_store_implicit(self.h, stage, _p(q), var)

    def advance(self, stage, q, var):
        lib().feo_rk_advance(self.h, stage, _p(q), var)
