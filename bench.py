#!/usr/bin/env python
"""bench.py -- DG dynamics DOF-updates/s of the nonhydro3d p=7 hot path on B200.

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line
(rank 0).  A "step" is one full dynamics time step (all RK stages, boundary conditions, modal filter,
final pressure diagnostic) of the regional density-current configuration (BASELINE.json configs[2],
32x32x16 elements, p=7, NONHYDRO3D_HEVE, ERK_SSP_4s3o) over the whole mesh.

  value   : 5*Np*Ne_global*K / t, state resident in HBM, t from CUDA events on the launching stream
  e2e     : the same metric through the host-buffer entry point (fedg_dyn_update_host): pinned host
            arrays -> H2D -> step -> D2H inside the timed region, one step per call
  roofline: fused stage kernel, algorithmic bytes (232 B/node/stage) / mean launch time
  cpu_baseline / --impl reference: the CPU oracle (reference-equivalent restatement; the Fortran
            reference cannot be built here) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

# The contract is ONE JSON line on stdout, but NCCL / torch write banners ("NCCL version ...") to file descriptor 1 from
# native code: everything goes to stderr until the result line is printed through the saved descriptor.
_STDOUT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    os.write(_STDOUT_FD, (json.dumps(line) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "DG dynamics DOF-updates/s (nonhydro3d p=7)"
UNIT = "DOF-updates/s"
ALG_BYTES_PER_NODE_STAGE = 232.0      # SURVEY.md 8(d), HEVE explicit stage fused with low-storage RK
WORKLOAD = dict(name="atm_nonhydro3d regional density current", p=7, NeX=32, NeY=32, NeZ=16,
                dom=(0.0, 25.6e3, 0.0, 25.6e3, 0.0, 6.4e3), dt=0.04, eqs="NONHYDRO3D_HEVE", tinteg="ERK_SSP_4s3o")


_ORIG_AFFINITY = None


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this process to the CPUs NVML reports as local to its GPU, BEFORE any pinned host buffer is allocated (first touch puts the
    pages on that NUMA node).  With one process per GPU and no binding the pinned staging buffers of the e2e leg land on whichever node the
    launcher started the process on, and at 8 ranks every host<->device copy crosses the socket interconnect (measured: 39 ms per
    one-step call at 8 GPUs against 7.3 ms at 1).  Best effort: any failure leaves the affinity as it was."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64 + 8)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            global _ORIG_AFFINITY
            _ORIG_AFFINITY = allowed
            os.sched_setaffinity(0, use)
        return sorted(use) if use else None
    except Exception:
        return None


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.gpu), "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def wait_started(self, timeout=1.0):
        """nvidia-smi needs a few hundred ms before its first line: wait for it, so that the samples fall INTO the timed region."""
        t0 = time.perf_counter()
        while self.p is not None and time.perf_counter() - t0 < timeout:
            if self.p.poll() is not None:          # nvidia-smi exited (no device / no permission): nothing to wait for
                return
            try:
                if os.path.getsize(self.f.name) > 0:
                    return
            except OSError:
                return
            time.sleep(0.02)

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def _oracle_modules():
    """The CPU oracle (oracle/, test infrastructure) is imported by the two CPU timing legs only (cpu_baseline, --impl reference),
    in its timing build (-O3, FMA contraction on; oracle/Makefile)."""
    os.environ["FEO_VARIANT"] = "perf"
    od = os.path.join(ROOT, "oracle")
    if od not in sys.path:
        sys.path.insert(0, od)
    import oracle_api
    import oracle_cases
    return oracle_api, oracle_cases


def workload_case(args, hevi=False, tiles=(1, 1), tile=(0, 0)):
    from fe_project_b200.cases import DensityCurrentCase
    NX, NY = tiles
    x0, x1, y0, y1, z0, z1 = WORKLOAD["dom"]
    dom = (x0, x0 + (x1 - x0) * NX, y0, y0 + (y1 - y0) * NY, z0, z1)
    return DensityCurrentCase(p=WORKLOAD["p"], NeX=args.nex, NeY=args.ney, NeZ=args.nez, dom=dom,
                              dt=(1.5 * WORKLOAD["dt"] if hevi else WORKLOAD["dt"]), tinteg=("IMEX_ARK324" if hevi else WORKLOAD["tinteg"]),
                              modalfilter=True, NprcX=NX, NprcY=NY, pi=tile[0], pj=tile[1],
                              eqs=("NONHYDRO3D_HEVI" if hevi else "NONHYDRO3D_HEVE"))


def oracle_timed(case, warmup=1, steps=None, target_s=12.0, threads=None):
    """Times the CPU oracle on the SAME configuration as the GPU arm (one tile of it), all host threads: `steps` steps, or as many
    as fit into about target_s seconds (at least 3)."""
    oracle_api, oracle_cases = _oracle_modules()
    if _ORIG_AFFINITY:                     # the CPU leg uses every host core again (the GPU legs are over by now)
        os.sched_setaffinity(0, _ORIG_AFFINITY)
    cores = threads or os.cpu_count() or 1
    used = oracle_api.lib().feo_set_num_threads(cores)      # explicit: torchrun exports OMP_NUM_THREADS=1
    o = oracle_cases.make_oracle_regional(case)
    t0 = time.perf_counter(); o.update(max(1, warmup)); t1 = (time.perf_counter() - t0) / max(1, warmup)
    n = steps if steps else int(max(3, min(200, target_s / max(t1, 1e-6))))
    t0 = time.perf_counter(); o.update(n); dt = time.perf_counter() - t0
    dof = 5 * case.elem.Np * case.mesh.Ne
    m = case.mesh
    return dict(value=dof * n / dt, steps=n, seconds=dt, cores=used, lib=oracle_api._variant_so(),
                sample=f"the bench configuration itself, one tile: density current {m.NeX}x{m.NeY}x{m.NeZ} elements p={case.p} ({dof} DOF), "
                       f"{n} steps of {case.tinteg} ({case.eqs}) + modal filter, {used} OpenMP threads, {oracle_api._variant_so()}")


def run_reference(args, rank, world):
    """CPU arm: the reference-equivalent C++ restatement (the Fortran reference needs gfortran + MPI + SCALE 5.5.5, none of which exist
    in this image) on the bench configuration, all host threads, timing build."""
    if rank != 0:
        return
    hevi = args.eqs == "hevi"
    case = workload_case(args, hevi=hevi)
    W, K = max(1, min(args.warmup, 3)), max(3, min(args.steps, 40 if not hevi else 10))
    r = oracle_timed(case, warmup=W, steps=K)
    line = dict(metric=METRIC, value=r["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * r["seconds"] / r["steps"], higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic", impl="reference",
                config=dict(workload=f"{WORKLOAD['name']}, {args.nex}x{args.ney}x{args.nez} elements p=7, {case.eqs}, {case.tinteg}, "
                                     f"dt={case.dt}, modal filter on (the GPU arm's configuration; {r['steps']} timed steps)"),
                cpu_baseline=dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port", sample=r["sample"]),
                e2e=dict(value=r["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note="CPU oracle = reference-equivalent C++ restatement, timing build (g++ -O3 -ffp-contract=fast -fopenmp, "
                     f"{r['lib']}); the Fortran reference needs gfortran+MPI+SCALE 5.5.5, none of which exist in this image")
    emit(line)


def run_advect3d(args):
    """configs[0]: sample/advect3d, 8x8x8 elements p=3, ERK_4s4o, dt 0.008 (test.conf); one variable."""
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    from fe_project_b200.advect3d import Advect3D, gaussian_hill
    from fe_project_b200.element import HexElement
    from fe_project_b200.mesh import LocalMeshCube
    ne = (8, 8, 8)
    e = HexElement(3)
    mesh = LocalMeshCube(e, *ne, 0, 1, 0, 1, 0, 1, periodic=(True, True, True))
    g = Advect3D(e, mesh, "ERK_4s4o", 0.008)
    q = gaussian_hill(mesh)
    u = np.zeros_like(q); u[:mesh.Ne] = 0.5
    g.set(q, u, u, u)
    W, K = max(3, args.warmup), args.steps
    g.update(W)
    sampler = ClockSampler(0); sampler.start(); sampler.wait_started()
    g.update(K)
    tm = g.last_timing()
    clocks = sampler.stop()
    dof = e.Np * mesh.Ne
    value = dof * K / (tm["ms_total"] * 1e-3)
    # e2e: host q in, K2 steps, host q out per call (the sample keeps q on the host between history writes)
    t0 = time.perf_counter()
    for _ in range(5):
        g.set(q, u, u, u); g.update(10); q2 = g.get()
    t_e2e = (time.perf_counter() - t0) / 5
    nbytes = 4 * q.size * 8
    cpu = None
    if not args.no_cpu_baseline:
        oracle_api, _ = _oracle_modules()
        Oracle, OracleAdvect3D = oracle_api.Oracle, oracle_api.OracleAdvect3D
        o = Oracle(3, *ne, (0, 1, 0, 1, 0, 1), periodic=(True, True, True))
        a = OracleAdvect3D(o, "ERK_4s4o", 0.008)
        a.arr("q")[:] = q.reshape(-1)
        for nm in "uvw":
            a.arr(nm)[:] = u.reshape(-1)
        a.update(10)
        t0 = time.perf_counter(); a.update(2000); dtc = time.perf_counter() - t0
        cpu = dict(value=dof * 2000 / dtc, unit=UNIT, cores=os.cpu_count(), kind="port", sample="the full case, 2000 steps")
    # dominant kernel: advect_stage_kernel; algorithmic bytes per node and stage: q, u, v, w, q0, varTmp in + q, varTmp out
    peak, peak_src = read_peaks()
    alg = 8 * 8.0 * dof
    ms_stage = tm["ms_total"] / (K * 4)
    emit(dict(
        metric=METRIC.replace("nonhydro3d p=7", "advect3d p=3"), value=value, unit=UNIT, n_gpus=1, steps=K, warmup=W,
        ms_per_step=tm["ms_total"] / K, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
        config=dict(workload="sample/advect3d 8x8x8 elements p=3, ERK_4s4o, dt=0.008, gaussian hill, u=v=w=0.5, triply periodic",
                    dof=dof, l2_policy="the whole case (1 MB) is L2 resident by construction: launch-latency bound, one CUDA graph per step"),
        clocks=clocks, e2e=dict(value=dof * 10 / t_e2e, unit=UNIT, h2d_bytes_per_step=nbytes // 10, d2h_bytes_per_step=q.size * 8 // 10, steps_per_call=10),
        gpu_launches=tm["launches"],
        roofline=dict(bound="hbm", achieved=alg / (ms_stage * 1e-3) / 1e9, peak=peak, unit="GB/s", frac=alg / (ms_stage * 1e-3) / 1e9 / peak,
                      traffic=None, kernel="advect_stage_kernel (halo kernel + stage kernel per RK stage, graph replay)", ms_per_launch=ms_stage,
                      algorithmic_bytes_per_launch=alg, peak_source=peak_src,
                      note="512 elements = 128 blocks < 148 SMs: the case cannot fill the device; the number documents launch latency"),
        cpu_baseline=cpu, finite=bool(np.isfinite(q2).all())))


def run_sphere(args, rank=0, world=1, local_rank=0):
    """configs[3]: global cubed sphere 6 x 32 x 32 x 12 elements p=7, GLOBALNONHYDRO3D_HEVI + IMEX_ARK324, the six panels as local
    meshes (fedg_group_update); synthetic state: balanced solid-body rotation + perturbations.  With --gpus 2 / 3 / 6 the panels
    are spread over the ranks (whole panels, the reference's rule) and the panel edges between ranks travel over NCCL: the
    sphere is fixed, so these lines are STRONG scaling."""
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        bind_to_gpu_numa_node(local_rank)
    dist = None
    bcast = None
    if world > 1:
        if world not in (2, 3, 4, 6, 8):
            if rank == 0:
                emit(dict(metric=METRIC, unavailable="the sphere is spread as whole panels (2, 3, 6 ranks) or as 2 x 2 tiles per panel "
                          "(4, 8 ranks)", n_gpus=world))
            return
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

        def bcast(raw):
            obj = [raw]
            dist.broadcast_object_list(obj, src=0)
            return obj[0]
    from fe_project_b200.cases import GlobalSphereCase
    from fe_project_b200.dyncore import PROG_NAMES
    # configs[3]: NeGX = NeGY = 32, NeZ = 12; configs[4] (--sphere-ne N): the same run with N elements per panel edge, sized by the caller
    # so that the state fills the HBM of the GPUs it runs on
    ne, nez = (args.sphere_ne or 32), (args.nez if args.nez != WORKLOAD["NeZ"] else 12)
    from fe_project_b200.cubedsphere import panel_owner
    ntile = args.sphere_ntile or (2 if world in (4, 8) else 1)          # 4 / 8 GPUs: 2 x 2 tiles per panel (24 local meshes), the same sphere
    own_ids = [t for t, r in enumerate(panel_owner(world, ntile)) if r == rank]
    # run.conf of test/case/baroclinic_wave_global: Jablonowski-Williamson state, lumped mass matrix, stretched FZ, eta_c = 0, sponge
    t_setup = time.perf_counter()
    case = GlobalSphereCase.config4(Ne=ne // ntile, NeZ=nez, ntile=ntile, fields_for=own_ids, init=args.sphere_init)
    g = case.make_driver(rank=rank, nranks=world, bcast=bcast)
    t_setup = time.perf_counter() - t_setup
    W, K = max(3, args.warmup), args.steps
    g.Update(W)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank); sampler.start(); sampler.wait_started()
    # barrier RIGHT before the timed steps: starting the clock sampler takes a different time on every rank, and a rank that enters the
    # region early spends the difference waiting in its first panel-edge exchange (measured: 20-40 ms on 8 GPUs, profiles/README.md)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    g.Update(K)
    torch.cuda.synchronize()
    tm = g.last_timing()
    clocks = sampler.stop()
    ms_total = tm["ms_total"]
    if dist:
        dist.barrier()
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    Np = case.elem.Np
    nel = sum(m.Ne for m in case.cs.panels)
    dof = 5 * Np * nel
    value = dof * K / (ms_total * 1e-3)
    own = [case.cs.panels[P] for P in g.panel_ids]
    states = [d.get_prog() for d in g.panels]
    finite = all(np.isfinite(st[k][:Np * m.Ne]).all() for st, m in zip(states, own) for k in PROG_NAMES)
    free_b, total_b = torch.cuda.mem_get_info()
    hbm_used_gb = (total_b - free_b) / 1e9
    # e2e: host state of the own panels in, one step, host state out
    ne2e = 1 if args.sphere_ne and args.sphere_ne > 48 else 3
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(ne2e):
        for d, st in zip(g.panels, states):
            d.set_prog(*(st[k] for k in PROG_NAMES))
        g.Update(1)
        states = [d.get_prog() for d in g.panels]
    t_e2e = (time.perf_counter() - t0) / ne2e
    nbytes = sum(5 * d.n_field * 8 for d in g.panels)
    if dist:
        t = torch.tensor([t_e2e, float(nbytes), 0.0 if finite else 1.0], device="cuda", dtype=torch.float64)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        t_e2e, nbytes, finite = float(tmax[0].item()), int(tsum[1].item()), tmax[2].item() == 0.0
    if rank == 0:
        emit(dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms_total / K, higher_is_better=True,
            scaling=("weak" if (world == 1 or args.sphere_ne) else "strong"), vs_baseline=None, dtype="f64", data="synthetic",
            config=dict(workload=("atm_nonhydro3d global baroclinic wave (Jablonowski-Williamson)" if args.sphere_init == "jw" else
                                  "atm_nonhydro3d global model, balanced solid-body rotation + perturbations (run.conf of baroclinic_wave_global)")
                                 + f" on the cubed sphere 6x{ne}x{ne}x{nez} elements p=7, "
                                 f"GLOBALNONHYDRO3D_HEVI, IMEX_ARK324, dt={case.dt}, lumped mass matrix, stretched FZ, modal filter eta_c=0, sponge layer, {len(own_ids)} {'panel' if ntile == 1 else 'tile (2x2 per panel)'}(s) per GPU as local meshes with linked halos"
                                 + (", panel edges between ranks over NCCL" if world > 1 else ""),
                        dof=dof, dof_per_gpu=dof // world, l2_policy="inputs larger than L2 (>= 50 MB per field and panel)", hbm_used_gb_rank0=round(hbm_used_gb, 1),
                        host_setup_s_rank0=round(t_setup, 1),
                        vi_kernel=os.environ.get("FEDG_VI_KERNEL", "2")),
            clocks=clocks, e2e=dict(value=dof / t_e2e, unit=UNIT, h2d_bytes_per_step=nbytes, d2h_bytes_per_step=nbytes, steps_per_call=1),
            gpu_launches=tm["launches"],
            roofline=dict(bound="fp64", achieved=None, peak=34.07, unit="TFLOP/s", frac=None, traffic=None,
                          kernel="vi_column_kernel (see the global_panel line for its per-launch figures)", ms_per_launch=None),
            cpu_baseline=None, finite=bool(finite)))
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nex", type=int, default=WORKLOAD["NeX"])
    ap.add_argument("--ney", type=int, default=WORKLOAD["NeY"])
    ap.add_argument("--nez", type=int, default=WORKLOAD["NeZ"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--numdiff", action="store_true", help="density_current as SHIPPED (run.conf: NUMDIFF_FLAG, ND_LAPLACIAN_NUM = 1, ND_COEF = 75): "
                    "numerical diffusion after every step, inside the timed region (extra line, not the headline)")
    ap.add_argument("--sphere-ntile", type=int, default=0, help="global_sphere: k x k tiles per panel (default: 1, or 2 on 4 / 8 GPUs)")
    ap.add_argument("--sphere-init", default="jw", choices=["jw", "solid_body"], help="global_sphere: initial state (configs[3] = jw; the configs[4] sizes use the cheap analytic one)")
    ap.add_argument("--sphere-ne", type=int, default=0, help="global_sphere: elements per panel edge (default 32 = configs[3]; configs[4] sizes it to the HBM)")
    ap.add_argument("--eqs", default="heve", choices=["heve", "hevi"], help="hevi: NONHYDRO3D_HEVI + IMEX_ARK324 (extra, not the headline)")
    ap.add_argument("--workload", default="density_current", choices=["density_current", "sound_wave", "global_panel", "global_sphere", "advect3d"],
                    help="density_current = BASELINE configs[2] (headline); the others are extra measurement lines: sound_wave = "
                         "configs[1] rate variant 16x16x16, global_panel = one 32x32x12 panel of configs[3], advect3d = configs[0]")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "advect3d":
        run_advect3d(args)
        return
    if args.workload == "global_sphere":
        run_sphere(args, rank, world, local_rank)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        print(f"[bench] rank {rank}: host CPUs {('%d-%d (%d)' % (numa_cpus[0], numa_cpus[-1], len(numa_cpus))) if numa_cpus else 'not bound'}", file=sys.stderr, flush=True)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from fe_project_b200.setup_aux import calc_phyd_hgrad
    from fe_project_b200.dyncore import PROG_NAMES, rk_tables

    W = max(3, args.warmup)
    K = args.steps
    # weak scaling: one 32x32x16 tile per GPU, NprcX x NprcY tiles (the reference's horizontal decomposition,
    # mod_atmos_mesh_rm.F90:104-106); the domain grows with the tile count
    NX, NY = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}.get(world, (world, 1))
    pi, pj = rank % NX, rank // NX
    hevi = args.eqs == "hevi" or args.workload != "density_current"
    wl_name = WORKLOAD["name"]
    if args.workload == "density_current":
        case = workload_case(args, hevi=hevi, tiles=(NX, NY), tile=(pi, pj))
    else:
        if world > 1:
            raise SystemExit("the extra workloads are single-GPU measurement lines")
        from fe_project_b200.cases import SoundWaveCase, GlobalPanelCase
        if args.workload == "sound_wave":     # configs[1]: sample/euler3d_hevi, rate variant 16x16x16
            args.nex = args.ney = args.nez = 16
            case = SoundWaveCase(p=7, NeX=16, NeY=16, NeZ=16, dt=0.015, tinteg="IMEX_ARK232", amplitude=1.0e-3)
            wl_name = "sample/euler3d_hevi sound wave (library HEVI path)"
        else:                                  # configs[3]: one panel of the 6x32x32x12 cubed sphere
            args.nex, args.ney, args.nez = 32, 32, 12
            case = GlobalPanelCase(p=7, NeX=32, NeY=32, NeZ=12, dt=5.0, tinteg="IMEX_ARK324", modalfilter=True)
            wl_name = "atm_nonhydro3d global, one cubed-sphere panel (lateral halo = own face values)"
    d = case.make_driver(None)
    if args.numdiff:
        if args.workload != "density_current":
            raise SystemExit("--numdiff belongs to the density_current workload")
        d.numdiff_init(ND_LAPLACIAN_NUM=1, ND_COEF_h=75.0, ND_COEF_v=75.0, apply_in_update=True)
        wl_name += " + numerical diffusion as shipped (ND_LAPLACIAN_NUM=1, ND_COEF=75)"
    if world > 1:
        def bcast(raw):
            obj = [raw]
            dist.broadcast_object_list(obj, src=0)
            return obj[0]
        d.init_comm(rank, world, bcast)
    if args.workload == "density_current":
        # DPhydDx / DPhydDy are registered on every rank count, so that the per-GPU work (two more fields per stage) is the same
        # at N = 1 and N > 1.  The background is horizontally uniform: the gradients vanish to round-off and do not depend on
        # the neighbours, so a tile evaluates them on a stand-alone copy of its own mesh (its lateral halo = own face values)
        # instead of exchanging PRES_hyd with the other ranks.
        if world == 1:
            gmesh = case.mesh
        else:
            from fe_project_b200.mesh import LocalMeshCube
            cm = case.mesh
            gmesh = LocalMeshCube(case.elem, cm.NeX, cm.NeY, cm.NeZ, cm.xmin, cm.xmax, cm.ymin, cm.ymax, cm.zmin, cm.zmax, FZ=cm.FZ,
                                  periodic=(False, False, False))
        gx, gy = calc_phyd_hgrad(case.elem, gmesh, case.fields["PRES_hyd"])
        d.set_phyd_hgrad(gx, gy)
    Np, Ne = case.elem.Np, case.mesh.Ne
    dof = 5 * Np * Ne * world
    nstage = rk_tables(case.tinteg)["nstage"]

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    def max_over_ranks(x):
        if not dist:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    d.Update(W)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_started()
    # The timed region is EXACTLY K steps between two barriers, CUDA events on the launching stream, max over ranks.  It is repeated
    # (R regions, every one timed the same way) until about a second of GPU time has passed, so that the clock sampler sees the
    # load; the reported region is the MEDIAN one.
    regions = []
    t_wall0 = time.perf_counter()
    f0 = [case.fields[k] for k in PROG_NAMES]
    while True:
        d.set_prog(*f0)                 # every region starts from the initial state (outside the timed region)
        barrier()
        d.Update(K)
        torch.cuda.synchronize()
        tm = d.last_timing()
        regions.append((max_over_ranks(tm["ms_total"]), tm["ms_stage_kernels"], tm["launches"]))
        stop = (time.perf_counter() - t_wall0 >= 1.0 and len(regions) >= 3) or len(regions) >= 200
        if dist:       # every rank takes the same decision
            t = torch.tensor([1.0 if stop else 0.0], device="cuda"); dist.broadcast(t, src=0); stop = bool(t.item() > 0.5)
        if stop:
            break
    clocks = sampler.stop()
    barrier()
    order = sorted(range(len(regions)), key=lambda i: regions[i][0])
    ms_total, ms_stage_sum, launches = regions[order[len(order) // 2]]
    value = dof * K / (ms_total * 1e-3)
    n_regions = K * nstage                      # one event-bracketed region per stage: the dominant kernel
    ms_stage = ms_stage_sum / max(1, n_regions)
    peak, peak_src = read_peaks()
    alg_bytes = ALG_BYTES_PER_NODE_STAGE * Np * Ne
    achieved = alg_bytes / (ms_stage * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "heve_stage_traffic.json")
    if os.path.exists(tpath) and not hevi and args.workload == "density_current" and (args.nex, args.ney, args.nez) == (32, 32, 16):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
        traffic_src = "ncu --set full capture of this kernel at this size, profiles/heve_stage_traffic.json (not measured in this run)"
    state = d.get_prog()
    finite = all(np.isfinite(state[k][:Np * Ne]).all() for k in PROG_NAMES)

    # ---- e2e: the reference-facing entry point with HOST buffers.  One dynamics step per call; pinned host arrays; the upload of the
    # step's five input fields and the download of its five output fields are inside the timed region.  The caller double-buffers
    # (three sets of host arrays = tiles / ensemble members / the physics side working on the other sets): call n + 1 is issued before
    # call n is waited for, so PCIe runs in both directions at once (fedg_dyn_update_host_async / _wait).  The blocking single call
    # (fedg_dyn_update_host) is reported next to it.
    nall = d.n_field
    NSLOT = 3
    sets = []
    for _ in range(NSLOT):
        pin = {k: torch.empty(nall, dtype=torch.float64).pin_memory() for k in PROG_NAMES}
        pout = {k: torch.empty(nall, dtype=torch.float64).pin_memory() for k in PROG_NAMES}
        hin, hout = {k: pin[k].numpy() for k in PROG_NAMES}, {k: pout[k].numpy() for k in PROG_NAMES}
        for k in PROG_NAMES:
            hin[k][:] = case.fields[k].reshape(-1)
        sets.append((pin, pout, hin, hout))
    d.Update_host(sets[0][2], 1)
    for s_ in range(NSLOT):                 # first use allocates the staging buffers and touches the pinned output arrays
        d.Update_host_async(sets[s_][2], sets[s_][3], 1, slot=s_)
    for s_ in range(NSLOT):
        d.Update_host_wait(s_)
    ncall = 18
    barrier()
    t0 = time.perf_counter()
    for i in range(ncall):
        s_ = i % NSLOT
        if i >= NSLOT:
            d.Update_host_wait(s_)
        d.Update_host_async(sets[s_][2], sets[s_][3], 1, slot=s_)
    for s_ in range(NSLOT):
        d.Update_host_wait(s_)
    torch.cuda.synchronize()
    t_pipe = max_over_ranks((time.perf_counter() - t0) / ncall)
    barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        d.Update_host(sets[0][2], 1)
    torch.cuda.synchronize()
    t_block = max_over_ranks((time.perf_counter() - t0) / 4)
    e2e_finite = all(np.isfinite(sets[s_][3][k][:Np * Ne]).all() for s_ in range(NSLOT) for k in PROG_NAMES)
    nbytes = 5 * Np * Ne * 8          # the (Np, Ne) interior of the five variables travels each way
    e2e = dict(value=dof / t_pipe, unit=UNIT, h2d_bytes_per_step=nbytes, d2h_bytes_per_step=nbytes, steps_per_call=1,
               mode="fedg_dyn_update_host_async/_wait, three host buffer sets in rotation (upload, step and download of consecutive calls overlap)",
               blocking_call_value=dof / t_block, ms_per_call=t_pipe * 1e3, ms_per_blocking_call=t_block * 1e3, finite=bool(e2e_finite),
               host_cpus_rank0=(f"{numa_cpus[0]}-{numa_cpus[-1]} ({len(numa_cpus)} CPUs local to the GPU, NVML)" if numa_cpus else "not bound"))

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            ccase = case if world == 1 and args.workload == "density_current" else workload_case(args, hevi=hevi)
            r = oracle_timed(ccase, warmup=1, target_s=12.0)
            cpu = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port", sample=r["sample"])
        # flops a vertical-implicit launch EXECUTES per column-element: two-lane block elimination (default) 16.7 k (ncu thread-instruction
        # counts of the implicit launch, gpurun_out/r02_vi2_raw.csv: 2 x DFMA + DADD + DMUL = 1.749e10 per launch at 32x32x16; it rebuilds
        # the rows from constant tables instead of loading a stored block), eight-lane kernel (FEDG_VI_KERNEL=1) 11.0 k
        vi_k2 = os.environ.get("FEDG_VI_KERNEL", "2") != "1"
        # ms_per_launch averages ALL vertical-implicit launches of a step: the implicit stages and the explicit evaluation of the
        # stages with a_im(s,s) = 0 (first stage of the ARK schemes: operator only, ~1.2 kflop per column-element, no solve), so the flop
        # counts are averaged over the same launches
        rk_t = rk_tables(case.tinteg) if hevi else None
        n_impl = sum(1 for s_ in range(nstage) if float(np.asarray(rk_t["a_im"]).reshape(nstage, nstage)[s_, s_]) != 0.0) if hevi else 0
        n_expl = nstage - n_impl
        avg = lambda impl: (n_impl * impl + n_expl * 1.2e3) / max(1, nstage)
        vi_flops_ref, vi_flops_exec = avg(15.1e3) * Ne * 64, avg(16.7e3 if vi_k2 else 11.0e3) * Ne * 64
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms_total / K,
            higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
            config=dict(workload=f"{wl_name}, {args.nex}x{args.ney}x{args.nez} elements per GPU p=7, {case.eqs}, "
                                 f"{case.tinteg}, dt={case.dt}, modal filter {'on' if case.modalfilter else 'off'}, tiles {NX}x{NY}",
                        dof=dof, l2_policy="inputs larger than L2 (67 MB per field, >1 GB touched per stage)",
                        specialisation="flat mesh (Gsqrt=1, GI3=0) and dry thermodynamics detected at registration; "
                                       "roofline uses the unspecialised 232 B/node/stage",
                        timed_regions=len(regions), region_ms=[round(regions[i][0], 4) for i in order[:1] + order[len(order) // 2:len(order) // 2 + 1] + order[-1:]],
                        halo=("direct peer memory (pack kernel stores into the neighbour's halo staging area over NVLink, flag per face)"
                              if world > 1 and os.environ.get("FEDG_HALO", "") != "nccl" else ("NCCL send/recv" if world > 1 else "none"))),
            clocks=clocks, e2e=e2e, gpu_launches=launches,
            roofline=(dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic, traffic_source=traffic_src,
                           kernel="stage_p7_kernel<flat,dry,HEVE> (DMMA contractions, TMA-staged inputs and z-face neighbours)", ms_per_launch=ms_stage,
                           algorithmic_bytes_per_launch=alg_bytes, peak_source=peak_src) if not hevi else
                      # vertical-implicit column solve: FP64 bound.  achieved / frac use the flops the kernel EXECUTES per column-element
                      # (see vi_flops_exec above); the reference
                      # algorithm's count (24x24 LU 9.2 k + substitutions 4.6 k + coupling 0.6 k + (u,v) 0.7 k = 15.1 kflop) is
                      # reported next to it.  ms_per_launch averages the implicit stages and the explicit-evaluation stage of the scheme
                      dict(bound="fp64", achieved=vi_flops_exec / (ms_stage * 1e-3) / 1e12, peak=34.07, unit="TFLOP/s",
                           frac=vi_flops_exec / (ms_stage * 1e-3) / 1e12 / 34.07, traffic=None,
                           kernel=("vi_column2_kernel (block-Thomas over the column; per element: density in closed form, theta block and Schur complement in w by partial-pivot Gauss-Jordan, two lanes per column)"
                                   if vi_k2 else "vi_column_kernel (block-Thomas over the column, DDENS eliminated, partial-pivot Gauss-Jordan per element, eight lanes per column)"), ms_per_launch=ms_stage,
                           executed_flops_per_launch=vi_flops_exec, reference_algorithm_flops_per_launch=vi_flops_ref,
                           launches_per_step=dict(implicit=n_impl, explicit_evaluation=n_expl),
                           frac_on_reference_algorithm_flops=vi_flops_ref / (ms_stage * 1e-3) / 1e12 / 34.07,
                           peak_source="measured DFMA peak, profiles/r01_fp64_peak.txt")),
            cpu_baseline=cpu, finite=bool(finite))
        emit(line)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
