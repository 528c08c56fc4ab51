"""A/B timing of stage_p7 tuning knobs inside ONE process on ONE box (boxes differ by several per cent).
usage: python tools/ab_stage.py "name:ENV=val,ENV2=val" ...   (knobs are re-read by the library at every launch)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import fe_project_b200._lib as _L
if os.environ.get("AB_LIB"):
    _L.LIB_PATH = os.path.join(ROOT, "fe_project_b200", os.environ["AB_LIB"])
import bench
from fe_project_b200.cases import DensityCurrentCase
from fe_project_b200.dyncore import rk_tables

W = bench.WORKLOAD
hevi = os.environ.get("AB_EQS", "heve") == "hevi"
case = DensityCurrentCase(p=7, NeX=32, NeY=32, NeZ=16, dom=W["dom"], dt=(1.5 * W["dt"] if hevi else W["dt"]),
                          tinteg=("IMEX_ARK324" if hevi else W["tinteg"]), modalfilter=True,
                          eqs=("NONHYDRO3D_HEVI" if hevi else "NONHYDRO3D_HEVE"))
d = case.make_driver(None)
if os.environ.get("AB_PHYD", "1") == "1":   # as bench.py does on one tile: DPhydDx / DPhydDy registered (two more fields per stage)
    from fe_project_b200.setup_aux import calc_phyd_hgrad
    d.set_phyd_hgrad(*calc_phyd_hgrad(case.elem, case.mesh, case.fields["PRES_hyd"]))
nstage = rk_tables(case.tinteg)["nstage"]
K = int(os.environ.get("AB_STEPS", "40"))
variants = []
for a in sys.argv[1:]:
    name, _, envs = a.partition(":")
    variants.append((name, dict(kv.split("=") for kv in envs.split(",") if kv)))
keys = sorted({k for _, e in variants for k in e})
for rep in range(int(os.environ.get("AB_REPS", "2"))):
    for name, env in variants:
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(env)
        d.Update(3); torch.cuda.synchronize()
        d.Update(K); torch.cuda.synchronize()
        tm = d.last_timing()
        print(f"rep{rep} {name:14s} ms/step {tm['ms_total'] / K:.4f}  ms/stage-kernel {tm['ms_stage_kernels'] / (K * nstage):.4f}", flush=True)
