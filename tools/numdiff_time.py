"""Time of AtmDyn_Nonhydro3D_Numdiff%Apply (fedg_numdiff_apply) on the bench tile, shipped coefficients (ND_LAPLACIAN_NUM = 1, ND_COEF = 75)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fe_project_b200._lib as _L
if os.environ.get("AB_LIB"):
    _L.LIB_PATH = os.path.join(ROOT, "fe_project_b200", os.environ["AB_LIB"])
import bench
from fe_project_b200.cases import DensityCurrentCase
W = bench.WORKLOAD
case = DensityCurrentCase(p=7, NeX=32, NeY=32, NeZ=16, dom=W["dom"], dt=W["dt"], tinteg=W["tinteg"], modalfilter=True)
d = case.make_driver(None)
d.numdiff_init(ND_LAPLACIAN_NUM=int(os.environ.get("ND_LAP", "1")), ND_COEF_h=75.0, ND_COEF_v=75.0, apply_in_update=False)
for _ in range(3):
    d.numdiff_apply()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
t0 = time.perf_counter()
N = 20
for _ in range(N):
    d.numdiff_apply()
torch.cuda.synchronize()
print(f"FEDG_ND_KERNEL={os.environ.get('FEDG_ND_KERNEL', 'default(p7 tensor-core)')} ND_LAP={os.environ.get('ND_LAP', '1')} numdiff_apply: {(time.perf_counter() - t0) / N * 1e3:.3f} ms per call (5 variables x {2 * int(os.environ.get('ND_LAP', '1'))} launches... wall clock around synchronous calls)")
