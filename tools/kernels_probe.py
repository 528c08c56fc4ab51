"""Launches every kernel of the library once or twice on the bench mesh (32x32x16, p = 7) for an ncu pass over the kernels other
than the two dominant ones: IMEX combinations, combination + filter, numerical diffusion, halo fill, pressure, monitors, tracer
advection, hydrostatic pressure gradient.  Run under  ncu --metrics ...  (tools/gpu_r2_small_kernels.sh)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from fe_project_b200.cases import DensityCurrentCase

dom = (0.0, 25.6e3, 0.0, 25.6e3, 0.0, 6.4e3)
ADIA = {k: "ADIABAT" for k in ("south", "east", "north", "west", "btm", "top")}
# HEVI step: vi (explicit + implicit), stage kernel, lincomb, lincomb_filter
case = DensityCurrentCase(p=7, NeX=32, NeY=32, NeZ=16, dom=dom, dt=0.06, eqs="NONHYDRO3D_HEVI", tinteg="IMEX_ARK324", modalfilter=True)
d = case.make_driver(None)
d.Update(2)
d.monitor(); d.get_pres(); d.update_phyd_hgrad(); d.exchange_halo(True)
del d
# HEVE step with the shipped numerical diffusion + tracer advection on the same state
case = DensityCurrentCase(p=7, NeX=32, NeY=32, NeZ=16, dom=dom, dt=0.04, modalfilter=True)
d = case.make_driver(None)
d.numdiff_init(2, 75.0 * 300.0 ** 2, 75.0 * 300.0 ** 2, therm_bc=ADIA, apply_in_update=True)
d.Update(2)
d.modalfilter_apply()
d.trcadv_init("ERK_SSP_3s3o", 0.04, MODALFILTER_FLAG=True)
q = np.zeros(d.n_field); q[: d.n_int] = 1.0e-3
d.trcadv_update(q, 1)
torch.cuda.synchronize()
print("probe done")
