"""CPU experiment on the oracle: elimination of the vertical-implicit blocks in a fixed order without pivot search
(FEO_VI_STATIC_PIVOT=1, oracle/dyn_hevi.cpp) against the reference's partial pivoting.  Run twice and compare:
  FEO_VI_STATIC_PIVOT=0 python tools/vi_static_pivot_experiment.py piv; FEO_VI_STATIC_PIVOT=1 python tools/vi_static_pivot_experiment.py sta
Result of round 1 (5 steps, worst relative L2 over the variables): density current CFLv~1.5: 3.7e-7, CFLv~15: 3.5e-12,
sound wave dt=10 s: 1.4e-1, global panel dt=20 s: 4.8e-13, dt=75 s: 1.3e-11 -> a search-free device solver is NOT an option."""
import os, sys, subprocess, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
mode=sys.argv[1]
from cases import DensityCurrentCase, SoundWaveCase, GlobalPanelCase
out={}
cases={
 "density_current CFLv~1.5": DensityCurrentCase(p=7, NeX=2, NeY=2, NeZ=6, perturb=2.0, eqs="NONHYDRO3D_HEVI", tinteg="IMEX_ARK324", dt=1.0),
 "density_current CFLv~15": DensityCurrentCase(p=7, NeX=2, NeY=2, NeZ=6, perturb=2.0, eqs="NONHYDRO3D_HEVI", tinteg="IMEX_ARK232", dt=0.6, dom=(0.0,25.6e3,0.0,12.8e3,0.0,640.0)),
 "sound_wave dt=10s (CFLv~30)": SoundWaveCase(p=7, NeX=1, NeY=1, NeZ=80, dt=10.0, tinteg="IMEX_ARK232", amplitude=1.0e-3),
 "global_panel dt=20s": GlobalPanelCase(p=7, NeX=2, NeY=2, NeZ=3, dt=20.0),
 "global_panel dt=75s NeZ=12": GlobalPanelCase(p=7, NeX=2, NeY=2, NeZ=12, dt=75.0),
}
for k,c in cases.items():
    o=c.make_oracle(); o.update(5)
    out[k]={nm: o.arr(nm)[:c.mesh.Ne*c.elem.Np].tolist() for nm in ("DDENS","MOMZ","DRHOT","MOMX")}
np.save(f"/tmp/sp_{mode}.npy", out, allow_pickle=True)
