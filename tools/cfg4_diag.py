"""Config-4 (JW sphere, shipped 6x8x8x4) own-norm errors against the oracle for each VI kernel / pow setting (diagnostic)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from cases import GlobalSphereCase
PROG = ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT")
case = GlobalSphereCase.config4(Ne=8, NeZ=4)
s = case.make_oracle(); s.update(5)
ref = {(P, nm): s.panels[P].arr(nm)[:s.panels[P].Ne * s.panels[P].Np].copy() for P in range(6) for nm in PROG}
for env in ({"FEDG_VI_KERNEL": "1"}, {"FEDG_VI_KERNEL": "2"}, {}, {"FEDG_VI_KERNEL": "1", "FEDG_EXACT_POW": "1"}, {"FEDG_VI_KERNEL": "2", "FEDG_EXACT_POW": "1"}):
    for k in ("FEDG_VI_KERNEL", "FEDG_EXACT_POW"): os.environ.pop(k, None)
    os.environ.update(env)
    g = case.make_driver(); g.Update(5)
    out = {}
    for nm in PROG:
        scale = max(np.abs(ref[(P, nm)]).max() for P in range(6))
        worst = 0.0
        for P, d in enumerate(g.panels):
            r = ref[(P, nm)]; a = d.get_prog()[nm][:r.size]
            worst = max(worst, np.linalg.norm(a - r) / max(np.linalg.norm(r), 1e-3 * scale * np.sqrt(r.size)))
        out[nm] = f"{worst:.2e}"
    print(env, out, flush=True)
    del g
