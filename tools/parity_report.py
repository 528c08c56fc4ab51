"""Prints the relative L2 error of every prognostic variable vs the CPU oracle after N steps (GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from cases import DensityCurrentCase, rel_l2

for (p, dims, nstep) in ((7, (4, 2, 3), 20), (7, (6, 2, 4), 100), (3, (6, 4, 4), 50)):
    case = DensityCurrentCase(p=p, NeX=dims[0], NeY=dims[1], NeZ=dims[2], perturb=2.0, intrp_order=min(11, p + 4),
                              dt=0.08 if p == 7 else 0.2)
    o = case.make_oracle(); d = case.make_driver(o)
    o.update(nstep); d.Update(nstep)
    g = d.get_prog(); n = case.mesh.Ne * case.elem.Np
    print(f"p={p} {dims} N={nstep}:", " ".join(f"{nm}={rel_l2(g[nm][:n], o.arr(nm)[:n]):.2e}" for nm in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT")))
