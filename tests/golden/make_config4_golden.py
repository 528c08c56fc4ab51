"""Generates tests/golden/config4_jw_6x32x32x12.npz: the CPU oracle's state of BASELINE configs[3] at its full size (global
Jablonowski-Williamson baroclinic wave, 6 x 32 x 32 x 12 elements, p = 7, GLOBALNONHYDRO3D_HEVI, IMEX_ARK324, lumped mass matrix,
stretched FZ, modal filter with eta_c = 0, sponge layer) after NSTEPS steps, at every STRIDE-th node of every panel and variable,
plus the L2 norms of the full fields.  The oracle needs ~45 GB and ~70 s per step on 8 cores at this size, so the GPU test compares
against this fixture instead of running it on the GPU box (tests/test_gpu_config_sizes.py).  ORACLE OUTPUT, not reference output: the
Fortran reference cannot be built in this image."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

NE, NEZ, NSTEPS, STRIDE = 32, 12, 2, 4099


def main():
    from cases import GlobalSphereCase
    import oracle_api
    ne = int(sys.argv[1]) if len(sys.argv) > 1 else NE
    oracle_api.lib().feo_set_num_threads(os.cpu_count())
    t0 = time.time()
    case = GlobalSphereCase.config4(Ne=ne, NeZ=NEZ)
    o = case.make_oracle()
    print("set-up", time.time() - t0, flush=True)
    o.update(NSTEPS)
    print("steps", time.time() - t0, flush=True)
    out = dict(ne=ne, nez=NEZ, nsteps=NSTEPS, stride=STRIDE, dt=case.dt)
    for P, pn in enumerate(o.panels):
        n = pn.Ne * pn.Np
        for nm in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT"):
            a = pn.arr(nm)[:n]
            out[f"s_{P}_{nm}"] = a[::STRIDE].copy()
            out[f"n_{P}_{nm}"] = float(np.linalg.norm(a))
    np.savez_compressed(os.path.join(HERE, f"config4_jw_6x{ne}x{ne}x{NEZ}.npz"), **out)
    print("done", time.time() - t0)


if __name__ == "__main__":
    main()
