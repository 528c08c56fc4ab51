"""Generates tests/golden/config4_jw_6x32x32x12.npz: the CPU oracle's state of BASELINE configs[3] at its full size (global
Jablonowski-Williamson baroclinic wave, 6 x 32 x 32 x 12 elements, p = 7, GLOBALNONHYDRO3D_HEVI, IMEX_ARK324, lumped mass matrix,
stretched FZ, modal filter with eta_c = 0, sponge layer) after NSTEPS steps, at every STRIDE-th node of every panel and variable,
plus the L2 norms of the full fields.  The oracle needs ~45 GB and ~70 s per step on 8 cores at this size, so the GPU test compares
against this fixture instead of running it on the GPU box (tests/test_gpu_config_sizes.py).  ORACLE OUTPUT, not reference output: the
Fortran reference cannot be built in this image.
A second run of the same oracle from an initial state whose MOMX / MOMY are moved by at most one unit in the last place (u_P_name) measures
how far round-off alone carries each variable: DDENS, DRHOT and MOMZ of this balanced state are near-zero perturbations and the test judges
them against that measured sensitivity instead of pretending they are determined to 1e-10 of their own norm."""
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
    del o
    # the oracle's own sensitivity to round-off: same run, MOMX / MOMY of the initial state moved by -1 / 0 / +1 ulp
    o = case.make_oracle()
    rng = np.random.default_rng(1)
    for pn in o.panels:
        for nm in ("MOMX", "MOMY"):
            a = pn.arr(nm)
            a *= 1.0 + 2.2e-16 * rng.integers(-1, 2, a.shape)
    o.update(NSTEPS)
    print("perturbed steps", time.time() - t0, flush=True)
    for P, pn in enumerate(o.panels):
        n = pn.Ne * pn.Np
        for nm in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT"):
            out[f"u_{P}_{nm}"] = pn.arr(nm)[:n][::STRIDE].copy()
    np.savez_compressed(os.path.join(HERE, f"config4_jw_6x{ne}x{ne}x{NEZ}.npz"), **out)
    print("done", time.time() - t0)


if __name__ == "__main__":
    main()
