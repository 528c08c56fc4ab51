"""Multi-GPU parity (run under torchrun, one rank per GPU): NprcX x NprcY tiles with NCCL halo exchange against the
single-domain CPU oracle on the same global mesh.  Rank 0 prints one line per variable and exits non-zero on failure.
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_parity.py [hevi] [numdiff]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    import torch.distributed as dist
    from cases import DensityCurrentCase, rel_l2
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    hevi = "hevi" in sys.argv[1:]
    numdiff = "numdiff" in sys.argv[1:] or "numdiff1" in sys.argv[1:]   # numerical diffusion after every step, adiabatic walls
    nd_lap = 1 if "numdiff1" in sys.argv[1:] else 2    # numdiff1: one Laplacian (the shipped setting; on p = 7 the five variables of a half-step share a launch)
    NX, NY = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}[world]
    pi, pj = rank % NX, rank // NX
    # HEVI: dt keeps the horizontally explicit part stable (acoustic CFL ~0.4) while the vertical CFL is ~1.5
    nex, ney, nez = (3, 2, 6) if hevi else (3, 2, 3)
    dom = (0.0, 25.6e3, 0.0, 12.8e3, 0.0, 6.4e3)
    kw = dict(p=7, dom=dom, perturb=2.0, dt=0.15 if hevi else 0.05, periodic=(False, True, False))
    if hevi:
        kw.update(eqs="NONHYDRO3D_HEVI", tinteg="IMEX_ARK232")
    tile = DensityCurrentCase(NeX=nex, NeY=ney, NeZ=nez, NprcX=NX, NprcY=NY, pi=pi, pj=pj, **kw)
    d = tile.make_driver(None)

    def bcast(raw):
        obj = [raw]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]
    d.init_comm(rank, world, bcast)
    if numdiff:
        d.numdiff_init(nd_lap, 75.0 * (300.0 ** 2 if nd_lap == 2 else 1.0), 75.0 * (300.0 ** 2 if nd_lap == 2 else 1.0), therm_bc={k: "ADIABAT" for k in ("south", "east", "north", "west", "btm", "top")},
                       apply_in_update=True)
    nsteps = 10
    d.Update(nsteps)
    g = d.get_prog()
    mon = d.monitor()
    t = torch.tensor(mon, device="cuda", dtype=torch.float64)
    dist.all_reduce(t)
    mon_g = t.cpu().numpy()

    # reference: the whole domain on the CPU oracle (every rank computes it: small)
    glob = DensityCurrentCase(NeX=nex * NX, NeY=ney * NY, NeZ=nez, **kw)
    o = glob.make_oracle()
    if numdiff:
        o.set_numdiff(True, nd_lap, 75.0 * (300.0 ** 2 if nd_lap == 2 else 1.0), 75.0 * (300.0 ** 2 if nd_lap == 2 else 1.0), therm_bc=(1,) * 6)
    o.update(nsteps)
    Np = tile.elem.Np
    # global element index of each tile element
    ex, ey, ez = tile.mesh.ex + pi * nex, tile.mesh.ey + pj * ney, tile.mesh.ez
    ke_g = ex + ey * (nex * NX) + ez * (nex * NX) * (ney * NY)
    worst = 0.0
    for nm in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT"):
        ref = o.arr(nm)[: glob.mesh.Ne * Np].reshape(-1, Np)[ke_g].reshape(-1)
        err = rel_l2(g[nm][: tile.mesh.Ne * Np], ref)
        worst = max(worst, err)
    errs = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    mo = o.monitor()
    ok = errs.item() <= 1e-10 and abs(mon_g[1] - mo[1]) <= 1e-12 * abs(mo[1])
    if rank == 0:
        print(f"mgpu_parity world={world} tiles={NX}x{NY} hevi={hevi} numdiff={numdiff}: worst rel L2 = {errs.item():.3e}; "
              f"ENGT tiles {mon_g[1]:.15e} oracle {mo[1]:.15e} -> {'OK' if ok else 'FAIL'}")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
