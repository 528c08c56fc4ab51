"""Whole cubed sphere over several GPUs (run under torchrun, 2 / 3 / 6 ranks, one per GPU): every rank owns a block of panels,
panel edges between ranks travel over NCCL (fedg_link_halo_send / _recv).  Checked against the single-process CPU oracle of the
whole sphere: halo contents after one exchange, then N steps.  Rank 0 prints one line and exits non-zero on failure.
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tests/mgpu_sphere_parity.py [heve]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
PROG = ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT")


def main():
    import torch
    import torch.distributed as dist
    from cases import GlobalSphereCase, rel_l2
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    heve = "heve" in sys.argv[1:]
    kw = dict(eqs="GLOBALNONHYDRO3D_HEVE", tinteg="ERK_SSP_4s3o", dt=2.0) if heve else dict(tinteg="IMEX_ARK324", dt=20.0)
    case = GlobalSphereCase(p=7, Ne=2, NeZ=3, **kw)

    def bcast(raw):
        obj = [raw]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]
    g = case.make_driver(rank=rank, nranks=world, bcast=bcast)
    s = case.make_oracle()                               # the whole sphere on every rank: small
    Np = case.elem.Np

    # 1. halo contents after one exchange
    for o in s.panels:
        o.piece("pressure")
    s.exchange(with_dpres=True)
    for d in g.panels:
        d.get_pres()
    g.exchange_halo(apply_bc=False)
    halo_err = 0.0
    for d, P in zip(g.panels, g.panel_ids):
        m, o = case.cs.panels[P], s.panels[P]
        got = d.get_prog()
        n = m.Ne * Np
        for nm in PROG:
            ref = o.arr(nm)
            halo_err = max(halo_err, float(np.abs(got[nm][n:n + m.Nhalo] - ref[n:n + m.Nhalo]).max() / np.abs(ref[:n]).max()))

    # 2. N steps
    nsteps = 5
    g.Update(nsteps)
    s.update(nsteps)
    worst = 0.0
    for d, P in zip(g.panels, g.panel_ids):
        m, o = case.cs.panels[P], s.panels[P]
        got = d.get_prog()
        n = m.Ne * Np
        for nm in PROG:
            worst = max(worst, rel_l2(got[nm][:n], o.arr(nm)[:n]))
    t = torch.tensor([halo_err, worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    halo_err, worst = t.tolist()
    ok = halo_err <= 1e-13 and worst <= 1e-10
    if rank == 0:
        print(f"mgpu_sphere_parity world={world} eqs={case.eqs} {case.tinteg}: halo max err {halo_err:.3e}, worst rel L2 after "
              f"{nsteps} steps = {worst:.3e} -> {'OK' if ok else 'FAIL'}")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
