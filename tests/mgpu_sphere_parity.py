"""Whole cubed sphere over several GPUs (run under torchrun, one rank per GPU): every rank owns a block of local meshes, the
edges between meshes of different ranks travel over NCCL (fedg_link_halo_send / _recv).  Whole panels: 2 / 3 / 6 ranks;
with `tiles` every panel is cut into 2 x 2 tiles (24 local meshes): 1 / 2 / 4 / 8 ranks.  Checked against the single-process CPU
oracle of the whole sphere (whole panels): halo contents after one exchange (whole panels only), then N steps.  Rank 0 prints one
line and exits non-zero on failure.
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tests/mgpu_sphere_parity.py [heve] [tiles]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
PROG = ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT")


def run_tiles(kw, rank, world, dist, torch, jw=False):
    """2 x 2 tiles per panel on the device(s) against the whole-panel oracle.  jw: BASELINE config 4 in small (Jablonowski-Williamson
    state, lumped mass matrix, stretched FZ, eta_c = 0, sponge): the background fields and the hydrostatic pressure gradient cross the
    tile and rank boundaries too (fedg_group_exchange_aux, fedg_update_phyd_hgrad)."""
    from cases import GlobalSphereCase, rel_l2
    if jw:
        case = GlobalSphereCase.config4(Ne=2, NeZ=4, ntile=2)
        ref = GlobalSphereCase.config4(Ne=4, NeZ=4, ntile=1)
        assert case.dt == ref.dt
    else:
        case = GlobalSphereCase(p=7, Ne=2, NeZ=2, ntile=2, **kw)          # 24 tiles of 2 x 2 x 2 elements
        ref = GlobalSphereCase(p=7, Ne=4, NeZ=2, ntile=1, **kw)           # the same sphere as six whole panels

    def bcast(raw):
        obj = [raw]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]
    g = case.make_driver(rank=rank, nranks=world, bcast=bcast if world > 1 else None)
    s = ref.make_oracle()
    nsteps = 5
    g.Update(nsteps)
    s.update(nsteps)
    Np = case.elem.Np
    worst = 0.0
    for d, t in zip(g.panels, g.panel_ids):
        m = case.cs.panels[t]
        k, ti, tj = m.sub
        nxp = k * m.NeX
        kep = (m.ex + ti * m.NeX) + (m.ey + tj * m.NeY) * nxp + m.ez * nxp * (k * m.NeY)
        o = s.panels[case.cs.panel_of[t]]
        got = d.get_prog()
        for nm in PROG:
            exp = o.arr(nm)[: ref.cs.panels[0].Ne * Np].reshape(-1, Np)[kep].reshape(-1)
            if jw and nm in ("DDENS", "DRHOT", "MOMZ"):
                # near-zero perturbations of a balanced state: judged against the full-field scale (tests/test_gpu_config_sizes.py)
                full = {"DDENS": 1.2, "DRHOT": 1.2 * 300.0, "MOMZ": 30.0}[nm]
                worst = max(worst, float(np.linalg.norm(got[nm][: m.Ne * Np] - exp) / (full * np.sqrt(exp.size))))
            else:
                worst = max(worst, rel_l2(got[nm][: m.Ne * Np], exp))
    if world > 1:
        t = torch.tensor([worst], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst = t.item()
    ok = worst <= 1e-10
    if rank == 0:
        print(f"mgpu_sphere_parity {'config4-JW ' if jw else ''}tiles=2x2 per panel ({24 // world} local meshes per rank) world={world} eqs={case.eqs} {case.tinteg}: "
              f"worst rel L2 after {nsteps} steps = {worst:.3e} -> {'OK' if ok else 'FAIL'}")
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


def main():
    import torch
    import torch.distributed as dist
    from cases import GlobalSphereCase, rel_l2
    rank, world, lrank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lrank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    heve = "heve" in sys.argv[1:]
    tiles = "tiles" in sys.argv[1:]
    kw = dict(eqs="GLOBALNONHYDRO3D_HEVE", tinteg="ERK_SSP_4s3o", dt=2.0) if heve else dict(tinteg="IMEX_ARK324", dt=20.0)
    if tiles:
        return run_tiles(kw, rank, world, dist, torch, jw="jw" in sys.argv[1:])
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
