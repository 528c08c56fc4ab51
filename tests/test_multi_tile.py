"""Multi-tile halo exchange.  CPU part (world_size-2 gloo): the tile graph, the VMapB ordering and the tag-free
matching rule used by the NCCL path (sends in ascending own face id, receives in ascending face id of the sender)
reproduce the single-domain halo.  GPU part: launches tests/mgpu_parity.py under torchrun when two GPUs are visible."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, outq):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    from fe_project_b200.element import HexElement
    from fe_project_b200.mesh import LocalMeshCube
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        e = HexElement(2)
        NX, NY = 2, 1
        dom = (0.0, 4.0, 0.0, 3.0, 0.0, 2.0)
        per = (True, True, False)            # periodic x with two tiles: both lateral x-faces talk to the same peer
        tile = LocalMeshCube(e, 2, 3, 2, *dom, periodic=per, NprcX=NX, NprcY=NY, pi=rank, pj=0)
        glob = LocalMeshCube(e, 4, 3, 2, *dom, periodic=per)
        Np = e.Np

        def f(x, y, z):
            return 1.0 + x + 10.0 * y + 100.0 * z + 0.01 * x * y * z

        q = np.zeros(tile.NeA * Np)
        q[: tile.Ne * Np] = f(*tile.pos_en).reshape(-1)
        nint = tile.Ne * Np
        tile_rank = lambda qi, qj: qi + qj * NX
        # same-rank faces
        for fc in range(6):
            (qi, qj), fo = tile.tile_neighbors[fc]
            if tile_rank(qi, qj) == rank:
                o, oo, s = tile.halo_face_off[fc], tile.halo_face_off[fo], tile.halo_face_size[fc]
                q[nint + o: nint + o + s] = q[tile.VMapB[oo:oo + s]]
        remote = [fc for fc in range(6) if tile_rank(*tile.tile_neighbors[fc][0]) != rank]
        sends = sorted(remote)                                                      # ascending own face id
        recvs = sorted(remote, key=lambda fc: tile.tile_neighbors[fc][1])           # ascending face id of the sender
        reqs, bufs = [], []
        for fc in sends:
            o, s = tile.halo_face_off[fc], tile.halo_face_size[fc]
            t = torch.from_numpy(q[tile.VMapB[o:o + s]].copy())
            bufs.append(t)
            reqs.append(dist.isend(t, tile_rank(*tile.tile_neighbors[fc][0])))
        rb = []
        for fc in recvs:
            t = torch.empty(int(tile.halo_face_size[fc]), dtype=torch.float64)
            rb.append((fc, t))
            reqs.append(dist.irecv(t, tile_rank(*tile.tile_neighbors[fc][0])))
        for r in reqs:
            r.wait()
        for fc, t in rb:
            o, s = tile.halo_face_off[fc], tile.halo_face_size[fc]
            q[nint + o: nint + o + s] = t.numpy()
        # expected: value seen through VMapP in the single-domain mesh
        qg = np.zeros(glob.NeA * Np)
        qg[: glob.Ne * Np] = f(*glob.pos_en).reshape(-1)
        glob.exchange_halo_numpy(qg)
        ex, ey, ez = tile.ex + rank * 2, tile.ey, tile.ez
        ke_g = ex + ey * 4 + ez * 12
        got = q[tile.VMapP]                      # (Ne, NfpTot)
        exp = qg[glob.VMapP][ke_g]
        ok = np.array_equal(got, exp)
        # tile graph symmetry
        for fc in range(6):
            (qi, qj), fo = tile.tile_neighbors[fc]
            ok = ok and tile.halo_face_size[fc] == tile.halo_face_size[fo]
        q_ok = bool(ok)
    except Exception as exc:  # noqa: BLE001
        q_ok = repr(exc)
    outq.put((rank, q_ok))
    dist.destroy_process_group()


def test_two_rank_exchange_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok in res:
        assert ok is True, (rank, ok)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["heve", "hevi"])
def test_two_gpu_nccl_parity(mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run: gpurun --gpus 2 -- python -m pytest tests -m gpu -k nccl)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tests", "mgpu_parity.py")] + (["hevi"] if mode == "hevi" else [])
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "OK" in out.stdout


# ------------------------------------------------------------------------------ cubed sphere: panels spread over ranks
def _sphere_worker(rank, world, port, outq, ntile=1):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    from fe_project_b200.element import HexElement
    from fe_project_b200.cubedsphere import CubedSphere, exchange_plan, panel_owner
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        e = HexElement(2)
        cs = CubedSphere(e, 2, 2, 1.0e4, 6.37122e6, ntile=ntile)
        owner = panel_owner(world, ntile)
        nmesh = len(cs.panels)
        names = ("DDENS", "MOMX", "MOMY")
        rng = np.random.default_rng(7)                        # the same on every rank: the whole sphere, for the expectation
        full = [{n: rng.standard_normal(m.NeA * e.Np) for n in names} for m in cs.panels]
        mine = {P: {n: full[P][n].copy() for n in names} for P in range(nmesh) if owner[P] == rank}
        for P in mine:                                         # nothing of another rank's panels may be used below
            for n in names:
                mine[P][n][cs.panels[P].Ne * e.Np:] = np.nan
        local, recvs, sends = exchange_plan(cs.links, owner, rank)

        def put(U, g, vals, rot):
            m = cs.panels[U]
            nint, o = m.Ne * e.Np, m.halo_face_off[g]
            sl = slice(nint + o, nint + o + vals["DDENS"].size)
            mine[U]["DDENS"][sl] = vals["DDENS"]
            if rot is None:                                   # neighbour tile inside a panel: same basis
                mine[U]["MOMX"][sl], mine[U]["MOMY"][sl] = vals["MOMX"], vals["MOMY"]
            else:
                mine[U]["MOMX"][sl] = rot[:, 0, 0] * vals["MOMX"] + rot[:, 0, 1] * vals["MOMY"]
                mine[U]["MOMY"][sl] = rot[:, 1, 0] * vals["MOMX"] + rot[:, 1, 1] * vals["MOMY"]

        for U, g, T in local:
            _, src, rot = cs.links[U][g]
            put(U, g, {n: mine[T][n][src] for n in names}, rot)
        # tag-free matching as NCCL does it: per peer, messages in posting order; both sides post in ascending msg_id
        reqs, keep, rb = [], [], []
        for T, peer, mid, U, g in sends:
            src = cs.links[U][g][1]
            t = torch.from_numpy(np.concatenate([mine[T][n][src] for n in names]))
            keep.append(t)
            reqs.append(dist.isend(t, peer))
        for U, g, peer, mid in recvs:
            t = torch.empty(3 * cs.links[U][g][1].size, dtype=torch.float64)
            rb.append((U, g, t))
            reqs.append(dist.irecv(t, peer))
        for r in reqs:
            r.wait()
        for U, g, t in rb:
            a = t.numpy().reshape(3, -1)
            put(U, g, dict(zip(names, a)), cs.links[U][g][2])
        cs.exchange_numpy(full)
        ok = True
        for P in mine:
            m = cs.panels[P]
            lat = slice(m.Ne * e.Np, m.Ne * e.Np + m.halo_face_off[4])       # the four lateral halo faces
            for n in names:
                ok = ok and np.array_equal(mine[P][n][lat], full[P][n][lat])
        # every linked face is covered exactly once
        ok = ok and len(local) + len(recvs) == 4 * len(mine)
        res = bool(ok)
    except Exception as exc:  # noqa: BLE001
        res = repr(exc)
    outq.put((rank, res))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,ntile", [(2, 1), (3, 1), (4, 2)])
def test_sphere_panel_exchange_plan_gloo(world, ntile):
    """Panels spread over 2 / 3 ranks, 2 x 2 tiles per panel over 4 ranks: local links + the send / receive plan (ascending
    msg_id per peer, no tags) reproduce the single-process exchange bit for bit."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + (os.getpid() + 7 * world) % 2000
    procs = [ctx.Process(target=_sphere_worker, args=(r, world, port, q, ntile)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok in res:
        assert ok is True, (rank, ok)


def test_panel_owner_follows_the_reference_rule():
    from fe_project_b200.cubedsphere import panel_owner
    assert panel_owner(1) == [0] * 6 and panel_owner(2) == [0, 0, 0, 1, 1, 1] and panel_owner(6) == list(range(6))
    with pytest.raises(ValueError):
        panel_owner(4)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["hevi", "heve"])
def test_two_gpu_sphere_parity(mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run: gpurun --gpus 2 -- python -m pytest tests -m gpu -k sphere_parity)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 300), os.path.join(ROOT, "tests", "mgpu_sphere_parity.py")] + (["heve"] if mode == "heve" else [])
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "OK" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["hevi", "heve"])
def test_sphere_sub_panel_tiles_one_gpu(mode):
    """24 local meshes (2 x 2 tiles per panel) advanced together on one GPU == the whole-panel oracle."""
    cmd = [sys.executable, os.path.join(ROOT, "tests", "mgpu_sphere_parity.py"), "tiles"] + (["heve"] if mode == "heve" else [])
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "OK" in out.stdout


@pytest.mark.gpu
def test_two_gpu_sphere_tiles_parity():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29800 + os.getpid() % 150), os.path.join(ROOT, "tests", "mgpu_sphere_parity.py"), "tiles"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "OK" in out.stdout
