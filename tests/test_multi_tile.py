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
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
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
