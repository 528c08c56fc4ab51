"""Oracle pins, mesh maps and halo exchange.

* The reference's halo test restated (FElib/test/FE/field_cubedom3d_hexahedral/test_field_cubedom3d_hexahedral.f90:
  3x3x3 elements, p=2, triply periodic, field value = tileID*1e6 + ke*1e3 + p; halo contents per face, :261-420).
* Cross-check of the two independent mesh restatements (C++ nearest-node search as in MeshUtil3D_BuildInteriorMap vs
  the closed-form NumPy maps)."""
import numpy as np
import pytest

from fe_project_b200.element import HexElement
from fe_project_b200.mesh import LocalMeshCube
from oracle_api import Oracle


def _val(tile, ke, p):  # get_field_val_func, 1-based ke and p
    return tile * 1000000 + ke * 1000 + p


def test_reference_halo_pattern():
    p, N = 2, 3
    o = Oracle(p, N, N, N, (-1, 1, -1, 1, -1, 1), periodic=(True, True, True), lumped=True)
    Np, Ne, Nfp = o.Np, o.Ne, (p + 1) ** 2
    q = o.arr("DDENS")
    ke, pp = np.meshgrid(np.arange(1, Ne + 1), np.arange(1, Np + 1), indexing="ij")
    q[:Ne * Np] = _val(1, ke, pp).reshape(-1)
    o.piece("exchange")
    Fmask = o.iarr("Fmask").reshape(6, Nfp) + 1
    halo = q[Ne * Np:Ne * Np + o.Nhalo].astype(np.int64)
    off = 0
    # south halo <- y+ face (Fmask_h(:,3)) of the northern row; east <- x- face (4) of the first column; ...
    def expect(face_elems, fm):
        return np.array([[_val(1, e, f) for f in Fmask[fm]] for e in face_elems]).reshape(-1)
    south = [i + (N - 1) * N + (k - 1) * N * N for k in range(1, N + 1) for i in range(1, N + 1)]
    east = [1 + (j - 1) * N + (k - 1) * N * N for k in range(1, N + 1) for j in range(1, N + 1)]
    north = [i + (k - 1) * N * N for k in range(1, N + 1) for i in range(1, N + 1)]
    west = [N + (j - 1) * N + (k - 1) * N * N for k in range(1, N + 1) for j in range(1, N + 1)]
    bottom = [i + (j - 1) * N + (N - 1) * N * N for j in range(1, N + 1) for i in range(1, N + 1)]
    top = [i + (j - 1) * N for j in range(1, N + 1) for i in range(1, N + 1)]
    for elems, fm in ((south, 2), (east, 3), (north, 0), (west, 1), (bottom, 5), (top, 4)):
        n = len(elems) * Nfp
        assert np.array_equal(halo[off:off + n], expect(elems, fm))
        off += n
    assert off == o.Nhalo
    # interior untouched
    assert np.array_equal(q[:Ne * Np], _val(1, ke, pp).reshape(-1))


@pytest.mark.parametrize("p,dims,per", [(2, (3, 3, 3), (True, True, True)), (3, (4, 2, 3), (False, True, False)),
                                        (7, (2, 3, 2), (False, False, False))])
def test_two_mesh_restatements_agree(p, dims, per):
    dom = (0.0, 4.0, -1.0, 2.0, 0.0, 3.0)
    o = Oracle(p, *dims, dom, periodic=per)
    e = HexElement(p)
    m = LocalMeshCube(e, *dims, *dom, periodic=per)
    assert (o.Ne, o.NeA, o.Nhalo, o.Ne2D) == (m.Ne, m.NeA, m.Nhalo, m.Ne2D)
    assert np.array_equal(o.iarr("vmapM"), m.VMapM.reshape(-1))
    assert np.array_equal(o.iarr("vmapP"), m.VMapP.reshape(-1))
    assert np.array_equal(o.iarr("vmapB"), m.VMapB)
    assert np.array_equal(o.iarr("emap2d"), m.EMap3Dto2D)
    for d, nm in enumerate(("pos_x", "pos_y", "pos_z")):
        assert np.abs(o.arr(nm) - m.pos_en[d].reshape(-1)).max() < 1e-13
    for nm, a in (("E11", m.Escale[0, 0]), ("E22", m.Escale[1, 1]), ("E33", m.Escale[2, 2]), ("J", m.J),
                  ("Fscale", m.Fscale), ("nx", m.normal_fn[0]), ("ny", m.normal_fn[1]), ("nz", m.normal_fn[2])):
        assert np.allclose(o.arr(nm), a.reshape(-1), rtol=1e-14, atol=0), nm
    # same-rank exchange agrees
    rng = np.random.default_rng(1)
    q = np.zeros(m.NeA * e.Np); q[:m.Ne * e.Np] = rng.standard_normal(m.Ne * e.Np)
    o.arr("MOMX")[:] = q
    o.piece("exchange")
    m.exchange_halo_numpy(q)
    assert np.array_equal(o.arr("MOMX"), q)


def test_vmap_geometry():
    """Every interior face node pairs with a node at the same position; boundary nodes point into the halo."""
    p = 3
    e = HexElement(p)
    m = LocalMeshCube(e, 3, 2, 2, 0, 3, 0, 2, 0, 2, periodic=(False, False, False))
    pos = m.pos_en.reshape(3, -1)
    nint = m.Ne * e.Np
    iM, iP = m.VMapM.reshape(-1), m.VMapP.reshape(-1)
    inner = iP < nint
    assert np.abs(pos[:, iM[inner]] - pos[:, iP[inner]]).max() < 1e-13
    halo = iP[~inner] - nint
    assert np.array_equal(np.sort(halo), np.arange(m.Nhalo))
    assert np.array_equal(m.VMapB[halo], iM[~inner])


def test_multi_tile_graph():
    """MeshUtil3D_buildGlobalMap (scale_meshutil_3d.F90:750-877): 2x2 tiles, periodic in y only."""
    e = HexElement(1)
    tiles = {(i, j): LocalMeshCube(e, 2, 2, 1, 0, 4, 0, 4, 0, 1, periodic=(False, True, False), NprcX=2, NprcY=2, pi=i, pj=j)
             for i in range(2) for j in range(2)}
    t = tiles[(0, 0)]
    assert t.tile_neighbors[0] == ((0, 1), 2)   # south wraps (periodic y)
    assert t.tile_neighbors[1] == ((1, 0), 3)   # east neighbour, its west face
    assert t.tile_neighbors[3] == ((0, 0), 3)   # west: domain boundary -> itself, same face
    assert t.tile_neighbors[4] == ((0, 0), 4) and t.tile_neighbors[5] == ((0, 0), 5)
    assert np.isclose(tiles[(1, 1)].xmin, 2.0) and np.isclose(tiles[(1, 1)].ymax, 4.0)
