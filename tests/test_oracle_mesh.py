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


def reference_vmaps(n1, NeX, NeY, NeZ):
    """VMapM_h/v_ans, VMapP_h/v_ans of check_connectivity (test_mesh_cubedom3d_hexahedral.f90:120-190, identical in
    test_mesh_cubedsphere3d.f90:110-180), 1-based, faces in the order south, east, north, west, bottom, top."""
    Np, Nfp, Ne = n1 ** 3, n1 ** 2, NeX * NeY * NeZ

    def elem_id(i, j, k):
        return i + (j - 1) * NeX + (k - 1) * NeX * NeY

    def node_id(i, j, k):
        return i + (j - 1) * n1 + (k - 1) * n1 ** 2

    vM = np.zeros((Ne, 6 * Nfp), dtype=np.int64)
    vP = np.zeros((Ne, 6 * Nfp), dtype=np.int64)
    for k in range(1, NeZ + 1):
        for j in range(1, NeY + 1):
            for i in range(1, NeX + 1):
                ke = elem_id(i, j, k)
                etoe = [elem_id(i, j - 1, k), elem_id(i + 1, j, k), elem_id(i, j + 1, k), elem_id(i - 1, j, k),
                        elem_id(i, j, k - 1), elem_id(i, j, k + 1)]
                for q in range(1, n1 + 1):
                    for p in range(1, n1 + 1):
                        n = p + (q - 1) * n1
                        own_h = [node_id(p, 1, q), node_id(n1, p, q), node_id(p, n1, q), node_id(1, p, q)]
                        nbr_h = [node_id(p, n1, q), node_id(1, p, q), node_id(p, 1, q), node_id(n1, p, q)]
                        own_v = [node_id(p, q, 1), node_id(p, q, n1)]
                        nbr_v = [node_id(p, q, n1), node_id(p, q, 1)]
                        for f in range(4):
                            vM[ke - 1, f * Nfp + n - 1] = (ke - 1) * Np + own_h[f]
                            vP[ke - 1, f * Nfp + n - 1] = (etoe[f] - 1) * Np + nbr_h[f]
                        for f in range(2):
                            vM[ke - 1, (4 + f) * Nfp + n - 1] = (ke - 1) * Np + own_v[f]
                            vP[ke - 1, (4 + f) * Nfp + n - 1] = (etoe[4 + f] - 1) * Np + nbr_v[f]
                pp = np.arange(1, Nfp + 1)
                if j == 1:
                    vP[ke - 1, 0 * Nfp:1 * Nfp] = Ne * Np + (pp + ((i - 1) + (k - 1) * NeX) * Nfp)
                if i == NeX:
                    vP[ke - 1, 1 * Nfp:2 * Nfp] = Ne * Np + Nfp * NeX * NeZ + (pp + ((j - 1) + (k - 1) * NeY) * Nfp)
                if j == NeY:
                    vP[ke - 1, 2 * Nfp:3 * Nfp] = Ne * Np + Nfp * (NeX + NeY) * NeZ + (pp + ((i - 1) + (k - 1) * NeX) * Nfp)
                if i == 1:
                    vP[ke - 1, 3 * Nfp:4 * Nfp] = Ne * Np + Nfp * (2 * NeX + NeY) * NeZ + (pp + ((j - 1) + (k - 1) * NeY) * Nfp)
                if k == 1:
                    vP[ke - 1, 4 * Nfp:5 * Nfp] = Ne * Np + 2 * Nfp * (NeX + NeY) * NeZ + (pp + ((i - 1) + (j - 1) * NeX) * Nfp)
                if k == NeZ:
                    vP[ke - 1, 5 * Nfp:6 * Nfp] = Ne * Np + 2 * Nfp * (NeX + NeY) * NeZ + NeX * NeY * Nfp + (pp + ((i - 1) + (j - 1) * NeX) * Nfp)
    return vM, vP


@pytest.mark.parametrize("dims", [(3, 3, 3), (4, 2, 3)])
def test_reference_connectivity_answers(dims):
    """Known-answer pin from the reference: check_connectivity of FElib/test/FE/mesh_cubedom3d_hexahedral/
    test_mesh_cubedom3d_hexahedral.f90:84-222 (3 x 3 x 3 elements, p = 2, one tile, periodic in all directions) writes VMapM
    and VMapP of every element in closed form -- interior faces pair with the opposite face of the neighbour, faces on the
    tile boundary point into the halo in the order south, east, north, west, bottom, top.  Restated 1-based as there and held
    against both mesh restatements."""
    e = HexElement(2)
    NeX, NeY, NeZ = dims
    m = LocalMeshCube(e, NeX, NeY, NeZ, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0, periodic=(True, True, True))
    o = Oracle(2, NeX, NeY, NeZ, (0, 1, 0, 1, 0, 1), periodic=(True, True, True))
    vM, vP = reference_vmaps(e.np1, NeX, NeY, NeZ)
    Ne = m.Ne
    assert np.array_equal(m.abi_vmapM().reshape(Ne, -1), vM)
    assert np.array_equal(m.abi_vmapP().reshape(Ne, -1), vP)
    assert np.array_equal(o.iarr("vmapM").reshape(Ne, -1) + 1, vM)
    assert np.array_equal(o.iarr("vmapP").reshape(Ne, -1) + 1, vP)


def test_reference_connectivity_answers_cubed_sphere_local_meshes():
    """The same answers for the six local meshes of MeshCubedSphereDom3D (FElib/test/FE/mesh_cubedsphere3d/
    test_mesh_cubedsphere3d.f90:74-215: 3 x 3 x 2 elements per panel, p = 2): every lateral panel face is a tile boundary."""
    from fe_project_b200.mesh import LocalMeshCubedSpherePanel
    e = HexElement(2)
    vM, vP = reference_vmaps(e.np1, 3, 3, 2)
    for pid in range(1, 7):
        m = LocalMeshCubedSpherePanel(e, pid, 3, 3, 2, 10.0e3, 6.37122e6)
        assert np.array_equal(m.abi_vmapM().reshape(m.Ne, -1), vM), pid
        assert np.array_equal(m.abi_vmapP().reshape(m.Ne, -1), vP), pid
        o = Oracle(2, 3, 3, 2, panel=dict(panelID=pid, ztop=10.0e3, RPlanet=6.37122e6))
        assert np.array_equal(o.iarr("vmapP").reshape(m.Ne, -1) + 1, vP), pid
