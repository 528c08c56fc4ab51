"""Host-side set-up of the cubed-sphere panel graph: which face of which panel fills a halo face, in which order, and
how the horizontal momentum is re-expressed in the receiving panel's basis.

Restates, for one tile per panel (the reference's `Nprc = 6` layout):

* `MeshUtilCubedSphere2D_getPanelConnectivity` (FElib/src/mesh/scale_meshutil_cubedsphere2d.F90:251-290): neighbour
  panel and *destination* face of every lateral panel face; a negative face id asks for the send buffer to be
  reverted along the edge.
* the same-rank path of `MeshFieldCommBase_exchange_core` (FElib/src/data/scale_meshfieldcomm_base.F90:870-895): the
  boundary data of (tile T, face f) lands in the halo of face |s_faceID| of tile s_tileID.
* `push_localsendbuf` / `revert_hori` (FElib/src/data/scale_meshfieldcomm_cubedspheredom3d.F90:492-540).
* `CubedSphereCoordCnv_CS2LonLatVec`, `_LonLat2CSVec`, `_CS2CartPos`
  (FElib/src/common/scale_cubedsphere_coord_cnv.F90:150-236, 314-401, 406-488) with `gam = 1` (shallow atmosphere):
  the sender turns (MOMX, MOMY) into lon-lat components at its face nodes, the receiver turns them into its own
  contravariant components at its own face nodes (`MeshFieldCommCubedSphereDom3D_exchange` :226-420).

The product of the two conversions is a 2x2 matrix per halo node; the device only gathers and applies it
(`fedg_link_halo`).  `cs2cart` gives an independent, purely geometric check of the tables (tests/test_cubedsphere.py).
"""
from __future__ import annotations

import numpy as np

from .mesh import LocalMeshCubedSpherePanel

EPS = 2.220446e-16


def panel_connectivity():
    """(panel_connectivity, face_connectivity), 1-based ids, index [face-1][panel-1]."""
    pc = np.zeros((4, 6), dtype=int)
    fc = np.zeros((4, 6), dtype=int)
    zonal = [4, 1, 2, 3, 4, 1]
    for n in range(1, 5):
        pc[0, n - 1] = 6
        fc[0, n - 1] = zonal[4 - n] if zonal[4 - n] > 2 else -zonal[4 - n]
        pc[1, n - 1] = zonal[n + 1]
        fc[1, n - 1] = 4
        pc[2, n - 1] = 5
        fc[2, n - 1] = -n if n > 2 else n
        pc[3, n - 1] = zonal[n - 1]
        fc[3, n - 1] = 2
    pc[:, 4] = (1, 2, 3, 4); fc[:, 4] = (3, 3, -3, -3)
    pc[:, 5] = (3, 2, 1, 4); fc[:, 5] = (-1, -1, 1, 1)
    return pc, fc


def cs2cart(panel, a, b, R=1.0):
    x1, x2 = np.tan(a), np.tan(b)
    fac = R / np.sqrt(1.0 + x1 ** 2 + x2 ** 2)
    if panel == 1: return np.stack([fac, fac * x1, fac * x2])
    if panel == 2: return np.stack([-fac * x1, fac, fac * x2])
    if panel == 3: return np.stack([-fac, -fac * x1, fac * x2])
    if panel == 4: return np.stack([fac * x1, -fac, fac * x2])
    if panel == 5: return np.stack([-fac * x2, fac * x1, fac])
    return np.stack([fac * x2, fac * x1, -fac])


def cs2lonlat(panel, a, b):
    x, y, z = cs2cart(panel, a, b)
    return np.arctan2(y, x), np.arcsin(np.clip(z, -1.0, 1.0))


def _coslat(panel, X, Y, a, b):
    if panel <= 4:
        return np.cos(np.arctan(np.tan(b) * np.cos(a)))
    s = 1.0 if panel == 5 else -1.0
    return np.cos(np.arctan(s / np.maximum(np.sqrt(X ** 2 + Y ** 2), EPS)))


def cs2lonlat_vec(panel, a, b, va, vb, R):
    X, Y = np.tan(a), np.tan(b)
    del2 = 1.0 + X ** 2 + Y ** 2
    cl = _coslat(panel, X, Y, a, b)
    if panel <= 4:
        return va * cl * R, (-X * Y * va + (1.0 + Y ** 2) * vb) * R * np.sqrt(1.0 + X ** 2) / del2
    r = R if panel == 5 else -R
    h2 = np.maximum(X ** 2 + Y ** 2, EPS)
    vlon = (-Y * (1.0 + X ** 2) * va + X * (1.0 + Y ** 2) * vb) * r / h2 * cl
    vlat = (-X * (1.0 + X ** 2) * va - Y * (1.0 + Y ** 2) * vb) * r / (del2 * np.maximum(np.sqrt(X ** 2 + Y ** 2), EPS))
    return vlon, vlat


def lonlat2cs_vec(panel, a, b, vlon, vlat, R):
    X, Y = np.tan(a), np.tan(b)
    del2 = 1.0 + X ** 2 + Y ** 2
    uc = vlon / _coslat(panel, X, Y, a, b)
    if panel <= 4:
        return uc / R, (X * Y * uc + del2 / np.sqrt(1.0 + X ** 2) * vlat) / (R * (1.0 + Y ** 2))
    r = R if panel == 5 else -R
    sq = np.sqrt(np.maximum(del2 - 1.0, EPS))
    return ((-Y * uc - del2 * X / sq * vlat) / (r * (1.0 + X ** 2)),
            (X * uc - del2 * Y / sq * vlat) / (r * (1.0 + Y ** 2)))


def revert_hori(idx, np1, nv, nex, nez):
    """revert_hori of push_localsendbuf: the face buffer is (Nnode_h1D, Nnode_v, NeX, NeZ), first index fastest."""
    a = np.asarray(idx).reshape(nez, nex, nv, np1)
    return a[:, ::-1, :, ::-1].reshape(-1)


class CubedSphere:
    """The panel tiles + the halo links between them.  ntile = k: every panel is cut into k x k tiles (24 local meshes for
    k = 2, the layout that spreads over 4 or 8 GPUs); `panels` lists the tiles panel by panel, tile (ti, tj) of panel P at
    index P k^2 + tj k + ti, and `links[U][g] = (T, src, rot)` says that the halo of face g of tile U is node-wise the interior
    nodes `src` of tile T, with (MOMX, MOMY) multiplied by `rot` (None inside a panel: same basis)."""

    def __init__(self, elem, Ne, NeZ, ztop, RPlanet, FZ=None, ntile=1, build=None):
        """Ne: elements per tile edge (the panel has ntile * Ne).  build: the tiles whose full geometry is needed (default all; a rank
        of a multi-GPU run passes the ones it owns): the others are skeletons (sizes, send-buffer maps and face coordinates, enough
        to link halos to them), and only the links with an end in `build` are worked out."""
        self.elem, self.Ne_h, self.NeZ, self.R, self.ntile = elem, Ne, NeZ, RPlanet, int(ntile)
        k = self.ntile
        self.build = None if build is None else set(int(t) for t in build)
        self.panels = [LocalMeshCubedSpherePanel(elem, pid, Ne, Ne, NeZ, ztop, RPlanet, FZ=FZ, sub=(k, ti, tj),
                                                 skeleton=(self.build is not None and (pid - 1) * k * k + tj * k + ti not in self.build))
                       for pid in range(1, 7) for tj in range(k) for ti in range(k)]
        self.panel_of = [P for P in range(6) for _ in range(k * k)]
        self.links = self._build_links()

    def tile_index(self, P, ti, tj):
        return P * self.ntile ** 2 + tj * self.ntile + ti

    def _face_nodes(self, mesh, f):
        """0-based flat interior indices of the boundary nodes of tile face f (0-based) in halo order."""
        o, n = mesh.halo_face_off[f], mesh.halo_face_size[f]
        return mesh.VMapB[o:o + n]

    @staticmethod
    def _edge_tile(k, f, e):
        """(ti, tj) of the e-th tile along panel face f (faces: 0 south, 1 east, 2 north, 3 west)."""
        return ((e, 0), (k - 1, e), (e, k - 1), (0, e))[f]

    def _build_links(self):
        """Panel edges follow `panel_connectivity`; a reverted edge (negative face id) reverts the order of the tiles along the
        edge as well as the nodes inside each tile (`revert_hori`).  Inside a panel the neighbour tile's opposite face feeds
        the halo in the same order (the tile graph of MeshCubeDom3D)."""
        pc, fc = panel_connectivity()
        e = self.elem
        npts, nv = e.np1, e.np1
        k = self.ntile
        links = [dict() for _ in self.panels]
        # tile faces inside a panel: (dx, dy) of the neighbour and its opposite face
        inner = ((0, -1, 2), (1, 0, 3), (0, 1, 0), (-1, 0, 1))
        for P in range(6):
            for tj in range(k):
                for ti in range(k):
                    U = self.tile_index(P, ti, tj)
                    for g, (dx, dy, fo) in enumerate(inner):
                        qi, qj = ti + dx, tj + dy
                        if 0 <= qi < k and 0 <= qj < k:
                            T = self.tile_index(P, qi, qj)
                            if self.build is not None and U not in self.build and T not in self.build:
                                continue
                            links[U][g] = (T, self._face_nodes(self.panels[T], fo).copy(), None)
        for Tp in range(6):
            for f in range(4):
                Up = pc[f, Tp] - 1
                g = abs(fc[f, Tp]) - 1
                rev = fc[f, Tp] < 0
                for et in range(k):
                    T = self.tile_index(Tp, *self._edge_tile(k, f, et))
                    U = self.tile_index(Up, *self._edge_tile(k, g, k - 1 - et if rev else et))
                    if self.build is not None and U not in self.build and T not in self.build:
                        continue
                    mT, mU = self.panels[T], self.panels[U]
                    src = self._face_nodes(mT, f)
                    if rev:
                        src = revert_hori(src, npts, nv, self.Ne_h, self.NeZ)
                    own = self._face_nodes(mU, g)
                    assert own.size == src.size
                    rot = None
                    if self.build is None or U in self.build:     # the basis change is applied by the receiving tile
                        # positions: horizontal coordinates of the 3D nodes
                        aT, bT = mT.hpos(src)
                        aU, bU = mU.hpos(own)
                        rot = np.empty((own.size, 2, 2))
                        for c, (va, vb) in enumerate(((1.0, 0.0), (0.0, 1.0))):
                            vl, vt = cs2lonlat_vec(Tp + 1, aT, bT, np.full(own.size, va), np.full(own.size, vb), self.R)
                            ua, ub = lonlat2cs_vec(Up + 1, aU, bU, vl, vt, self.R)
                            rot[:, 0, c], rot[:, 1, c] = ua, ub
                    assert g not in links[U], "two faces feed the same halo face"
                    links[U][g] = (T, src.copy(), rot)
        for U in range(len(self.panels)):
            if self.build is None or U in self.build:
                assert sorted(links[U]) == [0, 1, 2, 3]
        return links

    def exchange_numpy(self, fields, vector_pairs=(("MOMX", "MOMY"),)):
        """fields: list of 6 dicts name -> (NeA*Np,) arrays.  Fills the lateral halos (reference semantics) in place."""
        vec = {n for pr in vector_pairs for n in pr}
        for U in range(len(self.panels)):
            mU = self.panels[U]
            nint = mU.Ne * self.elem.Np
            for g, (T, src, rot) in self.links[U].items():
                o = mU.halo_face_off[g]
                sl = slice(nint + o, nint + o + src.size)
                for name, arr in fields[U].items():
                    if name not in vec:
                        arr[sl] = fields[T][name][src]
                for nx_, ny_ in vector_pairs:
                    if nx_ in fields[U]:
                        sx, sy = fields[T][nx_][src], fields[T][ny_][src]
                        if rot is None:
                            fields[U][nx_][sl], fields[U][ny_][sl] = sx, sy
                        else:
                            fields[U][nx_][sl] = rot[:, 0, 0] * sx + rot[:, 0, 1] * sy
                            fields[U][ny_][sl] = rot[:, 1, 0] * sx + rot[:, 1, 1] * sy


def panel_owner(nranks: int, ntile: int = 1):
    """Rank owning each local mesh (tile), contiguous blocks in tile order.  ntile = 1: whole panels over 1, 2, 3 or 6 ranks,
    as the reference distributes them (`MeshCubedSphereDom2D_check_division_params`,
    FElib/src/mesh/scale_mesh_cubedspheredom2d.F90:193-247).  ntile = k > 1: the reference's 6 k^2-tile graph with several
    local meshes per rank (its `NLocalMeshPerPrc`), any rank count that divides 6 k^2 (4 and 8 GPUs at k = 2)."""
    n = 6 * ntile * ntile
    if ntile == 1 and nranks not in (1, 2, 3, 6):
        raise ValueError("whole panels are distributed over 1, 2, 3 or 6 ranks (the reference's rule); use ntile = 2 for 4 or 8")
    if n % nranks:
        raise ValueError(f"{nranks} ranks do not divide the {n} tiles")
    per = n // nranks
    return [t // per for t in range(n)]


def exchange_plan(links, owner, rank):
    """Panel-edge messages of `rank`: (local, recvs, sends).
    local: (U, g, T) both on this rank -> fedg_link_halo;  recvs: (U, g, peer, msg_id) halo face g of own panel U filled by
    rank peer -> fedg_link_halo_recv;  sends: (T, peer, msg_id, U, g) own panel T feeds (U, g) on rank peer ->
    fedg_link_halo_send.  msg_id = 6 U + g identifies the linked face on both sides; NCCL matches the messages of a rank pair in
    ascending msg_id (tests/test_multi_tile.py runs the plan over gloo)."""
    local, recvs, sends = [], [], []
    for U in range(len(links)):
        for g, (T, _src, _rot) in links[U].items():
            mid = 6 * U + g
            if owner[U] == rank and owner[T] == rank:
                local.append((U, g, T))
            elif owner[U] == rank:
                recvs.append((U, g, owner[T], mid))
            elif owner[T] == rank:
                sends.append((T, owner[U], mid, U, g))
    recvs.sort(key=lambda r: (r[2], r[3]))
    sends.sort(key=lambda r: (r[1], r[2]))
    return local, recvs, sends


class GlobalSphereDriver:
    """The six local meshes of the global model on one GPU: one `AtmDynDGMDriver_nonhydro3d` per panel, linked halos,
    stage-synchronous stepping (`fedg_group_update`), i.e. what `AtmDynDGMDriver_nonhydro3d%Update` does with
    `LOCAL_MESH_NUM = 6` (fluid_dyn_solver/scale_atm_dyn_dgm_driver_nonhydro3d.F90:703-921)."""

    def __init__(self, cs: CubedSphere, consts: dict, vel_bc=None, rank: int = 0, nranks: int = 1, bcast=None):
        """rank / nranks: the panels are spread over the ranks (`panel_owner`), one GPU per rank; bcast broadcasts the NCCL
        unique id (see AtmDynDGMDriver_nonhydro3d.init_comm).  `panels` holds the drivers of the OWN panels, `panel_ids` their
        0-based panel numbers."""
        import ctypes as C
        from . import _lib
        from .dyncore import AtmDynDGMDriver_nonhydro3d
        self.cs, self.L = cs, _lib.load()
        self.rank, self.nranks = rank, nranks
        self.owner = panel_owner(nranks, cs.ntile)
        self.panel_ids = [P for P in range(len(cs.panels)) if self.owner[P] == rank]      # local meshes (tiles) of this rank
        bc = vel_bc or dict(btm="SLIP", top="SLIP")
        self.panels = [AtmDynDGMDriver_nonhydro3d(cs.elem, cs.panels[P], consts, vel_bc=bc, my_rank=rank) for P in self.panel_ids]
        drv = dict(zip(self.panel_ids, self.panels))
        self._keep = []
        local, recvs, sends = exchange_plan(cs.links, self.owner, rank)
        vp = C.c_void_p
        for U, g, T in local:
            _, src, rot = cs.links[U][g]
            idx = np.ascontiguousarray(src + 1, dtype=np.int32)
            r = None if rot is None else np.ascontiguousarray(rot.reshape(-1, 4), dtype=np.float64)   # [r00, r01, r10, r11] per node
            self._keep += [idx, r]
            _lib.check(self.L.fedg_link_halo(drv[U].h, g + 1, drv[T].h, idx.ctypes.data_as(vp), None if r is None else r.ctypes.data_as(vp)))
        for U, g, peer, mid in recvs:
            rot = cs.links[U][g][2]
            r = None if rot is None else np.ascontiguousarray(rot.reshape(-1, 4), dtype=np.float64)
            self._keep.append(r)
            _lib.check(self.L.fedg_link_halo_recv(drv[U].h, g + 1, peer, mid, None if r is None else r.ctypes.data_as(vp)))
        for T, peer, mid, U, g in sends:
            idx = np.ascontiguousarray(cs.links[U][g][1] + 1, dtype=np.int32)
            self._keep.append(idx)
            _lib.check(self.L.fedg_link_halo_send(drv[T].h, peer, mid, idx.ctypes.data_as(vp), idx.size))
        if nranks > 1:
            self.panels[0].init_comm(rank, nranks, bcast)      # the group's communicator lives on its first mesh
        self._h = (C.c_void_p * len(self.panels))(*[d.h for d in self.panels])

    def Init(self, *a, **kw):
        for d in self.panels:
            d.Init(*a, **kw)

    def Update(self, nsteps=1):
        from . import _lib
        _lib.check(self.L.fedg_group_update(self._h, len(self.panels), int(nsteps)))

    def exchange_aux(self):
        """AUX_VARS exchange over the linked faces (DENS_hyd, PRES_hyd, THERM_hyd); after set_aux on every own panel."""
        from . import _lib
        _lib.check(self.L.fedg_group_exchange_aux(self._h, len(self.panels)))

    def exchange_halo(self, apply_bc=False):
        """MeshFieldComm_Exchange of the prognostic variables of every own panel (remote panel edges included)."""
        from . import _lib
        _lib.check(self.L.fedg_group_exchange_halo(self._h, len(self.panels), int(bool(apply_bc))))

    def last_timing(self):
        return self.panels[0].last_timing()
