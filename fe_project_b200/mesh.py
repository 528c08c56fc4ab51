"""Host-side local mesh of a regional cube domain (one tile of `MeshCubeDom3D`).

Builds the arrays the dynamics kernels read from `LocalMesh3D`
(FElib/src/mesh/scale_localmesh_3d.F90:31-68): `Escale, Fscale, normal_fn, Gsqrt,
GI3, GsqrtH, J, VMapM, VMapP, VMapB, EMap3Dto2D, pos_en`, with the same element
numbering (`ke = i + (j-1)*NeX + (k-1)*NeX*NeY`), the same face order
(y-, x+, y+, x-, z-, z+) and the same halo ordering as
`MeshUtil3D_genPatchBoundaryMap` (FElib/src/mesh/scale_meshutil_3d.F90:478-641):
halo slots are face nodes, grouped by tile face 1..6, elements ascending inside a
face.  Geometry follows `MeshCubeDom3D_coord_conv` / `MeshBase3D_setGeometricInfo`
(scale_mesh_cubedom3d.F90:545-602, scale_mesh_base3d.F90:175-331).

The tile graph (`tile_neighbors`) restates `MeshUtil3D_buildGlobalMap`
(scale_meshutil_3d.F90:750-877): a face with no neighbour points back to the tile
itself with the same face id, which is what makes the halo exchange copy a tile's
own boundary values into its halo at a physical boundary.

Index maps are produced 0-based for numpy and exported 1-based (Fortran) through
`LocalMeshCube.abi_*` for the C ABI, which takes the reference's conventions.
"""
from __future__ import annotations

import numpy as np

from .element import HexElement

# tile face ids, 0-based here: 0 y-, 1 x+, 2 y+, 3 x-, 4 z-, 5 z+
OPPOSITE_FACE = (2, 3, 0, 1, 5, 4)

BND_NOSPEC, BND_PERIODIC, BND_SLIP, BND_NOSLIP = 0, 1, 2, 3   # scale_mesh_bndinfo.F90:48-54


class LocalMeshCube:
    def __init__(self, elem: HexElement, NeX: int, NeY: int, NeZ: int,
                 xmin: float, xmax: float, ymin: float, ymax: float,
                 zmin: float, zmax: float, FZ: np.ndarray | None = None,
                 periodic=(False, False, False),
                 NprcX: int = 1, NprcY: int = 1, pi: int = 0, pj: int = 0, skeleton: bool = False):
        """One tile (pi, pj) of an NprcX x NprcY horizontal decomposition; NeX/NeY/NeZ are per tile.
        skeleton = True: only what OTHER tiles need from this one to link their halos (sizes, vertex coordinates, `VMapB`, `hpos`):
        no per-node geometry, no VMapM / VMapP -- a rank of a multi-GPU run builds the tiles of the other ranks this way."""
        self.skeleton = bool(skeleton)
        self.elem = elem
        self.NeX, self.NeY, self.NeZ = NeX, NeY, NeZ
        self.NprcX, self.NprcY, self.pi, self.pj = NprcX, NprcY, pi, pj
        self.periodic = tuple(bool(b) for b in periodic)
        n = elem.np1
        Np, Nfp = elem.Np, elem.Nfp
        self.Ne = Ne = NeX * NeY * NeZ
        self.Ne2D = NeX * NeY
        self.NeA = Ne + 2 * (NeX + NeY) * NeZ + 2 * NeX * NeY
        self.Ne2DA = NeX * NeY + 2 * (NeX + NeY)
        delx = (xmax - xmin) / NprcX
        dely = (ymax - ymin) / NprcY
        self.xmin, self.xmax = xmin + pi * delx, xmin + (pi + 1) * delx
        self.ymin, self.ymax = ymin + pj * dely, ymin + (pj + 1) * dely
        if FZ is None:
            FZ = zmin + (zmax - zmin) * np.arange(NeZ + 1) / NeZ
        self.FZ = np.asarray(FZ, dtype=np.float64)
        assert self.FZ.size == NeZ + 1
        self.zmin, self.zmax = self.FZ[0], self.FZ[-1]

        # vertex coordinates exactly as MeshUtil3D_genCubeDomain computes them
        vx = (self.xmax - self.xmin) * np.arange(NeX + 1) / NeX + self.xmin
        vy = (self.ymax - self.ymin) * np.arange(NeY + 1) / NeY + self.ymin
        vz = self.FZ

        ez, ey, ex = np.meshgrid(np.arange(NeZ), np.arange(NeY), np.arange(NeX), indexing="ij")
        ex, ey, ez = ex.reshape(-1), ey.reshape(-1), ez.reshape(-1)
        self.ex, self.ey, self.ez = ex, ey, ez
        self.EMap3Dto2D = ex + ey * NeX

        self._vx, self._vy = vx, vy
        if skeleton:
            self._build_vmapB()
            self._build_tile_graph()
            return

        x0, x1 = vx[ex], vx[ex + 1]
        y0, y1 = vy[ey], vy[ey + 1]
        z0, z1 = vz[ez], vz[ez + 1]
        self.pos_en = np.empty((3, Ne, Np))
        self.pos_en[0] = x0[:, None] + 0.5 * (elem.x1[None, :] + 1.0) * (x1 - x0)[:, None]
        self.pos_en[1] = y0[:, None] + 0.5 * (elem.x2[None, :] + 1.0) * (y1 - y0)[:, None]
        self.pos_en[2] = z0[:, None] + 0.5 * (elem.x3[None, :] + 1.0) * (z1 - z0)[:, None]

        xX, yY, zZ = 0.5 * (x1 - x0), 0.5 * (y1 - y0), 0.5 * (z1 - z0)
        J = xX * yY * zZ
        # J and Escale are constant inside an element of this mapping: kept per element and handed out as read-only broadcast views
        # of the reference's shapes (the ABI copy in dyncore.py expands them for the call and drops them afterwards) -- 80 B per node
        # less host memory, which is what lets a 180 GB tile (2.4e8 nodes) be set up from Python
        self.J = np.broadcast_to(J[:, None], (Ne, Np))
        # Escale(:,ke,d,d): only the diagonal is non-zero for this mapping
        esc = np.zeros((3, 3, Ne, 1))
        esc[0, 0, :, 0] = (yY * zZ) / J
        esc[1, 1, :, 0] = (xX * zZ) / J
        esc[2, 2, :, 0] = (xX * yY) / J
        self.Escale = np.broadcast_to(esc, (3, 3, Ne, Np))

        # normals / Fscale (MeshCubeDom3D_calc_normal + setGeometricInfo)
        self.normal_fn = np.zeros((3, Ne, elem.NfpTot))
        self.Fscale = np.empty((Ne, elem.NfpTot))
        self.sJ = np.empty((Ne, elem.NfpTot))
        e11, e22, e33 = self.Escale[0, 0, :, 0], self.Escale[1, 1, :, 0], self.Escale[2, 2, :, 0]
        for f, (d, sgn, esc) in enumerate(((1, -1.0, e22), (0, 1.0, e11), (1, 1.0, e22),
                                           (0, -1.0, e11), (2, -1.0, e33), (2, 1.0, e33))):
            sl = slice(f * Nfp, (f + 1) * Nfp)
            nvec = sgn * esc
            sj = np.sqrt(nvec ** 2)
            self.normal_fn[d, :, sl] = (nvec / sj)[:, None]
            self.sJ[:, sl] = (sj * J)[:, None]
            self.Fscale[:, sl] = ((sj * J) / J)[:, None]

        # metric factors of the (flat) terrain-following map: defaults of setGeometricInfo
        self.Gsqrt = np.ones((self.NeA, Np))
        self.GI3 = np.zeros((2, self.NeA, Np))
        self.GsqrtH = np.ones((self.Ne2D, Nfp))
        self.gam = np.ones((self.NeA, Np))
        self.zlev = self.pos_en[2].copy()

        self._build_maps()
        self._build_tile_graph()

    # ------------------------------------------------------------------
    def hpos(self, idx):
        """Horizontal coordinates (x, y) of the interior nodes with flat 0-based indices idx, from the vertex coordinates (the same
        expression as pos_en, so the values are bit-identical); available on skeleton tiles too."""
        idx = np.asarray(idx)
        ke, p = idx // self.elem.Np, idx % self.elem.Np
        ex, ey = self.ex[ke], self.ey[ke]
        x0, x1 = self._vx[ex], self._vx[ex + 1]
        y0, y1 = self._vy[ey], self._vy[ey + 1]
        return x0 + 0.5 * (self.elem.x1[p] + 1.0) * (x1 - x0), y0 + 0.5 * (self.elem.x2[p] + 1.0) * (y1 - y0)

    def _boundary_faces(self):
        """Per tile face: (elements on it in ascending order, their rank along the face), genPatchBoundaryMap ordering."""
        NeX, NeY, NeZ = self.NeX, self.NeY, self.NeZ
        ex, ey, ez = self.ex, self.ey, self.ez
        on = (ey == 0, ex == NeX - 1, ey == NeY - 1, ex == 0, ez == 0, ez == NeZ - 1)
        rank = (ex + ez * NeX, ey + ez * NeY, ex + ez * NeX, ey + ez * NeY, ex + ey * NeX, ex + ey * NeX)
        return [(np.nonzero(on[f])[0], rank[f]) for f in range(6)]

    def _build_vmapB(self):
        """Halo layout (sizes, offsets) and VMapB: the interior node feeding each slot of the send buffer / owning each halo slot."""
        e = self.elem
        NeX, NeY, NeZ = self.NeX, self.NeY, self.NeZ
        Np, Nfp = e.Np, e.Nfp
        sizes = np.array([NeX * NeZ, NeY * NeZ, NeX * NeZ, NeY * NeZ, NeX * NeY, NeX * NeY]) * Nfp
        self.halo_face_size = sizes
        self.halo_face_off = np.concatenate([[0], np.cumsum(sizes)[:-1]])
        self.Nhalo = int(sizes.sum())
        vmapB = np.empty(self.Nhalo, dtype=np.int64)
        for f, (sel, rank) in enumerate(self._boundary_faces()):
            base = self.halo_face_off[f] + rank[sel] * Nfp
            vmapB[(base[:, None] + np.arange(Nfp)[None, :]).reshape(-1)] = (sel[:, None] * Np + e.Fmask[f][None, :]).reshape(-1)
        self.VMapB = vmapB

    def _build_maps(self):
        e = self.elem
        NeX, NeY, NeZ, Ne = self.NeX, self.NeY, self.NeZ, self.Ne
        Np, Nfp, NfpTot = e.Np, e.Nfp, e.NfpTot
        ke = np.arange(Ne)
        ex, ey, ez = self.ex, self.ey, self.ez
        vmapM = np.empty((Ne, NfpTot), dtype=np.int64)
        vmapP = np.empty((Ne, NfpTot), dtype=np.int64)
        for f in range(6):
            vmapM[:, f * Nfp:(f + 1) * Nfp] = ke[:, None] * Np + e.Fmask[f][None, :]
        # interior connections: the matching node of the neighbour's opposite face
        nb = [
            (ey > 0, ke - NeX), (ex < NeX - 1, ke + 1), (ey < NeY - 1, ke + NeX),
            (ex > 0, ke - 1), (ez > 0, ke - NeX * NeY), (ez < NeZ - 1, ke + NeX * NeY),
        ]
        on_bnd = []
        for f, (has, kn) in enumerate(nb):
            fo = OPPOSITE_FACE[f]
            sl = slice(f * Nfp, (f + 1) * Nfp)
            kn_ = np.where(has, kn, ke)
            fo_ = np.where(has, fo, f)
            vmapP[:, sl] = kn_[:, None] * Np + e.Fmask[fo_]
            on_bnd.append(~has)
        # tile-boundary faces -> halo slots (genPatchBoundaryMap ordering)
        sizes = np.array([NeX * NeZ, NeY * NeZ, NeX * NeZ, NeY * NeZ, NeX * NeY, NeX * NeY]) * Nfp
        self.halo_face_size = sizes
        self.halo_face_off = np.concatenate([[0], np.cumsum(sizes)[:-1]])
        self.Nhalo = int(sizes.sum())
        rank = [ex + ez * NeX, ey + ez * NeY, ex + ez * NeX, ey + ez * NeY, ex + ey * NeX, ex + ey * NeX]
        vmapB = np.empty(self.Nhalo, dtype=np.int64)
        for f in range(6):
            sel = np.nonzero(on_bnd[f])[0]                 # ascending ke
            r = rank[f][sel]
            assert np.array_equal(np.sort(r), r)
            base = self.halo_face_off[f] + r * Nfp
            sl = slice(f * Nfp, (f + 1) * Nfp)
            vmapP[sel, sl] = Np * Ne + base[:, None] + np.arange(Nfp)[None, :]
            vmapB[(base[:, None] + np.arange(Nfp)[None, :]).reshape(-1)] = \
                (sel[:, None] * Np + e.Fmask[f][None, :]).reshape(-1)
        self.VMapM, self.VMapP, self.VMapB = vmapM, vmapP, vmapB

    def _build_tile_graph(self):
        """(neighbour tile (pi,pj), neighbour face) per tile face; self+same face when none."""
        px, py = self.periodic[0], self.periodic[1]
        pi, pj, NX, NY = self.pi, self.pj, self.NprcX, self.NprcY
        nbr = []
        for f, (di, dj) in enumerate(((0, -1), (1, 0), (0, 1), (-1, 0))):
            qi, qj = pi + di, pj + dj
            if 0 <= qi < NX and 0 <= qj < NY:
                nbr.append(((qi, qj), OPPOSITE_FACE[f]))
            elif (di != 0 and px) or (dj != 0 and py):
                nbr.append(((qi % NX, qj % NY), OPPOSITE_FACE[f]))
            else:
                nbr.append(((pi, pj), f))
        for f in (4, 5):
            nbr.append(((pi, pj), OPPOSITE_FACE[f] if self.periodic[2] else f))
        self.tile_neighbors = nbr

    # ------------------------------------------------------------------
    def halo_bc_types(self, vel_bc: dict | None = None) -> np.ndarray:
        """Per tile face velocity BC id (bnd_Init_lc, scale_atm_dyn_dgm_bnd.F90:788-838).

        vel_bc maps 'south','east','north','west','btm','top' -> BND_* id.  Only faces whose
        neighbour is the tile itself with the same face id receive the BC; others stay NOSPEC.
        """
        names = ("south", "east", "north", "west", "btm", "top")
        out = np.zeros(6, dtype=np.int32)
        vel_bc = vel_bc or {}
        for f, nm in enumerate(names):
            (qi, qj), fo = self.tile_neighbors[f]
            if (qi, qj) == (self.pi, self.pj) and fo == f:
                out[f] = vel_bc.get(nm, BND_NOSPEC)
        return out

    def self_exchange_src(self) -> np.ndarray:
        """For a single-tile run: interior node index feeding each halo slot (0-based).

        Restates Put/Exchange/Get of MeshFieldCommCubeDom3D for same-rank neighbours
        (scale_meshfieldcomm_base.F90 exchange_core same-rank copy): halo slots of face f
        receive the sender's VMapB-ordered data of face f' = tile_neighbors[f].face.
        """
        src = np.empty(self.Nhalo, dtype=np.int64)
        for f in range(6):
            (qi, qj), fo = self.tile_neighbors[f]
            assert (qi, qj) == (self.pi, self.pj), "self_exchange_src needs a single-tile graph"
            assert self.halo_face_size[f] == self.halo_face_size[fo]
            o, oo, s = self.halo_face_off[f], self.halo_face_off[fo], self.halo_face_size[f]
            src[o:o + s] = self.VMapB[oo:oo + s]
        return src

    def exchange_halo_numpy(self, q: np.ndarray) -> None:
        """In-place single-tile halo fill of a (NeA*Np,) or (NeA, Np) field."""
        flat = q.reshape(-1)
        flat[self.Ne * self.elem.Np: self.Ne * self.elem.Np + self.Nhalo] = flat[self.self_exchange_src()]

    # ---- exports in the reference's conventions (Fortran order, 1-based) ----
    def abi_vmapM(self): return (self.VMapM + 1).astype(np.int32)
    def abi_vmapP(self): return (self.VMapP + 1).astype(np.int32)
    def abi_vmapB(self): return (self.VMapB + 1).astype(np.int32)
    def abi_emap3dto2d(self): return (self.EMap3Dto2D + 1).astype(np.int32)


class LocalMeshCubedSpherePanel(LocalMeshCube):
    """One whole panel of `MeshCubedSphereDom3D` as a single tile (shallow-atmosphere approximation, no topography).

    In the reference a panel is a cube mesh in the central angles (alpha, beta) in [-pi/4, pi/4]^2 and z
    (`MeshCubedSphereDom3D_coord_conv`, FElib/src/mesh/scale_mesh_cubedspheredom3d.F90:527-565), so everything of
    `LocalMeshCube` carries over; on top of it come the horizontal metric of the equiangular gnomonic map
    (`CubedSphereCoordCnv_GetMetric`, FElib/src/common/scale_cubedsphere_coord_cnv.F90:735-787) and
    `MeshCubedSphereDom3D_set_metric` (scale_mesh_cubedspheredom3d.F90:599-655): `G_ij, GIJ, GsqrtH` on the 2D nodes,
    `gam = 1`, `Gsqrt(:,ke) = GsqrtH(IndexH2Dto3D, ke2D)`, halo metric = own face value (`fill_halo_metric`, :657-678).
    The lateral halo of the tile holds its own face values: the panel-edge exchange is not part of this class."""

    def __init__(self, elem: HexElement, panelID: int, NeX: int, NeY: int, NeZ: int, ztop: float, RPlanet: float,
                 FZ: np.ndarray | None = None, sub=(1, 0, 0), skeleton: bool = False):
        """sub = (k, ti, tj): the tile (ti, tj) of a k x k decomposition of the panel (NeX, NeY are per tile), the layout the
        reference uses beyond six processes (`MeshCubedSphereDom2D` with NprcX = NprcY = k tiles per panel,
        scale_mesh_cubedspheredom2d.F90:193-247).  The tile is its own local mesh: all four lateral faces are filled by links
        (fe_project_b200/cubedsphere.py), none by the tile graph of `LocalMeshCube`."""
        q = 0.25 * np.pi
        k, ti, tj = sub
        assert 0 <= ti < k and 0 <= tj < k
        w = 2.0 * q / k
        super().__init__(elem, NeX, NeY, NeZ, -q + ti * w, -q + (ti + 1) * w if ti + 1 < k else q,
                         -q + tj * w, -q + (tj + 1) * w if tj + 1 < k else q, 0.0, ztop, FZ=FZ, periodic=(False, False, False),
                         skeleton=skeleton)
        self.sub = (int(k), int(ti), int(tj))
        assert 1 <= panelID <= 6
        self.panelID, self.RPlanet = int(panelID), float(RPlanet)
        if skeleton:
            return
        Np, Nfp, Ne = elem.Np, elem.Nfp, self.Ne
        # 2D nodes = bottom-layer elements, k = 0 plane
        a = self.pos_en[0][: self.Ne2D, :Nfp]
        b = self.pos_en[1][: self.Ne2D, :Nfp]
        self.pos2D = np.stack([a, b])                     # (2, Ne2D, Nfp)
        X, Y = np.tan(a), np.tan(b)
        r2 = 1.0 + X ** 2 + Y ** 2
        ox, oy = 1.0 + X ** 2, 1.0 + Y ** 2
        fac = ox * oy * (RPlanet / r2) ** 2
        self.G_ij = np.empty((2, 2, self.Ne2D, Nfp))
        self.G_ij[0, 0], self.G_ij[0, 1], self.G_ij[1, 0], self.G_ij[1, 1] = fac * ox, -fac * (X * Y), -fac * (X * Y), fac * oy
        self.GsqrtH = RPlanet ** 2 * ox * oy / (r2 * np.sqrt(r2))
        f2 = 1.0 / self.GsqrtH ** 2
        self.GIJ = np.empty((2, 2, self.Ne2D, Nfp))
        self.GIJ[0, 0], self.GIJ[0, 1], self.GIJ[1, 0], self.GIJ[1, 1] = (f2 * self.G_ij[1, 1], -f2 * self.G_ij[0, 1],
                                                                        -f2 * self.G_ij[1, 0], f2 * self.G_ij[0, 0])
        self.gam = np.ones((self.NeA, Np))
        self.Gsqrt = np.ones((self.NeA, Np))
        self.Gsqrt[:Ne] = self.GsqrtH[self.EMap3Dto2D][:, elem.IndexH2Dto3D]
        g = self.Gsqrt.reshape(-1)
        halo = self.VMapP >= Ne * Np
        g[self.VMapP[halo]] = g[self.VMapM[halo]]
        if self.panelID <= 4:
            # CubedSphereCoordCnv_CS2LonLatPos, equatorial panels (scale_cubedsphere_coord_cnv.F90:106-119)
            self.lon2D = a + 0.5 * np.pi * (self.panelID - 1)
            self.lat2D = np.arctan(np.tan(b) * np.cos(a))

    def lonlat_to_cs_vec(self, u_lon, v_lat):
        """CubedSphereCoordCnv_LonLat2CSVec for an equatorial panel (scale_cubedsphere_coord_cnv.F90:349-369), on the 2D
        nodes: physical (zonal, meridional) components -> contravariant (alpha, beta) components, gam = 1."""
        assert self.panelID <= 4
        X, Y = np.tan(self.pos2D[0]), np.tan(self.pos2D[1])
        del2 = 1.0 + X ** 2 + Y ** 2
        uc = u_lon / np.cos(self.lat2D)
        return uc / self.RPlanet, (X * Y * uc + del2 / np.sqrt(1.0 + X ** 2) * v_lat) / (self.RPlanet * (1.0 + Y ** 2))
