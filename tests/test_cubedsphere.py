"""Panel-edge exchange of the cubed sphere (row e, global part) and the six-panel step of the oracle.
The tables of the reference (getPanelConnectivity, revert_hori, CS2LonLatVec / LonLat2CSVec) are restated twice
(NumPy: fe_project_b200/cubedsphere.py, C++: oracle/sphere.cpp) and pinned geometrically: the source node of every
halo slot is the same physical point as the receiver's own face node, and globally continuous scalar / vector fields
come back from the exchange equal to the own-face values."""
import numpy as np
import pytest

from cases import GlobalSphereCase
from fe_project_b200.cubedsphere import CubedSphere, cs2cart, panel_connectivity
from fe_project_b200.element import HexElement


def test_connectivity_is_geometrically_consistent():
    pc, fc = panel_connectivity()
    for T in range(6):                         # the graph is symmetric
        for f in range(4):
            U, g = pc[f, T] - 1, abs(fc[f, T]) - 1
            assert pc[g, U] - 1 == T and abs(fc[g, U]) - 1 == f
    cs = CubedSphere(HexElement(3), 3, 2, 30.0e3, 6.37122e6)
    for U in range(6):
        mU = cs.panels[U]
        for g, (T, src, rot) in cs.links[U].items():
            mT = cs.panels[T]
            own = cs._face_nodes(mU, g)
            pU = cs2cart(U + 1, mU.pos_en[0].reshape(-1)[own], mU.pos_en[1].reshape(-1)[own])
            pT = cs2cart(T + 1, mT.pos_en[0].reshape(-1)[src], mT.pos_en[1].reshape(-1)[src])
            assert np.abs(pU - pT).max() <= 1e-14
            assert np.abs(mU.pos_en[2].reshape(-1)[own] - mT.pos_en[2].reshape(-1)[src]).max() <= 1e-9
            assert np.abs(np.linalg.det(rot)).min() > 0.3      # the basis change never degenerates on an edge


def test_exchange_two_restatements_and_continuity():
    case = GlobalSphereCase(p=3, Ne=3, NeZ=2)
    s = case.make_oracle()
    s.exchange(with_dpres=False)
    Np = case.elem.Np
    fields = [{k: f[k].reshape(-1).copy() for k in ("DDENS", "DRHOT", "MOMZ", "MOMX", "MOMY")} for f in case.fields]
    case.cs.exchange_numpy(fields)
    for P, m in enumerate(case.cs.panels):
        nint = m.Ne * Np
        lat = slice(nint, nint + m.halo_face_off[4])            # the four lateral faces
        for k in ("DDENS", "DRHOT", "MOMZ", "MOMX", "MOMY"):
            a = s.panels[P].arr(k)
            sc = np.abs(a[:nint]).max()
            assert np.abs(a[lat] - fields[P][k][lat]).max() <= 1e-13 * sc, (P, k)          # C++ == NumPy
            own = a[m.VMapB[: m.halo_face_off[4]]]
            assert np.abs(a[lat] - own).max() <= 1e-12 * sc, (P, k)                          # continuous field: halo == own face


def _sphere_tend(case):
    s = case.make_oracle()
    for o in s.panels:
        o.piece("pressure")
    s.exchange(with_dpres=True)
    out = []
    for o, m in zip(s.panels, case.cs.panels):
        o.piece("bc"); o.piece("tend_ex")
        n, N = m.Ne * case.elem.Np, m.NeA * case.elem.Np
        out.append(o.arr("tend_ex")[:5 * N].reshape(5, -1)[:, :n].copy())
    return s, out


def test_balanced_rotation_is_steady_on_the_whole_sphere():
    """Including the polar panels and every panel edge: converges spectrally with p."""
    cor = 2 * 7.292e-5 * 30.0 / 6.37122e6
    errs = []
    for p in (3, 5, 7):
        _, te = _sphere_tend(GlobalSphereCase(p=p, Ne=2, NeZ=2, perturb=0.0))
        errs.append(max(max(np.abs(t[3]).max(), np.abs(t[4]).max()) for t in te))
    assert errs[0] < 0.2 * cor and errs[1] < 0.1 * errs[0] and errs[2] < 0.1 * errs[1], errs


def test_mass_is_conserved_over_the_closed_sphere():
    """Sum over the six panels of the DDENS tendency integral vanishes to round-off: the flux leaving a panel edge enters the
    neighbour (it does not on a single panel, whose lateral halo mirrors its own values)."""
    case = GlobalSphereCase(p=4, Ne=2, NeZ=2)
    _, te = _sphere_tend(case)
    tot, scale = 0.0, 0.0
    for t, m in zip(te, case.cs.panels):
        n = m.Ne * case.elem.Np
        w = np.tile(case.elem.IntWeight_lgl, m.Ne) * m.J.reshape(-1) * m.Gsqrt.reshape(-1)[:n]
        tot += np.sum(w * t[0]); scale += np.sum(w * np.abs(t[0]))
    assert abs(tot) <= 1e-12 * scale, (tot, scale)


def test_six_panel_steps_stay_finite():
    case = GlobalSphereCase(p=5, Ne=2, NeZ=3, dt=30.0)
    s = case.make_oracle()
    s.update(4)
    for o, m in zip(s.panels, case.cs.panels):
        n = m.Ne * case.elem.Np
        for k in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT"):
            assert np.isfinite(o.arr(k)[:n]).all()
        assert np.abs(o.arr("MOMZ")[:n]).max() < 2.0


def tile_to_panel_elements(cs_tiles, t):
    """Element index inside the whole panel (k Ne x k Ne x NeZ elements) of every element of tile t."""
    m = cs_tiles.panels[t]
    k, ti, tj = m.sub
    nxp = k * m.NeX
    return (m.ex + ti * m.NeX) + (m.ey + tj * m.NeY) * nxp + m.ez * nxp * (k * m.NeY)


def test_sub_panel_tiles_reproduce_the_whole_panel_exchange():
    """2 x 2 tiles per panel (24 local meshes: the layout for 4 and 8 GPUs): after the tile exchange every face node sees
    through VMapP exactly what it sees on the whole-panel mesh -- neighbour tiles inside a panel, reverted and non-reverted panel
    edges, the basis change of (MOMX, MOMY)."""
    e = HexElement(2)
    whole = CubedSphere(e, 4, 2, 30.0e3, 6.37122e6)
    tiles = CubedSphere(e, 2, 2, 30.0e3, 6.37122e6, ntile=2)
    assert len(tiles.panels) == 24
    Np = e.Np
    rng = np.random.default_rng(3)
    names = ("DDENS", "MOMX", "MOMY")
    fw = [{n: np.zeros(m.NeA * Np) for n in names} for m in whole.panels]
    for P, m in enumerate(whole.panels):
        for n in names:
            fw[P][n][: m.Ne * Np] = rng.standard_normal(m.Ne * Np)
    ft = []
    for t, m in enumerate(tiles.panels):
        P, kep = tiles.panel_of[t], tile_to_panel_elements(tiles, t)
        d = {}
        for n in names:
            a = np.full(m.NeA * Np, np.nan)
            a[: m.Ne * Np] = fw[P][n][: whole.panels[P].Ne * Np].reshape(-1, Np)[kep].reshape(-1)
            d[n] = a
        ft.append(d)
        # the tile covers the same nodes as its part of the panel
        assert np.abs(m.pos_en[0] - whole.panels[P].pos_en[0][kep]).max() <= 1e-15
        assert np.abs(m.pos_en[1] - whole.panels[P].pos_en[1][kep]).max() <= 1e-15
    whole.exchange_numpy(fw)
    tiles.exchange_numpy(ft)
    for t, m in enumerate(tiles.panels):
        P, kep = tiles.panel_of[t], tile_to_panel_elements(tiles, t)
        mp = whole.panels[P]
        lat = slice(0, 4 * e.Nfp)                              # the four lateral faces
        for n in names:
            got = ft[t][n][m.VMapP][:, lat]
            exp = fw[P][n][mp.VMapP][kep][:, lat]
            assert np.isfinite(got).all(), (t, n)
            if n == "DDENS":
                assert np.array_equal(got, exp), (t, n)
            else:                                              # rot from tile node positions: equal to round-off
                assert np.abs(got - exp).max() <= 1e-13 * np.abs(exp).max(), (t, n)


@pytest.mark.parametrize("nranks", [4, 8])
def test_tile_owner_and_plan_cover_every_face(nranks):
    from fe_project_b200.cubedsphere import exchange_plan, panel_owner
    tiles = CubedSphere(HexElement(1), 1, 1, 1.0e4, 6.37122e6, ntile=2)
    owner = panel_owner(nranks, 2)
    assert [owner.count(r) for r in range(nranks)] == [24 // nranks] * nranks
    seen_send, seen_recv = set(), set()
    for r in range(nranks):
        local, recvs, sends = exchange_plan(tiles.links, owner, r)
        mine = [t for t in range(24) if owner[t] == r]
        assert len(local) + len(recvs) == 4 * len(mine)
        seen_recv |= {(peer, r, mid) for _, _, peer, mid in recvs}
        seen_send |= {(r, peer, mid) for _, peer, mid, _, _ in sends}
    assert seen_send == seen_recv                      # every message has exactly one sender and one receiver
    with pytest.raises(ValueError):
        panel_owner(4, 1)


def _ref_get_index(s_face, Ne1D, n1, i1D, k, fph, fpv):
    """get_index of the reference's own halo test (FElib/test/FE/field_cubedspheredom3d/test_field_cubedspheredom3d.f90:367-407):
    source element and face-node number (1-based) behind halo slot (fph, fpv) of the i1D-th element along the edge in layer k,
    for the signed source face id of tileFaceID_globalMap."""
    a = abs(s_face)
    rev = s_face < 0
    ii = Ne1D - i1D + 1 if rev else i1D
    base = (k - 1) * Ne1D ** 2
    if a == 1:
        ke = ii + base
    elif a == 2:
        ke = Ne1D + (ii - 1) * Ne1D + base
    elif a == 3:
        ke = ii + (Ne1D - 1) * Ne1D + base
    else:
        ke = 1 + (ii - 1) * Ne1D + base
    fp = (n1 - fph + 1 if rev else fph) + (fpv - 1) * n1
    return ke, fp


def test_halo_sources_match_the_reference_field_test():
    """Known-answer pin from the reference: test_field_cubedspheredom3d.f90:243-350 fills q with a value that encodes (tile,
    element, node) and asserts the lateral halo of every tile face against get_index (3 x 3 x 2 elements per panel, six local
    meshes).  Restated here on the link tables: the source tile and the source node of every halo slot are the ones that test
    expects, for all 24 panel faces (reverted and not)."""
    e = HexElement(2)
    Ne, NeZ = 3, 2
    cs = CubedSphere(e, Ne, NeZ, 30.0e3, 6.37122e6)
    pc, fc = panel_connectivity()                       # tileID_globalMap / tileFaceID_globalMap for one tile per panel
    n1, Np = e.np1, e.Np
    for U in range(6):
        for g in range(4):
            s_face, T_ref = int(fc[g, U]), int(pc[g, U]) - 1
            T, src, _ = cs.links[U][g]
            assert T == T_ref, (U, g)
            exp = []
            for k in range(1, NeZ + 1):
                for i in range(1, Ne + 1):
                    for fpv in range(1, n1 + 1):
                        for fph in range(1, n1 + 1):
                            ke, fp = _ref_get_index(s_face, Ne, n1, i, k, fph, fpv)
                            exp.append((ke - 1) * Np + e.Fmask[abs(s_face) - 1][fp - 1])     # Fmask_h(fp, |s_face|)
            assert np.array_equal(src, np.asarray(exp)), (U, g, s_face)
    # and the exchange moves exactly those values: q = 1e6 tile + 1e3 element + node, as get_field_val builds it
    fields = []
    for P, m in enumerate(cs.panels):
        q = np.zeros(m.NeA * Np)
        q[: m.Ne * Np] = (1e6 * (P + 1) + 1e3 * (np.arange(m.Ne)[:, None] + 1) + (np.arange(Np)[None, :] + 1)).reshape(-1)
        fields.append({"q": q})
    cs.exchange_numpy(fields, vector_pairs=())
    for U, m in enumerate(cs.panels):
        nint = m.Ne * Np
        for g in range(4):
            s_face, T_ref = int(fc[g, U]), int(pc[g, U])
            o = m.halo_face_off[g]
            got = fields[U]["q"][nint + o: nint + o + m.halo_face_size[g]].reshape(NeZ, Ne, n1, n1)
            for k in range(1, NeZ + 1):
                for i in range(1, Ne + 1):
                    for fpv in range(1, n1 + 1):
                        for fph in range(1, n1 + 1):
                            ke, fp = _ref_get_index(s_face, Ne, n1, i, k, fph, fpv)
                            ans = 1e6 * T_ref + 1e3 * ke + (e.Fmask[abs(s_face) - 1][fp - 1] + 1)
                            assert got[k - 1, i - 1, fpv - 1, fph - 1] == ans, (U, g, k, i, fpv, fph)


@pytest.mark.parametrize("ntile,own", [(1, [2, 3]), (2, [9, 10, 11])])
def test_skeleton_tiles_give_the_same_links(ntile, own):
    """A rank of a multi-GPU run builds only its own tiles in full; the others are skeletons (sizes, VMapB, face coordinates).  The
    links with an end on an own tile -- source indices and the (MOMX, MOMY) basis change -- must be bit-identical to the ones of the
    fully built sphere, hpos() must reproduce pos_en, and the exchange plan of the rank must not change."""
    from fe_project_b200.cubedsphere import CubedSphere, exchange_plan, panel_owner
    from fe_project_b200.element import HexElement
    e = HexElement(3)
    FZ = np.array([0.0, 1000.0, 3000.0])
    full = CubedSphere(e, 2, 2, 3000.0, 6.37122e6, FZ=FZ, ntile=ntile)
    part = CubedSphere(e, 2, 2, 3000.0, 6.37122e6, FZ=FZ, ntile=ntile, build=own)
    for t, (a, b) in enumerate(zip(full.panels, part.panels)):
        assert b.skeleton == (t not in own)
        assert np.array_equal(a.VMapB, b.VMapB) and a.Nhalo == b.Nhalo and a.Ne == b.Ne
        idx = np.arange(a.Ne * e.Np)
        x, y = b.hpos(idx)
        assert np.array_equal(x, a.pos_en[0].reshape(-1)) and np.array_equal(y, a.pos_en[1].reshape(-1))
    for U in own:
        assert sorted(part.links[U]) == [0, 1, 2, 3]
        for g in range(4):
            T, src, rot = full.links[U][g]
            T2, src2, rot2 = part.links[U][g]
            assert T == T2 and np.array_equal(src, src2)
            assert (rot is None) == (rot2 is None) and (rot is None or np.array_equal(rot, rot2))
    n = 6 * ntile * ntile
    owner = [0 if t in own else 1 for t in range(n)]
    pf, pp = exchange_plan(full.links, owner, 0), exchange_plan(part.links, owner, 0)
    assert pf == pp
    for T, peer, mid, U, g in pp[2]:                      # what this rank sends: source indices on its own tiles
        assert np.array_equal(full.links[U][g][1], part.links[U][g][1])


@pytest.mark.parametrize("world,ntile", [(2, 1), (3, 1), (6, 1), (4, 2), (8, 2)])
def test_every_rank_gets_its_exchange_plan_from_skeletons(world, ntile):
    """What bench.py does on N GPUs: rank r builds its own tiles in full and the rest as skeletons.  Its exchange plan (local links,
    receives, sends) must be the plan the fully built sphere gives, and the plans of all ranks must pair up: every send of rank a to
    rank b with message id m is a receive of rank b from rank a with the same id and the same number of nodes."""
    from fe_project_b200.cubedsphere import CubedSphere, exchange_plan, panel_owner
    e = HexElement(3)
    FZ = np.array([0.0, 1000.0, 3000.0])
    full = CubedSphere(e, 2, 2, 3000.0, 6.37122e6, FZ=FZ, ntile=ntile)
    owner = panel_owner(world, ntile)
    sends, recvs = {}, {}
    for r in range(world):
        own = [t for t, o in enumerate(owner) if o == r]
        part = CubedSphere(e, 2, 2, 3000.0, 6.37122e6, FZ=FZ, ntile=ntile, build=own)
        plan = exchange_plan(part.links, owner, r)
        assert plan == exchange_plan(full.links, owner, r)
        for T, peer, mid, U, g in plan[2]:
            sends[(r, peer, mid)] = part.links[U][g][1].size
        for U, g, peer, mid in plan[1]:
            recvs[(peer, r, mid)] = part.links[U][g][1].size
            assert part.links[U][g][2] is None or part.links[U][g][2].shape == (part.links[U][g][1].size, 2, 2)
    assert sends == recvs and len(sends) > 0
