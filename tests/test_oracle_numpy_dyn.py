"""two independent restatements of the dynamics rows agree.  oracle/numpy_dyn.py (NumPy, written from the Fortran)
against oracle/dyn_heve.cpp (C++) for pressure, numerical flux + explicit HEVE tendency on flat and terrain-following meshes; the
vertical-implicit Newton step has its own second restatement in tests/test_vi_block_host.py (different elimination, different code)."""
import numpy as np
import pytest

from cases import DensityCurrentCase, rel_l2
import numpy_dyn


@pytest.mark.parametrize("p,dims,periodic", [(7, (3, 2, 2), (False, True, False)), (3, (4, 3, 3), (True, True, False)), (5, (2, 2, 2), (False, False, False))])
def test_numpy_restatement_of_flux_and_tendency_equals_the_cpp_oracle(p, dims, periodic):
    case = DensityCurrentCase(p=p, NeX=dims[0], NeY=dims[1], NeZ=dims[2], perturb=2.0, periodic=periodic, intrp_order=min(11, p + 4))
    o = case.make_oracle()
    for w in ("exchange", "pressure", "bc", "tend_ex"):
        o.piece(w)
    e, m, c = case.elem, case.mesh, case.consts
    n, N = m.Ne * e.Np, m.NeA * e.Np
    q = {k: o.arr(k).copy() for k in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT")}            # halo + boundary condition applied by the oracle
    aux = {k: o.arr(k).copy() for k in ("DENS_hyd", "PRES_hyd", "THERM_hyd")}
    R, cv, cp = o.arr("Rtot"), o.arr("CVtot"), o.arr("CPtot")
    pres, dpres = numpy_dyn.drhot2pres(c, q["DRHOT"], aux["PRES_hyd"], aux["THERM_hyd"], R, cv, cp)
    assert rel_l2(pres[:n], o.arr("PRES")[:n]) <= 1e-15
    # the halo of DPRES is exchanged like a field: take the oracle's, whose interior we have just reproduced
    t = numpy_dyn.cal_tend_heve(e, m, c, q, aux, o.arr("DPRES"), o.arr("DPhydDx"), o.arr("DPhydDy"))
    te = o.arr("tend_ex")[:5 * N].reshape(5, -1)[:, :n]
    for nm, iv in (("DENS_dt", 0), ("RHOT_dt", 1), ("MOMZ_dt", 2), ("MOMX_dt", 3), ("MOMY_dt", 4)):
        assert rel_l2(t[nm].reshape(-1), te[iv]) <= 1e-13, nm


@pytest.mark.parametrize("p,dims,terrain", [(7, (3, 2, 2), False), (3, (4, 3, 3), False), (7, (3, 2, 3), True)])
def test_numpy_restatement_of_the_hevi_explicit_rows_equals_the_cpp_oracle(p, dims, terrain):
    """The horizontally explicit part of the HEVI equation set (numflux_get_generalvc of rhot_hevi_numflux.F90, cal_tend of rhot_hevi.F90)
    written a second time in NumPy from the Fortran, on the flat mesh and over the bell mountain, against oracle/dyn_hevi.cpp."""
    from cases import terrain_case, terrain_oracle
    kw = dict(eqs="NONHYDRO3D_HEVI", tinteg="IMEX_ARK232")
    if terrain:
        case = terrain_case(p, dims, **kw)
        o = terrain_oracle(case)
    else:
        case = DensityCurrentCase(p=p, NeX=dims[0], NeY=dims[1], NeZ=dims[2], perturb=2.0, periodic=(False, True, False),
                                  intrp_order=min(11, p + 4), **kw)
        o = case.make_oracle()
    for w in ("exchange", "pressure", "bc", "tend_ex"):
        o.piece(w)
    e, m, c = case.elem, case.mesh, case.consts
    n, N = m.Ne * e.Np, m.NeA * e.Np
    q = {k: o.arr(k).copy() for k in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT")}
    aux = {k: o.arr(k).copy() for k in ("DENS_hyd", "PRES_hyd", "THERM_hyd")}
    t = numpy_dyn.cal_tend_hevi(e, m, c, q, aux, o.arr("DPRES"), o.arr("DPhydDx"), o.arr("DPhydDy"))
    te = o.arr("tend_ex")[:5 * N].reshape(5, -1)[:, :n]
    for nm, iv in (("DENS_dt", 0), ("RHOT_dt", 1), ("MOMZ_dt", 2), ("MOMX_dt", 3), ("MOMY_dt", 4)):
        assert rel_l2(t[nm].reshape(-1), te[iv]) <= 1e-13, nm
    if terrain:
        assert np.abs(m.GI3[0]).max() > 1e-3


@pytest.mark.parametrize("panelID", [1, 4, 5, 6])
def test_numpy_restatement_of_the_global_hevi_explicit_rows_equals_the_cpp_oracle(panelID):
    """numflux_get_generalhvc + the explicit tendency of GLOBALNONHYDRO3D_HEVI on one cubed-sphere panel (metric, Christoffel and Coriolis
    terms; the Coriolis term changes form between the equatorial panels, panel 5 and panel 6), written a second time in NumPy from the
    Fortran, against oracle/dyn_global.cpp."""
    from cases import GlobalPanelCase
    # balanced = False: with the zonal flow in gradient-wind balance MOMY_dt is a 1e-12 residual of cancelling terms, nothing to compare
    case = GlobalPanelCase(p=7, panelID=1, NeX=2, NeY=2, NeZ=3, balanced=False)
    case.panelID = panelID
    case.mesh.panelID = panelID        # the metric tables are functions of the panel coordinates only
    o = case.make_oracle()
    for w in ("exchange", "pressure", "bc", "tend_ex"):
        o.piece(w)
    e, m, c = case.elem, case.mesh, case.consts
    n, N = m.Ne * e.Np, m.NeA * e.Np
    q = {k: o.arr(k).copy() for k in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT")}
    aux = {k: o.arr(k).copy() for k in ("DENS_hyd", "PRES_hyd", "THERM_hyd")}
    t = numpy_dyn.cal_tend_hevi_global(e, m, c, q, aux, o.arr("DPRES"), o.arr("DPhydDx"), o.arr("DPhydDy"))
    te = o.arr("tend_ex")[:5 * N].reshape(5, -1)[:, :n]
    for nm, iv in (("DENS_dt", 0), ("RHOT_dt", 1), ("MOMZ_dt", 2), ("MOMX_dt", 3), ("MOMY_dt", 4)):
        assert rel_l2(t[nm].reshape(-1), te[iv]) <= 1e-13, (panelID, nm)
    assert np.abs(te[3]).max() > 0.0 and np.abs(te[4]).max() > 0.0      # contravariant momenta: tendencies of order u / R per second


@pytest.mark.parametrize("terrain", [False, True])
@pytest.mark.parametrize("impl_fac", [0.0, 0.1])
def test_numpy_restatement_of_cal_vi_equals_the_cpp_oracle(impl_fac, terrain):
    """The vertical-implicit Newton step written a second time from the Fortran, with a different solver: the whole column (NeZ elements x
    8 nodes x 3 variables) as ONE dense system handed to numpy.linalg.solve, where the reference and oracle/dyn_hevi.cpp run a block-Thomas
    sweep with a partial-pivot LU per element.  Flat mesh and bell mountain (GsqrtV, G13, G23, the (MOMX, MOMY) pre-solve feeding the
    vertical mass flux): the pin of the terrain-following HEVI row."""
    from cases import terrain_case, terrain_oracle
    kw = dict(eqs="NONHYDRO3D_HEVI", tinteg="IMEX_ARK232")
    if terrain:
        case = terrain_case(7, (2, 2, 3), **kw)
        o = terrain_oracle(case)
    else:
        case = DensityCurrentCase(p=7, NeX=2, NeY=2, NeZ=3, perturb=2.0, periodic=(False, True, False), **kw)
        o = case.make_oracle()
    e, m, c = case.elem, case.mesh, case.consts
    n = m.Ne * e.Np
    ORD = ("DDENS", "DRHOT", "MOMZ", "MOMX", "MOMY")            # the oracle's variable order
    cur = {k: o.arr(k).copy() for k in ORD}
    rng = np.random.default_rng(4)
    var0 = np.stack([o.arr(k).copy() for k in ORD])
    if impl_fac != 0.0:
        var0[:, :n] += 1e-3 * rng.standard_normal((5, n)) * np.abs(var0[:, :n]).max(axis=1, keepdims=True)
    ref = o.cal_vi(impl_fac, case.dt, var0)[:, :n]
    aux = {k: o.arr(k).copy() for k in ("DENS_hyd", "PRES_hyd")}
    got = numpy_dyn.cal_vi(e, m, c, aux, cur, {k: var0[i] for i, k in enumerate(ORD)}, impl_fac)
    for i, k in enumerate(ORD):
        if impl_fac == 0.0:
            # MOMZ_dt is the residual of the near-cancelling -dDPRES/dz and -g drho: 1e-12 of the terms is 1e-11 of their difference
            assert rel_l2(got[k], ref[i]) <= (1e-11 if k == "MOMZ" else 1e-12), (k, terrain)
        else:       # compare the Newton iterate (the tendency cancels to round-off where the vertical operator is inactive)
            qs_ref, qs_got = cur[k][:n] + impl_fac * ref[i], cur[k][:n] + impl_fac * got[k]
            assert rel_l2(qs_got, qs_ref) <= 1e-11, (k, terrain)
            assert np.abs(got[k] - ref[i]).max() <= 1e-10 * max(np.abs(var0[i]).max(), np.abs(ref[i]).max()) / impl_fac, (k, terrain)


def test_numpy_cal_vi_on_a_cubed_sphere_panel():
    """The global twin of cal_vi (globalnonhydro3d_rhot_hevi.F90:873-1066) is the regional column solve with GsqrtV = Gsqrt / (gam^2 GsqrtH)
    (= 1 in the shallow atmosphere without topography): the NumPy restatement on a panel mesh against the oracle."""
    from cases import GlobalPanelCase
    case = GlobalPanelCase(p=7, panelID=1, NeX=2, NeY=2, NeZ=3, balanced=False)
    o = case.make_oracle()
    e, m, c = case.elem, case.mesh, case.consts
    n = m.Ne * e.Np
    ORD = ("DDENS", "DRHOT", "MOMZ", "MOMX", "MOMY")
    cur = {k: o.arr(k).copy() for k in ORD}
    rng = np.random.default_rng(5)
    var0 = np.stack([o.arr(k).copy() for k in ORD])
    var0[:, :n] += 1e-3 * rng.standard_normal((5, n)) * np.abs(var0[:, :n]).max(axis=1, keepdims=True)
    impl_fac = 5.0
    ref = o.cal_vi(impl_fac, case.dt, var0)[:, :n]
    aux = {k: o.arr(k).copy() for k in ("DENS_hyd", "PRES_hyd")}
    got = numpy_dyn.cal_vi(e, m, c, aux, cur, {k: var0[i] for i, k in enumerate(ORD)}, impl_fac)
    for i, k in enumerate(ORD):
        qs_ref, qs_got = cur[k][:n] + impl_fac * ref[i], cur[k][:n] + impl_fac * got[k]
        assert rel_l2(qs_got, qs_ref) <= 1e-11, k


@pytest.mark.parametrize("eqs,tinteg,dt,terrain,mf", [("NONHYDRO3D_HEVE", "ERK_SSP_4s3o", 0.05, False, True),
                                                       ("NONHYDRO3D_HEVE", "ERK_SSP_3s3o", 0.05, True, False),
                                                       ("NONHYDRO3D_HEVI", "IMEX_ARK232", 0.2, False, True),
                                                       ("NONHYDRO3D_HEVI", "IMEX_ARK324", 0.2, True, False)])
def test_numpy_restatement_of_the_full_step_equals_the_cpp_oracle(eqs, tinteg, dt, terrain, mf):
    """The whole dynamics step a second time in NumPy: stage loop in Butcher form, halo exchange, pressure, slip-wall boundary
    condition (with the terrain-following metric), explicit tendency, vertical-implicit Newton step as a dense column solve, modal
    filter -- three steps against the C++ oracle's driver (low-storage / one-buffer Runge-Kutta forms, block-Thomas + LU)."""
    import oracle_api
    from cases import terrain_case, terrain_oracle
    kw = dict(eqs=eqs, tinteg=tinteg, dt=dt, modalfilter=mf)
    if terrain:
        case = terrain_case(7, (2, 2, 3), **kw)
        o = terrain_oracle(case)
    else:
        case = DensityCurrentCase(p=7, NeX=2, NeY=2, NeZ=3, perturb=2.0, periodic=(False, True, False), **kw)
        o = case.make_oracle()
    e, m, c = case.elem, case.mesh, case.consts
    n = m.Ne * e.Np
    q = {k: o.arr(k).copy() for k in numpy_dyn.PROG}
    aux = {k: o.arr(k).copy() for k in ("DENS_hyd", "PRES_hyd", "THERM_hyd")}
    filt = None
    if mf:
        from cases import MF
        filt = (e.filter1d(MF["MF_ETAC_h"], MF["MF_ALPHA_h"], MF["MF_ORDER_h"]), e.filter1d(MF["MF_ETAC_v"], MF["MF_ALPHA_v"], MF["MF_ORDER_v"]))
    bc6 = m.halo_bc_types({k: 2 for k in ("south", "east", "north", "west", "btm", "top")})
    numpy_dyn.update(e, m, c, q, aux, oracle_api.rk_tables(tinteg), dt, bc6, hevi=eqs.endswith("HEVI"), filt=filt, nsteps=3,
                     DPhydDx=o.arr("DPhydDx").copy(), DPhydDy=o.arr("DPhydDy").copy())
    o.update(3)
    for k in numpy_dyn.PROG:
        assert rel_l2(q[k][:n], o.arr(k)[:n]) <= 1e-11, (k, eqs, terrain)


@pytest.mark.parametrize("panelID", [2, 5, 6])
def test_numpy_restatement_of_the_global_heve_rows_equals_the_cpp_oracle(panelID):
    """GLOBALNONHYDRO3D_HEVE, shallow atmosphere: numflux_get_generalhvc of the HEVE set + cal_tend_shallow_atm in NumPy against
    oracle/dyn_global.cpp."""
    from cases import GlobalPanelCase
    case = GlobalPanelCase(p=7, panelID=1, NeX=2, NeY=2, NeZ=3, balanced=False, eqs="GLOBALNONHYDRO3D_HEVE", tinteg="ERK_SSP_4s3o", dt=0.5)
    case.panelID = panelID
    case.mesh.panelID = panelID
    o = case.make_oracle()
    for w in ("exchange", "pressure", "bc", "tend_ex"):
        o.piece(w)
    e, m, c = case.elem, case.mesh, case.consts
    n, N = m.Ne * e.Np, m.NeA * e.Np
    q = {k: o.arr(k).copy() for k in ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT")}
    aux = {k: o.arr(k).copy() for k in ("DENS_hyd", "PRES_hyd", "THERM_hyd")}
    t = numpy_dyn.cal_tend_hevi_global(e, m, c, q, aux, o.arr("DPRES"), o.arr("DPhydDx"), o.arr("DPhydDy"), hevi=False)
    te = o.arr("tend_ex")[:5 * N].reshape(5, -1)[:, :n]
    for nm, iv in (("DENS_dt", 0), ("RHOT_dt", 1), ("MOMZ_dt", 2), ("MOMX_dt", 3), ("MOMY_dt", 4)):
        # MOMZ_dt: residual of the near-cancelling vertical pressure gradient and buoyancy (see the cal_vi test)
        assert rel_l2(t[nm].reshape(-1), te[iv]) <= (1e-11 if nm == "MOMZ_dt" else 1e-13), (panelID, nm)


def test_numpy_restatement_of_the_six_panel_step_equals_the_cpp_oracle():
    """The whole cubed sphere stepped a second time in NumPy (panel-edge exchange by cubedsphere.exchange_numpy, the global explicit
    tendency, the dense-column Newton step, modal filter) against the C++ oracle's six-panel driver: two steps of IMEX_ARK232."""
    import oracle_api
    from cases import GlobalSphereCase, MF
    case = GlobalSphereCase(p=3, Ne=2, NeZ=2, dt=40.0, tinteg="IMEX_ARK232", modalfilter=True)
    s = case.make_oracle()
    e, c = case.elem, case.consts
    qs = [{k: o.arr(k).copy() for k in numpy_dyn.PROG} for o in s.panels]
    auxs = [{k: o.arr(k).copy() for k in ("DENS_hyd", "PRES_hyd", "THERM_hyd")} for o in s.panels]
    dphyd = [(o.arr("DPhydDx").copy(), o.arr("DPhydDy").copy()) for o in s.panels]
    filt = (e.filter1d(MF["MF_ETAC_h"], MF["MF_ALPHA_h"], MF["MF_ORDER_h"]), e.filter1d(MF["MF_ETAC_v"], MF["MF_ALPHA_v"], MF["MF_ORDER_v"]))
    numpy_dyn.update_sphere(e, case.cs, c, qs, auxs, oracle_api.rk_tables("IMEX_ARK232"), case.dt, dphyd, filt=filt, nsteps=2)
    s.update(2)
    for P, (o, m) in enumerate(zip(s.panels, case.cs.panels)):
        n = m.Ne * e.Np
        for k in numpy_dyn.PROG:
            ref = o.arr(k)[:n]
            scale = max(np.abs(op.arr(k)[:n]).max() for op in s.panels)
            assert np.linalg.norm(qs[P][k][:n] - ref) <= 1e-11 * max(np.linalg.norm(ref), 1e-3 * scale * np.sqrt(n)), (P, k)
