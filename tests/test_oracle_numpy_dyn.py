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
