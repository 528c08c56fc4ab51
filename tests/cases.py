"""Case builders of the package + their CPU checker (oracle/, test infrastructure): `case.make_oracle()`.

The case classes live in fe_project_b200/cases.py (bench.py and smoke() use them too); this module attaches the oracle
builders so that a test reads `o = case.make_oracle(); d = case.make_driver(o)`."""
from fe_project_b200.cases import (C0, MF, SLIP6, DensityCurrentCase, GlobalPanelCase, GlobalSphereCase,  # noqa: F401
                                   SoundWaveCase, rel_l2)
import oracle_cases

DensityCurrentCase.make_oracle = oracle_cases.make_oracle_regional      # SoundWaveCase inherits it
GlobalPanelCase.make_oracle = oracle_cases.make_oracle_panel
GlobalSphereCase.make_oracle = oracle_cases.make_oracle_sphere


# ---- terrain-following regional mesh (bell mountain), shared by the CPU oracle tests and the GPU parity tests
def terrain_case(p, dims, h0=600.0, **kw):
    """Bell mountain h(x) = h0 / (1 + ((x - xc)/a)^2 + ((y - yc)/b)^2) with the linear terrain-following map
    z = zeta + h (1 - zeta / zTop): GsqrtV = 1 - h / zTop, G13 = -(1 - zeta/zTop) h_x / GsqrtV, G23 likewise (the
    quantities MeshTopography%SetVCoordinate hands to Set_geometric_with_vcoord, mesh/scale_mesh_topography.F90:101-264,
    evaluated analytically here: the test is about the metric terms of the tendency, not about the mesh generator)."""
    case = DensityCurrentCase(p=p, NeX=dims[0], NeY=dims[1], NeZ=dims[2], perturb=2.0, intrp_order=min(11, p + 4), **kw)
    m = case.mesh
    x, y, zeta = m.pos_en[0], m.pos_en[1], m.pos_en[2]
    zT, a, b, xc, yc = m.zmax, 5.0e3, 4.0e3, 12.0e3, 3.0e3
    den = 1.0 + ((x - xc) / a) ** 2 + ((y - yc) / b) ** 2
    h = h0 / den
    hx = -h0 / den ** 2 * 2.0 * (x - xc) / a ** 2
    hy = -h0 / den ** 2 * 2.0 * (y - yc) / b ** 2
    gv = 1.0 - h / zT
    Ne = m.Ne
    m.Gsqrt[:Ne] = gv
    m.GI3[0, :Ne] = -(1.0 - zeta / zT) * hx / gv
    m.GI3[1, :Ne] = -(1.0 - zeta / zT) * hy / gv
    for arr in (m.Gsqrt, m.GI3[0], m.GI3[1]):
        m.exchange_halo_numpy(arr.reshape(-1))
    return case      # zlev (used by the potential-energy monitor only) stays the computational height on both sides


def terrain_oracle(case):
    o = case.make_oracle()
    m = case.mesh
    o.arr("Gsqrt")[:] = m.Gsqrt.reshape(-1)
    o.arr("G13")[:] = m.GI3[0].reshape(-1)
    o.arr("G23")[:] = m.GI3[1].reshape(-1)
    o.prepare()
    return o
