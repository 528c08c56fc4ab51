"""Regression fixtures of the ORACLE (not outputs of the reference, which cannot be built here: see DESIGN.md section 2).
They freeze what oracle/ computes today for a few tiny cases, so that a later edit of the restatement that changes results is
noticed (tests/test_oracle_golden.py).  Regenerate deliberately with:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
PROG = ("DDENS", "MOMX", "MOMY", "MOMZ", "DRHOT")


def cases():
    from cases import DensityCurrentCase, GlobalPanelCase, SoundWaveCase
    return {
        "heve_p3": (DensityCurrentCase(p=3, NeX=2, NeY=2, NeZ=2, perturb=2.0, dt=0.2, tinteg="ERK_SSP_3s3o", intrp_order=7), 3),
        "heve_p7_filter": (DensityCurrentCase(p=7, NeX=2, NeY=1, NeZ=2, perturb=2.0, dt=0.05, modalfilter=True), 2),
        "hevi_p7": (DensityCurrentCase(p=7, NeX=1, NeY=1, NeZ=3, perturb=2.0, eqs="NONHYDRO3D_HEVI", tinteg="IMEX_ARK232", dt=0.5), 2),
        "sound_wave_p7": (SoundWaveCase(p=7, NeX=1, NeY=1, NeZ=6, dt=1.0, tinteg="IMEX_ARK232", amplitude=1.0e-3), 2),
        "global_panel2_hevi": (GlobalPanelCase(p=7, panelID=2, NeX=1, NeY=1, NeZ=2, dt=20.0), 2),
    }


def compute():
    out = {}
    for name, (case, nsteps) in cases().items():
        o = case.make_oracle()
        o.update(nsteps)
        n = case.mesh.Ne * case.elem.Np
        for nm in PROG:
            out[f"{name}/{nm}"] = o.arr(nm)[:n].copy()
    # tracer advection (prescribed mass flux, limiters on)
    from fe_project_b200.advect3d import gaussian_hill
    from fe_project_b200.element import HexElement
    from fe_project_b200.mesh import LocalMeshCube
    from oracle_api import Oracle
    dom, per = (0, 1, 0, 1, 0, 1), (True, True, True)
    o = Oracle(3, 2, 2, 2, dom, periodic=per)
    mesh = LocalMeshCube(HexElement(3), 2, 2, 2, *dom, periodic=per)
    n = o.Np * o.Ne
    o.arr("DENS_hyd")[:n] = 1.0
    o.arr("DDENS")[:n] = 0.2 * np.sin(2 * np.pi * mesh.pos_en[0].reshape(-1))
    for nm, v in zip(("MOMX", "MOMY", "MOMZ"), (0.5, 0.3, -0.2)):
        o.arr(nm)[:n] = v
    q = np.zeros(o.Np * o.NeA)
    q[:n] = gaussian_hill(mesh, 0.4, 0.5, 0.5, width=0.1)[:mesh.Ne].reshape(-1)
    o.trcadv_update(q, "ERK_SSP_3s3o", 0.01, nsteps=4, modalfilter=(0.0, 1.0, 16, 0.0, 1.0, 16))
    out["tracer_p3/QTRC"] = q[:n].copy()
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "oracle_regression.npz"), **compute())
    print("written", os.path.join(HERE, "oracle_regression.npz"))
