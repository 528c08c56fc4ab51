"""CPU checker of a synthetic case (fe_project_b200/cases.py): the oracle fed with the same arrays the CUDA path gets.
Test infrastructure only (tests/, __graft_entry__.smoke(), the cpu_baseline / --impl reference legs of bench.py)."""
from __future__ import annotations

from oracle_api import Oracle, OracleSphere


def _mf(case):
    m = case.mf
    return (m["MF_ETAC_h"], m["MF_ALPHA_h"], m["MF_ORDER_h"], m["MF_ETAC_v"], m["MF_ALPHA_v"], m["MF_ORDER_v"])


def _dry(o, c):
    o.arr("Rtot")[:] = c["Rdry"]; o.arr("CVtot")[:] = c["CVdry"]; o.arr("CPtot")[:] = c["CPdry"]


def make_oracle_regional(case):
    m = case.mesh
    o = Oracle(case.p, m.NeX, m.NeY, m.NeZ, case.dom, periodic=case.periodic, lumped=case.elem.lumped)
    o.set_consts(case.consts)
    for k, v in case.fields.items():
        o.arr(k)[:] = v.reshape(-1)
    _dry(o, case.consts)
    o.setup_dyn(case.eqs, case.tinteg, case.dt, case.modalfilter, _mf(case), (2, 2, 2, 2, 2, 2))
    o.prepare()
    return o


def make_oracle_panel(case):
    m, c = case.mesh, case.consts
    o = Oracle(case.p, m.NeX, m.NeY, m.NeZ, lumped=case.elem.lumped,
               panel=dict(panelID=case.panelID, ztop=case.ztop, RPlanet=c["RPlanet"]))
    o.set_consts(c)
    for k, v in case.fields.items():
        o.arr(k)[:] = v.reshape(-1)
    _dry(o, c)
    o.setup_dyn(case.eqs, case.tinteg, case.dt, case.modalfilter, _mf(case), (0, 0, 0, 0, 2, 2))
    o.prepare()
    return o


def make_oracle_sphere(case):
    assert case.cs.ntile == 1, "the oracle steps whole panels"
    c = case.consts
    panels = []
    for P, m in enumerate(case.cs.panels):
        o = Oracle(case.p, m.NeX, m.NeY, m.NeZ, lumped=case.elem.lumped, FZ=case.FZ,
                   panel=dict(panelID=P + 1, ztop=case.ztop, RPlanet=c["RPlanet"]))
        o.set_consts(c)
        for k, v in case.fields[P].items():
            o.arr(k)[:] = v.reshape(-1)
        _dry(o, c)
        o.setup_dyn(case.eqs, case.tinteg, case.dt, case.modalfilter, _mf(case), (0, 0, 0, 0, 2, 2))
        if case.sponge:
            o.set_sponge(**case.sponge)
        o.prepare()
        panels.append(o)
    s = OracleSphere(panels)
    s.exchange_aux()
    return s
