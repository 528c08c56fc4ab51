"""Tracer advection with a prescribed mass flux (row f4) on the GPU against the CPU oracle; run as a script in its own process
(tests/test_gpu_tracer.py): the kernels of fe_project_b200/csrc/tracer.cu had not been run on hardware when they were committed, so
a device fault here must not take the CUDA context of the other GPU tests with it.  Prints one line per case, exit code 0 = all OK."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    from cases import DensityCurrentCase, rel_l2
    ok_all = True
    mf = (0.0, 1.0, 16, 0.0, 1.0, 16)
    for p, dims, limiter_off, filt, positive in [(3, (3, 2, 3), True, False, True), (3, (3, 2, 3), False, True, False),
                                                  (7, (2, 2, 2), False, False, False), (7, (2, 1, 2), True, True, True)]:
        case = DensityCurrentCase(p=p, NeX=dims[0], NeY=dims[1], NeZ=dims[2], perturb=20.0, dt=0.1, intrp_order=min(11, p + 4), modalfilter=False)
        o = case.make_oracle()
        d = case.make_driver(o)
        m, e = case.mesh, case.elem
        n, N = m.Ne * e.Np, m.NeA * e.Np
        x, y, z = (m.pos_en[k].reshape(-1) for k in range(3))
        prof = np.sin(2 * np.pi * x / 25.6e3) * np.cos(2 * np.pi * y / 6.4e3) * np.sin(np.pi * z / 6.4e3)
        q0 = 1.0 + 0.5 * prof if positive else np.maximum(0.0, prof)          # zeros in half of the domain: the limiters act
        qo = np.zeros(N); qo[:n] = q0
        qg = qo.copy()
        dt, nsteps = 10.0, 5
        o.trcadv_update(qo, "ERK_SSP_3s3o", dt, nsteps=nsteps, modalfilter=mf if filt else None, disable_limiter=limiter_off)
        d.trcadv_init("ERK_SSP_3s3o", dt, MODALFILTER_FLAG=filt, disable_limiter=limiter_off)
        d.trcadv_update(qg, nsteps)
        err = rel_l2(qg[:n], qo[:n])
        moved = rel_l2(qo[:n], q0)
        ok = np.isfinite(err) and err <= 1e-10 and moved > 1e-3
        ok_all = ok_all and ok
        print(f"tracer parity p={p} dims={dims} limiter={'off' if limiter_off else 'on'} filter={filt}: rel L2 {err:.3e} "
              f"(field changed by {moved:.2e}, min {qg[:n].min():.3e}) -> {'OK' if ok else 'FAIL'}", flush=True)
    # ---- coupled mode: dynamics step (mass flux and alphDens averaged over the stages) followed by one tracer step, three times
    for eqs, tinteg, dt, q_uniform in [("NONHYDRO3D_HEVE", "ERK_SSP_4s3o", 0.2, True), ("NONHYDRO3D_HEVE", "ERK_SSP_3s3o", 0.2, False),
                                       ("NONHYDRO3D_HEVI", "IMEX_ARK232", 0.5, False), ("NONHYDRO3D_HEVI", "IMEX_ARK324", 0.5, False)]:
        case = DensityCurrentCase(p=7, NeX=2, NeY=2, NeZ=3, perturb=2.0, dt=dt, eqs=eqs, tinteg=tinteg, modalfilter=not q_uniform)   # the filter of the step changes DDENS after DDENS_TRC was taken: a uniform q is then rescaled
        o = case.make_oracle()
        o.set_tracer_coupling(True)
        d = case.make_driver(o)
        d.trcadv_init("ERK_SSP_3s3o", dt, MODALFILTER_FLAG=False, disable_limiter=q_uniform)
        d.trcadv_couple(True)
        m, e = case.mesh, case.elem
        n, N = m.Ne * e.Np, m.NeA * e.Np
        x, y, z = (m.pos_en[k].reshape(-1) for k in range(3))
        prof = np.sin(2 * np.pi * x / 25.6e3) * np.cos(2 * np.pi * y / 6.4e3) * np.sin(np.pi * z / 6.4e3)
        q0 = np.full(n, 0.7) if q_uniform else np.maximum(0.0, prof)
        qo = np.zeros(N); qo[:n] = q0
        qg = qo.copy()
        for _ in range(3):
            o.update(1); d.Update(1)
            o.trcadv_update_coupled(qo, "ERK_SSP_3s3o", dt, disable_limiter=q_uniform)
            d.trcadv_update_coupled(qg)
        err = rel_l2(qg[:n], qo[:n])
        g = d.get_prog()
        derr = max(rel_l2(g[k][:n], o.arr(k)[:n]) for k in ("DDENS", "MOMX", "MOMZ", "DRHOT"))
        ok = np.isfinite(err) and err <= 1e-10 and derr <= 1e-10
        if q_uniform:                                          # a uniform mixing ratio follows the density of the dynamics exactly
            ok = ok and np.abs(qg[:n] - 0.7).max() <= 1e-13
        ok_all = ok_all and ok
        print(f"coupled tracer parity {eqs} {tinteg}: rel L2 {err:.3e}, dynamics {derr:.1e}, "
              f"uniform-q deviation {np.abs(qg[:n] - 0.7).max() if q_uniform else float('nan'):.1e} -> {'OK' if ok else 'FAIL'}", flush=True)
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
