"""NumPy transliteration of the per-thread arithmetic of fe_project_b200/csrc/tracer.cu (index maps of the face nodes, outward
normals, flux, FCT coefficient, limited face flux, tensor-product Div with the 1D lift weights, density-weighted low-storage RK,
filter passes, TMAR) held against the oracle.  It cannot replace running the kernels -- synchronisation and shared-memory layout
are not covered -- but the kernels were committed before they could run on hardware, and this guards every formula and index in
them on the CPU.  Keep the two in step: the function below mirrors trc_alphdens_kernel / trc_fct_kernel / trc_stage_kernel line by
line, on the arrays fedg_trcadv_update hands to them."""
import numpy as np
import pytest

from cases import DensityCurrentCase, rel_l2
from fe_project_b200.dyncore import rk_tables


def face_node(f, fp, n):
    a, b, n2 = fp % n, fp // n, n * n
    return [a + b * n2, (n - 1) + a * n + b * n2, a + (n - 1) * n + b * n2, a * n + b * n2, fp, fp + (n - 1) * n2][f]


NORMAL = [(0, -1, 0), (1, 0, 0), (0, 1, 0), (-1, 0, 0), (0, 0, -1), (0, 0, 1)]


def emulate(case, o, q_host, scheme, dt, nsteps, filt, disable_limiter):
    e, m = case.elem, case.mesh
    n1, Np, Nfp, NfpTot, Ne = e.np1, e.Np, e.Nfp, e.NfpTot, m.Ne
    nint = Np * Ne
    # what the device holds: state with halo + boundary condition, aux halo, maps 0-based, per-element / per-face scales
    st = {k: o.arr(k).copy() for k in ("DDENS", "MOMX", "MOMY", "MOMZ", "DENS_hyd")}
    for k in ("DDENS", "DENS_hyd"):
        m.exchange_halo_numpy(st[k])
    o.piece("exchange"); o.piece("bc")                                   # oracle pieces: halo + ApplyBC_PROGVARS on its own state
    mfx, mfy, mfz = o.arr("MOMX").copy(), o.arr("MOMY").copy(), o.arr("MOMZ").copy()
    vmapP = m.VMapP.reshape(Ne, NfpTot)
    D = o.arr("D1D").reshape(n1, n1)
    Lw = o.arr("lift1d").reshape(n1, 2)
    w3 = o.arr("IntWeight")
    w1d = w3[:n1] * 2.0 / w3[:n1].sum()
    jac = o.arr("J")
    esc = np.stack([o.arr("E11").reshape(Ne, Np)[:, 0], o.arr("E22").reshape(Ne, Np)[:, 0], o.arr("E33").reshape(Ne, Np)[:, 0]])
    fsc = o.arr("Fscale").reshape(Ne, NfpTot)[:, ::Nfp].T.copy()          # [6][Ne]
    Fh = e.filter1d(0.0, 1.0, 16) if filt else np.eye(n1)
    Fv = Fh
    ddens, dh = st["DDENS"], st["DENS_hyd"]
    fidx = np.array([[face_node(f, fp, n1) for fp in range(Nfp)] for f in range(6)])          # [6][Nfp]
    ke_off = (np.arange(Ne) * Np)[:, None, None]
    iM = (ke_off + fidx[None]).reshape(Ne, NfpTot)
    iP = vmapP
    nrm = np.repeat(np.array(NORMAL, dtype=float), Nfp, axis=0)                               # [NfpTot][3]
    # trc_alphdens_kernel
    densM, densP = ddens[iM] + dh[iM], ddens[iP] + dh[iP]
    FM = mfx[iM] * nrm[:, 0] + mfy[iM] * nrm[:, 1] + mfz[iM] * nrm[:, 2]
    FP = mfx[iP] * nrm[:, 0] + mfy[iP] * nrm[:, 1] + mfz[iP] * nrm[:, 2]
    alpha = np.maximum(np.abs(FM / densM), np.abs(FP / densP))
    alphM, alphP = alpha * densM, alpha * densP
    W2 = (w1d[None, :] * w1d[:, None]).reshape(-1)                                            # w1d[a] * w1d[b], fp = a + b n
    tabs = rk_tables(scheme)
    ns = tabs["nstage"]
    sig, gam, aex = np.array(tabs["sig"]).reshape(ns + 1, ns), np.array(tabs["gam"]).reshape(ns + 1, ns), np.array(tabs["a_ex"]).reshape(ns, ns)
    q = q_host.copy()
    fct = np.ones_like(q)
    var0, vartmp = np.zeros(nint), np.zeros(nint)
    dens = (dh + ddens)[:nint]
    wq = jac[:nint] * np.tile(w3, Ne)
    for _ in range(nsteps):
        for s in range(ns):
            m.exchange_halo_numpy(q)
            QM, QP = q[iM], q[iP]
            num = 0.5 * ((QP * FP + QM * FM) - alphP * QP + alphM * QM)
            outw = np.zeros((Ne, 6))
            for f in range(6):
                acc = np.zeros(Ne)
                for fp in range(Nfp):                                     # ascending face-node order, one accumulator per face
                    acc = acc + W2[fp] * (jac[iM[:, f * Nfp + fp]] * fsc[f] * num[:, f * Nfp + fp])
                outw[:, f] = acc
            c_ssm1 = aex[s].sum()
            if not disable_limiter:
                dttmp = dt * gam[s + 1, s] / sig[s + 1, s]
                net = np.maximum(0.0, outw).sum(axis=1)
                Qs = (wq * (dens * q[:nint] / dttmp)).reshape(Ne, Np).sum(axis=1)
                fct[:nint] = np.repeat(np.maximum(0.0, np.minimum(1.0, Qs / (net + 1e-10))), Np)
            m.exchange_halo_numpy(fct)
            RM, RP = fct[iM], fct[iP]
            sgn = np.copysign(1.0, np.repeat(outw, Nfp, axis=1))
            dele = np.repeat(fsc.T, Nfp, axis=1) * (num * 0.5 * (RP + RM - (RP - RM) * sgn) - QM * FM)        # [Ne][NfpTot]
            qi = q[:nint].reshape(Ne, n1, n1, n1)                          # [ke][k][j][i]
            Fx, Fy, Fz = (mf[:nint].reshape(Ne, n1, n1, n1) * qi for mf in (mfx, mfy, mfz))
            dx = np.einsum("il,ekjl->ekji", D, Fx)
            dy = np.einsum("jl,ekli->ekji", D, Fy)
            dz = np.einsum("kl,elji->ekji", D, Fz)
            dl = dele.reshape(Ne, 6, n1, n1)                                # [ke][f][b][a]
            lift = (Lw[None, None, :, None, 0] * dl[:, 0][:, :, None, :] + Lw[None, None, None, :, 1] * dl[:, 1][:, :, :, None]
                    + Lw[None, None, :, None, 1] * dl[:, 2][:, :, None, :] + Lw[None, None, None, :, 0] * dl[:, 3][:, :, :, None]
                    + Lw[None, :, None, None, 0] * dl[:, 4][:, None, :, :] + Lw[None, :, None, None, 1] * dl[:, 5][:, None, :, :])
            tend = -(esc[0][:, None, None, None] * dx + esc[1][:, None, None, None] * dy + esc[2][:, None, None, None] * dz + lift).reshape(-1)
            sig_ss, gam_ss, sig_Ns, gam_Ns = sig[s + 1, s], dt * gam[s + 1, s], sig[ns, s], dt * gam[ns, s]
            qs = q[:nint]
            if s == ns - 1:
                qn = (vartmp + sig_ss * qs * dens + gam_ss * tend) / dens
            else:
                if s == 0:
                    var0, vartmp = qs * dens, np.zeros(nint)
                if abs(sig_Ns) > 2.2e-16 or abs(gam[ns, s]) > 2.2e-16:
                    vartmp = vartmp + sig_Ns * qs * dens + gam_Ns * tend
                qn = ((1.0 - sig_ss) * var0 + sig_ss * qs * dens + gam_ss * tend) / dens
            if s == ns - 1 and filt:
                t = (dens * qn).reshape(Ne, n1, n1, n1)
                t = np.einsum("il,ekjl->ekji", Fh, t)
                t = np.einsum("jl,ekli->ekji", Fh, t)
                t = np.einsum("kl,elji->ekji", Fv, t)
                qn = t.reshape(-1) / dens
            if s == ns - 1 and not disable_limiter:
                w = wq * dens
                Q0 = (w * qn).reshape(Ne, Np).sum(axis=1)
                Q1 = (w * np.maximum(0.0, qn)).reshape(Ne, Np).sum(axis=1)
                qn = np.repeat(Q0 / (Q1 + 1e-32), Np) * np.maximum(0.0, qn)
            q[:nint] = qn
    return q


@pytest.mark.parametrize("p,dims,limiter_off,filt,positive", [(3, (3, 2, 3), True, False, True), (3, (3, 2, 3), False, True, False),
                                                               (7, (2, 1, 2), False, False, False)])
def test_kernel_arithmetic_matches_the_oracle(p, dims, limiter_off, filt, positive):
    case = DensityCurrentCase(p=p, NeX=dims[0], NeY=dims[1], NeZ=dims[2], perturb=20.0, dt=0.1, intrp_order=min(11, p + 4), modalfilter=False)
    o = case.make_oracle()
    m, e = case.mesh, case.elem
    n, N = m.Ne * e.Np, m.NeA * e.Np
    x, y, z = (m.pos_en[k].reshape(-1) for k in range(3))
    prof = np.sin(2 * np.pi * x / 25.6e3) * np.cos(2 * np.pi * y / 6.4e3) * np.sin(np.pi * z / 6.4e3)
    q0 = np.zeros(N)
    q0[:n] = 1.0 + 0.5 * prof if positive else np.maximum(0.0, prof)
    qo = q0.copy()
    o2 = case.make_oracle()
    o2.trcadv_update(qo, "ERK_SSP_3s3o", 10.0, nsteps=3, modalfilter=(0.0, 1.0, 16, 0.0, 1.0, 16) if filt else None, disable_limiter=limiter_off)
    qe = emulate(case, o, q0, "ERK_SSP_3s3o", 10.0, 3, filt, limiter_off)
    assert rel_l2(qe[:n], qo[:n]) <= 1e-12
    assert rel_l2(qo[:n], q0[:n]) > 1e-2
