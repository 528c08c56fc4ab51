// TEST INFRASTRUCTURE ONLY -- CPU restatement of FE-Project's DG tracer advection ("next").
// No device kernel exists for this row yet; the restatement comes first so that the kernel has something to be held against.
//
//   fluid_dyn_solver/scale_atm_dyn_dgm_trcadvect3d_heve.F90   (cal_tend :149-231, calc_fct_coef :234-306, TMAR :311-340,
//                                                              cal_alphdens_advtest :404-455, get_delflux_generalhvc :554-674,
//                                                              get_netOutwardFlux_generalhvc :678-777, Init / FaceIntMat :75-137)
//   fluid_dyn_solver/scale_atm_dyn_dgm_driver_trcadv3d.F90    (update :312-559, the ONLY_TRACERADV_FLAG branch: prescribed mass flux)
//   fluid_dyn_solver/scale_atm_dyn_dgm_modalfilter.F90        (tracer_modalfilter_apply :232-270)
//   common/scale_timeint_rk.F90                               (rk_advance_trcvar_low_storage2D)
//
// PARITY STATUS: the reference holds no golden vectors for the tracer path either; the restatement is cross-checked in
// tests/test_oracle_tracer.py (constant preservation, conservation of tracer mass, positivity under the FCT + TMAR limiters,
// and equality with the independently restated sample/advect3d kernel when the limiter is off and the density is uniform).
#include <algorithm>
#include <cmath>
#include <stdexcept>

#include "fe_oracle.hpp"

namespace feo {

// IntWeight(f, fp) of atm_dyn_dgm_trcadvect3d_heve_Init (:75-137): LGL surface weights, row f non-zero on its own face nodes
void trc_face_int_weight(const Element& e, vec& W) {
  const int np = e.np, Nfp = e.Nfp, NfpTot = e.NfpTot;
  W.assign(size_t(6) * NfpTot, 0.0);
  for (int f = 0; f < 6; ++f)
    for (int b = 0; b < np; ++b)
      for (int a = 0; a < np; ++a) W[size_t(f) * NfpTot + f * Nfp + a + b * np] = e.w1d[a] * e.w1d[b];
}

// atm_dyn_dgm_trcadvect3d_heve_cal_alphdens_advtest (:404-455)
void trc_cal_alphdens_advtest(const Element& e, const Mesh& m, const double* DDENS, const double* MOMX, const double* MOMY,
                              const double* MOMZ, const double* DENS_hyd, vec& alphM, vec& alphP) {
  const size_t nf = size_t(e.NfpTot) * m.Ne;
  alphM.resize(nf); alphP.resize(nf);
  for (size_t i = 0; i < nf; ++i) {
    const int iM = m.vmapM[i], iP = m.vmapP[i];
    const double densM = DDENS[iM] + DENS_hyd[iM], densP = DDENS[iP] + DENS_hyd[iP];
    const double VelM = (MOMX[iM] * m.nx[i] + MOMY[iM] * m.ny[i] + MOMZ[iM] * m.nz[i]) / densM;
    const double VelP = (MOMX[iP] * m.nx[i] + MOMY[iP] * m.ny[i] + MOMZ[iP] * m.nz[i]) / densP;
    const double alpha = std::max(std::fabs(VelM), std::fabs(VelP));
    alphM[i] = alpha * densM * m.Gsqrt[iM];
    alphP[i] = alpha * densP * m.Gsqrt[iP];
  }
}

namespace {
// face quantities shared by get_delflux_generalhvc (:620-655) and get_netOutwardFlux_generalhvc (:735-765)
struct FaceFlux { double QM, QP, MomFlxM, MomFlxP, numflux; };
inline FaceFlux face_flux(const Element& e, const Mesh& m, int ke, int fp, const double* Q, const double* MX, const double* MY,
                          const double* MZ, const vec& alphM, const vec& alphP) {
  const int np = e.np, Nfp = e.Nfp;
  const size_t f = size_t(fp) + size_t(ke) * e.NfpTot;
  const int iM = m.vmapM[f], iP = m.vmapP[f];
  // IndexH2Dto3D_bnd: horizontal node of a face node
  const int fl = fp % Nfp, fc = fp / Nfp;
  const int h2d = fc < 4 ? (fc == 0 ? fl % np : fc == 1 ? (np - 1) + (fl % np) * np : fc == 2 ? fl % np + (np - 1) * np : (fl % np) * np) : fl;
  const double gH = m.GsqrtH[h2d + size_t(m.emap2d[ke]) * Nfp];
  const double GsM = m.Gsqrt[iM], GsP = m.Gsqrt[iP];
  const double GvM = GsM / gH, GvP = GsP / gH;
  const double gMXM = GsM * MX[iM], gMXP = GsP * MX[iP], gMYM = GsM * MY[iM], gMYP = GsP * MY[iP], gMZM = GsM * MZ[iM], gMZP = GsP * MZ[iP];
  FaceFlux r;
  r.QM = Q[iM]; r.QP = Q[iP];
  r.MomFlxM = (gMXM * m.nx[f] + gMYM * m.ny[f] + ((gMZM / GvM + m.G13[iM] * gMXM + m.G23[iM] * gMYM) * m.nz[f]));
  r.MomFlxP = (gMXP * m.nx[f] + gMYP * m.ny[f] + ((gMZP / GvP + m.G13[iP] * gMXP + m.G23[iP] * gMYP) * m.nz[f]));
  r.numflux = 0.5 * ((r.QP * r.MomFlxP + r.QM * r.MomFlxM) - alphP[f] * r.QP + alphM[f] * r.QM);
  return r;
}
// sparsemat_matmul(FaceIntMat, J(iM) * Fscale * numflux): one sum per face, fp ascending
inline void outward_flux(const Element& e, const Mesh& m, int ke, const vec& W, const double* numflux, double out[6]) {
  const int NfpTot = e.NfpTot, Nfp = e.Nfp;
  for (int f = 0; f < 6; ++f) {
    double s = 0.0;
    for (int fp = f * Nfp; fp < (f + 1) * Nfp; ++fp) {
      const size_t g = size_t(fp) + size_t(ke) * NfpTot;
      s += W[size_t(f) * NfpTot + fp] * (m.J[m.vmapM[g]] * m.Fscale[g] * numflux[fp]);
    }
    out[f] = s;
  }
}
}  // namespace

// atm_dyn_dgm_trcadvect3d_heve_get_netOutwardFlux_generalhvc (:678-777)
void trc_net_outward_flux(const Element& e, const Mesh& m, const vec& W, const double* Q, const double* MX, const double* MY,
                          const double* MZ, const vec& alphM, const vec& alphP, vec& net) {
  net.resize(m.Ne);
  vec numflux(e.NfpTot);
  for (int ke = 0; ke < m.Ne; ++ke) {
    for (int fp = 0; fp < e.NfpTot; ++fp) numflux[fp] = face_flux(e, m, ke, fp, Q, MX, MY, MZ, alphM, alphP).numflux;
    double o6[6];
    outward_flux(e, m, ke, W, numflux.data(), o6);
    double s = 0.0;
    for (int f = 0; f < 6; ++f) s += std::max(0.0, o6[f]);
    net[ke] = s;
  }
}

// atm_dyn_dgm_trcadvect3d_heve_calc_fct_coef (:234-306); fct (Np,NeA), interior part written
void trc_calc_fct_coef(const Element& e, const Mesh& m, const vec& W, const double* Q, const double* MX, const double* MY,
                       const double* MZ, const double* RHOQ_tp, const vec& alphM, const vec& alphP, const double* DENS_hyd,
                       const double* DDENS, const double* DDENS0, double rk_c_ssm1, double dt, bool disable_limiter, double* fct) {
  const int Np = e.Np;
  if (disable_limiter) {
    for (size_t i = 0; i < size_t(Np) * m.Ne; ++i) fct[i] = 1.0;
    return;
  }
  vec net;
  trc_net_outward_flux(e, m, W, Q, MX, MY, MZ, alphM, alphP, net);
  for (int ke = 0; ke < m.Ne; ++ke) {
    const size_t o = size_t(ke) * Np;
    double Qs = 0.0;
    for (int p = 0; p < Np; ++p) {
      const double dens_ssm1 = DENS_hyd[o + p] + (1.0 - rk_c_ssm1) * DDENS0[o + p] + rk_c_ssm1 * DDENS[o + p];
      Qs += m.Gsqrt[o + p] * m.J[o + p] * e.IntWeight[p] * (dens_ssm1 * Q[o + p] / dt + RHOQ_tp[o + p]);
    }
    const double c = std::max(0.0, std::min(1.0, Qs / (net[ke] + 1.0e-10)));
    for (int p = 0; p < Np; ++p) fct[o + p] = c;
  }
}

// atm_dyn_dgm_trcadvect3d_heve_get_delflux_generalhvc (:554-674) + cal_tend (:149-231)
void trc_cal_tend(const Element& e, const Mesh& m, const vec& W, const double* Q, const double* MX, const double* MY, const double* MZ,
                  const vec& alphM, const vec& alphP, const double* fct, const double* RHOQ_tp, double* Q_dt) {
  const int Np = e.Np, Nfp = e.Nfp, NfpTot = e.NfpTot;
#pragma omp parallel
  {
    vec numflux(NfpTot), del(NfpTot), Flux(size_t(Np) * 3), DFlux(size_t(Np) * 4);
    std::vector<FaceFlux> ff(NfpTot);
#pragma omp for
    for (int ke = 0; ke < m.Ne; ++ke) {
      const size_t o = size_t(ke) * Np;
      for (int fp = 0; fp < NfpTot; ++fp) { ff[fp] = face_flux(e, m, ke, fp, Q, MX, MY, MZ, alphM, alphP); numflux[fp] = ff[fp].numflux; }
      double o6[6];
      outward_flux(e, m, ke, W, numflux.data(), o6);
      for (int fp = 0; fp < NfpTot; ++fp) {
        const size_t g = size_t(fp) + size_t(ke) * NfpTot;
        const double RM = fct[m.vmapM[g]], RP = fct[m.vmapP[g]];
        const double sgn = std::copysign(1.0, o6[fp / Nfp]);           // sign(1.0, outward_flux_tmp(f))
        del[fp] = m.Fscale[g] * (numflux[fp] * 0.5 * (RP + RM - (RP - RM) * sgn) - ff[fp].QM * ff[fp].MomFlxM);
      }
      const int ke2d = m.emap2d[ke];
      for (int p = 0; p < Np; ++p) {
        const double G = m.Gsqrt[o + p];
        const double RGv = m.GsqrtH[(p % Nfp) + size_t(ke2d) * Nfp] * (1.0 / G);
        Flux[p] = G * MX[o + p] * Q[o + p];
        Flux[Np + p] = G * MY[o + p] * Q[o + p];
        Flux[2 * Np + p] = G * (MZ[o + p] * RGv + m.G13[o + p] * MX[o + p] + m.G23[o + p] * MY[o + p]) * Q[o + p];
      }
      op_div(e, Flux.data(), del.data(), DFlux.data());
      for (int p = 0; p < Np; ++p)
        Q_dt[o + p] = -(m.E11[o + p] * DFlux[p] + m.E22[o + p] * DFlux[Np + p] + m.E33[o + p] * DFlux[2 * Np + p] + DFlux[3 * Np + p]) /
                          m.Gsqrt[o + p] +
                      RHOQ_tp[o + p];
    }
  }
}

// atm_dyn_dgm_trcadvect3d_TMAR (:311-340): truncation of negative values + mass-aware rescaling per element
void trc_tmar(const Element& e, const Mesh& m, const double* DENS_hyd, const double* DDENS, double* Q) {
  const int Np = e.Np;
  for (int ke = 0; ke < m.Ne; ++ke) {
    const size_t o = size_t(ke) * Np;
    double Q0 = 0.0, Q1 = 0.0;
    for (int p = 0; p < Np; ++p) {
      const double w = m.Gsqrt[o + p] * m.J[o + p] * e.IntWeight[p] * (DENS_hyd[o + p] + DDENS[o + p]);
      Q0 += w * Q[o + p];
      Q1 += w * std::max(0.0, Q[o + p]);
    }
    for (int p = 0; p < Np; ++p) Q[o + p] = Q0 / (Q1 + 1.0e-32) * std::max(0.0, Q[o + p]);
  }
}

// atm_dyn_dgm_tracer_modalfilter_apply (modalfilter.F90:232-270); `ef` carries the TRACER filter matrices
void trc_modalfilter(const Element& ef, const Mesh& m, const double* DENS_hyd, const double* DDENS, double* Q) {
  const int Np = ef.Np;
  vec tmp(Np), work(Np), out(Np), wgt(Np);
  for (int ke = 0; ke < m.Ne; ++ke) {
    const size_t o = size_t(ke) * Np;
    for (int p = 0; p < Np; ++p) { wgt[p] = m.Gsqrt[o + p] * (DENS_hyd[o + p] + DDENS[o + p]); tmp[p] = wgt[p] * Q[o + p]; }
    op_modal_filter(ef, tmp.data(), work.data(), out.data());
    for (int p = 0; p < Np; ++p) Q[o + p] = out[p] / wgt[p];
  }
}

// rk_advance_trcvar_low_storage2D (common/scale_timeint_rk.F90): the integrator advances rho*q, q is recovered with the density
// interpolated to the stage time (coef_c_ex); var0 / varTmp / tend are (n) work arrays of the tracer integrator
void rk_advance_trcvar_low_storage(const RKScheme& sc, double dt, int stage /*0-based*/, size_t n, double* q, const double* DDENS,
                                   const double* DDENS0, const double* DENS_hyd, double* var0, double* varTmp, const double* tend) {
  if (!sc.low_storage || sc.imex) throw std::runtime_error("tracer advance: low-storage explicit schemes only");
  const int ns = sc.nstage;
  const double EPS = 2.220446e-16;
  const double sig_ss = sc.SIG(stage + 1, stage), sig_Ns = sc.SIG(ns, stage);
  const double gam_ss = dt * sc.GAM(stage + 1, stage), gam_Ns = dt * sc.GAM(ns, stage);
  const double c_ssm1 = sc.c_ex[stage];
  if (stage == ns - 1) {
    for (size_t i = 0; i < n; ++i) {
      const double dens_ssm1 = DENS_hyd[i] + DDENS0[i] + c_ssm1 * (DDENS[i] - DDENS0[i]);
      q[i] = (varTmp[i] + sig_ss * q[i] * dens_ssm1 + gam_ss * tend[i]) / (DENS_hyd[i] + DDENS[i]);
    }
    return;
  }
  const double c_ss = sc.c_ex[stage + 1];
  if (stage == 0)
    for (size_t i = 0; i < n; ++i) { var0[i] = q[i] * (DENS_hyd[i] + DDENS0[i]); varTmp[i] = 0.0; }
  if (std::fabs(sig_Ns) > EPS || std::fabs(gam_Ns) > EPS)
    for (size_t i = 0; i < n; ++i) {
      const double dens_ssm1 = DENS_hyd[i] + DDENS0[i] + c_ssm1 * (DDENS[i] - DDENS0[i]);
      varTmp[i] = varTmp[i] + sig_Ns * q[i] * dens_ssm1 + gam_Ns * tend[i];
    }
  for (size_t i = 0; i < n; ++i) {
    const double dens_ssm1 = DENS_hyd[i] + DDENS0[i] + c_ssm1 * (DDENS[i] - DDENS0[i]);
    const double dens_ss = DENS_hyd[i] + DDENS0[i] + c_ss * (DDENS[i] - DDENS0[i]);
    q[i] = ((1.0 - sig_ss) * var0[i] + sig_ss * q[i] * dens_ssm1 + gam_ss * tend[i]) / dens_ss;
  }
}

// atm_dyn_dgm_trcadvect3d_heve_cal_alphdens_dyn (:460-551): dissipation coefficient x density of the dynamics' Rusanov flux at
// the stage state, accumulated with the Butcher weights (horizontal faces: b_ex, vertical faces: b_im for HEVI)
void trc_cal_alphdens_dyn(const Element& e, const Mesh& m, const Consts& c, const DynState& s, double w_h, double w_v, bool is_hevi,
                          vec& alphM, vec& alphP) {
  const int np = e.np, Nfp = e.Nfp, NfpTot = e.NfpTot;
  const double gamm = c.CPdry / c.CVdry;
  for (int ke = 0; ke < m.Ne; ++ke)
    for (int fp = 0; fp < NfpTot; ++fp) {
      const size_t f = size_t(fp) + size_t(ke) * NfpTot;
      const int iM = m.vmapM[f], iP = m.vmapP[f];
      const int fl = fp % Nfp, fc = fp / Nfp;
      const int h2d = fc < 4 ? (fc == 0 ? fl % np : fc == 1 ? (np - 1) + (fl % np) * np : fc == 2 ? fl % np + (np - 1) * np : (fl % np) * np) : fl;
      const size_t h = h2d + size_t(m.emap2d[ke]) * Nfp;
      const double G11 = m.is_global ? m.GIJ11[h] : 1.0, G22 = m.is_global ? m.GIJ22[h] : 1.0;
      const double anx = std::fabs(m.nx[f]), any = std::fabs(m.ny[f]), anz = std::fabs(m.nz[f]);
      const double Gnn = is_hevi ? G11 * anx + G22 * any : G11 * anx + G22 * any + anz;
      const double densM = s.DDENS[iM] + s.DENS_hyd[iM], densP = s.DDENS[iP] + s.DENS_hyd[iP];
      const double VelM = (s.MOMX[iM] * m.nx[f] + s.MOMY[iM] * m.ny[f] + s.MOMZ[iM] * m.nz[f]) / densM;
      const double VelP = (s.MOMX[iP] * m.nx[f] + s.MOMY[iP] * m.ny[f] + s.MOMZ[iP] * m.nz[f]) / densP;
      const double alpha = std::max(std::sqrt(Gnn * gamm * (s.PRES_hyd[iM] + s.DPRES[iM]) / densM) + std::fabs(VelM),
                                    std::sqrt(Gnn * gamm * (s.PRES_hyd[iP] + s.DPRES[iP]) / densP) + std::fabs(VelP));
      const double w = w_h * (anx + any) + w_v * anz;
      alphM[f] = alphM[f] + w * alpha * densM * m.Gsqrt[iM];
      alphP[f] = alphP[f] + w * alpha * densP * m.Gsqrt[iP];
    }
}

// atm_dyn_dgm_trcadvect3d_save_massflux (:343-401), called by the dynamics driver in every RK stage before Advance
// (driver_nonhydro3d.F90:900-917) with tavg_weight_h = b_ex(stage), tavg_weight_v = b_im(stage) (IMEX) or b_ex(stage)
void trc_save_massflux(Driver& d, int stage) {
  const Element& e = d.elem; const Mesh& m = d.mesh; const DynState& s = d.st;
  const size_t nint = size_t(e.Np) * m.Ne, nall = size_t(e.Np) * m.NeA, nf = size_t(e.NfpTot) * m.Ne;
  const double w_h = d.tint.sc.b_ex[stage], w_v = d.tint.sc.imex ? d.tint.sc.b_im[stage] : d.tint.sc.b_ex[stage];
  if (stage == 0) {
    d.MFLX_x.assign(nall, 0.0); d.MFLX_y.assign(nall, 0.0); d.MFLX_z.assign(nall, 0.0);
    d.alphM_tavg.assign(nf, 0.0); d.alphP_tavg.assign(nf, 0.0);
    for (size_t i = 0; i < nint; ++i) { d.MFLX_x[i] = w_h * s.MOMX[i]; d.MFLX_y[i] = w_h * s.MOMY[i]; d.MFLX_z[i] = w_v * s.MOMZ[i]; }
  } else {
    for (size_t i = 0; i < nint; ++i) {
      d.MFLX_x[i] = d.MFLX_x[i] + w_h * s.MOMX[i]; d.MFLX_y[i] = d.MFLX_y[i] + w_h * s.MOMY[i]; d.MFLX_z[i] = d.MFLX_z[i] + w_v * s.MOMZ[i];
    }
  }
  trc_cal_alphdens_dyn(e, m, d.cst, s, w_h, w_v, d.hevi, d.alphM_tavg, d.alphP_tavg);
}

namespace {
struct TrcInputs { vec MFx, MFy, MFz, DD, DD0, alphM, alphP; };
// the per-tracer part of AtmDynDGMDriver_trcadv3d_update (driver_trcadv3d.F90:426-539)
void trc_core(Driver& d, const Element& elem_trcfilter, const RKScheme& sc, double dt, bool modalfilter, bool disable_limiter,
              bool apply_tmar, TrcInputs& in, double* QTRC, const double* RHOQ_tp) {
  const Element& e = d.elem;
  const Mesh& m = d.mesh;
  DynState& s = d.st;
  const size_t nint = size_t(e.Np) * m.Ne, nall = size_t(e.Np) * m.NeA;
  vec W;
  trc_face_int_weight(e, W);
  vec Qtmp(QTRC, QTRC + nall), fct(nall, 1.0), var0(nint), varTmp(nint), tend(nint);
  const int ns = sc.nstage;
  for (int st = 0; st < ns; ++st) {
    m.exchange_halo(e, Qtmp.data());                                             // TRCVAR3D_manager%MeshFieldComm_Exchange
    const double dttmp = dt * sc.GAM(st + 1, st) / sc.SIG(st + 1, st);
    trc_calc_fct_coef(e, m, W, Qtmp.data(), in.MFx.data(), in.MFy.data(), in.MFz.data(), RHOQ_tp, in.alphM, in.alphP, s.DENS_hyd.data(),
                      in.DD.data(), in.DD0.data(), sc.c_ex[st], dttmp, disable_limiter, fct.data());
    m.exchange_halo(e, fct.data());                                              // AUXTRCVAR3D_manager%MeshFieldComm_Exchange
    trc_cal_tend(e, m, W, Qtmp.data(), in.MFx.data(), in.MFy.data(), in.MFz.data(), in.alphM, in.alphP, fct.data(), RHOQ_tp, tend.data());
    rk_advance_trcvar_low_storage(sc, dt, st, nint, Qtmp.data(), in.DD.data(), in.DD0.data(), s.DENS_hyd.data(), var0.data(),
                                  varTmp.data(), tend.data());
    if (st == ns - 1 && modalfilter) trc_modalfilter(elem_trcfilter, m, s.DENS_hyd.data(), in.DD.data(), Qtmp.data());
    if (st == ns - 1 && apply_tmar) trc_tmar(e, m, s.DENS_hyd.data(), in.DD.data(), Qtmp.data());   // ONLY_TRACERADV_FLAG and limiter on
  }
  // QTRC = (DENS_hyd + DDENS_TRC) / (DENS_hyd + DDENS) * QTRC_tmp (:530-537): DDENS may have been filtered after DDENS_TRC was taken
  for (size_t i = 0; i < nint; ++i) QTRC[i] = (s.DENS_hyd[i] + in.DD[i]) / (s.DENS_hyd[i] + s.DDENS[i]) * Qtmp[i];
}
// exchange of the mass fluxes + ApplyBC_PROGVARS_lc on them (:404-420)
void trc_massflux_halo(Driver& d, TrcInputs& in) {
  const Element& e = d.elem; const Mesh& m = d.mesh;
  for (vec* f : {&in.MFx, &in.MFy, &in.MFz}) m.exchange_halo(e, f->data());
  DynState t = d.st;
  t.MOMX = in.MFx; t.MOMY = in.MFy; t.MOMZ = in.MFz; t.DDENS = in.DD;
  apply_bc_progvars(e, m, d.bnd, t);
  in.MFx = t.MOMX; in.MFy = t.MOMY; in.MFz = t.MOMZ;
}
}  // namespace

// AtmDynDGMDriver_trcadv3d_update (driver_trcadv3d.F90:312-559) after a dynamics step that accumulated the mass fluxes
// (Driver::tracer = true): DDENS0_TRC = density at the start of the step, DDENS_TRC = density after the RK loop
// (driver_nonhydro3d.F90:926-937).  Without the negative fixer and the pressure / specific-heat update that follow in the model.
void trcadv_update_coupled(Driver& d, const Element& elem_trcfilter, const RKScheme& sc, double dt, bool modalfilter,
                           bool disable_limiter, double* QTRC, const double* RHOQ_tp) {
  if (!d.tracer || d.MFLX_x.empty()) throw std::runtime_error("no accumulated mass flux: enable the tracer coupling and run a dynamics step first");
  TrcInputs in;
  in.MFx = d.MFLX_x; in.MFy = d.MFLX_y; in.MFz = d.MFLX_z; in.DD = d.DENS_TRC; in.DD0 = d.DENS0_TRC;
  in.alphM = d.alphM_tavg; in.alphP = d.alphP_tavg;
  trc_massflux_halo(d, in);
  trc_core(d, elem_trcfilter, sc, dt, modalfilter, disable_limiter, false, in, QTRC, RHOQ_tp);
}

// The same with ONLY_TRACERADV_FLAG = .true. (:376-399): the mass flux is the momentum of the (frozen) dynamical state,
// DDENS_TRC = DDENS0_TRC = DDENS; one tracer QTRC (Np,NeA) advanced by one step.
void trcadv_update_advtest(Driver& d, const Element& elem_trcfilter, const RKScheme& sc, double dt, bool modalfilter,
                           bool disable_limiter, double* QTRC, const double* RHOQ_tp) {
  const Element& e = d.elem;
  const Mesh& m = d.mesh;
  DynState& s = d.st;
  // The reference evaluates alphDens on the halo of the dynamical state as the caller left it; here that halo is made valid first
  // (exchange + boundary condition), so that the function is defined by its interior inputs alone.
  TrcInputs in;
  in.MFx = s.MOMX; in.MFy = s.MOMY; in.MFz = s.MOMZ; in.DD = s.DDENS;
  vec DH(s.DENS_hyd);
  m.exchange_halo(e, in.DD.data()); m.exchange_halo(e, DH.data());
  trc_massflux_halo(d, in);
  in.DD0 = in.DD;
  trc_cal_alphdens_advtest(e, m, in.DD.data(), in.MFx.data(), in.MFy.data(), in.MFz.data(), DH.data(), in.alphM, in.alphP);
  trc_core(d, elem_trcfilter, sc, dt, modalfilter, disable_limiter, !disable_limiter, in, QTRC, RHOQ_tp);
}

}  // namespace feo
