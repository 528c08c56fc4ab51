// TEST INFRASTRUCTURE ONLY (see fe_oracle.hpp).  Regional HEVE dynamics: pressure, boundary
// conditions, Rusanov flux, explicit tendency, modal filter.
#include "fe_oracle.hpp"

#include <algorithm>

namespace feo {

void DynState::alloc(size_t n, size_t n2d) {
  for (vec* v : {&DDENS, &MOMX, &MOMY, &MOMZ, &DRHOT, &DENS_hyd, &PRES_hyd, &THERM_hyd, &PRES_hyd_ref, &Rtot,
                 &CVtot, &CPtot, &PRES, &DPRES, &DPhydDx, &DPhydDy, &DENS_tp, &MOMX_tp, &MOMY_tp, &MOMZ_tp, &RHOT_tp, &RHOH_p})
    v->assign(n, 0.0);
  CORIOLIS.assign(n2d, 0.0);
}

// fluid_dyn_solver/scale_atm_dyn_dgm_spongelayer.F90:129-220 (add_tend, calc_wdampcoef); cfg.tau / cfg.height already resolved
// as Init does (:97-118: SL_WDAMP_LAYER -> height of the first node of that layer, tau < 0 -> 10 dt)
void sponge_add_tend(const Element& e, const Mesh& m, const SpongeCfg& cfg, const DynState& s, double* dt5[5]) {
  const double PI = 3.14159265358979323846;
  const int Np = e.Np, Nfp = e.Nfp, np = e.np;
  const double sflag = cfg.hveldamp ? 1.0 : 0.0, r_tau = 1.0 / cfg.tau;
  for (int ke = 0; ke < m.Ne; ++ke) {
    const int keZtop = (ke % (m.NeX * m.NeY)) + (m.NeZ - 1) * m.NeX * m.NeY;
    for (int p = 0; p < Np; ++p) {
      const size_t i = size_t(ke) * Np + p;
      const double z = m.pos[2][i], zTop = m.pos[2][size_t(keZtop) * Np + (p % Nfp) + (np - 1) * Nfp];
      const double coef = 0.25 * r_tau * (1.0 + (z - cfg.height >= 0.0 ? 1.0 : -1.0)) * (1.0 - std::cos(PI * (z - cfg.height) / (zTop - cfg.height)));
      dt5[MOMX_VID][i] = dt5[MOMX_VID][i] - sflag * coef * s.MOMX[i];
      dt5[MOMY_VID][i] = dt5[MOMY_VID][i] - sflag * coef * s.MOMY[i];
      dt5[MOMZ_VID][i] = dt5[MOMZ_VID][i] - coef * s.MOMZ[i];
    }
  }
}

// fluid_dyn_solver/scale_atm_dyn_dgm_driver_nonhydro3d.F90:1098-1178 (add_phy_tend, CPU branch)
void add_phy_tend(const Element& e, const Mesh& m, const Consts& c, const DynState& s, bool entot_conserve, double* dt5[5]) {
  const double rP0 = 1.0 / c.PRES00;
  const size_t n = size_t(e.Np) * m.Ne;
#pragma omp parallel for
  for (size_t i = 0; i < n; ++i) {
    dt5[DENS_VID][i] = dt5[DENS_VID][i] + s.DENS_tp[i];
    dt5[MOMZ_VID][i] = dt5[MOMZ_VID][i] + s.MOMZ_tp[i];
    dt5[MOMX_VID][i] = dt5[MOMX_VID][i] + s.MOMX_tp[i];
    dt5[MOMY_VID][i] = dt5[MOMY_VID][i] + s.MOMY_tp[i];
    const double EXNER = std::pow(s.PRES[i] * rP0, s.Rtot[i] / s.CPtot[i]);
    if (entot_conserve) dt5[THERM_VID][i] = dt5[THERM_VID][i] + s.RHOH_p[i] + (s.CPtot[i] * EXNER) * s.RHOT_tp[i];
    else dt5[THERM_VID][i] = dt5[THERM_VID][i] + s.RHOT_tp[i] + s.RHOH_p[i] / (s.CPtot[i] * EXNER);
  }
}

// fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_common.F90:428-479 (CPU branch)
void drhot2pres(const Element& e, const Mesh& m, const Consts& c, DynState& s) {
  const double rP0 = 1.0 / c.PRES00;
  const size_t n = size_t(e.Np) * m.Ne;
#pragma omp parallel for
  for (size_t i = 0; i < n; ++i) {
    double rhot = s.THERM_hyd[i] + s.DRHOT[i];
    s.PRES[i] = c.PRES00 * std::pow(s.Rtot[i] * rP0 * rhot, s.CPtot[i] / s.CVtot[i]);
    s.DPRES[i] = s.PRES[i] - s.PRES_hyd[i];
  }
}

// common.F90:584-619
void calc_rhot_hyd(const Element& e, const Mesh& m, const Consts& c, DynState& s) {
  const size_t n = size_t(e.Np) * m.Ne;
  for (size_t i = 0; i < n; ++i)
    s.THERM_hyd[i] = c.PRES00 / c.Rdry * std::pow(s.PRES_hyd[i] / c.PRES00, c.CVdry / c.CPdry);
}

// common.F90:624-777 (calc_phyd_hgrad_lc + get_phyd_hgrad_numflux_generalhvc), gam = 1 on the cube
void calc_phyd_hgrad(const Element& e, const Mesh& m, DynState& s) {
  const int Np = e.Np, NfpTot = e.NfpTot, np = e.np;
  vec delx(NfpTot), dely(NfpTot), Flux(3 * Np), Fz(Np), D1(4 * Np), D2z(Np), L2(Np), RGsqrtV(Np);
  for (int ke = 0; ke < m.Ne; ++ke) {
    int ke2d = m.emap2d[ke];
    for (int fp = 0; fp < NfpTot; ++fp) {
      size_t f = size_t(fp) + size_t(ke) * NfpTot;
      int iM = m.vmapM[f], iP = m.vmapP[f];
      int fl = fp % e.Nfp, face = fp / e.Nfp;
      int h2d = face < 4 ? (face == 0 ? fl % np : face == 1 ? (np - 1) + (fl % np) * np
                            : face == 2 ? fl % np + (np - 1) * np : (fl % np) * np) : fl;
      double gh = m.GsqrtH[h2d + size_t(ke2d) * e.Nfp];
      double GvM = m.Gsqrt[iM] / gh, GvP = m.Gsqrt[iP] / gh;
      double dpM = s.PRES_hyd[iM] - s.PRES_hyd_ref[iM], dpP = s.PRES_hyd[iP] - s.PRES_hyd_ref[iP];
      double t1 = m.Fscale[f] * 0.5 * GvP * dpP, t2 = m.Fscale[f] * 0.5 * GvM * dpM;
      delx[fp] = (m.nx[f] + m.G13[iP] * m.nz[f]) * t1 - (m.nx[f] + m.G13[iM] * m.nz[f]) * t2;
      dely[fp] = (m.ny[f] + m.G23[iP] * m.nz[f]) * t1 - (m.ny[f] + m.G23[iM] * m.nz[f]) * t2;
    }
    for (int p = 0; p < Np; ++p) {
      size_t i = size_t(p) + size_t(ke) * Np;
      double Gv = m.Gsqrt[i] / m.GsqrtH[(p % e.Nfp) + size_t(ke2d) * e.Nfp];
      RGsqrtV[p] = 1.0 / Gv;
      Flux[p] = Gv * (s.PRES_hyd[i] - s.PRES_hyd_ref[i]);
      Flux[Np + p] = Flux[p];
      Flux[2 * Np + p] = m.G13[i] * Flux[p];
      Fz[p] = m.G23[i] * Flux[p];
    }
    op_div(e, Flux.data(), delx.data(), D1.data());
    op_dz(e, Fz.data(), D2z.data());
    op_lift(e, dely.data(), L2.data());
    for (int p = 0; p < Np; ++p) {
      size_t i = size_t(p) + size_t(ke) * Np;
      double gx = m.E11[i] * D1[p] + m.E33[i] * D1[2 * Np + p] + D1[3 * Np + p];
      double gy = m.E22[i] * D1[Np + p] + m.E33[i] * D2z[p] + L2[p];
      s.DPhydDx[i] = gx * RGsqrtV[p];
      s.DPhydDy[i] = gy * RGsqrtV[p];
    }
  }
}

// fluid_dyn_solver/scale_atm_dyn_dgm_bnd.F90:270-367 (GIJ = identity on the cube)
void apply_bc_progvars(const Element& e, const Mesh& m, const BndCfg& b, DynState& s) {
  const int NfpTot = e.NfpTot, Np = e.Np, np = e.np;
  for (int ke = 0; ke < m.Ne; ++ke)
    for (int fp = 0; fp < NfpTot; ++fp) {
      size_t f = size_t(fp) + size_t(ke) * NfpTot;
      int iP = m.vmapP[f], i_ = iP - Np * m.Ne;
      if (i_ < 0) continue;
      int face = 0;
      while (i_ >= m.halo_off[face + 1]) ++face;
      // a face carries the BC only if its tile neighbour is the tile itself with the same face (bnd_Init_lc)
      int bc = (m.nbr_face[face] == face) ? b.vel_bc[face] : 0;
      int iM = m.vmapM[f];
      if (bc == 2) {
        int fl = fp % e.Nfp, fc = fp / e.Nfp;
        int h2d = fc < 4 ? (fc == 0 ? fl % np : fc == 1 ? (np - 1) + (fl % np) * np
                            : fc == 2 ? fl % np + (np - 1) * np : (fl % np) * np) : fl;
        double GsqrtV = m.Gsqrt[iM] / m.GsqrtH[h2d + size_t(m.emap2d[ke]) * e.Nfp];
        double G11 = 1.0, G12 = 0.0, G22 = 1.0;
        double G13 = m.G13[iM], G23 = m.G23[iM];
        double MOMW = s.MOMZ[iM] / GsqrtV + G13 * s.MOMX[iM] + G23 * s.MOMY[iM];
        double fac = m.nz[f] * GsqrtV * GsqrtV /
                     (1.0 + G11 * (GsqrtV * G13) * (GsqrtV * G13) + 2.0 * G12 * (GsqrtV * GsqrtV * G13 * G23) +
                      G22 * (GsqrtV * G23) * (GsqrtV * G23));
        double mn = s.MOMX[iM] * m.nx[f] + s.MOMY[iM] * m.ny[f] + MOMW * m.nz[f];
        s.MOMX[iP] = s.MOMX[iM] - 2.0 * mn * (m.nx[f] + fac * (G11 * G13 + G12 * G23));
        s.MOMY[iP] = s.MOMY[iM] - 2.0 * mn * (m.ny[f] + fac * (G12 * G13 + G22 * G23));
        s.MOMZ[iP] = s.MOMZ[iM] - 2.0 * mn * fac / GsqrtV;
      } else if (bc == 3) {
        s.MOMX[iP] = -s.MOMX[iM]; s.MOMY[iP] = -s.MOMY[iM]; s.MOMZ[iP] = -s.MOMZ[iM];
      }
    }
}

// fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_rhot_heve_numflux.F90:946-1138
void heve_numflux_generalvc(const Element& e, const Mesh& m, const Consts& c, const DynState& s, vec& del_flux) {
  const int NfpTot = e.NfpTot;
  const double gamm = c.CPdry / c.CVdry;
  del_flux.resize(size_t(NfpTot) * PRGVAR_NUM * m.Ne);
#pragma omp parallel for
  for (int ke = 0; ke < m.Ne; ++ke) {
    double* df = &del_flux[size_t(ke) * NfpTot * PRGVAR_NUM];
    for (int fp = 0; fp < NfpTot; ++fp) {
      size_t f = size_t(fp) + size_t(ke) * NfpTot;
      const int id[2] = {m.vmapM[f], m.vmapP[f]};
      const double nx = m.nx[f], ny = m.ny[f], nz = m.nz[f];
      double Gs[2], RGv[2], G13[2], G23[2], gDD[2], gMX[2], gMY[2], gMZ[2], gDR[2], Phyd[2], dp[2], gDens[2], gRhot[2], Vel[2];
      for (int t = 0; t < 2; ++t) {
        int i = id[t];
        Gs[t] = m.Gsqrt[i]; RGv[t] = 1.0 / Gs[t]; G13[t] = m.G13[i]; G23[t] = m.G23[i];
        gDD[t] = Gs[t] * s.DDENS[i]; gMX[t] = Gs[t] * s.MOMX[i]; gMY[t] = Gs[t] * s.MOMY[i];
        gMZ[t] = Gs[t] * s.MOMZ[i]; gDR[t] = Gs[t] * s.DRHOT[i];
        Phyd[t] = s.PRES_hyd[i]; dp[t] = s.DPRES[i];
        gDens[t] = gDD[t] + Gs[t] * s.DENS_hyd[i];
        gRhot[t] = Gs[t] * s.THERM_hyd[i] + gDR[t];
        Vel[t] = (gMX[t] * nx + gMY[t] * ny + ((gMZ[t] * RGv[t] + G13[t] * gMX[t] + G23[t] * gMY[t]) * nz)) / gDens[t];
      }
      double t1 = std::fabs(nx) + std::fabs(ny);
      double GnnM = t1 + (1.0 * RGv[0] * RGv[0] + G13[0] * G13[0] + G23[0] * G23[0]) * std::fabs(nz);
      double GnnP = t1 + (1.0 * RGv[1] * RGv[1] + G13[1] * G13[1] + G23[1] * G23[1]) * std::fabs(nz);
      double alpha = std::max(std::sqrt(GnnM * gamm * (Phyd[0] + dp[0]) * Gs[0] / gDens[0]) + std::fabs(Vel[0]),
                              std::sqrt(GnnP * gamm * (Phyd[1] + dp[1]) * Gs[1] / gDens[1]) + std::fabs(Vel[1]));
      double hf = m.Fscale[f] * 0.5;
      df[fp + DENS_VID * NfpTot] = hf * (gDens[1] * Vel[1] - gDens[0] * Vel[0] + (-alpha * (gDD[1] - gDD[0])));
      df[fp + RHOT_VID * NfpTot] = hf * (gRhot[1] * Vel[1] - gRhot[0] * Vel[0] + (-alpha * (gDR[1] - gDR[0])));
      double t3 = Gs[1] * dp[1], t4 = Gs[0] * dp[0];
      double mz = (t3 * RGv[1] - t4 * RGv[0]) * nz;
      double mx = (nx + G13[1] * nz) * t3 - (nx + G13[0] * nz) * t4;
      double my = (ny + G23[1] * nz) * t3 - (ny + G23[0] * nz) * t4;
      df[fp + MOMZ_VID * NfpTot] = hf * (gMZ[1] * Vel[1] - gMZ[0] * Vel[0] + mz + (-alpha * (gMZ[1] - gMZ[0])));
      df[fp + MOMX_VID * NfpTot] = hf * (gMX[1] * Vel[1] - gMX[0] * Vel[0] + mx + (-alpha * (gMX[1] - gMX[0])));
      df[fp + MOMY_VID * NfpTot] = hf * (gMY[1] * Vel[1] - gMY[0] * Vel[0] + my + (-alpha * (gMY[1] - gMY[0])));
    }
  }
}

// fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_rhot_heve.F90:292-489
void heve_cal_tend(const Element& e, const Mesh& m, const Consts& c, const DynState& s, double* dt5[5]) {
  const int Np = e.Np, NfpTot = e.NfpTot, Nfp = e.Nfp;
  vec del_flux;
  heve_numflux_generalvc(e, m, c, s, del_flux);
#pragma omp parallel
  {
    vec Flux(size_t(Np) * 3 * 5), DFlux(size_t(Np) * 4 * 5), drho(Np), RGsqrtV(Np), RGsqrt(Np), RDENS(Np);
#pragma omp for
    for (int ke = 0; ke < m.Ne; ++ke) {
      const int ke2d = m.emap2d[ke];
      const size_t o = size_t(ke) * Np;
      auto F = [&](int p, int d, int v) -> double& { return Flux[p + Np * (d + 3 * v)]; };
      auto DF = [&](int p, int d, int v) -> double& { return DFlux[p + Np * (d + 4 * v)]; };
      for (int p = 0; p < Np; ++p) {
        double GsqrtV = m.Gsqrt[o + p] / m.GsqrtH[(p % Nfp) + size_t(ke2d) * Nfp];
        RGsqrtV[p] = 1.0 / GsqrtV;
        RGsqrt[p] = 1.0 / m.Gsqrt[o + p];
        RDENS[p] = 1.0 / (s.DDENS[o + p] + s.DENS_hyd[o + p]);
      }
      for (int p = 0; p < Np; ++p) {
        double G = m.Gsqrt[o + p];
        F(p, 0, DENS_VID) = G * s.MOMX[o + p];
        F(p, 1, DENS_VID) = G * s.MOMY[o + p];
        F(p, 2, DENS_VID) = G * (s.MOMZ[o + p] * RGsqrtV[p] + m.G13[o + p] * s.MOMX[o + p] + m.G23[o + p] * s.MOMY[o + p]);
      }
      for (int p = 0; p < Np; ++p) {
        double pt = (s.THERM_hyd[o + p] + s.DRHOT[o + p]) * RDENS[p];
        for (int d = 0; d < 3; ++d) F(p, d, RHOT_VID) = F(p, d, DENS_VID) * pt;
        double w = s.MOMZ[o + p] * RDENS[p];
        F(p, 0, MOMZ_VID) = F(p, 0, DENS_VID) * w;
        F(p, 1, MOMZ_VID) = F(p, 1, DENS_VID) * w;
        F(p, 2, MOMZ_VID) = F(p, 2, DENS_VID) * w + m.Gsqrt[o + p] * s.DPRES[o + p] * RGsqrtV[p];
      }
      for (int p = 0; p < Np; ++p) {
        double GP = m.Gsqrt[o + p] * s.DPRES[o + p];
        double u = s.MOMX[o + p] * RDENS[p];
        F(p, 0, MOMX_VID) = F(p, 0, DENS_VID) * u + GP;
        F(p, 1, MOMX_VID) = F(p, 1, DENS_VID) * u;
        F(p, 2, MOMX_VID) = F(p, 2, DENS_VID) * u + GP * m.G13[o + p];
        double v = s.MOMY[o + p] * RDENS[p];
        F(p, 0, MOMY_VID) = F(p, 0, DENS_VID) * v;
        F(p, 1, MOMY_VID) = F(p, 1, DENS_VID) * v + GP;
        F(p, 2, MOMY_VID) = F(p, 2, DENS_VID) * v + GP * m.G23[o + p];
      }
      for (int v = 0; v < 5; ++v)
        op_div(e, &Flux[size_t(Np) * 3 * v], &del_flux[(size_t(ke) * PRGVAR_NUM + v) * NfpTot], &DFlux[size_t(Np) * 4 * v]);
      op_matz(e, e.VPOrdM1.data(), &s.DDENS[o], drho.data());
      auto div = [&](int p, int v) {
        return (m.E11[o + p] * DF(p, 0, v) + m.E22[o + p] * DF(p, 1, v) + m.E33[o + p] * DF(p, 2, v) + DF(p, 3, v)) * RGsqrt[p];
      };
      for (int p = 0; p < Np; ++p) {
        dt5[DENS_VID][o + p] = -div(p, DENS_VID);
        dt5[RHOT_VID][o + p] = -div(p, RHOT_VID);
        dt5[MOMZ_VID][o + p] = -div(p, MOMZ_VID) - c.GRAV * drho[p];
        double cor = s.CORIOLIS[(p % Nfp) + size_t(ke2d) * Nfp];
        double mx = -s.DPhydDx[o + p] + cor * s.MOMY[o + p];
        double my = -s.DPhydDy[o + p] - cor * s.MOMX[o + p];
        dt5[MOMX_VID][o + p] = mx - div(p, MOMX_VID);
        dt5[MOMY_VID][o + p] = my - div(p, MOMY_VID);
      }
    }
  }
}

// fluid_dyn_solver/scale_atm_dyn_dgm_modalfilter.F90:49-130 (do_weight_Gsqrt = .true.)
void modalfilter_apply(const Element& e, const Mesh& m, DynState& s) {
  const int Np = e.Np;
#pragma omp parallel
  {
    vec tmp(Np), work(Np), out(Np);
#pragma omp for
    for (int ke = 0; ke < m.Ne; ++ke) {
      const size_t o = size_t(ke) * Np;
      for (double* q : {s.DDENS.data(), s.MOMX.data(), s.MOMY.data(), s.MOMZ.data(), s.DRHOT.data()}) {
        for (int p = 0; p < Np; ++p) tmp[p] = m.Gsqrt[o + p] * q[o + p];
        op_modal_filter(e, tmp.data(), work.data(), out.data());
        for (int p = 0; p < Np; ++p) q[o + p] = out[p] * (1.0 / m.Gsqrt[o + p]);
      }
    }
  }
}

}  // namespace feo
