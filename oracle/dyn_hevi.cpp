// TEST INFRASTRUCTURE ONLY (see fe_oracle.hpp).  Regional HEVI: horizontally explicit tendency,
// and the vertically implicit column solve (one Newton iteration, block-tridiagonal LU).
//
// Restates, under FElib/src/fluid_dyn_solver:
//   scale_atm_dyn_dgm_nonhydro3d_rhot_hevi_numflux.F90:232-416   (numflux_get_generalvc)
//   scale_atm_dyn_dgm_nonhydro3d_rhot_hevi.F90:289-482           (cal_tend)
//   scale_atm_dyn_dgm_nonhydro3d_rhot_hevi.F90:772-965           (cal_vi)
//   scale_atm_dyn_dgm_nonhydro3d_rhot_hevi_common_2.F90:111-326  (eval_Ax), :329-479 (solve),
//       :482-586 (eval_Ax_uv), :589-689 (solve_uv), :693-901 (construct_matbnd),
//       :904-1011 (construct_matbnd_uv), :1016-1168 (vi_cal_del_flux_dyn_uv), :1171-1328 (vi_cal_del_flux_dyn)
//   scale_atm_dyn_dgm_hevi_common_linalgebra.F90:2142-2445       (solve_Nnode8_uv / solve_Nnode8_var3; the
//       other Nnode variants are the same algorithm for another block size)
//   mesh/scale_localmesh_3d.F90:188-242                          (GetVmapZ3D: column-local maps; at the bottom
//       and top of the column vmapP == vmapM)
// The reference vectorises over `im` horizontal nodes (first array index); the arithmetic per column
// is what is restated here, one column (ke2D, ij) at a time.
#include <cstdlib>
#include "fe_oracle.hpp"

#include <algorithm>
#include <stdexcept>

namespace feo {

// ---------------------------------------------------------------------------------------------
void hevi_numflux_generalvc(const Element& e, const Mesh& m, const Consts& c, const DynState& s, vec& del_flux) {
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
      double Gs[2], RGv[2], G13[2], G23[2], gDD[2], gMX[2], gMY[2], gMZ[2], gDR[2], Phyd[2], dp[2], gDens[2], gRhot[2], Velh[2], Vel[2];
      for (int t = 0; t < 2; ++t) {
        int i = id[t];
        Gs[t] = m.Gsqrt[i]; RGv[t] = 1.0 / Gs[t]; G13[t] = m.G13[i]; G23[t] = m.G23[i];   // GsqrtV_ = Gsqrt_ (:331)
        gDD[t] = Gs[t] * s.DDENS[i]; gMX[t] = Gs[t] * s.MOMX[i]; gMY[t] = Gs[t] * s.MOMY[i];
        gMZ[t] = Gs[t] * s.MOMZ[i]; gDR[t] = Gs[t] * s.DRHOT[i];
        Phyd[t] = s.PRES_hyd[i]; dp[t] = s.DPRES[i];
        gDens[t] = gDD[t] + Gs[t] * s.DENS_hyd[i];
        gRhot[t] = Gs[t] * s.THERM_hyd[i] + gDR[t];
        Velh[t] = (gMX[t] * nx + gMY[t] * ny) / gDens[t];
        Vel[t] = Velh[t] + ((gMZ[t] * RGv[t] + G13[t] * gMX[t] + G23[t] * gMY[t]) * nz) / gDens[t];
      }
      double swV = 1.0 - nz * nz;
      double alpha = swV * std::max(std::sqrt(gamm * (Phyd[0] + dp[0]) * Gs[0] / gDens[0]) + std::fabs(Vel[0]),
                                    std::sqrt(gamm * (Phyd[1] + dp[1]) * Gs[1] / gDens[1]) + std::fabs(Vel[1]));
      double hf = m.Fscale[f] * 0.5;
      df[fp + DENS_VID * NfpTot] = hf * (gDens[1] * Velh[1] - gDens[0] * Velh[0] + (-alpha * (gDD[1] - gDD[0])));
      df[fp + RHOT_VID * NfpTot] = hf * (gRhot[1] * Velh[1] - gRhot[0] * Velh[0] + (-alpha * (gDR[1] - gDR[0])));
      df[fp + MOMZ_VID * NfpTot] = hf * (gMZ[1] * Vel[1] - gMZ[0] * Vel[0] + (-alpha * (gMZ[1] - gMZ[0])));
      double t3 = Gs[1] * dp[1], t4 = Gs[0] * dp[0];
      double mx = (nx + G13[1] * nz) * t3 - (nx + G13[0] * nz) * t4;
      double my = (ny + G23[1] * nz) * t3 - (ny + G23[0] * nz) * t4;
      df[fp + MOMX_VID * NfpTot] = hf * (gMX[1] * Vel[1] - gMX[0] * Vel[0] + mx + (-alpha * (gMX[1] - gMX[0])));
      df[fp + MOMY_VID * NfpTot] = hf * (gMY[1] * Vel[1] - gMY[0] * Vel[0] + my + (-alpha * (gMY[1] - gMY[0])));
    }
  }
}

// ---------------------------------------------------------------------------------------------
void hevi_cal_tend(const Element& e, const Mesh& m, const Consts& c, const DynState& s, double* dt5[5]) {
  const int Np = e.Np, NfpTot = e.NfpTot, Nfp = e.Nfp;
  vec del_flux;
  hevi_numflux_generalvc(e, m, c, s, del_flux);
#pragma omp parallel
  {
    vec Flux(size_t(Np) * 3 * 5, 0.0), DFlux(size_t(Np) * 4 * 5), RGsqrtV(Np), RGsqrt(Np), RDENS(Np);
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
        F(p, 0, RHOT_VID) = F(p, 0, DENS_VID) * pt;
        F(p, 1, RHOT_VID) = F(p, 1, DENS_VID) * pt;
        // Flux(p,3,RHOT) is not set in the reference (:398) and its derivative is not used; 0 here.
        F(p, 2, RHOT_VID) = 0.0;
        double w = s.MOMZ[o + p] * RDENS[p];
        F(p, 0, MOMZ_VID) = F(p, 0, DENS_VID) * w;
        F(p, 1, MOMZ_VID) = F(p, 1, DENS_VID) * w;
        F(p, 2, MOMZ_VID) = F(p, 2, DENS_VID) * w;
      }
      for (int p = 0; p < Np; ++p) {
        double GP = m.Gsqrt[o + p] * s.DPRES[o + p];
        double u = s.MOMX[o + p] * RDENS[p], v = s.MOMY[o + p] * RDENS[p];
        F(p, 0, MOMX_VID) = F(p, 0, DENS_VID) * u + GP;
        F(p, 1, MOMX_VID) = F(p, 1, DENS_VID) * u;
        F(p, 2, MOMX_VID) = F(p, 2, DENS_VID) * u + GP * m.G13[o + p];
        F(p, 0, MOMY_VID) = F(p, 0, DENS_VID) * v;
        F(p, 1, MOMY_VID) = F(p, 1, DENS_VID) * v + GP;
        F(p, 2, MOMY_VID) = F(p, 2, DENS_VID) * v + GP * m.G23[o + p];
      }
      for (int v = 0; v < 5; ++v)
        op_div(e, &Flux[size_t(Np) * 3 * v], &del_flux[(size_t(ke) * PRGVAR_NUM + v) * NfpTot], &DFlux[size_t(Np) * 4 * v]);
      for (int p = 0; p < Np; ++p) {
        double E11 = m.E11[o + p], E22 = m.E22[o + p], E33 = m.E33[o + p];
        dt5[DENS_VID][o + p] = -(E11 * DF(p, 0, DENS_VID) + E22 * DF(p, 1, DENS_VID) + DF(p, 3, DENS_VID)) * RGsqrt[p];
        dt5[RHOT_VID][o + p] = -(E11 * DF(p, 0, RHOT_VID) + E22 * DF(p, 1, RHOT_VID) + DF(p, 3, RHOT_VID)) * RGsqrt[p];
        dt5[MOMZ_VID][o + p] =
            -(E11 * DF(p, 0, MOMZ_VID) + E22 * DF(p, 1, MOMZ_VID) + E33 * DF(p, 2, MOMZ_VID) + DF(p, 3, MOMZ_VID)) * RGsqrt[p];
        double cor = s.CORIOLIS[(p % Nfp) + size_t(ke2d) * Nfp];
        double mx = -s.DPhydDx[o + p] + cor * s.MOMY[o + p];
        double my = -s.DPhydDy[o + p] - cor * s.MOMX[o + p];
        dt5[MOMX_VID][o + p] =
            mx - (E11 * DF(p, 0, MOMX_VID) + E22 * DF(p, 1, MOMX_VID) + E33 * DF(p, 2, MOMX_VID) + DF(p, 3, MOMX_VID)) * RGsqrt[p];
        dt5[MOMY_VID][o + p] =
            my - (E11 * DF(p, 0, MOMY_VID) + E22 * DF(p, 1, MOMY_VID) + E33 * DF(p, 2, MOMY_VID) + DF(p, 3, MOMY_VID)) * RGsqrt[p];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
namespace {

// In-place LU with partial pivoting exactly as solve_Nnode*_var3 does it (linalgebra.F90:2320-2355):
// pivot = first row with the strictly largest |A(i,k)|, whole rows swapped, reciprocal pivot stored
// on the diagonal, unit-lower multipliers below it.  A is n x n, row-major.
void lu_factor(double* A, int n, int* ipiv) {
  for (int k = 0; k < n; ++k) {
    double best = std::fabs(A[k * n + k]);
    ipiv[k] = k;
    for (int i = k + 1; i < n; ++i) {
      double t = std::fabs(A[i * n + k]);
      if (t > best) { ipiv[k] = i; best = t; }
    }
    if (ipiv[k] != k)
      for (int j = 0; j < n; ++j) std::swap(A[k * n + j], A[ipiv[k] * n + j]);
    double inv = 1.0 / A[k * n + k];
    A[k * n + k] = inv;
    for (int i = k + 1; i < n; ++i) A[i * n + k] = A[i * n + k] * inv;
    for (int j = k + 1; j < n; ++j)
      for (int i = k + 1; i < n; ++i) A[i * n + j] = A[i * n + j] - A[i * n + k] * A[k * n + j];
  }
}
// rhs: n x nrhs row-major (linalgebra.F90:2357-2440)
void lu_solve(const double* A, int n, const int* ipiv, double* rhs, int nrhs) {
  for (int i = 0; i < n - 1; ++i)
    if (ipiv[i] != i)
      for (int r = 0; r < nrhs; ++r) std::swap(rhs[i * nrhs + r], rhs[ipiv[i] * nrhs + r]);
  for (int i = 0; i < n; ++i)
    for (int r = 0; r < nrhs; ++r) {
      double t = rhs[i * nrhs + r];
      for (int j = 0; j < i; ++j) t = t - rhs[j * nrhs + r] * A[i * n + j];
      rhs[i * nrhs + r] = t;
    }
  for (int i = n - 1; i >= 0; --i)
    for (int r = 0; r < nrhs; ++r) {
      double t = rhs[i * nrhs + r];
      for (int j = i + 1; j < n; ++j) t = t - A[i * n + j] * rhs[j * nrhs + r];
      rhs[i * nrhs + r] = t * A[i * n + i];
    }
}

// Experiment switch (tests only, FEO_VI_STATIC_PIVOT=1): eliminate in the fixed order DDENS_0..n, MOMZ_0..n, DRHOT_0..n
// without any pivot search, to measure how far a search-free device solver may deviate from the reference's partial pivoting.
bool vi_static_pivot() {
  static int v = -1;
  if (v < 0) { const char* e = std::getenv("FEO_VI_STATIC_PIVOT"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
// symmetric permutation to block order, LU without pivoting, solve, permute back.  A: n x n row-major (destroyed); rhs n x nrhs.
void solve_static_order(double* A, int n, double* rhs, int nrhs) {
  const int np = n / 3;
  std::vector<int> ord(n);
  for (int v = 0; v < 3; ++v) for (int pv = 0; pv < np; ++pv) ord[v * np + pv] = 3 * pv + v;
  std::vector<double> B(size_t(n) * n), r(size_t(n) * nrhs);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) B[size_t(i) * n + j] = A[size_t(ord[i]) * n + ord[j]];
    for (int q = 0; q < nrhs; ++q) r[size_t(i) * nrhs + q] = rhs[size_t(ord[i]) * nrhs + q];
  }
  for (int k = 0; k < n; ++k) {                      // Gauss-Jordan, pivot (k, k)
    const double inv = 1.0 / B[size_t(k) * n + k];
    for (int i = 0; i < n; ++i) {
      if (i == k) continue;
      const double m = B[size_t(i) * n + k] * inv;
      for (int j = k + 1; j < n; ++j) B[size_t(i) * n + j] -= m * B[size_t(k) * n + j];
      for (int q = 0; q < nrhs; ++q) r[size_t(i) * nrhs + q] -= m * r[size_t(k) * nrhs + q];
    }
    for (int q = 0; q < nrhs; ++q) r[size_t(k) * nrhs + q] *= inv;
    for (int j = k + 1; j < n; ++j) B[size_t(k) * n + j] *= inv;
  }
  for (int i = 0; i < n; ++i) for (int q = 0; q < nrhs; ++q) rhs[size_t(ord[i]) * nrhs + q] = r[size_t(i) * nrhs + q];
}

// Experiment switch (tests only, FEO_VI_BLOCK=1): the elimination order a thread-per-column device solver would use -- DDENS
// eliminated with static pivots, then the DRHOT block (8 x 8, partial pivoting inside), then the Schur complement in MOMZ
// (8 x 8, partial pivoting) -- to measure its deviation from the reference's partial pivoting over the whole block.
bool vi_block_order() {
  static int v = -1;
  if (v < 0) { const char* e = std::getenv("FEO_VI_BLOCK"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
void solve_block_order(double* A, int n, double* rhs, int nrhs) {
  const int np = n / 3;
  std::vector<int> ord(n);
  for (int v = 0; v < 3; ++v) for (int pv = 0; pv < np; ++pv) ord[v * np + pv] = 3 * pv + v;      // [rho | w | theta]
  std::vector<double> B(size_t(n) * n), r(size_t(n) * nrhs);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) B[size_t(i) * n + j] = A[size_t(ord[i]) * n + ord[j]];
    for (int q = 0; q < nrhs; ++q) r[size_t(i) * nrhs + q] = rhs[size_t(ord[i]) * nrhs + q];
  }
  auto elim = [&](int k, int p) {     // Gauss-Jordan step: pivot row p for column k (all rows of the whole system are updated)
    const double inv = 1.0 / B[size_t(p) * n + k];
    for (int i = 0; i < n; ++i) {
      if (i == p) continue;
      const double m = B[size_t(i) * n + k] * inv;
      if (m == 0.0) continue;
      for (int j = 0; j < n; ++j) if (j != k) B[size_t(i) * n + j] -= m * B[size_t(p) * n + j];
      B[size_t(i) * n + k] = 0.0;
      for (int q = 0; q < nrhs; ++q) r[size_t(i) * nrhs + q] -= m * r[size_t(p) * nrhs + q];
    }
    for (int q = 0; q < nrhs; ++q) r[size_t(p) * nrhs + q] *= inv;
    for (int j = 0; j < n; ++j) if (j != k) B[size_t(p) * n + j] *= inv;
    B[size_t(p) * n + k] = 1.0;
  };
  std::vector<int> rowof(n, -1);
  std::vector<char> used(n, 0);
  for (int k = 0; k < np; ++k) { elim(k, k); used[k] = 1; rowof[k] = k; }            // rho: static pivots
  for (int blk : {2, 1})                                                            // theta block, then w: partial pivoting inside the block rows
    for (int k = blk * np; k < (blk + 1) * np; ++k) {
      int p = -1; double best = -1.0;
      for (int i = blk * np; i < (blk + 1) * np; ++i) if (!used[i] && std::fabs(B[size_t(i) * n + k]) > best) { best = std::fabs(B[size_t(i) * n + k]); p = i; }
      elim(k, p); used[p] = 1; rowof[k] = p;
    }
  for (int k = 0; k < n; ++k) for (int q = 0; q < nrhs; ++q) rhs[size_t(ord[k]) * nrhs + q] = r[size_t(rowof[k]) * nrhs + q];
}

struct VIWork {
  // PROG_VARS (the Newton iterate), indexed like the fields: [VID][Np*Ne]
  vec pv[5];
  vec alph;     // (2*Nfp, Ne): vertical faces only (bottom, top); horizontal faces have nz = 0 -> alph = 0
  vec GsqrtV;   // (Np, Ne)
  vec DENS, W, WT, POT, DPDRHOT;  // (Np, Ne)
  vec t_dens, t_momz, t_rhot, t_momx, t_momy;
};

inline double rhot_hyd_dry(const Consts& c, double pres_hyd) {
  return c.PRES00 / c.Rdry * std::pow(pres_hyd / c.PRES00, c.CVdry / c.CPdry);
}

// vi_cal_del_flux_dyn_uv (hevi_common_2.F90:1016-1168) + eval_Ax_uv (:482-586); G_ij = identity, gam = 1 on the cube.
// Computes alph on the vertical faces (from var0) and the (MOMX, MOMY) residual tendencies.
void eval_ax_uv(const Element& e, const Mesh& m, const Consts& c, const DynState& s, const double* var0[5], VIWork& w) {
  const int Np = e.Np, Nfp = e.Nfp, np = e.np, Ne2D = m.Ne2D;
  const double gamm = c.CPdry / c.CVdry, rP0 = 1.0 / c.PRES00;
#pragma omp parallel for
  for (int ke = 0; ke < m.Ne; ++ke) {
    const int kz = ke / Ne2D;
    for (int side = 0; side < 2; ++side) {
      const int f = 4 + side;
      for (int ij = 0; ij < Nfp; ++ij) {
        const int iM = ke * Np + ij + (side ? (np - 1) * Nfp : 0);
        int iP = iM;  // GetVmapZ3D
        if (side == 0 && kz > 0) iP = (ke - Ne2D) * Np + ij + (np - 1) * Nfp;
        if (side == 1 && kz < m.NeZ - 1) iP = (ke + Ne2D) * Np + ij;
        const double nz = m.nz[size_t(ke) * e.NfpTot + f * Nfp + ij];
        double a2[2];
        const int id[2] = {iM, iP};
        for (int t = 0; t < 2; ++t) {
          const int i = id[t];
          double RGv = 1.0 / w.GsqrtV[i], G13 = m.G13[i], G23 = m.G23[i];
          double rdens0 = 1.0 / (s.DENS_hyd[i] + var0[DENS_VID][i]);
          double wt0 = (var0[MOMZ_VID][i] * RGv + G13 * var0[MOMX_VID][i] + G23 * var0[MOMY_VID][i]) * rdens0;
          double pres0 = c.PRES00 * std::pow(s.Rtot[i] * rP0 * (rhot_hyd_dry(c, s.PRES_hyd[i]) + var0[RHOT_VID][i]),
                                             s.CPtot[i] / s.CVtot[i]);
          double Gxz = 1.0 * (1.0 * G13 + 0.0 * G23), Gyz = 1.0 * (0.0 * G13 + 1.0 * G23);
          double Gnn = (1.0 * RGv * RGv + G13 * Gxz + G23 * Gyz) * std::fabs(nz);
          a2[t] = std::fabs(wt0) + std::sqrt(Gnn * gamm * pres0 * rdens0);
        }
        w.alph[size_t(ke) * 2 * Nfp + side * Nfp + ij] = nz * nz * std::max(a2[0], a2[1]);
      }
    }
  }
#pragma omp parallel
  {
    vec df(size_t(2) * e.NfpTot, 0.0), L(size_t(2) * Np);
#pragma omp for
    for (int ke = 0; ke < m.Ne; ++ke) {
      const int kz = ke / Ne2D;
      for (int side = 0; side < 2; ++side)
        for (int ij = 0; ij < Nfp; ++ij) {
          const int f = 4 + side;
          const int iM = ke * Np + ij + (side ? (np - 1) * Nfp : 0);
          int iP = iM;
          if (side == 0 && kz > 0) iP = (ke - Ne2D) * Np + ij + (np - 1) * Nfp;
          if (side == 1 && kz < m.NeZ - 1) iP = (ke + Ne2D) * Np + ij;
          double t1 = -0.5 * m.Fscale[size_t(ke) * e.NfpTot + f * Nfp + ij] * w.alph[size_t(ke) * 2 * Nfp + side * Nfp + ij];
          df[f * Nfp + ij] = t1 * (w.pv[MOMX_VID][iP] - w.pv[MOMX_VID][iM]);
          df[e.NfpTot + f * Nfp + ij] = t1 * (w.pv[MOMY_VID][iP] - w.pv[MOMY_VID][iM]);
        }
      op_lift(e, &df[0], &L[0]);
      op_lift(e, &df[e.NfpTot], &L[Np]);
      for (int p = 0; p < Np; ++p) {
        double RGv = 1.0 / w.GsqrtV[size_t(ke) * Np + p];
        w.t_momx[size_t(ke) * Np + p] = -L[p] * RGv;
        w.t_momy[size_t(ke) * Np + p] = -L[Np + p] * RGv;
      }
    }
  }
}

// construct_matbnd_uv + solve_uv
void solve_uv(const Element& e, const Mesh& m, const DynState& s, double impl_fac, VIWork& w) {
  const int Np = e.Np, Nfp = e.Nfp, np = e.np, Ne2D = m.Ne2D, NeZ = m.NeZ;
  const double* cur[2] = {s.MOMX.data(), s.MOMY.data()};
  const int vid[2] = {MOMX_VID, MOMY_VID};
  const vec* tt[2] = {&w.t_momx, &w.t_momy};
#pragma omp parallel
  {
    vec D(size_t(np) * np), Lm(np), G(size_t(NeZ) * np), b(size_t(NeZ) * np * 2), rhs(size_t(np) * 3);
    std::vector<int> ipiv(np);
#pragma omp for collapse(2)
    for (int k2 = 0; k2 < Ne2D; ++k2)
      for (int ij = 0; ij < Nfp; ++ij) {
        for (int kz = 0; kz < NeZ; ++kz) {
          const int ke = k2 + kz * Ne2D;
          for (int pv = 0; pv < np; ++pv) {
            const size_t n = size_t(ke) * Np + ij + pv * Nfp;
            for (int r = 0; r < 2; ++r)
              b[(size_t(kz) * np + pv) * 2 + r] = impl_fac * (*tt[r])[n] - w.pv[vid[r]][n] + cur[r][n];
          }
          for (int a = 0; a < np; ++a) for (int bq = 0; bq < np; ++bq) D[a * np + bq] = (a == bq) ? 1.0 : 0.0;
          for (int pv = 0; pv < np; ++pv) G[size_t(kz) * np + pv] = 0.0;
          for (int f1 = 0; f1 < 2; ++f1) {
            const bool bc = (kz == 0 && f1 == 0) || (kz == NeZ - 1 && f1 == 1);
            if (bc) continue;
            const int kz2 = f1 == 0 ? kz - 1 : kz + 1, pv1 = f1 == 0 ? 0 : np - 1, f2 = 1 - f1;
            const int ke2 = k2 + kz2 * Ne2D;
            for (int pv = 0; pv < np; ++pv) {
              const size_t n = size_t(ke) * Np + ij + pv * Nfp;
              double fac = 0.5 * impl_fac / w.GsqrtV[n];
              double t1 = fac * e.lift1d[pv * 2 + f1] * m.Fscale[size_t(ke) * e.NfpTot + (4 + f1) * Nfp + ij] *
                          std::max(w.alph[size_t(ke) * 2 * Nfp + f1 * Nfp + ij], w.alph[size_t(ke2) * 2 * Nfp + f2 * Nfp + ij]);
              D[pv * np + pv1] += t1;
              if (f1 == 0) Lm[pv] = -t1; else G[size_t(kz) * np + pv] = -t1;   // BndMatU is stored in G (:632)
            }
          }
          if (kz > 0)
            for (int pv = 0; pv < np; ++pv) {
              double t = Lm[pv];
              D[pv * np + 0] = D[pv * np + 0] - t * G[size_t(kz - 1) * np + (np - 1)];
              for (int r = 0; r < 2; ++r)
                b[(size_t(kz) * np + pv) * 2 + r] = b[(size_t(kz) * np + pv) * 2 + r] - t * b[(size_t(kz - 1) * np + (np - 1)) * 2 + r];
            }
          lu_factor(D.data(), np, ipiv.data());
          const int nr = (kz == NeZ - 1) ? 2 : 3;
          for (int pv = 0; pv < np; ++pv) {
            rhs[pv * nr + 0] = b[(size_t(kz) * np + pv) * 2];
            rhs[pv * nr + 1] = b[(size_t(kz) * np + pv) * 2 + 1];
            if (nr == 3) rhs[pv * nr + 2] = G[size_t(kz) * np + pv];
          }
          lu_solve(D.data(), np, ipiv.data(), rhs.data(), nr);
          for (int pv = 0; pv < np; ++pv) {
            b[(size_t(kz) * np + pv) * 2] = rhs[pv * nr + 0];
            b[(size_t(kz) * np + pv) * 2 + 1] = rhs[pv * nr + 1];
            if (nr == 3) G[size_t(kz) * np + pv] = rhs[pv * nr + 2];
          }
        }
        for (int kz = NeZ - 2; kz >= 0; --kz)
          for (int pv = 0; pv < np; ++pv) {
            double t = G[size_t(kz) * np + pv];
            for (int r = 0; r < 2; ++r)
              b[(size_t(kz) * np + pv) * 2 + r] = b[(size_t(kz) * np + pv) * 2 + r] - t * b[(size_t(kz + 1) * np + 0) * 2 + r];
          }
        for (int kz = 0; kz < NeZ; ++kz)
          for (int pv = 0; pv < np; ++pv) {
            const size_t n = size_t(k2 + kz * Ne2D) * Np + ij + pv * Nfp;
            w.pv[MOMX_VID][n] = w.pv[MOMX_VID][n] + b[(size_t(kz) * np + pv) * 2];
            w.pv[MOMY_VID][n] = w.pv[MOMY_VID][n] + b[(size_t(kz) * np + pv) * 2 + 1];
          }
      }
  }
}

// vi_cal_del_flux_dyn (:1171-1328) + eval_Ax (:111-326)
void eval_ax(const Element& e, const Mesh& m, const Consts& c, const DynState& s, VIWork& w) {
  const int Np = e.Np, Nfp = e.Nfp, np = e.np, Ne2D = m.Ne2D, NfpTot = e.NfpTot;
  const double rP0 = 1.0 / c.PRES00;
#pragma omp parallel
  {
    vec df(size_t(3) * NfpTot, 0.0), Flux(size_t(3) * Np), DF(size_t(6) * Np), RHOT(Np), drho(Np);
#pragma omp for
    for (int ke = 0; ke < m.Ne; ++ke) {
      const int kz = ke / Ne2D;
      const size_t o = size_t(ke) * Np;
      for (int side = 0; side < 2; ++side)
        for (int ij = 0; ij < Nfp; ++ij) {
          const int f = 4 + side;
          const int iM = ke * Np + ij + (side ? (np - 1) * Nfp : 0);
          int iP = iM;
          if (side == 0 && kz > 0) iP = (ke - Ne2D) * Np + ij + (np - 1) * Nfp;
          if (side == 1 && kz < m.NeZ - 1) iP = (ke + Ne2D) * Np + ij;
          const size_t ff = size_t(ke) * NfpTot + f * Nfp + ij;
          const double nz = m.nz[ff];
          const int id[2] = {iM, iP};
          double DD[2], MZ[2], MW[2], DR[2], pott[2], dpres[2];
          for (int t = 0; t < 2; ++t) {
            const int i = id[t];
            DD[t] = w.pv[DENS_VID][i]; MZ[t] = w.pv[MOMZ_VID][i]; DR[t] = w.pv[RHOT_VID][i];
            double dens = s.DENS_hyd[i] + DD[t];
            pott[t] = (rhot_hyd_dry(c, s.PRES_hyd[i]) + DR[t]) / dens;
            dpres[t] = c.PRES00 * std::pow(s.Rtot[i] * rP0 * dens * pott[t], s.CPtot[i] / s.CVtot[i]);
            dpres[t] = dpres[t] - s.PRES_hyd[i];
            MW[t] = MZ[t] + w.GsqrtV[i] * m.G13[i] * w.pv[MOMX_VID][i] + w.GsqrtV[i] * m.G23[i] * w.pv[MOMY_VID][i];
          }
          if ((kz == 0 || kz == m.NeZ - 1) && iM == iP) {
            MZ[1] = -MZ[0] - 2.0 * w.GsqrtV[iM] * (m.G13[iM] * w.pv[MOMX_VID][iM] + m.G23[iM] * w.pv[MOMY_VID][iM]);
            MW[1] = -MW[0];
          }
          const double t1 = 0.5 * m.Fscale[ff], al = w.alph[size_t(ke) * 2 * Nfp + side * Nfp + ij];
          df[0 * NfpTot + f * Nfp + ij] = t1 * ((MW[1] - MW[0]) * nz - al * (DD[1] - DD[0]));
          df[1 * NfpTot + f * Nfp + ij] = t1 * ((dpres[1] - dpres[0]) * nz - al * (MZ[1] - MZ[0]));
          df[2 * NfpTot + f * Nfp + ij] = t1 * ((pott[1] * MW[1] - pott[0] * MW[0]) * nz - al * (DR[1] - DR[0]));
        }
      for (int p = 0; p < Np; ++p) {
        const size_t n = o + p;
        Flux[p] = w.pv[MOMZ_VID][n] + w.GsqrtV[n] * m.G13[n] * w.pv[MOMX_VID][n] + w.GsqrtV[n] * m.G23[n] * w.pv[MOMY_VID][n];
        RHOT[p] = rhot_hyd_dry(c, s.PRES_hyd[n]) + w.pv[RHOT_VID][n];
        double pt = RHOT[p] / (w.pv[DENS_VID][n] + s.DENS_hyd[n]);
        Flux[Np + p] = pt * Flux[p];                                                             // RHOT
        Flux[2 * Np + p] = c.PRES00 * std::pow(s.Rtot[n] * rP0 * RHOT[p], s.CPtot[n] / s.CVtot[n]) - s.PRES_hyd[n];  // MOMZ
      }
      // DFlux(:,1,v) = Dz Flux_v ; DFlux(:,2,v) = Lift del_flux_v   (v: DENS, RHOT, MOMZ)
      op_dz(e, &Flux[0], &DF[0]);           op_lift(e, &df[0 * NfpTot], &DF[Np]);
      op_dz(e, &Flux[Np], &DF[2 * Np]);     op_lift(e, &df[2 * NfpTot], &DF[3 * Np]);
      op_dz(e, &Flux[2 * Np], &DF[4 * Np]); op_lift(e, &df[1 * NfpTot], &DF[5 * Np]);
      op_matz(e, e.VPOrdM1.data(), &w.pv[DENS_VID][o], drho.data());
      for (int p = 0; p < Np; ++p) {
        const size_t n = o + p;
        const double E33 = m.E33[n], RGv = 1.0 / w.GsqrtV[n];
        w.t_dens[n] = -(E33 * DF[p] + DF[Np + p]) * RGv;
        w.t_rhot[n] = -(E33 * DF[2 * Np + p] + DF[3 * Np + p]) * RGv;
        w.t_momz[n] = -(E33 * DF[4 * Np + p] + DF[5 * Np + p]) * RGv - c.GRAV * drho[p];
        double dens = s.DENS_hyd[n] + w.pv[DENS_VID][n];
        w.DENS[n] = dens;
        w.POT[n] = RHOT[p] / dens;
        w.W[n] = w.pv[MOMZ_VID][n] / dens;
        w.WT[n] = Flux[p] / dens;
        double g = s.CPtot[n] / s.CVtot[n];
        w.DPDRHOT[n] = g * c.PRES00 * std::pow(s.Rtot[n] / c.PRES00 * RHOT[p], g) / RHOT[p];
      }
    }
  }
}

// construct_matbnd (:693-901) + solve (:329-479): block-tridiagonal system with blocks of 3*np, ordering
// row = var + 3*pv with var (DENS, MOMZ, RHOT) = (0, 1, 2).
void solve_var3(const Element& e, const Mesh& m, const Consts& c, const DynState& s, double impl_fac, VIWork& w) {
  const int Np = e.Np, Nfp = e.Nfp, np = e.np, Ne2D = m.Ne2D, NeZ = m.NeZ, nb = 3 * np;
  const double* cur[3] = {s.DDENS.data(), s.MOMZ.data(), s.DRHOT.data()};
  const int vid[3] = {DENS_VID, MOMZ_VID, RHOT_VID};
  const vec* tt[3] = {&w.t_dens, &w.t_momz, &w.t_rhot};
#pragma omp parallel
  {
    vec D(size_t(nb) * nb), Lm(size_t(nb) * 3), G(size_t(NeZ) * nb * 3), b(size_t(NeZ) * nb), rhs(size_t(nb) * 4), fac_dz(size_t(np) * np);
    std::vector<int> ipiv(nb);
#pragma omp for collapse(2)
    for (int k2 = 0; k2 < Ne2D; ++k2)
      for (int ij = 0; ij < Nfp; ++ij) {
        auto node = [&](int kz, int pv) { return size_t(k2 + kz * Ne2D) * Np + ij + pv * Nfp; };
        for (int kz = 0; kz < NeZ; ++kz) {
          const int ke = k2 + kz * Ne2D;
          for (int pv = 0; pv < np; ++pv) {
            const size_t n = node(kz, pv);
            for (int v = 0; v < 3; ++v) b[size_t(kz) * nb + 3 * pv + v] = impl_fac * (*tt[v])[n] - w.pv[vid[v]][n] + cur[v][n];
          }
          auto Dm = [&](int v, int pv, int v2, int pv2) -> double& { return D[size_t(3 * pv + v) * nb + 3 * pv2 + v2]; };
          for (int pv = 0; pv < np; ++pv)
            for (int pv2 = 0; pv2 < np; ++pv2) {
              const size_t n = node(kz, pv), n2 = node(kz, pv2);
              const double Dx3 = impl_fac * e.D1D[pv * np + pv2];
              const double fdz = m.E33[n] / w.GsqrtV[n] * Dx3;
              const double Id = (pv == pv2) ? 1.0 : 0.0;
              Dm(0, pv, 0, pv2) = Id;
              Dm(0, pv, 1, pv2) = fdz;
              Dm(0, pv, 2, pv2) = 0.0;
              Dm(1, pv, 1, pv2) = Id;
              Dm(1, pv, 0, pv2) = impl_fac * c.GRAV * e.VPOrdM1[pv * np + pv2];
              Dm(1, pv, 2, pv2) = fdz * w.DPDRHOT[n2];
              Dm(2, pv, 0, pv2) = -fdz * w.POT[n2] * w.WT[n2];
              Dm(2, pv, 1, pv2) = fdz * w.POT[n2];
              Dm(2, pv, 2, pv2) = Id + fdz * w.WT[n2];
            }
          for (size_t q = 0; q < size_t(nb) * 3; ++q) G[size_t(kz) * nb * 3 + q] = 0.0;
          for (int f1 = 0; f1 < 2; ++f1) {
            const bool bc = (kz == 0 && f1 == 0) || (kz == NeZ - 1 && f1 == 1);
            int kz2 = f1 == 0 ? std::max(kz - 1, 0) : std::min(kz + 1, NeZ - 1);
            int pv1 = f1 == 0 ? 0 : np - 1, pv2 = f1 == 0 ? np - 1 : 0, f2 = 1 - f1;
            if (bc) { pv2 = pv1; f2 = f1; }
            const int ke2 = k2 + kz2 * Ne2D;
            const size_t n1 = node(kz, pv1), n2 = node(kz2, pv2);
            for (int pv = 0; pv < np; ++pv) {
              const size_t n = node(kz, pv);
              const size_t ff = size_t(ke) * e.NfpTot + (4 + f1) * Nfp + ij;
              double fac = 0.5 * impl_fac / w.GsqrtV[n] * e.lift1d[pv * 2 + f1] * m.Fscale[ff];
              double t1 = fac * std::max(w.alph[size_t(ke) * 2 * Nfp + f1 * Nfp + ij], w.alph[size_t(ke2) * 2 * Nfp + f2 * Nfp + ij]);
              double t2 = fac * m.nz[ff];
              if (bc) {
                Dm(2, pv, 0, pv1) = Dm(2, pv, 0, pv1) + 2.0 * t2 * w.POT[n1] * w.WT[n1];
                Dm(0, pv, 1, pv1) = Dm(0, pv, 1, pv1) - 2.0 * t2;
                Dm(1, pv, 1, pv1) = Dm(1, pv, 1, pv1) + 2.0 * t1;
                Dm(2, pv, 1, pv1) = Dm(2, pv, 1, pv1) - 2.0 * t2 * w.POT[n1];
                Dm(2, pv, 2, pv1) = Dm(2, pv, 2, pv1) - 2.0 * t2 * w.WT[n1];
              } else {
                Dm(0, pv, 0, pv1) = Dm(0, pv, 0, pv1) + t1;
                Dm(2, pv, 0, pv1) = Dm(2, pv, 0, pv1) + t2 * w.POT[n1] * w.WT[n1];
                Dm(0, pv, 1, pv1) = Dm(0, pv, 1, pv1) - t2;
                Dm(1, pv, 1, pv1) = Dm(1, pv, 1, pv1) + t1;
                Dm(2, pv, 1, pv1) = Dm(2, pv, 1, pv1) - t2 * w.POT[n1];
                Dm(1, pv, 2, pv1) = Dm(1, pv, 2, pv1) - t2 * w.DPDRHOT[n1];
                Dm(2, pv, 2, pv1) = Dm(2, pv, 2, pv1) + t1 - t2 * w.WT[n1];
                double* X = (f1 == 0) ? &Lm[size_t(3 * pv) * 3] : &G[size_t(kz) * nb * 3 + size_t(3 * pv) * 3];
                // X[(row var) * 3 + (column var)]
                X[0 * 3 + 0] = -t1;  X[1 * 3 + 0] = 0.0;                 X[2 * 3 + 0] = -t2 * w.POT[n2] * w.WT[n2];
                X[0 * 3 + 1] = t2;   X[1 * 3 + 1] = -t1;                 X[2 * 3 + 1] = t2 * w.POT[n2];
                X[0 * 3 + 2] = 0.0;  X[1 * 3 + 2] = t2 * w.DPDRHOT[n2];  X[2 * 3 + 2] = -t1 + t2 * w.WT[n2];
              }
            }
          }
          if (kz > 0) {
            const double* Gp = &G[size_t(kz - 1) * nb * 3];
            const double* bp = &b[size_t(kz - 1) * nb];
            const int p1 = 3 * (np - 1);
            for (int r = 0; r < nb; ++r) {
              const double a0 = Lm[r * 3], a1 = Lm[r * 3 + 1], a2 = Lm[r * 3 + 2];
              for (int cc = 0; cc < 3; ++cc)
                D[size_t(r) * nb + cc] = D[size_t(r) * nb + cc] - a0 * Gp[(p1) * 3 + cc] - a1 * Gp[(p1 + 1) * 3 + cc] - a2 * Gp[(p1 + 2) * 3 + cc];
              b[size_t(kz) * nb + r] = b[size_t(kz) * nb + r] - a0 * bp[p1] - a1 * bp[p1 + 1] - a2 * bp[p1 + 2];
            }
          }
          const int nr = (kz == NeZ - 1) ? 1 : 4;
          for (int r = 0; r < nb; ++r) {
            rhs[r * nr] = b[size_t(kz) * nb + r];
            if (nr == 4) for (int cc = 0; cc < 3; ++cc) rhs[r * nr + 1 + cc] = G[size_t(kz) * nb * 3 + r * 3 + cc];
          }
          if (vi_static_pivot()) solve_static_order(D.data(), nb, rhs.data(), nr);
          else if (vi_block_order()) solve_block_order(D.data(), nb, rhs.data(), nr);
          else {
            lu_factor(D.data(), nb, ipiv.data());
            lu_solve(D.data(), nb, ipiv.data(), rhs.data(), nr);
          }
          for (int r = 0; r < nb; ++r) {
            b[size_t(kz) * nb + r] = rhs[r * nr];
            if (nr == 4) for (int cc = 0; cc < 3; ++cc) G[size_t(kz) * nb * 3 + r * 3 + cc] = rhs[r * nr + 1 + cc];
          }
        }
        for (int kz = NeZ - 2; kz >= 0; --kz)
          for (int r = 0; r < nb; ++r) {
            const double* g = &G[size_t(kz) * nb * 3 + r * 3];
            const double* bn = &b[size_t(kz + 1) * nb];
            b[size_t(kz) * nb + r] = b[size_t(kz) * nb + r] - g[0] * bn[0] - g[1] * bn[1] - g[2] * bn[2];
          }
        for (int kz = 0; kz < NeZ; ++kz)
          for (int pv = 0; pv < np; ++pv) {
            const size_t n = node(kz, pv);
            for (int v = 0; v < 3; ++v) w.pv[vid[v]][n] = w.pv[vid[v]][n] + b[size_t(kz) * nb + 3 * pv + v];
          }
      }
  }
}

}  // namespace

// rhot_hevi.F90:772-965
void hevi_cal_vi(const Element& e, const Mesh& m, const Consts& c, const DynState& s, const double* var0[5],
                 double impl_fac, double dt, double* dt5[5]) {
  (void)dt;
  const int Np = e.Np, Nfp = e.Nfp;
  const size_t nint = size_t(Np) * m.Ne;
  VIWork w;
  for (int v = 0; v < 5; ++v) w.pv[v].assign(var0[v], var0[v] + nint);
  w.alph.assign(size_t(2) * Nfp * m.Ne, 0.0);
  w.GsqrtV.resize(nint);
  for (vec* x : {&w.DENS, &w.W, &w.WT, &w.POT, &w.DPDRHOT, &w.t_dens, &w.t_momz, &w.t_rhot, &w.t_momx, &w.t_momy}) x->assign(nint, 0.0);
  for (int ke = 0; ke < m.Ne; ++ke)
    for (int p = 0; p < Np; ++p) {
      // regional: Gsqrt / GsqrtH (rhot_hevi.F90:864); global: Gsqrt / (gam^2 GsqrtH) (globalnonhydro3d_rhot_hevi.F90:965)
      const double g2 = m.is_global ? m.gam[size_t(ke) * Np + p] * m.gam[size_t(ke) * Np + p] : 1.0;
      w.GsqrtV[size_t(ke) * Np + p] = m.Gsqrt[size_t(ke) * Np + p] / (g2 * m.GsqrtH[(p % Nfp) + size_t(m.emap2d[ke]) * Nfp]);
    }
  const double* cur[5];
  cur[DENS_VID] = s.DDENS.data(); cur[RHOT_VID] = s.DRHOT.data(); cur[MOMZ_VID] = s.MOMZ.data();
  cur[MOMX_VID] = s.MOMX.data(); cur[MOMY_VID] = s.MOMY.data();

  eval_ax_uv(e, m, c, s, var0, w);
  if (std::fabs(impl_fac) > 0.0) {
    solve_uv(e, m, s, impl_fac, w);
    eval_ax(e, m, c, s, w);
    solve_var3(e, m, c, s, impl_fac, w);
    for (int v = 0; v < 5; ++v)
      for (size_t n = 0; n < nint; ++n) dt5[v][n] = (w.pv[v][n] - cur[v][n]) / impl_fac;
  } else {
    eval_ax(e, m, c, s, w);
    for (size_t n = 0; n < nint; ++n) {
      dt5[MOMX_VID][n] = w.t_momx[n]; dt5[MOMY_VID][n] = w.t_momy[n];
      dt5[DENS_VID][n] = w.t_dens[n]; dt5[MOMZ_VID][n] = w.t_momz[n]; dt5[RHOT_VID][n] = w.t_rhot[n];
    }
  }
}

}  // namespace feo
