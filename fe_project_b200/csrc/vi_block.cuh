// Vertical-implicit block of ONE column-element for p = 7, written as host/device functions of the row index, so that the same
// arithmetic runs (i) in vi_column2_kernel, where the two lanes of a column own rows 0..3 / 4..7, and (ii) on the CPU in
// tests/vi_block_host.cpp, where a plain loop over the rows is compared with the oracle (no GPU needed to validate the algebra).
//
// Reference: construct_matbnd / eval_Ax / vi_cal_del_flux_dyn / solve of scale_atm_dyn_dgm_nonhydro3d_rhot_hevi_common_2.F90:111-1328
// and the dense solver solve_Nnode8_var3 of scale_atm_dyn_dgm_hevi_common_linalgebra.F90:2296-2445.  The reference stores the 24 x 24
// block D and the 24 x 3 couplings L, U of every column-element and factorises D by partial-pivot LU.  Here the block is never
// stored.  With x = (rho_l, w_l, theta_l), l = 0..7 the vertical nodes, its rows are (flat geometry, GsqrtV = 1)
//
//   rho_l   + dfac (D w)_l                                                            + lw0_l B0[0].zb + lw1_l B7[0].zt = Rrho_l
//   w_l     + gfac (VP rho)_l + dfac (D (dpd o theta))_l                              + lw0_l B0[1].zb + lw1_l B7[1].zt = Rw_l
//   theta_l - dfac (D (s o rho))_l + dfac (D (pot o w))_l + dfac (D (wt o theta))_l   + lw0_l B0[2].zb + lw1_l B7[2].zt = Rth_l
//
// dfac = impl_fac E33, gfac = impl_fac GRAV, D = D1D, VP = IntrpMat_VPOrdM1, s = pot wt, lw0 / lw1 the lifting weights of the bottom /
// top face, zb = (rho_0, w_0, theta_0), zt = (rho_7, w_7, theta_7), B0 / B7 the 3 x 3 face matrices (Rusanov penalty + the
// block-Thomas elimination of the element below, which only touches zb), four right-hand sides (b and the three columns of U).
// Elimination order: rho (its rows are the identity + the constant matrix dfac D + two lifted face terms: closed form), then theta
// (8 x 8, partial pivoting), then the Schur complement in w (8 x 8, partial pivoting).  Against the reference's pivoting over the
// whole block the solution differs by <= 5e-13 relative for vertical acoustic CFL up to ~100 (tools/vi_block_experiment.py).
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define VI_HD __host__ __device__ __forceinline__
#define VI_UNROLL _Pragma("unroll")
#else
#define VI_HD inline
#define VI_UNROLL
#endif

namespace fedg {
namespace vib {

constexpr int N = 8;          // vertical nodes
constexpr int NR = 4;         // right-hand sides: b, U(:,rho), U(:,w), U(:,theta)
constexpr int NLF = 13;       // columns of a linear form: w_0..7, theta_0, rhs 0..3

// constant tables of the element (the same for every column): built once per block / once on the host
struct Tables {
  double D[N * N];        // D1D[l][j]
  double VP[N * N];       // VPOrdM1[l][j]
  double VPD[N * N];      // VP . D
  double lw0[N], lw1[N];  // lifting weights of the bottom / top face
  double VPlw0[N], VPlw1[N];
};
VI_HD void build_tables(const double* D, const double* VP, const double* Lw /* [l][side] */, Tables& T) {
  for (int i = 0; i < N * N; ++i) { T.D[i] = D[i]; T.VP[i] = VP[i]; }
  for (int l = 0; l < N; ++l) { T.lw0[l] = Lw[2 * l]; T.lw1[l] = Lw[2 * l + 1]; }
  for (int l = 0; l < N; ++l) {
    double a0 = 0.0, a1 = 0.0;
    for (int m = 0; m < N; ++m) { a0 += VP[l * N + m] * Lw[2 * m]; a1 += VP[l * N + m] * Lw[2 * m + 1]; }
    T.VPlw0[l] = a0; T.VPlw1[l] = a1;
    for (int j = 0; j < N; ++j) {
      double a = 0.0;
      for (int m = 0; m < N; ++m) a += VP[l * N + m] * D[m * N + j];
      T.VPD[l * N + j] = a;
    }
  }
}

// state of the neighbouring nodes across the two faces
struct FaceNbr {
  // node below the bottom face (top node of the element below) / above the top face (bottom node of the element above): Jacobian factors
  double potn_b, wtn_b, dpdn_b;
  double potn_t, wtn_t, dpdn_t;
  double g[3][NR];        // solution of the element below at its top node: unknown (rho, w, theta) x (b, G columns); used when !bot
};

// per-column-element scalars of the linear system
struct Coef {
  double dfac, gfac;
  double B0[3][3], B7[3][3];       // face matrices: row variable x (rho, w, theta) of the face node
  double rb[3];                    // rhs 0: lw0_l rb[a]   (block-Thomas: - L d)
  double U[3][3];                  // rhs 1+b: lw1_l U[a][b]   (coupling to the element above)
  // rho elimination
  double k00, k01, k10, k11;       // inverse of the 2 x 2 system in (rho_0, rho_7)
  double kap0b, kap0t, kap7b, kap7t;
};

// Face matrices (construct_matbnd :774-871 with the signs of nz = -1 / +1 folded in) and the block-Thomas terms (solve :385-416).
//   hb2 = 0.5 impl_fac Fscale_bottom, ht2 = 0.5 impl_fac Fscale_top; alph_b / alph_t the frozen dissipation coefficients;
//   pot, wt, dpd, s at the own face nodes 0 / 7.
VI_HD void face_coef(Coef& C, bool bot, bool top, double hb2, double ht2, double alph_b, double alph_t, double pot0, double wt0,
                     double dpd0, double pot7, double wt7, double dpd7, const FaceNbr& F) {
  const double s0 = pot0 * wt0, s7 = pot7 * wt7;
  double Lam[3][NR];
  VI_UNROLL for (int a = 0; a < 3; ++a) VI_UNROLL for (int c = 0; c < NR; ++c) Lam[a][c] = 0.0;
  if (!bot) {
    const double sn = F.potn_b * F.wtn_b;
    // L = lw0_l hb2 lam, lam = [[-alph_b, -1, 0], [0, -alph_b, -dpdn], [sn, -potn, -alph_b - wtn]];  Lam[a][c] = - sum_b lam[a][b] g[b][c]
    VI_UNROLL for (int c = 0; c < NR; ++c) {
      Lam[0][c] = alph_b * F.g[0][c] + F.g[1][c];
      Lam[1][c] = alph_b * F.g[1][c] + F.dpdn_b * F.g[2][c];
      Lam[2][c] = -sn * F.g[0][c] + F.potn_b * F.g[1][c] + (alph_b + F.wtn_b) * F.g[2][c];
    }
  }
  // column order of Lam: 0 = rhs, 1 = rho_0, 2 = w_0, 3 = theta_0
  C.B0[0][0] = (bot ? 0.0 : hb2 * alph_b) + hb2 * Lam[0][1];
  C.B0[0][1] = (bot ? 2.0 * hb2 : hb2) + hb2 * Lam[0][2];
  C.B0[0][2] = hb2 * Lam[0][3];
  C.B0[1][0] = hb2 * Lam[1][1];
  C.B0[1][1] = (bot ? 2.0 * hb2 * alph_b : hb2 * alph_b) + hb2 * Lam[1][2];
  C.B0[1][2] = (bot ? 0.0 : hb2 * dpd0) + hb2 * Lam[1][3];
  C.B0[2][0] = (bot ? -2.0 * hb2 * s0 : -hb2 * s0) + hb2 * Lam[2][1];
  C.B0[2][1] = (bot ? 2.0 * hb2 * pot0 : hb2 * pot0) + hb2 * Lam[2][2];
  C.B0[2][2] = (bot ? 2.0 * hb2 * wt0 : hb2 * (alph_b + wt0)) + hb2 * Lam[2][3];
  VI_UNROLL for (int a = 0; a < 3; ++a) C.rb[a] = hb2 * Lam[a][0];
  C.B7[0][0] = top ? 0.0 : ht2 * alph_t;
  C.B7[0][1] = top ? -2.0 * ht2 : -ht2;
  C.B7[0][2] = 0.0;
  C.B7[1][0] = 0.0;
  C.B7[1][1] = top ? 2.0 * ht2 * alph_t : ht2 * alph_t;
  C.B7[1][2] = top ? 0.0 : -ht2 * dpd7;
  C.B7[2][0] = (top ? 2.0 : 1.0) * ht2 * s7;
  C.B7[2][1] = top ? -2.0 * ht2 * pot7 : -ht2 * pot7;
  C.B7[2][2] = top ? -2.0 * ht2 * wt7 : ht2 * (alph_t - wt7);
  VI_UNROLL for (int a = 0; a < 3; ++a) VI_UNROLL for (int b = 0; b < 3; ++b) C.U[a][b] = 0.0;
  if (!top) {
    const double sn = F.potn_t * F.wtn_t;
    C.U[0][0] = -ht2 * alph_t; C.U[0][1] = ht2;             C.U[0][2] = 0.0;
    C.U[1][0] = 0.0;           C.U[1][1] = -ht2 * alph_t;   C.U[1][2] = ht2 * F.dpdn_t;
    C.U[2][0] = -ht2 * sn;     C.U[2][1] = ht2 * F.potn_t;  C.U[2][2] = ht2 * (-alph_t + F.wtn_t);
  }
}

// rho elimination, step 1: the 2 x 2 system of the face nodes
VI_HD void rho_pivots(Coef& C, const Tables& T) {
  const double a00 = 1.0 + T.lw0[0] * C.B0[0][0], a01 = T.lw1[0] * C.B7[0][0];
  const double a10 = T.lw0[7] * C.B0[0][0], a11 = 1.0 + T.lw1[7] * C.B7[0][0];
  const double rdet = 1.0 / (a00 * a11 - a01 * a10);
  C.k00 = a11 * rdet; C.k01 = -a01 * rdet; C.k10 = -a10 * rdet; C.k11 = a00 * rdet;
  C.kap0b = C.k00 * T.lw0[0] + C.k01 * T.lw0[7]; C.kap0t = C.k00 * T.lw1[0] + C.k01 * T.lw1[7];
  C.kap7b = C.k10 * T.lw0[0] + C.k11 * T.lw0[7]; C.kap7t = C.k10 * T.lw1[0] + C.k11 * T.lw1[7];
}

// The four linear forms the eliminated density leaves behind, over the columns c = (w_0..7, theta_0, rhs 0..3):
//   rho_0, rho_7 and phi0 = B0[0].zb, phi7 = B7[0].zt  with  rho_l = Rrho_l - dfac (D w)_l - lw0_l phi0 - lw1_l phi7.
// The density rows have the right-hand sides Rrho_l[0] = Rrho0[l] (rhs 0) and Rrho_l[1+b] = lw1_l U[0][b].  out[4] = column c of
// (rho0, rho7, phi0, phi7); the rhs entries are the VALUES of the forms for unit right-hand side r (they move to the other side with a
// minus sign in the rows).
VI_HD double rrho(const Coef& C, const Tables& T, int l, int r, double Rrho0_l) { return r == 0 ? Rrho0_l : T.lw1[l] * C.U[0][r - 1]; }
VI_HD void rho_form_col(const Coef& C, const Tables& T, int c, double Rrho0_0, double Rrho0_7, double* out /* [4] */) {
  double x0, x7;
  if (c < N) {
    x0 = -C.dfac * T.D[0 * N + c] - (c == 0 ? T.lw0[0] * C.B0[0][1] : 0.0) - (c == 7 ? T.lw1[0] * C.B7[0][1] : 0.0);
    x7 = -C.dfac * T.D[7 * N + c] - (c == 0 ? T.lw0[7] * C.B0[0][1] : 0.0) - (c == 7 ? T.lw1[7] * C.B7[0][1] : 0.0);
  } else if (c == 8) {
    x0 = -T.lw0[0] * C.B0[0][2]; x7 = -T.lw0[7] * C.B0[0][2];
  } else {
    x0 = rrho(C, T, 0, c - 9, Rrho0_0); x7 = rrho(C, T, 7, c - 9, Rrho0_7);
  }
  out[0] = C.k00 * x0 + C.k01 * x7;
  out[1] = C.k10 * x0 + C.k11 * x7;
  out[2] = C.B0[0][0] * out[0] + (c == 0 ? C.B0[0][1] : 0.0) + (c == 8 ? C.B0[0][2] : 0.0);
  out[3] = C.B7[0][0] * out[1] + (c == 7 ? C.B7[0][1] : 0.0);
}
VI_HD void rho_forms(const Coef& C, const Tables& T, double Rrho0_0, double Rrho0_7, double (*LF)[NLF]) {
  VI_UNROLL for (int c = 0; c < NLF; ++c) {
    double o[4];
    rho_form_col(C, T, c, Rrho0_0, Rrho0_7, o);
    VI_UNROLL for (int f = 0; f < 4; ++f) LF[f][c] = o[f];
  }
}

// right-hand sides of the three rows of node l: base[a] = impl_fac * A_v - var0 + q (eval_Ax :306-317) for rhs 0
VI_HD void row_rhs(const Coef& C, const Tables& T, int l, const double base[3], double R[3][NR]) {
  VI_UNROLL for (int a = 0; a < 3; ++a) {
    R[a][0] = base[a] + T.lw0[l] * C.rb[a];
    VI_UNROLL for (int b = 0; b < 3; ++b) R[a][1 + b] = T.lw1[l] * C.U[a][b];
  }
}

// Row l of [S_thth | S_thw | RHS_th] after the density has been eliminated (20 entries).
//   pot, wt, s: node vectors; Rrho0[m]: rhs 0 of the density rows; Rth[r]: right-hand sides of the own theta row; LF: rho_forms
VI_HD void theta_row(const Coef& C, const Tables& T, int l, const double* pot, const double* wt, const double* s,
                     const double* Rrho0, const double* Rth, const double (*LF)[NLF], double* A /* [20] */) {
  const double* Dl = T.D + l * N;
  double a0 = 0.0, a7 = 0.0;                    // dfac (D (s o lw0))_l, dfac (D (s o lw1))_l
  VI_UNROLL for (int m = 0; m < N; ++m) { a0 += Dl[m] * (s[m] * T.lw0[m]); a7 += Dl[m] * (s[m] * T.lw1[m]); }
  a0 *= C.dfac; a7 *= C.dfac;
  const double c0 = T.lw0[l] * C.B0[2][0], c7 = T.lw1[l] * C.B7[2][0];      // face terms on rho_0 / rho_7
  // lf[c] = a0 phi0[c] + a7 phi7[c] + c0 rho0[c] + c7 rho7[c]
  double lf[NLF];
  VI_UNROLL for (int c = 0; c < NLF; ++c) lf[c] = a0 * LF[2][c] + a7 * LF[3][c] + c0 * LF[0][c] + c7 * LF[1][c];
  // theta columns
  VI_UNROLL for (int j = 0; j < N; ++j) A[j] = (j == l ? 1.0 : 0.0) + C.dfac * Dl[j] * wt[j];
  A[0] += T.lw0[l] * C.B0[2][2] + lf[8];
  A[7] += T.lw1[l] * C.B7[2][2];
  // w columns: dfac D_lj pot_j + dfac^2 sum_m D_lm s_m D_mj + lf[j] + face
  double t[N];
  VI_UNROLL for (int m = 0; m < N; ++m) t[m] = C.dfac * Dl[m] * s[m];
  VI_UNROLL for (int j = 0; j < N; ++j) {
    double acc = Dl[j] * pot[j];
    VI_UNROLL for (int m = 0; m < N; ++m) acc += t[m] * T.D[m * N + j];
    A[8 + j] = C.dfac * acc + lf[j];
  }
  A[8 + 0] += T.lw0[l] * C.B0[2][1];
  A[8 + 7] += T.lw1[l] * C.B7[2][1];
  // right-hand sides: Rth + dfac (D (s o Rrho))_l - lf[9 + r];  Rrho[:, 1+b] = lw1 U[0][b]  ->  dfac (D (s o lw1))_l U[0][b] = a7 U[0][b]
  {
    double acc = 0.0;
    VI_UNROLL for (int m = 0; m < N; ++m) acc += t[m] * Rrho0[m];
    A[16] = Rth[0] + acc - lf[9];
  }
  VI_UNROLL for (int b = 0; b < 3; ++b) A[17 + b] = Rth[1 + b] + a7 * C.U[0][b] - lf[10 + b];
}

// Row l of the Schur complement [H | rhs_H] = [S_ww | RHS_w] - S_wth X, X = S_thth^-1 [S_thw | RHS_th] (8 x 12, row = theta unknown).
VI_HD void schur_row(const Coef& C, const Tables& T, int l, const double* dpd, const double* Rrho0, const double* Rw,
                     const double (*LF)[NLF], const double (*X)[12], double* H /* [12] */) {
  const double* Dl = T.D + l * N;
  const double a0 = -C.gfac * T.VPlw0[l], a7 = -C.gfac * T.VPlw1[l];
  const double c0 = T.lw0[l] * C.B0[1][0];
  double lf[NLF];
  VI_UNROLL for (int c = 0; c < NLF; ++c) lf[c] = a0 * LF[2][c] + a7 * LF[3][c] + c0 * LF[0][c];
  // S_wth row: dfac D_lj dpd_j + face (theta_0, theta_7) + lf[8] on theta_0
  double sw[N];
  VI_UNROLL for (int j = 0; j < N; ++j) sw[j] = C.dfac * Dl[j] * dpd[j];
  sw[0] += T.lw0[l] * C.B0[1][2] + lf[8];
  sw[7] += T.lw1[l] * C.B7[1][2];
  const double gd = C.gfac * C.dfac;
  VI_UNROLL for (int j = 0; j < N; ++j) {
    double h = (j == l ? 1.0 : 0.0) - gd * T.VPD[l * N + j] + lf[j];
    if (j == 0) h += T.lw0[l] * C.B0[1][1];
    if (j == 7) h += T.lw1[l] * C.B7[1][1];
    VI_UNROLL for (int m = 0; m < N; ++m) h -= sw[m] * X[m][j];
    H[j] = h;
  }
  VI_UNROLL for (int r = 0; r < NR; ++r) {
    double vp = 0.0;
    if (r == 0) { VI_UNROLL for (int m = 0; m < N; ++m) vp += T.VP[l * N + m] * Rrho0[m]; }
    else vp = T.VPlw1[l] * C.U[0][r - 1];
    double h = Rw[r] - C.gfac * vp - lf[9 + r];
    VI_UNROLL for (int m = 0; m < N; ++m) h -= sw[m] * X[m][8 + r];
    H[8 + r] = h;
  }
}

// theta_k[r] = X[k][8 + r] - sum_j X[k][j] w_j[r]
VI_HD void theta_solve(int k, const double (*X)[12], const double (*w)[NR], double* th /* [NR] */) {
  VI_UNROLL for (int r = 0; r < NR; ++r) {
    double a = X[k][8 + r];
    VI_UNROLL for (int j = 0; j < N; ++j) a -= X[k][j] * w[j][r];
    th[r] = a;
  }
}

// rho_l[r] = Rrho_l[r] - dfac (D w[:, r])_l - lw0_l phi0[r] - lw1_l phi7[r];  phi[r] = sum_j phi_w[j] w_j[r] + phi_t theta_0[r] + phi_r[r]
VI_HD void rho_solve(const Coef& C, const Tables& T, int l, double Rrho0_l, const double (*LF)[NLF], const double (*w)[NR],
                     const double* th0 /* [NR] */, double* rho /* [NR] */) {
  VI_UNROLL for (int r = 0; r < NR; ++r) {
    double dw = 0.0, p0 = LF[2][8] * th0[r] + LF[2][9 + r], p7 = LF[3][8] * th0[r] + LF[3][9 + r];
    VI_UNROLL for (int j = 0; j < N; ++j) { dw += T.D[l * N + j] * w[j][r]; p0 += LF[2][j] * w[j][r]; p7 += LF[3][j] * w[j][r]; }
    rho[r] = rrho(C, T, l, r, Rrho0_l) - C.dfac * dw - T.lw0[l] * p0 - T.lw1[l] * p7;
  }
}

}  // namespace vib
}  // namespace fedg
