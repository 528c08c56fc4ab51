// TEST INFRASTRUCTURE ONLY (see fe_oracle.hpp).  Runge-Kutta tableaux and stage updates.
#include "fe_oracle.hpp"

#include <stdexcept>

namespace feo {

// common/scale_timeint_rk_butcher_tab.F90:60-324.  Shu-Osher (sig, gam) coefficients for the SSP
// schemes, Butcher (a, b) for ERK_4s4o and the two additive IMEX schemes; a/b of the SSP schemes
// follow from ShuOsher2Butcher (:297-324): (I - sig) a = gam.
bool RKScheme::init(const std::string& scheme) {
  name = scheme; low_storage = false; imex = false;
  if (scheme == "ERK_1s1o" || scheme == "ERK_Euler") { nstage = 1; tend_buf_size = 1; }
  else if (scheme == "ERK_4s4o" || scheme == "ERK_RK4") { nstage = 4; tend_buf_size = 1; }
  else if (scheme == "ERK_SSP_2s2o") { nstage = 2; tend_buf_size = 1; low_storage = true; }
  else if (scheme == "ERK_SSP_3s3o") { nstage = 3; tend_buf_size = 1; low_storage = true; }
  else if (scheme == "ERK_SSP_4s3o") { nstage = 4; tend_buf_size = 1; low_storage = true; }
  else if (scheme == "ERK_SSP_5s3o_2N2*") { nstage = 5; tend_buf_size = 1; low_storage = true; }
  else if (scheme == "ERK_SSP_10s4o_2N") { nstage = 10; tend_buf_size = 1; low_storage = true; }
  else if (scheme == "IMEX_ARK232") { nstage = 3; tend_buf_size = 3; imex = true; }
  else if (scheme == "IMEX_ARK324") { nstage = 4; tend_buf_size = 4; imex = true; }
  else return false;
  const int s = nstage;
  a_ex.assign(size_t(s) * s, 0.0); b_ex.assign(s, 0.0); c_ex.assign(s, 0.0);
  a_im.assign(size_t(s) * s, 0.0); b_im.assign(s, 0.0); c_im.assign(s, 0.0);
  sig.assign(size_t(s + 1) * s, 0.0); gam.assign(size_t(s + 1) * s, 0.0);
  indmap.assign(s, 0);
  auto A = [&](int i, int j) -> double& { return a_ex[(i - 1) * s + (j - 1)]; };
  auto AI = [&](int i, int j) -> double& { return a_im[(i - 1) * s + (j - 1)]; };
  auto S = [&](int i, int j) -> double& { return sig[(i - 1) * s + (j - 1)]; };
  auto G = [&](int i, int j) -> double& { return gam[(i - 1) * s + (j - 1)]; };
  bool shu_osher = false;
  if (scheme == "ERK_1s1o" || scheme == "ERK_Euler") {
    b_ex[0] = 1.0; S(2, 1) = 1.0; G(2, 1) = 1.0;
  } else if (scheme == "ERK_4s4o" || scheme == "ERK_RK4") {
    A(2, 1) = 0.5; A(3, 2) = 0.5; A(4, 3) = 1.0;
    b_ex = {1.0 / 6.0, 2.0 / 6.0, 2.0 / 6.0, 1.0 / 6.0};
  } else if (scheme == "ERK_SSP_2s2o") {
    S(2, 1) = 1.0; S(3, 1) = 0.5; S(3, 2) = 0.5; G(2, 1) = 1.0; G(3, 2) = 0.5; shu_osher = true;
  } else if (scheme == "ERK_SSP_3s3o") {
    S(2, 1) = 1.0; S(3, 1) = 3.0 / 4.0; S(3, 2) = 1.0 / 4.0; S(4, 1) = 1.0 / 3.0; S(4, 3) = 2.0 / 3.0;
    G(2, 1) = 1.0; G(3, 2) = 1.0 / 4.0; G(4, 3) = 2.0 / 3.0; shu_osher = true;
  } else if (scheme == "ERK_SSP_4s3o") {
    S(2, 1) = 1.0; S(3, 2) = 1.0; S(4, 1) = 2.0 / 3.0; S(4, 3) = 1.0 / 3.0; S(5, 4) = 1.0;
    G(2, 1) = 0.5; G(3, 2) = 0.5; G(4, 3) = 1.0 / 6.0; G(5, 4) = 0.5; shu_osher = true;
  } else if (scheme == "ERK_SSP_5s3o_2N2*") {
    S(2, 1) = 1.0; S(3, 2) = 1.0; S(4, 1) = 0.682342861037239; S(4, 3) = 0.317657138962761; S(5, 4) = 1.0;
    S(6, 1) = 0.045230974482400; S(6, 5) = 0.954769025517600;
    G(2, 1) = 0.465388589249323; G(3, 2) = 0.465388589249323; G(4, 3) = 0.124745797313998;
    G(5, 4) = 0.465388589249323; G(6, 5) = 0.154263303748666; shu_osher = true;
  } else if (scheme == "ERK_SSP_10s4o_2N") {
    for (int n = 1; n <= 4; ++n) { S(n + 1, n) = 1.0; G(n + 1, n) = 1.0 / 6.0; }
    S(6, 1) = 3.0 / 5.0; S(6, 5) = 2.0 / 5.0; G(6, 5) = 1.0 / 15.0;
    for (int n = 6; n <= 9; ++n) { S(n + 1, n) = 1.0; G(n + 1, n) = 1.0 / 6.0; }
    S(11, 1) = 0.2 * 0.2; S(11, 5) = 1.8 * 0.2; S(11, 10) = 3.0 * 0.2;
    G(11, 5) = 1.8 / 30.0; G(11, 10) = 3.0 / 30.0; shu_osher = true;
  } else if (scheme == "IMEX_ARK232") {
    double alp = (3.0 + 2.0 * std::sqrt(2.0)) / 6.0, gm = 1.0 - 1.0 / std::sqrt(2.0), del = 1.0 / (2.0 * std::sqrt(2.0));
    A(2, 1) = 2.0 * gm; A(3, 1) = 1.0 - alp; A(3, 2) = alp;
    b_ex = {del, del, gm};
    AI(2, 1) = gm; AI(2, 2) = gm; AI(3, 1) = del; AI(3, 2) = del; AI(3, 3) = gm;
    b_im = b_ex; indmap = {0, 1, 2};
  } else if (scheme == "IMEX_ARK324") {
    A(2, 1) = 1767732205903.0 / 2027836641118.0;
    A(3, 1) = 5535828885825.0 / 10492691773637.0; A(3, 2) = 788022342437.0 / 10882634858940.0;
    A(4, 1) = 6485989280629.0 / 16251701735622.0; A(4, 2) = -4246266847089.0 / 9704473918619.0;
    A(4, 3) = 10755448449292.0 / 10357097424841.0;
    b_ex = {1471266399579.0 / 7840856788654.0, -4482444167858.0 / 7529755066697.0,
            11266239266428.0 / 11593286722821.0, 1767732205903.0 / 4055673282236.0};
    AI(2, 1) = 1767732205903.0 / 4055673282236.0; AI(2, 2) = AI(2, 1);
    AI(3, 1) = 2746238789719.0 / 10658868560708.0; AI(3, 2) = -640167445237.0 / 6845629431997.0;
    AI(3, 3) = 1767732205903.0 / 4055673282236.0;
    for (int j = 1; j <= 4; ++j) AI(4, j) = b_ex[j - 1];
    b_im = b_ex; indmap = {0, 1, 2, 3};
  }
  if (shu_osher) {
    // solve (I - sig(1:s,:)) a = gam(1:s,:): the matrix is unit lower triangular -> forward substitution
    for (int i = 1; i <= s; ++i)
      for (int j = 1; j <= s; ++j) {
        double v = G(i, j);
        for (int k = 1; k < i; ++k) v += S(i, k) * A(k, j);
        A(i, j) = v;  // S(i,i) = 0 for all schemes here
      }
    for (int n = 1; n <= s; ++n) {
      double v = G(s + 1, n);
      for (int k = 1; k <= s; ++k) v += S(s + 1, k) * A(k, n);
      b_ex[n - 1] = v;
    }
  }
  for (int n = 1; n <= s; ++n) { double c = 0; for (int j = 1; j <= n; ++j) c += A(n, j); c_ex[n - 1] = c; }
  return true;
}

void TimeIntRK::init(const std::string& scheme, double dt_, int nvar_, size_t n_) {
  if (!sc.init(scheme)) throw std::runtime_error("unsupported RK scheme " + scheme);
  dt = dt_; nvar = nvar_; n = n_;
  tend_ex.assign(size_t(sc.tend_buf_size) * nvar * n, 0.0);
  if (sc.imex) tend_im.assign(size_t(sc.tend_buf_size) * nvar * n, 0.0);
  var0.assign(size_t(nvar) * n, 0.0); varTmp.assign(size_t(nvar) * n, 0.0);
}

void TimeIntRK::store_var0(const double* q, int var, size_t is, size_t ie) {  // scale_timeint_rk.F90:624
  double* v0 = &var0[size_t(var) * n];
  for (size_t i = is; i < ie; ++i) v0[i] = q[i];
}

// scale_timeint_rk.F90:2510-2560 (rk_storeimpl_general2D)
void TimeIntRK::store_implicit(int stage, double* q, int var, size_t is, size_t ie) {
  if (!sc.imex) return;
  double* v0 = &var0[size_t(var) * n]; double* vt = &varTmp[size_t(var) * n];
  const double* ki = tend_im_buf(var, sc.indmap[stage]);
  double c = dt * sc.A_im(stage, stage);
  if (stage == 0) for (size_t i = is; i < ie; ++i) { v0[i] = q[i]; vt[i] = q[i]; }
  for (size_t i = is; i < ie; ++i) q[i] = q[i] + c * ki[i];
}

// scale_timeint_rk.F90:1182-1266 (low storage) and :2201-2355 (general / IMEX)
void TimeIntRK::advance(int stage, double* q, int var, size_t is, size_t ie) {
  const int s = sc.nstage;
  double* v0 = &var0[size_t(var) * n]; double* vt = &varTmp[size_t(var) * n];
  if (sc.low_storage) {
    const double* k = tend_ex_buf(var, 0);
    double sig_ss = sc.SIG(stage + 1, stage), sig_Ns = sc.SIG(s, stage);
    double one_m = 1.0 - sig_ss, gam_ss = dt * sc.GAM(stage + 1, stage), gam_Ns = dt * sc.GAM(s, stage);
    if (stage == s - 1) {
      for (size_t i = is; i < ie; ++i) q[i] = vt[i] + sig_ss * q[i] + gam_ss * k[i];
      return;
    }
    if (stage == 0) for (size_t i = is; i < ie; ++i) { v0[i] = q[i]; vt[i] = 0.0; }
    const double EPS = 2.220446e-16;
    if (std::fabs(sig_Ns) > EPS || std::fabs(gam_Ns) > EPS)
      for (size_t i = is; i < ie; ++i) vt[i] = vt[i] + sig_Ns * q[i] + gam_Ns * k[i];
    for (size_t i = is; i < ie; ++i) q[i] = one_m * v0[i] + sig_ss * q[i] + gam_ss * k[i];
    return;
  }
  const int ind = sc.indmap[stage];
  if (s == 1) for (size_t i = is; i < ie; ++i) vt[i] = q[i];
  if (stage == s - 1) {
    double bex = sc.b_ex[stage] * dt;
    const double* ke = tend_ex_buf(var, ind);
    if (sc.imex) {
      double bim = sc.b_im[stage] * dt; const double* ki = tend_im_buf(var, ind);
      for (size_t i = is; i < ie; ++i) q[i] = vt[i] + bex * ke[i] + bim * ki[i];
    } else {
      for (size_t i = is; i < ie; ++i) q[i] = vt[i] + bex * ke[i];
    }
    return;
  }
  if (stage == 0 && !sc.imex) for (size_t i = is; i < ie; ++i) { v0[i] = q[i]; vt[i] = q[i]; }
  {
    double bex = sc.b_ex[stage] * dt; const double* ke = tend_ex_buf(var, ind);
    if (sc.imex) {
      double bim = sc.b_im[stage] * dt; const double* ki = tend_im_buf(var, ind);
      for (size_t i = is; i < ie; ++i) { q[i] = v0[i]; vt[i] = vt[i] + bex * ke[i] + bim * ki[i]; }
    } else {
      for (size_t i = is; i < ie; ++i) { q[i] = v0[i]; vt[i] = vt[i] + bex * ke[i]; }
    }
  }
  if (sc.tend_buf_size == 1 && !sc.imex) {
    double a = dt * sc.A_ex(stage + 1, stage); const double* ke = tend_ex_buf(var, 0);
    for (size_t i = is; i < ie; ++i) q[i] = v0[i] + a * ke[i];
  } else if (!sc.imex) {
    for (int ss = 0; ss <= stage; ++ss) {
      double a = dt * sc.A_ex(stage + 1, ss); const double* ke = tend_ex_buf(var, ss);
      for (size_t i = is; i < ie; ++i) q[i] = q[i] + a * ke[i];
    }
  } else {
    for (int ss = 0; ss <= stage; ++ss) {
      double ae = dt * sc.A_ex(stage + 1, ss), ai = dt * sc.A_im(stage + 1, ss);
      const double* ke = tend_ex_buf(var, ss); const double* ki = tend_im_buf(var, ss);
      for (size_t i = is; i < ie; ++i) q[i] = q[i] + ae * ke[i] + ai * ki[i];
    }
  }
}

}  // namespace feo
