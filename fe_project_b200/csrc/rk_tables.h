// Host-side Runge-Kutta scheme tables of the time integrator the dynamics step is written against
// (reference: FElib/src/common/scale_timeint_rk_butcher_tab.F90:27-324, `timeint_rk` in
// scale_timeint_rk.F90:35-98).  Same scheme names, same coefficient semantics:
//   low-storage SSP schemes are given in Shu-Osher form (sig, gam) and converted to Butcher form,
//   ERK_4s4o and the additive IMEX schemes are given in Butcher form.
#pragma once
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

namespace fedg {

struct RKTable {
  std::string name;
  int nstage = 0, tend_buf_size = 1;
  bool low_storage = false, imex = false;
  std::vector<double> a_ex, b_ex, a_im, b_im, sig, gam;  // row-major
  std::vector<int> indmap;

  double& aex(int i, int j) { return a_ex[size_t(i) * nstage + j]; }
  double& aim(int i, int j) { return a_im[size_t(i) * nstage + j]; }
  double& sg(int i, int j) { return sig[size_t(i) * nstage + j]; }
  double& gm(int i, int j) { return gam[size_t(i) * nstage + j]; }
  double aex(int i, int j) const { return a_ex[size_t(i) * nstage + j]; }
  double aim(int i, int j) const { return a_im[size_t(i) * nstage + j]; }
  double sg(int i, int j) const { return sig[size_t(i) * nstage + j]; }
  double gm(int i, int j) const { return gam[size_t(i) * nstage + j]; }

  void alloc(int s, int bufsize, bool ls, bool ix) {
    nstage = s; tend_buf_size = bufsize; low_storage = ls; imex = ix;
    a_ex.assign(size_t(s) * s, 0.0); a_im.assign(size_t(s) * s, 0.0);
    b_ex.assign(s, 0.0); b_im.assign(s, 0.0);
    sig.assign(size_t(s + 1) * s, 0.0); gam.assign(size_t(s + 1) * s, 0.0);
    indmap.assign(s, 0);
  }

  // (I - sig[0:s]) a = gam[0:s]; b = gam[s] + sig[s] a      (ShuOsher2Butcher)
  void shu_osher_to_butcher() {
    const int s = nstage;
    for (int j = 0; j < s; ++j)
      for (int i = 0; i < s; ++i) {
        double r = gm(i, j);
        for (int k = 0; k < i; ++k) r += sg(i, k) * aex(k, j);
        aex(i, j) = r;
      }
    for (int j = 0; j < s; ++j) {
      double r = gm(s, j);
      for (int k = 0; k < s; ++k) r += sg(s, k) * aex(k, j);
      b_ex[j] = r;
    }
  }

  bool init(const std::string& scheme) {
    name = scheme;
    struct SO { int row, col; double sig, gam; };
    auto fill = [&](std::initializer_list<SO> l) {
      for (const SO& e : l) { if (e.sig != 0.0) sg(e.row, e.col) = e.sig; if (e.gam != 0.0) gm(e.row, e.col) = e.gam; }
      shu_osher_to_butcher();
    };
    if (scheme == "ERK_1s1o" || scheme == "ERK_Euler") {
      alloc(1, 1, false, false); b_ex[0] = 1.0; sg(1, 0) = 1.0; gm(1, 0) = 1.0;
    } else if (scheme == "ERK_4s4o" || scheme == "ERK_RK4") {
      alloc(4, 1, false, false);
      aex(1, 0) = 0.5; aex(2, 1) = 0.5; aex(3, 2) = 1.0;
      b_ex = {1.0 / 6.0, 2.0 / 6.0, 2.0 / 6.0, 1.0 / 6.0};
    } else if (scheme == "ERK_SSP_2s2o") {
      alloc(2, 1, true, false);
      fill({{1, 0, 1.0, 1.0}, {2, 0, 0.5, 0.0}, {2, 1, 0.5, 0.5}});
    } else if (scheme == "ERK_SSP_3s3o") {
      alloc(3, 1, true, false);
      fill({{1, 0, 1.0, 1.0}, {2, 0, 0.75, 0.0}, {2, 1, 0.25, 0.25}, {3, 0, 1.0 / 3.0, 0.0}, {3, 2, 2.0 / 3.0, 2.0 / 3.0}});
    } else if (scheme == "ERK_SSP_4s3o") {
      alloc(4, 1, true, false);
      fill({{1, 0, 1.0, 0.5}, {2, 1, 1.0, 0.5}, {3, 0, 2.0 / 3.0, 0.0}, {3, 2, 1.0 / 3.0, 1.0 / 6.0}, {4, 3, 1.0, 0.5}});
    } else if (scheme == "ERK_SSP_5s3o_2N2*") {
      alloc(5, 1, true, false);
      fill({{1, 0, 1.0, 0.465388589249323}, {2, 1, 1.0, 0.465388589249323},
            {3, 0, 0.682342861037239, 0.0}, {3, 2, 0.317657138962761, 0.124745797313998},
            {4, 3, 1.0, 0.465388589249323}, {5, 0, 0.045230974482400, 0.0},
            {5, 4, 0.954769025517600, 0.154263303748666}});
    } else if (scheme == "ERK_SSP_10s4o_2N") {
      alloc(10, 1, true, false);
      for (int n = 0; n < 4; ++n) { sg(n + 1, n) = 1.0; gm(n + 1, n) = 1.0 / 6.0; }
      sg(5, 0) = 3.0 / 5.0; sg(5, 4) = 2.0 / 5.0; gm(5, 4) = 1.0 / 15.0;
      for (int n = 5; n < 9; ++n) { sg(n + 1, n) = 1.0; gm(n + 1, n) = 1.0 / 6.0; }
      sg(10, 0) = 0.2 * 0.2; sg(10, 4) = 1.8 * 0.2; sg(10, 9) = 3.0 * 0.2;
      gm(10, 4) = 1.8 / 30.0; gm(10, 9) = 3.0 / 30.0;
      shu_osher_to_butcher();
    } else if (scheme == "IMEX_ARK232") {
      alloc(3, 3, false, true);
      const double r2 = std::sqrt(2.0);
      const double alp = (3.0 + 2.0 * r2) / 6.0, g = 1.0 - 1.0 / r2, del = 1.0 / (2.0 * r2);
      aex(1, 0) = 2.0 * g; aex(2, 0) = 1.0 - alp; aex(2, 1) = alp;
      b_ex = {del, del, g};
      aim(1, 0) = g; aim(1, 1) = g; aim(2, 0) = del; aim(2, 1) = del; aim(2, 2) = g;
      b_im = b_ex; indmap = {0, 1, 2};
    } else if (scheme == "IMEX_ARK324") {
      alloc(4, 4, false, true);
      aex(1, 0) = 1767732205903.0 / 2027836641118.0;
      aex(2, 0) = 5535828885825.0 / 10492691773637.0; aex(2, 1) = 788022342437.0 / 10882634858940.0;
      aex(3, 0) = 6485989280629.0 / 16251701735622.0; aex(3, 1) = -4246266847089.0 / 9704473918619.0;
      aex(3, 2) = 10755448449292.0 / 10357097424841.0;
      b_ex = {1471266399579.0 / 7840856788654.0, -4482444167858.0 / 7529755066697.0,
              11266239266428.0 / 11593286722821.0, 1767732205903.0 / 4055673282236.0};
      const double d = 1767732205903.0 / 4055673282236.0;
      aim(1, 0) = d; aim(1, 1) = d;
      aim(2, 0) = 2746238789719.0 / 10658868560708.0; aim(2, 1) = -640167445237.0 / 6845629431997.0; aim(2, 2) = d;
      for (int j = 0; j < 4; ++j) aim(3, j) = b_ex[j];
      b_im = b_ex; indmap = {0, 1, 2, 3};
    } else {
      return false;
    }
    return true;
  }
};

}  // namespace fedg
