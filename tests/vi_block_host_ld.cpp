// long-double build of the host harness (test infrastructure): a 19-digit solution of the same column systems, against which the
// accuracy of BOTH double-precision solvers -- the oracle's partial-pivot LU and the block elimination of vi_block.cuh -- is measured
// (tests/test_vi_block_host.py::test_block_elimination_is_as_accurate_as_the_reference_lu).
#include <cmath>
#include <cstddef>
#include <vector>
#define double long double
#define vib_cal_vi vib_cal_vi_ld_impl
#include "vi_block_host.cpp"
#undef double
#undef vib_cal_vi
extern "C" int vib_cal_vi_ld(int Ne2D, int NeZ, const double* const* q0, const double* const* qcur, const double* dens_hyd, const double* pres_hyd,
               const double* escale33, const double* fscale_b, const double* fscale_t, const double* D, const double* VP, const double* Lw,
               const double* consts5, double ifac, double* const* kim) {
  const size_t n = size_t(Ne2D) * NeZ * 512, ne = size_t(Ne2D) * NeZ;
  auto cv = [](const double* p, size_t m) { return std::vector<long double>(p, p + m); };
  std::vector<long double> a[5], b[5], o[5];
  const long double* pa[5]; const long double* pb[5]; long double* po[5];
  for (int v = 0; v < 5; ++v) { a[v] = cv(q0[v], n); b[v] = cv(qcur[v], n); o[v].assign(n, 0); pa[v] = a[v].data(); pb[v] = b[v].data(); po[v] = o[v].data(); }
  auto dh = cv(dens_hyd, n), ph = cv(pres_hyd, n), e3 = cv(escale33, ne), fb = cv(fscale_b, ne), ft = cv(fscale_t, ne), d = cv(D, 64), vp = cv(VP, 64), lw = cv(Lw, 16), cs = cv(consts5, 5);
  int rc = vib_cal_vi_ld_impl(Ne2D, NeZ, pa, pb, dh.data(), ph.data(), e3.data(), fb.data(), ft.data(), d.data(), vp.data(), lw.data(), cs.data(), (long double)ifac, po);
  for (int v = 0; v < 5; ++v) for (size_t i = 0; i < n; ++i) kim[v][i] = (double)o[v][i];
  return rc;
}
