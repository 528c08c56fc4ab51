// CPU harness of fe_project_b200/csrc/vi_block.cuh (test infrastructure): the block-eliminated vertical-implicit solve of every
// column of a flat mesh, built from the SAME host/device row functions the CUDA kernel vi_column2_kernel calls, so that the algebra of
// the kernel is validated against the oracle without a GPU (tests/test_vi_block_host.py).  Plain loops, no lanes, no shared memory.
//   g++ -O2 -ffp-contract=off -shared -fPIC -I fe_project_b200/csrc tests/vi_block_host.cpp -o tests/_vi_block_host.so
#include <cmath>
#include <cstddef>
#include <vector>

#include "vi_block.cuh"

using namespace fedg::vib;

namespace {
struct Consts { double GRAV, Rdry, CPdry, CVdry, PRES00; };
struct NodeQ { double rho0, w0, th0, u0, v0, dens, rhot, pot, wt, dpd, dpres_vol, a, dpf; };

NodeQ node_q(const Consts& c, double rho0, double w0, double th0, double u0, double v0, double dh, double rh, double ph) {
  NodeQ q;
  const double gm = c.CPdry / c.CVdry, rP0 = 1.0 / c.PRES00;
  q.rho0 = rho0; q.w0 = w0; q.th0 = th0; q.u0 = u0; q.v0 = v0;
  q.dens = dh + rho0; q.rhot = rh + th0; q.pot = q.rhot / q.dens;
  const double ptot = c.PRES00 * std::pow(c.Rdry * rP0 * q.rhot, gm);
  q.dpres_vol = ptot - ph;
  q.wt = w0 / q.dens;
  q.dpd = gm * ptot / q.rhot;
  q.a = std::fabs(w0 / q.dens) + std::sqrt(gm * ptot / q.dens);
  q.dpf = c.PRES00 * std::pow(c.Rdry * rP0 * q.dens * q.pot, gm) - ph;
  return q;
}

// Gauss-Jordan with partial pivoting on the leading n x n block of A (n rows, W columns): pivot = largest magnitude among the rows
// not used yet (first one wins ties); on return row rowof[k] scaled by 1 / pivot holds unknown k in the columns n .. W-1
void gauss_jordan(double* A, int n, int W, int* rowof) {
  bool used[N] = {false};
  for (int k = 0; k < n; ++k) {
    int p = -1; double best = -1.0;
    for (int i = 0; i < n; ++i) if (!used[i] && std::fabs(A[i * W + k]) > best) { best = std::fabs(A[i * W + k]); p = i; }
    used[p] = true; rowof[k] = p;
    const double rp = 1.0 / A[p * W + k];
    for (int j = k + 1; j < W; ++j) A[p * W + j] *= rp;
    for (int i = 0; i < n; ++i) {
      if (i == p) continue;
      const double m = A[i * W + k];
      for (int j = k + 1; j < W; ++j) A[i * W + j] -= m * A[p * W + j];
    }
  }
}
}  // namespace

extern "C" {

// fields: (Np * Ne) node-major arrays of the mesh (Ne = Ne2D * NeZ, element ke = ke2d + kz * Ne2D, node n = ij + 64 l)
// q0[5], qcur[5], kim[5] in the order DDENS, MOMX, MOMY, MOMZ, DRHOT.  escale33[Ne], fscale_b[Ne], fscale_t[Ne].
int vib_cal_vi(int Ne2D, int NeZ, const double* const* q0, const double* const* qcur, const double* dens_hyd, const double* pres_hyd,
               const double* escale33, const double* fscale_b, const double* fscale_t, const double* D, const double* VP, const double* Lw,
               const double* consts5, double ifac, double* const* kim) {
  const Consts c{consts5[0], consts5[1], consts5[2], consts5[3], consts5[4]};
  Tables T;
  build_tables(D, VP, Lw, T);
  const int ncol = Ne2D * 64;
  enum { V_DDENS = 0, V_MOMX = 1, V_MOMY = 2, V_MOMZ = 3, V_DRHOT = 4 };
#pragma omp parallel for
  for (int col = 0; col < ncol; ++col) {
    const int ke2d = col / 64, ij = col % 64;
    auto node = [&](int kz, int l) { return (size_t(ke2d) + size_t(kz) * Ne2D) * 512 + ij + 64 * l; };
    auto load = [&](int kz, NodeQ* q) {
      for (int l = 0; l < N; ++l) {
        const size_t n = node(kz, l);
        const double rh = c.PRES00 / c.Rdry * std::pow(pres_hyd[n] / c.PRES00, c.CVdry / c.CPdry);
        q[l] = node_q(c, q0[V_DDENS][n], q0[V_MOMZ][n], q0[V_DRHOT][n], q0[V_MOMX][n], q0[V_MOMY][n], dens_hyd[n], rh, pres_hyd[n]);
      }
    };
    std::vector<double> sd(size_t(NeZ) * 24), sG(size_t(NeZ) * 72), suv(size_t(NeZ) * 24);   // d, G, (du, dv, guv)
    NodeQ q[N], qn[N], prev{};
    double gprev[3][NR] = {{0}}, uvprev[3] = {0, 0, 0};
    load(0, q);
    for (int kz = 0; kz < NeZ; ++kz) {
      const int ke = ke2d + kz * Ne2D;
      const bool bot = kz == 0, top = kz == NeZ - 1;
      if (!top) load(kz + 1, qn);
      const double E33 = escale33[ke], Fs_b = fscale_b[ke], Fs_t = fscale_t[ke];
      const double alph_b = bot ? q[0].a : std::fmax(q[0].a, prev.a);
      const double alph_t = top ? q[7].a : std::fmax(q[7].a, qn[0].a);
      // exterior states (vi_cal_del_flux_dyn :1262-1322): slip-wall mirror at the column ends
      const NodeQ &M0 = q[0], &M7 = q[7];
      double rP_b, wP_b, tP_b, pP_b, dP_b, uP_b, vP_b, rP_t, wP_t, tP_t, pP_t, dP_t, uP_t, vP_t;
      if (bot) { rP_b = M0.rho0; wP_b = -M0.w0; tP_b = M0.th0; pP_b = M0.pot; dP_b = M0.dpf; uP_b = M0.u0; vP_b = M0.v0; }
      else { rP_b = prev.rho0; wP_b = prev.w0; tP_b = prev.th0; pP_b = prev.pot; dP_b = prev.dpf; uP_b = prev.u0; vP_b = prev.v0; }
      if (top) { rP_t = M7.rho0; wP_t = -M7.w0; tP_t = M7.th0; pP_t = M7.pot; dP_t = M7.dpf; uP_t = M7.u0; vP_t = M7.v0; }
      else { rP_t = qn[0].rho0; wP_t = qn[0].w0; tP_t = qn[0].th0; pP_t = qn[0].pot; dP_t = qn[0].dpf; uP_t = qn[0].u0; vP_t = qn[0].v0; }
      const double hb = 0.5 * Fs_b, ht = 0.5 * Fs_t;
      const double dl_r_b = hb * ((wP_b - M0.w0) * (-1.0) - alph_b * (rP_b - M0.rho0));
      const double dl_w_b = hb * ((dP_b - M0.dpf) * (-1.0) - alph_b * (wP_b - M0.w0));
      const double dl_t_b = hb * ((pP_b * wP_b - M0.pot * M0.w0) * (-1.0) - alph_b * (tP_b - M0.th0));
      const double dl_r_t = ht * ((wP_t - M7.w0) - alph_t * (rP_t - M7.rho0));
      const double dl_w_t = ht * ((dP_t - M7.dpf) - alph_t * (wP_t - M7.w0));
      const double dl_t_t = ht * ((pP_t * wP_t - M7.pot * M7.w0) - alph_t * (tP_t - M7.th0));
      const double dl_u_b = (-0.5 * Fs_b * alph_b) * (uP_b - M0.u0), dl_v_b = (-0.5 * Fs_b * alph_b) * (vP_b - M0.v0);
      const double dl_u_t = (-0.5 * Fs_t * alph_t) * (uP_t - M7.u0), dl_v_t = (-0.5 * Fs_t * alph_t) * (vP_t - M7.v0);
      double t_r[N], t_w[N], t_t[N], t_u[N], t_v[N];
      for (int l = 0; l < N; ++l) {
        double dz_r = 0, dz_t = 0, dz_w = 0, drho = 0;
        for (int p = 0; p < N; ++p) {
          dz_r += T.D[l * N + p] * q[p].w0; dz_t += T.D[l * N + p] * (q[p].pot * q[p].w0); dz_w += T.D[l * N + p] * q[p].dpres_vol;
          drho += T.VP[l * N + p] * q[p].rho0;
        }
        t_r[l] = -(E33 * dz_r + (T.lw0[l] * dl_r_b + T.lw1[l] * dl_r_t));
        t_t[l] = -(E33 * dz_t + (T.lw0[l] * dl_t_b + T.lw1[l] * dl_t_t));
        t_w[l] = -(E33 * dz_w + (T.lw0[l] * dl_w_b + T.lw1[l] * dl_w_t)) - c.GRAV * drho;
        t_u[l] = -(T.lw0[l] * dl_u_b + T.lw1[l] * dl_u_t); t_v[l] = -(T.lw0[l] * dl_v_b + T.lw1[l] * dl_v_t);
      }
      if (ifac == 0.0) {
        for (int l = 0; l < N; ++l) {
          const size_t n = node(kz, l);
          kim[V_DDENS][n] = t_r[l]; kim[V_MOMZ][n] = t_w[l]; kim[V_DRHOT][n] = t_t[l]; kim[V_MOMX][n] = t_u[l]; kim[V_MOMY][n] = t_v[l];
        }
      } else {
        Coef C;
        C.dfac = E33 * ifac; C.gfac = ifac * c.GRAV;
        FaceNbr F{};
        if (!bot) { F.potn_b = prev.pot; F.wtn_b = prev.wt; F.dpdn_b = prev.dpd; for (int a = 0; a < 3; ++a) for (int r = 0; r < NR; ++r) F.g[a][r] = gprev[a][r]; }
        if (!top) { F.potn_t = qn[0].pot; F.wtn_t = qn[0].wt; F.dpdn_t = qn[0].dpd; }
        const double hb2 = 0.5 * ifac * Fs_b, ht2 = 0.5 * ifac * Fs_t;
        face_coef(C, bot, top, hb2, ht2, alph_b, alph_t, q[0].pot, q[0].wt, q[0].dpd, q[7].pot, q[7].wt, q[7].dpd, F);
        rho_pivots(C, T);
        double pot[N], wt[N], dpd[N], s[N], R[N][3][NR], Rrho0[N];
        for (int l = 0; l < N; ++l) {
          const size_t n = node(kz, l);
          pot[l] = q[l].pot; wt[l] = q[l].wt; dpd[l] = q[l].dpd; s[l] = q[l].pot * q[l].wt;
          const double base[3] = {ifac * t_r[l] - q[l].rho0 + qcur[V_DDENS][n], ifac * t_w[l] - q[l].w0 + qcur[V_MOMZ][n],
                                  ifac * t_t[l] - q[l].th0 + qcur[V_DRHOT][n]};
          row_rhs(C, T, l, base, R[l]);
          Rrho0[l] = R[l][0][0];
        }
        double LF[4][NLF];
        rho_forms(C, T, Rrho0[0], Rrho0[7], LF);
        double A[N][20];
        for (int l = 0; l < N; ++l) theta_row(C, T, l, pot, wt, s, Rrho0, R[l][2], LF, A[l]);
        int rowof[N];
        gauss_jordan(&A[0][0], N, 20, rowof);
        double X[N][12];
        for (int k = 0; k < N; ++k) for (int cix = 0; cix < 12; ++cix) X[k][cix] = A[rowof[k]][8 + cix];
        double H[N][12];
        for (int l = 0; l < N; ++l) schur_row(C, T, l, dpd, Rrho0, R[l][1], LF, X, H[l]);
        gauss_jordan(&H[0][0], N, 12, rowof);
        double w[N][NR], th[N][NR], rho[N][NR];
        for (int k = 0; k < N; ++k) for (int r = 0; r < NR; ++r) w[k][r] = H[rowof[k]][8 + r];
        for (int k = 0; k < N; ++k) theta_solve(k, X, w, th[k]);
        for (int l = 0; l < N; ++l) rho_solve(C, T, l, Rrho0[l], LF, w, th[0], rho[l]);
        for (int l = 0; l < N; ++l) {
          const double* v3[3] = {rho[l], w[l], th[l]};
          for (int a = 0; a < 3; ++a) {
            sd[(size_t(kz) * N + l) * 3 + a] = v3[a][0];
            for (int b = 0; b < 3; ++b) sG[((size_t(kz) * N + l) * 3 + a) * 3 + b] = v3[a][1 + b];
          }
        }
        for (int r = 0; r < NR; ++r) { gprev[0][r] = rho[7][r]; gprev[1][r] = w[7][r]; gprev[2][r] = th[7][r]; }
        // (MOMX, MOMY): (I + ua0 e0^T + ua7 e7^T) x = [bu | bv | bg]   (construct_matbnd_uv :960-1003, solve_uv :640-674)
        double ua0[N], ua7[N], rhs[N][3];
        for (int l = 0; l < N; ++l) {
          const size_t n = node(kz, l);
          const double t1b = hb2 * T.lw0[l] * alph_b, t1t = ht2 * T.lw1[l] * alph_t;
          ua0[l] = bot ? 0.0 : t1b; ua7[l] = top ? 0.0 : t1t;
          rhs[l][0] = ifac * t_u[l] - q[l].u0 + qcur[V_MOMX][n]; rhs[l][1] = ifac * t_v[l] - q[l].v0 + qcur[V_MOMY][n];
          rhs[l][2] = top ? 0.0 : -t1t;
          if (!bot) { const double Luv = -t1b; ua0[l] -= Luv * uvprev[2]; rhs[l][0] -= Luv * uvprev[0]; rhs[l][1] -= Luv * uvprev[1]; }
        }
        const double a00 = 1.0 + ua0[0], a01 = ua7[0], a10 = ua0[7], a11 = 1.0 + ua7[7], rdet = 1.0 / (a00 * a11 - a01 * a10);
        for (int r = 0; r < 3; ++r) {
          const double x0 = (a11 * rhs[0][r] - a01 * rhs[7][r]) * rdet, x7 = (-a10 * rhs[0][r] + a00 * rhs[7][r]) * rdet;
          for (int l = 0; l < N; ++l) suv[(size_t(kz) * N + l) * 3 + r] = (l == 0) ? x0 : (l == 7) ? x7 : rhs[l][r] - ua0[l] * x0 - ua7[l] * x7;
        }
        for (int r = 0; r < 3; ++r) uvprev[r] = suv[(size_t(kz) * N + 7) * 3 + r];
      }
      prev = q[7];
      for (int l = 0; l < N; ++l) q[l] = qn[l];
    }
    if (ifac == 0.0) continue;
    // backward sweep (solve :429-444, solve_uv :661-674) and the tendency (rhot_hevi.F90:931-940)
    double nb[3] = {0, 0, 0}, nbu = 0, nbv = 0;
    for (int kz = NeZ - 1; kz >= 0; --kz) {
      double x0[3] = {0, 0, 0}, xu0 = 0, xv0 = 0;
      for (int l = 0; l < N; ++l) {
        const size_t n = node(kz, l);
        double d[3], du = suv[(size_t(kz) * N + l) * 3 + 0], dv = suv[(size_t(kz) * N + l) * 3 + 1];
        for (int a = 0; a < 3; ++a) {
          d[a] = sd[(size_t(kz) * N + l) * 3 + a];
          if (kz < NeZ - 1) for (int b = 0; b < 3; ++b) d[a] -= sG[((size_t(kz) * N + l) * 3 + a) * 3 + b] * nb[b];
        }
        if (kz < NeZ - 1) { const double g = suv[(size_t(kz) * N + l) * 3 + 2]; du -= g * nbu; dv -= g * nbv; }
        if (l == 0) { x0[0] = d[0]; x0[1] = d[1]; x0[2] = d[2]; xu0 = du; xv0 = dv; }
        kim[V_DDENS][n] = (q0[V_DDENS][n] + d[0] - qcur[V_DDENS][n]) / ifac;
        kim[V_MOMZ][n] = (q0[V_MOMZ][n] + d[1] - qcur[V_MOMZ][n]) / ifac;
        kim[V_DRHOT][n] = (q0[V_DRHOT][n] + d[2] - qcur[V_DRHOT][n]) / ifac;
        kim[V_MOMX][n] = (q0[V_MOMX][n] + du - qcur[V_MOMX][n]) / ifac;
        kim[V_MOMY][n] = (q0[V_MOMY][n] + dv - qcur[V_MOMY][n]) / ifac;
      }
      nb[0] = x0[0]; nb[1] = x0[1]; nb[2] = x0[2]; nbu = xu0; nbv = xv0;
    }
  }
  return 0;
}

}  // extern "C"
