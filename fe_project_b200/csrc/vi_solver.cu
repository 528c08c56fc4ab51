// HEVI vertical-implicit column solve for p = 7 (sm_100a, FP64).
//
// One Newton iteration of   q* - q + impl_fac * A_v(q*) = 0   about var0 (rows a9-a12 of SURVEY.md 8):
//   atm_dyn_dgm_nonhydro3d_rhot_hevi_cal_vi                    scale_atm_dyn_dgm_nonhydro3d_rhot_hevi.F90:772-965
//   eval_Ax(_uv), vi_cal_del_flux_dyn(_uv), construct_matbnd(_uv), solve(_uv)
//                                                             ..._rhot_hevi_common_2.F90:111-1328
//   solve_Nnode8_uv / solve_Nnode8_var3                        scale_atm_dyn_dgm_hevi_common_linalgebra.F90:2142-2445
//
// Mapping: a column (ke2D, ij) of NeZ stacked elements is owned by a group of 8 lanes, lane l <-> vertical node l
// (four columns per warp, columns consecutive in ij so that every global access is a full 32-byte sector).  The
// group marches upwards through the column (block-Thomas forward sweep) and back down.  Per element the lane holds
// the three equations of its node (DDENS, MOMZ, DRHOT rows of the 24x24 block, 168 registers with the four
// right-hand sides [b | U]) and the 8-lane group eliminates them with partial-pivot Gauss-Jordan: the pivot row is
// broadcast by warp shuffles, no shared-memory matrix.  The reference factorises with partial-pivot LU and then
// substitutes four right-hand sides; Gauss-Jordan on the augmented block costs the same flops here, keeps all 24
// rows busy at every step and needs no triangular solves.  The (MOMX, MOMY) system is an 8x8 block with a scalar
// coupling and runs in the same sweep with one row per lane.
//
// Scope of this kernel: flat MeshCubeDom3D geometry (GsqrtV = 1, G13 = G23 = 0), the configuration the regional
// HEVI cases run on; terrain-following HEVI is rejected at fedg_dyn_init.
#include <cstdint>

#include "fedg_internal.h"

namespace fedg {

constexpr unsigned FULL = 0xffffffffu;

struct NodeQ {            // quantities of one node evaluated on var0 (the Newton linearisation point)
  double rho0, w0, th0, u0, v0;   // DDENS, MOMZ, DRHOT, MOMX, MOMY of var0
  double dens, rhot, pot, wt, dpd, dpres_vol, a;
};

template <bool MOIST>
__device__ __forceinline__ NodeQ node_q(const VIParams& P, size_t n) {
  NodeQ q;
  q.rho0 = P.q0[V_DDENS][n]; q.w0 = P.q0[V_MOMZ][n]; q.th0 = P.q0[V_DRHOT][n]; q.u0 = P.q0[V_MOMX][n]; q.v0 = P.q0[V_MOMY][n];
  const double R = MOIST ? P.rtot[n] : P.c.Rdry;
  const double gm = MOIST ? P.cptot[n] / P.cvtot[n] : P.c.CPovCV;
  q.dens = P.dens_hyd[n] + q.rho0;
  q.rhot = P.rhot_hyd_vi[n] + q.th0;
  q.pot = q.rhot / q.dens;
  const double ptot = P.c.PRES00 * pow(R * P.c.rP0 * q.rhot, gm);
  q.dpres_vol = ptot - P.pres_hyd[n];
  q.wt = q.w0 / q.dens;
  q.dpd = gm * ptot / q.rhot;
  const double rdens0 = 1.0 / q.dens;
  q.a = fabs(q.w0 * rdens0) + sqrt(P.c.gamm * ptot * rdens0);
  return q;
}

// face form of the pressure perturbation (vi_cal_del_flux_dyn :1262-1266: dens * pott instead of RHOT)
template <bool MOIST>
__device__ __forceinline__ double dpres_face(const VIParams& P, size_t n, const NodeQ& q) {
  const double R = MOIST ? P.rtot[n] : P.c.Rdry;
  const double gm = MOIST ? P.cptot[n] / P.cvtot[n] : P.c.CPovCV;
  return P.c.PRES00 * pow(R * P.c.rP0 * q.dens * q.pot, gm) - P.pres_hyd[n];
}

__device__ __forceinline__ double sel3(int s, double a, double b, double c) { return s == 0 ? a : (s == 1 ? b : c); }

// Partial-pivot Gauss-Jordan on [A | R] (24 x (24+4)), rows 3l..3l+2 on lane l of the 8-lane group.
// Pivot = first row with the strictly largest magnitude among the rows not used yet (the reference's rule,
// linalgebra.F90:2322-2331).  On return sol[k*4 + r] (shared memory of the group) holds unknown k of RHS r.
__device__ __forceinline__ void gauss_jordan_24(double (&A)[3][28], int l8, double* __restrict__ sol) {
  unsigned used = 0;
  int kk[3] = {0, 0, 0};
  double rpiv[3] = {1.0, 1.0, 1.0};
#pragma unroll
  for (int k = 0; k < 24; ++k) {
    double best = -1.0;
    int cand = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const double v = fabs(A[s][k]);
      if (!((used >> s) & 1u) && v > best) { best = v; cand = 3 * l8 + s; }
    }
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) {
      const double ob = __shfl_xor_sync(FULL, best, off, 8);
      const int oc = __shfl_xor_sync(FULL, cand, off, 8);
      if (ob > best || (ob == best && oc < cand)) { best = ob; cand = oc; }
    }
    const int pl = cand / 3, ps = cand - 3 * pl;
    const bool mine = (pl == l8);
    const double piv = __shfl_sync(FULL, sel3(ps, A[0][k], A[1][k], A[2][k]), pl, 8);
    const double rp = 1.0 / piv;
    double m[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const bool is_piv = mine && ps == s;
      m[s] = is_piv ? 0.0 : A[s][k] * rp;
      if (is_piv) { used |= 1u << s; kk[s] = k; rpiv[s] = rp; }
    }
#pragma unroll
    for (int j = k + 1; j < 28; ++j) {
      const double pj = __shfl_sync(FULL, sel3(ps, A[0][j], A[1][j], A[2][j]), pl, 8);
#pragma unroll
      for (int s = 0; s < 3; ++s) A[s][j] -= m[s] * pj;
    }
  }
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int r = 0; r < 4; ++r) sol[kk[s] * 4 + r] = A[s][24 + r] * rpiv[s];
}

// 8 x (8+3) system, one row per lane
__device__ __forceinline__ void gauss_jordan_8(double (&A)[11], int l8, double* __restrict__ sol) {
  bool used = false;
  int kk = 0;
  double rpiv = 1.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    double best = used ? -1.0 : fabs(A[k]);
    int cand = l8;
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) {
      const double ob = __shfl_xor_sync(FULL, best, off, 8);
      const int oc = __shfl_xor_sync(FULL, cand, off, 8);
      if (ob > best || (ob == best && oc < cand)) { best = ob; cand = oc; }
    }
    const bool mine = (cand == l8);
    const double piv = __shfl_sync(FULL, A[k], cand, 8);
    const double rp = 1.0 / piv;
    const double m = mine ? 0.0 : A[k] * rp;
    if (mine) { used = true; kk = k; rpiv = rp; }
#pragma unroll
    for (int j = k + 1; j < 11; ++j) {
      const double pj = __shfl_sync(FULL, A[j], cand, 8);
      A[j] -= m * pj;
    }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) sol[kk * 3 + r] = A[8 + r] * rpiv;
}

// sum_l M[row][l] * x_l with x_l held by lane l of the group, l ascending
__device__ __forceinline__ double group_matvec(const double* __restrict__ Mrow, double x) {
  double s = Mrow[0] * __shfl_sync(FULL, x, 0, 8);
#pragma unroll
  for (int l = 1; l < 8; ++l) s += Mrow[l] * __shfl_sync(FULL, x, l, 8);
  return s;
}

constexpr int VI_THREADS = 128;   // 16 column groups per block
constexpr int VI_SOL = 96 + 24;   // per group: sol[24][4] + sol_uv[8][3]

template <bool MOIST>
__global__ void __launch_bounds__(VI_THREADS) vi_column_kernel(const __grid_constant__ VIParams P) {
  const int tid = threadIdx.x, grp = tid >> 3, l8 = tid & 7;
  const int ncol = P.Ne2D * 64;
  const int col = blockIdx.x * (VI_THREADS / 8) + grp;     // grid is sized so that col < ncol (Ne2D*64 % 16 == 0)
  const int ke2d = col >> 6, ij = col & 63;
  const int NeZ = P.NeZ, Ne2D = P.Ne2D;
  const double ifac = P.impl_fac;

  extern __shared__ __align__(16) double smem[];
  double* sD = smem;          // D1D[pv][l]
  double* sVP = smem + 64;    // VPOrdM1
  double* sLw = smem + 128;   // lift1d[pv][side]
  double* sSol = smem + 144 + size_t(grp) * VI_SOL;
  double* sSolUV = sSol + 96;
  for (int m = tid; m < 144; m += VI_THREADS) smem[m] = m < 64 ? P.tab->D[m] : (m < 128 ? P.tab->VP[m - 64] : P.tab->Lw[m - 128]);
  __syncthreads();
  double Drow[8], VProw[8];
#pragma unroll
  for (int l = 0; l < 8; ++l) { Drow[l] = sD[l8 * 8 + l]; VProw[l] = sVP[l8 * 8 + l]; }
  const double lw0 = sLw[l8 * 2], lw1 = sLw[l8 * 2 + 1];

  auto node = [&](int kz) { return (size_t(ke2d) + size_t(kz) * Ne2D) * 512 + ij + 64 * l8; };
  double* scr = P.scratch;   // var3: [kz][12][8][ncol], then uv: [kz][3][8][ncol]
  const size_t scr_uv = size_t(NeZ) * 96 * ncol;

  NodeQ q = node_q<MOIST>(P, node(0));
  double a_top_prev = 0.0;                       // a() of the top node of the element below
  NodeQ qb_prev = q;                             // node 7 of the element below (valid from kz = 1)
  double dpf_prev = 0.0;                         // its face-form pressure perturbation

  // ---------------- forward sweep
  for (int kz = 0; kz < NeZ; ++kz) {
    const int ke = ke2d + kz * Ne2D;
    const size_t n = node(kz);
    const bool bot_bc = (kz == 0), top_bc = (kz == NeZ - 1);
    NodeQ qn = q;                                // next element (lookahead); self when at the top
    if (!top_bc) qn = node_q<MOIST>(P, node(kz + 1));
    // face-form pressure of the own end nodes and of the node above
    const double dpf_own = dpres_face<MOIST>(P, n, q);
    const double dpf_next = top_bc ? 0.0 : dpres_face<MOIST>(P, node(kz + 1), qn);

    const double E33 = P.escale[2 * size_t(P.Ne) + ke];
    const double Fs_b = P.fscale[4 * size_t(P.Ne) + ke], Fs_t = P.fscale[5 * size_t(P.Ne) + ke];
    // dissipation coefficient of the two faces (nz^2 = 1)
    const double a0 = __shfl_sync(FULL, q.a, 0, 8), a7 = __shfl_sync(FULL, q.a, 7, 8), an0 = __shfl_sync(FULL, qn.a, 0, 8);
    const double alph_b = bot_bc ? fmax(a0, a0) : fmax(a0, a_top_prev);
    const double alph_t = top_bc ? fmax(a7, a7) : fmax(a7, an0);

    // ---- exterior states of the two faces (interior = own node 0 / node 7)
    //   M side values broadcast from lanes 0 / 7, P side from the neighbour element or the slip-wall mirror
    const double rM_b = __shfl_sync(FULL, q.rho0, 0, 8), wM_b = __shfl_sync(FULL, q.w0, 0, 8), tM_b = __shfl_sync(FULL, q.th0, 0, 8);
    const double pM_b = __shfl_sync(FULL, q.pot, 0, 8), dM_b = __shfl_sync(FULL, dpf_own, 0, 8);
    const double uM_b = __shfl_sync(FULL, q.u0, 0, 8), vM_b = __shfl_sync(FULL, q.v0, 0, 8);
    const double rM_t = __shfl_sync(FULL, q.rho0, 7, 8), wM_t = __shfl_sync(FULL, q.w0, 7, 8), tM_t = __shfl_sync(FULL, q.th0, 7, 8);
    const double pM_t = __shfl_sync(FULL, q.pot, 7, 8), dM_t = __shfl_sync(FULL, dpf_own, 7, 8);
    const double uM_t = __shfl_sync(FULL, q.u0, 7, 8), vM_t = __shfl_sync(FULL, q.v0, 7, 8);
    double rP_b, wP_b, mwP_b, tP_b, pP_b, dP_b, uP_b, vP_b, rP_t, wP_t, mwP_t, tP_t, pP_t, dP_t, uP_t, vP_t;
    if (bot_bc) { rP_b = rM_b; wP_b = -wM_b; mwP_b = -wM_b; tP_b = tM_b; pP_b = pM_b; dP_b = dM_b; uP_b = uM_b; vP_b = vM_b; }
    else {
      rP_b = __shfl_sync(FULL, qb_prev.rho0, 7, 8); wP_b = __shfl_sync(FULL, qb_prev.w0, 7, 8); mwP_b = wP_b;
      tP_b = __shfl_sync(FULL, qb_prev.th0, 7, 8); pP_b = __shfl_sync(FULL, qb_prev.pot, 7, 8); dP_b = __shfl_sync(FULL, dpf_prev, 7, 8);
      uP_b = __shfl_sync(FULL, qb_prev.u0, 7, 8); vP_b = __shfl_sync(FULL, qb_prev.v0, 7, 8);
    }
    if (top_bc) { rP_t = rM_t; wP_t = -wM_t; mwP_t = -wM_t; tP_t = tM_t; pP_t = pM_t; dP_t = dM_t; uP_t = uM_t; vP_t = vM_t; }
    else {
      rP_t = __shfl_sync(FULL, qn.rho0, 0, 8); wP_t = __shfl_sync(FULL, qn.w0, 0, 8); mwP_t = wP_t;
      tP_t = __shfl_sync(FULL, qn.th0, 0, 8); pP_t = __shfl_sync(FULL, qn.pot, 0, 8); dP_t = __shfl_sync(FULL, dpf_next, 0, 8);
      uP_t = __shfl_sync(FULL, qn.u0, 0, 8); vP_t = __shfl_sync(FULL, qn.v0, 0, 8);
    }
    // flux jumps (vi_cal_del_flux_dyn :1306-1322, _uv :1158-1161); nz = -1 at the bottom face, +1 at the top face
    const double hb = 0.5 * Fs_b, ht = 0.5 * Fs_t;
    const double dl_r_b = hb * ((mwP_b - wM_b) * (-1.0) - alph_b * (rP_b - rM_b));
    const double dl_w_b = hb * ((dP_b - dM_b) * (-1.0) - alph_b * (wP_b - wM_b));
    const double dl_t_b = hb * ((pP_b * mwP_b - pM_b * wM_b) * (-1.0) - alph_b * (tP_b - tM_b));
    const double dl_r_t = ht * ((mwP_t - wM_t) * (1.0) - alph_t * (rP_t - rM_t));
    const double dl_w_t = ht * ((dP_t - dM_t) * (1.0) - alph_t * (wP_t - wM_t));
    const double dl_t_t = ht * ((pP_t * mwP_t - pM_t * wM_t) * (1.0) - alph_t * (tP_t - tM_t));
    const double dl_u_b = (-0.5 * Fs_b * alph_b) * (uP_b - uM_b), dl_v_b = (-0.5 * Fs_b * alph_b) * (vP_b - vM_b);
    const double dl_u_t = (-0.5 * Fs_t * alph_t) * (uP_t - uM_t), dl_v_t = (-0.5 * Fs_t * alph_t) * (vP_t - vM_t);

    // ---- vertical operator at var0 (eval_Ax :224-262, eval_Ax_uv :546-553); GsqrtV = 1
    const double dz_r = group_matvec(Drow, q.w0);
    const double dz_t = group_matvec(Drow, q.pot * q.w0);
    const double dz_w = group_matvec(Drow, q.dpres_vol);
    const double drho = group_matvec(VProw, q.rho0);
    const double t_r = -(E33 * dz_r + (lw0 * dl_r_b + lw1 * dl_r_t));
    const double t_t = -(E33 * dz_t + (lw0 * dl_t_b + lw1 * dl_t_t));
    const double t_w = -(E33 * dz_w + (lw0 * dl_w_b + lw1 * dl_w_t)) - P.c.GRAV * drho;
    const double t_u = -(lw0 * dl_u_b + lw1 * dl_u_t), t_v = -(lw0 * dl_v_b + lw1 * dl_v_t);

    if (ifac == 0.0) {   // explicit evaluation only (first IMEX stage): k_im = -A_v(q)
      P.kim[V_DDENS][n] = t_r; P.kim[V_MOMZ][n] = t_w; P.kim[V_DRHOT][n] = t_t; P.kim[V_MOMX][n] = t_u; P.kim[V_MOMY][n] = t_v;
    } else {
      const double cr = P.qcur[V_DDENS][n], cw = P.qcur[V_MOMZ][n], ct = P.qcur[V_DRHOT][n], cu = P.qcur[V_MOMX][n], cv = P.qcur[V_MOMY][n];
      // ---- Jacobian block of this element (construct_matbnd :750-772), rows of the own node
      double A[3][28];
      const double potl = q.pot, wtl = q.wt, dpdl = q.dpd;
#pragma unroll
      for (int p2 = 0; p2 < 8; ++p2) {
        const double fdz = E33 / 1.0 * (ifac * Drow[p2]);
        const double id = (p2 == l8) ? 1.0 : 0.0;
        const double pot2 = __shfl_sync(FULL, potl, p2, 8), wt2 = __shfl_sync(FULL, wtl, p2, 8), dpd2 = __shfl_sync(FULL, dpdl, p2, 8);
        A[0][3 * p2 + 0] = id;                         A[0][3 * p2 + 1] = fdz;        A[0][3 * p2 + 2] = 0.0;
        A[1][3 * p2 + 0] = ifac * P.c.GRAV * VProw[p2]; A[1][3 * p2 + 1] = id;         A[1][3 * p2 + 2] = fdz * dpd2;
        A[2][3 * p2 + 0] = -fdz * pot2 * wt2;          A[2][3 * p2 + 1] = fdz * pot2; A[2][3 * p2 + 2] = id + fdz * wt2;
      }
      double Lm[3][3], Um[3][3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) { Lm[a][b] = 0.0; Um[a][b] = 0.0; }
      // face couplings (construct_matbnd :774-871)
      const double pot0 = __shfl_sync(FULL, potl, 0, 8), wt0 = __shfl_sync(FULL, wtl, 0, 8), dpd0 = __shfl_sync(FULL, dpdl, 0, 8);
      const double pot7 = __shfl_sync(FULL, potl, 7, 8), wt7 = __shfl_sync(FULL, wtl, 7, 8), dpd7 = __shfl_sync(FULL, dpdl, 7, 8);
      const double facb = 0.5 * ifac / 1.0 * lw0 * Fs_b, fact = 0.5 * ifac / 1.0 * lw1 * Fs_t;
      const double t1b = facb * fmax(alph_b, alph_b), t2b = facb * (-1.0);
      const double t1t = fact * fmax(alph_t, alph_t), t2t = fact * (1.0);
      if (bot_bc) {
        A[2][0] += 2.0 * t2b * pot0 * wt0; A[0][1] -= 2.0 * t2b; A[1][1] += 2.0 * t1b; A[2][1] -= 2.0 * t2b * pot0; A[2][2] -= 2.0 * t2b * wt0;
      } else {
        A[0][0] += t1b; A[2][0] += t2b * pot0 * wt0; A[0][1] -= t2b; A[1][1] += t1b; A[2][1] -= t2b * pot0;
        A[1][2] -= t2b * dpd0; A[2][2] += t1b - t2b * wt0;
        const double potn = __shfl_sync(FULL, qb_prev.pot, 7, 8), wtn = __shfl_sync(FULL, qb_prev.wt, 7, 8), dpdn = __shfl_sync(FULL, qb_prev.dpd, 7, 8);
        Lm[0][0] = -t1b; Lm[1][0] = 0.0;         Lm[2][0] = -t2b * potn * wtn;
        Lm[0][1] = t2b;  Lm[1][1] = -t1b;        Lm[2][1] = t2b * potn;
        Lm[0][2] = 0.0;  Lm[1][2] = t2b * dpdn;  Lm[2][2] = -t1b + t2b * wtn;
      }
      if (top_bc) {
        A[2][21] += 2.0 * t2t * pot7 * wt7; A[0][22] -= 2.0 * t2t; A[1][22] += 2.0 * t1t; A[2][22] -= 2.0 * t2t * pot7; A[2][23] -= 2.0 * t2t * wt7;
      } else {
        A[0][21] += t1t; A[2][21] += t2t * pot7 * wt7; A[0][22] -= t2t; A[1][22] += t1t; A[2][22] -= t2t * pot7;
        A[1][23] -= t2t * dpd7; A[2][23] += t1t - t2t * wt7;
        const double potn = __shfl_sync(FULL, qn.pot, 0, 8), wtn = __shfl_sync(FULL, qn.wt, 0, 8), dpdn = __shfl_sync(FULL, qn.dpd, 0, 8);
        Um[0][0] = -t1t; Um[1][0] = 0.0;         Um[2][0] = -t2t * potn * wtn;
        Um[0][1] = t2t;  Um[1][1] = -t1t;        Um[2][1] = t2t * potn;
        Um[0][2] = 0.0;  Um[1][2] = t2t * dpdn;  Um[2][2] = -t1t + t2t * wtn;
      }
      // right-hand sides: b = impl_fac * A_t - PROG_VARS + q00 (eval_Ax :306-317), PROG_VARS = var0
      A[0][24] = ifac * t_r - q.rho0 + cr;
      A[1][24] = ifac * t_w - q.w0 + cw;
      A[2][24] = ifac * t_t - q.th0 + ct;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) A[a][25 + b] = Um[a][b];
      // (MOMX, MOMY) block (construct_matbnd_uv :960-1003): I + columns 0 / 7, scalar couplings
      double B[11];
#pragma unroll
      for (int p2 = 0; p2 < 8; ++p2) B[p2] = (p2 == l8) ? 1.0 : 0.0;
      double Luv = 0.0;
      B[10] = 0.0;
      if (!bot_bc) { B[0] += t1b; Luv = -t1b; }
      if (!top_bc) { B[7] += t1t; B[10] = -t1t; }
      B[8] = ifac * t_u - q.u0 + cu;
      B[9] = ifac * t_v - q.v0 + cv;

      // ---- eliminate the coupling to the element below with its G = D^-1 U and b (solve :385-416, solve_uv :640-655)
      if (!bot_bc) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            A[a][c] = A[a][c] - Lm[a][0] * sSol[21 * 4 + 1 + c] - Lm[a][1] * sSol[22 * 4 + 1 + c] - Lm[a][2] * sSol[23 * 4 + 1 + c];
          A[a][24] = A[a][24] - Lm[a][0] * sSol[21 * 4] - Lm[a][1] * sSol[22 * 4] - Lm[a][2] * sSol[23 * 4];
        }
        B[0] = B[0] - Luv * sSolUV[7 * 3 + 2];
        B[8] = B[8] - Luv * sSolUV[7 * 3 + 0];
        B[9] = B[9] - Luv * sSolUV[7 * 3 + 1];
      }
      __syncwarp();
      gauss_jordan_24(A, l8, sSol);
      gauss_jordan_8(B, l8, sSolUV);
      __syncwarp();
      // ---- keep b and G of this element for the backward sweep
#pragma unroll
      for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int r = 0; r < 4; ++r) scr[((size_t(kz) * 12 + v * 4 + r) * 8 + l8) * ncol + col] = sSol[(3 * l8 + v) * 4 + r];
#pragma unroll
      for (int r = 0; r < 3; ++r) scr[scr_uv + ((size_t(kz) * 3 + r) * 8 + l8) * ncol + col] = sSolUV[l8 * 3 + r];
    }
    // roll the lookahead
    a_top_prev = a7;
    qb_prev = q;
    dpf_prev = dpf_own;
    q = qn;
  }

  // ---------------- backward sweep, update, outputs
  double nb_r = 0.0, nb_w = 0.0, nb_t = 0.0, nb_u = 0.0, nb_v = 0.0;   // solution at node 0 of the element above
  for (int kz = NeZ - 1; kz >= 0; --kz) {
    const size_t n = node(kz);
    const double cr = P.qcur[V_DDENS][n], cw = P.qcur[V_MOMZ][n], ct = P.qcur[V_DRHOT][n], cu = P.qcur[V_MOMX][n], cv = P.qcur[V_MOMY][n];
    double qr = cr, qw = cw, qt = ct, qu = cu, qv = cv;
    if (ifac != 0.0) {
      double d[3], gq[3][3];
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        d[v] = scr[((size_t(kz) * 12 + v * 4) * 8 + l8) * ncol + col];
#pragma unroll
        for (int r = 0; r < 3; ++r) gq[v][r] = scr[((size_t(kz) * 12 + v * 4 + 1 + r) * 8 + l8) * ncol + col];
      }
      double du = scr[scr_uv + ((size_t(kz) * 3 + 0) * 8 + l8) * ncol + col], dv = scr[scr_uv + ((size_t(kz) * 3 + 1) * 8 + l8) * ncol + col];
      const double guv = scr[scr_uv + ((size_t(kz) * 3 + 2) * 8 + l8) * ncol + col];
      if (kz < NeZ - 1) {   // solve :429-444, solve_uv :661-674
#pragma unroll
        for (int v = 0; v < 3; ++v) d[v] = d[v] - gq[v][0] * nb_r - gq[v][1] * nb_w - gq[v][2] * nb_t;
        du = du - guv * nb_u;
        dv = dv - guv * nb_v;
      }
      nb_r = __shfl_sync(FULL, d[0], 0, 8); nb_w = __shfl_sync(FULL, d[1], 0, 8); nb_t = __shfl_sync(FULL, d[2], 0, 8);
      nb_u = __shfl_sync(FULL, du, 0, 8); nb_v = __shfl_sync(FULL, dv, 0, 8);
      // PROG_VARS = var0 + delta;  tendency = (PROG_VARS - q) / impl_fac  (rhot_hevi.F90:931-940); StoreImplicit: q += impl_fac * k
      const double pr = P.q0[V_DDENS][n] + d[0], pw = P.q0[V_MOMZ][n] + d[1], pth = P.q0[V_DRHOT][n] + d[2];
      const double pu = P.q0[V_MOMX][n] + du, pvv = P.q0[V_MOMY][n] + dv;
      const double kr = (pr - cr) / ifac, kw = (pw - cw) / ifac, kt = (pth - ct) / ifac, ku = (pu - cu) / ifac, kv = (pvv - cv) / ifac;
      P.kim[V_DDENS][n] = kr; P.kim[V_MOMZ][n] = kw; P.kim[V_DRHOT][n] = kt; P.kim[V_MOMX][n] = ku; P.kim[V_MOMY][n] = kv;
      qr = cr + ifac * kr; qw = cw + ifac * kw; qt = ct + ifac * kt; qu = cu + ifac * ku; qv = cv + ifac * kv;
    }
    P.qout[V_DDENS][n] = qr; P.qout[V_MOMZ][n] = qw; P.qout[V_DRHOT][n] = qt; P.qout[V_MOMX][n] = qu; P.qout[V_MOMY][n] = qv;
    {  // DPRES of the updated state for the explicit part of this stage (DRHOT2PRES, nonhydro3d_common.F90:467-474)
      const double R = MOIST ? P.rtot[n] : P.c.Rdry;
      const double gm = MOIST ? P.cptot[n] / P.cvtot[n] : P.c.CPovCV;
      P.dpout[n] = P.c.PRES00 * pow(R * P.c.rP0 * (P.therm_hyd[n] + qt), gm) - P.pres_hyd[n];
    }
  }
}

void launch_vi(const VIParams& p, bool moist, cudaStream_t s) {
  const int ncol = p.Ne2D * 64;
  const int groups = VI_THREADS / 8;
  dim3 grid(ncol / groups), block(VI_THREADS);
  const size_t shmem = (144 + size_t(groups) * VI_SOL) * sizeof(double);
  if (moist) vi_column_kernel<true><<<grid, block, shmem, s>>>(p);
  else vi_column_kernel<false><<<grid, block, shmem, s>>>(p);
}

// IMEX / general stage combination  q = base + sum_m coef[m] * k[m]  (rk_advance_general2D, scale_timeint_rk.F90:2201-2355,
// evaluated in the reference's accumulation order), for the five variables at once.
__global__ void lincomb_kernel(const __grid_constant__ LinCombParams L) {
  const size_t n = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n >= L.n) return;
#pragma unroll
  for (int v = 0; v < NVAR; ++v) {
    double r = L.base[v][n];
    for (int m = 0; m < L.nterm; ++m) r = r + L.coef[m] * L.k[m][v][n];
    L.out[v][n] = r;
  }
}
void launch_lincomb(const LinCombParams& L, cudaStream_t s) {
  const int block = 256;
  lincomb_kernel<<<unsigned((L.n + block - 1) / block), block, 0, s>>>(L);
}

}  // namespace fedg
