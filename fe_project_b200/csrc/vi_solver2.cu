// HEVI vertical-implicit column solve for p = 7, second design (sm_100a, FP64): TWO lanes per column, block elimination.
//
// Rows a9-a12 of SURVEY.md 8: atm_dyn_dgm_nonhydro3d_rhot_hevi_cal_vi (scale_atm_dyn_dgm_nonhydro3d_rhot_hevi.F90:772-965), eval_Ax,
// vi_cal_del_flux_dyn, construct_matbnd, solve (..._rhot_hevi_common_2.F90:111-1328), solve_Nnode8_var3
// (scale_atm_dyn_dgm_hevi_common_linalgebra.F90:2296-2445).
//
// Why a second design.  The first kernel (vi_solver.cu) spreads the 16 x 20 augmented block of a column-element over 8 lanes, two rows
// each: every one of the 16 pivot steps publishes a row in shared memory and reads it back in all lanes, so one 8-byte operand
// crosses lanes per two multiply-adds and ncu showed the shared-memory / shuffle data pipe at 82 % with the FP64 pipe at 23 %
// (profiles/r01_v5_vi_column_details.txt).  The data pipe moves 128 B per clock and SM whatever the sharing pattern, so the only cure
// is fewer operands crossing lanes per multiply-add: here a column is owned by TWO lanes (rows 0..3 / 4..7 of every block), and the
// system is never held as a 24 x 24 or 16 x 16 block: the density is eliminated in closed form (its rows are I + dfac D + two lifted
// face terms), then theta (8 x 8, partial pivoting) and the Schur complement in w (8 x 8, partial pivoting): vi_block.cuh, whose row
// functions are validated against the oracle on the CPU (tests/test_vi_block_host.py).  A lane holds 4 x 20 and then 4 x 12 matrix
// entries; the pivot row travels by one shuffle per entry between the two lanes of the pair; the constant operator tables sit in
// constant memory, where an index known at compile time costs no load at all.
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "fedg_internal.h"
#include "vi_block.cuh"

namespace fedg {

namespace {
using namespace vib;

__constant__ Tables cVT;
Tables g_vt_loaded{};
bool g_vt_valid = false;

constexpr unsigned FULL = 0xffffffffu;
constexpr int V2_COLS = 32;                  // columns per block
constexpr int V2_THREADS = 2 * V2_COLS;

struct NodeQ {            // quantities of one node evaluated on var0 (the Newton linearisation point)
  double rho0, w0, th0, u0, v0, dens, rhot, pot, wt, dpd, dpres_vol, a, dpf;
};
constexpr int NQ = 13;

__device__ __noinline__ double vi_pow_slow(double x, double e) { return pow(x, e); }
__device__ __forceinline__ double vi_pow(double x, double e, int exact) {
  if (!exact && x > 0.25 && x < 2.0) return exp(e * log(x));
  return vi_pow_slow(x, e);
}
// DPRES of a state value (DRHOT2PRES, nonhydro3d_common.F90:467-474), out of line: used once per node in the backward sweep
__device__ __noinline__ double dpres_of(double R, double rP0, double rhot, double gm, double P00, double ph, int exact) {
  return P00 * vi_pow(R * rP0 * rhot, gm, exact) - ph;
}
// raw inputs of one node: var0 (5), DENS_hyd, RHOT_hyd (dry, as the solver recomputes it), PRES_hyd, and for moist runs Rtot, CPtot / CVtot
struct RawQ { double rho0, w0, th0, u0, v0, dh, rh, ph, R, gm; };
template <bool MOIST>
__device__ __forceinline__ RawQ load_raw(const VIParams& P, size_t n) {
  RawQ r;
  r.rho0 = P.q0[V_DDENS][n]; r.w0 = P.q0[V_MOMZ][n]; r.th0 = P.q0[V_DRHOT][n]; r.u0 = P.q0[V_MOMX][n]; r.v0 = P.q0[V_MOMY][n];
  r.dh = P.dens_hyd[n]; r.rh = P.rhot_hyd_vi[n]; r.ph = P.pres_hyd[n];
  r.R = MOIST ? P.rtot[n] : P.c.Rdry;
  r.gm = MOIST ? P.cptot[n] / P.cvtot[n] : P.c.CPovCV;
  return r;
}
// out of line: one copy of the pow / exp / log / sqrt / division sequences
__device__ __noinline__ NodeQ node_q(const RawQ r, double P00, double rP0, double gamm, int exact) {
  NodeQ q;
  q.rho0 = r.rho0; q.w0 = r.w0; q.th0 = r.th0; q.u0 = r.u0; q.v0 = r.v0;
  q.dens = r.dh + q.rho0;
  q.rhot = r.rh + q.th0;
  q.pot = q.rhot / q.dens;
  const double ptot = P00 * vi_pow(r.R * rP0 * q.rhot, r.gm, exact);
  q.dpres_vol = ptot - r.ph;
  q.wt = q.w0 / q.dens;
  q.dpd = r.gm * ptot / q.rhot;
  const double rdens0 = 1.0 / q.dens;
  q.a = fabs(q.w0 * rdens0) + sqrt(gamm * ptot * rdens0);
  q.dpf = P00 * vi_pow(r.R * rP0 * q.dens * q.pot, r.gm, exact) - r.ph;
  return q;
}
template <bool MOIST>
__device__ __forceinline__ NodeQ load_node(const VIParams& P, size_t n) {
  return node_q(load_raw<MOIST>(P, n), P.c.PRES00, P.c.rP0, P.c.gamm, P.exact_pow);
}
__device__ __forceinline__ void put_node(double* s, const NodeQ& q) {
  s[0] = q.rho0; s[1] = q.w0; s[2] = q.th0; s[3] = q.u0; s[4] = q.v0; s[5] = q.dens; s[6] = q.rhot; s[7] = q.pot; s[8] = q.wt; s[9] = q.dpd;
  s[10] = q.dpres_vol; s[11] = q.a; s[12] = q.dpf;
}
__device__ __forceinline__ NodeQ get_node(const double* s) {
  NodeQ q;
  q.rho0 = s[0]; q.w0 = s[1]; q.th0 = s[2]; q.u0 = s[3]; q.v0 = s[4]; q.dens = s[5]; q.rhot = s[6]; q.pot = s[7]; q.wt = s[8]; q.dpd = s[9];
  q.dpres_vol = s[10]; q.a = s[11]; q.dpf = s[12];
  return q;
}

// per-column shared memory (the two lanes of the column read and write it; they sit in one warp: __syncwarp orders the accesses)
struct ColSm {
  Coef C;                                   // 40
  double LF[4][NLF];                        // 52
  double vec[32];                           // pot, wt, s, dpd [8 each]; the solution w[8][4] lives here once dpd is dead
  double Rrho0[8];
  union {
    struct { double nq7[NQ]; double w0[8], pw[8], dpv[8], rho0[8]; } a;    // node-7 quantities + the vectors of the operator evaluation
    double X[8][12];                        // S_thth^-1 [S_thw | RHS_th]
  } u;
  double th0[4];                            // theta_0 of the four right-hand sides
  double prev[NQ];                          // top node of the element below
  double g[3][NR];                          // solution of the element below at its top node
  double uvp[3];                            // the same for the (MOMX, MOMY) system: du, dv, guv
  double nq0[2][NQ];                        // node 0 of the current / the next element (parity of kz)
  double pad[4];
};
static_assert(sizeof(ColSm) % 8 == 0, "doubles");
static_assert((sizeof(ColSm) / 4) % 32 == 4, "bank layout: column stride = 4 words mod 32");

__device__ __forceinline__ void pf_l2(const double* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void pair_sync() {
  __syncwarp();
  asm volatile("" ::: "memory");
}
// One row of the theta block / of the Schur complement, assembled OUT OF LINE into the lane's staging slot: one copy of the row code
// instead of four, and the row's temporaries cannot be interleaved with the rows the caller already holds in registers (inlined four
// times the kernel spilled 1.2 KB per thread).
__device__ __noinline__ void theta_row_staged(ColSm& sm, int h, int l, double Rth0) {
  const Tables& T = cVT;
  const Coef& C = sm.C;
  double Rth[NR], A[20];
  Rth[0] = Rth0;
#pragma unroll
  for (int b = 0; b < 3; ++b) Rth[1 + b] = T.lw1[l] * C.U[2][b];
  theta_row(C, T, l, sm.vec, sm.vec + 8, sm.vec + 16, sm.Rrho0, Rth, sm.LF, A);
#pragma unroll
  double* st = &sm.u.X[4][0] + 20 * h;      // staging: the upper half of the X area (free until the theta block is solved)
#pragma unroll
  for (int j = 0; j < 20; ++j) st[j] = A[j];
}
__device__ __noinline__ void schur_row_staged(ColSm& sm, int h, int l, double Rw0) {
  const Tables& T = cVT;
  const Coef& C = sm.C;
  double Rw[NR], H[12];
  Rw[0] = Rw0;
#pragma unroll
  for (int b = 0; b < 3; ++b) Rw[1 + b] = T.lw1[l] * C.U[1][b];
  schur_row(C, T, l, sm.vec + 24, sm.Rrho0, Rw, sm.LF, sm.u.X, H);
#pragma unroll
  double* st = sm.vec + 12 * h;             // staging: pot / wt / s are dead once the theta rows exist (dpd, vec[24..31], is not)
#pragma unroll
  for (int j = 0; j < 12; ++j) st[j] = H[j];
}
__device__ __forceinline__ double sel4(int c, double a0, double a1, double a2, double a3) {
  const double lo = (c & 1) ? a1 : a0, hi = (c & 1) ? a3 : a2;
  return (c & 2) ? hi : lo;
}

// Partial-pivot Gauss-Jordan on the leading 8 x 8 block of a system whose rows 4h .. 4h+3 sit on lane h of the pair (W columns).
// Pivot of column k = the largest magnitude among the rows not used yet, the lowest row index winning ties (the host harness and the
// reference's rule).  The pivot row is taken out of the owner's registers by selects and handed to the partner by one shuffle per
// entry.  The matrix is SHIFTED one column to the left with every step (the update writes A[s][j-1]), so that the pivot column is
// always column 0 and the eight steps are one rolled loop body: an eighth of the code of the unrolled form (the unrolled kernel spent
// 2.5 cycles per issued instruction waiting for the instruction cache).  On return row s of this lane solves unknown kk[s]:
// x_r = A[s][r] * rpiv[s], r = 0 .. W-9.
template <int W>
__device__ __forceinline__ void gauss_jordan_pair(double (&A)[4][W], int h, int (&kk)[4], double (&rpiv)[4]) {
  unsigned used = 0;
#pragma unroll
  for (int s = 0; s < 4; ++s) { kk[s] = 0; rpiv[s] = 1.0; }
#pragma unroll 1
  for (int k = 0; k < 8; ++k) {
    double best = -1.0;
    int cs = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const double v = fabs(A[s][0]);
      if (!((used >> s) & 1u) && v > best) { best = v; cs = s; }
    }
    const double ob = __shfl_xor_sync(FULL, best, 1);
    const bool mine = (best > ob) || (best == ob && h == 0);
    double pk;
    {
      const double c = sel4(cs, A[0][0], A[1][0], A[2][0], A[3][0]);
      const double o = __shfl_xor_sync(FULL, c, 1);
      pk = mine ? c : o;
    }
    const double rp = 1.0 / pk;
    double m[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const bool is_piv = mine && (cs == s);
      m[s] = is_piv ? 0.0 : A[s][0] * rp;
      if (is_piv) { used |= 1u << s; kk[s] = k; rpiv[s] = rp; }
    }
#pragma unroll
    for (int j = 1; j < W; ++j) {
      const double c = sel4(cs, A[0][j], A[1][j], A[2][j], A[3][j]);
      const double o = __shfl_xor_sync(FULL, c, 1);
      const double pj = mine ? c : o;
#pragma unroll
      for (int s = 0; s < 4; ++s) A[s][j - 1] = A[s][j] - m[s] * pj;
    }
  }
}

template <bool MOIST, bool IMPLICIT>
__global__ void __launch_bounds__(V2_THREADS, IMPLICIT ? 3 : 6) vi_column2_kernel(const __grid_constant__ VIParams P) {
  extern __shared__ __align__(16) unsigned char smraw[];
  ColSm& sm = reinterpret_cast<ColSm*>(smraw)[threadIdx.x >> 1];
  const Tables& T = cVT;
  const int h = threadIdx.x & 1;
  const int col = blockIdx.x * V2_COLS + (threadIdx.x >> 1);
  const int ncol = P.Ne2D * 64;
  const int ke2d = col >> 6, ij = col & 63;
  const int NeZ = P.NeZ, Ne2D = P.Ne2D;
  const double ifac = P.impl_fac;
  auto node = [&](int kz, int l) { return (size_t(ke2d) + size_t(kz) * Ne2D) * 512 + ij + 64 * l; };
  double* scr = P.scratch;   // var3: [kz][12][8][ncol], then uv: [kz][3][8][ncol]
  const size_t scr_uv = size_t(NeZ) * 96 * ncol;
  double* vpot = sm.vec, *vwt = sm.vec + 8, *vs = sm.vec + 16, *vdpd = sm.vec + 24;

  // prologue: node 0 of the bottom element (lane 0 evaluates it; afterwards every element hands the next one its node 0)
  if (h == 0) put_node(sm.nq0[0], load_node<MOIST>(P, node(0, 0)));
  pair_sync();

  // ---------------- forward sweep
  for (int kz = 0; kz < NeZ; ++kz) {
    const int ke = ke2d + kz * Ne2D;
    const bool bot = (kz == 0), top = (kz == NeZ - 1);
    const double* nq0 = sm.nq0[kz & 1];
    double* nq0n = sm.nq0[(kz + 1) & 1];
    // ---- node quantities: lane 0 evaluates nodes 1, 2, 3 and node 0 of the element above, lane 1 nodes 4 .. 7 (four each)
    // L2 prefetch of the inputs of the element above (one lane per 128-byte line: 16 columns share it)
    if (kz + 1 < NeZ && ((threadIdx.x >> 1) & 15) == 0) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const size_t n = node(kz + 1, 4 * h + a);
        pf_l2(P.q0[V_DDENS] + n); pf_l2(P.q0[V_MOMZ] + n); pf_l2(P.q0[V_DRHOT] + n); pf_l2(P.q0[V_MOMX] + n); pf_l2(P.q0[V_MOMY] + n);
        pf_l2(P.dens_hyd + n); pf_l2(P.rhot_hyd_vi + n); pf_l2(P.pres_hyd + n);
        if (MOIST) { pf_l2(P.rtot + n); pf_l2(P.cptot + n); pf_l2(P.cvtot + n); }
        if (IMPLICIT) { pf_l2(P.qcur[V_DDENS] + n); pf_l2(P.qcur[V_MOMZ] + n); pf_l2(P.qcur[V_DRHOT] + n); pf_l2(P.qcur[V_MOMX] + n); pf_l2(P.qcur[V_MOMY] + n); }
      }
    }
    double base_in[4][5];                    // var0 of the own rows (rho0, w0, th0, u0, v0)
    RawQ raw[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {            // all loads of the element first (the four L2 latencies overlap), then the arithmetic
      const bool nxt = (h == 0 && a == 0);
      raw[a] = load_raw<MOIST>(P, node(nxt ? min(kz + 1, NeZ - 1) : kz, 4 * h + a));
    }
    double qc[4][5];                         // state entering the stage at the own rows
    if (IMPLICIT) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const size_t n = node(kz, 4 * h + a);
        qc[a][0] = P.qcur[V_DDENS][n]; qc[a][1] = P.qcur[V_MOMZ][n]; qc[a][2] = P.qcur[V_DRHOT][n]; qc[a][3] = P.qcur[V_MOMX][n]; qc[a][4] = P.qcur[V_MOMY][n];
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      // slot 0 of lane 0: the evaluation goes to the hand-over slot of the element above, row 0 itself comes from this element's slot
      const bool nxt = (h == 0 && a == 0);
      const int l = 4 * h + a;
      NodeQ q = node_q(raw[a], P.c.PRES00, P.c.rP0, P.c.gamm, P.exact_pow);
      if (nxt) { put_node(nq0n, q); q = get_node(nq0); }
      vpot[l] = q.pot; vwt[l] = q.wt; vs[l] = q.pot * q.wt; vdpd[l] = q.dpd;
      sm.u.a.w0[l] = q.w0; sm.u.a.pw[l] = q.pot * q.w0; sm.u.a.dpv[l] = q.dpres_vol; sm.u.a.rho0[l] = q.rho0;
      if (l == 7) put_node(sm.u.a.nq7, q);
      base_in[a][0] = q.rho0; base_in[a][1] = q.w0; base_in[a][2] = q.th0; base_in[a][3] = q.u0; base_in[a][4] = q.v0;
    }
    pair_sync();

    // ---- face states and flux jumps (vi_cal_del_flux_dyn :1262-1322, _uv :1158-1161); nz = -1 at the bottom, +1 at the top
    const double E33 = P.escale[2 * size_t(P.Ne) + ke];
    const double Fs_b = P.fscale[4 * size_t(P.Ne) + ke], Fs_t = P.fscale[5 * size_t(P.Ne) + ke];
    const NodeQ M0 = get_node(nq0), M7 = get_node(sm.u.a.nq7);
    const NodeQ Pb = bot ? M0 : get_node(sm.prev), Pt = top ? M7 : get_node(nq0n);
    const double alph_b = bot ? M0.a : fmax(M0.a, Pb.a), alph_t = top ? M7.a : fmax(M7.a, Pt.a);
    const double wP_b = bot ? -M0.w0 : Pb.w0, wP_t = top ? -M7.w0 : Pt.w0;
    const double hb = 0.5 * Fs_b, ht = 0.5 * Fs_t;
    const double dl_r_b = hb * ((wP_b - M0.w0) * (-1.0) - alph_b * (Pb.rho0 - M0.rho0));
    const double dl_w_b = hb * ((Pb.dpf - M0.dpf) * (-1.0) - alph_b * (wP_b - M0.w0));
    const double dl_t_b = hb * ((Pb.pot * wP_b - M0.pot * M0.w0) * (-1.0) - alph_b * (Pb.th0 - M0.th0));
    const double dl_r_t = ht * ((wP_t - M7.w0) - alph_t * (Pt.rho0 - M7.rho0));
    const double dl_w_t = ht * ((Pt.dpf - M7.dpf) - alph_t * (wP_t - M7.w0));
    const double dl_t_t = ht * ((Pt.pot * wP_t - M7.pot * M7.w0) - alph_t * (Pt.th0 - M7.th0));
    const double dl_u_b = (-0.5 * Fs_b * alph_b) * (Pb.u0 - M0.u0), dl_v_b = (-0.5 * Fs_b * alph_b) * (Pb.v0 - M0.v0);
    const double dl_u_t = (-0.5 * Fs_t * alph_t) * (Pt.u0 - M7.u0), dl_v_t = (-0.5 * Fs_t * alph_t) * (Pt.v0 - M7.v0);

    // ---- vertical operator at var0 on the own rows (eval_Ax :224-262, eval_Ax_uv :546-553); GsqrtV = 1
    double t_r[4], t_w[4], t_t[4], t_u[4], t_v[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int l = 4 * h + a;
      const double* Dl = T.D + l * N;
      const double* Vl = T.VP + l * N;
      double dz_r = 0.0, dz_t = 0.0, dz_w = 0.0, drho = 0.0;
#pragma unroll
      for (int p = 0; p < N; ++p) {
        const double d = Dl[p];
        dz_r += d * sm.u.a.w0[p]; dz_t += d * sm.u.a.pw[p]; dz_w += d * sm.u.a.dpv[p]; drho += Vl[p] * sm.u.a.rho0[p];
      }
      const double l0 = T.lw0[l], l1 = T.lw1[l];
      t_r[a] = -(E33 * dz_r + (l0 * dl_r_b + l1 * dl_r_t));
      t_t[a] = -(E33 * dz_t + (l0 * dl_t_b + l1 * dl_t_t));
      t_w[a] = -(E33 * dz_w + (l0 * dl_w_b + l1 * dl_w_t)) - P.c.GRAV * drho;
      t_u[a] = -(l0 * dl_u_b + l1 * dl_u_t); t_v[a] = -(l0 * dl_v_b + l1 * dl_v_t);
    }

    if (!IMPLICIT) {   // explicit evaluation only (first IMEX stage): k_im = -A_v(q)
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const size_t n = node(kz, 4 * h + a);
        P.kim[V_DDENS][n] = t_r[a]; P.kim[V_MOMZ][n] = t_w[a]; P.kim[V_DRHOT][n] = t_t[a]; P.kim[V_MOMX][n] = t_u[a]; P.kim[V_MOMY][n] = t_v[a];
        // this kernel is also the StoreImplicit of the stage (impl_fac = 0): the stage state is the input state
        const double qr = P.qcur[V_DDENS][n], qw = P.qcur[V_MOMZ][n], qt = P.qcur[V_DRHOT][n], qu = P.qcur[V_MOMX][n], qv = P.qcur[V_MOMY][n];
        P.qout[V_DDENS][n] = qr; P.qout[V_MOMZ][n] = qw; P.qout[V_DRHOT][n] = qt; P.qout[V_MOMX][n] = qu; P.qout[V_MOMY][n] = qv;
        const double R = MOIST ? P.rtot[n] : P.c.Rdry;
        const double gm = MOIST ? P.cptot[n] / P.cvtot[n] : P.c.CPovCV;
        P.dpout[n] = dpres_of(R, P.c.rP0, P.therm_hyd[n] + qt, gm, P.c.PRES00, P.pres_hyd[n], P.exact_pow);
      }
      pair_sync();
      if (h == 1) put_node(sm.prev, M7);
      pair_sync();
      continue;
    }

    // ---- coefficients of the block: face matrices, block-Thomas terms of the element below, pivots of the density rows
    const double hb2 = 0.5 * ifac * Fs_b, ht2 = 0.5 * ifac * Fs_t;
    if (h == 0) {
      Coef C;
      C.dfac = E33 * ifac; C.gfac = ifac * P.c.GRAV;
      FaceNbr F;
      F.potn_b = Pb.pot; F.wtn_b = Pb.wt; F.dpdn_b = Pb.dpd;
      F.potn_t = Pt.pot; F.wtn_t = Pt.wt; F.dpdn_t = Pt.dpd;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int r = 0; r < NR; ++r) F.g[a][r] = sm.g[a][r];
      face_coef(C, bot, top, hb2, ht2, alph_b, alph_t, M0.pot, M0.wt, M0.dpd, M7.pot, M7.wt, M7.dpd, F);
      rho_pivots(C, T);
      sm.C = C;
    }
    // state entering the stage at the own rows, right-hand sides of rhs 0 (eval_Ax :306-317: impl_fac A_t - PROG_VARS + q00)
    double Rrho_own[4], Rw0[4], Rth0[4], bu[4], bv[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      Rrho_own[a] = ifac * t_r[a] - base_in[a][0] + qc[a][0];
      Rw0[a] = ifac * t_w[a] - base_in[a][1] + qc[a][1];
      Rth0[a] = ifac * t_t[a] - base_in[a][2] + qc[a][2];
      bu[a] = ifac * t_u[a] - base_in[a][3] + qc[a][3];
      bv[a] = ifac * t_v[a] - base_in[a][4] + qc[a][4];
    }
    // ---- (MOMX, MOMY): (I + ua0 e0^T + ua7 e7^T) x = [bu | bv | bg]  (construct_matbnd_uv :960-1003, solve_uv :640-674); the
    // solution goes to the scratch arrays at once
    {
      double ua0[4], ua7[4], bg[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int l = 4 * h + a;
        const double t1b = hb2 * T.lw0[l] * alph_b, t1t = ht2 * T.lw1[l] * alph_t;
        ua0[a] = bot ? 0.0 : t1b; ua7[a] = top ? 0.0 : t1t; bg[a] = top ? 0.0 : -t1t;
        if (!bot) { const double Luv = -t1b; ua0[a] -= Luv * sm.uvp[2]; bu[a] -= Luv * sm.uvp[0]; bv[a] -= Luv * sm.uvp[1]; }
      }
      // rows 0 (lane 0, slot 0) and 7 (lane 1, slot 3)
      const double e0 = (h == 0) ? ua0[0] : ua0[3], e7 = (h == 0) ? ua7[0] : ua7[3];
      const double eu = (h == 0) ? bu[0] : bu[3], ev = (h == 0) ? bv[0] : bv[3], eg = (h == 0) ? bg[0] : bg[3];
      const double o0 = __shfl_xor_sync(FULL, e0, 1), o7 = __shfl_xor_sync(FULL, e7, 1);
      const double ou = __shfl_xor_sync(FULL, eu, 1), ov = __shfl_xor_sync(FULL, ev, 1), og = __shfl_xor_sync(FULL, eg, 1);
      const double a00 = 1.0 + ((h == 0) ? e0 : o0), a01 = (h == 0) ? e7 : o7, a10 = (h == 0) ? o0 : e0, a11 = 1.0 + ((h == 0) ? o7 : e7);
      const double rdet = 1.0 / (a00 * a11 - a01 * a10);
      const double u0r = (h == 0) ? eu : ou, u7r = (h == 0) ? ou : eu, v0r = (h == 0) ? ev : ov, v7r = (h == 0) ? ov : ev;
      const double g0r = (h == 0) ? eg : og, g7r = (h == 0) ? og : eg;
      const double xu0 = (a11 * u0r - a01 * u7r) * rdet, xu7 = (-a10 * u0r + a00 * u7r) * rdet;
      const double xv0 = (a11 * v0r - a01 * v7r) * rdet, xv7 = (-a10 * v0r + a00 * v7r) * rdet;
      const double xg0 = (a11 * g0r - a01 * g7r) * rdet, xg7 = (-a10 * g0r + a00 * g7r) * rdet;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int l = 4 * h + a;
        const double xu = (l == 0) ? xu0 : (l == 7) ? xu7 : bu[a] - ua0[a] * xu0 - ua7[a] * xu7;
        const double xv = (l == 0) ? xv0 : (l == 7) ? xv7 : bv[a] - ua0[a] * xv0 - ua7[a] * xv7;
        const double xg = (l == 0) ? xg0 : (l == 7) ? xg7 : bg[a] - ua0[a] * xg0 - ua7[a] * xg7;
        scr[scr_uv + ((size_t(kz) * 3 + 0) * 8 + l) * ncol + col] = xu;
        scr[scr_uv + ((size_t(kz) * 3 + 1) * 8 + l) * ncol + col] = xv;
        scr[scr_uv + ((size_t(kz) * 3 + 2) * 8 + l) * ncol + col] = xg;
      }
      // every read of the old hand-over values precedes the shuffles above: the top node's solution can take their place
      if (h == 1) { sm.uvp[0] = xu7; sm.uvp[1] = xv7; sm.uvp[2] = xg7; }
    }
    pair_sync();
    if (h == 1) put_node(sm.prev, M7);       // every read of the old hand-over node is done (face states, face_coef)
    const Coef& C = sm.C;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int l = 4 * h + a;
      sm.Rrho0[l] = Rrho_own[a] + T.lw0[l] * C.rb[0];
      Rw0[a] += T.lw0[l] * C.rb[1]; Rth0[a] += T.lw0[l] * C.rb[2];
    }
    pair_sync();
    // ---- linear forms of the eliminated density: the 13 columns are split over the two lanes
    {
      const double r00 = sm.Rrho0[0], r07 = sm.Rrho0[7];
#pragma unroll
      for (int cc = 0; cc < 7; ++cc) {
        const int c = 2 * cc + h;
        if (c < NLF) {
          double o[4];
          rho_form_col(C, T, c, r00, r07, o);
          sm.LF[0][c] = o[0]; sm.LF[1][c] = o[1]; sm.LF[2][c] = o[2]; sm.LF[3][c] = o[3];
        }
      }
    }
    pair_sync();

    // ---- theta block: [S_thth | S_thw | RHS_th] rows of the own nodes, Gauss-Jordan, X to shared memory
    {
      double A[4][20];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        theta_row_staged(sm, h, 4 * h + a, Rth0[a]);
#pragma unroll
        for (int j = 0; j < 20; ++j) A[a][j] = (&sm.u.X[4][0] + 20 * h)[j];
      }
      pair_sync();                            // the operator-evaluation vectors (union with X) are dead on both lanes
      int kk[4];
      double rpiv[4];
      gauss_jordan_pair<20>(A, h, kk, rpiv);
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int c = 0; c < 12; ++c) sm.u.X[kk[s]][c] = A[s][c] * rpiv[s];
    }
    pair_sync();
    // ---- Schur complement in w: rows of the own nodes, Gauss-Jordan, solution to shared memory
    {
      double Hm[4][12];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        schur_row_staged(sm, h, 4 * h + a, Rw0[a]);
#pragma unroll
        for (int j = 0; j < 12; ++j) Hm[a][j] = (sm.vec + 12 * h)[j];
      }
      int kw[4];
      double rpiv[4];
      gauss_jordan_pair<12>(Hm, h, kw, rpiv);
      pair_sync();                            // every read of pot / wt / s / dpd is done: w takes their place
      double (*wsol)[NR] = reinterpret_cast<double (*)[NR]>(sm.vec);
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int r = 0; r < NR; ++r) wsol[kw[s]][r] = Hm[s][r] * rpiv[s];
    }
    pair_sync();
    const double (*wsol)[NR] = reinterpret_cast<const double (*)[NR]>(sm.vec);
    double th[4][NR], rho[4][NR], wl[4][NR];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int l = 4 * h + a;
      theta_solve(l, sm.u.X, wsol, th[a]);
#pragma unroll
      for (int r = 0; r < NR; ++r) wl[a][r] = wsol[l][r];
    }
    if (h == 0) {
#pragma unroll
      for (int r = 0; r < NR; ++r) sm.th0[r] = th[0][r];
    }
    pair_sync();
#pragma unroll
    for (int a = 0; a < 4; ++a) rho_solve(C, T, 4 * h + a, sm.Rrho0[4 * h + a], sm.LF, wsol, sm.th0, rho[a]);

    // ---- keep b and G of this element for the backward sweep; hand the top node to the element above
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int l = 4 * h + a;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        scr[((size_t(kz) * 12 + 0 * 4 + r) * 8 + l) * ncol + col] = rho[a][r];
        scr[((size_t(kz) * 12 + 1 * 4 + r) * 8 + l) * ncol + col] = wl[a][r];
        scr[((size_t(kz) * 12 + 2 * 4 + r) * 8 + l) * ncol + col] = th[a][r];
      }
    }
    pair_sync();
    if (h == 1) {
#pragma unroll
      for (int r = 0; r < NR; ++r) { sm.g[0][r] = rho[3][r]; sm.g[1][r] = wl[3][r]; sm.g[2][r] = th[3][r]; }
    }
    pair_sync();
  }
  if (!IMPLICIT) return;

  // ---------------- backward sweep, update, outputs (solve :429-444, solve_uv :661-674, rhot_hevi.F90:931-940)
  double nb_r = 0.0, nb_w = 0.0, nb_t = 0.0, nb_u = 0.0, nb_v = 0.0;   // solution at node 0 of the element above
  for (int kz = NeZ - 1; kz >= 0; --kz) {
    double d[4][3], du[4], dv[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int l = 4 * h + a;
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        double x = scr[((size_t(kz) * 12 + v * 4) * 8 + l) * ncol + col];
        if (kz < NeZ - 1) {
          const double g0 = scr[((size_t(kz) * 12 + v * 4 + 1) * 8 + l) * ncol + col];
          const double g1 = scr[((size_t(kz) * 12 + v * 4 + 2) * 8 + l) * ncol + col];
          const double g2 = scr[((size_t(kz) * 12 + v * 4 + 3) * 8 + l) * ncol + col];
          x = x - g0 * nb_r - g1 * nb_w - g2 * nb_t;
        }
        d[a][v] = x;
      }
      du[a] = scr[scr_uv + ((size_t(kz) * 3 + 0) * 8 + l) * ncol + col];
      dv[a] = scr[scr_uv + ((size_t(kz) * 3 + 1) * 8 + l) * ncol + col];
      if (kz < NeZ - 1) {
        const double guv = scr[scr_uv + ((size_t(kz) * 3 + 2) * 8 + l) * ncol + col];
        du[a] = du[a] - guv * nb_u; dv[a] = dv[a] - guv * nb_v;
      }
    }
    // node 0 lives on lane 0 (slot 0) of the pair
    const int src = (threadIdx.x & 31) & ~1;
    nb_r = __shfl_sync(FULL, d[0][0], src); nb_w = __shfl_sync(FULL, d[0][1], src); nb_t = __shfl_sync(FULL, d[0][2], src);
    nb_u = __shfl_sync(FULL, du[0], src); nb_v = __shfl_sync(FULL, dv[0], src);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const size_t n = node(kz, 4 * h + a);
      const double cr = P.qcur[V_DDENS][n], cw = P.qcur[V_MOMZ][n], ct = P.qcur[V_DRHOT][n], cu = P.qcur[V_MOMX][n], cv = P.qcur[V_MOMY][n];
      // PROG_VARS = var0 + delta;  tendency = (PROG_VARS - q) / impl_fac;  StoreImplicit: q += impl_fac * k
      const double pr = P.q0[V_DDENS][n] + d[a][0], pw = P.q0[V_MOMZ][n] + d[a][1], pth = P.q0[V_DRHOT][n] + d[a][2];
      const double pu = P.q0[V_MOMX][n] + du[a], pvv = P.q0[V_MOMY][n] + dv[a];
      const double kr = (pr - cr) / ifac, kw = (pw - cw) / ifac, kt = (pth - ct) / ifac, ku = (pu - cu) / ifac, kv = (pvv - cv) / ifac;
      P.kim[V_DDENS][n] = kr; P.kim[V_MOMZ][n] = kw; P.kim[V_DRHOT][n] = kt; P.kim[V_MOMX][n] = ku; P.kim[V_MOMY][n] = kv;
      const double qr = cr + ifac * kr, qw = cw + ifac * kw, qt = ct + ifac * kt, qu = cu + ifac * ku, qv = cv + ifac * kv;
      P.qout[V_DDENS][n] = qr; P.qout[V_MOMZ][n] = qw; P.qout[V_DRHOT][n] = qt; P.qout[V_MOMX][n] = qu; P.qout[V_MOMY][n] = qv;
      // DPRES of the updated state for the explicit part of this stage (DRHOT2PRES, nonhydro3d_common.F90:467-474)
      const double R = MOIST ? P.rtot[n] : P.c.Rdry;
      const double gm = MOIST ? P.cptot[n] / P.cvtot[n] : P.c.CPovCV;
      P.dpout[n] = dpres_of(R, P.c.rP0, P.therm_hyd[n] + qt, gm, P.c.PRES00, P.pres_hyd[n], P.exact_pow);
    }
  }
}

}  // namespace

// returns false when the configuration is outside this kernel (the caller falls back to the first design)
bool launch_vi2(const VIParams& p, const ElemTables& tab, bool moist, cudaStream_t s) {
  const int ncol = p.Ne2D * 64;
  if (ncol % V2_COLS != 0 || tab.np != 8) return false;
  Tables T;
  build_tables(tab.D, tab.VP, tab.Lw, T);
  if (!g_vt_valid || std::memcmp(&g_vt_loaded, &T, sizeof(Tables)) != 0) {
    cudaMemcpyToSymbolAsync(cVT, &T, sizeof(Tables), 0, cudaMemcpyHostToDevice, s);   // pageable source: staged before the call returns
    g_vt_loaded = T;
    g_vt_valid = true;
  }
  dim3 grid(ncol / V2_COLS), block(V2_THREADS);
  const size_t shmem = size_t(V2_COLS) * sizeof(ColSm);
#define FEDG_VI2_LAUNCH(M, I)                                                                                   \
  do {                                                                                                          \
    static bool attr_set = false;                                                                               \
    if (!attr_set) {                                                                                            \
      cudaFuncSetAttribute(vi_column2_kernel<M, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shmem));   \
      attr_set = true;                                                                                          \
    }                                                                                                           \
    vi_column2_kernel<M, I><<<grid, block, shmem, s>>>(p);                                                      \
  } while (0)
  const bool implicit = p.impl_fac != 0.0;
  if (implicit) { if (moist) FEDG_VI2_LAUNCH(true, true); else FEDG_VI2_LAUNCH(false, true); }
  else { if (moist) FEDG_VI2_LAUNCH(true, false); else FEDG_VI2_LAUNCH(false, false); }
#undef FEDG_VI2_LAUNCH
  return true;
}

}  // namespace fedg
