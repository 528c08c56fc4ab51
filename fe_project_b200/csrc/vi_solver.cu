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
// TERRAIN = false: flat MeshCubeDom3D geometry (GsqrtV = 1, G13 = G23 = 0), the configuration the regional and global HEVI cases
// run on.  TERRAIN = true (regional mesh with topography): GsqrtV = Gsqrt / GsqrtH scales every row of the vertical operator, the
// vertical mass flux is MOMZ + GsqrtV (G13 MOMX + G23 MOMY) with the horizontal momenta AFTER their own implicit solve
// (rhot_hevi.F90:895-925 calls solve_uv before eval_Ax), and the dissipation coefficient carries Gnn = 1/GsqrtV^2 + G13^2 + G23^2
// (vi_cal_del_flux_dyn_uv :1100-1130).  Because the (MOMX, MOMY) solution of a node is only known after the backward sweep, the
// terrain path runs the kernel twice: pass 0 solves the horizontal momenta alone and stores them, pass 1 is the full solve.
#include <cstdint>
#include <cstdlib>

#include "fedg_internal.h"

namespace fedg {

constexpr unsigned FULL = 0xffffffffu;

struct NodeQ {            // quantities of one node evaluated on var0 (the Newton linearisation point)
  double rho0, w0, th0, u0, v0;   // DDENS, MOMZ, DRHOT, MOMX, MOMY of var0
  double dens, rhot, pot, wt, dpd, dpres_vol, a;
  double dpf;                     // face form of the pressure perturbation (vi_cal_del_flux_dyn :1262-1266: dens * pott instead of RHOT)
  double rgv, mw;                 // TERRAIN: 1 / GsqrtV and the vertical mass flux MOMZ + GsqrtV (G13 MOMX' + G23 MOMY'); flat: 1 and MOMZ
};

// raw inputs of one node: var0 (5), DENS_hyd, RHOT_hyd (dry, as the solver recomputes it), PRES_hyd, and for moist runs
// Rtot, CPtot, CVtot
struct RawQ {
  double rho0, w0, th0, u0, v0, dh, rh, ph, R, cp, cv;
  double gs, g13, g23, up, vp;    // TERRAIN: Gsqrt, G13, G23 and the horizontal momenta after their implicit solve (pass 1) / of var0
};
template <bool MOIST, bool TERRAIN>
__device__ __forceinline__ RawQ raw_load(const VIParams& P, size_t n) {
  RawQ r;
  r.rho0 = P.q0[V_DDENS][n]; r.w0 = P.q0[V_MOMZ][n]; r.th0 = P.q0[V_DRHOT][n]; r.u0 = P.q0[V_MOMX][n]; r.v0 = P.q0[V_MOMY][n];
  r.dh = P.dens_hyd[n]; r.rh = P.rhot_hyd_vi[n]; r.ph = P.pres_hyd[n];
  r.R = r.cp = r.cv = 0.0;
  if (MOIST) { r.R = P.rtot[n]; r.cp = P.cptot[n]; r.cv = P.cvtot[n]; }
  r.gs = 1.0; r.g13 = r.g23 = 0.0; r.up = r.u0; r.vp = r.v0;
  if (TERRAIN) {
    r.gs = P.gsqrt[n]; r.g13 = P.g13[n]; r.g23 = P.g23[n];
    if (P.pvu) { r.up = P.pvu[n]; r.vp = P.pvv[n]; }
  }
  return r;
}
// x^e as exp(e log x) for x = Rtot RHOT / P00 in [0.25, 2] (see eos_pres_fast in stage_common.cuh: <= 3.5e-16 relative, 64 % of
// the instructions of pow()); pow() outside the range or with FEDG_EXACT_POW=1
__device__ __forceinline__ double vi_pow(double x, double e, int exact) {
  if (!exact && x > 0.25 && x < 2.0) return exp(e * log(x));
  return pow(x, e);
}

// rgh = 1 / GsqrtH of the column (TERRAIN only)
template <bool MOIST, bool TERRAIN>
__device__ __forceinline__ NodeQ node_q(const VIParams& P, const RawQ& r, double rgh) {
  NodeQ q;
  q.rho0 = r.rho0; q.w0 = r.w0; q.th0 = r.th0; q.u0 = r.u0; q.v0 = r.v0;
  const double R = MOIST ? r.R : P.c.Rdry;
  const double gm = MOIST ? r.cp / r.cv : P.c.CPovCV;
  q.dens = r.dh + q.rho0;
  q.rhot = r.rh + q.th0;
  q.pot = q.rhot / q.dens;
  const double ptot = P.c.PRES00 * vi_pow(R * P.c.rP0 * q.rhot, gm, P.exact_pow);
  q.dpres_vol = ptot - r.ph;
  q.dpd = gm * ptot / q.rhot;
  const double rdens0 = 1.0 / q.dens;
  if (TERRAIN) {
    const double gv = r.gs * rgh;      // GsqrtV = Gsqrt / GsqrtH (rhot_hevi.F90:864)
    q.rgv = 1.0 / gv;
    q.mw = q.w0 + gv * r.g13 * r.up + gv * r.g23 * r.vp;                 // eval_Ax :224-230, with the updated horizontal momenta
    q.wt = q.mw / q.dens;
    const double wt0 = (q.w0 * q.rgv + r.g13 * q.u0 + r.g23 * q.v0) * rdens0;     // vi_cal_del_flux_dyn_uv :1100-1130, on var0
    const double Gnn = q.rgv * q.rgv + r.g13 * r.g13 + r.g23 * r.g23;
    q.a = fabs(wt0) + sqrt(Gnn * P.c.gamm * ptot * rdens0);
  } else {
    q.rgv = 1.0; q.mw = q.w0;
    q.wt = q.w0 / q.dens;
    q.a = fabs(q.w0 * rdens0) + sqrt(P.c.gamm * ptot * rdens0);
  }
  q.dpf = P.c.PRES00 * vi_pow(R * P.c.rP0 * q.dens * q.pot, gm, P.exact_pow) - r.ph;
  return q;
}

// 8-byte asynchronous global -> shared copy (LDGSTS): the forward sweep prefetches the inputs of the next element while
// the current one is eliminated, without holding them in registers
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ double sel3(int s, double a, double b, double c) { return s == 0 ? a : (s == 1 ? b : c); }

// The row buffer is written by one lane and read by the others: plain C++ accesses let the compiler forward its own
// stores / reuse earlier loads across __syncwarp (observed: wrong results, and selects + moves instead of loads), so the
// cross-lane shared-memory traffic of the solver is explicit PTX.
__device__ __forceinline__ void sts_pair(double* p, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ double2 lds_pair(const double* p) {
  double2 r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))) : "memory");
  return r;
}
__device__ __forceinline__ double lds_one(const double* p) {
  double r;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(p))) : "memory");
  return r;
}

// Partial-pivot Gauss-Jordan on the reduced system [S | R] (16 x (16+4)): unknowns (MOMZ_0..7, DRHOT_0..7) of one
// column-element after DDENS has been eliminated (see the kernel), rows (MOMZ_l, DRHOT_l) on lane l of the 8-lane group.
// Pivot = the largest magnitude among the rows not used yet (linalgebra.F90:2322-2331 applies the same rule to the
// 24 x 24 block).  The lane that owns the pivot row publishes it in the group's shared-memory row buffer (double-buffered
// over k: one __syncwarp per pivot) and every lane reads it back with broadcast 128-bit loads.  On return
// sol[id*4 + r] holds unknown id = 3*node + {1: MOMZ, 2: DRHOT} of RHS r (the reference's interleaved numbering).
constexpr int PROW = 24;   // doubles per row buffer (20 used)
__device__ __forceinline__ void gauss_jordan_16(double (&A)[2][20], int l8, double* prow, double* sol) {
  unsigned used = 0;
  int kk[2] = {0, 0};
  double rpiv[2] = {1.0, 1.0};
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    double best = -1.0;
    int cand = 0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const double v = fabs(A[s][k]);
      if (!((used >> s) & 1u) && v > best) { best = v; cand = 2 * l8 + s; }
    }
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) {
      const double ob = __shfl_xor_sync(FULL, best, off, 8);
      const int oc = __shfl_xor_sync(FULL, cand, off, 8);
      if (ob > best || (ob == best && oc < cand)) { best = ob; cand = oc; }
    }
    const int pl = cand >> 1, ps = cand & 1;
    const bool mine = (pl == l8);
    double* buf = prow + (k & 1) * PROW;
    const int j0 = k & ~1;
#pragma unroll
    for (int s = 0; s < 2; ++s)
      if (mine && ps == s) {
#pragma unroll
        for (int j = j0; j < 20; j += 2) sts_pair(buf + j, A[s][j], A[s][j + 1]);
      }
    __syncwarp();
    const double rp = 1.0 / lds_one(buf + k);
    double m[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const bool is_piv = mine && ps == s;
      m[s] = is_piv ? 0.0 : A[s][k] * rp;
      if (is_piv) { used |= 1u << s; kk[s] = k; rpiv[s] = rp; }
    }
#pragma unroll
    for (int j = (k + 1) & ~1; j < 20; j += 2) {
      const double2 t = lds_pair(buf + j);
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (j > k) A[s][j] -= m[s] * t.x;
        A[s][j + 1] -= m[s] * t.y;
      }
    }
  }
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int id = kk[s] < 8 ? 3 * kk[s] + 1 : 3 * (kk[s] - 8) + 2;
    sts_pair(sol + id * 4, A[s][16] * rpiv[s], A[s][17] * rpiv[s]);
    sts_pair(sol + id * 4 + 2, A[s][18] * rpiv[s], A[s][19] * rpiv[s]);
  }
}

// x = (lane == piv) ? n : x - f * n   : one Gauss-Jordan row operation with a pivot row that lives on lane `piv`
#define VI_ROWOP(x, n, f, is_p) x = (is_p) ? (n) : ((x) - (f) * (n))

// sum_l M[row][l] * x_l with x_l held by lane l of the group, l ascending
__device__ __forceinline__ double group_matvec(const double* __restrict__ Mrow, double x) {
  double s = Mrow[0] * __shfl_sync(FULL, x, 0, 8);
#pragma unroll
  for (int l = 1; l < 8; ++l) s += Mrow[l] * __shfl_sync(FULL, x, l, 8);
  return s;
}

constexpr int VI_THREADS = 128;   // 16 column groups per block
constexpr int VI_PREV = 12;       // per group: values of the top node of the element below (lane 7 writes, all lanes read)
constexpr int VI_YROW = 14;       // eliminated DDENS row of a node: [W_0..7 | T0, R0 | R1, R2 | R3, pad]
// per group: sol[24][4] + sol_uv[8][3] + two pivot-row buffers + prev + Y[8][14]; 292 = 4 mod 16 keeps the four groups of a
// warp on distinct banks for the broadcast 128-bit loads
constexpr int VI_SOL = 96 + 24 + 2 * PROW + VI_PREV + 8 * VI_YROW;
static_assert(VI_SOL == 292, "bank layout");
constexpr int VI_NQ = 13;         // doubles of a NodeQ parked in shared memory across the elimination (+ 2 with TERRAIN)
constexpr int VI_PF = 16;         // prefetch slots per thread: raw inputs of the node two elements up (11) + qcur of the next element (5)
constexpr int VI_NQ_T = 15, VI_PF_T = 21;   // TERRAIN: rgv, mw parked; Gsqrt, G13, G23, MOMX', MOMY' prefetched

__device__ __forceinline__ void park(double* s, const NodeQ& q) {
  s[0 * VI_THREADS] = q.rho0; s[1 * VI_THREADS] = q.w0; s[2 * VI_THREADS] = q.th0; s[3 * VI_THREADS] = q.u0; s[4 * VI_THREADS] = q.v0;
  s[5 * VI_THREADS] = q.dens; s[6 * VI_THREADS] = q.rhot; s[7 * VI_THREADS] = q.pot; s[8 * VI_THREADS] = q.wt; s[9 * VI_THREADS] = q.dpd;
  s[10 * VI_THREADS] = q.dpres_vol; s[11 * VI_THREADS] = q.a; s[12 * VI_THREADS] = q.dpf;
}
__device__ __forceinline__ void park_t(double* s, const NodeQ& q) { s[13 * VI_THREADS] = q.rgv; s[14 * VI_THREADS] = q.mw; }
__device__ __forceinline__ NodeQ unpark(const double* s) {
  NodeQ q;
  q.rho0 = s[0 * VI_THREADS]; q.w0 = s[1 * VI_THREADS]; q.th0 = s[2 * VI_THREADS]; q.u0 = s[3 * VI_THREADS]; q.v0 = s[4 * VI_THREADS];
  q.dens = s[5 * VI_THREADS]; q.rhot = s[6 * VI_THREADS]; q.pot = s[7 * VI_THREADS]; q.wt = s[8 * VI_THREADS]; q.dpd = s[9 * VI_THREADS];
  q.dpres_vol = s[10 * VI_THREADS]; q.a = s[11 * VI_THREADS]; q.dpf = s[12 * VI_THREADS];
  q.rgv = 1.0; q.mw = q.w0;
  return q;
}

// sum_l M[row][l] * x_l, x_l held by lane l of the group, l ascending; the row comes from the transposed table sMT[l*8 + row]
__device__ __forceinline__ double group_matvec_s(const double* __restrict__ sMT, int l8, double x) {
  double s = sMT[l8] * __shfl_sync(FULL, x, 0, 8);
#pragma unroll
  for (int l = 1; l < 8; ++l) s += sMT[l * 8 + l8] * __shfl_sync(FULL, x, l, 8);
  return s;
}

// IMPLICIT = false is the explicit evaluation k_im = -A_v(q) of a stage with a_im(s,s) = 0 (first stage of the ARK schemes): no
// elimination code, a fraction of the registers, so it runs at the occupancy of a streaming kernel.
template <bool MOIST, bool IMPLICIT, int MINB, bool TERRAIN>
__global__ void __launch_bounds__(VI_THREADS, MINB) vi_column_kernel(const __grid_constant__ VIParams P) {
  const int tid = threadIdx.x, grp = tid >> 3, l8 = tid & 7;
  const int ncol = P.Ne2D * 64;
  const int col = blockIdx.x * (VI_THREADS / 8) + grp;     // grid is sized so that col < ncol (Ne2D*64 % 16 == 0)
  const int ke2d = col >> 6, ij = col & 63;
  const int NeZ = P.NeZ, Ne2D = P.Ne2D;
  const double ifac = P.impl_fac;

  extern __shared__ __align__(16) double smem[];
  double* sDT = smem;         // D1D transposed: sDT[l*8 + pv] = D1D[pv][l]
  double* sVPT = smem + 64;   // VPOrdM1 transposed
  double* sLw = smem + 128;   // lift1d[pv][side]
  double* sSol = smem + 144 + size_t(grp) * VI_SOL;
  double* sSolUV = sSol + 96;
  double* sRow = sSol + 120;
  double* sPrev = sRow + 2 * PROW;
  double* sY = sPrev + VI_PREV;
  double* sQ = smem + 144 + size_t(VI_THREADS / 8) * VI_SOL + tid;   // [VI_NQ][VI_THREADS]
  double* sPF = sQ + size_t(TERRAIN ? VI_NQ_T : VI_NQ) * VI_THREADS;   // [VI_PF][VI_THREADS]
  const bool pass0 = TERRAIN && IMPLICIT && P.pass0;                  // terrain pass 0: only the (MOMX, MOMY) solve, stored in pvu_out / pvv_out
  const double rgh = TERRAIN ? 1.0 / P.gsqrtH[size_t(ke2d) * 64 + ij] : 1.0;
  for (int m = tid; m < 144; m += VI_THREADS) {
    if (m < 128) { const int r = (m & 63) >> 3, c = m & 7; smem[(m & 64) + c * 8 + r] = m < 64 ? P.tab->D[m] : P.tab->VP[m - 64]; }
    else smem[m] = P.tab->Lw[m - 128];
  }
  __syncthreads();
  const double lw0 = sLw[l8 * 2], lw1 = sLw[l8 * 2 + 1];

  auto node = [&](int kz) { return (size_t(ke2d) + size_t(kz) * Ne2D) * 512 + ij + 64 * l8; };
  double* scr = P.scratch;   // var3: [kz][12][8][ncol], then uv: [kz][3][8][ncol]
  const size_t scr_uv = size_t(NeZ) * 96 * ncol;

  // prefetch (asynchronous copies into this thread's slots of sPF): raw inputs of node(kr), qcur of node(kc)
  auto prefetch = [&](int kr, int kc) {
    if (kr < NeZ) {
      const size_t m = node(kr);
      cp_async8(sPF + 0 * VI_THREADS, P.q0[V_DDENS] + m); cp_async8(sPF + 1 * VI_THREADS, P.q0[V_MOMZ] + m);
      cp_async8(sPF + 2 * VI_THREADS, P.q0[V_DRHOT] + m); cp_async8(sPF + 3 * VI_THREADS, P.q0[V_MOMX] + m);
      cp_async8(sPF + 4 * VI_THREADS, P.q0[V_MOMY] + m); cp_async8(sPF + 5 * VI_THREADS, P.dens_hyd + m);
      cp_async8(sPF + 6 * VI_THREADS, P.rhot_hyd_vi + m); cp_async8(sPF + 7 * VI_THREADS, P.pres_hyd + m);
      if (MOIST) { cp_async8(sPF + 8 * VI_THREADS, P.rtot + m); cp_async8(sPF + 9 * VI_THREADS, P.cptot + m); cp_async8(sPF + 10 * VI_THREADS, P.cvtot + m); }
      if (TERRAIN) {
        cp_async8(sPF + 16 * VI_THREADS, P.gsqrt + m); cp_async8(sPF + 17 * VI_THREADS, P.g13 + m); cp_async8(sPF + 18 * VI_THREADS, P.g23 + m);
        cp_async8(sPF + 19 * VI_THREADS, (P.pvu ? P.pvu : P.q0[V_MOMX]) + m); cp_async8(sPF + 20 * VI_THREADS, (P.pvv ? P.pvv : P.q0[V_MOMY]) + m);
      }
    }
    if (IMPLICIT && kc < NeZ) {
      const size_t m = node(kc);
      cp_async8(sPF + 11 * VI_THREADS, P.qcur[V_DDENS] + m); cp_async8(sPF + 12 * VI_THREADS, P.qcur[V_MOMZ] + m);
      cp_async8(sPF + 13 * VI_THREADS, P.qcur[V_DRHOT] + m); cp_async8(sPF + 14 * VI_THREADS, P.qcur[V_MOMX] + m);
      cp_async8(sPF + 15 * VI_THREADS, P.qcur[V_MOMY] + m);
    }
    cp_async_commit();
  };
  prefetch(1, 0);
  NodeQ q = node_q<MOIST, TERRAIN>(P, raw_load<MOIST, TERRAIN>(P, node(0)), rgh);

  // ---------------- forward sweep
  for (int kz = 0; kz < NeZ; ++kz) {
    const int ke = ke2d + kz * Ne2D;
    const size_t n = node(kz);
    const bool bot_bc = (kz == 0), top_bc = (kz == NeZ - 1);
    // inputs of this element arrived while the previous one was eliminated
    cp_async_wait_all();
    NodeQ qn = q;                                // next element (lookahead); self when at the top
    if (!top_bc) {
      RawQ r;
      r.rho0 = sPF[0 * VI_THREADS]; r.w0 = sPF[1 * VI_THREADS]; r.th0 = sPF[2 * VI_THREADS]; r.u0 = sPF[3 * VI_THREADS];
      r.v0 = sPF[4 * VI_THREADS]; r.dh = sPF[5 * VI_THREADS]; r.rh = sPF[6 * VI_THREADS]; r.ph = sPF[7 * VI_THREADS];
      r.R = r.cp = r.cv = 0.0;
      if (MOIST) { r.R = sPF[8 * VI_THREADS]; r.cp = sPF[9 * VI_THREADS]; r.cv = sPF[10 * VI_THREADS]; }
      r.gs = 1.0; r.g13 = r.g23 = 0.0; r.up = r.u0; r.vp = r.v0;
      if (TERRAIN) { r.gs = sPF[16 * VI_THREADS]; r.g13 = sPF[17 * VI_THREADS]; r.g23 = sPF[18 * VI_THREADS]; r.up = sPF[19 * VI_THREADS]; r.vp = sPF[20 * VI_THREADS]; }
      qn = node_q<MOIST, TERRAIN>(P, r, rgh);
    }
    double cr = 0.0, cw_ = 0.0, ct = 0.0, cu = 0.0, cv = 0.0;   // state entering the stage at the own node
    if (IMPLICIT) { cr = sPF[11 * VI_THREADS]; cw_ = sPF[12 * VI_THREADS]; ct = sPF[13 * VI_THREADS]; cu = sPF[14 * VI_THREADS]; cv = sPF[15 * VI_THREADS]; }
    prefetch(kz + 2, kz + 1);
    // face-form pressure of the own end nodes and of the node above
    const double dpf_own = q.dpf;
    const double dpf_next = top_bc ? 0.0 : qn.dpf;

    const double E33 = P.escale[2 * size_t(P.Ne) + ke];
    const double Fs_b = P.fscale[4 * size_t(P.Ne) + ke], Fs_t = P.fscale[5 * size_t(P.Ne) + ke];
    // dissipation coefficient of the two faces (nz^2 = 1)
    const double a0 = __shfl_sync(FULL, q.a, 0, 8), a7 = __shfl_sync(FULL, q.a, 7, 8), an0 = __shfl_sync(FULL, qn.a, 0, 8);
    const double alph_b = bot_bc ? fmax(a0, a0) : fmax(a0, sPrev[9]);
    const double alph_t = top_bc ? fmax(a7, a7) : fmax(a7, an0);

    // ---- exterior states of the two faces (interior = own node 0 / node 7)
    //   M side values broadcast from lanes 0 / 7, P side from the neighbour element or the slip-wall mirror
    const double rM_b = __shfl_sync(FULL, q.rho0, 0, 8), wM_b = __shfl_sync(FULL, q.w0, 0, 8), tM_b = __shfl_sync(FULL, q.th0, 0, 8);
    const double pM_b = __shfl_sync(FULL, q.pot, 0, 8), dM_b = __shfl_sync(FULL, dpf_own, 0, 8);
    const double uM_b = __shfl_sync(FULL, q.u0, 0, 8), vM_b = __shfl_sync(FULL, q.v0, 0, 8);
    const double rM_t = __shfl_sync(FULL, q.rho0, 7, 8), wM_t = __shfl_sync(FULL, q.w0, 7, 8), tM_t = __shfl_sync(FULL, q.th0, 7, 8);
    const double pM_t = __shfl_sync(FULL, q.pot, 7, 8), dM_t = __shfl_sync(FULL, dpf_own, 7, 8);
    const double uM_t = __shfl_sync(FULL, q.u0, 7, 8), vM_t = __shfl_sync(FULL, q.v0, 7, 8);
    // vertical mass flux of the own face nodes (TERRAIN: MOMZ + GsqrtV (G13 MOMX + G23 MOMY); flat: MOMZ itself)
    const double mwM_b = TERRAIN ? __shfl_sync(FULL, q.mw, 0, 8) : wM_b, mwM_t = TERRAIN ? __shfl_sync(FULL, q.mw, 7, 8) : wM_t;
    double rP_b, wP_b, mwP_b, tP_b, pP_b, dP_b, uP_b, vP_b, rP_t, wP_t, mwP_t, tP_t, pP_t, dP_t, uP_t, vP_t;
    double potn_b = 0.0, wtn_b = 0.0, dpdn_b = 0.0;   // Jacobian factors of the node below the bottom face
    // slip wall (vi_cal_del_flux_dyn :1268-1274): MOMZ_P = -MOMZ_M - 2 GsqrtV (G13 MOMX + G23 MOMY) = MOMZ_M - 2 MW_M, MW_P = -MW_M
    if (bot_bc) { rP_b = rM_b; wP_b = TERRAIN ? wM_b - 2.0 * mwM_b : -wM_b; mwP_b = -mwM_b; tP_b = tM_b; pP_b = pM_b; dP_b = dM_b; uP_b = uM_b; vP_b = vM_b; }
    else {
      rP_b = sPrev[0]; wP_b = sPrev[1]; mwP_b = TERRAIN ? sPrev[10] : wP_b; tP_b = sPrev[2]; pP_b = sPrev[3]; dP_b = sPrev[8];
      uP_b = sPrev[4]; vP_b = sPrev[5]; potn_b = sPrev[3]; wtn_b = sPrev[6]; dpdn_b = sPrev[7];
    }
    if (top_bc) { rP_t = rM_t; wP_t = TERRAIN ? wM_t - 2.0 * mwM_t : -wM_t; mwP_t = -mwM_t; tP_t = tM_t; pP_t = pM_t; dP_t = dM_t; uP_t = uM_t; vP_t = vM_t; }
    else {
      rP_t = __shfl_sync(FULL, qn.rho0, 0, 8); wP_t = __shfl_sync(FULL, qn.w0, 0, 8); mwP_t = TERRAIN ? __shfl_sync(FULL, qn.mw, 0, 8) : wP_t;
      tP_t = __shfl_sync(FULL, qn.th0, 0, 8); pP_t = __shfl_sync(FULL, qn.pot, 0, 8); dP_t = __shfl_sync(FULL, dpf_next, 0, 8);
      uP_t = __shfl_sync(FULL, qn.u0, 0, 8); vP_t = __shfl_sync(FULL, qn.v0, 0, 8);
    }
    // flux jumps (vi_cal_del_flux_dyn :1306-1322, _uv :1158-1161); nz = -1 at the bottom face, +1 at the top face
    const double hb = 0.5 * Fs_b, ht = 0.5 * Fs_t;
    const double dl_r_b = hb * ((mwP_b - mwM_b) * (-1.0) - alph_b * (rP_b - rM_b));
    const double dl_w_b = hb * ((dP_b - dM_b) * (-1.0) - alph_b * (wP_b - wM_b));
    const double dl_t_b = hb * ((pP_b * mwP_b - pM_b * mwM_b) * (-1.0) - alph_b * (tP_b - tM_b));
    const double dl_r_t = ht * ((mwP_t - mwM_t) * (1.0) - alph_t * (rP_t - rM_t));
    const double dl_w_t = ht * ((dP_t - dM_t) * (1.0) - alph_t * (wP_t - wM_t));
    const double dl_t_t = ht * ((pP_t * mwP_t - pM_t * mwM_t) * (1.0) - alph_t * (tP_t - tM_t));
    const double dl_u_b = (-0.5 * Fs_b * alph_b) * (uP_b - uM_b), dl_v_b = (-0.5 * Fs_b * alph_b) * (vP_b - vM_b);
    const double dl_u_t = (-0.5 * Fs_t * alph_t) * (uP_t - uM_t), dl_v_t = (-0.5 * Fs_t * alph_t) * (vP_t - vM_t);

    // ---- vertical operator at var0 (eval_Ax :224-262, eval_Ax_uv :546-553); GsqrtV = 1
    const double dz_r = group_matvec_s(sDT, l8, TERRAIN ? q.mw : q.w0);
    const double dz_t = group_matvec_s(sDT, l8, q.pot * (TERRAIN ? q.mw : q.w0));
    const double dz_w = group_matvec_s(sDT, l8, q.dpres_vol);
    const double drho = group_matvec_s(sVPT, l8, q.rho0);
    double t_r = -(E33 * dz_r + (lw0 * dl_r_b + lw1 * dl_r_t));
    double t_t = -(E33 * dz_t + (lw0 * dl_t_b + lw1 * dl_t_t));
    double t_w = -(E33 * dz_w + (lw0 * dl_w_b + lw1 * dl_w_t));
    double t_u = -(lw0 * dl_u_b + lw1 * dl_u_t), t_v = -(lw0 * dl_v_b + lw1 * dl_v_t);
    if (TERRAIN) { t_r *= q.rgv; t_t *= q.rgv; t_w *= q.rgv; t_u *= q.rgv; t_v *= q.rgv; }   // eval_Ax :246-262, eval_Ax_uv :546-553: / GsqrtV
    t_w = t_w - P.c.GRAV * drho;

    if (!IMPLICIT) {   // explicit evaluation only (first IMEX stage): k_im = -A_v(q)
      P.kim[V_DDENS][n] = t_r; P.kim[V_MOMZ][n] = t_w; P.kim[V_DRHOT][n] = t_t; P.kim[V_MOMX][n] = t_u; P.kim[V_MOMY][n] = t_v;
      __syncwarp();
      if (l8 == 7) { sPrev[0] = q.rho0; sPrev[1] = q.w0; sPrev[2] = q.th0; sPrev[3] = q.pot; sPrev[4] = q.u0; sPrev[5] = q.v0;
                     sPrev[6] = q.wt; sPrev[7] = q.dpd; sPrev[8] = dpf_own; sPrev[9] = q.a; if (TERRAIN) sPrev[10] = q.mw; }
      __syncwarp();
      q = qn;
    } else {
      // ---- Jacobian block of this element (construct_matbnd :750-772), rows of the own node.  Unknowns are numbered
      // 3*node + {0: DDENS, 1: MOMZ, 2: DRHOT} in the reference.  The DDENS rows are the identity except in the columns
      // of the two face nodes, so DDENS is eliminated first with two static pivots (rows 0 and 7: the diagonal carries
      // the positive lift weight of the own face) and only the 16 x 16 system in (MOMZ, DRHOT) goes through the
      // partial-pivot elimination: 2944 instead of 8928 multiply-adds per column-element, 40 instead of 84 matrix
      // values per lane.  The solution is the reference's (same linear system) up to round-off.
      //   DDENS row l :  rho_l + ra0 rho_0 + ra7 rho_7 + sum_j rW[j] w_j + rT0 theta_0 = rR[0..3]
      //   A2[0] = MOMZ row, A2[1] = DRHOT row over the columns [w_0..7 | theta_0..7 | 4 RHS]; cw / cth = their DDENS columns
      double ra0 = 0.0, ra7 = 0.0, rT0 = 0.0, rW[8], rR[4], A2[2][20];
      double cw0 = 0.0, cth0 = 0.0, cth7 = 0.0;   // face / coupling corrections of the DDENS columns of node 0 and node 7
      const double potl = q.pot, wtl = q.wt, dpdl = q.dpd;
      // TERRAIN: every row of the block is divided by GsqrtV of its node (construct_matbnd :750-772, :795)
      const double gfac = ifac * P.c.GRAV, dfac = TERRAIN ? E33 * q.rgv * ifac : E33 / 1.0 * ifac;
#pragma unroll
      for (int p2 = 0; p2 < 8; ++p2) {
        const double fdz = dfac * sDT[p2 * 8 + l8];
        const double id = (p2 == l8) ? 1.0 : 0.0;
        const double pot2 = __shfl_sync(FULL, potl, p2, 8), wt2 = __shfl_sync(FULL, wtl, p2, 8), dpd2 = __shfl_sync(FULL, dpdl, p2, 8);
        rW[p2] = fdz;
        A2[0][p2] = id;         A2[0][8 + p2] = fdz * dpd2;     // DDENS columns: gfac * VP[l][p2] and -fdz * pot2 * wt2,
        A2[1][p2] = fdz * pot2; A2[1][8 + p2] = id + fdz * wt2; // rebuilt where they are used (product loop below)
      }
      double Lm[3][3], Um[3][3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) { Lm[a][b] = 0.0; Um[a][b] = 0.0; }
      // face couplings (construct_matbnd :774-871)
      const double pot0 = __shfl_sync(FULL, potl, 0, 8), wt0 = __shfl_sync(FULL, wtl, 0, 8), dpd0 = __shfl_sync(FULL, dpdl, 0, 8);
      const double pot7 = __shfl_sync(FULL, potl, 7, 8), wt7 = __shfl_sync(FULL, wtl, 7, 8), dpd7 = __shfl_sync(FULL, dpdl, 7, 8);
      const double facb = TERRAIN ? 0.5 * ifac * q.rgv * lw0 * Fs_b : 0.5 * ifac / 1.0 * lw0 * Fs_b;
      const double fact = TERRAIN ? 0.5 * ifac * q.rgv * lw1 * Fs_t : 0.5 * ifac / 1.0 * lw1 * Fs_t;
      const double t1b = facb * fmax(alph_b, alph_b), t2b = facb * (-1.0);
      const double t1t = fact * fmax(alph_t, alph_t), t2t = fact * (1.0);
      if (bot_bc) {
        cth0 += 2.0 * t2b * pot0 * wt0; rW[0] -= 2.0 * t2b; A2[0][0] += 2.0 * t1b; A2[1][0] -= 2.0 * t2b * pot0; A2[1][8] -= 2.0 * t2b * wt0;
      } else {
        ra0 += t1b; cth0 += t2b * pot0 * wt0; rW[0] -= t2b; A2[0][0] += t1b; A2[1][0] -= t2b * pot0;
        A2[0][8] -= t2b * dpd0; A2[1][8] += t1b - t2b * wt0;
        const double potn = potn_b, wtn = wtn_b, dpdn = dpdn_b;
        Lm[0][0] = -t1b; Lm[1][0] = 0.0;         Lm[2][0] = -t2b * potn * wtn;
        Lm[0][1] = t2b;  Lm[1][1] = -t1b;        Lm[2][1] = t2b * potn;
        Lm[0][2] = 0.0;  Lm[1][2] = t2b * dpdn;  Lm[2][2] = -t1b + t2b * wtn;
      }
      if (top_bc) {
        cth7 += 2.0 * t2t * pot7 * wt7; rW[7] -= 2.0 * t2t; A2[0][7] += 2.0 * t1t; A2[1][7] -= 2.0 * t2t * pot7; A2[1][15] -= 2.0 * t2t * wt7;
      } else {
        ra7 += t1t; cth7 += t2t * pot7 * wt7; rW[7] -= t2t; A2[0][7] += t1t; A2[1][7] -= t2t * pot7;
        A2[0][15] -= t2t * dpd7; A2[1][15] += t1t - t2t * wt7;
        const double potn = __shfl_sync(FULL, qn.pot, 0, 8), wtn = __shfl_sync(FULL, qn.wt, 0, 8), dpdn = __shfl_sync(FULL, qn.dpd, 0, 8);
        Um[0][0] = -t1t; Um[1][0] = 0.0;         Um[2][0] = -t2t * potn * wtn;
        Um[0][1] = t2t;  Um[1][1] = -t1t;        Um[2][1] = t2t * potn;
        Um[0][2] = 0.0;  Um[1][2] = t2t * dpdn;  Um[2][2] = -t1t + t2t * wtn;
      }
      // right-hand sides: b = impl_fac * A_t - PROG_VARS + q00 (eval_Ax :306-317), PROG_VARS = var0;  then the three
      // columns of U (coupling to the element above)
      rR[0] = ifac * t_r - q.rho0 + cr;
      A2[0][16] = ifac * t_w - q.w0 + cw_;
      A2[1][16] = ifac * t_t - q.th0 + ct;
#pragma unroll
      for (int b = 0; b < 3; ++b) { rR[1 + b] = Um[0][b]; A2[0][17 + b] = Um[1][b]; A2[1][17 + b] = Um[2][b]; }
      // right-hand sides and matrix of the (MOMX, MOMY) system (construct_matbnd_uv :960-1003): I + columns 0 / 7
      double bu = ifac * t_u - q.u0 + cu, bv = ifac * t_v - q.v0 + cv;
      double ua0 = bot_bc ? 0.0 : t1b, ua7 = top_bc ? 0.0 : t1t, bg = top_bc ? 0.0 : -t1t;

      // ---- eliminate the coupling to the element below with its G = D^-1 U and b (solve :385-416, solve_uv :640-655)
      if (!bot_bc) {
        const double2 g21a = lds_pair(sSol + 21 * 4), g21b = lds_pair(sSol + 21 * 4 + 2);   // (b, G col 0), (G col 1, G col 2)
        const double2 g22a = lds_pair(sSol + 22 * 4), g22b = lds_pair(sSol + 22 * 4 + 2);
        const double2 g23a = lds_pair(sSol + 23 * 4), g23b = lds_pair(sSol + 23 * 4 + 2);
        ra0 = ra0 - Lm[0][0] * g21a.y - Lm[0][1] * g22a.y - Lm[0][2] * g23a.y;
        rW[0] = rW[0] - Lm[0][0] * g21b.x - Lm[0][1] * g22b.x - Lm[0][2] * g23b.x;
        rT0 = rT0 - Lm[0][0] * g21b.y - Lm[0][1] * g22b.y - Lm[0][2] * g23b.y;
        rR[0] = rR[0] - Lm[0][0] * g21a.x - Lm[0][1] * g22a.x - Lm[0][2] * g23a.x;
        cw0 = cw0 - Lm[1][0] * g21a.y - Lm[1][1] * g22a.y - Lm[1][2] * g23a.y;
        A2[0][0] = A2[0][0] - Lm[1][0] * g21b.x - Lm[1][1] * g22b.x - Lm[1][2] * g23b.x;
        A2[0][8] = A2[0][8] - Lm[1][0] * g21b.y - Lm[1][1] * g22b.y - Lm[1][2] * g23b.y;
        A2[0][16] = A2[0][16] - Lm[1][0] * g21a.x - Lm[1][1] * g22a.x - Lm[1][2] * g23a.x;
        cth0 = cth0 - Lm[2][0] * g21a.y - Lm[2][1] * g22a.y - Lm[2][2] * g23a.y;
        A2[1][0] = A2[1][0] - Lm[2][0] * g21b.x - Lm[2][1] * g22b.x - Lm[2][2] * g23b.x;
        A2[1][8] = A2[1][8] - Lm[2][0] * g21b.y - Lm[2][1] * g22b.y - Lm[2][2] * g23b.y;
        A2[1][16] = A2[1][16] - Lm[2][0] * g21a.x - Lm[2][1] * g22a.x - Lm[2][2] * g23a.x;
        const double Luv = -t1b;
        const double guv = lds_one(sSolUV + 7 * 3 + 2);
        ua0 = ua0 - Luv * guv;
        bu = bu - Luv * lds_one(sSolUV + 7 * 3 + 0);
        bv = bv - Luv * lds_one(sSolUV + 7 * 3 + 1);
      }
      // park what the next element needs: the lookahead node and the top node of this element
      park(sQ, qn);
      if (TERRAIN) park_t(sQ, qn);
      __syncwarp();
      if (l8 == 7) { sPrev[0] = q.rho0; sPrev[1] = q.w0; sPrev[2] = q.th0; sPrev[3] = q.pot; sPrev[4] = q.u0; sPrev[5] = q.v0;
                     sPrev[6] = q.wt; sPrev[7] = q.dpd; sPrev[8] = dpf_own; sPrev[9] = q.a; if (TERRAIN) sPrev[10] = q.mw; }

      // ---- (MOMX, MOMY): (I + ua0 e0^T + ua7 e7^T) x = [bu | bv | bg], two static pivots
      {
        const bool is0 = (l8 == 0), is7 = (l8 == 7);
        const double rp0 = 1.0 / (1.0 + __shfl_sync(FULL, ua0, 0, 8));
        const double n7 = __shfl_sync(FULL, ua7, 0, 8) * rp0, nu = __shfl_sync(FULL, bu, 0, 8) * rp0,
                     nv = __shfl_sync(FULL, bv, 0, 8) * rp0, ng = __shfl_sync(FULL, bg, 0, 8) * rp0;
        VI_ROWOP(ua7, n7, ua0, is0); VI_ROWOP(bu, nu, ua0, is0); VI_ROWOP(bv, nv, ua0, is0); VI_ROWOP(bg, ng, ua0, is0);
        const double rp7 = 1.0 / (1.0 + __shfl_sync(FULL, ua7, 7, 8));
        const double mu = __shfl_sync(FULL, bu, 7, 8) * rp7, mv = __shfl_sync(FULL, bv, 7, 8) * rp7, mg = __shfl_sync(FULL, bg, 7, 8) * rp7;
        VI_ROWOP(bu, mu, ua7, is7); VI_ROWOP(bv, mv, ua7, is7); VI_ROWOP(bg, mg, ua7, is7);
      }
      // ---- DDENS rows: pivot on (row 0, rho_0), then (row 7, rho_7)   (terrain pass 0 skips the three-variable system)
      if (!pass0) {
      {
        const bool is0 = (l8 == 0), is7 = (l8 == 7);
        const double rp0 = 1.0 / (1.0 + __shfl_sync(FULL, ra0, 0, 8));
        {
          const double n = __shfl_sync(FULL, ra7, 0, 8) * rp0;
          VI_ROWOP(ra7, n, ra0, is0);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { const double n = __shfl_sync(FULL, rW[j], 0, 8) * rp0; VI_ROWOP(rW[j], n, ra0, is0); }
        { const double n = __shfl_sync(FULL, rT0, 0, 8) * rp0; VI_ROWOP(rT0, n, ra0, is0); }
#pragma unroll
        for (int r = 0; r < 4; ++r) { const double n = __shfl_sync(FULL, rR[r], 0, 8) * rp0; VI_ROWOP(rR[r], n, ra0, is0); }
        const double rp7 = 1.0 / (1.0 + __shfl_sync(FULL, ra7, 7, 8));
#pragma unroll
        for (int j = 0; j < 8; ++j) { const double n = __shfl_sync(FULL, rW[j], 7, 8) * rp7; VI_ROWOP(rW[j], n, ra7, is7); }
        { const double n = __shfl_sync(FULL, rT0, 7, 8) * rp7; VI_ROWOP(rT0, n, ra7, is7); }
#pragma unroll
        for (int r = 0; r < 4; ++r) { const double n = __shfl_sync(FULL, rR[r], 7, 8) * rp7; VI_ROWOP(rR[r], n, ra7, is7); }
      }
      // publish the eliminated DDENS row of the own node, then remove the DDENS columns from the MOMZ / DRHOT rows
      {
        double* y = sY + l8 * VI_YROW;
#pragma unroll
        for (int j = 0; j < 8; j += 2) sts_pair(y + j, rW[j], rW[j + 1]);
        sts_pair(y + 8, rT0, rR[0]); sts_pair(y + 10, rR[1], rR[2]); sts_pair(y + 12, rR[3], 0.0);
      }
      __syncwarp();
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const double* y = sY + m * VI_YROW;
        // DDENS columns of the MOMZ / DRHOT rows (construct_matbnd :757, :763) + the corrections of the face nodes
        const double potwt = __shfl_sync(FULL, potl * wtl, m, 8);
        double c0 = gfac * sVPT[m * 8 + l8], c1 = -(dfac * sDT[m * 8 + l8]) * potwt;
        if (m == 0) { c0 += cw0; c1 += cth0; }
        if (m == 7) c1 += cth7;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const double2 t = lds_pair(y + j);
          A2[0][j] -= c0 * t.x; A2[0][j + 1] -= c0 * t.y;
          A2[1][j] -= c1 * t.x; A2[1][j + 1] -= c1 * t.y;
        }
        const double2 t8 = lds_pair(y + 8), t10 = lds_pair(y + 10);
        const double t12 = lds_one(y + 12);
        A2[0][8] -= c0 * t8.x;  A2[0][16] -= c0 * t8.y;  A2[0][17] -= c0 * t10.x; A2[0][18] -= c0 * t10.y; A2[0][19] -= c0 * t12;
        A2[1][8] -= c1 * t8.x;  A2[1][16] -= c1 * t8.y;  A2[1][17] -= c1 * t10.x; A2[1][18] -= c1 * t10.y; A2[1][19] -= c1 * t12;
      }
      gauss_jordan_16(A2, l8, sRow, sSol);
      __syncwarp();
      // back-substitution of DDENS:  rho_l = R - sum_j W_j w_j - T0 theta_0   (own row re-read from shared memory)
      {
        const double* y = sY + l8 * VI_YROW;
        const double2 t8 = lds_pair(y + 8), t10 = lds_pair(y + 10);
        double x0 = t8.y, x1 = t10.x, x2 = t10.y, x3 = lds_one(y + 12);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const double wj = lds_one(y + j);
          const double2 sa = lds_pair(sSol + (3 * j + 1) * 4), sb = lds_pair(sSol + (3 * j + 1) * 4 + 2);
          x0 -= wj * sa.x; x1 -= wj * sa.y; x2 -= wj * sb.x; x3 -= wj * sb.y;
        }
        const double2 sa = lds_pair(sSol + 2 * 4), sb = lds_pair(sSol + 2 * 4 + 2);
        x0 -= t8.x * sa.x; x1 -= t8.x * sa.y; x2 -= t8.x * sb.x; x3 -= t8.x * sb.y;
        sts_pair(sSol + (3 * l8) * 4, x0, x1); sts_pair(sSol + (3 * l8) * 4 + 2, x2, x3);
      }
      }   // !pass0
      sSolUV[l8 * 3 + 0] = bu; sSolUV[l8 * 3 + 1] = bv; sSolUV[l8 * 3 + 2] = bg;
      __syncwarp();
      // ---- keep b and G of this element for the backward sweep
      if (!pass0)
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const double2 sa = lds_pair(sSol + (3 * l8 + v) * 4), sb = lds_pair(sSol + (3 * l8 + v) * 4 + 2);
        scr[((size_t(kz) * 12 + v * 4 + 0) * 8 + l8) * ncol + col] = sa.x;
        scr[((size_t(kz) * 12 + v * 4 + 1) * 8 + l8) * ncol + col] = sa.y;
        scr[((size_t(kz) * 12 + v * 4 + 2) * 8 + l8) * ncol + col] = sb.x;
        scr[((size_t(kz) * 12 + v * 4 + 3) * 8 + l8) * ncol + col] = sb.y;
      }
      scr[scr_uv + ((size_t(kz) * 3 + 0) * 8 + l8) * ncol + col] = bu;
      scr[scr_uv + ((size_t(kz) * 3 + 1) * 8 + l8) * ncol + col] = bv;
      scr[scr_uv + ((size_t(kz) * 3 + 2) * 8 + l8) * ncol + col] = bg;
      q = unpark(sQ);
      if (TERRAIN) { q.rgv = sQ[13 * VI_THREADS]; q.mw = sQ[14 * VI_THREADS]; }
    }
  }

  // ---------------- backward sweep, update, outputs.  The loads of element kz-1 are issued before element kz is processed
  // (register double buffer): the sweep is a chain of short dependent steps and was bound by global-load latency.
  struct BwdIn {
    double c[5], q0[5], d[3], gq[3][3], du, dv, guv, th, ph, R, gm;
  };
  auto bwd_load = [&](int kz) {
    BwdIn b;
    const size_t n = node(kz);
    b.c[0] = P.qcur[V_DDENS][n]; b.c[1] = P.qcur[V_MOMZ][n]; b.c[2] = P.qcur[V_DRHOT][n]; b.c[3] = P.qcur[V_MOMX][n]; b.c[4] = P.qcur[V_MOMY][n];
    b.th = P.therm_hyd[n]; b.ph = P.pres_hyd[n];
    b.R = MOIST ? P.rtot[n] : P.c.Rdry;
    b.gm = MOIST ? P.cptot[n] / P.cvtot[n] : P.c.CPovCV;
    if (IMPLICIT) {
      b.q0[0] = P.q0[V_DDENS][n]; b.q0[1] = P.q0[V_MOMZ][n]; b.q0[2] = P.q0[V_DRHOT][n]; b.q0[3] = P.q0[V_MOMX][n]; b.q0[4] = P.q0[V_MOMY][n];
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        b.d[v] = scr[((size_t(kz) * 12 + v * 4) * 8 + l8) * ncol + col];
#pragma unroll
        for (int r = 0; r < 3; ++r) b.gq[v][r] = scr[((size_t(kz) * 12 + v * 4 + 1 + r) * 8 + l8) * ncol + col];
      }
      b.du = scr[scr_uv + ((size_t(kz) * 3 + 0) * 8 + l8) * ncol + col];
      b.dv = scr[scr_uv + ((size_t(kz) * 3 + 1) * 8 + l8) * ncol + col];
      b.guv = scr[scr_uv + ((size_t(kz) * 3 + 2) * 8 + l8) * ncol + col];
    }
    return b;
  };
  double nb_r = 0.0, nb_w = 0.0, nb_t = 0.0, nb_u = 0.0, nb_v = 0.0;   // solution at node 0 of the element above
  BwdIn nxt = bwd_load(NeZ - 1);
  for (int kz = NeZ - 1; kz >= 0; --kz) {
    const size_t n = node(kz);
    const BwdIn in = nxt;
    if (kz > 0) nxt = bwd_load(kz - 1);
    const double cr = in.c[0], cw = in.c[1], ct = in.c[2], cu = in.c[3], cv = in.c[4];
    double qr = cr, qw = cw, qt = ct, qu = cu, qv = cv;
    if (IMPLICIT) {
      double d[3] = {in.d[0], in.d[1], in.d[2]};
      double du = in.du, dv = in.dv;
      if (kz < NeZ - 1) {   // solve :429-444, solve_uv :661-674
#pragma unroll
        for (int v = 0; v < 3; ++v) d[v] = d[v] - in.gq[v][0] * nb_r - in.gq[v][1] * nb_w - in.gq[v][2] * nb_t;
        du = du - in.guv * nb_u;
        dv = dv - in.guv * nb_v;
      }
      nb_r = __shfl_sync(FULL, d[0], 0, 8); nb_w = __shfl_sync(FULL, d[1], 0, 8); nb_t = __shfl_sync(FULL, d[2], 0, 8);
      nb_u = __shfl_sync(FULL, du, 0, 8); nb_v = __shfl_sync(FULL, dv, 0, 8);
      // PROG_VARS = var0 + delta;  tendency = (PROG_VARS - q) / impl_fac  (rhot_hevi.F90:931-940); StoreImplicit: q += impl_fac * k
      const double pr = in.q0[0] + d[0], pw = in.q0[1] + d[1], pth = in.q0[2] + d[2];
      const double pu = in.q0[3] + du, pvv = in.q0[4] + dv;
      if (pass0) { P.pvu_out[n] = pu; P.pvv_out[n] = pvv; continue; }   // terrain pass 0: MOMX', MOMY' for the mass flux of pass 1
      const double kr = (pr - cr) / ifac, kw = (pw - cw) / ifac, kt = (pth - ct) / ifac, ku = (pu - cu) / ifac, kv = (pvv - cv) / ifac;
      P.kim[V_DDENS][n] = kr; P.kim[V_MOMZ][n] = kw; P.kim[V_DRHOT][n] = kt; P.kim[V_MOMX][n] = ku; P.kim[V_MOMY][n] = kv;
      qr = cr + ifac * kr; qw = cw + ifac * kw; qt = ct + ifac * kt; qu = cu + ifac * ku; qv = cv + ifac * kv;
    }
    P.qout[V_DDENS][n] = qr; P.qout[V_MOMZ][n] = qw; P.qout[V_DRHOT][n] = qt; P.qout[V_MOMX][n] = qu; P.qout[V_MOMY][n] = qv;
    // DPRES of the updated state for the explicit part of this stage (DRHOT2PRES, nonhydro3d_common.F90:467-474)
    P.dpout[n] = P.c.PRES00 * vi_pow(in.R * P.c.rP0 * (in.th + qt), in.gm, P.exact_pow) - in.ph;
  }
}

void launch_vi(const VIParams& p, bool moist, cudaStream_t s) {
  if (p.gsqrt) {   // terrain-following mesh: the eight-lane kernel with the metric terms; implicit launches run pass 0 + pass 1
    const int ncol = p.Ne2D * 64, groups = VI_THREADS / 8;
    dim3 grid(ncol / groups), block(VI_THREADS);
    const size_t shmem = (144 + size_t(groups) * VI_SOL + size_t(VI_NQ_T + VI_PF_T) * VI_THREADS) * sizeof(double);
#define FEDG_VIT_LAUNCH(M, I, Q)                                                                                    \
  do {                                                                                                              \
    static bool attr_set = false;                                                                                   \
    if (!attr_set) {                                                                                                \
      cudaFuncSetAttribute(vi_column_kernel<M, I, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shmem)); \
      attr_set = true;                                                                                              \
    }                                                                                                               \
    vi_column_kernel<M, I, 2, true><<<grid, block, shmem, s>>>(Q);                                                  \
  } while (0)
    if (p.impl_fac == 0.0) {
      VIParams q = p; q.pvu = q.pvv = nullptr; q.pass0 = 0;
      if (moist) FEDG_VIT_LAUNCH(true, false, q); else FEDG_VIT_LAUNCH(false, false, q);
    } else {
      VIParams q0 = p; q0.pvu = q0.pvv = nullptr; q0.pass0 = 1;
      if (moist) FEDG_VIT_LAUNCH(true, true, q0); else FEDG_VIT_LAUNCH(false, true, q0);
      VIParams q1 = p; q1.pvu = p.pvu_out; q1.pvv = p.pvv_out; q1.pass0 = 0;
      if (moist) FEDG_VIT_LAUNCH(true, true, q1); else FEDG_VIT_LAUNCH(false, true, q1);
    }
#undef FEDG_VIT_LAUNCH
    return;
  }
  // Two kernels for the flat geometry: this file's eight-lane kernel (partial pivoting over the whole reduced block) and the two-lane
  // block elimination of vi_solver2.cu.  Measured (profiles/r02_ab_vi_kernels.txt): the two-lane kernel is faster for the implicit
  // launches (2.76 vs 2.9 ms at 32x32x16) and for the explicit evaluation of the first ARK stage (0.55 vs 0.92 ms); both pass every
  // parity test, and at config 4's vertical acoustic CFL both sit at the same distance from the oracle (tools/cfg4_diag.py: 3.56e-10 /
  // 1.20e-8 of DDENS / MOMZ's own norm with either kernel, 1.4x / 1.8x the oracle's own response to a one-ulp perturbation), so the
  // faster one is the default.  FEDG_VI_KERNEL=1 / 2 forces one kernel for every launch (read at every launch: in-process A/B runs).
  int which = 2;
  { const char* e = getenv("FEDG_VI_KERNEL"); if (e && (e[0] == '1' || e[0] == '2')) which = e[0] - '0'; }
  if (which == 2 && p.htab && launch_vi2(p, *p.htab, moist, s)) return;
  const int ncol = p.Ne2D * 64;
  const int groups = VI_THREADS / 8;
  dim3 grid(ncol / groups), block(VI_THREADS);
  const size_t shmem = (144 + size_t(groups) * VI_SOL + size_t(VI_NQ + VI_PF) * VI_THREADS) * sizeof(double);
  // resident blocks per SM the implicit kernel is compiled for: 2 (244 registers, no spills; default) or 3 (168 registers,
  // 144 B of spills).  Measured at 32x32x16: 2.49 vs 2.68 ms per launch (average over the stages of IMEX_ARK324); local-memory
  // traffic costs more than the fourth warp per scheduler brings.  FEDG_VI_MINB=3 selects the other build (A/B measurements).
  static int minb = -1;
  if (minb < 0) { const char* e = getenv("FEDG_VI_MINB"); minb = (e && e[0] == '3') ? 3 : 2; }
#define FEDG_VI_LAUNCH(M, I, B)                                                                              \
  do {                                                                                                       \
    static bool attr_set = false;                                                                            \
    if (!attr_set) {                                                                                         \
      cudaFuncSetAttribute(vi_column_kernel<M, I, B, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shmem)); \
      attr_set = true;                                                                                       \
    }                                                                                                        \
    vi_column_kernel<M, I, B, false><<<grid, block, shmem, s>>>(p);                                          \
  } while (0)
  const bool implicit = p.impl_fac != 0.0;
  if (!implicit) { if (moist) FEDG_VI_LAUNCH(true, false, 4); else FEDG_VI_LAUNCH(false, false, 4); }
  else if (minb == 2) { if (moist) FEDG_VI_LAUNCH(true, true, 2); else FEDG_VI_LAUNCH(false, true, 2); }
  else { if (moist) FEDG_VI_LAUNCH(true, true, 3); else FEDG_VI_LAUNCH(false, true, 3); }
#undef FEDG_VI_LAUNCH
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
// Last IMEX stage fused with the modal filter of the step (atm_dyn_dgm_modalfilter_apply, dyn_dgm_modalfilter.F90:49-130,
// called at driver_nonhydro3d.F90:940-951):  q = F3D(Gsqrt (base + sum coef k)) / Gsqrt.  One block per element, one
// thread per node, the five variables move through the three 1D passes together (three block barriers per element);
// the filter tables sit in shared memory transposed so that the lanes of a warp read consecutive words.  Saves the
// write + read of the unfiltered state and replaces the stand-alone filter kernel (1.04 ms -> inside a 0.5 ms pass).
__global__ void lincomb_filter_kernel(const __grid_constant__ LinCombParams L, const ElemTables* __restrict__ tab,
                                      const double* __restrict__ gsqrt, int weighted, int np) {
  extern __shared__ double sm[];
  const int N2 = np * np, N3 = N2 * np;
  double* sFhT = sm;            // sFhT[l*np + i] = Fh[i][l]
  double* sFvT = sm + N2;
  double* A = sm + 2 * N2;      // [5][N3]
  double* B = A + NVAR * N3;    // [5][N3]
  const int n = threadIdx.x;
  const int i = n % np, j = (n / np) % np, k = n / N2;
  if (n < N2) { const int r = n / np, c = n % np; sFhT[c * np + r] = tab->Fh[n]; sFvT[c * np + r] = tab->Fv[n]; }
  const size_t gi = size_t(blockIdx.x) * N3 + n;
  const double G = weighted ? gsqrt[gi] : 1.0;
#pragma unroll
  for (int v = 0; v < NVAR; ++v) {
    double r = L.base[v][gi];
    for (int m = 0; m < L.nterm; ++m) r = r + L.coef[m] * L.k[m][v][gi];
    A[v * N3 + n] = G * r;
  }
  __syncthreads();
#pragma unroll
  for (int v = 0; v < NVAR; ++v) {
    const double* s = A + v * N3 + j * np + k * N2;
    double a = sFhT[i] * s[0];
    for (int l = 1; l < np; ++l) a += sFhT[l * np + i] * s[l];
    B[v * N3 + n] = a;
  }
  __syncthreads();
#pragma unroll
  for (int v = 0; v < NVAR; ++v) {
    const double* w = B + v * N3 + i + k * N2;
    double b = w[0] * sFhT[j];
    for (int l = 1; l < np; ++l) b += w[l * np] * sFhT[l * np + j];
    A[v * N3 + n] = b;
  }
  __syncthreads();
  const double rG = 1.0 / G;
#pragma unroll
  for (int v = 0; v < NVAR; ++v) {
    const double* s = A + v * N3 + i + j * np;
    double r = s[0] * sFvT[k];
    for (int l = 1; l < np; ++l) r += s[l * N2] * sFvT[l * np + k];
    L.out[v][gi] = r * rG;
  }
}
void launch_lincomb_filter(const LinCombParams& L, const ElemTables* tab, const double* gsqrt, bool weighted, int Ne, int np,
                           cudaStream_t s) {
  const int N3 = np * np * np;
  const size_t shmem = (size_t(2) * np * np + size_t(2) * NVAR * N3) * sizeof(double);
  lincomb_filter_kernel<<<Ne, N3, shmem, s>>>(L, tab, gsqrt, weighted ? 1 : 0, np);
}

void launch_lincomb(const LinCombParams& L, cudaStream_t s) {
  const int block = 256;
  lincomb_kernel<<<unsigned((L.n + block - 1) / block), block, 0, s>>>(L);
}

}  // namespace fedg
