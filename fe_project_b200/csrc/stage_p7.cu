// Fused explicit stage kernel for p = 7 elements (8x8x8 nodes), FP64 tensor-core contractions (sm_100a).
//
// Why tensor cores here: with one thread per node and the three 8-term contractions read as rows from shared
// memory (stage_kernel<8,...> in heve_stage.cu) ncu shows the shared-memory data pipe as the binding unit
// (l1tex__data_pipe_lsu_wavefronts 80 %, profiles/r01_v2_*): every node re-reads 3 x 64 B of flux rows plus 64 B of
// D rows per variable.  mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4, measured 37.1 TFLOP/s vs 34.1 for DFMA) shares the
// operands inside the warp instead: each flux value is read from shared memory once per direction.
//
// Mapping (block = 256 threads = 8 warps, one element):  g = lane/4, t = lane%4
//   warp w owns the plane k = w; lane (g,t) owns the two nodes (i = 2t, 2t+1 ; j = g ; k = w), i.e. the C fragment
//   of the 8x8 tile  out^T[j][i]  of that plane.  With this orientation
//     x-derivative   out^T = Fx^T * D^T      (A = data, B = D[i=g][l=t]  constant fragment)
//     y-derivative   out^T += D * Fy^T       (A = D[j=g][l=t] constant,  B = data)
//   accumulate in the same registers, and both data fragments come from the warp's own plane (no block barrier).
//   The z-derivative uses tiles at fixed j = w:  out_j[k][i] = D * Fz_j  (A = D[k=g][l=t]); its C fragment belongs
//   to other warps' nodes, so z-results cross through shared memory once (two block barriers per stage, batched
//   over the five variables).  The lift is three more k=4 steps with (lift weights x face jumps) as operands.
//   E11/E22/E33 (constant per element on MeshCubeDom3D) are folded into the constant fragments.
#include <cstdint>
#include <cstdlib>

#include "fedg_internal.h"
#include "stage_common.cuh"

namespace fedg {

namespace p7 {
constexpr int NP = 8, N2 = 64, N3 = 512, NFT = 384;
constexpr int KS_FZ = 70;   // k-stride of the z-staging layout  [k][j][i]: conflict-free B-fragment reads at k = 2t, 2t+1
constexpr int KS_Z = 72;    // k-stride of the z-result layout   [k][j][i]: conflict-free 128-bit writes and reads
constexpr int PLS = 10;     // row stride of the per-warp plane  [j][i]: conflict-free B-fragment reads at j = 2t, 2t+1
constexpr int TAB = 4 * 64 + 16;                       // D, Lw, VP, Fh, Fv: the leading 272 doubles of ElemTables
constexpr int STASH = 9 * N3;
constexpr int ZREG = NVAR * NP * KS_FZ + NVAR * NP * KS_Z;   // sFz + sZ alias the stash
constexpr int REGA = STASH > ZREG ? STASH : ZREG;
constexpr int PLANES = 8 * 2 * NP * PLS;
constexpr int SM_DOUBLES = TAB + REGA + NVAR * NFT + PLANES;
constexpr size_t SMEM_BYTES = size_t(SM_DOUBLES) * sizeof(double) + 16;
}  // namespace p7

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// the nine fields staged per element, in stash order
__device__ __forceinline__ const double* stage_field(const StageParams& P, int l) {
  switch (l) {
    case 0: return P.qin[V_DDENS];
    case 1: return P.qin[V_MOMX];
    case 2: return P.qin[V_MOMY];
    case 3: return P.qin[V_MOMZ];
    case 4: return P.qin[V_DRHOT];
    case 5: return P.dens_hyd;
    case 6: return P.pres_hyd;
    case 7: return P.therm_hyd;
    default: return P.dpin;
  }
}

// PLAIN: no sponge layer, no physics tendencies, and the output the step of the equation set takes (HEVE: the new state,
// HEVI: the tendency) -- the dynamics-only step; the other branches and their code (a pow() per node among it) are compiled out.
template <bool TERRAIN, bool MOIST, bool HEVI, bool GLOBAL, bool PLAIN>
__global__ void __launch_bounds__(256, 3) stage_p7_kernel(const __grid_constant__ StageParams P) {
  using namespace p7;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int ke = P.elem_list ? P.elem_list[blockIdx.x] : int(blockIdx.x);
  const size_t eb = size_t(ke) * N3;
  const int n0 = 2 * t + 8 * g + 64 * w;     // own nodes n0, n0 + 1
  const size_t gn = eb + n0;

  extern __shared__ __align__(16) double smem[];
  // tables arrive as one bulk copy of the device ElemTables struct: D, Lw, VP, Fh, Fv (fedg_internal.h)
  double* sTabD = smem;
  double* sTabLw = smem + 64;
  double* sTabVP = smem + 80;
  double* sTabFh = smem + 144;
  double* sTabFv = smem + 208;
  double* sStash = smem + TAB;               // [9][512], dead after the face phase
  double* sFz = smem + TAB;                  // [5][8*KS_FZ]   aliases the stash
  double* sZ = sFz + NVAR * NP * KS_FZ;      // [5][8*KS_Z]
  double* sDel = smem + TAB + REGA;          // [5][384]
  double* sExt = sDel + NVAR * NFT;          // [2][9][64] exterior side of the z faces; aliases the planes (first written in phase 5)
  double* sPl = sDel + NVAR * NFT + size_t(w) * 2 * NP * PLS;   // this warp's two planes [j][i], row stride PLS, used alternately
  uint64_t* sBar = reinterpret_cast<uint64_t*>(smem + SM_DOUBLES);
  const size_t fb = size_t(ke) * NFT;
  // exterior z-face values by bulk copy when every z face of the mesh maps to 64 consecutive nodes (the element above / below
  // or a halo face: checked at fedg_create); the terrain instantiation keeps the gathers (it needs three metric fields more)
  const bool zext = !TERRAIN && P.zface_contig;

  // ---- phase 0: TMA bulk loads issued by one thread: nine input fields of the element, 2 x 9 exterior z-face rows of 512 B, the
  //      operator tables.  (Issuing the 28 copies from 28 lanes of warp 0 at once was measured SLOWER: 0.520 vs 0.471 ms per launch
  //      in the same process, profiles/r02_ab_stage_tma_lanes.txt -- the lanes diverge over the field switch and the address set-up.)
  if (tid == 0) mbar_init(sBar, 1);
  __syncthreads();   // barrier initialised before anybody polls it.  Placed here, ahead of every global load: the warps run
                     // independently from now to the end of the face phase (behind the gathers it made every warp wait for
                     // the slowest VMapP fetch of the block: 10 % of the stall samples)
  if (tid == 0) {
    constexpr uint32_t BYTES = N3 * sizeof(double), FBYTES = N2 * sizeof(double), TBYTES = TAB * sizeof(double);
    mbar_expect_tx(sBar, 9 * BYTES + TBYTES + (zext ? 18 * FBYTES : 0u));
#pragma unroll
    for (int l = 0; l < 9; ++l) tma_load_1d(sStash + l * N3, stage_field(P, l) + eb, BYTES, sBar);
    tma_load_1d(smem, P.tab, TBYTES, sBar);
    if (zext) {
      const size_t ib4 = size_t(P.vmapP[fb + 4 * N2]), ib5 = size_t(P.vmapP[fb + 5 * N2]);
#pragma unroll
      for (int l = 0; l < 9; ++l) {
        tma_load_1d(sExt + l * N2, stage_field(P, l) + ib4, FBYTES, sBar);
        tma_load_1d(sExt + (9 + l) * N2, stage_field(P, l) + ib5, FBYTES, sBar);
      }
    }
    // fields read per node late in the kernel (background pressure gradient, RK operands): pull the element's 4 KB of each
    // into L2 now, so that the loads in phase 5 find them there instead of paying a DRAM round trip per variable
    if (P.l2_prefetch) {
      if (P.has_phyd) { tma_prefetch_l2(P.dphydx + eb, BYTES); tma_prefetch_l2(P.dphydy + eb, BYTES); }
      if (PLAIN ? !HEVI : (P.tend_out[0] == nullptr)) {
        if (P.rk.use_q0) {
#pragma unroll
          for (int v = 0; v < NVAR; ++v) tma_prefetch_l2(P.q0[v] + eb, BYTES);
        }
        if (P.rk.add_vt || (P.rk.vt_update && !P.rk.vt_init)) {
#pragma unroll
          for (int v = 0; v < NVAR; ++v) tma_prefetch_l2(P.vt[v] + eb, BYTES);
        }
      }
    }
  }
  const double E11 = P.escale[ke], E22 = P.escale[P.Ne + ke], E33 = P.escale[2 * size_t(P.Ne) + ke];
  const int ke2d = P.emap2d[ke];
  double2 cor = make_double2(0.0, 0.0), Gn = make_double2(1.0, 1.0), G13n = make_double2(0.0, 0.0), G23n = make_double2(0.0, 0.0),
          gH = make_double2(1.0, 1.0);
  if (P.has_cor) cor = *reinterpret_cast<const double2*>(P.coriolis + size_t(ke2d) * N2 + 2 * t + 8 * g);
  if (TERRAIN) {
    Gn = *reinterpret_cast<const double2*>(P.gsqrt + gn);
    G13n = *reinterpret_cast<const double2*>(P.g13 + gn);
    G23n = *reinterpret_cast<const double2*>(P.g23 + gn);
    gH = *reinterpret_cast<const double2*>(P.gsqrtH + size_t(ke2d) * N2 + 2 * t + 8 * g);
  }
  // global equation set: horizontal metric of the own two nodes (2D tables, the same for every plane k)
  double2 G11n = make_double2(1.0, 1.0), G12n = make_double2(0.0, 0.0), G22n = make_double2(1.0, 1.0), Xn = make_double2(0.0, 0.0),
          Yn = make_double2(0.0, 0.0);
  const size_t n2d = size_t(P.Ne2D) * N2;
  if (GLOBAL) {
    const size_t h = size_t(ke2d) * N2 + 2 * t + 8 * g;
    Gn = *reinterpret_cast<const double2*>(P.g2d + h);
    gH = Gn;
    G11n = *reinterpret_cast<const double2*>(P.g2d + n2d + h);
    G12n = *reinterpret_cast<const double2*>(P.g2d + 2 * n2d + h);
    G22n = *reinterpret_cast<const double2*>(P.g2d + 3 * n2d + h);
    Xn = *reinterpret_cast<const double2*>(P.g2d + 4 * n2d + h);
    Yn = *reinterpret_cast<const double2*>(P.g2d + 5 * n2d + h);
  }
  // exterior-side gather of this thread's first face node, issued while the bulk copies are in flight
  // (computing the index of in-tile lateral neighbours arithmetically on structured meshes instead of fetching VMapP was
  //  measured slower: 0.4748 vs 0.4633 ms per launch, the integer divisions cost more than the L2 hit)
  RawSide<TERRAIN> pre;
  const size_t iP0 = size_t(P.vmapP[fb + tid]);
  pre.load(P, iP0);
  mbar_wait(sBar, 0);

  // ---- phase 2: face flux jumps (384 face nodes over 256 threads: one pass for all, a second one for warps 0-3)
  {
    const double gamm = P.c.gamm;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int m = tid + 256 * pass;
      if (m >= NFT) break;
      const int f = m >> 6, fp = m & 63, a = fp & 7, b = fp >> 3;
      int nloc;
      switch (f) {
        case 0: nloc = a + b * N2; break;
        case 1: nloc = (NP - 1) + a * NP + b * N2; break;
        case 2: nloc = a + (NP - 1) * NP + b * N2; break;
        case 3: nloc = a * NP + b * N2; break;
        case 4: nloc = fp; break;
        default: nloc = fp + (NP - 1) * N2; break;
      }
      RawSide<TERRAIN> ex;
      if (pass == 0) ex = pre;
      else if (zext) {
        const double* se = sExt + (f - 4) * 9 * N2 + fp;
        ex.dd = se[0]; ex.mx = se[N2]; ex.my = se[2 * N2]; ex.mz = se[3 * N2]; ex.dr = se[4 * N2];
        ex.dh = se[5 * N2]; ex.ph = se[6 * N2]; ex.th = se[7 * N2]; ex.dp = se[8 * N2];
        ex.Gs = 1.0; ex.G13 = 0.0; ex.G23 = 0.0;
      } else ex.load(P, size_t(P.vmapP[fb + m]));
      double GsM = 1.0, G13M = 0.0, G23M = 0.0;
      if (TERRAIN) { GsM = P.gsqrt[eb + nloc]; G13M = P.g13[eb + nloc]; G23M = P.g23[eb + nloc]; }
      double fG11 = 1.0, fG12 = 0.0, fG22 = 1.0;
      if (GLOBAL) {
        // Gsqrt of both sides: own 2D node; the neighbour's own value inside the tile, the own value in the halo
        // (fill_halo_metric, mesh_cubedspheredom3d.F90:657-678)
        const size_t h = size_t(ke2d) * N2 + (nloc & 63);
        GsM = P.g2d[h]; fG11 = P.g2d[n2d + h]; fG12 = P.g2d[2 * n2d + h]; fG22 = P.g2d[3 * n2d + h];
        const size_t iP = (pass == 0) ? iP0 : size_t(P.vmapP[fb + m]);
        ex.Gs = GsM;
        if (iP < size_t(P.Ne) * N3) ex.Gs = P.g2d[size_t(P.emap2d[iP >> 9]) * N2 + (iP & 63)];
      }
      FaceSide M, Q;
      make_side<TERRAIN>(M, sStash[0 * N3 + nloc], sStash[1 * N3 + nloc], sStash[2 * N3 + nloc], sStash[3 * N3 + nloc],
                         sStash[4 * N3 + nloc], sStash[5 * N3 + nloc], sStash[6 * N3 + nloc], sStash[7 * N3 + nloc],
                         sStash[8 * N3 + nloc], GsM, G13M, G23M);
      make_side<TERRAIN>(Q, ex.dd, ex.mx, ex.my, ex.mz, ex.dr, ex.dh, ex.ph, ex.th, ex.dp, ex.Gs, ex.G13, ex.G23);
      const double hf = P.fscale[size_t(f) * P.Ne + ke] * 0.5;
      double o5[NVAR];
      if (GLOBAL) {
        switch (f) {
          case 0: rusanov_global<1, HEVI>(M, Q, -1.0, gamm, hf, fG11, fG12, fG22, o5); break;
          case 1: rusanov_global<0, HEVI>(M, Q, 1.0, gamm, hf, fG11, fG12, fG22, o5); break;
          case 2: rusanov_global<1, HEVI>(M, Q, 1.0, gamm, hf, fG11, fG12, fG22, o5); break;
          case 3: rusanov_global<0, HEVI>(M, Q, -1.0, gamm, hf, fG11, fG12, fG22, o5); break;
          case 4: rusanov_global<2, HEVI>(M, Q, -1.0, gamm, hf, fG11, fG12, fG22, o5); break;
          default: rusanov_global<2, HEVI>(M, Q, 1.0, gamm, hf, fG11, fG12, fG22, o5); break;
        }
      } else
      switch (f) {
        case 0: rusanov<1, TERRAIN, HEVI>(M, Q, -1.0, gamm, hf, o5); break;
        case 1: rusanov<0, TERRAIN, HEVI>(M, Q, 1.0, gamm, hf, o5); break;
        case 2: rusanov<1, TERRAIN, HEVI>(M, Q, 1.0, gamm, hf, o5); break;
        case 3: rusanov<0, TERRAIN, HEVI>(M, Q, -1.0, gamm, hf, o5); break;
        case 4: rusanov<2, TERRAIN, HEVI>(M, Q, -1.0, gamm, hf, o5); break;
        default: rusanov<2, TERRAIN, HEVI>(M, Q, 1.0, gamm, hf, o5); break;
      }
#pragma unroll
      for (int v = 0; v < NVAR; ++v) sDel[v * NFT + m] = o5[v];
    }
  }

  // ---- own-node values (two adjacent nodes per lane: 128-bit shared loads).  Read after the face phase, right before the
  //      stash dies: held across the face phase they were spilled to local memory and reloaded
  const double2 dd = *reinterpret_cast<const double2*>(sStash + 0 * N3 + n0);
  const double2 mx = *reinterpret_cast<const double2*>(sStash + 1 * N3 + n0);
  const double2 my = *reinterpret_cast<const double2*>(sStash + 2 * N3 + n0);
  const double2 mz = *reinterpret_cast<const double2*>(sStash + 3 * N3 + n0);
  const double2 dr = *reinterpret_cast<const double2*>(sStash + 4 * N3 + n0);
  const double2 dp = *reinterpret_cast<const double2*>(sStash + 8 * N3 + n0);
  double2 rdens, pt;
  {
    const double2 dh = *reinterpret_cast<const double2*>(sStash + 5 * N3 + n0);
    const double2 th = *reinterpret_cast<const double2*>(sStash + 7 * N3 + n0);
    rdens.x = 1.0 / (dd.x + dh.x); rdens.y = 1.0 / (dd.y + dh.y);
    pt.x = (th.x + dr.x) * rdens.x; pt.y = (th.y + dr.y) * rdens.y;
  }
  double2 drho = make_double2(0.0, 0.0);
  if (!HEVI) {  // VFilterPM1 of DDENS along the column (rhot_heve.F90:442-443), l ascending
    const double* col = sStash + (n0 - 64 * w);
#pragma unroll
    for (int l = 0; l < NP; ++l) {
      const double2 c = *reinterpret_cast<const double2*>(col + 64 * l);
      const double vp = sTabVP[w * NP + l];
      if (l == 0) { drho.x = c.x * vp; drho.y = c.y * vp; }
      else { drho.x += c.x * vp; drho.y += c.y * vp; }
    }
  }

  // per-node flux building blocks
  double2 RGv = make_double2(1.0, 1.0), RGs = make_double2(1.0, 1.0);
  if (TERRAIN) { RGv.x = 1.0 / (Gn.x / gH.x); RGv.y = 1.0 / (Gn.y / gH.y); RGs.x = 1.0 / Gn.x; RGs.y = 1.0 / Gn.y; }
  if (GLOBAL) { RGs.x = 1.0 / Gn.x; RGs.y = 1.0 / Gn.y; }   // GsqrtV = 1: RGv stays 1
  const double2 fx0 = make_double2(Gn.x * mx.x, Gn.y * mx.y), fy0 = make_double2(Gn.x * my.x, Gn.y * my.y);
  double2 fz0 = mz;
  if (GLOBAL) fz0 = make_double2(Gn.x * mz.x, Gn.y * mz.y);
  if (TERRAIN) {
    fz0.x = Gn.x * (mz.x * RGv.x + G13n.x * mx.x + G23n.x * my.x);
    fz0.y = Gn.y * (mz.y * RGv.y + G13n.y * mx.y + G23n.y * my.y);
  }
  const double2 GP = make_double2(Gn.x * dp.x, Gn.y * dp.y);
  const double2 uu = make_double2(mx.x * rdens.x, mx.y * rdens.y), vv = make_double2(my.x * rdens.x, my.y * rdens.y),
                ww = make_double2(mz.x * rdens.x, mz.y * rdens.y);
  __syncthreads();   // stash dead (sFz / sZ alias it), sDel complete

  // ---- phase 3: stage the vertical fluxes of all variables   sFz[v][k*KS_FZ + j*8 + i]
  constexpr int order[NVAR] = {V_DDENS, V_DRHOT, V_MOMZ, V_MOMX, V_MOMY};
  const int ownFz = 2 * t + 8 * g + KS_FZ * w;
#pragma unroll
  for (int iv = 0; iv < NVAR; ++iv) {
    const int v = order[iv];
    if (HEVI && (v == V_DDENS || v == V_DRHOT)) continue;   // vertical mass / theta fluxes are implicit (rhot_hevi.F90:440-452)
    double2 Fz;
    if (v == V_DDENS) Fz = fz0;
    else if (v == V_DRHOT) Fz = make_double2(fz0.x * pt.x, fz0.y * pt.y);
    else if (v == V_MOMZ) Fz = HEVI ? make_double2(fz0.x * ww.x, fz0.y * ww.y) : make_double2(fz0.x * ww.x + GP.x * RGv.x, fz0.y * ww.y + GP.y * RGv.y);
    else if (v == V_MOMX) Fz = TERRAIN ? make_double2(fz0.x * uu.x + GP.x * G13n.x, fz0.y * uu.y + GP.y * G13n.y) : make_double2(fz0.x * uu.x, fz0.y * uu.y);
    else Fz = TERRAIN ? make_double2(fz0.x * vv.x + GP.x * G23n.x, fz0.y * vv.y + GP.y * G23n.y) : make_double2(fz0.x * vv.x, fz0.y * vv.y);
    *reinterpret_cast<double2*>(sFz + v * NP * KS_FZ + ownFz) = Fz;
  }
  // constant fragments: D[g][t], D[g][t+4] scaled by the element metric, lift weights Lw[g][s=t] (t < 2)
  // The contraction index of every 8-term product is split over the two k = 4 steps as l = 2t (first) and l = 2t + 1 (second):
  // with this order the A fragment of the x-derivative is the lane's own node pair, no exchange through shared memory.
  const double2 Dg = *reinterpret_cast<const double2*>(sTabD + g * NP + 2 * t);
  const double Dg0 = Dg.x, Dg1 = Dg.y;
  const double lwA = (t < 2) ? sTabLw[g * 2 + t] : 0.0;
  __syncthreads();

  // ---- phase 4: z-derivative + z-face lift, tile at fixed j = w:  out_j[k][i] = sum_l (E33 D)[k][l] Fz_j[l][i]
  {
    const double a0 = E33 * Dg0, a1 = E33 * Dg1;
#pragma unroll
    for (int iv = 0; iv < NVAR; ++iv) {
      const int v = order[iv];
      if (HEVI && (v == V_DDENS || v == V_DRHOT)) continue;
      const double* src = sFz + v * NP * KS_FZ + g + 8 * w;
      const double b0 = src[KS_FZ * 2 * t], b1 = src[KS_FZ * (2 * t + 1)];
      const double bl = (t < 2) ? sDel[v * NFT + (4 + t) * N2 + g + 8 * w] : 0.0;
      double c0 = 0.0, c1 = 0.0;
      dmma(c0, c1, a0, b0);
      dmma(c0, c1, a1, b1);
      dmma(c0, c1, lwA, bl);
      *reinterpret_cast<double2*>(sZ + v * NP * KS_Z + 2 * t + 8 * w + KS_Z * g) = make_double2(c0, c1);
    }
  }
  __syncthreads();

  // ---- phase 5: per variable x/y-derivative + lateral lift on the own plane, tendency, RK update, filter passes x/y
  const bool tend_mode = PLAIN ? HEVI : (P.tend_out[0] != nullptr);   // PLAIN: HEVI steps take the tendency, HEVE steps the new state
  const double bx0 = E11 * Dg0, bx1 = E11 * Dg1, ay0 = E22 * Dg0, ay1 = E22 * Dg1;
  const int ownP = PLS * g + 2 * t, ownZ = 2 * t + 8 * g + KS_Z * w;
  double2 drn = make_double2(0.0, 0.0);   // DRHOT of the new state, for its pressure
  // horizontal gradient of the background pressure: fetched one variable ahead of its use (MOMX, MOMY come last in `order`)
  double2 phx_c = make_double2(0.0, 0.0), phy_c = make_double2(0.0, 0.0);
  // background of the own nodes for the pressure of the new state: issued here, consumed after the loop
  double2 ph_pre = make_double2(0.0, 0.0), th_pre = make_double2(0.0, 0.0);
  if (!tend_mode) { ph_pre = *reinterpret_cast<const double2*>(P.pres_hyd + gn); th_pre = *reinterpret_cast<const double2*>(P.therm_hyd + gn); }
#pragma unroll
  for (int iv = 0; iv < NVAR; ++iv) {
    const int v = order[iv];
    double2 Fx, Fy, q;
    if (v == V_DDENS) { Fx = fx0; Fy = fy0; q = dd; }
    else if (v == V_DRHOT) { Fx = make_double2(fx0.x * pt.x, fx0.y * pt.y); Fy = make_double2(fy0.x * pt.x, fy0.y * pt.y); q = dr; }
    else if (v == V_MOMZ) { Fx = make_double2(fx0.x * ww.x, fx0.y * ww.y); Fy = make_double2(fy0.x * ww.x, fy0.y * ww.y); q = mz; }
    else if (v == V_MOMX) {
      if (GLOBAL) { Fx = make_double2(fx0.x * uu.x + G11n.x * GP.x, fx0.y * uu.y + G11n.y * GP.y); Fy = make_double2(fy0.x * uu.x + G12n.x * GP.x, fy0.y * uu.y + G12n.y * GP.y); }
      else { Fx = make_double2(fx0.x * uu.x + GP.x, fx0.y * uu.y + GP.y); Fy = make_double2(fy0.x * uu.x, fy0.y * uu.y); }
      q = mx;
    } else {
      if (GLOBAL) { Fx = make_double2(fx0.x * vv.x + G12n.x * GP.x, fx0.y * vv.y + G12n.y * GP.y); Fy = make_double2(fy0.x * vv.x + G22n.x * GP.x, fy0.y * vv.y + G22n.y * GP.y); }
      else { Fx = make_double2(fx0.x * vv.x, fx0.y * vv.y); Fy = make_double2(fy0.x * vv.x + GP.x, fy0.y * vv.y + GP.y); }
      q = my;
    }
    double2 q0v = make_double2(0.0, 0.0), vtv = make_double2(0.0, 0.0);
    if (!tend_mode) {
      if (P.rk.use_q0) q0v = *reinterpret_cast<const double2*>(P.q0[v] + gn);
      if (P.rk.add_vt || (P.rk.vt_update && !P.rk.vt_init)) vtv = *reinterpret_cast<const double2*>(P.vt[v] + gn);
    }
    if (P.has_phyd) {
      if (GLOBAL) { if (iv == 2) { phx_c = *reinterpret_cast<const double2*>(P.dphydx + gn); phy_c = *reinterpret_cast<const double2*>(P.dphydy + gn); } }
      else if (iv == 2) phx_c = *reinterpret_cast<const double2*>(P.dphydx + gn);
      else if (iv == 3) phy_c = *reinterpret_cast<const double2*>(P.dphydy + gn);
    }
    double* sPy = sPl + (iv & 1) * NP * PLS;   // alternate planes: the reads of variable iv - 2 finished before the __syncwarp of iv - 1
    if (P.do_filter) __syncwarp();              // ... but the filter of the previous variable read the other plane
    *reinterpret_cast<double2*>(sPy + ownP) = Fy;
    double c0 = 0.0, c1 = 0.0;
    if (!(HEVI && (v == V_DDENS || v == V_DRHOT))) {
      const double2 z = *reinterpret_cast<const double2*>(sZ + v * NP * KS_Z + ownZ);
      c0 = z.x; c1 = z.y;
    }
    __syncwarp();
    {
      // x: out^T[j][i] += Fx^T[j][l] * (E11 D)[i][l]      A = Fx(j = g, l = 2t | 2t+1) = the own pair,  B = const
      dmma(c0, c1, Fx.x, bx0);
      dmma(c0, c1, Fx.y, bx1);
      // y: out^T[j][i] += (E22 D)[j][l] * Fy^T[l][i]      A = const,  B = Fy(i = g, j = 2t | 2t+1)
      const double by0 = sPy[PLS * 2 * t + g], by1 = sPy[PLS * (2 * t + 1) + g];
      dmma(c0, c1, ay0, by0);
      dmma(c0, c1, ay1, by1);
      // lift, x faces (3: x-, 1: x+): A = jump(j = g, k = w) for s = t < 2, B = Lw[i = g][s]
      const double axl = (t < 2) ? sDel[v * NFT + (t == 0 ? 3 : 1) * N2 + g + 8 * w] : 0.0;
      dmma(c0, c1, axl, lwA);
      // lift, y faces (0: y-, 2: y+): A = Lw[j = g][s], B = jump(i = g, k = w)
      const double byl = (t < 2) ? sDel[v * NFT + (t == 0 ? 0 : 2) * N2 + g + 8 * w] : 0.0;
      dmma(c0, c1, lwA, byl);
    }
    const double2 div = make_double2(c0 * RGs.x, c1 * RGs.y);
    double2 tend;
    if (v == V_MOMZ) tend = HEVI ? make_double2(-div.x, -div.y) : make_double2(-div.x - P.c.GRAV * drho.x, -div.y - P.c.GRAV * drho.y);
    else if (GLOBAL && (v == V_MOMX || v == V_MOMY)) {
      // pressure-gradient of the background, metric (Christoffel) and Coriolis terms, globalnonhydro3d_rhot_hevi.F90:535-566
      double2 phx = make_double2(0.0, 0.0), phy = make_double2(0.0, 0.0);
      phx = phx_c; phy = phy_c;
      const double sg = (P.panel == 6) ? -1.0 : 1.0;
      const bool p14 = P.panel <= 4;
      double r[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const double X = c ? Xn.y : Xn.x, Y = c ? Yn.y : Yn.x, MX = c ? mx.y : mx.x, MY = c ? my.y : my.x;
        const double u = c ? uu.y : uu.x, w_ = c ? vv.y : vv.x;
        const double g11 = c ? G11n.y : G11n.x, g12 = c ? G12n.y : G12n.x, g22 = c ? G22n.y : G22n.x;
        const double px = c ? phx.y : phx.x, py = c ? phy.y : phy.x;
        const double two = 2.0 / (1.0 + X * X + Y * Y);
        double cori;
        if (v == V_MOMX) cori = sg * P.OHM * two * (-X * Y * MX + (1.0 + Y * Y) * MY);
        else cori = sg * P.OHM * two * (-(1.0 + X * X) * MX + X * Y * MY);
        if (p14) cori = sg * Y * cori;
        if (v == V_MOMX) r[c] = -(g11 * px + g12 * py) - two * Y * (X * Y * u - (1.0 + Y * Y) * w_) * MX + cori;
        else r[c] = -(g12 * px + g22 * py) - two * X * (-(1.0 + X * X) * u + X * Y * w_) * MY + cori;
      }
      tend = make_double2(r[0] - div.x, r[1] - div.y);
    } else if (v == V_MOMX) {
      double2 ph = make_double2(0.0, 0.0);
      ph = phx_c;
      tend = make_double2((-ph.x + cor.x * my.x) - div.x, (-ph.y + cor.y * my.y) - div.y);
    } else if (v == V_MOMY) {
      double2 ph = make_double2(0.0, 0.0);
      ph = phy_c;
      tend = make_double2((-ph.x - cor.x * mx.x) - div.x, (-ph.y - cor.y * mx.y) - div.y);
    } else tend = make_double2(-div.x, -div.y);

    if (!PLAIN && P.sponge && (v == V_MOMX || v == V_MOMY || v == V_MOMZ)) {   // AtmDynSpongeLayer%AddTend (spongelayer.F90:129-185)
      const double2 wc = *reinterpret_cast<const double2*>(P.sponge + gn);
      const double sf = (v == V_MOMZ) ? 1.0 : P.sponge_h;
      tend.x -= sf * wc.x * q.x; tend.y -= sf * wc.y * q.y;
    }
    if (!PLAIN && P.has_phyt) {   // add_phy_tend (driver_nonhydro3d.F90:1098-1178), non-conservative form: RHOT_tp + RHOH_p / (CP * EXNER)
      const int pv = (v == V_DDENS) ? 0 : (v == V_MOMX) ? 1 : (v == V_MOMY) ? 2 : (v == V_MOMZ) ? 3 : 4;
      const double2 tp = *reinterpret_cast<const double2*>(P.phyt[pv] + gn);
      tend.x += tp.x; tend.y += tp.y;
      if (v == V_DRHOT) {
        const double2 hh = *reinterpret_cast<const double2*>(P.phyt[5] + gn), ph = *reinterpret_cast<const double2*>(P.pres_hyd + gn);
        double2 R = make_double2(P.c.Rdry, P.c.Rdry), cp = make_double2(P.c.CPdry, P.c.CPdry);
        if (MOIST) { R = *reinterpret_cast<const double2*>(P.rtot + gn); cp = *reinterpret_cast<const double2*>(P.cptot + gn); }
        tend.x += hh.x / (cp.x * pow((ph.x + dp.x) * P.c.rP0, R.x / cp.x));
        tend.y += hh.y / (cp.y * pow((ph.y + dp.y) * P.c.rP0, R.y / cp.y));
      }
    }
    if (tend_mode) {
      *reinterpret_cast<double2*>(P.tend_out[v] + gn) = tend;
      continue;
    }
    // RK stage update (scale_timeint_rk.F90:1182-1266 low storage, :2201-2355 general with one buffer)
    double2 base = make_double2(0.0, 0.0);
    if (P.rk.use_q0) base = make_double2(P.rk.c_q0 * q0v.x, P.rk.c_q0 * q0v.y);
    if (P.rk.add_vt) base = vtv;
    double2 r = make_double2(base.x + P.rk.c_q * q.x + P.rk.c_k * tend.x, base.y + P.rk.c_q * q.y + P.rk.c_k * tend.y);
    if (P.rk.vt_update) {
      double2 vb;
      if (P.rk.vt_init) vb = make_double2(P.rk.vt_init_q * q.x, P.rk.vt_init_q * q.y);
      else vb = vtv;
      *reinterpret_cast<double2*>(P.vt[v] + gn) =
          make_double2(vb.x + P.rk.vt_q * q.x + P.rk.vt_k * tend.x, vb.y + P.rk.vt_q * q.y + P.rk.vt_k * tend.y);
    }
    if (P.do_filter) {
      // modal filter of the Gsqrt-weighted variable (dyn_dgm_modalfilter.F90:49-130): x and y passes on the own plane
      const double2 Fh2 = *reinterpret_cast<const double2*>(sTabFh + g * NP + 2 * t);
      double f0 = 0.0, f1 = 0.0;
      dmma(f0, f1, Gn.x * r.x, Fh2.x);                               // out^T[j][i] = g^T[j][l] Fh[i][l], A = the own pair
      dmma(f0, f1, Gn.y * r.y, Fh2.y);
      double* sPf = sPl + ((iv & 1) ^ 1) * NP * PLS;                 // the plane not holding this variable's Fy
      *reinterpret_cast<double2*>(sPf + ownP) = make_double2(f0, f1);
      __syncwarp();
      double h0 = 0.0, h1 = 0.0;
      dmma(h0, h1, Fh2.x, sPf[PLS * 2 * t + g]);                     // out^T[j][i] = Fh[j][l] r1^T[l][i]
      dmma(h0, h1, Fh2.y, sPf[PLS * (2 * t + 1) + g]);
      // stage for the z pass (sFz[v] is free: phase 4 has completed for every warp)
      *reinterpret_cast<double2*>(sFz + v * NP * KS_FZ + ownFz) = make_double2(h0, h1);
    }
    else {   // no filter: the stage output is final, store it now instead of carrying five results to the end of the loop
      *reinterpret_cast<double2*>(P.qout[v] + gn) = r;
      if (v == V_DRHOT) drn = r;
    }
  }

  if (tend_mode) return;

  if (P.do_filter) {
    __syncthreads();   // all planes staged; every warp has finished reading sZ
    const double a0 = sTabFv[g * NP + 2 * t], a1 = sTabFv[g * NP + 2 * t + 1];
#pragma unroll
    for (int v = 0; v < NVAR; ++v) {
      const double* src = sFz + v * NP * KS_FZ + g + 8 * w;
      double c0 = 0.0, c1 = 0.0;
      dmma(c0, c1, a0, src[KS_FZ * 2 * t]);                          // out_j[k][i] = Fv[k][l] r2_j[l][i]
      dmma(c0, c1, a1, src[KS_FZ * (2 * t + 1)]);
      *reinterpret_cast<double2*>(sZ + v * NP * KS_Z + 2 * t + 8 * w + KS_Z * g) = make_double2(c0, c1);
    }
    __syncthreads();
    const double2 rG = make_double2(1.0 / Gn.x, 1.0 / Gn.y);
#pragma unroll
    for (int v = 0; v < NVAR; ++v) {
      const double2 z = *reinterpret_cast<const double2*>(sZ + v * NP * KS_Z + ownZ);
      const double2 r = make_double2(z.x * rG.x, z.y * rG.y);
      *reinterpret_cast<double2*>(P.qout[v] + gn) = r;
      if (v == V_DRHOT) drn = r;
    }
  }

  {  // pressure of the new state: next stage's DPRES; PRES diagnostic at the end of Update (driver:954-959)
    const double2 ph = ph_pre, th = th_pre;
    double2 R = make_double2(P.c.Rdry, P.c.Rdry), e = make_double2(P.c.CPovCV, P.c.CPovCV);
    if (MOIST) {
      R = *reinterpret_cast<const double2*>(P.rtot + gn);
      const double2 cp = *reinterpret_cast<const double2*>(P.cptot + gn), cv = *reinterpret_cast<const double2*>(P.cvtot + gn);
      e = make_double2(cp.x / cv.x, cp.y / cv.y);
    }
    const double p0 = eos_pres_fast(R.x, P.c.rP0, th.x + drn.x, e.x, P.c.PRES00, P.exact_pow);
    const double p1 = eos_pres_fast(R.y, P.c.rP0, th.y + drn.y, e.y, P.c.PRES00, P.exact_pow);
    *reinterpret_cast<double2*>(P.dpout + gn) = make_double2(p0 - ph.x, p1 - ph.y);
    if (P.write_pres) *reinterpret_cast<double2*>(P.pres_out + gn) = make_double2(p0, p1);
  }
}

void launch_stage_p7(const StageParams& p, bool terrain, bool moist, bool hevi, cudaStream_t s) {
  const size_t shmem = p7::SMEM_BYTES;
  dim3 grid(p.elem_list ? p.nelem : p.Ne), block(256);
  const bool plain = !p.sponge && !p.has_phyt && ((p.tend_out[0] != nullptr) == hevi);
#define FEDG_LAUNCH2(T, M, H, G, PL)                                                                                  \
  do {                                                                                                                \
    static bool attr_set = false;                                                                                     \
    if (!attr_set) {                                                                                                  \
      cudaFuncSetAttribute(stage_p7_kernel<T, M, H, G, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem); \
      attr_set = true;                                                                                                \
    }                                                                                                                 \
    stage_p7_kernel<T, M, H, G, PL><<<grid, block, shmem, s>>>(p);                                                    \
  } while (0)
#define FEDG_LAUNCH(T, M, H, G)                                                \
  do {                                                                         \
    if (plain) FEDG_LAUNCH2(T, M, H, G, true); else FEDG_LAUNCH2(T, M, H, G, false); \
  } while (0)
  if (p.is_global) {   // GLOBALNONHYDRO3D_HEVI / _HEVE (flat, shallow atmosphere: enforced at fedg_create / fedg_dyn_init)
    if (hevi) { if (moist) FEDG_LAUNCH(false, true, true, true); else FEDG_LAUNCH(false, false, true, true); }
    else { if (moist) FEDG_LAUNCH(false, true, false, true); else FEDG_LAUNCH(false, false, false, true); }
  } else if (hevi) {
    if (terrain) { if (moist) FEDG_LAUNCH(true, true, true, false); else FEDG_LAUNCH(true, false, true, false); }
    else { if (moist) FEDG_LAUNCH(false, true, true, false); else FEDG_LAUNCH(false, false, true, false); }
  } else {
    if (terrain) { if (moist) FEDG_LAUNCH(true, true, false, false); else FEDG_LAUNCH(true, false, false, false); }
    else { if (moist) FEDG_LAUNCH(false, true, false, false); else FEDG_LAUNCH(false, false, false, false); }
  }
#undef FEDG_LAUNCH
#undef FEDG_LAUNCH2
}

}  // namespace fedg
