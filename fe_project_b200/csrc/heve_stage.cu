// Fused explicit stage kernel of the regional HEVE / HEVI-explicit equations (sm_100a, FP64).
//
//   Rusanov face flux (a5/a6) -> volume divergence + lift (a2) -> tendency (a7/a8) -> RK stage update (a13)
//   [-> modal filter (a15), last stage] -> pressure of the NEW state (a4), carried to the next stage
// reading the stage-input state once and writing the stage-output state once.
//
// Thread mapping: one thread per node, 512 threads per block (one p=7 element, or eight p=3 elements).
//   phase 0  one elected thread per element issues TMA bulk copies (cp.async.bulk + mbarrier) of the
//            element's nine 4 KB input fields into a shared-memory stash; the D / filter tables are
//            staged in shared memory as well (no divergent constant-bank reads)
//   phase 1  node values stash -> registers
//   phase 2  face nodes: interior side from the stash, exterior side gathered through VMapP; flux jumps -> smem
//   phase 3  per variable: fluxes -> three padded smem layouts (rows along x, y, z), one barrier, three
//            8-term contractions read as 128-bit rows (broadcast inside a warp), lift, tendency, RK update,
//            optional modal filter, store
// The stash is dead after phase 2 and is aliased by the flux buffers.
#include <cstdint>
#include <cstdlib>

#include "fedg_internal.h"
#include "stage_common.cuh"

namespace fedg {

template <int NP>
struct StageGeo {
  static constexpr int N2 = NP * NP, N3 = N2 * NP, NFT = 6 * N2;
  static constexpr int RS = NP + 2;          // padded row length (doubles): rows stay 16-byte aligned, bank spread
  static constexpr int PLN = N2 * RS;        // one padded layout
  static constexpr int EPB = 512 / N3;       // elements per block
  static constexpr int NSTASH = 9;           // DDENS MOMX MOMY MOMZ DRHOT DENS_hyd PRES_hyd THERM_hyd DPRES
  static constexpr int STASH = NSTASH * N3;
  static constexpr int FLUX = 6 * PLN;       // 2 buffers x 3 layouts
  static constexpr int UNI = STASH > FLUX + PLN ? STASH : FLUX + PLN;   // + z-rows of DDENS
  static constexpr int PER_EL = UNI + NVAR * NFT;
  static constexpr int NTAB = 4 * NP * NP + 2 * NP;   // D, Fh, Fv, VP, Lw
  static constexpr int TAB_PAD = (NTAB + 1) & ~1;
  static constexpr size_t SMEM_BYTES = (size_t(TAB_PAD) + size_t(EPB) * PER_EL) * sizeof(double) + 16 * EPB;
};

// sum_l row[l] * f[l], l ascending (the reference's unrolled left-to-right order, kernel.F90.erb:66-70)
template <int NP>
__device__ __forceinline__ double dot_row(const double* __restrict__ m, const double* __restrict__ f) {
  const double2* m2 = reinterpret_cast<const double2*>(m);
  const double2* f2 = reinterpret_cast<const double2*>(f);
  double2 a = m2[0], b = f2[0];
  double s = a.x * b.x;
  s += a.y * b.y;
#pragma unroll
  for (int q = 1; q < NP / 2; ++q) {
    a = m2[q]; b = f2[q];
    s += a.x * b.x;
    s += a.y * b.y;
  }
  return s;
}
template <int NP>
__device__ __forceinline__ double dot_reg(const double (&m)[NP], const double* __restrict__ f) {
  const double2* f2 = reinterpret_cast<const double2*>(f);
  double2 b = f2[0];
  double s = m[0] * b.x;
  s += m[1] * b.y;
#pragma unroll
  for (int q = 1; q < NP / 2; ++q) {
    b = f2[q];
    s += m[2 * q] * b.x;
    s += m[2 * q + 1] * b.y;
  }
  return s;
}

template <int NP, bool TERRAIN, bool MOIST, bool HEVI, int MINB>
__global__ void __launch_bounds__(512, MINB) stage_kernel(const __grid_constant__ StageParams P) {
  using G = StageGeo<NP>;
  constexpr int N2 = G::N2, N3 = G::N3, NFT = G::NFT, RS = G::RS, PLN = G::PLN, EPB = G::EPB;
  const int tid = threadIdx.x;
  const int el = tid / N3, n = tid - el * N3;
  const int i = n % NP, j = (n / NP) % NP, k = n / N2;
  const int nel = P.elem_list ? P.nelem : P.Ne;
  int slot = blockIdx.x * EPB + el;
  const bool live = slot < nel;
  if (!live) slot = nel - 1;
  const int ke = P.elem_list ? P.elem_list[slot] : slot;
  const size_t eb = size_t(ke) * N3;
  const size_t gn = eb + n;

  extern __shared__ __align__(16) double smem[];
  double* sTabD = smem;                    // D[i][l]
  double* sTabFh = sTabD + NP * NP;
  double* sTabFv = sTabFh + NP * NP;
  double* sTabVP = sTabFv + NP * NP;       // VPOrdM1[k][l]
  double* sTabLw = sTabVP + NP * NP;       // lift1d[m][side]
  double* sEl = smem + G::TAB_PAD + size_t(el) * G::PER_EL;
  double* sStash = sEl;                    // [NSTASH][N3]
  double* sFlux = sEl;                     // aliases the stash after phase 2
  double* sDel = sEl + G::UNI;             // [NVAR][NFT]
  uint64_t* sBar = reinterpret_cast<uint64_t*>(smem + G::TAB_PAD + size_t(EPB) * G::PER_EL) + 2 * el;

  // ---- phase 0: TMA bulk loads of the element's input fields
  if (n == 0) {
    mbar_init(sBar, 1);
    constexpr uint32_t BYTES = N3 * sizeof(double);
    mbar_expect_tx(sBar, G::NSTASH * BYTES);
    tma_load_1d(sStash + 0 * N3, P.qin[V_DDENS] + eb, BYTES, sBar);
    tma_load_1d(sStash + 1 * N3, P.qin[V_MOMX] + eb, BYTES, sBar);
    tma_load_1d(sStash + 2 * N3, P.qin[V_MOMY] + eb, BYTES, sBar);
    tma_load_1d(sStash + 3 * N3, P.qin[V_MOMZ] + eb, BYTES, sBar);
    tma_load_1d(sStash + 4 * N3, P.qin[V_DRHOT] + eb, BYTES, sBar);
    tma_load_1d(sStash + 5 * N3, P.dens_hyd + eb, BYTES, sBar);
    tma_load_1d(sStash + 6 * N3, P.pres_hyd + eb, BYTES, sBar);
    tma_load_1d(sStash + 7 * N3, P.therm_hyd + eb, BYTES, sBar);
    tma_load_1d(sStash + 8 * N3, P.dpin + eb, BYTES, sBar);
  }
  if (tid < G::NTAB) {
    double v;
    if (tid < NP * NP) v = P.tab->D[tid];
    else if (tid < 2 * NP * NP) v = P.tab->Fh[tid - NP * NP];
    else if (tid < 3 * NP * NP) v = P.tab->Fv[tid - 2 * NP * NP];
    else if (tid < 4 * NP * NP) v = P.tab->VP[tid - 3 * NP * NP];
    else v = P.tab->Lw[tid - 4 * NP * NP];
    smem[tid] = v;
  }
  // geometry of this element while the copies are in flight
  const double E11 = P.escale[ke], E22 = P.escale[P.Ne + ke], E33 = P.escale[2 * size_t(P.Ne) + ke];
  const int ke2d = P.emap2d[ke];
  const double cor = P.has_cor ? P.coriolis[size_t(ke2d) * N2 + (n % N2)] : 0.0;
  double Gn = 1.0, G13n = 0.0, G23n = 0.0, gH = 1.0;
  if (TERRAIN) { Gn = P.gsqrt[gn]; G13n = P.g13[gn]; G23n = P.g23[gn]; gH = P.gsqrtH[size_t(ke2d) * N2 + (n % N2)]; }
  __syncthreads();   // barrier init + tables visible
  mbar_wait(sBar, 0);

  // ---- phase 1: node values
  const double dd = sStash[0 * N3 + n], mx = sStash[1 * N3 + n], my = sStash[2 * N3 + n], mz = sStash[3 * N3 + n];
  const double dr = sStash[4 * N3 + n];
  const double dp = sStash[8 * N3 + n];
  const double rdens = 1.0 / (dd + sStash[5 * N3 + n]);
  const double pt = (sStash[7 * N3 + n] + dr) * rdens;

  // ---- phase 2: face flux jumps
  {
    const double gamm = P.c.gamm;
    const size_t fb = size_t(ke) * NFT;
    for (int m = n; m < NFT; m += N3) {
      const int f = m / N2, fp = m - f * N2, a = fp % NP, b = fp / NP;
      int nloc;
      switch (f) {
        case 0: nloc = a + b * N2; break;
        case 1: nloc = (NP - 1) + a * NP + b * N2; break;
        case 2: nloc = a + (NP - 1) * NP + b * N2; break;
        case 3: nloc = a * NP + b * N2; break;
        case 4: nloc = fp; break;
        default: nloc = fp + (NP - 1) * N2; break;
      }
      const size_t iP = size_t(P.vmapP[fb + m]);
      double GsM = 1.0, G13M = 0.0, G23M = 0.0, GsP = 1.0, G13P = 0.0, G23P = 0.0;
      if (TERRAIN) {
        GsM = P.gsqrt[eb + nloc]; G13M = P.g13[eb + nloc]; G23M = P.g23[eb + nloc];
        GsP = P.gsqrt[iP]; G13P = P.g13[iP]; G23P = P.g23[iP];
      }
      FaceSide M, Q;
      make_side<TERRAIN>(M, sStash[0 * N3 + nloc], sStash[1 * N3 + nloc], sStash[2 * N3 + nloc], sStash[3 * N3 + nloc],
                         sStash[4 * N3 + nloc], sStash[5 * N3 + nloc], sStash[6 * N3 + nloc], sStash[7 * N3 + nloc],
                         sStash[8 * N3 + nloc], GsM, G13M, G23M);
      make_side<TERRAIN>(Q, P.qin[V_DDENS][iP], P.qin[V_MOMX][iP], P.qin[V_MOMY][iP], P.qin[V_MOMZ][iP], P.qin[V_DRHOT][iP],
                         P.dens_hyd[iP], P.pres_hyd[iP], P.therm_hyd[iP], P.dpin[iP], GsP, G13P, G23P);
      const double hf = P.fscale[size_t(f) * P.Ne + ke] * 0.5;
      double o5[NVAR];
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
  __syncthreads();   // stash is dead from here on: the flux buffers alias it

  // ---- phase 3: volume terms, tendency, stage update, variable by variable
  double Di[NP], Dj[NP];
#pragma unroll
  for (int l = 0; l < NP; ++l) { Di[l] = sTabD[i * NP + l]; Dj[l] = sTabD[j * NP + l]; }
  const double lwi0 = sTabLw[i * 2], lwi1 = sTabLw[i * 2 + 1], lwj0 = sTabLw[j * 2], lwj1 = sTabLw[j * 2 + 1];
  const double lwk0 = sTabLw[k * 2], lwk1 = sTabLw[k * 2 + 1];
  const double RGv = TERRAIN ? 1.0 / (Gn / gH) : 1.0;
  const double RGs = TERRAIN ? 1.0 / Gn : 1.0;
  const double fx0 = Gn * mx, fy0 = Gn * my;
  const double fz0 = TERRAIN ? Gn * (mz * RGv + G13n * mx + G23n * my) : mz;
  const double GP = Gn * dp;
  const bool tend_mode = P.tend_out[0] != nullptr;
  const int rowX = (k * NP + j) * RS, rowY = (k * NP + i) * RS, rowZ = (j * NP + i) * RS;
  double* sDDz = sFlux + G::FLUX;   // z-rows of DDENS for the buoyancy term
  double drho = 0.0, dr_new = dr;

  const int order[NVAR] = {V_DDENS, V_DRHOT, V_MOMZ, V_MOMX, V_MOMY};
#pragma unroll
  for (int iv = 0; iv < NVAR; ++iv) {
    const int v = order[iv];
    const int b = P.do_filter ? 0 : (iv & 1);
    double* bx = sFlux + b * 3 * PLN;
    double* by = bx + PLN;
    double* bz = by + PLN;
    double Fx, Fy, Fz, q;
    if (v == V_DDENS) { Fx = fx0; Fy = fy0; Fz = fz0; q = dd; }
    else if (v == V_DRHOT) { Fx = fx0 * pt; Fy = fy0 * pt; Fz = fz0 * pt; q = dr; }
    else if (v == V_MOMZ) { const double w = mz * rdens; Fx = fx0 * w; Fy = fy0 * w; Fz = HEVI ? fz0 * w : fz0 * w + GP * RGv; q = mz; }
    else if (v == V_MOMX) { const double u = mx * rdens; Fx = fx0 * u + GP; Fy = fy0 * u; Fz = TERRAIN ? fz0 * u + GP * G13n : fz0 * u; q = mx; }
    else { const double vv = my * rdens; Fx = fx0 * vv; Fy = fy0 * vv + GP; Fz = TERRAIN ? fz0 * vv + GP * G23n : fz0 * vv; q = my; }
    // HEVI: the vertical mass / theta fluxes are treated implicitly (rhot_hevi.F90:440-452 drops E33*Dz)
    const bool need_z = !(HEVI && (v == V_DDENS || v == V_DRHOT));
    bx[rowX + i] = Fx;
    by[rowY + j] = Fy;
    if (need_z) bz[rowZ + k] = Fz;
    if (!HEVI && v == V_DDENS) sDDz[rowZ + k] = dd;
    __syncthreads();

    const double dx = dot_reg<NP>(Di, bx + rowX);
    const double dy = dot_reg<NP>(Dj, by + rowY);
    double dz = 0.0;
    if (need_z) dz = dot_row<NP>(sTabD + k * NP, bz + rowZ);
    if (!HEVI && v == V_DDENS) drho = dot_row<NP>(sTabVP + k * NP, sDDz + rowZ);   // VFilterPM1 (rhot_heve.F90:442-443)
    const double* sD = sDel + v * NFT;
    const double lift = lwj0 * sD[i + k * NP] + lwi1 * sD[N2 + j + k * NP] + lwj1 * sD[2 * N2 + i + k * NP] +
                        lwi0 * sD[3 * N2 + j + k * NP] + lwk0 * sD[4 * N2 + i + j * NP] + lwk1 * sD[5 * N2 + i + j * NP];
    const double div = need_z ? (E11 * dx + E22 * dy + E33 * dz + lift) * RGs : (E11 * dx + E22 * dy + lift) * RGs;
    double tend;
    if (v == V_MOMZ) tend = HEVI ? -div : -div - P.c.GRAV * drho;
    else if (v == V_MOMX) tend = ((P.has_phyd ? -P.dphydx[gn] : 0.0) + cor * my) - div;
    else if (v == V_MOMY) tend = ((P.has_phyd ? -P.dphydy[gn] : 0.0) - cor * mx) - div;
    else tend = -div;
    if (P.sponge && live && (v == V_MOMX || v == V_MOMY || v == V_MOMZ))   // AtmDynSpongeLayer%AddTend (spongelayer.F90:129-185)
      tend -= ((v == V_MOMZ) ? 1.0 : P.sponge_h) * P.sponge[gn] * q;
    if (P.has_phyt && live) {   // add_phy_tend (driver_nonhydro3d.F90:1098-1178), non-conservative form
      const int pv = (v == V_DDENS) ? 0 : (v == V_MOMX) ? 1 : (v == V_MOMY) ? 2 : (v == V_MOMZ) ? 3 : 4;
      tend += P.phyt[pv][gn];
      if (v == V_DRHOT) {
        const double R = MOIST ? P.rtot[gn] : P.c.Rdry, cp = MOIST ? P.cptot[gn] : P.c.CPdry;
        tend += P.phyt[5][gn] / (cp * pow((P.pres_hyd[gn] + dp) * P.c.rP0, R / cp));
      }
    }

    if (tend_mode) {
      if (live) P.tend_out[v][gn] = tend;
      continue;
    }
    // RK stage update (scale_timeint_rk.F90:1182-1266 low storage, :2201-2355 general with one buffer)
    double base = 0.0;
    if (P.rk.use_q0) base = P.rk.c_q0 * P.q0[v][gn];
    if (P.rk.add_vt) base = P.vt[v][gn];
    double r = base + P.rk.c_q * q + P.rk.c_k * tend;
    if (P.rk.vt_update) {
      const double vb = P.rk.vt_init ? P.rk.vt_init_q * q : P.vt[v][gn];
      if (live) P.vt[v][gn] = vb + P.rk.vt_q * q + P.rk.vt_k * tend;
    }
    if (P.do_filter) {  // modal filter of the Gsqrt-weighted variable (dyn_dgm_modalfilter.F90:49-130): x, y, z passes
      double* fxb = sFlux + 3 * PLN;
      double* fyb = fxb + PLN;
      double* fzb = fyb + PLN;
      fxb[rowX + i] = Gn * r;
      __syncthreads();
      const double r1 = dot_row<NP>(sTabFh + i * NP, fxb + rowX);
      fyb[rowY + j] = r1;
      __syncthreads();
      const double r2 = dot_row<NP>(sTabFh + j * NP, fyb + rowY);
      fzb[rowZ + k] = r2;
      __syncthreads();
      r = dot_row<NP>(sTabFv + k * NP, fzb + rowZ) * (1.0 / Gn);
    }
    if (live) P.qout[v][gn] = r;
    if (v == V_DRHOT) dr_new = r;
  }

  if (!tend_mode) {  // pressure of the new state: next stage's DPRES; PRES diagnostic at the end of Update (driver:954-959)
    const double ph = P.pres_hyd[gn], th = P.therm_hyd[gn];
    const double R = MOIST ? P.rtot[gn] : P.c.Rdry;
    const double e = MOIST ? P.cptot[gn] / P.cvtot[gn] : P.c.CPovCV;
    const double pr = eos_pres(R, P.c.rP0, th + dr_new, e, P.c.PRES00);
    if (live) {
      P.dpout[gn] = pr - ph;
      if (P.write_pres) P.pres_out[gn] = pr;
    }
  }
}

template <int NP, bool HEVI, int MINB>
static void launch_stage_np(const StageParams& p, bool terrain, bool moist, cudaStream_t s) {
  using G = StageGeo<NP>;
  const size_t shmem = G::SMEM_BYTES;
  const int nel = p.elem_list ? p.nelem : p.Ne;
  dim3 grid((nel + G::EPB - 1) / G::EPB), block(G::EPB * G::N3);   // 512 threads for p = 7, 3, 1; 2 x 216 for p = 5
#define FEDG_LAUNCH(T, M)                                                                                             \
  do {                                                                                                                \
    static bool attr_set = false;                                                                                     \
    if (!attr_set) {                                                                                                  \
      cudaFuncSetAttribute(stage_kernel<NP, T, M, HEVI, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem); \
      attr_set = true;                                                                                                \
    }                                                                                                                 \
    stage_kernel<NP, T, M, HEVI, MINB><<<grid, block, shmem, s>>>(p);                                                 \
  } while (0)
  if (terrain) { if (moist) FEDG_LAUNCH(true, true); else FEDG_LAUNCH(true, false); }
  else { if (moist) FEDG_LAUNCH(false, true); else FEDG_LAUNCH(false, false); }
#undef FEDG_LAUNCH
}

// Resident blocks per SM the kernel is compiled for (register cap 64 vs 128).  Tuning knob, default 2;
// FEDG_STAGE_MINB=1 selects the 128-register build.
static int stage_minb() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FEDG_STAGE_MINB"); v = (e && e[0] == '1') ? 1 : 2; }
  return v;
}

void launch_stage_p7(const StageParams& p, bool terrain, bool moist, bool hevi, cudaStream_t s);  // stage_p7.cu

// FEDG_STAGE_GENERIC=1 routes p = 7 through the generic node-per-thread kernel (A/B measurements).
static bool stage_generic() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FEDG_STAGE_GENERIC"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

void launch_stage(const StageParams& p, int np, bool terrain, bool moist, bool hevi, cudaStream_t s) {
  if (np == 8 && !stage_generic()) { launch_stage_p7(p, terrain, moist, hevi, s); return; }
  if (np == 8) {
    if (stage_minb() == 1) { if (hevi) launch_stage_np<8, true, 1>(p, terrain, moist, s); else launch_stage_np<8, false, 1>(p, terrain, moist, s); }
    else { if (hevi) launch_stage_np<8, true, 2>(p, terrain, moist, s); else launch_stage_np<8, false, 2>(p, terrain, moist, s); }
  } else if (np == 4) {
    if (hevi) launch_stage_np<4, true, 2>(p, terrain, moist, s); else launch_stage_np<4, false, 2>(p, terrain, moist, s);
  } else if (np == 6) {   // p = 5 (HEVE only: the vertical-implicit kernels are built for p = 7)
    launch_stage_np<6, false, 2>(p, terrain, moist, s);
  } else if (np == 2) {   // p = 1
    launch_stage_np<2, false, 2>(p, terrain, moist, s);
  }
  // other orders are rejected at fedg_create
}

}  // namespace fedg
