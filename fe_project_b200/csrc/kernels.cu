// CUDA kernels (sm_100a, FP64) of the DG dynamics hot path.
//
// Thread mapping of the fused stage kernel: one thread owns one vertical column (i,j) of an
// element and keeps its NP values along k in registers ("k-column" layout).  Global loads and
// stores of a plane k are NP*NP consecutive doubles (fully coalesced); the z-derivative, the
// vertical modal truncation and the vertical filter pass are register-only; x/y passes go through
// shared-memory planes padded to NP+1 to keep the 8-byte bank pattern conflict free.
#include <cstdio>

#include "fedg_internal.h"

namespace fedg {

__constant__ ElemTables cT;

void upload_tables(const ElemTables& t, cudaStream_t s) {
  cudaMemcpyToSymbolAsync(cT, &t, sizeof(ElemTables), 0, cudaMemcpyHostToDevice, s);
}

// PRES = P00 * (Rtot/P00 * RHOT)^(CPtot/CVtot)   (nonhydro3d_common.F90:467-474)
__device__ __forceinline__ double eos_pres(double R, double rP0, double rhot, double cpovcv, double P00) {
  return P00 * pow(R * rP0 * rhot, cpovcv);
}

// One side of a face node: Gsqrt-weighted state of rhot_heve_numflux.F90:1030-1069.
struct FaceSide {
  double gDD, gMX, gMY, gMZ, gDR, gDens, gRhot, Gs, RGv, G13, G23, Phyd, dp, Vel;
};

template <int AX, bool TERRAIN>
__device__ __forceinline__ void face_velocity(FaceSide& s, double sgn) {
  if (AX == 0) s.Vel = (s.gMX * sgn) / s.gDens;
  else if (AX == 1) s.Vel = (s.gMY * sgn) / s.gDens;
  else {
    double w = TERRAIN ? (s.gMZ * s.RGv + s.G13 * s.gMX + s.G23 * s.gMY) : s.gMZ;
    s.Vel = (w * sgn) / s.gDens;
  }
}

// Rusanov flux jump of the five variables at one face node (rhot_heve_numflux.F90:1071-1134).
// AX: axis of the (axis-aligned) face normal, sgn = +-1 its sign.
template <int AX, bool TERRAIN>
__device__ __forceinline__ void rusanov_heve(FaceSide& M, FaceSide& Pp, double sgn, double gamm, double hf, double* out5) {
  face_velocity<AX, TERRAIN>(M, sgn);
  face_velocity<AX, TERRAIN>(Pp, sgn);
  double GnnM = 1.0, GnnP = 1.0;
  if (AX == 2 && TERRAIN) {
    GnnM = M.RGv * M.RGv + M.G13 * M.G13 + M.G23 * M.G23;
    GnnP = Pp.RGv * Pp.RGv + Pp.G13 * Pp.G13 + Pp.G23 * Pp.G23;
  }
  double aM = sqrt(GnnM * gamm * (M.Phyd + M.dp) * M.Gs / M.gDens) + fabs(M.Vel);
  double aP = sqrt(GnnP * gamm * (Pp.Phyd + Pp.dp) * Pp.Gs / Pp.gDens) + fabs(Pp.Vel);
  double alpha = fmax(aM, aP);
  out5[V_DDENS] = hf * (Pp.gDens * Pp.Vel - M.gDens * M.Vel - alpha * (Pp.gDD - M.gDD));
  out5[V_DRHOT] = hf * (Pp.gRhot * Pp.Vel - M.gRhot * M.Vel - alpha * (Pp.gDR - M.gDR));
  double t3 = Pp.Gs * Pp.dp, t4 = M.Gs * M.dp;
  double pz = 0.0, px = 0.0, py = 0.0;
  if (AX == 2) {
    pz = (t3 * Pp.RGv - t4 * M.RGv) * sgn;
    if (TERRAIN) { px = (Pp.G13 * sgn) * t3 - (M.G13 * sgn) * t4; py = (Pp.G23 * sgn) * t3 - (M.G23 * sgn) * t4; }
  } else if (AX == 0) {
    px = sgn * t3 - sgn * t4;
  } else {
    py = sgn * t3 - sgn * t4;
  }
  out5[V_MOMZ] = hf * (Pp.gMZ * Pp.Vel - M.gMZ * M.Vel + pz - alpha * (Pp.gMZ - M.gMZ));
  out5[V_MOMX] = hf * (Pp.gMX * Pp.Vel - M.gMX * M.Vel + px - alpha * (Pp.gMX - M.gMX));
  out5[V_MOMY] = hf * (Pp.gMY * Pp.Vel - M.gMY * M.Vel + py - alpha * (Pp.gMY - M.gMY));
}

template <bool TERRAIN, bool MOIST>
__device__ __forceinline__ void load_side(const StageParams& P, size_t n, bool need_pres, double dp_known, FaceSide& s) {
  double Gs = 1.0, G13 = 0.0, G23 = 0.0;
  if (TERRAIN) { Gs = P.gsqrt[n]; G13 = P.g13[n]; G23 = P.g23[n]; }
  s.Gs = Gs; s.RGv = TERRAIN ? 1.0 / Gs : 1.0; s.G13 = G13; s.G23 = G23;
  double dd = P.qin[V_DDENS][n], mx = P.qin[V_MOMX][n], my = P.qin[V_MOMY][n], mz = P.qin[V_MOMZ][n], dr = P.qin[V_DRHOT][n];
  double dh = P.dens_hyd[n], ph = P.pres_hyd[n], th = P.therm_hyd[n];
  s.gDD = Gs * dd; s.gMX = Gs * mx; s.gMY = Gs * my; s.gMZ = Gs * mz; s.gDR = Gs * dr;
  s.gDens = s.gDD + Gs * dh;
  s.gRhot = Gs * th + s.gDR;
  s.Phyd = ph;
  if (need_pres) {
    double R = MOIST ? P.rtot[n] : P.c.Rdry;
    double e = MOIST ? P.cptot[n] / P.cvtot[n] : P.c.CPovCV;
    s.dp = eos_pres(R, P.c.rP0, th + dr, e, P.c.PRES00) - ph;
  } else {
    s.dp = dp_known;
  }
}

// Fused explicit stage of the regional HEVE equations:
//   pressure (a4) -> Rusanov face flux (a5) -> volume divergence + lift (a2) -> tendency (a7)
//   -> RK stage update (a13) [-> modal filter (a15) -> pressure diagnostic, on the last stage]
// reading the stage-input state once and writing the stage-output state once.
template <int NP, int EPB, bool TERRAIN, bool MOIST>
__global__ void __launch_bounds__(NP * NP * EPB, (NP == 8 ? 4 : 4))
heve_stage_kernel(const __grid_constant__ StageParams P) {
  constexpr int N2 = NP * NP, N3 = NP * N2, NFT = 6 * N2, PJ = NP + 1, PL = NP * PJ;
  const int tid = threadIdx.x;
  const int el = tid / N2, t = tid - el * N2, i = t % NP, j = t / NP;
  int ke = blockIdx.x * EPB + el;
  const bool live = ke < P.Ne;
  if (!live) ke = P.Ne - 1;
  const size_t eb = size_t(ke) * N3;

  extern __shared__ double smem[];
  // per element: 2 x (sFx,sFy) ping-pong planes, filter planes, dpres, del_flux
  constexpr int SM_PER_EL = 4 * NP * PL + 2 * NP * PL + N3 + NVAR * NFT;
  double* sm = smem + size_t(el) * SM_PER_EL;
  double* sF = sm;                      // [2][2][NP][PL]
  double* sG = sm + 4 * NP * PL;        // [2][NP][PL]
  double* sDP = sG + 2 * NP * PL;       // [N3]
  double* sDel = sDP + N3;              // [NVAR][NFT]

  // ---- operator rows of this thread
  double Di[NP], Dj[NP];
#pragma unroll
  for (int l = 0; l < NP; ++l) { Di[l] = cT.D[i * NP + l]; Dj[l] = cT.D[j * NP + l]; }
  const double lwi0 = cT.Lw[i * 2], lwi1 = cT.Lw[i * 2 + 1], lwj0 = cT.Lw[j * 2], lwj1 = cT.Lw[j * 2 + 1];

  // ---- 1. load the column, pressure
  double mx[NP], my[NP], mz[NP], rdens[NP], pt[NP], dp[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    size_t n = eb + t + k * N2;
    double dd = P.qin[V_DDENS][n], dr = P.qin[V_DRHOT][n];
    mx[k] = P.qin[V_MOMX][n]; my[k] = P.qin[V_MOMY][n]; mz[k] = P.qin[V_MOMZ][n];
    double dh = P.dens_hyd[n], ph = P.pres_hyd[n], th = P.therm_hyd[n];
    double R = MOIST ? P.rtot[n] : P.c.Rdry;
    double e = MOIST ? P.cptot[n] / P.cvtot[n] : P.c.CPovCV;
    double rhot = th + dr;
    dp[k] = eos_pres(R, P.c.rP0, rhot, e, P.c.PRES00) - ph;
    rdens[k] = 1.0 / (dd + dh);
    pt[k] = rhot * rdens[k];
    sDP[t + k * N2] = dp[k];
  }
  __syncthreads();

  // ---- 2. face flux jumps: thread t handles node t of each of the six faces
  {
    const int a = i, b = j;
    const size_t fb = size_t(ke) * NFT;
    const double gamm = P.c.gamm;
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      int nloc;
      switch (f) {
        case 0: nloc = a + b * N2; break;
        case 1: nloc = (NP - 1) + a * NP + b * N2; break;
        case 2: nloc = a + (NP - 1) * NP + b * N2; break;
        case 3: nloc = a * NP + b * N2; break;
        case 4: nloc = t; break;
        default: nloc = t + (NP - 1) * N2; break;
      }
      const int iP = P.vmapP[fb + f * N2 + t];
      FaceSide M, Q;
      load_side<TERRAIN, MOIST>(P, eb + nloc, false, sDP[nloc], M);
      load_side<TERRAIN, MOIST>(P, size_t(iP), true, 0.0, Q);
      const double hf = P.fscale[size_t(f) * P.Ne + ke] * 0.5;
      double o5[NVAR];
      if (f == 0) rusanov_heve<1, TERRAIN>(M, Q, -1.0, gamm, hf, o5);
      else if (f == 1) rusanov_heve<0, TERRAIN>(M, Q, 1.0, gamm, hf, o5);
      else if (f == 2) rusanov_heve<1, TERRAIN>(M, Q, 1.0, gamm, hf, o5);
      else if (f == 3) rusanov_heve<0, TERRAIN>(M, Q, -1.0, gamm, hf, o5);
      else if (f == 4) rusanov_heve<2, TERRAIN>(M, Q, -1.0, gamm, hf, o5);
      else rusanov_heve<2, TERRAIN>(M, Q, 1.0, gamm, hf, o5);
#pragma unroll
      for (int v = 0; v < NVAR; ++v) sDel[v * NFT + f * N2 + t] = o5[v];
    }
  }
  // (the first __syncthreads of the volume loop orders these writes before the lift reads)

  // ---- 3. volume terms, tendency, stage update, variable by variable
  const double E11 = P.escale[ke], E22 = P.escale[P.Ne + ke], E33 = P.escale[2 * size_t(P.Ne) + ke];
  const int ke2d = P.emap2d[ke];
  const double cor = P.has_cor ? P.coriolis[size_t(ke2d) * N2 + t] : 0.0;
  const double gH = TERRAIN ? P.gsqrtH[size_t(ke2d) * N2 + t] : 1.0;
  double qfin_dr[NP];  // filtered DRHOT of the last stage, for the pressure diagnostic

  const int order[NVAR] = {V_DDENS, V_DRHOT, V_MOMZ, V_MOMX, V_MOMY};
#pragma unroll
  for (int iv = 0; iv < NVAR; ++iv) {
    const int v = order[iv];
    double* bFx = sF + (iv & 1) * 2 * NP * PL;
    double* bFy = bFx + NP * PL;
    double Fz[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      double G = 1.0, RGv = 1.0, G13 = 0.0, G23 = 0.0;
      if (TERRAIN) { size_t n = eb + t + k * N2; G = P.gsqrt[n]; G13 = P.g13[n]; G23 = P.g23[n]; RGv = 1.0 / (G / gH); }
      double fx0 = G * mx[k], fy0 = G * my[k];
      double fz0 = TERRAIN ? G * (mz[k] * RGv + G13 * mx[k] + G23 * my[k]) : mz[k];
      double GP = G * dp[k];
      double Fx, Fy;
      if (v == V_DDENS) { Fx = fx0; Fy = fy0; Fz[k] = fz0; }
      else if (v == V_DRHOT) { Fx = fx0 * pt[k]; Fy = fy0 * pt[k]; Fz[k] = fz0 * pt[k]; }
      else if (v == V_MOMZ) { double w = mz[k] * rdens[k]; Fx = fx0 * w; Fy = fy0 * w; Fz[k] = fz0 * w + GP * RGv; }
      else if (v == V_MOMX) { double u = mx[k] * rdens[k]; Fx = fx0 * u + GP; Fy = fy0 * u; Fz[k] = TERRAIN ? fz0 * u + GP * G13 : fz0 * u; }
      else { double vv = my[k] * rdens[k]; Fx = fx0 * vv; Fy = fy0 * vv + GP; Fz[k] = TERRAIN ? fz0 * vv + GP * G23 : fz0 * vv; }
      bFx[k * PL + j * PJ + i] = Fx;
      bFy[k * PL + j * PJ + i] = Fy;
    }
    __syncthreads();

    double drho[NP];
    if (v == V_MOMZ) {  // VFilterPM1 of DDENS (rhot_heve.F90:442-443)
      double ddc[NP];
#pragma unroll
      for (int k = 0; k < NP; ++k) ddc[k] = P.qin[V_DDENS][eb + t + k * N2];
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        double s = ddc[0] * cT.VP[k * NP];
#pragma unroll
        for (int l = 1; l < NP; ++l) s += ddc[l] * cT.VP[k * NP + l];
        drho[k] = s;
      }
    }

    const double* sD = sDel + v * NFT;
    double qn[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const double* rx = bFx + k * PL + j * PJ;
      const double* ry = bFy + k * PL + i;
      double dx = Di[0] * rx[0], dy = ry[0] * Dj[0], dz = Fz[0] * cT.D[k * NP];
#pragma unroll
      for (int l = 1; l < NP; ++l) {
        dx += Di[l] * rx[l];
        dy += ry[l * PJ] * Dj[l];
        dz += Fz[l] * cT.D[k * NP + l];
      }
      double lift = lwj0 * sD[i + k * NP] + lwi1 * sD[N2 + j + k * NP] + lwj1 * sD[2 * N2 + i + k * NP] +
                    lwi0 * sD[3 * N2 + j + k * NP] + cT.Lw[k * 2] * sD[4 * N2 + t] + cT.Lw[k * 2 + 1] * sD[5 * N2 + t];
      const size_t n = eb + t + k * N2;
      double RGs = TERRAIN ? 1.0 / P.gsqrt[n] : 1.0;
      double div = (E11 * dx + E22 * dy + E33 * dz + lift) * RGs;
      double tend;
      if (v == V_MOMZ) tend = -div - P.c.GRAV * drho[k];
      else if (v == V_MOMX) tend = ((P.has_phyd ? -P.dphydx[n] : 0.0) + cor * my[k]) - div;
      else if (v == V_MOMY) tend = ((P.has_phyd ? -P.dphydy[n] : 0.0) - cor * mx[k]) - div;
      else tend = -div;

      if (P.tend_out[0] != nullptr) {
        if (live) P.tend_out[v][n] = tend;
        qn[k] = 0.0;
        continue;
      }
      double q = (v == V_MOMX) ? mx[k] : (v == V_MOMY) ? my[k] : (v == V_MOMZ) ? mz[k] : P.qin[v][n];
      double base = P.rk.use_q0 ? P.rk.c_q0 * P.q0[v][n] : 0.0;
      if (P.rk.add_vt) base = P.vt[v][n];
      double r = base + P.rk.c_q * q + P.rk.c_k * tend;
      if (P.rk.vt_update) {
        double vb = P.rk.vt_init ? P.rk.vt_init_q * q : P.vt[v][n];
        if (live) P.vt[v][n] = vb + P.rk.vt_q * q + P.rk.vt_k * tend;
      }
      qn[k] = r;
    }
    if (P.tend_out[0] != nullptr) continue;

    if (P.do_filter) {  // modal filter of Gsqrt-weighted variables (dyn_dgm_modalfilter.F90:49-130)
      double* gX = sG;
      double* gY = sG + NP * PL;
      double Fi[NP], Fj[NP];
#pragma unroll
      for (int l = 0; l < NP; ++l) { Fi[l] = cT.Fh[i * NP + l]; Fj[l] = cT.Fh[j * NP + l]; }
      double Gk[NP];
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        Gk[k] = TERRAIN ? P.gsqrt[eb + t + k * N2] : 1.0;
        gX[k * PL + j * PJ + i] = Gk[k] * qn[k];
      }
      __syncthreads();
      double r1[NP];
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        const double* rx = gX + k * PL + j * PJ;
        double s = Fi[0] * rx[0];
#pragma unroll
        for (int l = 1; l < NP; ++l) s += Fi[l] * rx[l];
        r1[k] = s;
        gY[k * PL + j * PJ + i] = s;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        const double* ry = gY + k * PL + i;
        double s = ry[0] * Fj[0];
#pragma unroll
        for (int l = 1; l < NP; ++l) s += ry[l * PJ] * Fj[l];
        r1[k] = s;
      }
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        double s = r1[0] * cT.Fv[k * NP];
#pragma unroll
        for (int l = 1; l < NP; ++l) s += r1[l] * cT.Fv[k * NP + l];
        qn[k] = s * (1.0 / Gk[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      if (live) P.qout[v][eb + t + k * N2] = qn[k];
      if (v == V_DRHOT) qfin_dr[k] = qn[k];
    }
  }

  if (P.write_pres && P.tend_out[0] == nullptr) {  // calc_pressure at the end of Update (driver_nonhydro3d.F90:954-959)
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      size_t n = eb + t + k * N2;
      double ph = P.pres_hyd[n], th = P.therm_hyd[n];
      double R = MOIST ? P.rtot[n] : P.c.Rdry;
      double e = MOIST ? P.cptot[n] / P.cvtot[n] : P.c.CPovCV;
      double pr = eos_pres(R, P.c.rP0, th + qfin_dr[k], e, P.c.PRES00);
      if (live) { P.pres_out[n] = pr; P.dpres_out[n] = pr - ph; }
    }
  }
}

template <int NP, int EPB>
static void launch_stage_np(const StageParams& p, bool terrain, bool moist, cudaStream_t s) {
  constexpr int N2 = NP * NP, N3 = NP * N2, NFT = 6 * N2, PL = NP * (NP + 1);
  constexpr int SM_PER_EL = 4 * NP * PL + 2 * NP * PL + N3 + NVAR * NFT;
  size_t shmem = size_t(EPB) * SM_PER_EL * sizeof(double);
  dim3 grid((p.Ne + EPB - 1) / EPB), block(N2 * EPB);
#define FEDG_LAUNCH(T, M)                                                                                    \
  do {                                                                                                       \
    static bool attr_set = false;                                                                            \
    if (!attr_set) {                                                                                         \
      cudaFuncSetAttribute(heve_stage_kernel<NP, EPB, T, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem); \
      attr_set = true;                                                                                       \
    }                                                                                                        \
    heve_stage_kernel<NP, EPB, T, M><<<grid, block, shmem, s>>>(p);                                         \
  } while (0)
  if (terrain) { if (moist) FEDG_LAUNCH(true, true); else FEDG_LAUNCH(true, false); }
  else { if (moist) FEDG_LAUNCH(false, true); else FEDG_LAUNCH(false, false); }
#undef FEDG_LAUNCH
}

void launch_heve_stage(const StageParams& p, int np, bool terrain, bool moist, cudaStream_t s) {
  switch (np) {
    case 8: launch_stage_np<8, 1>(p, terrain, moist, s); break;
    case 4: launch_stage_np<4, 8>(p, terrain, moist, s); break;
    default: break;  // validated at fedg_dyn_init
  }
}

// ---------------------------------------------------------------------------------------------
// Halo fill of the five prognostic variables for faces whose neighbour lives on the same rank
// (MeshFieldComm Put/Exchange/Get collapsed into one gather), fused with ApplyBC_PROGVARS_lc
// (scale_atm_dyn_dgm_bnd.F90:270-367) for faces on a physical boundary.
__global__ void halo_fill_kernel(const __grid_constant__ HaloParams H) {
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H.Nhalo) return;
  int src = H.src[h];
  if (src < 0) return;  // filled by the NCCL unpack kernel
  int f = 0;
  while (h >= H.face_off[f + 1]) ++f;
  const size_t dst = size_t(H.Np) * H.Ne + h;
  double dd = H.q[V_DDENS][src], mx = H.q[V_MOMX][src], my = H.q[V_MOMY][src], mz = H.q[V_MOMZ][src], dr = H.q[V_DRHOT][src];
  const int bc = H.bc[f];
  if (bc == FEDG_BND_SLIP) {
    // src == vmapB[h] on a physical boundary
    double nx = (f == 1) ? 1.0 : (f == 3) ? -1.0 : 0.0;
    double ny = (f == 2) ? 1.0 : (f == 0) ? -1.0 : 0.0;
    double nz = (f == 5) ? 1.0 : (f == 4) ? -1.0 : 0.0;
    double GsqrtV = 1.0, G13 = 0.0, G23 = 0.0;
    if (H.terrain) {
      int ke = src / H.Np, p = src - ke * H.Np;
      GsqrtV = H.gsqrt[src] / H.gsqrtH[size_t(H.emap2d[ke]) * H.Nfp + (p % H.Nfp)];
      G13 = H.g13[src]; G23 = H.g23[src];
    }
    double momw = mz / GsqrtV + G13 * mx + G23 * my;
    double fac = nz * GsqrtV * GsqrtV / (1.0 + (GsqrtV * G13) * (GsqrtV * G13) + (GsqrtV * G23) * (GsqrtV * G23));
    double mn = mx * nx + my * ny + momw * nz;
    double mxP = mx - 2.0 * mn * (nx + fac * G13);
    double myP = my - 2.0 * mn * (ny + fac * G23);
    double mzP = mz - 2.0 * mn * fac / GsqrtV;
    mx = mxP; my = myP; mz = mzP;
  } else if (bc == FEDG_BND_NOSLIP) {
    mx = -mx; my = -my; mz = -mz;
  }
  H.q[V_DDENS][dst] = dd; H.q[V_MOMX][dst] = mx; H.q[V_MOMY][dst] = my; H.q[V_MOMZ][dst] = mz; H.q[V_DRHOT][dst] = dr;
}

void launch_halo_fill(const HaloParams& p, cudaStream_t s) {
  if (p.Nhalo <= 0) return;
  int block = 256, grid = (p.Nhalo + block - 1) / block;
  halo_fill_kernel<<<grid, block, 0, s>>>(p);
}

// ---------------------------------------------------------------------------------------------
// atm_dyn_dgm_nonhydro3d_common_DRHOT2PRES stand-alone (nonhydro3d_common.F90:428-479)
__global__ void calc_pres_kernel(const double* __restrict__ drhot, const double* __restrict__ pres_hyd,
                                 const double* __restrict__ therm_hyd, const double* __restrict__ rtot,
                                 const double* __restrict__ cvtot, const double* __restrict__ cptot, int moist, PhysConst c,
                                 double* __restrict__ pres, double* __restrict__ dpres, long n) {
  long idx = long(blockIdx.x) * blockDim.x + threadIdx.x;
  long stride = long(gridDim.x) * blockDim.x;
  for (; idx < n; idx += stride) {
    double R = moist ? rtot[idx] : c.Rdry;
    double e = moist ? cptot[idx] / cvtot[idx] : c.CPovCV;
    double pr = eos_pres(R, c.rP0, therm_hyd[idx] + drhot[idx], e, c.PRES00);
    pres[idx] = pr;
    dpres[idx] = pr - pres_hyd[idx];
  }
}
void launch_calc_pres(const double* drhot, const double* pres_hyd, const double* therm_hyd, const double* rtot,
                      const double* cvtot, const double* cptot, bool moist, PhysConst c, double* pres, double* dpres,
                      long n, cudaStream_t s) {
  int block = 256;
  long g = (n + block - 1) / block;
  int grid = int(g < 148L * 16 ? g : 148L * 16);
  calc_pres_kernel<<<grid, block, 0, s>>>(drhot, pres_hyd, therm_hyd, rtot, cvtot, cptot, moist ? 1 : 0, c, pres, dpres, n);
}

// atm_dyn_dgm_nonhydro3d_common_calc_RHOT_hyd (nonhydro3d_common.F90:584-619)
__global__ void calc_rhot_hyd_kernel(const double* __restrict__ pres_hyd, PhysConst c, double* __restrict__ out, long n) {
  long idx = long(blockIdx.x) * blockDim.x + threadIdx.x;
  long stride = long(gridDim.x) * blockDim.x;
  for (; idx < n; idx += stride) out[idx] = c.PRES00 / c.Rdry * pow(pres_hyd[idx] / c.PRES00, c.CVdry / c.CPdry);
}
void launch_calc_rhot_hyd(const double* pres_hyd, PhysConst c, double* therm_hyd, long n, cudaStream_t s) {
  int block = 256;
  long g = (n + block - 1) / block;
  int grid = int(g < 148L * 16 ? g : 148L * 16);
  calc_rhot_hyd_kernel<<<grid, block, 0, s>>>(pres_hyd, c, therm_hyd, n);
}

// ---------------------------------------------------------------------------------------------
// Conservation monitors: block partial sums in a fixed order, then a single-block final pass
// (deterministic, no atomics).  file/scale_file_monitor_meshfield.F90:176-213.
__global__ void monitor_partial_kernel(const double* dd, const double* mx, const double* my, const double* mz,
                                       const double* dens_hyd, const double* pres, const double* rtot, int moist,
                                       const double* w3, const double* Jac, const double* gsqrt, int terrain,
                                       const double* zlev, PhysConst c, int Np, int Ne, double* partial) {
  __shared__ double sh[5][256];
  double acc[5] = {0, 0, 0, 0, 0};
  const int ke = blockIdx.x;
  for (int p = threadIdx.x; p < Np; p += blockDim.x) {
    size_t n = size_t(ke) * Np + p;
    double w = w3[p] * Jac[n] * (terrain ? gsqrt[n] : 1.0);
    double dens = dd[n] + dens_hyd[n];
    double engk = 0.5 * (mx[n] * mx[n] + my[n] * my[n] + mz[n] * mz[n]) / dens;
    double engi = pres[n] / (moist ? rtot[n] : c.Rdry) * c.CVdry;
    double engp = dens * c.GRAV * zlev[n];
    acc[0] += w * dd[n]; acc[1] += w * (engk + engi + engp); acc[2] += w * engk; acc[3] += w * engi; acc[4] += w * engp;
  }
  for (int m = 0; m < 5; ++m) sh[m][threadIdx.x] = acc[m];
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) for (int m = 0; m < 5; ++m) sh[m][threadIdx.x] += sh[m][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < 5) partial[size_t(threadIdx.x) * Ne + ke] = sh[threadIdx.x][0];
}
__global__ void monitor_final_kernel(const double* partial, int Ne, double* out5) {
  __shared__ double sh[256];
  const int m = blockIdx.x;
  double a = 0.0;
  for (int e = threadIdx.x; e < Ne; e += blockDim.x) a += partial[size_t(m) * Ne + e];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out5[m] = sh[0];
}
static double* g_partial = nullptr;
static size_t g_partial_n = 0;
void launch_monitor(const double* const q[NVAR], const double* dens_hyd, const double* pres, const double* rtot,
                    bool moist, const double* w3, const double* Jac, const double* gsqrt, bool terrain,
                    const double* zlev, PhysConst c, int Np, int Ne, double* out5, cudaStream_t s) {
  if (g_partial_n < size_t(5) * Ne) {
    if (g_partial) cudaFree(g_partial);
    cudaMalloc(&g_partial, size_t(5) * Ne * sizeof(double));
    g_partial_n = size_t(5) * Ne;
  }
  monitor_partial_kernel<<<Ne, 256, 0, s>>>(q[V_DDENS], q[V_MOMX], q[V_MOMY], q[V_MOMZ], dens_hyd, pres, rtot, moist ? 1 : 0,
                                            w3, Jac, gsqrt, terrain ? 1 : 0, zlev, c, Np, Ne, g_partial);
  monitor_final_kernel<<<5, 256, 0, s>>>(g_partial, Ne, out5);
}

// ---------------------------------------------------------------------------------------------
// ElementOperationBase3D conformance kernels: one block per element, one thread per node.
// op: 0 Dx, 1 Dy, 2 Dz, 3 Lift, 4 VFilterPM1, 5 ModalFilter
__global__ void elem_op_kernel(int op, const double* __restrict__ in, double* __restrict__ out, int np) {
  extern __shared__ double s[];
  const int N2 = np * np, N3 = N2 * np, NFT = 6 * N2;
  const int e = blockIdx.x, n = threadIdx.x;
  const int i = n % np, j = (n / np) % np, k = n / N2;
  const int nin = (op == 3) ? NFT : N3;
  for (int m = n; m < nin; m += blockDim.x) s[m] = in[size_t(e) * nin + m];
  __syncthreads();
  double r = 0.0;
  if (op == 0) { r = cT.D[i * np] * s[j * np + k * N2]; for (int l = 1; l < np; ++l) r += cT.D[i * np + l] * s[l + j * np + k * N2]; }
  else if (op == 1) { r = s[i + k * N2] * cT.D[j * np]; for (int l = 1; l < np; ++l) r += s[i + l * np + k * N2] * cT.D[j * np + l]; }
  else if (op == 2) { r = s[i + j * np] * cT.D[k * np]; for (int l = 1; l < np; ++l) r += s[i + j * np + l * N2] * cT.D[k * np + l]; }
  else if (op == 4) { r = s[i + j * np] * cT.VP[k * np]; for (int l = 1; l < np; ++l) r += s[i + j * np + l * N2] * cT.VP[k * np + l]; }
  else if (op == 3) {
    r = cT.Lw[j * 2] * s[i + k * np] + cT.Lw[i * 2 + 1] * s[N2 + j + k * np] + cT.Lw[j * 2 + 1] * s[2 * N2 + i + k * np] +
        cT.Lw[i * 2] * s[3 * N2 + j + k * np] + cT.Lw[k * 2] * s[4 * N2 + i + j * np] + cT.Lw[k * 2 + 1] * s[5 * N2 + i + j * np];
  } else {  // modal filter: x, y, z passes
    double* w = s + N3;
    double a = cT.Fh[i * np] * s[j * np + k * N2];
    for (int l = 1; l < np; ++l) a += cT.Fh[i * np + l] * s[l + j * np + k * N2];
    w[n] = a;
    __syncthreads();
    double b = w[i + k * N2] * cT.Fh[j * np];
    for (int l = 1; l < np; ++l) b += w[i + l * np + k * N2] * cT.Fh[j * np + l];
    __syncthreads();
    s[n] = b;
    __syncthreads();
    r = s[i + j * np] * cT.Fv[k * np];
    for (int l = 1; l < np; ++l) r += s[i + j * np + l * N2] * cT.Fv[k * np + l];
  }
  out[size_t(e) * N3 + n] = r;
}
void launch_elem_op(int op, const double* in, double* out, int nelem, int np, cudaStream_t s) {
  int N3 = np * np * np;
  size_t sh = size_t(2) * (6 * np * np > N3 ? 6 * np * np : N3) * sizeof(double);
  elem_op_kernel<<<nelem, N3, sh, s>>>(op, in, out, np);
}

}  // namespace fedg
