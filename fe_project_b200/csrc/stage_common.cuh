// Device helpers shared by the stage kernels: equation of state, TMA bulk copy + mbarrier, Rusanov flux.
#pragma once
#include <cstdint>

#include "fedg_internal.h"

namespace fedg {

// PRES = P00 * (Rtot/P00 * RHOT)^(CPtot/CVtot)   (nonhydro3d_common.F90:467-474)
__device__ __forceinline__ double eos_pres(double R, double rP0, double rhot, double cpovcv, double P00) {
  return P00 * pow(R * rP0 * rhot, cpovcv);
}

// The same with x^e evaluated as exp(e log x) while x = Rtot RHOT / P00 lies in [0.25, 2] (|log x| <= 1.39; the atmosphere up to
// ~60 km): 64 % of the instructions of pow().  Measured against a 64-bit-mantissa reference over that range the relative error
// is <= 3.5e-16 (1.6 ulp; pow itself: CUDA documents 2 ulp), outside the range and with exact != 0 (FEDG_EXACT_POW=1) pow() is used.
__device__ __forceinline__ double eos_pres_fast(double R, double rP0, double rhot, double cpovcv, double P00, int exact) {
  const double x = R * rP0 * rhot;
  if (!exact && x > 0.25 && x < 2.0) return P00 * exp(cpovcv * log(x));
  return P00 * pow(x, cpovcv);
}

// ---- TMA bulk copy + mbarrier (PTX ISA: cp.async.bulk, mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// L2 prefetch of a contiguous global range (no destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
// all lanes wait (each observes the phase completion itself: acquire of the async-proxy writes); back off
// between polls so that waiting warps do not eat issue slots of the other resident blocks
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(64);
}

// One side of a face node: Gsqrt-weighted state of rhot_heve_numflux.F90:1030-1069.
struct FaceSide {
  double gDD, gMX, gMY, gMZ, gDR, gDens, rgDens, gRhot, Gs, RGv, G13, G23, Phyd, dp, Vel, Velh;
};

template <bool TERRAIN>
__device__ __forceinline__ void make_side(FaceSide& s, double dd, double mx, double my, double mz, double dr, double dh, double ph,
                                          double th, double dp, double Gs, double G13, double G23) {
  s.Gs = Gs; s.RGv = TERRAIN ? 1.0 / Gs : 1.0; s.G13 = G13; s.G23 = G23;
  s.gDD = Gs * dd; s.gMX = Gs * mx; s.gMY = Gs * my; s.gMZ = Gs * mz; s.gDR = Gs * dr;
  s.gDens = s.gDD + Gs * dh;
  s.rgDens = 1.0 / s.gDens;   // one reciprocal per side instead of the reference's three divisions (<= 1 ulp apart)
  s.gRhot = Gs * th + s.gDR;
  s.Phyd = ph; s.dp = dp;
}

// raw exterior-side node values gathered from global memory through VMapP
template <bool TERRAIN>
struct RawSide {
  double dd, mx, my, mz, dr, dh, ph, th, dp, Gs, G13, G23;
  __device__ __forceinline__ void load(const StageParams& P, size_t iP) {
    dd = P.qin[V_DDENS][iP]; mx = P.qin[V_MOMX][iP]; my = P.qin[V_MOMY][iP]; mz = P.qin[V_MOMZ][iP]; dr = P.qin[V_DRHOT][iP];
    dh = P.dens_hyd[iP]; ph = P.pres_hyd[iP]; th = P.therm_hyd[iP]; dp = P.dpin[iP];
    Gs = 1.0; G13 = 0.0; G23 = 0.0;
    if (TERRAIN) { Gs = P.gsqrt[iP]; G13 = P.g13[iP]; G23 = P.g23[iP]; }
  }
};

// contravariant normal velocity; Velh = horizontal part only (HEVI mass / theta fluxes)
template <int AX, bool TERRAIN>
__device__ __forceinline__ void face_velocity(FaceSide& s, double sgn) {
  if (AX == 0) { s.Velh = (s.gMX * sgn) * s.rgDens; s.Vel = s.Velh; }
  else if (AX == 1) { s.Velh = (s.gMY * sgn) * s.rgDens; s.Vel = s.Velh; }
  else {
    double w = TERRAIN ? (s.gMZ * s.RGv + s.G13 * s.gMX + s.G23 * s.gMY) : s.gMZ;
    s.Velh = 0.0;
    s.Vel = s.Velh + (w * sgn) * s.rgDens;
  }
}

// Rusanov flux jump of the five variables at one face node.
//   HEVE: rhot_heve_numflux.F90:1071-1134.   HEVI: rhot_hevi_numflux.F90:363-411 (alpha *= 1 - nz^2; mass and
//   theta fluxes advect with the horizontal velocity only; no vertical pressure term in MOMZ).
template <int AX, bool TERRAIN, bool HEVI>
__device__ __forceinline__ void rusanov(FaceSide& M, FaceSide& Q, double sgn, double gamm, double hf, double* out5) {
  face_velocity<AX, TERRAIN>(M, sgn);
  face_velocity<AX, TERRAIN>(Q, sgn);
  double alpha;
  if (HEVI) {
    if (AX == 2) alpha = 0.0;
    else alpha = fmax(sqrt(gamm * (M.Phyd + M.dp) * M.Gs * M.rgDens) + fabs(M.Vel), sqrt(gamm * (Q.Phyd + Q.dp) * Q.Gs * Q.rgDens) + fabs(Q.Vel));
  } else {
    double GnnM = 1.0, GnnP = 1.0;
    if (AX == 2 && TERRAIN) {
      GnnM = M.RGv * M.RGv + M.G13 * M.G13 + M.G23 * M.G23;
      GnnP = Q.RGv * Q.RGv + Q.G13 * Q.G13 + Q.G23 * Q.G23;
    }
    alpha = fmax(sqrt(GnnM * gamm * (M.Phyd + M.dp) * M.Gs * M.rgDens) + fabs(M.Vel),
                 sqrt(GnnP * gamm * (Q.Phyd + Q.dp) * Q.Gs * Q.rgDens) + fabs(Q.Vel));
  }
  const double vM = HEVI ? M.Velh : M.Vel, vQ = HEVI ? Q.Velh : Q.Vel;
  out5[V_DDENS] = hf * (Q.gDens * vQ - M.gDens * vM - alpha * (Q.gDD - M.gDD));
  out5[V_DRHOT] = hf * (Q.gRhot * vQ - M.gRhot * vM - alpha * (Q.gDR - M.gDR));
  const double t3 = Q.Gs * Q.dp, t4 = M.Gs * M.dp;
  double pz = 0.0, px = 0.0, py = 0.0;
  if (AX == 2) {
    if (!HEVI) pz = (t3 * Q.RGv - t4 * M.RGv) * sgn;
    if (TERRAIN) { px = (Q.G13 * sgn) * t3 - (M.G13 * sgn) * t4; py = (Q.G23 * sgn) * t3 - (M.G23 * sgn) * t4; }
  } else if (AX == 0) {
    px = sgn * t3 - sgn * t4;
  } else {
    py = sgn * t3 - sgn * t4;
  }
  out5[V_MOMZ] = hf * (Q.gMZ * Q.Vel - M.gMZ * M.Vel + pz - alpha * (Q.gMZ - M.gMZ));
  out5[V_MOMX] = hf * (Q.gMX * Q.Vel - M.gMX * M.Vel + px - alpha * (Q.gMX - M.gMX));
  out5[V_MOMY] = hf * (Q.gMY * Q.Vel - M.gMY * M.Vel + py - alpha * (Q.gMY - M.gMY));
}

// Global (cubed-sphere) flux jump, specialised to the shallow atmosphere without topography (gam = 1, GsqrtV = 1,
// G13 = G23 = 0; checked at fedg_create): Gs is GsqrtH on both sides, G11/G12/G22 are the contravariant metric of the own
// face node (the reference uses the own element's values for both sides).
//   HEVI: rhot_hevi_numflux.F90:606-834 (numflux_get_generalhvc): alpha *= 1 - nz^2, mass / theta advect with the horizontal
//         velocity, no vertical pressure term.   HEVE: rhot_heve_numflux.F90:1543-1772.
template <int AX, bool HEVI>
__device__ __forceinline__ void rusanov_global(FaceSide& M, FaceSide& Q, double sgn, double gamm, double hf, double G11, double G12,
                                               double G22, double* out5) {
  face_velocity<AX, false>(M, sgn);
  face_velocity<AX, false>(Q, sgn);
  double alpha = 0.0, G1n = 0.0, G2n = 0.0;
  if (AX != 2 || !HEVI) {
    const double Gnn = (AX == 0) ? fabs(G11 * sgn) : (AX == 1) ? fabs(G22 * sgn) : 1.0;
    alpha = fmax(sqrt(Gnn * gamm * (M.Phyd + M.dp) * M.Gs * M.rgDens) + fabs(M.Vel),
                 sqrt(Gnn * gamm * (Q.Phyd + Q.dp) * Q.Gs * Q.rgDens) + fabs(Q.Vel));
  }
  if (AX != 2) {
    G1n = (AX == 0) ? G11 * sgn : G12 * sgn;
    G2n = (AX == 0) ? G12 * sgn : G22 * sgn;
  }
  const double vM = HEVI ? M.Velh : M.Vel, vQ = HEVI ? Q.Velh : Q.Vel;
  out5[V_DDENS] = hf * (Q.gDens * vQ - M.gDens * vM - alpha * (Q.gDD - M.gDD));
  out5[V_DRHOT] = hf * (Q.gRhot * vQ - M.gRhot * vM - alpha * (Q.gDR - M.gDR));
  const double t3 = Q.Gs * Q.dp, t4 = M.Gs * M.dp;
  const double pz = (AX == 2 && !HEVI) ? (t3 - t4) * sgn : 0.0;
  out5[V_MOMZ] = hf * (Q.gMZ * Q.Vel - M.gMZ * M.Vel + pz - alpha * (Q.gMZ - M.gMZ));
  out5[V_MOMX] = hf * (Q.gMX * Q.Vel - M.gMX * M.Vel + (G1n * t3 - G1n * t4) - alpha * (Q.gMX - M.gMX));
  out5[V_MOMY] = hf * (Q.gMY * Q.Vel - M.gMY * M.Vel + (G2n * t3 - G2n * t4) - alpha * (Q.gMY - M.gMY));
}

}  // namespace fedg
