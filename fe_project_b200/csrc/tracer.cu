// DG tracer advection with a prescribed mass flux on the GPU (the ONLY_TRACERADV_FLAG mode of
// AtmDynDGMDriver_trcadv3d_update, fluid_dyn_solver/scale_atm_dyn_dgm_driver_trcadv3d.F90:312-559) for the flat regional mesh.
//
// Parity with the CPU restatement: tests/test_gpu_tracer.py (green on hardware since the round-1 driver run).  Round 2 adds the COUPLED
// mode (driver_trcadv3d.F90:426-539 after a dynamics step): the mass flux and the dissipation coefficient are the stage averages the
// dynamics saved (trc_massflux_accum_kernel / trc_alphdens_dyn_kernel = atm_dyn_dgm_trcadvect3d_save_massflux :343-401 and
// ..._cal_alphdens_dyn :460-551, called from the stage loop at driver_nonhydro3d.F90:900-917), DDENS0_TRC / DDENS_TRC the density at the
// start of the step and after the RK loop.
//
//   trc_alphdens_kernel   atm_dyn_dgm_trcadvect3d_heve_cal_alphdens_advtest   trcadvect3d_heve.F90:404-455
//   trc_fct_kernel        ..._calc_fct_coef + get_netOutwardFlux_generalhvc   :234-306, :678-777
//   trc_stage_kernel      ..._cal_tend + get_delflux_generalhvc               :149-231, :554-674
//                         + rk_advance_trcvar_low_storage2D (common/scale_timeint_rk.F90)
//                         + atm_dyn_dgm_tracer_modalfilter_apply (modalfilter.F90:232-270) and ..._TMAR (:311-340) at the last stage
//
// One block per element, one thread per node (Np = 64 or 512).  The face sums that decide a sign (outward flux per face) are
// accumulated by one thread per face in ascending face-node order, as the reference's sparse product does.
#include "fedg_internal.h"

namespace fedg {

namespace {

__device__ __forceinline__ int trc_face_node(int f, int fp, int np) {
  const int a = fp % np, b = fp / np, n2 = np * np;
  switch (f) {
    case 0: return a + b * n2;
    case 1: return (np - 1) + a * np + b * n2;
    case 2: return a + (np - 1) * np + b * n2;
    case 3: return a * np + b * n2;
    case 4: return fp;
    default: return fp + (np - 1) * n2;
  }
}
__device__ __forceinline__ void trc_normal(int f, double& nx, double& ny, double& nz) {
  nx = (f == 1) ? 1.0 : (f == 3) ? -1.0 : 0.0;
  ny = (f == 2) ? 1.0 : (f == 0) ? -1.0 : 0.0;
  nz = (f == 5) ? 1.0 : (f == 4) ? -1.0 : 0.0;
}

// alphDens_M / alphDens_P of every face node: alpha = max |V.n| of the two sides, times the density of each side (Gsqrt = 1)
__global__ void trc_alphdens_kernel(const __grid_constant__ TracerParams P) {
  const size_t g = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= size_t(P.NfpTot) * P.Ne) return;
  const int ke = int(g / P.NfpTot), m = int(g - size_t(ke) * P.NfpTot), f = m / P.Nfp, fp = m - f * P.Nfp;
  const size_t iM = size_t(ke) * P.Np + trc_face_node(f, fp, P.np), iP = size_t(P.vmapP[g]);
  double nx, ny, nz;
  trc_normal(f, nx, ny, nz);
  const double densM = P.ddens[iM] + P.dens_hyd[iM], densP = P.ddens[iP] + P.dens_hyd[iP];
  const double VelM = (P.mfx[iM] * nx + P.mfy[iM] * ny + P.mfz[iM] * nz) / densM;
  const double VelP = (P.mfx[iP] * nx + P.mfy[iP] * ny + P.mfz[iP] * nz) / densP;
  const double alpha = fmax(fabs(VelM), fabs(VelP));
  P.alphM[g] = alpha * densM;
  P.alphP[g] = alpha * densP;
}

// numerical flux of one face node (get_delflux_generalhvc :620-655, flat mesh: Gsqrt = GsqrtV = 1, G13 = G23 = 0)
struct TrcFace { double QM, FM, num; };
__device__ __forceinline__ TrcFace trc_face(const TracerParams& P, int ke, int m) {
  const int f = m / P.Nfp, fp = m - f * P.Nfp;
  const size_t g = size_t(ke) * P.NfpTot + m;
  const size_t iM = size_t(ke) * P.Np + trc_face_node(f, fp, P.np), iP = size_t(P.vmapP[g]);
  double nx, ny, nz;
  trc_normal(f, nx, ny, nz);
  const double FM = P.mfx[iM] * nx + P.mfy[iM] * ny + P.mfz[iM] * nz;
  const double FP = P.mfx[iP] * nx + P.mfy[iP] * ny + P.mfz[iP] * nz;
  const double QM = P.q[iM], QP = P.q[iP];
  TrcFace r;
  r.QM = QM; r.FM = FM;
  r.num = 0.5 * ((QP * FP + QM * FM) - P.alphP[g] * QP + P.alphM[g] * QM);
  return r;
}
// sparsemat_matmul(FaceIntMat, J(iM) * Fscale * numflux) for face f: ascending face-node order
__device__ __forceinline__ double trc_outward(const TracerParams& P, int ke, int f, const double* sNum) {
  const double fs = P.fscale[size_t(f) * P.Ne + ke];
  double s = 0.0;
  for (int fp = 0; fp < P.Nfp; ++fp) {
    const int a = fp % P.np, b = fp / P.np;
    const size_t iM = size_t(ke) * P.Np + trc_face_node(f, fp, P.np);
    s += (P.w1d[a] * P.w1d[b]) * (P.jac[iM] * fs * sNum[f * P.Nfp + fp]);
  }
  return s;
}
// block-wide sum, the same value in every thread (tree over shared memory; blockDim.x is a power of two: 64 or 512)
__device__ __forceinline__ double trc_block_sum(double v, double* sRed) {
  const int n = threadIdx.x;
  __syncthreads();
  sRed[n] = v;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (n < s) sRed[n] = sRed[n] + sRed[n + s];
    __syncthreads();
  }
  const double r = sRed[0];
  __syncthreads();
  return r;
}

// dynamic shared memory: sNum[NfpTot] sDel[NfpTot] sQF[NfpTot] sOut[8] sF[3*Np] sRed[Np] sTab[2*np*np + 2*np]
__global__ void trc_fct_kernel(const __grid_constant__ TracerParams P) {
  extern __shared__ __align__(16) double sm[];
  double* sNum = sm;
  double* sOut = sm + 3 * P.NfpTot;
  double* sRed = sOut + 8 + 3 * P.Np;
  const int ke = blockIdx.x, n = threadIdx.x;
  const size_t gn = size_t(ke) * P.Np + n;
  if (P.disable_limiter) { P.fct[gn] = 1.0; return; }
  for (int m = n; m < P.NfpTot; m += P.Np) sNum[m] = trc_face(P, ke, m).num;
  __syncthreads();
  if (n < 6) sOut[n] = trc_outward(P, ke, n, sNum);
  __syncthreads();
  double net = 0.0;
  for (int f = 0; f < 6; ++f) net += fmax(0.0, sOut[f]);
  const double dens_ssm1 = P.dens_hyd[gn] + (1.0 - P.c_ssm1) * P.ddens0[gn] + P.c_ssm1 * P.ddens[gn];
  const double tp = P.rhoq_tp ? P.rhoq_tp[gn] : 0.0;
  const double Qs = trc_block_sum(P.jac[gn] * P.w3[n] * (dens_ssm1 * P.q[gn] / P.dttmp + tp), sRed);
  P.fct[gn] = fmax(0.0, fmin(1.0, Qs / (net + 1.0e-10)));
}

__global__ void trc_stage_kernel(const __grid_constant__ TracerParams P) {
  extern __shared__ __align__(16) double sm[];
  const int Np = P.Np, NfpTot = P.NfpTot, Nfp = P.Nfp, np = P.np;
  double* sNum = sm;
  double* sDel = sm + NfpTot;
  double* sQF = sm + 2 * NfpTot;
  double* sOut = sm + 3 * NfpTot;
  double* sF = sOut + 8;            // [3][Np]; reused by the filter passes
  double* sRed = sF + 3 * Np;
  double* sD = sRed + Np;           // D1D[i][l]
  double* sFh = sD + np * np;       // tracer modal filter, horizontal [i][l]
  double* sFv = sFh + np * np;      // vertical [k][l]
  double* sLw = sFv + np * np;      // lift1d[m][side]
  const int ke = blockIdx.x, n = threadIdx.x;
  const int i = n % np, j = (n / np) % np, k = n / (np * np);
  const size_t gn = size_t(ke) * Np + n;
  for (int m = n; m < np * np; m += Np) { sD[m] = P.tab->D[m]; sFh[m] = P.filt[m]; sFv[m] = P.filt[np * np + m]; }
  if (n < 2 * np) sLw[n] = P.tab->Lw[n];
  const double q = P.q[gn];
  sF[n] = P.mfx[gn] * q;
  sF[Np + n] = P.mfy[gn] * q;
  sF[2 * Np + n] = P.mfz[gn] * q;
  for (int m = n; m < NfpTot; m += Np) {
    const TrcFace t = trc_face(P, ke, m);
    sNum[m] = t.num;
    sQF[m] = t.QM * t.FM;
  }
  __syncthreads();
  if (n < 6) sOut[n] = trc_outward(P, ke, n, sNum);
  __syncthreads();
  for (int m = n; m < NfpTot; m += Np) {
    const int f = m / Nfp, fp = m - f * Nfp;
    const size_t g = size_t(ke) * NfpTot + m;
    const double RM = P.fct[size_t(ke) * Np + trc_face_node(f, fp, np)], RP = P.fct[P.vmapP[g]];
    const double sgn = copysign(1.0, sOut[f]);
    sDel[m] = P.fscale[size_t(f) * P.Ne + ke] * (sNum[m] * 0.5 * (RP + RM - (RP - RM) * sgn) - sQF[m]);
  }
  __syncthreads();
  // Div (tensorprod3D Div: three 1D derivative products + lift), then cal_tend :218-224
  double dx = 0.0, dy = 0.0, dz = 0.0;
  for (int l = 0; l < np; ++l) {
    dx += sD[i * np + l] * sF[l + j * np + k * np * np];
    dy += sD[j * np + l] * sF[Np + i + l * np + k * np * np];
    dz += sD[k * np + l] * sF[2 * Np + i + j * np + l * np * np];
  }
  const double lift = sLw[j * 2] * sDel[i + k * np] + sLw[i * 2 + 1] * sDel[Nfp + j + k * np] + sLw[j * 2 + 1] * sDel[2 * Nfp + i + k * np] +
                      sLw[i * 2] * sDel[3 * Nfp + j + k * np] + sLw[k * 2] * sDel[4 * Nfp + i + j * np] + sLw[k * 2 + 1] * sDel[5 * Nfp + i + j * np];
  const double E11 = P.escale[ke], E22 = P.escale[P.Ne + ke], E33 = P.escale[2 * size_t(P.Ne) + ke];
  const double tend = -(E11 * dx + E22 * dy + E33 * dz + lift) + (P.rhoq_tp ? P.rhoq_tp[gn] : 0.0);

  // rk_advance_trcvar_low_storage2D: rho*q is advanced, q recovered with the density interpolated to the stage time
  const double dh = P.dens_hyd[gn], d0 = P.ddens0[gn], d1 = P.ddens[gn];
  const double dens_ssm1 = dh + d0 + P.c_ssm1 * (d1 - d0);
  double qn;
  if (P.stage == P.nstage - 1) {
    qn = (P.vartmp[gn] + P.sig_ss * q * dens_ssm1 + P.gam_ss * tend) / (dh + d1);
  } else {
    double v0 = P.var0[gn], vt = P.vartmp[gn];
    if (P.stage == 0) { v0 = q * (dh + d0); vt = 0.0; P.var0[gn] = v0; }
    if (P.upd_vartmp) vt = vt + P.sig_Ns * q * dens_ssm1 + P.gam_Ns * tend;
    if (P.stage == 0 || P.upd_vartmp) P.vartmp[gn] = vt;
    const double dens_ss = dh + d0 + P.c_ss * (d1 - d0);
    qn = ((1.0 - P.sig_ss) * v0 + P.sig_ss * q * dens_ssm1 + P.gam_ss * tend) / dens_ss;
  }
  if (P.do_filter) {   // q <- F3D(rho q) / rho, passes x, y, z
    const double wgt = dh + d1;
    __syncthreads();
    sF[n] = wgt * qn;
    __syncthreads();
    double s = 0.0;
    for (int l = 0; l < np; ++l) s += sFh[i * np + l] * sF[l + j * np + k * np * np];
    sF[Np + n] = s;
    __syncthreads();
    s = 0.0;
    for (int l = 0; l < np; ++l) s += sF[Np + i + l * np + k * np * np] * sFh[j * np + l];
    sF[2 * Np + n] = s;
    __syncthreads();
    s = 0.0;
    for (int l = 0; l < np; ++l) s += sFv[k * np + l] * sF[2 * Np + i + j * np + l * np * np];
    qn = s / wgt;
  }
  if (P.do_tmar) {     // truncation + mass-aware rescaling of the element
    const double w = P.jac[gn] * P.w3[n] * (dh + d1);
    const double Q0 = trc_block_sum(w * qn, sRed);
    const double Q1 = trc_block_sum(w * fmax(0.0, qn), sRed);
    qn = Q0 / (Q1 + 1.0e-32) * fmax(0.0, qn);
  }
  P.qout[gn] = qn;
}

// save_massflux: MFLX = (first ? 0 : MFLX) + w MOM at the interior nodes
__global__ void trc_massflux_accum_kernel(const double* __restrict__ mx, const double* __restrict__ my, const double* __restrict__ mz,
                                          double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ fz, double w_h, double w_v,
                                          int first, size_t n) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (first) { fx[i] = w_h * mx[i]; fy[i] = w_h * my[i]; fz[i] = w_v * mz[i]; }
  else { fx[i] = fx[i] + w_h * mx[i]; fy[i] = fy[i] + w_h * my[i]; fz[i] = fz[i] + w_v * mz[i]; }
}
// cal_alphdens_dyn: the Rusanov coefficient of the dynamics at the stage state (halo filled, boundary condition applied, DPRES of
// the state) times the density of each side, accumulated with the stage weight of the face's direction; flat mesh (Gsqrt = G11 = G22 = 1)
__global__ void trc_alphdens_dyn_kernel(const double* __restrict__ ddens, const double* __restrict__ dens_hyd, const double* __restrict__ mx,
                                        const double* __restrict__ my, const double* __restrict__ mz, const double* __restrict__ pres_hyd,
                                        const double* __restrict__ dpres, const int* __restrict__ vmapP, double* __restrict__ alphM,
                                        double* __restrict__ alphP, double w_h, double w_v, double gamm, int hevi, int first, int Np, int Nfp,
                                        int NfpTot, int np, int Ne) {
  const size_t g = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= size_t(NfpTot) * Ne) return;
  const int ke = int(g / NfpTot), m = int(g - size_t(ke) * NfpTot), f = m / Nfp, fp = m - f * Nfp;
  const size_t iM = size_t(ke) * Np + trc_face_node(f, fp, np), iP = size_t(vmapP[g]);
  double nx, ny, nz;
  trc_normal(f, nx, ny, nz);
  const double anx = fabs(nx), any = fabs(ny), anz = fabs(nz);
  const double Gnn = hevi ? anx + any : anx + any + anz;
  const double densM = ddens[iM] + dens_hyd[iM], densP = ddens[iP] + dens_hyd[iP];
  const double VelM = (mx[iM] * nx + my[iM] * ny + mz[iM] * nz) / densM;
  const double VelP = (mx[iP] * nx + my[iP] * ny + mz[iP] * nz) / densP;
  const double alpha = fmax(sqrt(Gnn * gamm * (pres_hyd[iM] + dpres[iM]) / densM) + fabs(VelM),
                            sqrt(Gnn * gamm * (pres_hyd[iP] + dpres[iP]) / densP) + fabs(VelP));
  const double w = w_h * (anx + any) + w_v * anz;
  const double aM = first ? 0.0 : alphM[g], aP = first ? 0.0 : alphP[g];
  alphM[g] = aM + w * alpha * densM;
  alphP[g] = aP + w * alpha * densP;
}
// QTRC = (DENS_hyd + DDENS_TRC) / (DENS_hyd + DDENS) * QTRC_tmp (driver_trcadv3d.F90:530-537)
__global__ void trc_rescale_kernel(double* __restrict__ q, const double* __restrict__ dens_hyd, const double* __restrict__ dd_trc,
                                   const double* __restrict__ ddens, size_t n) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  q[i] = (dens_hyd[i] + dd_trc[i]) / (dens_hyd[i] + ddens[i]) * q[i];
}

size_t trc_smem_bytes(const TracerParams& P) {
  return (size_t(3) * P.NfpTot + 8 + size_t(4) * P.Np + size_t(3) * P.np * P.np + 2 * P.np + 8) * sizeof(double);
}

}  // namespace

cudaError_t launch_trc_save_massflux(const double* const prog[NVAR], const double* dens_hyd, const double* pres_hyd, const double* dpres,
                                     const int* vmapP, double* const mflx[3], double* alphM, double* alphP, double w_h, double w_v, double gamm,
                                     bool hevi, bool first, int Np, int Nfp, int NfpTot, int np, int Ne, cudaStream_t s) {
  const size_t n = size_t(Np) * Ne, nf = size_t(NfpTot) * Ne;
  trc_massflux_accum_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(prog[V_MOMX], prog[V_MOMY], prog[V_MOMZ], mflx[0], mflx[1], mflx[2], w_h, w_v,
                                                                      first ? 1 : 0, n);
  trc_alphdens_dyn_kernel<<<unsigned((nf + 255) / 256), 256, 0, s>>>(prog[V_DDENS], dens_hyd, prog[V_MOMX], prog[V_MOMY], prog[V_MOMZ], pres_hyd, dpres,
                                                                     vmapP, alphM, alphP, w_h, w_v, gamm, hevi ? 1 : 0, first ? 1 : 0, Np, Nfp, NfpTot, np, Ne);
  return cudaGetLastError();
}
cudaError_t launch_trc_rescale(double* q, const double* dens_hyd, const double* dd_trc, const double* ddens, size_t n, cudaStream_t s) {
  trc_rescale_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(q, dens_hyd, dd_trc, ddens, n);
  return cudaGetLastError();
}
cudaError_t launch_trc_alphdens(const TracerParams& P, cudaStream_t s) {
  const size_t n = size_t(P.NfpTot) * P.Ne;
  trc_alphdens_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(P);
  return cudaGetLastError();
}
cudaError_t launch_trc_fct(const TracerParams& P, cudaStream_t s) {
  const size_t shmem = trc_smem_bytes(P);
  static size_t attr = 0;
  if (shmem > attr) {
    cudaError_t e = cudaFuncSetAttribute(trc_fct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shmem));
    if (e != cudaSuccess) return e;
    attr = shmem;
  }
  trc_fct_kernel<<<P.Ne, P.Np, shmem, s>>>(P);
  return cudaGetLastError();
}
cudaError_t launch_trc_stage(const TracerParams& P, cudaStream_t s) {
  const size_t shmem = trc_smem_bytes(P);
  static size_t attr = 0;
  if (shmem > attr) {
    cudaError_t e = cudaFuncSetAttribute(trc_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shmem));
    if (e != cudaSuccess) return e;
    attr = shmem;
  }
  trc_stage_kernel<<<P.Ne, P.Np, shmem, s>>>(P);
  return cudaGetLastError();
}

}  // namespace fedg
