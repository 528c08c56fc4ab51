// sample/advect3d on the GPU (BASELINE config 1; SURVEY.md rows a1 + a18).
//
//   advect_stage_kernel   advect3d_kernel_cal_tend (sample/advect3d/mod_advect3d_kernel.f90:34-166) fused with the
//                         tint%Advance that follows it in the stage loop (sample/advect3d/test_advect3d.f90:108-121):
//                         upwind element-boundary flux -> Dx/Dy/Dz/Lift products -> tendency -> RK update.
//   ell_spmv_kernel       sparsemat_matmul for the ELL storage (FElib/src/common/scale_sparsemat.F90:554-634),
//                         batched over right-hand sides; conformance entry for row a1.
//
// The sample hands Dx, Dy, Dz, Lift over as `sparsemat` objects, so the kernel consumes exactly those: the ELL
// arrays val(M*col_size), colIdx(M*col_size) with slot-major storage l = i + k*M (scale_sparsemat.F90:172), staged
// in shared memory once per block.  One thread per node; a block holds EPB elements (4 at p = 3).  The products
// run slot by slot in ascending k like the reference's loop, so the summation order is the reference's.
#include "fedg_internal.h"

namespace fedg {

namespace {

__device__ __forceinline__ int face_node(int f, int fp, int np) {
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

// dynamic shared memory:  [ELL values 3*Np*csD + Np*csL] [sF 3*EPB*Np] [sFl EPB*NfpTot] | [ELL columns (int)]
__global__ void advect_stage_kernel(const __grid_constant__ AdvectParams P) {
  extern __shared__ __align__(16) double sm[];
  const int Np = P.Np, NfpTot = P.NfpTot, Nfp = P.Nfp, np = P.np, EPB = P.epb;
  const int nD = Np * P.colsz[0], nL = Np * P.colsz[3];
  double* sVal = sm;                                  // Dx | Dy | Dz | Lift
  double* sF = sVal + 3 * nD + nL;                    // [3][EPB*Np]  q*u, q*v, q*w
  double* sFl = sF + 3 * EPB * Np;                    // [EPB*NfpTot] Fscale * boundary flux
  int* sCol = reinterpret_cast<int*>(sFl + EPB * NfpTot);
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int m = tid; m < 3 * nD + nL; m += nthr) {
    const int which = m < 3 * nD ? m / nD : 3, off = m < 3 * nD ? m - which * nD : m - 3 * nD;
    sVal[m] = P.ellval[which][off];
    sCol[m] = P.ellcol[which][off];
  }
  const int el = tid / Np, n = tid - el * Np;
  const int ke = blockIdx.x * EPB + el;
  const bool live = ke < P.Ne;
  const size_t gn = size_t(live ? ke : 0) * Np + n;
  double q = 0.0;
  if (live) {
    q = P.q[gn];
    sF[0 * EPB * Np + tid] = q * P.u[gn];
    sF[1 * EPB * Np + tid] = q * P.v[gn];
    sF[2 * EPB * Np + tid] = q * P.w[gn];
    // cal_elembnd_flux (mod_advect3d_kernel.f90:131-166): alpha = 0.5 |VelP + VelM|
    for (int m = n; m < NfpTot; m += Np) {
      const int f = m / Nfp, fp = m - f * Nfp;
      const size_t iM = size_t(ke) * Np + face_node(f, fp, np);
      const size_t iP = size_t(P.vmapP[size_t(ke) * NfpTot + m]);
      const double nx = (f == 1) ? 1.0 : (f == 3) ? -1.0 : 0.0;
      const double ny = (f == 2) ? 1.0 : (f == 0) ? -1.0 : 0.0;
      const double nz = (f == 5) ? 1.0 : (f == 4) ? -1.0 : 0.0;
      const double VelM = P.u[iM] * nx + P.v[iM] * ny + P.w[iM] * nz;
      const double VelP = P.u[iP] * nx + P.v[iP] * ny + P.w[iP] * nz;
      const double qM = P.q[iM], qP = P.q[iP];
      const double alpha = 0.5 * fabs(VelP + VelM);
      const double fl = 0.5 * ((qP * VelP - qM * VelM) - alpha * (qP - qM));
      sFl[el * NfpTot + m] = P.fscale[size_t(f) * P.Ne + ke] * fl;
    }
  }
  __syncthreads();
  if (!live) return;
  // cal_dqdt (mod_advect3d_kernel.f90:86-127): four sparsemat products, slots ascending
  double Fx = 0.0, Fy = 0.0, Fz = 0.0, L = 0.0;
  const double* bx = sF + 0 * EPB * Np + el * Np;
  const double* by = sF + 1 * EPB * Np + el * Np;
  const double* bz = sF + 2 * EPB * Np + el * Np;
  for (int k = 0; k < P.colsz[0]; ++k) {
    const int l = n + k * Np;
    Fx += sVal[l] * bx[sCol[l]];
    Fy += sVal[nD + l] * by[sCol[nD + l]];
    Fz += sVal[2 * nD + l] * bz[sCol[2 * nD + l]];
  }
  for (int k = 0; k < P.colsz[3]; ++k) {
    const int l = 3 * nD + n + k * Np;
    L += sVal[l] * sFl[el * NfpTot + sCol[l]];
  }
  const double E11 = P.escale[ke], E22 = P.escale[P.Ne + ke], E33 = P.escale[2 * size_t(P.Ne) + ke];
  const double tend = -(E11 * Fx + E22 * Fy + E33 * Fz + L);
  if (P.tend_out) { P.tend_out[gn] = tend; return; }
  // tint%Advance (scale_timeint_rk.F90:1182-1266 low storage, :2201-2355 general with one tendency buffer)
  double base = 0.0;
  if (P.rk.use_q0) base = P.rk.c_q0 * P.q0[gn];
  if (P.rk.add_vt) base = P.vt[gn];
  if (P.rk.vt_update) {
    const double vb = P.rk.vt_init ? P.rk.vt_init_q * q : P.vt[gn];
    P.vt[gn] = vb + P.rk.vt_q * q + P.rk.vt_k * tend;
  }
  P.qout[gn] = base + P.rk.c_q * q + P.rk.c_k * tend;
}

// halo of q, u, v, w for faces whose neighbour is this tile (periodic wrap / self map): fields_comm Put/Exchange/Get
// of test_advect3d.f90:97-101 collapsed into one gather
__global__ void advect_halo_kernel(double* q, double* u, double* v, double* w, const int* __restrict__ src, size_t nint, int nhalo,
                                   int with_vel) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= nhalo) return;
  const int s = src[h];
  if (s < 0) return;
  q[nint + h] = q[s];
  if (with_vel) { u[nint + h] = u[s]; v[nint + h] = v[s]; w[nint + h] = w[s]; }
}

// c(:,j) = A b(:,j), ELL storage, one block per right-hand side (b staged in shared memory)
__global__ void ell_spmv_kernel(int M, int N, int col_size, const double* __restrict__ val, const int* __restrict__ col,
                                const double* __restrict__ b, double* __restrict__ c) {
  extern __shared__ double sb[];
  const double* bj = b + size_t(blockIdx.x) * N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) sb[i] = bj[i];
  __syncthreads();
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < col_size; ++k) { const size_t l = size_t(i) + size_t(k) * M; s += val[l] * sb[col[l]]; }
    c[size_t(blockIdx.x) * M + i] = s;
  }
}

// General sparsemat product for the conformance entries (row a1): CSR or ELL storage, one or two operand vectors, vectors stored
// one after the other (matmul1, batched) or interleaved (matmul2: b(NQ,N), c(NQ,M)).  One thread per (row, vector); the sums run in
// the reference's order (CSR: j ascending, scale_sparsemat.F90:439-474; ELL: slots ascending, :554-634).
//   b[col * sb_col + q * sb_q],  c[row * sc_row + q * sc_q]
__global__ void sparsemat_general_kernel(int M, int nq, int col_size, const double* __restrict__ val, const int* __restrict__ col,
                                         const int* __restrict__ rowptr, const double* __restrict__ b1, const double* __restrict__ b2,
                                         double* __restrict__ c, size_t sb_col, size_t sb_q, size_t sc_row, size_t sc_q) {
  const size_t x = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (x >= size_t(M) * nq) return;
  const int i = int(x % M), q = int(x / M);
  const double* bq1 = b1 + q * sb_q;
  const double* bq2 = b2 ? b2 + q * sb_q : nullptr;
  double s = 0.0;
  if (rowptr) {
    for (int j = rowptr[i]; j < rowptr[i + 1]; ++j) {
      const size_t cj = size_t(col[j]) * sb_col;
      s = bq2 ? s + val[j] * bq1[cj] * bq2[cj] : s + val[j] * bq1[cj];
    }
  } else {
    for (int k = 0; k < col_size; ++k) {
      const size_t l = size_t(i) + size_t(k) * M, cj = size_t(col[l]) * sb_col;
      s = bq2 ? s + val[l] * bq1[cj] * bq2[cj] : s + val[l] * bq1[cj];
    }
  }
  c[i * sc_row + q * sc_q] = s;
}

}  // namespace

cudaError_t launch_sparsemat_general(int M, int nq, int col_size, const double* val, const int* col, const int* rowptr, const double* b1,
                                     const double* b2, double* c, size_t sb_col, size_t sb_q, size_t sc_row, size_t sc_q, cudaStream_t s) {
  const size_t n = size_t(M) * nq;
  sparsemat_general_kernel<<<unsigned((n + 127) / 128), 128, 0, s>>>(M, nq, col_size, val, col, rowptr, b1, b2, c, sb_col, sb_q, sc_row, sc_q);
  return cudaGetLastError();
}

size_t advect_smem_bytes(const AdvectParams& P) {
  const size_t nell = size_t(3) * P.Np * P.colsz[0] + size_t(P.Np) * P.colsz[3];
  return (nell + size_t(3) * P.epb * P.Np + size_t(P.epb) * P.NfpTot) * sizeof(double) + nell * sizeof(int);
}

cudaError_t launch_advect_stage(const AdvectParams& P, cudaStream_t s) {
  const size_t shmem = advect_smem_bytes(P);
  static size_t attr = 0;
  if (shmem > attr) {
    cudaError_t e = cudaFuncSetAttribute(advect_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shmem));
    if (e != cudaSuccess) return e;
    attr = shmem;
  }
  const int grid = (P.Ne + P.epb - 1) / P.epb;
  advect_stage_kernel<<<grid, P.epb * P.Np, shmem, s>>>(P);
  return cudaGetLastError();
}

void launch_advect_halo(double* q, double* u, double* v, double* w, const int* src, size_t nint, int nhalo, bool with_vel,
                        cudaStream_t s) {
  if (nhalo <= 0) return;
  advect_halo_kernel<<<(nhalo + 255) / 256, 256, 0, s>>>(q, u, v, w, src, nint, nhalo, with_vel ? 1 : 0);
}

cudaError_t launch_ell_spmv(int M, int N, int col_size, const double* val, const int* col, const double* b, double* c, int nvec,
                            cudaStream_t s) {
  const size_t shmem = size_t(N) * sizeof(double);
  if (shmem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ell_spmv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shmem));
    if (e != cudaSuccess) return e;
  }
  ell_spmv_kernel<<<nvec, 256, shmem, s>>>(M, N, col_size, val, col, b, c);
  return cudaGetLastError();
}

}  // namespace fedg
