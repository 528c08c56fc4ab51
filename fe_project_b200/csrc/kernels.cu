// CUDA kernels (sm_100a, FP64) of the DG dynamics hot path.
//
// The fused stage kernel lives in heve_stage.cu; this file holds the small kernels around it: halo fill +
// boundary condition, stand-alone pressure, conservation monitors, element-operator conformance kernels.
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
  H.dp[dst] = H.dp[src];
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
// atm_dyn_dgm_modalfilter_apply stand-alone (scale_atm_dyn_dgm_modalfilter.F90:49-130), five variables in place; used by the
// HEVI path (the HEVE path applies the filter inside its last stage kernel).  One block per element, one thread per node.
__global__ void modal_filter5_kernel(double* q0, double* q1, double* q2, double* q3, double* q4, const double* __restrict__ gsqrt,
                                     int terrain, int np) {
  extern __shared__ double s[];
  const int N2 = np * np, N3 = N2 * np;
  double* w = s + N3;
  const int e = blockIdx.x, n = threadIdx.x;
  const int i = n % np, j = (n / np) % np, k = n / N2;
  double* q[NVAR] = {q0, q1, q2, q3, q4};
  const size_t gi = size_t(e) * N3 + n;
  const double G = terrain ? gsqrt[gi] : 1.0;
  for (int v = 0; v < NVAR; ++v) {
    __syncthreads();
    s[n] = G * q[v][gi];
    __syncthreads();
    double a = cT.Fh[i * np] * s[j * np + k * N2];
    for (int l = 1; l < np; ++l) a += cT.Fh[i * np + l] * s[l + j * np + k * N2];
    w[n] = a;
    __syncthreads();
    double b = w[i + k * N2] * cT.Fh[j * np];
    for (int l = 1; l < np; ++l) b += w[i + l * np + k * N2] * cT.Fh[j * np + l];
    __syncthreads();
    s[n] = b;
    __syncthreads();
    double r = s[i + j * np] * cT.Fv[k * np];
    for (int l = 1; l < np; ++l) r += s[i + j * np + l * N2] * cT.Fv[k * np + l];
    q[v][gi] = r * (1.0 / G);
  }
}
void launch_modal_filter5(double* const q[NVAR], const double* gsqrt, bool terrain, int Ne, int np, cudaStream_t s) {
  const int N3 = np * np * np;
  modal_filter5_kernel<<<Ne, N3, size_t(2) * N3 * sizeof(double), s>>>(q[0], q[1], q[2], q[3], q[4], gsqrt, terrain ? 1 : 0, np);
}


// atm_dyn_dgm_nonhydro3d_common_calc_phyd_hgrad_lc + get_phyd_hgrad_numflux_generalhvc (nonhydro3d_common.F90:624-777): horizontal
// gradient of the hydrostatic pressure (minus a reference profile), the set-up product the driver refreshes after a restart is read
// (driver_nonhydro3d.F90:1060-1095) and every explicit tendency subtracts.  One block per element, one thread per node.
//   DPhydDx = [E11 Dx(Gv P) + E33 Dz(G13 Gv P) + Lift(delx)] / Gv,   Gv = Gsqrt / GsqrtH   (= 1 without topography)
//   delx = (nx + G13_P nz) Fscale/2 Gv_P P_P - (nx + G13_M nz) Fscale/2 Gv_M P_M        (central flux jump)
struct PhydParams {
  const double *pres_hyd, *pres_ref;      // (Np*Ne + halo); pres_ref may be NULL
  const double *gsqrt, *g13, *g23, *gsqrtH;
  const double *escale, *fscale;
  const int *vmapP, *emap2d;
  double *outx, *outy;
  int np, Ne, terrain;
};
__global__ void phyd_hgrad_kernel(const __grid_constant__ PhydParams P) {
  extern __shared__ double sm[];
  const int np = P.np, N2 = np * np, N3 = N2 * np, NFT = 6 * N2;
  double* F = sm;               // Gv * dP
  double* F13 = sm + N3;        // G13 * F   (terrain)
  double* F23 = sm + 2 * N3;    // G23 * F
  double* dlx = sm + 3 * N3;    // [NFT]
  double* dly = dlx + NFT;
  const int ke = blockIdx.x, n = threadIdx.x;
  const int i = n % np, j = (n / np) % np, k = n / N2;
  const size_t g = size_t(ke) * N3 + n;
  const int ke2d = P.emap2d[ke];
  auto dpres = [&](size_t idx) { return P.pres_hyd[idx] - (P.pres_ref ? P.pres_ref[idx] : 0.0); };
  double Gv = 1.0, g13 = 0.0, g23 = 0.0;
  if (P.terrain) { Gv = P.gsqrt[g] / P.gsqrtH[size_t(ke2d) * N2 + (n % N2)]; g13 = P.g13[g]; g23 = P.g23[g]; }
  const double f = Gv * dpres(g);
  F[n] = f; F13[n] = g13 * f; F23[n] = g23 * f;
  for (int fp = n; fp < NFT; fp += N3) {
    const int face = fp / N2, fl = fp % N2;
    int nloc, h2d;
    switch (face) {
      case 0: nloc = (fl % np) + (fl / np) * N2; h2d = fl % np; break;
      case 1: nloc = (np - 1) + (fl % np) * np + (fl / np) * N2; h2d = (np - 1) + (fl % np) * np; break;
      case 2: nloc = (fl % np) + (np - 1) * np + (fl / np) * N2; h2d = (fl % np) + (np - 1) * np; break;
      case 3: nloc = (fl % np) * np + (fl / np) * N2; h2d = (fl % np) * np; break;
      case 4: nloc = fl; h2d = fl; break;
      default: nloc = fl + (np - 1) * N2; h2d = fl; break;
    }
    const size_t iM = size_t(ke) * N3 + nloc, iP = size_t(P.vmapP[size_t(ke) * NFT + fp]);
    const double nx = (face == 1) ? 1.0 : (face == 3) ? -1.0 : 0.0;
    const double ny = (face == 2) ? 1.0 : (face == 0) ? -1.0 : 0.0;
    const double nz = (face == 5) ? 1.0 : (face == 4) ? -1.0 : 0.0;
    double GvM = 1.0, GvP = 1.0, g13M = 0.0, g13P = 0.0, g23M = 0.0, g23P = 0.0;
    if (P.terrain) {
      const double gh = P.gsqrtH[size_t(ke2d) * N2 + h2d];
      GvM = P.gsqrt[iM] / gh; GvP = P.gsqrt[iP] / gh;
      g13M = P.g13[iM]; g13P = P.g13[iP]; g23M = P.g23[iM]; g23P = P.g23[iP];
    }
    const double fs = P.fscale[size_t(face) * P.Ne + ke];
    const double t1 = fs * 0.5 * GvP * dpres(iP), t2 = fs * 0.5 * GvM * dpres(iM);
    dlx[fp] = (nx + g13P * nz) * t1 - (nx + g13M * nz) * t2;
    dly[fp] = (ny + g23P * nz) * t1 - (ny + g23M * nz) * t2;
  }
  __syncthreads();
  double dx = 0.0, dy = 0.0, dz1 = 0.0, dz2 = 0.0;
  for (int l = 0; l < np; ++l) {
    dx += cT.D[i * np + l] * F[l + j * np + k * N2];
    dy += cT.D[j * np + l] * F[i + l * np + k * N2];
    if (P.terrain) { dz1 += cT.D[k * np + l] * F13[i + j * np + l * N2]; dz2 += cT.D[k * np + l] * F23[i + j * np + l * N2]; }
  }
  // Lift: the same tensor-product weights as elem_op_kernel (op 3)
  const double lx = cT.Lw[j * 2] * dlx[i + k * np] + cT.Lw[i * 2 + 1] * dlx[N2 + j + k * np] + cT.Lw[j * 2 + 1] * dlx[2 * N2 + i + k * np] +
                    cT.Lw[i * 2] * dlx[3 * N2 + j + k * np] + cT.Lw[k * 2] * dlx[4 * N2 + i + j * np] + cT.Lw[k * 2 + 1] * dlx[5 * N2 + i + j * np];
  const double ly = cT.Lw[j * 2] * dly[i + k * np] + cT.Lw[i * 2 + 1] * dly[N2 + j + k * np] + cT.Lw[j * 2 + 1] * dly[2 * N2 + i + k * np] +
                    cT.Lw[i * 2] * dly[3 * N2 + j + k * np] + cT.Lw[k * 2] * dly[4 * N2 + i + j * np] + cT.Lw[k * 2 + 1] * dly[5 * N2 + i + j * np];
  const double E11 = P.escale[ke], E22 = P.escale[size_t(P.Ne) + ke], E33 = P.escale[2 * size_t(P.Ne) + ke];
  P.outx[g] = (E11 * dx + E33 * dz1 + lx) / Gv;
  P.outy[g] = (E22 * dy + E33 * dz2 + ly) / Gv;
}
void launch_phyd_hgrad(const double* pres_hyd, const double* pres_ref, const double* gsqrt, const double* g13, const double* g23,
                       const double* gsqrtH, const double* escale, const double* fscale, const int* vmapP, const int* emap2d,
                       double* outx, double* outy, int np, int Ne, bool terrain, cudaStream_t s) {
  PhydParams P{pres_hyd, pres_ref, gsqrt, g13, g23, gsqrtH, escale, fscale, vmapP, emap2d, outx, outy, np, Ne, terrain ? 1 : 0};
  const int N3 = np * np * np;
  phyd_hgrad_kernel<<<Ne, N3, (size_t(3) * N3 + 12 * np * np) * sizeof(double), s>>>(P);
}

void launch_elem_op(int op, const double* in, double* out, int nelem, int np, cudaStream_t s) {
  int N3 = np * np * np;
  size_t sh = size_t(2) * (6 * np * np > N3 ? 6 * np * np : N3) * sizeof(double);
  elem_op_kernel<<<nelem, N3, sh, s>>>(op, in, out, np);
}

}  // namespace fedg
