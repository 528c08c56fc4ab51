// Numerical diffusion of the prognostic variables, sm_100a FP64.
//
//   AtmDyn_Nonhydro3D_Numdiff%Apply      fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_numdiff.F90:214-376
//   numdiff_cal_flx / cal_del_gradDiffVar                                                 :600-720   -> MODE_FLX
//   numdiff_cal_laplacian / cal_del_flux_lap                                              :505-597   -> MODE_LAP
//   numdiff_tend / cal_del_flux_lap_with_coef + the update var += dt * tend               :379-502   -> MODE_TEND
//   ApplyBC_numdiff_even_lc / _odd_lc     fluid_dyn_solver/scale_atm_dyn_dgm_bnd.F90:370-508          -> evaluated at the face node
//
// The reference runs, per variable, a boundary-condition pass over the halo, a face-flux pass, an element pass and an update
// pass for each half-step of the local-DG Laplacian.  Here one launch per half-step does all of it: one block per element,
// one thread per node; the exterior value of a face node on a physical boundary is formed from the interior value by the
// boundary rule instead of being written to the halo first; elsewhere it is read from the halo slots (filled by the
// exchange) or from the neighbour element.
#include <cstdint>
#include <cstdlib>

#include "fedg_internal.h"

namespace fedg {

namespace {
enum { MODE_FLX = 0, MODE_LAP = 1, MODE_TEND = 2 };

__device__ __forceinline__ int nd_face_node(int f, int fp, int np) {
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

template <int MODE>
__global__ void __launch_bounds__(512, 3) numdiff_kernel(const __grid_constant__ NumdiffParams P) {
  extern __shared__ double sm[];
  const int np = P.np, N2 = np * np, Np = P.Np, Nfp = P.Nfp, NfpTot = P.NfpTot;
  double* sD = sm;                  // D1D[i][l], row-major: a node reads its rows as 128-bit loads (the transposed table, conflict-free
                                    // per load but eight 64-bit loads per row, measured slower: 3.81 vs 3.25 ms per Apply, tools/numdiff_time.py)
  double* sLw = sD + N2;            // lift1d[m][side]
  double* sV = sLw + 2 * np;        // [3][Np] volume operands
  double* sJ = sV + 3 * Np;         // [3][NfpTot] Fscale * face jumps
  const int n = threadIdx.x, ke = blockIdx.x;
  const int i = n % np, j = (n / np) % np, k = n / N2;
  if (n < N2) sD[n] = P.tab->D[n];
  if (n < 2 * np) sLw[n] = P.tab->Lw[n];
  const size_t gn = size_t(ke) * Np + n;
  const bool dens = P.dens_flag != 0;
  double rho = 1.0;
  if (dens) rho = P.ddens[gn] + P.dens_hyd[gn];
  if (MODE == MODE_FLX) {
    const double w = dens ? 1.0 / rho : 1.0;
    const double vh = P.in0[gn] * w, vv = P.in1[gn] * w;
    sV[n] = vh; sV[Np + n] = vv;
  } else if (MODE == MODE_LAP) {
    sV[n] = P.in0[gn]; sV[Np + n] = P.in1[gn]; sV[2 * Np + n] = P.in2[gn];
  } else {
    const double ch = dens ? P.coef_h * rho : P.coef_h, cv = dens ? P.coef_v * rho : P.coef_v;
    sV[n] = ch * P.in0[gn]; sV[Np + n] = ch * P.in1[gn]; sV[2 * Np + n] = cv * P.in2[gn];
  }
  // ---- face jumps
  for (int m = n; m < NfpTot; m += Np) {
    const int f = m / Nfp, fp = m - f * Nfp;
    const size_t iM = size_t(ke) * Np + nd_face_node(f, fp, np);
    const size_t iP = size_t(P.vmapP[size_t(ke) * NfpTot + m]);
    const double nx = (f == 1) ? 1.0 : (f == 3) ? -1.0 : 0.0;
    const double ny = (f == 2) ? 1.0 : (f == 0) ? -1.0 : 0.0;
    const double nz = (f == 5) ? 1.0 : (f == 4) ? -1.0 : 0.0;
    int vel = 0, therm = 0;
    if (iP >= P.nint) {            // halo slot: which tile face, which boundary condition
      const int h = int(iP - P.nint);
      int tf = 0;
      while (h >= P.face_off[tf + 1]) ++tf;
      vel = P.vel_bc[tf]; therm = P.therm_bc[tf];
    }
    const double hf = P.fscale[size_t(f) * P.Ne + ke];
    if (MODE == MODE_FLX) {
      // ApplyBC_numdiff_even_lc (on Varh; on Varv too in the first half-step where both are the variable itself)
      const bool is_bound = (vel == FEDG_BND_SLIP || vel == FEDG_BND_NOSLIP);
      double hM = P.in0[iM], vM = P.in1[iM], hP = P.in0[iP], vP = P.in1[iP];
      if (is_bound) {
        const bool mom = (P.varid == V_MOMX || P.varid == V_MOMY || P.varid == V_MOMZ);
        const double nn = (P.varid == V_MOMX) ? nx : (P.varid == V_MOMY) ? ny : nz;
        double eh = hP, ev = vP;
        if (vel == FEDG_BND_SLIP && mom) { eh = hM - 2.0 * (hM * nn) * nn; ev = vM - 2.0 * (vM * nn) * nn; }
        else if (vel == FEDG_BND_NOSLIP && mom) { eh = -hM; ev = -vM; }
        hP = eh;
        if (P.bc_on_v) vP = ev;
      }
      double wP = 1.0, wM = 1.0;
      if (dens) { wP = 1.0 / (P.ddens[iP] + P.dens_hyd[iP]); wM = 1.0 / (P.ddens[iM] + P.dens_hyd[iM]); }
      const double dh = 0.5 * (hP * wP - hM * wM), dv = 0.5 * (vP * wP - vM * wM);
      // not on a boundary: the jump enters only through the faces with a negative normal (alternating flux)
      const double sx = is_bound ? 1.0 : (1.0 - (nx >= 0.0 ? 1.0 : -1.0)), sy = is_bound ? 1.0 : (1.0 - (ny >= 0.0 ? 1.0 : -1.0)),
                   sz = is_bound ? 1.0 : (1.0 - (nz >= 0.0 ? 1.0 : -1.0));
      sJ[m] = hf * (sx * dh * nx); sJ[NfpTot + m] = hf * (sy * dh * ny); sJ[2 * NfpTot + m] = hf * (sz * dv * nz);
    } else {
      // ApplyBC_numdiff_odd_lc
      const bool is_bound = (vel == FEDG_BND_SLIP) || (therm == 1);
      const double xM = P.in0[iM], yM = P.in1[iM], zM = P.in2[iM];
      double xP = P.in0[iP], yP = P.in1[iP], zP = P.in2[iP];
      if (is_bound) {
        const double gnrm = xM * nx + yM * ny + zM * nz;
        if (vel == FEDG_BND_SLIP) {
          if (P.varid == V_MOMX) { yP = yM - 2.0 * gnrm * ny; zP = zM - 2.0 * gnrm * nz; }
          else if (P.varid == V_MOMY) { xP = xM - 2.0 * gnrm * nx; zP = zM - 2.0 * gnrm * nz; }
          else if (P.varid == V_MOMZ) { xP = xM - 2.0 * gnrm * nx; yP = yM - 2.0 * gnrm * ny; }
        }
        if (therm == 1 && (P.varid == V_DDENS || P.varid == V_DRHOT)) {
          xP = xM - 2.0 * gnrm * nx; yP = yM - 2.0 * gnrm * ny; zP = zM - 2.0 * gnrm * nz;
        }
      }
      const double sx = is_bound ? 1.0 : (1.0 + (nx >= 0.0 ? 1.0 : -1.0)), sy = is_bound ? 1.0 : (1.0 + (ny >= 0.0 ? 1.0 : -1.0)),
                   sz = is_bound ? 1.0 : (1.0 + (nz >= 0.0 ? 1.0 : -1.0));
      if (MODE == MODE_LAP) {
        sJ[m] = hf * (0.5 * (sx * (xP - xM) * nx + sy * (yP - yM) * ny));
        sJ[NfpTot + m] = hf * (0.5 * sz * (zP - zM) * nz);
      } else {
        double wM = 0.5, wP = 0.5;
        if (dens) { wM = 0.5 * (P.dens_hyd[iM] + P.ddens[iM]); wP = 0.5 * (P.dens_hyd[iP] + P.ddens[iP]); }
        sJ[m] = hf * (sx * P.coef_h * (wP * xP - wM * xM) * nx + sy * P.coef_h * (wP * yP - wM * yM) * ny +
                      sz * P.coef_v * (wP * zP - wM * zM) * nz);
      }
    }
  }
  __syncthreads();
  // ---- element operators: tensor-product derivatives (rows from shared memory) + lift
  const double* Di = sD + i * np; const double* Dj = sD + j * np; const double* Dk = sD + k * np;
  auto lift = [&](const double* d) {
    return sLw[j * 2] * d[i + k * np] + sLw[i * 2 + 1] * d[Nfp + j + k * np] + sLw[j * 2 + 1] * d[2 * Nfp + i + k * np] +
           sLw[i * 2] * d[3 * Nfp + j + k * np] + sLw[k * 2] * d[4 * Nfp + i + j * np] + sLw[k * 2 + 1] * d[5 * Nfp + i + j * np];
  };
  const double E11 = P.escale[ke], E22 = P.escale[P.Ne + ke], E33 = P.escale[2 * size_t(P.Ne) + ke];
  if (MODE == MODE_FLX) {
    double dx = 0.0, dy = 0.0, dz = 0.0;
    for (int l = 0; l < np; ++l) {
      dx += Di[l] * sV[l + j * np + k * N2];
      dy += Dj[l] * sV[i + l * np + k * N2];
      dz += Dk[l] * sV[Np + i + j * np + l * N2];
    }
    P.out0[gn] = E11 * dx + lift(sJ);
    P.out1[gn] = E22 * dy + lift(sJ + NfpTot);
    P.out2[gn] = E33 * dz + lift(sJ + 2 * NfpTot);
  } else {
    double dx = 0.0, dy = 0.0, dz = 0.0;
    for (int l = 0; l < np; ++l) {
      dx += Di[l] * sV[l + j * np + k * N2];
      dy += Dj[l] * sV[Np + i + l * np + k * N2];
      dz += Dk[l] * sV[2 * Np + i + j * np + l * N2];
    }
    if (MODE == MODE_LAP) {
      P.out0[gn] = (E11 * dx + E22 * dy + lift(sJ));
      P.out1[gn] = (E33 * dz + lift(sJ + NfpTot));
    } else {
      const double tend = (E11 * dx + E22 * dy + E33 * dz + lift(sJ));
      P.var[gn] = P.var[gn] + P.dt * tend;
    }
  }
}

// ---- p = 7: the same half-step on the FP64 tensor cores, in the mapping of stage_p7.cu (block = 256 threads = 8 warps = one element,
// warp w owns the plane k = w, lane (g, t) the node pair (i = 2t, 2t+1; j = g); x / y contractions on the own plane, the z contraction
// on the tile of fixed j = w, lift as one more k = 4 step).  The node-per-thread kernel above reads 60 shared-memory words per node and
// is bound by that pipe (0.3 ms per launch at 32x32x16, 0.18-0.24 of the HBM roofline; A/B in tools/numdiff_time.py): here every operand
// crosses shared memory once.  Face jumps: the code of the kernel above, interior side from the staged element, exterior side gathered.
namespace p7nd {
constexpr int NP = 8, N2 = 64, N3 = 512, NFT = 384, KS_FZ = 70, KS_Z = 72, PLS = 10;
constexpr int SM_DOUBLES = 80 + 4 * N3 + NFT + NP * KS_FZ + NP * KS_Z + 8 * NP * PLS;
}
__device__ __forceinline__ void nd_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 nd_ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
// (Gathering the exterior side of the face nodes with 8-byte cp.async copies issued at block start, to overlap them with the element's own
// loads, was measured slower: 1.91 vs 1.43 ms per Apply -- 1920 eight-byte asynchronous copies per element cost more than the latency.)

// one variable of one element; the caller has loaded the tables, the face-node indices and the density of the own node pair
template <int MODE>
__device__ __forceinline__ void nd_p7_body(const NumdiffParams& P, const NdVar& V, double* sm, size_t iPa, size_t iPb, double2 rho_own) {
  using namespace p7nd;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int ke = blockIdx.x;
  const size_t eb = size_t(ke) * N3;
  const int n0 = 2 * t + 8 * g + 64 * w;
  const size_t gn = eb + n0;
  double* sD = sm; double* sLw = sm + 64; double* sRaw = sm + 80; double* sDel = sRaw + 4 * N3; double* sFz = sDel + NFT;
  double* sZ = sFz + NP * KS_FZ; double* sPl = sZ + NP * KS_Z + w * NP * PLS;
  const bool dens = V.dens_flag != 0;
  const double2 a0 = nd_ld2(V.in0 + gn), a1 = nd_ld2(V.in1 + gn);
  double2 a2 = make_double2(0.0, 0.0), rho = make_double2(1.0, 1.0);
  if (MODE != MODE_FLX) a2 = nd_ld2(V.in2 + gn);
  if (dens) rho = rho_own;
  *reinterpret_cast<double2*>(sRaw + n0) = a0;
  *reinterpret_cast<double2*>(sRaw + N3 + n0) = a1;
  *reinterpret_cast<double2*>(sRaw + 2 * N3 + n0) = (MODE == MODE_FLX) ? rho : a2;
  if (MODE == MODE_TEND) *reinterpret_cast<double2*>(sRaw + 3 * N3 + n0) = rho;
  // volume operands of the own node pair
  double2 Fx, Fy, Fz;
  if (MODE == MODE_FLX) {
    const double wx = dens ? 1.0 / rho.x : 1.0, wy = dens ? 1.0 / rho.y : 1.0;
    Fx = make_double2(a0.x * wx, a0.y * wy); Fy = Fx; Fz = make_double2(a1.x * wx, a1.y * wy);
  } else if (MODE == MODE_LAP) {
    Fx = a0; Fy = a1; Fz = a2;
  } else {
    const double chx = dens ? P.coef_h * rho.x : P.coef_h, chy = dens ? P.coef_h * rho.y : P.coef_h;
    const double cvx = dens ? P.coef_v * rho.x : P.coef_v, cvy = dens ? P.coef_v * rho.y : P.coef_v;
    Fx = make_double2(chx * a0.x, chy * a0.y); Fy = make_double2(chx * a1.x, chy * a1.y); Fz = make_double2(cvx * a2.x, cvy * a2.y);
  }
  *reinterpret_cast<double2*>(sPl + PLS * g + 2 * t) = Fy;
  *reinterpret_cast<double2*>(sFz + 2 * t + 8 * g + KS_FZ * w) = Fz;
  __syncthreads();

  // ---- face jumps (384 face nodes over 256 threads)
#pragma unroll 1
  for (int m = tid; m < NFT; m += 256) {
    const int f = m >> 6, fp = m & 63;
    const int nloc = nd_face_node(f, fp, NP);
    const size_t iP = (m < 256) ? iPa : iPb;
    const double nx = (f == 1) ? 1.0 : (f == 3) ? -1.0 : 0.0;
    const double ny = (f == 2) ? 1.0 : (f == 0) ? -1.0 : 0.0;
    const double nz = (f == 5) ? 1.0 : (f == 4) ? -1.0 : 0.0;
    int vel = 0, therm = 0;
    if (iP >= P.nint) {
      const int h = int(iP - P.nint);
      int tf = 0;
      while (h >= P.face_off[tf + 1]) ++tf;
      vel = P.vel_bc[tf]; therm = P.therm_bc[tf];
    }
    const double hf = P.fscale[size_t(f) * P.Ne + ke];
    double jump;
    if (MODE == MODE_FLX) {
      const bool is_bound = (vel == FEDG_BND_SLIP || vel == FEDG_BND_NOSLIP);
      const double hM = sRaw[nloc], vM = sRaw[N3 + nloc];
      double hP = V.in0[iP], vP = V.in1[iP];
      if (is_bound) {
        const bool mom = (V.varid == V_MOMX || V.varid == V_MOMY || V.varid == V_MOMZ);
        const double nn = (V.varid == V_MOMX) ? nx : (V.varid == V_MOMY) ? ny : nz;
        double eh = hP, ev = vP;
        if (vel == FEDG_BND_SLIP && mom) { eh = hM - 2.0 * (hM * nn) * nn; ev = vM - 2.0 * (vM * nn) * nn; }
        else if (vel == FEDG_BND_NOSLIP && mom) { eh = -hM; ev = -vM; }
        hP = eh;
        if (P.bc_on_v) vP = ev;
      }
      double wP = 1.0, wM = 1.0;
      if (dens) { wP = 1.0 / (P.ddens[iP] + P.dens_hyd[iP]); wM = 1.0 / sRaw[2 * N3 + nloc]; }
      const double dh = 0.5 * (hP * wP - hM * wM), dv = 0.5 * (vP * wP - vM * wM);
      const double sx = is_bound ? 1.0 : (1.0 - (nx >= 0.0 ? 1.0 : -1.0)), sy = is_bound ? 1.0 : (1.0 - (ny >= 0.0 ? 1.0 : -1.0)),
                   sz = is_bound ? 1.0 : (1.0 - (nz >= 0.0 ? 1.0 : -1.0));
      jump = (f == 1 || f == 3) ? hf * (sx * dh * nx) : (f == 0 || f == 2) ? hf * (sy * dh * ny) : hf * (sz * dv * nz);
    } else {
      const bool is_bound = (vel == FEDG_BND_SLIP) || (therm == 1);
      const double xM = sRaw[nloc], yM = sRaw[N3 + nloc], zM = sRaw[2 * N3 + nloc];
      double xP = V.in0[iP], yP = V.in1[iP], zP = V.in2[iP];
      if (is_bound) {
        const double gnrm = xM * nx + yM * ny + zM * nz;
        if (vel == FEDG_BND_SLIP) {
          if (V.varid == V_MOMX) { yP = yM - 2.0 * gnrm * ny; zP = zM - 2.0 * gnrm * nz; }
          else if (V.varid == V_MOMY) { xP = xM - 2.0 * gnrm * nx; zP = zM - 2.0 * gnrm * nz; }
          else if (V.varid == V_MOMZ) { xP = xM - 2.0 * gnrm * nx; yP = yM - 2.0 * gnrm * ny; }
        }
        if (therm == 1 && (V.varid == V_DDENS || V.varid == V_DRHOT)) {
          xP = xM - 2.0 * gnrm * nx; yP = yM - 2.0 * gnrm * ny; zP = zM - 2.0 * gnrm * nz;
        }
      }
      const double sx = is_bound ? 1.0 : (1.0 + (nx >= 0.0 ? 1.0 : -1.0)), sy = is_bound ? 1.0 : (1.0 + (ny >= 0.0 ? 1.0 : -1.0)),
                   sz = is_bound ? 1.0 : (1.0 + (nz >= 0.0 ? 1.0 : -1.0));
      if (MODE == MODE_LAP) {
        jump = (f < 4) ? hf * (0.5 * (sx * (xP - xM) * nx + sy * (yP - yM) * ny)) : hf * (0.5 * sz * (zP - zM) * nz);
      } else {
        double wM = 0.5, wP = 0.5;
        if (dens) { wM = 0.5 * sRaw[3 * N3 + nloc]; wP = 0.5 * (P.dens_hyd[iP] + P.ddens[iP]); }
        jump = hf * (sx * P.coef_h * (wP * xP - wM * xM) * nx + sy * P.coef_h * (wP * yP - wM * yM) * ny +
                     sz * P.coef_v * (wP * zP - wM * zM) * nz);
      }
    }
    sDel[m] = jump;
  }
  const double E11 = P.escale[ke], E22 = P.escale[P.Ne + ke], E33 = P.escale[2 * size_t(P.Ne) + ke];
  const double2 Dg = nd_ld2(sD + g * NP + 2 * t);          // D[g][2t], D[g][2t+1]: the contraction index is split l = 2t | 2t + 1
  const double lwA = (t < 2) ? sLw[g * 2 + t] : 0.0;
  __syncthreads();

  // ---- z contraction + z-face lift on the tile of fixed j = w:  out_j[k][i] = sum_l (E33 D)[k][l] Fz_j[l][i] + Lw[k][s] jump_s[i]
  double z0 = 0.0, z1 = 0.0;
  {
    const double* src = sFz + g + 8 * w;
    const double bl = (t < 2) ? sDel[(4 + t) * N2 + g + 8 * w] : 0.0;
    nd_dmma(z0, z1, E33 * Dg.x, src[KS_FZ * 2 * t]);
    nd_dmma(z0, z1, E33 * Dg.y, src[KS_FZ * (2 * t + 1)]);
    nd_dmma(z0, z1, lwA, bl);
  }
  const size_t gz = eb + 2 * t + 8 * w + 64 * g;            // the nodes of the z tile's C fragment: (i = 2t, 2t+1; j = w; k = g)
  if (MODE == MODE_FLX) *reinterpret_cast<double2*>(V.out2 + gz) = make_double2(z0, z1);
  else if (MODE == MODE_LAP) *reinterpret_cast<double2*>(V.out1 + gz) = make_double2(z0, z1);
  else {
    *reinterpret_cast<double2*>(sZ + 2 * t + 8 * w + KS_Z * g) = make_double2(z0, z1);
    __syncthreads();
  }
  // ---- x / y contractions + lateral lift on the own plane
  const double bx0 = E11 * Dg.x, bx1 = E11 * Dg.y, ay0 = E22 * Dg.x, ay1 = E22 * Dg.y;
  const double by0 = sPl[PLS * 2 * t + g], by1 = sPl[PLS * (2 * t + 1) + g];
  const double axl = (t < 2) ? sDel[(t == 0 ? 3 : 1) * N2 + g + 8 * w] : 0.0;   // x faces (3: x-, 1: x+) at (j = g, k = w)
  const double byl = (t < 2) ? sDel[(t == 0 ? 0 : 2) * N2 + g + 8 * w] : 0.0;   // y faces (0: y-, 2: y+) at (i = g, k = w)
  if (MODE == MODE_FLX) {
    double c0 = 0.0, c1 = 0.0;
    nd_dmma(c0, c1, Fx.x, bx0); nd_dmma(c0, c1, Fx.y, bx1); nd_dmma(c0, c1, axl, lwA);
    *reinterpret_cast<double2*>(V.out0 + gn) = make_double2(c0, c1);
    double d0 = 0.0, d1 = 0.0;
    nd_dmma(d0, d1, ay0, by0); nd_dmma(d0, d1, ay1, by1); nd_dmma(d0, d1, lwA, byl);
    *reinterpret_cast<double2*>(V.out1 + gn) = make_double2(d0, d1);
  } else {
    double c0 = 0.0, c1 = 0.0;
    if (MODE == MODE_TEND) { const double2 z = nd_ld2(sZ + 2 * t + 8 * g + KS_Z * w); c0 = z.x; c1 = z.y; }
    nd_dmma(c0, c1, Fx.x, bx0); nd_dmma(c0, c1, Fx.y, bx1);
    nd_dmma(c0, c1, ay0, by0); nd_dmma(c0, c1, ay1, by1);
    nd_dmma(c0, c1, axl, lwA); nd_dmma(c0, c1, lwA, byl);
    if (MODE == MODE_LAP) *reinterpret_cast<double2*>(V.out0 + gn) = make_double2(c0, c1);
    else {
      const double2 v = nd_ld2(V.var + gn);
      *reinterpret_cast<double2*>(V.var + gn) = make_double2(v.x + P.dt * c0, v.y + P.dt * c1);
    }
  }
}

// the shared part of a block: tables, face-node indices (issued first: their latency passes under the element's own loads), density
#define ND_P7_PROLOGUE                                                                                               \
  using namespace p7nd;                                                                                              \
  extern __shared__ __align__(16) double sm[];                                                                       \
  const int tid = threadIdx.x;                                                                                       \
  const int ke = blockIdx.x;                                                                                         \
  const size_t gn0 = size_t(ke) * N3 + 2 * (tid & 3) + 8 * ((tid & 31) >> 2) + 64 * (tid >> 5);                      \
  const size_t iPa = size_t(P.vmapP[size_t(ke) * NFT + tid]);                                                        \
  const size_t iPb = (tid < NFT - 256) ? size_t(P.vmapP[size_t(ke) * NFT + 256 + tid]) : 0;                          \
  if (tid < 64) sm[tid] = P.tab->D[tid];                                                                             \
  if (tid < 16) sm[64 + tid] = P.tab->Lw[tid];                                                                       \
  const double2 dd_ = nd_ld2(P.ddens + gn0), dh_ = nd_ld2(P.dens_hyd + gn0);                                         \
  const double2 rho_own = make_double2(dd_.x + dh_.x, dd_.y + dh_.y);

template <int MODE>
__global__ void __launch_bounds__(256, 5) numdiff_p7_kernel(const __grid_constant__ NumdiffParams P) {
  ND_P7_PROLOGUE
  NdVar V;
  V.in0 = P.in0; V.in1 = P.in1; V.in2 = P.in2; V.out0 = P.out0; V.out1 = P.out1; V.out2 = P.out2; V.var = P.var;
  V.varid = P.varid; V.dens_flag = P.dens_flag;
  nd_p7_body<MODE>(P, V, sm, iPa, iPb, rho_own);
}

// several variables of the same half-step in one launch: the density, the tables and the face indices are loaded once, and the
// fields of the next variable are pulled into L2 while the current one is worked on (one variable per launch left the kernel waiting
// for two dependent global latencies per element: profiles/r02_numdiff_p7_details.txt)
template <int MODE>
__global__ void __launch_bounds__(256, 5) numdiff_p7_multi_kernel(const __grid_constant__ NumdiffMulti M) {
  const NumdiffParams& P = M.P;
  ND_P7_PROLOGUE
  for (int iv = 0; iv < M.nvar; ++iv) {
    if (iv + 1 < M.nvar && tid < 3) {
      const NdVar& N = M.v[iv + 1];
      const double* f = (tid == 0) ? N.in0 : (tid == 1) ? N.in1 : N.in2;
      if (f && (tid == 0 || f != N.in0))
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(f + size_t(ke) * N3), "r"(uint32_t(N3 * sizeof(double))) : "memory");
    }
    nd_p7_body<MODE>(P, M.v[iv], sm, iPa, iPb, rho_own);
    __syncthreads();       // the staging areas are reused by the next variable
  }
}
}  // namespace

// several variables of one half-step (p = 7 only): FLX or TEND of M.nvar variables in one launch
void launch_numdiff_multi(int mode, const NumdiffMulti& M, cudaStream_t s) {
  const size_t shmem = size_t(p7nd::SM_DOUBLES) * sizeof(double);
  if (mode == MODE_FLX) numdiff_p7_multi_kernel<MODE_FLX><<<M.P.Ne, 256, shmem, s>>>(M);
  else if (mode == MODE_LAP) numdiff_p7_multi_kernel<MODE_LAP><<<M.P.Ne, 256, shmem, s>>>(M);
  else numdiff_p7_multi_kernel<MODE_TEND><<<M.P.Ne, 256, shmem, s>>>(M);
}

void launch_numdiff(int mode, const NumdiffParams& P, cudaStream_t s) {
  // p = 7: tensor-core kernel (FEDG_ND_KERNEL=1 selects the node-per-thread kernel: A/B runs, read at every launch)
  if (P.np == 8) {
    const char* e = getenv("FEDG_ND_KERNEL");
    if (!(e && e[0] == '1')) {
      const size_t shmem = size_t(p7nd::SM_DOUBLES) * sizeof(double);
      if (mode == MODE_FLX) numdiff_p7_kernel<MODE_FLX><<<P.Ne, 256, shmem, s>>>(P);
      else if (mode == MODE_LAP) numdiff_p7_kernel<MODE_LAP><<<P.Ne, 256, shmem, s>>>(P);
      else numdiff_p7_kernel<MODE_TEND><<<P.Ne, 256, shmem, s>>>(P);
      return;
    }
  }
  const size_t shmem = (size_t(P.np) * P.np + 2 * P.np + size_t(3) * P.Np + size_t(3) * P.NfpTot) * sizeof(double);
  if (mode == MODE_FLX) numdiff_kernel<MODE_FLX><<<P.Ne, P.Np, shmem, s>>>(P);
  else if (mode == MODE_LAP) numdiff_kernel<MODE_LAP><<<P.Ne, P.Np, shmem, s>>>(P);
  else numdiff_kernel<MODE_TEND><<<P.Ne, P.Np, shmem, s>>>(P);
}

}  // namespace fedg
