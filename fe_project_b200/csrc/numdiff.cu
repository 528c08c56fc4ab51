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
__global__ void numdiff_kernel(const __grid_constant__ NumdiffParams P) {
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
}  // namespace

void launch_numdiff(int mode, const NumdiffParams& P, cudaStream_t s) {
  const size_t shmem = (size_t(P.np) * P.np + 2 * P.np + size_t(3) * P.Np + size_t(3) * P.NfpTot) * sizeof(double);
  if (mode == MODE_FLX) numdiff_kernel<MODE_FLX><<<P.Ne, P.Np, shmem, s>>>(P);
  else if (mode == MODE_LAP) numdiff_kernel<MODE_LAP><<<P.Ne, P.Np, shmem, s>>>(P);
  else numdiff_kernel<MODE_TEND><<<P.Ne, P.Np, shmem, s>>>(P);
}

}  // namespace fedg
