// TEST INFRASTRUCTURE ONLY (see fe_oracle.hpp).  Global (cubed-sphere) HEVI: one panel tile.
//
// Restates, under FElib/src:
//   common/scale_cubedsphere_coord_cnv.F90:735-787                 (CubedSphereCoordCnv_GetMetric)
//   mesh/scale_mesh_cubedspheredom3d.F90:527-565, 599-678          (coord_conv: the panel is a cube mesh in the central
//       angles (alpha, beta) and z; set_metric: GsqrtH, G_ij, GIJ, gam, Gsqrt = GsqrtH on the 3D nodes; fill_halo_metric:
//       halo metric = own face value)
//   fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_rhot_hevi_numflux.F90:606-834   (numflux_get_generalhvc)
//   fluid_dyn_solver/scale_atm_dyn_dgm_globalnonhydro3d_rhot_hevi.F90:337-583     (cal_tend)
//   fluid_dyn_solver/scale_atm_dyn_dgm_globalnonhydro3d_rhot_hevi.F90:873-1066    (cal_vi: the regional column solve
//       with GsqrtV = Gsqrt / (gam^2 GsqrtH), :965 -- see hevi_cal_vi in dyn_hevi.cpp)
// Scope: one panel tile whose lateral halo holds its own face values (the panel-edge exchange with index reversal and
// the lon-lat rotation of (MOMX, MOMY), data/scale_meshfieldcomm_cubedspheredom3d.F90, is not restated yet).
#include "fe_oracle.hpp"

#include <algorithm>
#include <stdexcept>

namespace feo {

void Mesh::init_cubedsphere_panel(const Element& e, int panel, int nex, int ney, int nez, const double* FZ, double ztop,
                                  double radius, bool shallow) {
  const double PI = 3.14159265358979323846;
  const bool per[3] = {false, false, false};
  init_cube(e, nex, ney, nez, -0.25 * PI, 0.25 * PI, -0.25 * PI, 0.25 * PI, 0.0, ztop, FZ, per);
  panelID = panel; RPlanet = radius; is_global = true;
  const int Np = e.Np, Nfp = e.Nfp, NfpTot = e.NfpTot;
  const size_t n2 = size_t(Nfp) * Ne2D;
  alpha2D.resize(n2); beta2D.resize(n2);
  Gij11.resize(n2); Gij12.resize(n2); Gij22.resize(n2); GIJ11.resize(n2); GIJ12.resize(n2); GIJ22.resize(n2);
  for (int ke2d = 0; ke2d < Ne2D; ++ke2d)
    for (int h = 0; h < Nfp; ++h) {
      const size_t i2 = size_t(h) + size_t(ke2d) * Nfp, i3 = size_t(h) + size_t(ke2d) * Np;   // bottom layer, k = 0 plane
      alpha2D[i2] = pos[0][i3]; beta2D[i2] = pos[1][i3];
      // GetMetric (cubedsphere_coord_cnv.F90:760-783)
      const double X = std::tan(alpha2D[i2]), Y = std::tan(beta2D[i2]);
      const double r2 = 1.0 + X * X + Y * Y, OnePlusX2 = 1.0 + X * X, OnePlusY2 = 1.0 + Y * Y;
      double fac = OnePlusX2 * OnePlusY2 * ((radius / r2) * (radius / r2));
      Gij11[i2] = fac * OnePlusX2; Gij12[i2] = -fac * (X * Y); Gij22[i2] = fac * OnePlusY2;
      GsqrtH[i2] = radius * radius * OnePlusX2 * OnePlusY2 / (r2 * std::sqrt(r2));
      fac = 1.0 / (GsqrtH[i2] * GsqrtH[i2]);
      GIJ11[i2] = fac * Gij22[i2]; GIJ12[i2] = -fac * Gij12[i2]; GIJ22[i2] = fac * Gij11[i2];
    }
  gam.assign(size_t(Np) * NeA, 1.0);
  for (int ke = 0; ke < Ne; ++ke)
    for (int p = 0; p < Np; ++p) {
      const size_t i = size_t(p) + size_t(ke) * Np;
      gam[i] = shallow ? 1.0 : 1.0 + pos[2][i] / radius;
      Gsqrt[i] = GsqrtH[(p % Nfp) + size_t(emap2d[ke]) * Nfp];
    }
  for (size_t f = 0; f < size_t(NfpTot) * Ne; ++f) {   // fill_halo_metric
    const int iM = vmapM[f], iP = vmapP[f];
    if (iP >= Np * Ne) { Gsqrt[iP] = Gsqrt[iM]; gam[iP] = gam[iM]; }
  }
}

// hevi_numflux.F90:606-834
// hevi = true: rhot_hevi_numflux.F90:606-834;  hevi = false: rhot_heve_numflux.F90:1543-1772 (full dissipation, full normal
// velocity in the mass / theta fluxes, vertical pressure term in MOMZ)
void global_numflux_generalhvc(const Element& e, const Mesh& m, const Consts& c, const DynState& s, bool hevi, vec& del_flux) {
  const int NfpTot = e.NfpTot, Nfp = e.Nfp, np = e.np;
  const double gamm = c.CPdry / c.CVdry;
  del_flux.resize(size_t(NfpTot) * PRGVAR_NUM * m.Ne);
#pragma omp parallel for
  for (int ke = 0; ke < m.Ne; ++ke) {
    const int ke2d = m.emap2d[ke];
    double* df = &del_flux[size_t(ke) * NfpTot * PRGVAR_NUM];
    for (int fp = 0; fp < NfpTot; ++fp) {
      const size_t f = size_t(fp) + size_t(ke) * NfpTot;
      const int id[2] = {m.vmapM[f], m.vmapP[f]};
      const double nx = m.nx[f], ny = m.ny[f], nz = m.nz[f];
      const int h2d = (id[0] - ke * e.Np) % Nfp;           // IndexH2Dto3D_bnd: horizontal node of the own face node
      (void)np;
      const size_t i2 = size_t(h2d) + size_t(ke2d) * Nfp;
      const double G11 = m.GIJ11[i2], G12 = m.GIJ12[i2], G22 = m.GIJ22[i2], GsH = m.GsqrtH[i2];
      double Gs[2], rgam2[2], RGv[2], G13[2], G23[2], gDD[2], gMX[2], gMY[2], gMZ[2], gDR[2], Phyd[2], dp[2], gDens[2], gRhot[2],
          Velh[2], Vel[2], Gxz[2], Gyz[2], G1n[2], G2n[2], Gnn[2];
      for (int t = 0; t < 2; ++t) {
        const int i = id[t];
        Gs[t] = m.Gsqrt[i];
        rgam2[t] = 1.0 / (m.gam[i] * m.gam[i]);
        const double GsV = Gs[t] * rgam2[t] / GsH;
        RGv[t] = 1.0 / GsV;
        G13[t] = m.G13[i]; G23[t] = m.G23[i];
        gDD[t] = Gs[t] * s.DDENS[i]; gMX[t] = Gs[t] * s.MOMX[i]; gMY[t] = Gs[t] * s.MOMY[i];
        gMZ[t] = Gs[t] * s.MOMZ[i]; gDR[t] = Gs[t] * s.DRHOT[i];
        Phyd[t] = s.PRES_hyd[i]; dp[t] = s.DPRES[i];
        gDens[t] = gDD[t] + Gs[t] * s.DENS_hyd[i];
        gRhot[t] = Gs[t] * s.THERM_hyd[i] + gDR[t];
        Velh[t] = (gMX[t] * nx + gMY[t] * ny) / gDens[t];
        Vel[t] = Velh[t] + (((gMZ[t] * RGv[t] + G13[t] * gMX[t] + G23[t] * gMY[t]) * nz)) / gDens[t];
        Gxz[t] = rgam2[t] * (G11 * G13[t] + G12 * G23[t]);
        Gyz[t] = rgam2[t] * (G12 * G13[t] + G22 * G23[t]);
        G1n[t] = rgam2[t] * (G11 * nx + G12 * ny);
        G2n[t] = rgam2[t] * (G12 * nx + G22 * ny);
      }
      const double tmp1 = std::fabs(G11 * nx) + std::fabs(G22 * ny);
      for (int t = 0; t < 2; ++t)
        Gnn[t] = rgam2[t] * tmp1 + (1.0 * (RGv[t] * RGv[t]) + G13[t] * Gxz[t] + G23[t] * Gyz[t]) * std::fabs(nz);
      const double swV = hevi ? 1.0 - nz * nz : 1.0;
      const double alpha = swV * std::max(std::sqrt(Gnn[0] * gamm * (Phyd[0] + dp[0]) * Gs[0] / gDens[0]) + std::fabs(Vel[0]),
                                          std::sqrt(Gnn[1] * gamm * (Phyd[1] + dp[1]) * Gs[1] / gDens[1]) + std::fabs(Vel[1]));
      const double hf = m.Fscale[f] * 0.5;
      const double* Vm = hevi ? Velh : Vel;
      df[fp + DENS_VID * NfpTot] = hf * (gDens[1] * Vm[1] - gDens[0] * Vm[0] + (-alpha * (gDD[1] - gDD[0])));
      df[fp + RHOT_VID * NfpTot] = hf * (gRhot[1] * Vm[1] - gRhot[0] * Vm[0] + (-alpha * (gDR[1] - gDR[0])));
      const double t3 = Gs[1] * dp[1], t4 = Gs[0] * dp[0];
      const double mz = hevi ? 0.0 : (t3 * RGv[1] - t4 * RGv[0]) * nz;
      df[fp + MOMZ_VID * NfpTot] = hf * (gMZ[1] * Vel[1] - gMZ[0] * Vel[0] + mz + (-alpha * (gMZ[1] - gMZ[0])));
      const double mx = (G1n[1] + Gxz[1] * nz) * t3 - (G1n[0] + Gxz[0] * nz) * t4;
      const double my = (G2n[1] + Gyz[1] * nz) * t3 - (G2n[0] + Gyz[0] * nz) * t4;
      df[fp + MOMX_VID * NfpTot] = hf * (gMX[1] * Vel[1] - gMX[0] * Vel[0] + mx + (-alpha * (gMX[1] - gMX[0])));
      df[fp + MOMY_VID * NfpTot] = hf * (gMY[1] * Vel[1] - gMY[0] * Vel[0] + my + (-alpha * (gMY[1] - gMY[0])));
    }
  }
}

// globalnonhydro3d_rhot_hevi.F90:337-583
// hevi = true: globalnonhydro3d_rhot_hevi.F90:337-583;  hevi = false: globalnonhydro3d_rhot_heve.F90:338-600 (cal_tend_shallow_atm)
void global_cal_tend(const Element& e, const Mesh& m, const Consts& c, const DynState& s, bool hevi, double* dt5[5]) {
  const int Np = e.Np, NfpTot = e.NfpTot, Nfp = e.Nfp;
  vec del_flux;
  global_numflux_generalhvc(e, m, c, s, hevi, del_flux);
  double sgn = 1.0;
  bool is_panel1to4 = true;
  if (m.panelID == 5) is_panel1to4 = false;
  else if (m.panelID == 6) { is_panel1to4 = false; sgn = -1.0; }
  const double OHM = c.OHM;
#pragma omp parallel
  {
    vec Flux(size_t(Np) * 3 * 5, 0.0), DFlux(size_t(Np) * 4 * 5), RGsqrtV(Np), RGsqrt(Np), RDENS(Np), G11(Np), G12(Np), G22(Np), drho(Np, 0.0);
#pragma omp for
    for (int ke = 0; ke < m.Ne; ++ke) {
      const int ke2d = m.emap2d[ke];
      const size_t o = size_t(ke) * Np;
      auto F = [&](int p, int d, int v) -> double& { return Flux[p + Np * (d + 3 * v)]; };
      auto DF = [&](int p, int d, int v) -> double& { return DFlux[p + Np * (d + 4 * v)]; };
      for (int p = 0; p < Np; ++p) {
        const size_t i2 = size_t(p % Nfp) + size_t(ke2d) * Nfp;
        const double Rgam2 = 1.0 / (m.gam[o + p] * m.gam[o + p]);
        G11[p] = m.GIJ11[i2] * Rgam2; G12[p] = m.GIJ12[i2] * Rgam2; G22[p] = m.GIJ22[i2] * Rgam2;
        const double GsqrtV = m.Gsqrt[o + p] * Rgam2 / m.GsqrtH[i2];
        RGsqrtV[p] = 1.0 / GsqrtV;
        RGsqrt[p] = 1.0 / m.Gsqrt[o + p];
        RDENS[p] = 1.0 / (s.DDENS[o + p] + s.DENS_hyd[o + p]);
      }
      for (int p = 0; p < Np; ++p) {
        const double G = m.Gsqrt[o + p];
        F(p, 0, DENS_VID) = G * s.MOMX[o + p];
        F(p, 1, DENS_VID) = G * s.MOMY[o + p];
        F(p, 2, DENS_VID) = G * (s.MOMZ[o + p] * RGsqrtV[p] + m.G13[o + p] * s.MOMX[o + p] + m.G23[o + p] * s.MOMY[o + p]);
      }
      for (int p = 0; p < Np; ++p) {
        const double pt = (s.THERM_hyd[o + p] + s.DRHOT[o + p]) * RDENS[p];
        F(p, 0, RHOT_VID) = F(p, 0, DENS_VID) * pt;
        F(p, 1, RHOT_VID) = F(p, 1, DENS_VID) * pt;
        F(p, 2, RHOT_VID) = hevi ? 0.0 : F(p, 2, DENS_VID) * pt;   // HEVI: not set in the reference (:474), derivative unused
        const double w = s.MOMZ[o + p] * RDENS[p];
        F(p, 0, MOMZ_VID) = F(p, 0, DENS_VID) * w;
        F(p, 1, MOMZ_VID) = F(p, 1, DENS_VID) * w;
        F(p, 2, MOMZ_VID) = F(p, 2, DENS_VID) * w + (hevi ? 0.0 : m.Gsqrt[o + p] * RGsqrtV[p] * s.DPRES[o + p]);
      }
      for (int p = 0; p < Np; ++p) {
        const double GP = m.Gsqrt[o + p] * s.DPRES[o + p];
        const double u = s.MOMX[o + p] * RDENS[p], v = s.MOMY[o + p] * RDENS[p];
        F(p, 0, MOMX_VID) = F(p, 0, DENS_VID) * u + G11[p] * GP;
        F(p, 1, MOMX_VID) = F(p, 1, DENS_VID) * u + G12[p] * GP;
        F(p, 2, MOMX_VID) = F(p, 2, DENS_VID) * u + GP * (G11[p] * m.G13[o + p] + G12[p] * m.G23[o + p]);
        F(p, 0, MOMY_VID) = F(p, 0, DENS_VID) * v + G12[p] * GP;
        F(p, 1, MOMY_VID) = F(p, 1, DENS_VID) * v + G22[p] * GP;
        F(p, 2, MOMY_VID) = F(p, 2, DENS_VID) * v + GP * (G12[p] * m.G13[o + p] + G22[p] * m.G23[o + p]);
      }
      for (int v = 0; v < 5; ++v)
        op_div(e, &Flux[size_t(Np) * 3 * v], &del_flux[(size_t(ke) * PRGVAR_NUM + v) * NfpTot], &DFlux[size_t(Np) * 4 * v]);
      if (!hevi) op_matz(e, e.VPOrdM1.data(), &s.DDENS[o], drho.data());     // VFilterPM1 (rhot_heve.F90:507-508)
      for (int p = 0; p < Np; ++p) {
        const double E11 = m.E11[o + p], E22 = m.E22[o + p], E33 = m.E33[o + p];
        if (hevi) {
          dt5[DENS_VID][o + p] = -(E11 * DF(p, 0, DENS_VID) + E22 * DF(p, 1, DENS_VID) + DF(p, 3, DENS_VID)) * RGsqrt[p];
          dt5[RHOT_VID][o + p] = -(E11 * DF(p, 0, RHOT_VID) + E22 * DF(p, 1, RHOT_VID) + DF(p, 3, RHOT_VID)) * RGsqrt[p];
          dt5[MOMZ_VID][o + p] =
              -(E11 * DF(p, 0, MOMZ_VID) + E22 * DF(p, 1, MOMZ_VID) + E33 * DF(p, 2, MOMZ_VID) + DF(p, 3, MOMZ_VID)) * RGsqrt[p];
        } else {
          dt5[DENS_VID][o + p] = -(E11 * DF(p, 0, DENS_VID) + E22 * DF(p, 1, DENS_VID) + E33 * DF(p, 2, DENS_VID) + DF(p, 3, DENS_VID)) * RGsqrt[p];
          dt5[RHOT_VID][o + p] = -(E11 * DF(p, 0, RHOT_VID) + E22 * DF(p, 1, RHOT_VID) + E33 * DF(p, 2, RHOT_VID) + DF(p, 3, RHOT_VID)) * RGsqrt[p];
          dt5[MOMZ_VID][o + p] =
              -(E11 * DF(p, 0, MOMZ_VID) + E22 * DF(p, 1, MOMZ_VID) + E33 * DF(p, 2, MOMZ_VID) + DF(p, 3, MOMZ_VID)) * RGsqrt[p] - c.GRAV * drho[p];
        }
        const size_t i2 = size_t(p % Nfp) + size_t(ke2d) * Nfp;
        const double X = std::tan(m.alpha2D[i2]), Y = std::tan(m.beta2D[i2]);
        const double twoOVdel2 = 2.0 / (1.0 + X * X + Y * Y);
        const double MX = s.MOMX[o + p], MY = s.MOMY[o + p];
        double CORI1 = sgn * OHM * twoOVdel2 * (-X * Y * MX + (1.0 + Y * Y) * MY);
        double CORI2 = sgn * OHM * twoOVdel2 * (-(1.0 + X * X) * MX + X * Y * MY);
        if (is_panel1to4) { CORI1 = sgn * Y * CORI1; CORI2 = sgn * Y * CORI2; }
        const double u = MX * RDENS[p], v = MY * RDENS[p];
        double mx = -(G11[p] * s.DPhydDx[o + p] + G12[p] * s.DPhydDy[o + p]) - twoOVdel2 * Y * (X * Y * u - (1.0 + Y * Y) * v) * MX + CORI1;
        double my = -(G12[p] * s.DPhydDx[o + p] + G22[p] * s.DPhydDy[o + p]) - twoOVdel2 * X * (-(1.0 + X * X) * u + X * Y * v) * MY + CORI2;
        dt5[MOMX_VID][o + p] =
            mx - (E11 * DF(p, 0, MOMX_VID) + E22 * DF(p, 1, MOMX_VID) + E33 * DF(p, 2, MOMX_VID) + DF(p, 3, MOMX_VID)) * RGsqrt[p];
        dt5[MOMY_VID][o + p] =
            my - (E11 * DF(p, 0, MOMY_VID) + E22 * DF(p, 1, MOMY_VID) + E33 * DF(p, 2, MOMY_VID) + DF(p, 3, MOMY_VID)) * RGsqrt[p];
      }
    }
  }
}

}  // namespace feo

namespace feo {
void global_hevi_numflux_generalhvc(const Element& e, const Mesh& m, const Consts& c, const DynState& s, vec& del_flux) {
  global_numflux_generalhvc(e, m, c, s, true, del_flux);
}
void global_hevi_cal_tend(const Element& e, const Mesh& m, const Consts& c, const DynState& s, double* dt5[5]) { global_cal_tend(e, m, c, s, true, dt5); }
}  // namespace feo
