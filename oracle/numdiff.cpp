// TEST INFRASTRUCTURE ONLY (see fe_oracle.hpp).  Numerical diffusion of the prognostic variables.
//
// Restates, under FElib/src/fluid_dyn_solver:
//   scale_atm_dyn_dgm_nonhydro3d_numdiff.F90:214-241   (Apply: exchange, then THERM, MOMZ, MOMX, MOMY, DENS)
//   scale_atm_dyn_dgm_nonhydro3d_numdiff.F90:245-376   (apply_numfilter: (-1)^(n+1) coef * Laplacian^n, LDG alternating fluxes)
//   scale_atm_dyn_dgm_nonhydro3d_numdiff.F90:379-502   (numdiff_tend, cal_del_flux_lap_with_coef)
//   scale_atm_dyn_dgm_nonhydro3d_numdiff.F90:505-597   (numdiff_cal_laplacian, cal_del_flux_lap)
//   scale_atm_dyn_dgm_nonhydro3d_numdiff.F90:600-720   (numdiff_cal_flx, cal_del_gradDiffVar)
//   scale_atm_dyn_dgm_bnd.F90:370-508                  (ApplyBC_numdiff_odd_lc / _even_lc)
// Called after the dynamics step (model/atm_nonhydro3d/src/atmos/mod_atmos_dyn.F90:343-349).
#include "fe_oracle.hpp"

namespace feo {

namespace {
inline double sgn(double x) { return x >= 0.0 ? 1.0 : -1.0; }   // Fortran sign(1.0, x); the normals are never -0.0

struct BcView {
  const Element& e; const Mesh& m; const NumdiffCfg& cfg;
  // boundary-condition ids of the halo slot behind face node f (0 when the face is not a physical boundary)
  void ids(size_t f, int& vel, int& therm) const {
    vel = therm = 0;
    const int i_ = m.vmapP[f] - e.Np * m.Ne;
    if (i_ < 0) return;
    int face = 0;
    while (i_ >= m.halo_off[face + 1]) ++face;
    if (m.nbr_face[face] != face) return;
    vel = cfg.vel_bc[face]; therm = cfg.therm_bc[face];
  }
};

// bnd.F90:445-508
void bc_even(const BcView& b, double* var, int varid, std::vector<char>& is_bound) {
  const Element& e = b.e; const Mesh& m = b.m;
  const size_t nf = size_t(e.NfpTot) * m.Ne;
  is_bound.assign(nf, 0);
  for (size_t f = 0; f < nf; ++f) {
    const int iP = m.vmapP[f];
    if (iP - e.Np * m.Ne < 0) continue;
    int vel, therm; b.ids(f, vel, therm);
    const int iM = m.vmapM[f];
    if (vel == 2) {
      if (varid == MOMX_VID) var[iP] = var[iM] - 2.0 * (var[iM] * m.nx[f]) * m.nx[f];
      else if (varid == MOMY_VID) var[iP] = var[iM] - 2.0 * (var[iM] * m.ny[f]) * m.ny[f];
      else if (varid == MOMZ_VID) var[iP] = var[iM] - 2.0 * (var[iM] * m.nz[f]) * m.nz[f];
      is_bound[f] = 1;
    } else if (vel == 3) {
      if (varid == MOMX_VID || varid == MOMY_VID || varid == MOMZ_VID) var[iP] = -var[iM];
      is_bound[f] = 1;
    }
  }
}
// bnd.F90:370-442
void bc_odd(const BcView& b, double* gx, double* gy, double* gz, int varid, std::vector<char>& is_bound) {
  const Element& e = b.e; const Mesh& m = b.m;
  const size_t nf = size_t(e.NfpTot) * m.Ne;
  is_bound.assign(nf, 0);
  for (size_t f = 0; f < nf; ++f) {
    const int iP = m.vmapP[f];
    if (iP - e.Np * m.Ne < 0) continue;
    int vel, therm; b.ids(f, vel, therm);
    const int iM = m.vmapM[f];
    const double gn = gx[iM] * m.nx[f] + gy[iM] * m.ny[f] + gz[iM] * m.nz[f];
    if (vel == 2) {
      if (varid == MOMX_VID) { gy[iP] = gy[iM] - 2.0 * gn * m.ny[f]; gz[iP] = gz[iM] - 2.0 * gn * m.nz[f]; }
      else if (varid == MOMY_VID) { gx[iP] = gx[iM] - 2.0 * gn * m.nx[f]; gz[iP] = gz[iM] - 2.0 * gn * m.nz[f]; }
      else if (varid == MOMZ_VID) { gx[iP] = gx[iM] - 2.0 * gn * m.nx[f]; gy[iP] = gy[iM] - 2.0 * gn * m.ny[f]; }
      is_bound[f] = 1;
    }
    if (therm == 1) {   // BND_TYPE_ADIABAT
      if (varid == DENS_VID || varid == RHOT_VID) {
        gx[iP] = gx[iM] - 2.0 * gn * m.nx[f]; gy[iP] = gy[iM] - 2.0 * gn * m.ny[f]; gz[iP] = gz[iM] - 2.0 * gn * m.nz[f];
      }
      is_bound[f] = 1;
    }
  }
}

// numdiff.F90:600-720
void cal_flx(const Element& e, const Mesh& m, const double* varh, const double* varv, const double* ddens, const double* dens_hyd,
             const std::vector<char>& is_bound, bool divide_dens, double* gx, double* gy, double* gz) {
  const int Np = e.Np, NfpTot = e.NfpTot;
#pragma omp parallel for
  for (int ke = 0; ke < m.Ne; ++ke) {
    vec d1(NfpTot), d2(NfpTot), d3(NfpTot), vh(Np), vv(Np), F(Np), L(Np), fl(NfpTot);
    for (int p = 0; p < NfpTot; ++p) {
      const size_t f = size_t(ke) * NfpTot + p;
      const int iM = m.vmapM[f], iP = m.vmapP[f];
      double wP = 1.0, wM = 1.0;
      if (divide_dens) { wP = 1.0 / (ddens[iP] + dens_hyd[iP]); wM = 1.0 / (ddens[iM] + dens_hyd[iM]); }
      const double dh = 0.5 * (varh[iP] * wP - varh[iM] * wM), dv = 0.5 * (varv[iP] * wP - varv[iM] * wM);
      if (is_bound[f]) { d1[p] = dh * m.nx[f]; d2[p] = dh * m.ny[f]; d3[p] = dv * m.nz[f]; }
      else {
        d1[p] = (1.0 - sgn(m.nx[f])) * dh * m.nx[f]; d2[p] = (1.0 - sgn(m.ny[f])) * dh * m.ny[f]; d3[p] = (1.0 - sgn(m.nz[f])) * dv * m.nz[f];
      }
    }
    const size_t o = size_t(ke) * Np;
    for (int n = 0; n < Np; ++n) {
      if (divide_dens) { vh[n] = varh[o + n] / (ddens[o + n] + dens_hyd[o + n]); vv[n] = varv[o + n] / (ddens[o + n] + dens_hyd[o + n]); }
      else { vh[n] = varh[o + n]; vv[n] = varv[o + n]; }
    }
    auto lift = [&](const vec& d) { for (int p = 0; p < NfpTot; ++p) fl[p] = m.Fscale[size_t(ke) * NfpTot + p] * d[p]; op_lift(e, fl.data(), L.data()); };
    op_dx(e, vh.data(), F.data()); lift(d1);
    for (int n = 0; n < Np; ++n) gx[o + n] = m.E11[o + n] * F[n] + L[n];
    op_dy(e, vh.data(), F.data()); lift(d2);
    for (int n = 0; n < Np; ++n) gy[o + n] = m.E22[o + n] * F[n] + L[n];
    op_dz(e, vv.data(), F.data()); lift(d3);
    for (int n = 0; n < Np; ++n) gz[o + n] = m.E33[o + n] * F[n] + L[n];
  }
}

// numdiff.F90:505-597
void cal_laplacian(const Element& e, const Mesh& m, const double* gx, const double* gy, const double* gz, const std::vector<char>& is_bound,
                   double* lap_h, double* lap_v) {
  const int Np = e.Np, NfpTot = e.NfpTot;
#pragma omp parallel for
  for (int ke = 0; ke < m.Ne; ++ke) {
    vec dh(NfpTot), dv(NfpTot), Fx(Np), Fy(Np), Fz(Np), L(Np), fl(NfpTot);
    for (int p = 0; p < NfpTot; ++p) {
      const size_t f = size_t(ke) * NfpTot + p;
      const int iM = m.vmapM[f], iP = m.vmapP[f];
      if (is_bound[f]) {
        dh[p] = 0.5 * ((gx[iP] - gx[iM]) * m.nx[f] + (gy[iP] - gy[iM]) * m.ny[f]);
        dv[p] = 0.5 * (gz[iP] - gz[iM]) * m.nz[f];
      } else {
        dh[p] = 0.5 * ((1.0 + sgn(m.nx[f])) * (gx[iP] - gx[iM]) * m.nx[f] + (1.0 + sgn(m.ny[f])) * (gy[iP] - gy[iM]) * m.ny[f]);
        dv[p] = 0.5 * (1.0 + sgn(m.nz[f])) * (gz[iP] - gz[iM]) * m.nz[f];
      }
    }
    const size_t o = size_t(ke) * Np;
    op_dx(e, gx + o, Fx.data()); op_dy(e, gy + o, Fy.data());
    for (int p = 0; p < NfpTot; ++p) fl[p] = m.Fscale[size_t(ke) * NfpTot + p] * dh[p];
    op_lift(e, fl.data(), L.data());
    for (int n = 0; n < Np; ++n) lap_h[o + n] = (m.E11[o + n] * Fx[n] + m.E22[o + n] * Fy[n] + L[n]);
    op_dz(e, gz + o, Fz.data());
    for (int p = 0; p < NfpTot; ++p) fl[p] = m.Fscale[size_t(ke) * NfpTot + p] * dv[p];
    op_lift(e, fl.data(), L.data());
    for (int n = 0; n < Np; ++n) lap_v[o + n] = (m.E33[o + n] * Fz[n] + L[n]);
  }
}

// numdiff.F90:379-502
void cal_tend(const Element& e, const Mesh& m, const double* gx, const double* gy, const double* gz, const double* ddens,
              const double* dens_hyd, double coef_h, double coef_v, const std::vector<char>& is_bound, bool mul_dens, double* tend) {
  const int Np = e.Np, NfpTot = e.NfpTot;
#pragma omp parallel for
  for (int ke = 0; ke < m.Ne; ++ke) {
    vec df(NfpTot), in(Np), Fx(Np), Fy(Np), Fz(Np), L(Np), fl(NfpTot);
    for (int p = 0; p < NfpTot; ++p) {
      const size_t f = size_t(ke) * NfpTot + p;
      const int iM = m.vmapM[f], iP = m.vmapP[f];
      double wM = 0.5, wP = 0.5;
      if (mul_dens) { wM = 0.5 * (dens_hyd[iM] + ddens[iM]); wP = 0.5 * (dens_hyd[iP] + ddens[iP]); }
      if (is_bound[f])
        df[p] = coef_h * (wP * gx[iP] - wM * gx[iM]) * m.nx[f] + coef_h * (wP * gy[iP] - wM * gy[iM]) * m.ny[f] +
                coef_v * (wP * gz[iP] - wM * gz[iM]) * m.nz[f];
      else
        df[p] = (1.0 + sgn(m.nx[f])) * coef_h * (wP * gx[iP] - wM * gx[iM]) * m.nx[f] +
                (1.0 + sgn(m.ny[f])) * coef_h * (wP * gy[iP] - wM * gy[iM]) * m.ny[f] +
                (1.0 + sgn(m.nz[f])) * coef_v * (wP * gz[iP] - wM * gz[iM]) * m.nz[f];
    }
    const size_t o = size_t(ke) * Np;
    auto coef = [&](int n, double c) { return mul_dens ? c * (dens_hyd[o + n] + ddens[o + n]) : c; };
    for (int n = 0; n < Np; ++n) in[n] = coef(n, coef_h) * gx[o + n];
    op_dx(e, in.data(), Fx.data());
    for (int n = 0; n < Np; ++n) in[n] = coef(n, coef_h) * gy[o + n];
    op_dy(e, in.data(), Fy.data());
    for (int n = 0; n < Np; ++n) in[n] = coef(n, coef_v) * gz[o + n];
    op_dz(e, in.data(), Fz.data());
    for (int p = 0; p < NfpTot; ++p) fl[p] = m.Fscale[size_t(ke) * NfpTot + p] * df[p];
    op_lift(e, fl.data(), L.data());
    for (int n = 0; n < Np; ++n) tend[o + n] = (m.E11[o + n] * Fx[n] + m.E22[o + n] * Fy[n] + m.E33[o + n] * Fz[n] + L[n]);
  }
}
}  // namespace

// numdiff.F90:245-376 for one variable
static void apply_numfilter(const Element& e, const Mesh& m, const NumdiffCfg& cfg, DynState& s, int varid) {
  const size_t N = size_t(e.Np) * m.NeA, nint = size_t(e.Np) * m.Ne;
  const double nd_sign = ((cfg.laplacian_num + 1) % 2 == 0) ? 1.0 : -1.0;      // (-1)**mod(n+1, 2)
  const bool dens_weight = varid != DENS_VID;
  double* var = s.prog(varid);
  vec gx(N, 0.0), gy(N, 0.0), gz(N, 0.0), lh(N, 0.0), lv(N, 0.0), tend(N, 0.0);
  std::vector<char> is_bound;
  BcView b{e, m, cfg};
  bc_even(b, var, varid, is_bound);
  cal_flx(e, m, var, var, s.DDENS.data(), s.DENS_hyd.data(), is_bound, dens_weight, gx.data(), gy.data(), gz.data());
  for (vec* g : {&gx, &gy, &gz}) m.exchange_halo(e, g->data());
  for (int it = 1; it <= cfg.laplacian_num - 1; ++it) {
    bc_odd(b, gx.data(), gy.data(), gz.data(), varid, is_bound);
    cal_laplacian(e, m, gx.data(), gy.data(), gz.data(), is_bound, lh.data(), lv.data());
    m.exchange_halo(e, lh.data()); m.exchange_halo(e, lv.data());
    bc_even(b, lh.data(), varid, is_bound);
    cal_flx(e, m, lh.data(), lv.data(), s.DDENS.data(), s.DENS_hyd.data(), is_bound, false, gx.data(), gy.data(), gz.data());
    for (vec* g : {&gx, &gy, &gz}) m.exchange_halo(e, g->data());
  }
  bc_odd(b, gx.data(), gy.data(), gz.data(), varid, is_bound);
  cal_tend(e, m, gx.data(), gy.data(), gz.data(), s.DDENS.data(), s.DENS_hyd.data(), nd_sign * cfg.coef_h, nd_sign * cfg.coef_v, is_bound,
           dens_weight, tend.data());
  for (size_t i = 0; i < nint; ++i) var[i] = var[i] + cfg.dt * tend[i];
}

// numdiff.F90:214-241
void numdiff_apply(const Element& e, const Mesh& m, const NumdiffCfg& cfg, DynState& s) {
  for (int v = 0; v < 5; ++v) m.exchange_halo(e, s.prog(v));
  for (int varid : {THERM_VID, MOMZ_VID, MOMX_VID, MOMY_VID, DENS_VID}) apply_numfilter(e, m, cfg, s, varid);
}

}  // namespace feo
