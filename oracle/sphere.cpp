// TEST INFRASTRUCTURE ONLY (see fe_oracle.hpp).  The whole cubed sphere: six panel tiles (one per "rank" of the reference's
// Nprc = 6 layout) and the panel-edge halo exchange.
//
// Restates, under FElib/src:
//   mesh/scale_meshutil_cubedsphere2d.F90:251-290   (getPanelConnectivity: neighbour panel, destination face, sign = revert)
//   data/scale_meshfieldcomm_base.F90:870-895        (same-rank exchange: boundary data of (tile, face) -> recv buffer of
//                                                     face |s_faceID| of tile s_tileID)
//   data/scale_meshfieldcomm_cubedspheredom3d.F90:226-420, 492-540  (exchange: CS -> lon-lat on the sender's face nodes,
//                                                     revert_hori, lon-lat -> CS on the receiver's face nodes)
//   common/scale_cubedsphere_coord_cnv.F90:150-236, 314-401          (CS2LonLatVec, LonLat2CSVec; gam = 1)
//   fluid_dyn_solver/scale_atm_dyn_dgm_driver_nonhydro3d.F90:614-963 (Update with several local meshes)
#include <stdexcept>

#include "fe_oracle.hpp"

namespace feo {

namespace {
const double EPS = 2.220446e-16;

void panel_connectivity(int pc[4][6], int fc[4][6]) {
  const int zonal[6] = {4, 1, 2, 3, 4, 1};
  for (int n = 1; n <= 4; ++n) {
    pc[0][n - 1] = 6;
    fc[0][n - 1] = zonal[4 - n] > 2 ? zonal[4 - n] : -zonal[4 - n];
    pc[1][n - 1] = zonal[n + 1]; fc[1][n - 1] = 4;
    pc[2][n - 1] = 5; fc[2][n - 1] = n > 2 ? -n : n;
    pc[3][n - 1] = zonal[n - 1]; fc[3][n - 1] = 2;
  }
  const int p5[4] = {1, 2, 3, 4}, f5[4] = {3, 3, -3, -3}, p6[4] = {3, 2, 1, 4}, f6[4] = {-1, -1, 1, 1};
  for (int f = 0; f < 4; ++f) { pc[f][4] = p5[f]; fc[f][4] = f5[f]; pc[f][5] = p6[f]; fc[f][5] = f6[f]; }
}

double cos_lat(int panel, double X, double Y, double a, double b) {
  if (panel <= 4) return std::cos(std::atan(std::tan(b) * std::cos(a)));
  const double s = panel == 5 ? 1.0 : -1.0;
  return std::cos(std::atan(s / std::max(std::sqrt(X * X + Y * Y), EPS)));
}
// cubedsphere_coord_cnv.F90:150-236
void cs2lonlat_vec(int panel, double a, double b, double R, double va, double vb, double& vlon, double& vlat) {
  const double X = std::tan(a), Y = std::tan(b), del2 = 1.0 + X * X + Y * Y, cl = cos_lat(panel, X, Y, a, b);
  if (panel <= 4) {
    vlon = va * cl * R;
    vlat = (-X * Y * va + (1.0 + Y * Y) * vb) * R * std::sqrt(1.0 + X * X) / del2;
  } else {
    const double r = panel == 5 ? R : -R;
    vlon = (-Y * (1.0 + X * X) * va + X * (1.0 + Y * Y) * vb) * r / std::max(X * X + Y * Y, EPS) * cl;
    vlat = (-X * (1.0 + X * X) * va - Y * (1.0 + Y * Y) * vb) * r / (del2 * std::max(std::sqrt(X * X + Y * Y), EPS));
  }
}
// cubedsphere_coord_cnv.F90:314-401
void lonlat2cs_vec(int panel, double a, double b, double R, double vlon, double vlat, double& va, double& vb) {
  const double X = std::tan(a), Y = std::tan(b), del2 = 1.0 + X * X + Y * Y;
  const double uc = vlon / cos_lat(panel, X, Y, a, b);
  if (panel <= 4) {
    va = uc / R;
    vb = (X * Y * uc + del2 / std::sqrt(1.0 + X * X) * vlat) / (R * (1.0 + Y * Y));
  } else {
    const double r = panel == 5 ? R : -R, sq = std::sqrt(std::max(del2 - 1.0, EPS));
    va = (-Y * uc - del2 * X / sq * vlat) / (r * (1.0 + X * X));
    vb = (X * uc - del2 * Y / sq * vlat) / (r * (1.0 + Y * Y));
  }
}
}  // namespace

// One exchange of the listed scalar fields and of one horizontal vector (u1, u2) over the six panels.
// scal[p][k] / u1[p] / u2[p]: field arrays (Np*NeA) of panel p (0-based).
void sphere_exchange(const Element& e, Mesh* const mesh[6], const std::vector<double*> scal[6], double* const u1[6], double* const u2[6]) {
  int pc[4][6], fc[4][6];
  panel_connectivity(pc, fc);
  const int np = e.np, Nfp = e.Nfp;
  for (int T = 0; T < 6; ++T) {
    const Mesh& mT = *mesh[T];
    for (int f = 0; f < 4; ++f) {
      const int U = pc[f][T] - 1, g = std::abs(fc[f][T]) - 1;
      const bool rev = fc[f][T] < 0;
      Mesh& mU = *mesh[U];
      const int cnt = mT.halo_off[f + 1] - mT.halo_off[f];
      if (cnt != mU.halo_off[g + 1] - mU.halo_off[g]) throw std::runtime_error("panel faces do not match (NeX must equal NeY)");
      const int nez = mT.NeZ, nex = cnt / (Nfp * nez);
      const size_t nintU = size_t(e.Np) * mU.Ne;
      for (int m = 0; m < cnt; ++m) {
        int ms = m;
        if (rev) {   // revert_hori: buffer (Nnode_h1D, Nnode_v, NeX, NeZ)
          const int p1 = m % np, p3 = (m / np) % np, i = (m / Nfp) % nex, k = m / (Nfp * nex);
          ms = (np - 1 - p1) + np * (p3 + np * ((nex - 1 - i) + nex * k));
        }
        const int src = mT.vmapB[mT.halo_off[f] + ms];
        const size_t dst = nintU + mU.halo_off[g] + m;
        for (size_t k = 0; k < scal[T].size(); ++k) scal[U][k][dst] = scal[T][k][src];
        if (u1[T]) {
          double vlon, vlat, va, vb;
          cs2lonlat_vec(mT.panelID, mT.pos[0][src], mT.pos[1][src], mT.RPlanet, u1[T][src], u2[T][src], vlon, vlat);
          const int own = mU.vmapB[mU.halo_off[g] + m];
          lonlat2cs_vec(mU.panelID, mU.pos[0][own], mU.pos[1][own], mU.RPlanet, vlon, vlat, va, vb);
          u1[U][dst] = va; u2[U][dst] = vb;
        }
      }
    }
  }
}

// driver_nonhydro3d.F90:614-963 with LOCAL_MESH_NUM = 6
void sphere_update(Driver* d[6]) {
  const Element& e = d[0]->elem;
  Mesh* mesh[6];
  for (int p = 0; p < 6; ++p) mesh[p] = &d[p]->mesh;
  const int rkvar[5] = {DENS_VID, THERM_VID, MOMZ_VID, MOMX_VID, MOMY_VID};
  auto exchange_prog = [&](bool with_dpres) {
    std::vector<double*> sc[6]; double* u1[6]; double* u2[6];
    for (int p = 0; p < 6; ++p) {
      DynState& s = d[p]->st;
      for (int v = 0; v < 5; ++v) d[p]->mesh.exchange_halo(e, s.prog(v));       // bottom / top faces: own values
      sc[p] = {s.DDENS.data(), s.DRHOT.data(), s.MOMZ.data()};
      if (with_dpres) { d[p]->mesh.exchange_halo(e, s.DPRES.data()); sc[p].push_back(s.DPRES.data()); }
      u1[p] = s.MOMX.data(); u2[p] = s.MOMY.data();
    }
    sphere_exchange(e, mesh, sc, u1, u2);
  };
  const int ns = d[0]->tint.sc.nstage;
  const bool hevi = d[0]->hevi;
  if (hevi)
    for (int p = 0; p < 6; ++p) {
      const size_t nint = size_t(e.Np) * d[p]->mesh.Ne;
      for (int v : rkvar) d[p]->tint.store_var0(d[p]->st.prog(v), v, 0, nint);
    }
  for (int stage = 0; stage < ns; ++stage) {
    for (int p = 0; p < 6; ++p) {
      if (!hevi) { drhot2pres(d[p]->elem, d[p]->mesh, d[p]->cst, d[p]->st); continue; }
      Driver& D = *d[p];
      const size_t nint = size_t(e.Np) * D.mesh.Ne;
      const int ind = D.tint.sc.indmap[stage];
      double* out[5]; const double* v0[5];
      for (int v = 0; v < 5; ++v) { out[v] = D.tint.tend_im_buf(v, ind); v0[v] = &D.tint.var0[size_t(v) * D.tint.n]; }
      hevi_cal_vi(D.elem, D.mesh, D.cst, D.st, v0, D.tint.implicit_diagfac(stage), D.tint.dt, out);
      for (int v : rkvar) D.tint.store_implicit(stage, D.st.prog(v), v, 0, nint);
      drhot2pres(D.elem, D.mesh, D.cst, D.st);
    }
    exchange_prog(true);
    for (int p = 0; p < 6; ++p) {
      Driver& D = *d[p];
      const size_t nint = size_t(e.Np) * D.mesh.Ne;
      const int ind = D.tint.sc.indmap[stage];
      apply_bc_progvars(D.elem, D.mesh, D.bnd, D.st);
      double* out[5];
      for (int v = 0; v < 5; ++v) out[v] = D.tint.tend_ex_buf(v, ind);
      global_cal_tend(D.elem, D.mesh, D.cst, D.st, hevi, out);
      if (D.sponge.on) sponge_add_tend(D.elem, D.mesh, D.sponge, D.st, out);               // driver_nonhydro3d.F90:830-841
      if (D.phytend) add_phy_tend(D.elem, D.mesh, D.cst, D.st, D.entot_conserve, out);
      for (int v : rkvar) D.tint.advance(stage, D.st.prog(v), v, 0, nint);
    }
  }
  for (int p = 0; p < 6; ++p) {
    if (d[p]->modalfilter) modalfilter_apply(d[p]->elem, d[p]->mesh, d[p]->st);
    drhot2pres(d[p]->elem, d[p]->mesh, d[p]->cst, d[p]->st);
  }
}

}  // namespace feo
