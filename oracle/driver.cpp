// TEST INFRASTRUCTURE ONLY (see fe_oracle.hpp).  Stage sequencing of one dynamics step + monitors.
#include "fe_oracle.hpp"

namespace feo {

// fluid_dyn_solver/scale_atm_dyn_dgm_driver_nonhydro3d.F90:614-963 (hide_mpi_comm_flag = .false.,
// QA = 0, no sponge layer, zero physics tendencies)
void Driver::update() {
  const size_t nint = size_t(elem.Np) * mesh.Ne;
  const int rkvar[5] = {DENS_VID, THERM_VID, MOMZ_VID, MOMX_VID, MOMY_VID};
  if (hevi)
    for (int v : rkvar) tint.store_var0(st.prog(v), v, 0, nint);                       // :703-713
  if (tracer) DENS0_TRC = st.DDENS;                                                    // = tint%var0_2D(:,:,DENS_VID), :926-937
  for (int stage = 0; stage < tint.sc.nstage; ++stage) {
    const int ind = tint.sc.indmap[stage];
    if (hevi) {                                                                        // :730-766
      double impl_fac = tint.implicit_diagfac(stage);
      double* out[5]; const double* v0[5];
      for (int v = 0; v < 5; ++v) { out[v] = tint.tend_im_buf(v, ind); v0[v] = &tint.var0[size_t(v) * tint.n]; }
      hevi_cal_vi(elem, mesh, cst, st, v0, impl_fac, tint.dt, out);
      for (int v : rkvar) tint.store_implicit(stage, st.prog(v), v, 0, nint);
    }
    for (int v = 0; v < 5; ++v) mesh.exchange_halo(elem, st.prog(v));                  // :770
    drhot2pres(elem, mesh, cst, st);                                                   // :777
    mesh.exchange_halo(elem, st.DPRES.data());                                         // :787
    apply_bc_progvars(elem, mesh, bnd, st);                                            // :797
    double* out[5];
    for (int v = 0; v < 5; ++v) out[v] = tint.tend_ex_buf(v, ind);
    if (global) global_cal_tend(elem, mesh, cst, st, hevi, out);
    else if (hevi) hevi_cal_tend(elem, mesh, cst, st, out);
    else heve_cal_tend(elem, mesh, cst, st, out);                                       // :815
    if (sponge.on) sponge_add_tend(elem, mesh, sponge, st, out);                        // :830-841
    if (phytend) add_phy_tend(elem, mesh, cst, st, entot_conserve, out);                // :843-857
    if (tracer) trc_save_massflux(*this, stage);                                       // :900-917
    for (int v : rkvar) tint.advance(stage, st.prog(v), v, 0, nint);                   // :920
  }
  if (tracer) DENS_TRC = st.DDENS;                                                     // before the modal filter, :926-937
  if (modalfilter) modalfilter_apply(elem, mesh, st);                                  // :940-951
  drhot2pres(elem, mesh, cst, st);                                                     // :954
  // numerical diffusion follows the dynamics step (model mod_atmos_dyn.F90:343-349)
  if (numdiff) { numdiff_apply(elem, mesh, nd, st); drhot2pres(elem, mesh, cst, st); }
}

// file/scale_file_monitor_meshfield.F90:176-213 (cal_total_lc) over the fields of
// model/atm_nonhydro3d/src/atmos/mod_atmos_vars_container.F90:1281-1357 (dry: QDRY = 1, G_ij = I)
void monitor_sums(const Driver& d, double out[5]) {
  const Element& e = d.elem; const Mesh& m = d.mesh; const DynState& s = d.st;
  for (int i = 0; i < 5; ++i) out[i] = 0.0;
  for (int ke = 0; ke < m.Ne; ++ke)
    for (int p = 0; p < e.Np; ++p) {
      size_t i = size_t(p) + size_t(ke) * e.Np;
      double w = e.IntWeight[p] * m.J[i] * m.Gsqrt[i];
      double dens = s.DDENS[i] + s.DENS_hyd[i];
      double engk = 0.5 * (s.MOMX[i] * s.MOMX[i] + s.MOMY[i] * s.MOMY[i] + s.MOMZ[i] * s.MOMZ[i]) / dens;
      double engi = s.PRES[i] / s.Rtot[i] * d.cst.CVdry;
      double engp = dens * d.cst.GRAV * m.pos[2][i];
      out[0] += w * s.DDENS[i];
      out[1] += w * (engk + engi + engp);
      out[2] += w * engk;
      out[3] += w * engi;
      out[4] += w * engp;
    }
}

}  // namespace feo
