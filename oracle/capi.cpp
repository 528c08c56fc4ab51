// TEST INFRASTRUCTURE ONLY (see fe_oracle.hpp).  C entry points used by tests/ and bench.py through ctypes.
#include <cstring>
#include <stdexcept>
#include <cstdio>

#include <omp.h>

#include "fe_oracle.hpp"

using namespace feo;

namespace {
struct Handle {
  Driver d;
  std::string err;
};
thread_local std::string g_err;
template <class F> int guard(F&& f) {
  try { f(); return 0; } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
}  // namespace

extern "C" {

const char* feo_last_error() { return g_err.c_str(); }
// OpenMP team size of the restatement's element loops (torchrun exports OMP_NUM_THREADS=1 before libgomp is initialised, so
// the timing legs of bench.py set the thread count explicitly); returns the value in effect
int feo_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); return omp_get_max_threads(); }

void* feo_create(int p, int lumped, int NeX, int NeY, int NeZ, const double* dom, const double* FZ, const int* periodic) {
  Handle* h = nullptr;
  int rc = guard([&] {
    h = new Handle;
    h->d.elem.init(p, lumped != 0);
    bool per[3] = {periodic[0] != 0, periodic[1] != 0, periodic[2] != 0};
    h->d.mesh.init_cube(h->d.elem, NeX, NeY, NeZ, dom[0], dom[1], dom[2], dom[3], dom[4], dom[5], FZ, per);
    h->d.st.alloc(size_t(h->d.elem.Np) * h->d.mesh.NeA, size_t(h->d.elem.Nfp) * h->d.mesh.Ne2D);
  });
  if (rc) { delete h; return nullptr; }
  return h;
}
// one cubed-sphere panel tile: (alpha, beta) in [-pi/4, pi/4]^2, z in [0, ztop] (or FZ), shallow-atmosphere metric
void* feo_create_panel(int p, int lumped, int panelID, int NeX, int NeY, int NeZ, double ztop, const double* FZ, double radius, int shallow) {
  Handle* h = nullptr;
  int rc = guard([&] {
    h = new Handle;
    h->d.elem.init(p, lumped != 0);
    h->d.mesh.init_cubedsphere_panel(h->d.elem, panelID, NeX, NeY, NeZ, FZ, ztop, radius, shallow != 0);
    h->d.st.alloc(size_t(h->d.elem.Np) * h->d.mesh.NeA, size_t(h->d.elem.Nfp) * h->d.mesh.Ne2D);
  });
  if (rc) { delete h; return nullptr; }
  return h;
}
void feo_destroy(void* hv) { delete static_cast<Handle*>(hv); }
// element / mesh of a handle, for the sample restatements that live on the same mesh (advect3d.cpp)
const void* feo_elem_ptr(void* hv) { return &static_cast<Handle*>(hv)->d.elem; }
const void* feo_mesh_ptr(void* hv) { return &static_cast<Handle*>(hv)->d.mesh; }

void feo_dims(void* hv, int* out) {
  auto& d = static_cast<Handle*>(hv)->d;
  out[0] = d.elem.Np; out[1] = d.elem.NfpTot; out[2] = d.mesh.Ne; out[3] = d.mesh.NeA; out[4] = d.mesh.Nhalo;
  out[5] = d.elem.np; out[6] = d.mesh.Ne2D;
}

// name -> pointer/size of an internal array (double)
double* feo_array(void* hv, const char* name, long* n) {
  auto& d = static_cast<Handle*>(hv)->d;
  std::string s(name);
  vec* v = nullptr;
  if (s == "x1d") v = &d.elem.x1d; else if (s == "w1d") v = &d.elem.w1d; else if (s == "D1D") v = &d.elem.D1D;
  else if (s == "lift1d") v = &d.elem.lift1d; else if (s == "VPOrdM1") v = &d.elem.VPOrdM1;
  else if (s == "V1") v = &d.elem.V1; else if (s == "invV1") v = &d.elem.invV1;
  else if (s == "filt_h") v = &d.elem.filt_h; else if (s == "filt_v") v = &d.elem.filt_v;
  else if (s == "IntWeight") v = &d.elem.IntWeight;
  else if (s == "pos_x") v = &d.mesh.pos[0]; else if (s == "pos_y") v = &d.mesh.pos[1]; else if (s == "pos_z") v = &d.mesh.pos[2];
  else if (s == "E11") v = &d.mesh.E11; else if (s == "E22") v = &d.mesh.E22; else if (s == "E33") v = &d.mesh.E33;
  else if (s == "J") v = &d.mesh.J; else if (s == "nx") v = &d.mesh.nx; else if (s == "ny") v = &d.mesh.ny;
  else if (s == "nz") v = &d.mesh.nz; else if (s == "Fscale") v = &d.mesh.Fscale;
  else if (s == "Gsqrt") v = &d.mesh.Gsqrt; else if (s == "G13") v = &d.mesh.G13; else if (s == "G23") v = &d.mesh.G23;
  else if (s == "GsqrtH") v = &d.mesh.GsqrtH;
  else if (s == "gam") v = &d.mesh.gam; else if (s == "alpha2D") v = &d.mesh.alpha2D; else if (s == "beta2D") v = &d.mesh.beta2D;
  else if (s == "GIJ11") v = &d.mesh.GIJ11; else if (s == "GIJ12") v = &d.mesh.GIJ12; else if (s == "GIJ22") v = &d.mesh.GIJ22;
  else if (s == "Gij11") v = &d.mesh.Gij11; else if (s == "Gij12") v = &d.mesh.Gij12; else if (s == "Gij22") v = &d.mesh.Gij22;
  else if (s == "DDENS") v = &d.st.DDENS; else if (s == "MOMX") v = &d.st.MOMX; else if (s == "MOMY") v = &d.st.MOMY;
  else if (s == "MOMZ") v = &d.st.MOMZ; else if (s == "DRHOT") v = &d.st.DRHOT;
  else if (s == "DENS_hyd") v = &d.st.DENS_hyd; else if (s == "PRES_hyd") v = &d.st.PRES_hyd;
  else if (s == "THERM_hyd") v = &d.st.THERM_hyd; else if (s == "PRES_hyd_ref") v = &d.st.PRES_hyd_ref;
  else if (s == "Rtot") v = &d.st.Rtot; else if (s == "CVtot") v = &d.st.CVtot; else if (s == "CPtot") v = &d.st.CPtot;
  else if (s == "PRES") v = &d.st.PRES; else if (s == "DPRES") v = &d.st.DPRES;
  else if (s == "DPhydDx") v = &d.st.DPhydDx; else if (s == "DPhydDy") v = &d.st.DPhydDy;
  else if (s == "CORIOLIS") v = &d.st.CORIOLIS;
  else if (s == "DENS_tp") v = &d.st.DENS_tp; else if (s == "MOMX_tp") v = &d.st.MOMX_tp; else if (s == "MOMY_tp") v = &d.st.MOMY_tp;
  else if (s == "MOMZ_tp") v = &d.st.MOMZ_tp; else if (s == "RHOT_tp") v = &d.st.RHOT_tp; else if (s == "RHOH_p") v = &d.st.RHOH_p;
  else if (s == "tend_ex") v = &d.tint.tend_ex; else if (s == "tend_im") v = &d.tint.tend_im;
  if (!v) { *n = 0; return nullptr; }
  *n = long(v->size());
  return v->data();
}
int* feo_iarray(void* hv, const char* name, long* n) {
  auto& d = static_cast<Handle*>(hv)->d;
  std::string s(name);
  ivec* v = nullptr;
  if (s == "vmapM") v = &d.mesh.vmapM; else if (s == "vmapP") v = &d.mesh.vmapP; else if (s == "vmapB") v = &d.mesh.vmapB;
  else if (s == "emap2d") v = &d.mesh.emap2d; else if (s == "Fmask") v = &d.elem.Fmask;
  if (!v) { *n = 0; return nullptr; }
  *n = long(v->size());
  return v->data();
}

void feo_set_consts(void* hv, const double* c) {
  auto& k = static_cast<Handle*>(hv)->d.cst;
  k.GRAV = c[0]; k.Rdry = c[1]; k.CPdry = c[2]; k.CVdry = c[3]; k.PRES00 = c[4]; k.OHM = c[5];
}

// eqs: "NONHYDRO3D_HEVE" | "NONHYDRO3D_HEVI"; mf = {etac_h, alpha_h, ord_h, etac_v, alpha_v, ord_v}
int feo_setup_dyn(void* hv, const char* eqs, const char* tinteg, double dt, int modalfilter, const double* mf, const int* vel_bc) {
  auto* h = static_cast<Handle*>(hv);
  return guard([&] {
    auto& d = h->d;
    std::string e(eqs);
    d.global = false;
    if (e == "NONHYDRO3D_HEVE") d.hevi = false; else if (e == "NONHYDRO3D_HEVI") d.hevi = true;
    else if (e == "GLOBALNONHYDRO3D_HEVI") {
      if (!d.mesh.is_global) throw std::runtime_error("GLOBALNONHYDRO3D_HEVI needs a cubed-sphere panel mesh");
      d.hevi = true; d.global = true;
    }
    else if (e == "GLOBALNONHYDRO3D_HEVE") {
      if (!d.mesh.is_global) throw std::runtime_error("GLOBALNONHYDRO3D_HEVE needs a cubed-sphere panel mesh");
      d.hevi = false; d.global = true;
    }
    else throw std::runtime_error("unsupported EQS_TYPE " + e);
    if (d.mesh.is_global && !d.global) throw std::runtime_error("a cubed-sphere panel mesh needs a GLOBALNONHYDRO3D equation set");
    d.tint.init(tinteg, dt, 5, size_t(d.elem.Np) * d.mesh.NeA);
    if (d.hevi != d.tint.sc.imex) throw std::runtime_error("HEVI needs an IMEX scheme and HEVE an explicit one");
    d.modalfilter = modalfilter != 0;
    if (d.modalfilter) d.elem.setup_filter(mf[0], mf[1], int(mf[2]), mf[3], mf[4], int(mf[5]));
    for (int f = 0; f < 6; ++f) d.bnd.vel_bc[f] = vel_bc[f];
  });
}

// Tracer advection with a prescribed mass flux (ONLY_TRACERADV_FLAG, driver_trcadv3d.F90:312-559): advances QTRC (Np*NeA, in/out)
// by nsteps of `tinteg` with the momentum / density of the handle's state; mf[6] = tracer modal filter (etac_h, alpha_h, ord_h,
// etac_v, alpha_v, ord_v) when modalfilter != 0; RHOQ_tp may be NULL (zero).
void feo_set_tracer_coupling(void* hv, int on) { static_cast<Handle*>(hv)->d.tracer = on != 0; }
// one tracer step with the mass fluxes the last dynamics step accumulated (feo_set_tracer_coupling(1) before feo_update)
int feo_trcadv_update_coupled(void* hv, const char* tinteg, double dt, int modalfilter, const double* mf, int disable_limiter,
                              double* QTRC, const double* RHOQ_tp) {
  auto* h = static_cast<Handle*>(hv);
  return guard([&] {
    auto& d = h->d;
    RKScheme sc;
    if (!sc.init(tinteg)) throw std::runtime_error(std::string("unsupported RK scheme ") + tinteg);
    Element ef = d.elem;
    if (modalfilter) ef.setup_filter(mf[0], mf[1], int(mf[2]), mf[3], mf[4], int(mf[5]));
    vec zero;
    if (!RHOQ_tp) { zero.assign(size_t(d.elem.Np) * d.mesh.NeA, 0.0); RHOQ_tp = zero.data(); }
    trcadv_update_coupled(d, ef, sc, dt, modalfilter != 0, disable_limiter != 0, QTRC, RHOQ_tp);
  });
}
int feo_trcadv_update(void* hv, const char* tinteg, double dt, int nsteps, int modalfilter, const double* mf, int disable_limiter,
                      double* QTRC, const double* RHOQ_tp) {
  auto* h = static_cast<Handle*>(hv);
  return guard([&] {
    auto& d = h->d;
    RKScheme sc;
    if (!sc.init(tinteg)) throw std::runtime_error(std::string("unsupported RK scheme ") + tinteg);
    Element ef = d.elem;
    if (modalfilter) ef.setup_filter(mf[0], mf[1], int(mf[2]), mf[3], mf[4], int(mf[5]));
    vec zero;
    if (!RHOQ_tp) { zero.assign(size_t(d.elem.Np) * d.mesh.NeA, 0.0); RHOQ_tp = zero.data(); }
    for (int n = 0; n < nsteps; ++n) trcadv_update_advtest(d, ef, sc, dt, modalfilter != 0, disable_limiter != 0, QTRC, RHOQ_tp);
  });
}

// restart-time preparation: model/atm_nonhydro3d mod_atmos_vars.F90:553-636 (THERM_hyd, pressure, hyd halos, DPhydDx/y)
int feo_prepare(void* hv) {
  auto* h = static_cast<Handle*>(hv);
  return guard([&] {
    auto& d = h->d;
    calc_rhot_hyd(d.elem, d.mesh, d.cst, d.st);
    drhot2pres(d.elem, d.mesh, d.cst, d.st);
    for (vec* v : {&d.st.PRES_hyd, &d.st.DENS_hyd, &d.st.THERM_hyd, &d.st.PRES_hyd_ref, &d.st.Rtot, &d.st.CVtot, &d.st.CPtot,
                   &d.st.PRES, &d.st.DPRES})
      d.mesh.exchange_halo(d.elem, v->data());
    calc_phyd_hgrad(d.elem, d.mesh, d.st);
  });
}

// numerical diffusion after every step: PARAM_ATMOS_DYN_NUMDIFF (ND_LAPLACIAN_NUM, ND_COEF_h, ND_COEF_v), therm_bc: 1 = ADIABAT
void feo_set_numdiff(void* hv, int on, int laplacian_num, double coef_h, double coef_v, const int* therm_bc) {
  auto& d = static_cast<Handle*>(hv)->d;
  d.numdiff = on != 0;
  d.nd.laplacian_num = laplacian_num; d.nd.coef_h = coef_h; d.nd.coef_v = coef_v; d.nd.dt = d.tint.dt;
  for (int f = 0; f < 6; ++f) { d.nd.vel_bc[f] = d.bnd.vel_bc[f]; d.nd.therm_bc[f] = therm_bc ? therm_bc[f] : 0; }
}
// one application on the current state (no dynamics step)
int feo_numdiff_apply(void* hv) {
  auto* h = static_cast<Handle*>(hv);
  return guard([&] { numdiff_apply(h->d.elem, h->d.mesh, h->d.nd, h->d.st); });
}

// PARAM_ATMOS_DYN_SPONGELAYER (spongelayer.F90:55-118); call after setup_dyn
void feo_set_sponge(void* hv, int on, double tau, double height, int layer, int hveldamp) {
  auto& d = static_cast<Handle*>(hv)->d;
  d.sponge.on = on != 0; d.sponge.tau = tau; d.sponge.height = height; d.sponge.layer = layer; d.sponge.hveldamp = hveldamp != 0;
  if (layer > 0) d.sponge.height = d.mesh.pos[2][size_t((layer - 1) * d.mesh.NeX * d.mesh.NeY) * d.elem.Np];
  if (d.sponge.tau < 0.0) d.sponge.tau = d.tint.dt * 10.0;
}

// physics tendencies on / off (the arrays are DENS_tp ... RHOH_p of feo_array)
void feo_set_phytend(void* hv, int on) { static_cast<Handle*>(hv)->d.phytend = on != 0; }

// ---- whole cubed sphere: six panel handles (created with feo_create_panel, panel ids 1..6 in order)
int feo_sphere_exchange(void** hv6, int with_dpres) {
  return guard([&] {
    Mesh* mesh[6]; std::vector<double*> sc[6]; double* u1[6]; double* u2[6];
    for (int p = 0; p < 6; ++p) {
      auto& d = static_cast<Handle*>(hv6[p])->d;
      if (d.mesh.panelID != p + 1) throw std::runtime_error("panels must be passed in the order 1..6");
      mesh[p] = &d.mesh;
      for (int v = 0; v < 5; ++v) d.mesh.exchange_halo(d.elem, d.st.prog(v));
      sc[p] = {d.st.DDENS.data(), d.st.DRHOT.data(), d.st.MOMZ.data()};
      if (with_dpres) { d.mesh.exchange_halo(d.elem, d.st.DPRES.data()); sc[p].push_back(d.st.DPRES.data()); }
      u1[p] = d.st.MOMX.data(); u2[p] = d.st.MOMY.data();
    }
    sphere_exchange(static_cast<Handle*>(hv6[0])->d.elem, mesh, sc, u1, u2);
  });
}
// exchange of the background fields after set-up (scalars only)
int feo_sphere_exchange_aux(void** hv6) {
  return guard([&] {
    Mesh* mesh[6]; std::vector<double*> sc[6]; double* u1[6]; double* u2[6];
    for (int p = 0; p < 6; ++p) {
      auto& d = static_cast<Handle*>(hv6[p])->d;
      mesh[p] = &d.mesh;
      sc[p] = {d.st.DENS_hyd.data(), d.st.PRES_hyd.data(), d.st.THERM_hyd.data(), d.st.PRES_hyd_ref.data(), d.st.Rtot.data(),
               d.st.CVtot.data(), d.st.CPtot.data(), d.st.PRES.data(), d.st.DPRES.data()};
      u1[p] = u2[p] = nullptr;
    }
    sphere_exchange(static_cast<Handle*>(hv6[0])->d.elem, mesh, sc, u1, u2);
    // update_phyd_hgrad follows the exchange of the background fields (model mod_atmos_vars.F90:553-636, driver_nonhydro3d.F90:1060-1095):
    // at a panel edge the exterior value of PRES_hyd is the neighbour panel's, not the own face value feo_prepare had
    for (int p = 0; p < 6; ++p) {
      auto& d = static_cast<Handle*>(hv6[p])->d;
      calc_phyd_hgrad(d.elem, d.mesh, d.st);
    }
  });
}
int feo_sphere_update(void** hv6, int nsteps) {
  return guard([&] {
    Driver* d[6];
    for (int p = 0; p < 6; ++p) { d[p] = &static_cast<Handle*>(hv6[p])->d; if (!d[p]->global) throw std::runtime_error("sphere needs GLOBALNONHYDRO3D_HEVI"); }
    for (int n = 0; n < nsteps; ++n) sphere_update(d);
  });
}

int feo_update(void* hv, int nsteps) {
  auto* h = static_cast<Handle*>(hv);
  return guard([&] { for (int n = 0; n < nsteps; ++n) h->d.update(); });
}

// vertical-implicit tendency (cal_vi seam) of the current state about var0; arrays (5, Np*NeA) in the oracle's
// variable order DENS, RHOT, MOMZ, MOMX, MOMY
int feo_cal_vi(void* hv, double impl_fac, double dt, const double* var0, double* out) {
  auto* h = static_cast<Handle*>(hv);
  return guard([&] {
    auto& d = h->d;
    const size_t n = size_t(d.elem.Np) * d.mesh.NeA;
    const double* v0[5]; double* o[5];
    for (int v = 0; v < 5; ++v) { v0[v] = var0 + size_t(v) * n; o[v] = out + size_t(v) * n; }
    hevi_cal_vi(d.elem, d.mesh, d.cst, d.st, v0, impl_fac, dt, o);
  });
}

void feo_monitor(void* hv, double* out) { monitor_sums(static_cast<Handle*>(hv)->d, out); }

// single pieces of the step, for kernel-by-kernel parity tests
int feo_stage_piece(void* hv, const char* what) {
  auto* h = static_cast<Handle*>(hv);
  return guard([&] {
    auto& d = h->d; std::string w(what);
    if (w == "exchange") { for (int v = 0; v < 5; ++v) d.mesh.exchange_halo(d.elem, d.st.prog(v)); }
    else if (w == "pressure") { drhot2pres(d.elem, d.mesh, d.cst, d.st); d.mesh.exchange_halo(d.elem, d.st.DPRES.data()); }
    else if (w == "bc") apply_bc_progvars(d.elem, d.mesh, d.bnd, d.st);
    else if (w == "tend_ex") {
      double* out[5]; for (int v = 0; v < 5; ++v) out[v] = d.tint.tend_ex_buf(v, 0);
      if (d.global) global_cal_tend(d.elem, d.mesh, d.cst, d.st, d.hevi, out);
      else if (d.hevi) hevi_cal_tend(d.elem, d.mesh, d.cst, d.st, out); else heve_cal_tend(d.elem, d.mesh, d.cst, d.st, out);
      if (d.sponge.on) sponge_add_tend(d.elem, d.mesh, d.sponge, d.st, out);
      if (d.phytend) add_phy_tend(d.elem, d.mesh, d.cst, d.st, d.entot_conserve, out);
    }
    else if (w == "modalfilter") modalfilter_apply(d.elem, d.mesh, d.st);
    else throw std::runtime_error("unknown piece " + w);
  });
}

// per-element operators on caller arrays: name in {Dx,Dy,Dz,Lift,Div,VFilterPM1,ModalFilter}
int feo_elem_op(void* hv, const char* name, const double* in, const double* in2, double* out) {
  auto* h = static_cast<Handle*>(hv);
  return guard([&] {
    const Element& e = h->d.elem; std::string s(name);
    if (s == "Dx") op_dx(e, in, out); else if (s == "Dy") op_dy(e, in, out); else if (s == "Dz") op_dz(e, in, out);
    else if (s == "Lift") op_lift(e, in, out); else if (s == "Div") op_div(e, in, in2, out);
    else if (s == "VFilterPM1") op_matz(e, e.VPOrdM1.data(), in, out);
    else if (s == "ModalFilter") { vec w(e.Np); op_modal_filter(e, in, w.data(), out); }
    else throw std::runtime_error("unknown op " + s);
  });
}
void feo_lift_dense(void* hv, double* out) {
  vec L = static_cast<Handle*>(hv)->d.elem.lift_dense();
  std::memcpy(out, L.data(), L.size() * sizeof(double));
}
void feo_dmat_dense(void* hv, int dir, double* out) {
  vec D = static_cast<Handle*>(hv)->d.elem.dmat_dense(dir);
  std::memcpy(out, D.data(), D.size() * sizeof(double));
}

// sparse matrix (a1): build from a dense row-major matrix, return c = A b and A(i,j) lookups
int feo_sparsemat_matmul(const double* A, int M, int N, double eps, int ell, const double* b, double* c, double* getval_MN) {
  return guard([&] {
    SparseMat sm; sm.init(A, M, N, eps, ell != 0);
    sm.matmul(b, c);
    if (getval_MN) for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) getval_MN[size_t(i) * N + j] = sm.get(i, j);
  });
}

// c = A (b1 .* b2) (mode 1) or c(NQ,M) = A b(NQ,N) (mode 2), CSR or ELL storage; also hands out the CSR arrays (1-based like the
// reference's colIdx / rowPtr) when csr_val != nullptr (sized nnz, nnz, M + 1 by the caller; nnz returned through *nnz_out)
int feo_sparsemat_matmul_ex(const double* A, int M, int N, double eps, int ell, int mode, int NQ, const double* b1, const double* b2,
                            double* c, int* nnz_out, double* csr_val, int* csr_col, int* csr_rowptr) {
  return guard([&] {
    SparseMat sm; sm.init(A, M, N, eps, ell != 0);
    if (mode == 1) sm.matmul_1_2(b1, b2, c);
    else if (mode == 2) sm.matmul2(b1, c, NQ);
    else sm.matmul(b1, c);
    if (nnz_out) *nnz_out = sm.nnz;
    if (csr_val && !sm.ell) {
      for (int k = 0; k < sm.nnz; ++k) { csr_val[k] = sm.val[k]; csr_col[k] = sm.colIdx[k] + 1; }
      for (int i = 0; i <= M; ++i) csr_rowptr[i] = sm.rowPtr[i] + 1;
    }
  });
}

// RK scheme tables: out arrays sized by the caller from nstage
int feo_rk_info(const char* scheme, int* nstage, int* tend_buf_size, int* low_storage, int* imex) {
  RKScheme sc; if (!sc.init(scheme)) return 1;
  *nstage = sc.nstage; *tend_buf_size = sc.tend_buf_size; *low_storage = sc.low_storage; *imex = sc.imex; return 0;
}
int feo_rk_coef(const char* scheme, double* a_ex, double* b_ex, double* a_im, double* b_im, double* sig, double* gam) {
  RKScheme sc; if (!sc.init(scheme)) return 1;
  std::memcpy(a_ex, sc.a_ex.data(), sc.a_ex.size() * 8); std::memcpy(b_ex, sc.b_ex.data(), sc.b_ex.size() * 8);
  std::memcpy(a_im, sc.a_im.data(), sc.a_im.size() * 8); std::memcpy(b_im, sc.b_im.data(), sc.b_im.size() * 8);
  std::memcpy(sig, sc.sig.data(), sc.sig.size() * 8); std::memcpy(gam, sc.gam.data(), sc.gam.size() * 8);
  return 0;
}
// stand-alone integrator object for the harmonic-oscillator tests (FElib/test/common/timeint_rk)
void* feo_rk_create(const char* scheme, double dt, int nvar, long n) {
  auto* t = new TimeIntRK;
  if (guard([&] { t->init(scheme, dt, nvar, size_t(n)); })) { delete t; return nullptr; }
  return t;
}
void feo_rk_destroy(void* t) { delete static_cast<TimeIntRK*>(t); }
double* feo_rk_tend(void* tv, int im, int var, int stage) {
  auto* t = static_cast<TimeIntRK*>(tv);
  return im ? t->tend_im_buf(var, t->sc.indmap[stage]) : t->tend_ex_buf(var, t->sc.indmap[stage]);
}
double feo_rk_implicit_fac(void* tv, int stage) { return static_cast<TimeIntRK*>(tv)->implicit_diagfac(stage); }
void feo_rk_store_implicit(void* tv, int stage, double* q, int var) { auto* t = static_cast<TimeIntRK*>(tv); t->store_implicit(stage, q, var, 0, t->n); }
void feo_rk_advance(void* tv, int stage, double* q, int var) { auto* t = static_cast<TimeIntRK*>(tv); t->advance(stage, q, var, 0, t->n); }

}  // extern "C"
