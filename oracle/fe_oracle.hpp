// TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of FE-Project's DG dynamics hot path.
//
// Nothing under fe_project_b200/ (the product) may include, link or call this code.  It is used
// by tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py.
//
// PARITY STATUS: the reference (Fortran 2008 + SCALE 5.5.5 + MPI + LAPACK + NetCDF) cannot be
// compiled in this environment (no Fortran compiler, SCALE library not vendored), so this oracle
// is pinned by (i) the reference's own known-answer unit tests restated in tests/ (element
// operators on 4x^p+3y^p+2z^p, sparse-matrix 5x5 case, Butcher tableaux of SSP3s3o/SSP4s3o, RK
// convergence orders, halo fill pattern, VMapM / VMapP and halo sources in closed form, LGL points) and
// (ii) independent NumPy restatements: of the set-up code (fe_project_b200/element.py, mesh.py,
// cubedsphere.py) and, since round 2, of the dynamics rows themselves (oracle/numpy_dyn.py: pressure,
// flux + tendency of HEVE / HEVI / global HEVI, boundary condition, modal filter, the vertical-implicit
// Newton step as a dense column solve, the whole step; tests/test_oracle_numpy_dyn.py: agreement to
// round-off, flat and terrain-following).  Numerical flux, tendency assembly, pressure, boundary
// conditions, the vertical-implicit solver and the full step have NO golden vectors in the reference:
// for those rows parity is UNPINNED BY REFERENCE VECTORS -- two restatements written from the same
// Fortran agree, nothing the reference itself produced was compared ("Oracle").
//
// Every function cites the reference file:line it restates.  Paths are relative to
// /root/reference/FElib/src unless stated otherwise.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace feo {

using vec = std::vector<double>;
using ivec = std::vector<int>;

// Physical constants of SCALE's scale_const (external library, not vendored; values from its
// public source, passed in by the caller, never hard-coded in kernels).
struct Consts {
  double GRAV = 9.80665, Rdry = 287.04, CPdry = 1004.64, CVdry = 1004.64 - 287.04,
         PRES00 = 1.0e5, OHM = 7.2920e-5, RPlanet = 6.37122e6, EPS = 2.220446e-16;
};

// prognostic variable ids, fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_common.F90:68-77 (0-based here)
enum { DENS_VID = 0, THERM_VID = 1, RHOT_VID = 1, MOMZ_VID = 2, MOMX_VID = 3, MOMY_VID = 4, PRGVAR_NUM = 5 };

// ---------------------------------------------------------------- element
// element/scale_element_line.F90, scale_element_hexahedral.F90:52-404,
// scale_element_operation_tensorprod3D.F90.erb:508-565
struct Element {
  int p = 0, np = 0, Np = 0, Nfp = 0, NfpTot = 0;
  bool lumped = false;
  vec x1d, w1d;         // LGL nodes / weights
  vec V1, invV1;        // 1D orthonormal-Legendre Vandermonde, row-major (node, mode)
  vec D1D;              // D1D[i*np+l] = d l_l/dx (x_i)
  vec invM1, M1;        // 1D mass matrices
  vec lift1d;           // lift1d[m*2+side]
  vec VPOrdM1;          // removes highest mode, row-major
  vec filt_h, filt_v;   // 1D modal filter matrices, row-major (identity until setup_filter)
  vec IntWeight;        // Np
  ivec Fmask;           // Fmask[f*Nfp + fp], 0-based node id
  void init(int order, bool lumped_mass);
  void setup_filter(double etac_h, double alpha_h, int ord_h, double etac_v, double alpha_v, int ord_v);
  // dense (Np x NfpTot) lift, row-major, built the way hexahedral.F90:331-400 builds it (invM * Emat)
  vec lift_dense() const;
  // dense (Np x Np) Dx1/Dx2/Dx3, row-major (hexahedral.F90:251-281)
  vec dmat_dense(int dir) const;
};

// per-element operators, element/scale_element_operation_tensorprod3D_kernel.F90.erb
void op_dx(const Element& e, const double* in, double* out);
void op_dy(const Element& e, const double* in, double* out);
void op_dz(const Element& e, const double* in, double* out);
void op_matz(const Element& e, const double* Mat_rowmajor, const double* in, double* out);
void op_lift(const Element& e, const double* in_face, double* out);
void op_div(const Element& e, const double* flux3, const double* del_flux, double* dflux4);  // Div: (Np,3),(NfpTot)->(Np,4)
void op_modal_filter(const Element& e, const double* in, double* work, double* out);

// ---------------------------------------------------------------- sparse matrix (a1)
// common/scale_sparsemat.F90:33-55, 100-250, 439-474, 554-634
struct SparseMat {
  int M = 0, N = 0, nnz = 0, col_size = 0;
  bool ell = false;
  vec val;
  ivec colIdx, rowPtr;
  void init(const double* dense_rowmajor, int M, int N, double eps, bool ell_format);
  double get(int i, int j) const;
  void matmul(const double* b, double* c) const;
  void matmul_1_2(const double* b1, const double* b2, double* c) const;   // c = A (b1 .* b2)            (:386-408)
  void matmul2(const double* b, double* c, int NQ) const;                  // b(NQ,N) -> c(NQ,M), Fortran order (:411-431)
};

// ---------------------------------------------------------------- mesh
// mesh/scale_mesh_cubedom3d.F90:339-469,545-602; scale_mesh_base3d.F90:175-331;
// scale_meshutil_3d.F90:38-118,264-641,750-877
struct Mesh {
  int NeX = 0, NeY = 0, NeZ = 0, Ne = 0, NeA = 0, Ne2D = 0, Nhalo = 0;
  bool periodic[3] = {false, false, false};
  vec pos[3];           // (Np,Ne) each
  vec E11, E22, E33;    // Escale diagonal, (Np,Ne)
  vec J;                // (Np,Ne)
  vec nx, ny, nz, Fscale;  // (NfpTot,Ne)
  vec Gsqrt, G13, G23;  // (Np,NeA)
  vec GsqrtH;           // (Nfp,Ne2D)
  ivec vmapM, vmapP;    // (NfpTot,Ne) 0-based flat index into (Np*NeA)
  ivec vmapB;           // (Nhalo) 0-based
  ivec emap2d;          // (Ne)
  int halo_off[7] = {0};
  int nbr_face[6] = {0};  // single tile: face whose VMapB data fills halo face f
  void init_cube(const Element& e, int nex, int ney, int nez, double xmin, double xmax, double ymin,
                 double ymax, double zmin, double zmax, const double* FZ, const bool per[3]);
  void exchange_halo(const Element& e, double* field) const;  // single-tile Put/Exchange/Get
  // cubed-sphere panel tile (mesh/scale_mesh_cubedspheredom3d.F90): cube mesh in (alpha, beta, z) + horizontal metric
  bool is_global = false;
  int panelID = 0;
  double RPlanet = 0.0;
  vec alpha2D, beta2D, Gij11, Gij12, Gij22, GIJ11, GIJ12, GIJ22;   // (Nfp,Ne2D)
  vec gam;                                                           // (Np,NeA)
  void init_cubedsphere_panel(const Element& e, int panel, int nex, int ney, int nez, const double* FZ, double ztop,
                              double radius, bool shallow);
};

// ---------------------------------------------------------------- time integrator (a13)
// common/scale_timeint_rk_butcher_tab.F90:60-324, scale_timeint_rk.F90(.erb)
struct RKScheme {
  std::string name;
  int nstage = 0, tend_buf_size = 0;
  bool low_storage = false, imex = false;
  vec a_ex, b_ex, c_ex, a_im, b_im, c_im;  // row-major (nstage x nstage)
  vec sig, gam;                            // row-major ((nstage+1) x nstage)
  ivec indmap;                             // 0-based
  bool init(const std::string& scheme);
  double A_ex(int i, int j) const { return a_ex[i * nstage + j]; }
  double A_im(int i, int j) const { return a_im[i * nstage + j]; }
  double SIG(int i, int j) const { return sig[i * nstage + j]; }
  double GAM(int i, int j) const { return gam[i * nstage + j]; }
};

struct TimeIntRK {
  RKScheme sc;
  double dt = 0;
  int nvar = 0;
  size_t n = 0;  // values per variable
  vec tend_ex, tend_im, var0, varTmp;
  void init(const std::string& scheme, double dt_, int nvar_, size_t n_);
  double* tend_ex_buf(int var, int ind) { return &tend_ex[(size_t(ind) * nvar + var) * n]; }
  double* tend_im_buf(int var, int ind) { return &tend_im[(size_t(ind) * nvar + var) * n]; }
  double implicit_diagfac(int stage) const { return sc.imex ? sc.A_im(stage, stage) * dt : 0.0; }
  void store_var0(const double* q, int var, size_t is, size_t ie);
  void store_implicit(int stage, double* q, int var, size_t is, size_t ie);
  void advance(int stage, double* q, int var, size_t is, size_t ie);
};

// ---------------------------------------------------------------- dynamics
struct BndCfg { int vel_bc[6] = {0, 0, 0, 0, 0, 0}; };  // 0 nospec, 1 periodic, 2 slip, 3 noslip

struct DynState {
  // prognostic, (Np,NeA)
  vec DDENS, MOMX, MOMY, MOMZ, DRHOT;
  // auxiliary
  vec DENS_hyd, PRES_hyd, THERM_hyd, PRES_hyd_ref, Rtot, CVtot, CPtot, PRES, DPRES, DPhydDx, DPhydDy;
  vec CORIOLIS;  // (Nfp, Ne2D)
  // physics tendencies handed to the dynamics (driver_nonhydro3d.F90:843-857), (Np,NeA)
  vec DENS_tp, MOMX_tp, MOMY_tp, MOMZ_tp, RHOT_tp, RHOH_p;
  double* prog(int v) {
    switch (v) { case DENS_VID: return DDENS.data(); case RHOT_VID: return DRHOT.data();
      case MOMZ_VID: return MOMZ.data(); case MOMX_VID: return MOMX.data(); default: return MOMY.data(); }
  }
  void alloc(size_t n, size_t n2d);
};

void drhot2pres(const Element& e, const Mesh& m, const Consts& c, DynState& s);
void calc_rhot_hyd(const Element& e, const Mesh& m, const Consts& c, DynState& s);
void calc_phyd_hgrad(const Element& e, const Mesh& m, DynState& s);
void apply_bc_progvars(const Element& e, const Mesh& m, const BndCfg& b, DynState& s);
void heve_numflux_generalvc(const Element& e, const Mesh& m, const Consts& c, const DynState& s, vec& del_flux);
void heve_cal_tend(const Element& e, const Mesh& m, const Consts& c, const DynState& s, double* dt5[5]);
void modalfilter_apply(const Element& e, const Mesh& m, DynState& s);
void add_phy_tend(const Element& e, const Mesh& m, const Consts& c, const DynState& s, bool entot_conserve, double* dt5[5]);
// sponge layer (fluid_dyn_solver/scale_atm_dyn_dgm_spongelayer.F90:55-220)
struct SpongeCfg { bool on = false; double tau = -1.0, height = -1.0; int layer = -1; bool hveldamp = false; };
void sponge_add_tend(const Element& e, const Mesh& m, const SpongeCfg& cfg, const DynState& s, double* dt5[5]);

// HEVI (a6, a8-a12)
void hevi_numflux_generalvc(const Element& e, const Mesh& m, const Consts& c, const DynState& s, vec& del_flux);
void hevi_cal_tend(const Element& e, const Mesh& m, const Consts& c, const DynState& s, double* dt5[5]);
void hevi_cal_vi(const Element& e, const Mesh& m, const Consts& c, const DynState& s, const double* var0[5],
                 double impl_fac, double dt, double* dt5[5]);

// numerical diffusion (numdiff.cpp): PARAM_ATMOS_DYN_NUMDIFF + the boundary ids it reads
struct NumdiffCfg {
  int laplacian_num = 1;
  double coef_h = 0.0, coef_v = 0.0, dt = 0.0;
  int vel_bc[6] = {0, 0, 0, 0, 0, 0};     // 0 nospec, 2 slip, 3 noslip
  int therm_bc[6] = {0, 0, 0, 0, 0, 0};   // 1 adiabat
};
void numdiff_apply(const Element& e, const Mesh& m, const NumdiffCfg& cfg, DynState& s);

// global (cubed-sphere) HEVI: explicit part (dyn_global.cpp); the column solve is hevi_cal_vi
void global_hevi_numflux_generalhvc(const Element& e, const Mesh& m, const Consts& c, const DynState& s, vec& del_flux);
void global_hevi_cal_tend(const Element& e, const Mesh& m, const Consts& c, const DynState& s, double* dt5[5]);
void global_numflux_generalhvc(const Element& e, const Mesh& m, const Consts& c, const DynState& s, bool hevi, vec& del_flux);
void global_cal_tend(const Element& e, const Mesh& m, const Consts& c, const DynState& s, bool hevi, double* dt5[5]);

struct Driver {
  Element elem;
  Mesh mesh;
  Consts cst;
  BndCfg bnd;
  DynState st;
  TimeIntRK tint;
  bool hevi = false, modalfilter = false, global = false, phytend = false, entot_conserve = false, numdiff = false;
  NumdiffCfg nd;
  SpongeCfg sponge;
  // tracer coupling (QA > 0): stage-averaged mass fluxes and dissipation coefficients saved by the dynamics stages
  // (driver_nonhydro3d.F90:900-917, 926-937) for AtmDynDGMDriver_trcadv3d_update
  bool tracer = false;
  vec MFLX_x, MFLX_y, MFLX_z, alphM_tavg, alphP_tavg, DENS_TRC, DENS0_TRC;
  void update();  // fluid_dyn_solver/scale_atm_dyn_dgm_driver_nonhydro3d.F90:614-963
};

// tracer advection (tracer.cpp; row f4, restatement only: no device kernel yet)
void trc_face_int_weight(const Element& e, vec& W);
void trc_cal_alphdens_advtest(const Element& e, const Mesh& m, const double* DDENS, const double* MOMX, const double* MOMY,
                              const double* MOMZ, const double* DENS_hyd, vec& alphM, vec& alphP);
void trc_net_outward_flux(const Element& e, const Mesh& m, const vec& W, const double* Q, const double* MX, const double* MY,
                          const double* MZ, const vec& alphM, const vec& alphP, vec& net);
void trc_calc_fct_coef(const Element& e, const Mesh& m, const vec& W, const double* Q, const double* MX, const double* MY,
                       const double* MZ, const double* RHOQ_tp, const vec& alphM, const vec& alphP, const double* DENS_hyd,
                       const double* DDENS, const double* DDENS0, double rk_c_ssm1, double dt, bool disable_limiter, double* fct);
void trc_cal_tend(const Element& e, const Mesh& m, const vec& W, const double* Q, const double* MX, const double* MY, const double* MZ,
                  const vec& alphM, const vec& alphP, const double* fct, const double* RHOQ_tp, double* Q_dt);
void trc_tmar(const Element& e, const Mesh& m, const double* DENS_hyd, const double* DDENS, double* Q);
void trc_modalfilter(const Element& ef, const Mesh& m, const double* DENS_hyd, const double* DDENS, double* Q);
void rk_advance_trcvar_low_storage(const RKScheme& sc, double dt, int stage, size_t n, double* q, const double* DDENS,
                                   const double* DDENS0, const double* DENS_hyd, double* var0, double* varTmp, const double* tend);
struct Driver;
void trc_save_massflux(Driver& d, int stage);
void trcadv_update_coupled(Driver& d, const Element& elem_trcfilter, const RKScheme& sc, double dt, bool modalfilter,
                           bool disable_limiter, double* QTRC, const double* RHOQ_tp);
void trcadv_update_advtest(Driver& d, const Element& elem_trcfilter, const RKScheme& sc, double dt, bool modalfilter,
                           bool disable_limiter, double* QTRC, const double* RHOQ_tp);

// whole cubed sphere (sphere.cpp): panel-edge exchange and the six-panel step
void sphere_exchange(const Element& e, Mesh* const mesh[6], const std::vector<double*> scal[6], double* const u1[6], double* const u2[6]);
void sphere_update(Driver* d[6]);

// monitors: file/scale_file_monitor_meshfield.F90:176-213 + model mod_atmos_vars_container.F90:1281-1357
void monitor_sums(const Driver& d, double out[5]);  // DDENS mass, ENGT, ENGK, ENGI, ENGP

}  // namespace feo
