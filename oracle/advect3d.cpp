// TEST INFRASTRUCTURE ONLY (see fe_oracle.hpp).
// CPU restatement of sample/advect3d (BASELINE config 1; SURVEY.md rows a1 + a18):
//   advect3d_kernel_cal_tend     /root/reference/sample/advect3d/mod_advect3d_kernel.f90:34-82
//   cal_dqdt   (ELL/CSR SpMV)    mod_advect3d_kernel.f90:86-127
//   cal_elembnd_flux (upwind)    mod_advect3d_kernel.f90:131-166
//   stage loop                   /root/reference/sample/advect3d/test_advect3d.f90:81-126
// The derivative and lifting matrices are sparsemat objects built from the dense element matrices exactly as the
// sample does (Dx%Init(refElem%Dx1, storage_format='ELL'), test_advect3d.f90 init()), and every product goes through
// SparseMat::matmul (FElib/src/common/scale_sparsemat.F90:439-474, 554-634).
#include <stdexcept>

#include "fe_oracle.hpp"

namespace feo {

struct Advect3D {
  SparseMat Dx, Dy, Dz, Lift;
  TimeIntRK tint;
  vec q, u, v, w;   // (Np,NeA)
};

// mod_advect3d_kernel.f90:131-166
static void advect_elembnd_flux(const Element& e, const Mesh& m, const double* q, const double* u, const double* v, const double* w,
                                vec& flux) {
  const int NfpTot = e.NfpTot;
  flux.assign(size_t(NfpTot) * m.Ne, 0.0);
#pragma omp parallel for
  for (int ke = 0; ke < m.Ne; ++ke)
    for (int p = 0; p < NfpTot; ++p) {
      const size_t i = size_t(ke) * NfpTot + p;
      const int iM = m.vmapM[i], iP = m.vmapP[i];
      const double VelM = u[iM] * m.nx[i] + v[iM] * m.ny[i] + w[iM] * m.nz[i];
      const double VelP = u[iP] * m.nx[i] + v[iP] * m.ny[i] + w[iP] * m.nz[i];
      const double qM = q[iM], qP = q[iP];
      const double alpha = 0.5 * std::fabs(VelP + VelM);
      flux[i] = 0.5 * ((qP * VelP - qM * VelM) - alpha * (qP - qM));
    }
}

// mod_advect3d_kernel.f90:86-127
static void advect_dqdt(const Element& e, const Mesh& m, const Advect3D& a, const vec& flux, double* dqdt) {
  const int Np = e.Np, NfpTot = e.NfpTot;
#pragma omp parallel for
  for (int ke = 0; ke < m.Ne; ++ke) {
    vec in(Np), Fx(Np), Fy(Np), Fz(Np), fl(NfpTot), L(Np);
    const size_t b = size_t(ke) * Np;
    for (int n = 0; n < Np; ++n) in[n] = a.q[b + n] * a.u[b + n];
    a.Dx.matmul(in.data(), Fx.data());
    for (int n = 0; n < Np; ++n) in[n] = a.q[b + n] * a.v[b + n];
    a.Dy.matmul(in.data(), Fy.data());
    for (int n = 0; n < Np; ++n) in[n] = a.q[b + n] * a.w[b + n];
    a.Dz.matmul(in.data(), Fz.data());
    for (int p = 0; p < NfpTot; ++p) fl[p] = m.Fscale[size_t(ke) * NfpTot + p] * flux[size_t(ke) * NfpTot + p];
    a.Lift.matmul(fl.data(), L.data());
    for (int n = 0; n < Np; ++n)
      dqdt[b + n] = -(m.E11[b + n] * Fx[n] + m.E22[b + n] * Fy[n] + m.E33[b + n] * Fz[n] + L[n]);
  }
}

}  // namespace feo

using namespace feo;

namespace {
struct AdvHandle {
  const Element* e;
  const Mesh* m;
  Advect3D a;
};
}  // namespace

// element / mesh of a feo_create handle (defined in capi.cpp)
extern "C" const void* feo_elem_ptr(void* hv);
extern "C" const void* feo_mesh_ptr(void* hv);

extern "C" {

void* feo_advect3d_create(void* hv, const char* scheme, double dt, int ell) {
  auto* h = new AdvHandle;
  h->e = static_cast<const Element*>(feo_elem_ptr(hv));
  h->m = static_cast<const Mesh*>(feo_mesh_ptr(hv));
  const Element& e = *h->e;
  const double eps = 2.220446e-16 * 500.0;   // scale_sparsemat.F90:130 (CONST_EPS * 500)
  try {
    vec D;
    D = e.dmat_dense(0); h->a.Dx.init(D.data(), e.Np, e.Np, eps, ell != 0);
    D = e.dmat_dense(1); h->a.Dy.init(D.data(), e.Np, e.Np, eps, ell != 0);
    D = e.dmat_dense(2); h->a.Dz.init(D.data(), e.Np, e.Np, eps, ell != 0);
    D = e.lift_dense(); h->a.Lift.init(D.data(), e.Np, e.NfpTot, eps, ell != 0);
    const size_t n = size_t(e.Np) * h->m->NeA;
    h->a.tint.init(scheme, dt, 1, n);
    h->a.q.assign(n, 0.0); h->a.u.assign(n, 0.0); h->a.v.assign(n, 0.0); h->a.w.assign(n, 0.0);
  } catch (const std::exception&) { delete h; return nullptr; }
  return h;
}
void feo_advect3d_destroy(void* av) { delete static_cast<AdvHandle*>(av); }

double* feo_advect3d_array(void* av, const char* name, long* n) {
  auto* h = static_cast<AdvHandle*>(av);
  std::string s(name);
  vec* v = s == "q" ? &h->a.q : s == "u" ? &h->a.u : s == "v" ? &h->a.v : s == "w" ? &h->a.w : nullptr;
  if (!v) { *n = 0; return nullptr; }
  *n = long(v->size());
  return v->data();
}

// ELL / CSR arrays of the four operators, for the GPU conformance test of row a1: which = 0..3 (Dx, Dy, Dz, Lift)
void feo_advect3d_sparsemat(void* av, int which, int* M, int* N, int* col_size, const double** val, const int** colIdx) {
  auto* h = static_cast<AdvHandle*>(av);
  const SparseMat& s = which == 0 ? h->a.Dx : which == 1 ? h->a.Dy : which == 2 ? h->a.Dz : h->a.Lift;
  *M = s.M; *N = s.N; *col_size = s.col_size; *val = s.val.data(); *colIdx = s.colIdx.data();
}

// one evaluation of advect3d_kernel_cal_tend on the current q,u,v,w (halo exchanged first): dqdt (Np,Ne)
void feo_advect3d_cal_tend(void* av, double* dqdt) {
  auto* h = static_cast<AdvHandle*>(av);
  for (vec* f : {&h->a.q, &h->a.u, &h->a.v, &h->a.w}) h->m->exchange_halo(*h->e, f->data());
  vec flux;
  advect_elembnd_flux(*h->e, *h->m, h->a.q.data(), h->a.u.data(), h->a.v.data(), h->a.w.data(), flux);
  advect_dqdt(*h->e, *h->m, h->a, flux, dqdt);
}

// test_advect3d.f90:81-126: per stage  exchange -> cal_tend -> tint%Advance
void feo_advect3d_update(void* av, int nsteps) {
  auto* h = static_cast<AdvHandle*>(av);
  const size_t nint = size_t(h->e->Np) * h->m->Ne;
  vec flux;
  for (int step = 0; step < nsteps; ++step)
    for (int s = 0; s < h->a.tint.sc.nstage; ++s) {
      for (vec* f : {&h->a.q, &h->a.u, &h->a.v, &h->a.w}) h->m->exchange_halo(*h->e, f->data());
      advect_elembnd_flux(*h->e, *h->m, h->a.q.data(), h->a.u.data(), h->a.v.data(), h->a.w.data(), flux);
      advect_dqdt(*h->e, *h->m, h->a, flux, h->a.tint.tend_ex_buf(0, h->a.tint.sc.indmap[s]));
      h->a.tint.advance(s, h->a.q.data(), 0, 0, nint);
    }
}

}  // extern "C"
