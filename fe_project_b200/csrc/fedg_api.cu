// C-ABI layer + host-side driver of the dynamics step (see include/fedg.h).
//
// The host side mirrors AtmDynDGMDriver_nonhydro3d%Update
// (FElib/src/fluid_dyn_solver/scale_atm_dyn_dgm_driver_nonhydro3d.F90:614-963): per RK stage a halo
// fill (+ boundary condition) and ONE fused stage kernel; the prognostic state ping-pongs between
// device buffers so that neighbours always read the stage-input state.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "fedg_internal.h"
#include "rk_tables.h"

using namespace fedg;

namespace {
thread_local std::string g_err;

int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return fail(FEDG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));              \
  } while (0)

struct DevBuf {
  double* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t n_) {
    release();
    n = n_;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(double));
    // the memset runs on the legacy stream, the context's streams are non-blocking: without this wait an upload that
    // follows on the context stream could be overtaken by the zero fill (seen as a test-order dependent parity failure)
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};
}  // namespace

// sample/advect3d run on the mesh of a context (fedg_advect3d_*)
struct AdvectState {
  bool ready = false;
  RKTable rk;
  std::vector<RKStage> stages;
  bool vt_used = false;
  double dt = 0;
  DevBuf q[3], u, v, w, vt, tend, ellval[4];
  int* ellcol[4] = {nullptr, nullptr, nullptr, nullptr};
  int colsz[4] = {0, 0, 0, 0};
  int cur = 0, epb = 1;
  cudaGraphExec_t step_graph = nullptr;
  void release() {
    if (step_graph) { cudaGraphExecDestroy(step_graph); step_graph = nullptr; }
    for (auto& b : q) b.release();
    for (DevBuf* b : {&u, &v, &w, &vt, &tend}) b->release();
    for (auto& b : ellval) b.release();
    for (auto& p : ellcol) { if (p) cudaFree(p); p = nullptr; }
    ready = false;
  }
};

// tracer advection with a prescribed mass flux (fedg_trcadv_*)
struct TracerState {
  bool ready = false;
  RKTable rk;
  double dt = 0;
  bool disable_limiter = false, do_filter = false;
  DevBuf q[2], fct, var0, vartmp, alphM, alphP, rhoq, filt;
  double w1d[MAXNP] = {0};
  // coupling with the dynamics (fedg_trcadv_couple): stage-averaged mass flux and alphDens, density at the start of the step / after the RK loop
  bool couple = false, have_flux = false;
  int stage = 0;
  DevBuf mflx[3], alphM_t, alphP_t, dens0, dens1;
  void release() {
    for (DevBuf* b : {&q[0], &q[1], &fct, &var0, &vartmp, &alphM, &alphP, &rhoq, &filt}) b->release();
    ready = false;
  }
  void release_coupling() {
    for (DevBuf* b : {&mflx[0], &mflx[1], &mflx[2], &alphM_t, &alphP_t, &dens0, &dens1}) b->release();
    couple = have_flux = false;
  }
};

struct fedg_ctx {
  int np = 0, Np = 0, Nfp = 0, NfpTot = 0;
  int Ne = 0, NeA = 0, NeX = 0, NeY = 0, NeZ = 0, Ne2D = 0, Nhalo = 0;
  size_t nint = 0, nall = 0;  // Np*Ne, Np*Ne + Nhalo
  bool terrain = false, moist = false, has_cor = false, has_phyd = false;
  bool zface_contig = false;
  bool global = false; int panel = 0;   // cubed-sphere panel tile (GLOBALNONHYDRO3D_HEVI)
  bool dyn_ready = false, aux_ready = false;
  PhysConst c{};
  double OHM = 0;
  ElemTables tab{};
  RKTable rk;
  std::vector<RKStage> stages;
  bool vt_used = false, hevi = false, modalfilter = false;
  double dt = 0;
  int my_rank = 0, nbr_rank[6], nbr_face[6], vel_bc[6];
  int face_off[7];
  cudaStream_t stream = nullptr;
  // device data
  DevBuf prog[3][NVAR], dp[3], vt[NVAR], tendbuf[NVAR];
  bool dp_valid[3] = {false, false, false};  // dp[b] holds DPRES of prog[b] (interior)
  ElemTables* d_tab = nullptr;
  bool tab_dirty = true;
  DevBuf dens_hyd, pres_hyd, therm_hyd, rtot, cvtot, cptot, gsqrt, g13, g23, gsqrtH, dphydx, dphydy, coriolis;
  DevBuf escale, fscale, pres, w3, Jac, zlev, mon, g2d, phyt[6];
  bool has_phyt = false;
  // HEVI: stage tendencies k_ex / k_im [stage][var], var0-based IMEX combination, column-solver scratch
  std::vector<DevBuf> kex, kim;
  DevBuf rhot_hyd_vi, vi_scratch, vi_pvu, vi_pvv;
  double last_ms_vi = 0;
  int* d_vmapP = nullptr; int* d_emap2d = nullptr; int* d_vmapB = nullptr; int* d_halo_src = nullptr;
  // multi-GPU: NCCL state and the interior / tile-boundary element lists used to overlap the exchange
  CommState comm;
  int* d_elem_inner = nullptr; int* d_elem_bnd = nullptr; int n_inner = 0, n_bnd = 0;
  int cur = 0;
  AdvectState adv;
  // numerical diffusion (fedg_numdiff_init): PARAM_ATMOS_DYN_NUMDIFF, thermal BC ids, work fields
  struct { bool on = false, in_update = false; int lap_num = 1; double coef_h = 0, coef_v = 0; int therm_bc[6] = {0, 0, 0, 0, 0, 0}; } nd;
  DevBuf nd_g[3], nd_lap[2], nd_g5[15];
  DevBuf sponge; double sponge_h = 0.0; bool has_sponge = false;
  // halo faces filled from another local mesh on the same device (cubed-sphere panel edges): fedg_link_halo
  //   src != nullptr: gathered from that mesh;  recvbuf != nullptr: the data of a mesh on another rank, shipped by NCCL into
  //   recvbuf[6][cnt] (fedg_link_halo_recv) and scattered with the same rotation
  struct HaloLink { fedg_ctx* src = nullptr; int* d_src = nullptr; double* d_rot = nullptr; int off = 0, cnt = 0;
                    double* recvbuf = nullptr; int peer = -1, msg_id = 0; } link[6];
  // interior nodes of this mesh that feed a halo face of a mesh on another rank (fedg_link_halo_send)
  struct OutMsg { int peer = -1, msg_id = 0, cnt = 0; int* d_idx = nullptr; double* sendbuf = nullptr; };
  std::vector<OutMsg> outmsg;
  TracerState trc;
  // pipelined host update (fedg_dyn_update_host_async / _wait): device staging of two slots, one copy stream per direction
  static constexpr int NSLOT = 3;   // one slot uploads, one computes, one downloads
  struct HostPipe {
    bool ready = false;
    cudaStream_t h2d = nullptr, d2h = nullptr;
    DevBuf in[NSLOT][NVAR], out[NSLOT][NVAR];
    cudaEvent_t in_ready[NSLOT] = {}, in_free[NSLOT] = {}, out_ready[NSLOT] = {}, out_done[NSLOT] = {};
    bool pending[NSLOT] = {};
  } hp;
  int xbuf = 0;                    // buffer that holds the state other local meshes gather from (stage input of the explicit part)
  struct { int i0, in, mid, nxt; } hs{0, 0, 0, 0};   // buffer cursor of the HEVI stage pieces
  // stage-level seams (fedg_rk_store_var0 ... fedg_rk_advance): a step driven from outside, piece by piece
  struct { bool in_step = false, halo_inflight = false, halo_ready = false, vi_done = false; int halo_buf = 0; } seam;
  // timing
  bool profile = true;
  std::vector<cudaEvent_t> ev;
  double last_ms_total = 0, last_ms_stage = 0; long last_launches = 0;
  ~fedg_ctx() {
    for (auto& e : ev) cudaEventDestroy(e);
    for (int k = 0; k < NSLOT; ++k) {
      for (auto& b : hp.in[k]) b.release();
      for (auto& b : hp.out[k]) b.release();
      for (cudaEvent_t e : {hp.in_ready[k], hp.in_free[k], hp.out_ready[k], hp.out_done[k]}) if (e) cudaEventDestroy(e);
    }
    if (hp.h2d) cudaStreamDestroy(hp.h2d);
    if (hp.d2h) cudaStreamDestroy(hp.d2h);
    trc.release(); trc.release_coupling();
    for (auto& l : link) { if (l.d_src) cudaFree(l.d_src); if (l.d_rot) cudaFree(l.d_rot); if (l.recvbuf) cudaFree(l.recvbuf); }
    for (auto& m : outmsg) { if (m.d_idx) cudaFree(m.d_idx); if (m.sendbuf) cudaFree(m.sendbuf); }
    if (d_vmapP) cudaFree(d_vmapP);
    if (d_emap2d) cudaFree(d_emap2d);
    if (d_vmapB) cudaFree(d_vmapB);
    if (d_halo_src) cudaFree(d_halo_src);
    if (d_tab) cudaFree(d_tab);
    if (d_elem_inner) cudaFree(d_elem_inner);
    if (d_elem_bnd) cudaFree(d_elem_bnd);
    comm_destroy(comm);
    adv.release();
    for (auto& b : phyt) b.release();
    for (auto& b : nd_g) b.release();
    for (auto& b : nd_g5) b.release();
    for (auto& b : nd_lap) b.release();
    sponge.release();
    for (auto& b : dp) b.release();
    for (auto& s : prog) for (auto& b : s) b.release();
    for (auto& b : vt) b.release();
    for (auto& b : tendbuf) b.release();
    for (auto& b : kex) b.release();
    for (auto& b : kim) b.release();
    rhot_hyd_vi.release(); vi_scratch.release(); vi_pvu.release(); vi_pvv.release();
    for (DevBuf* b : {&dens_hyd, &pres_hyd, &therm_hyd, &rtot, &cvtot, &cptot, &gsqrt, &g13, &g23, &gsqrtH, &dphydx, &dphydy,
                      &coriolis, &escale, &fscale, &pres, &w3, &Jac, &zlev, &mon, &g2d})
      b->release();
    if (stream) cudaStreamDestroy(stream);
  }
};

static ElemTables g_loaded_tab{};
static bool g_tab_valid = false;
static void ensure_tables(fedg_ctx* c) {
  if (!c->d_tab) { cudaMalloc(&c->d_tab, sizeof(ElemTables)); c->tab_dirty = true; }
  if (c->tab_dirty) {
    cudaMemcpyAsync(c->d_tab, &c->tab, sizeof(ElemTables), cudaMemcpyHostToDevice, c->stream);  // pageable source: staged before return
    c->tab_dirty = false;
  }
  if (!g_tab_valid || std::memcmp(&g_loaded_tab, &c->tab, sizeof(ElemTables)) != 0) {
    upload_tables(c->tab, c->stream);
    g_loaded_tab = c->tab;
    g_tab_valid = true;
  }
}

static int upload(fedg_ctx* c, DevBuf& b, const double* host, size_t n) {
  if (b.n < n) CUDA_TRY(b.alloc(n));
  CUDA_TRY(cudaMemcpyAsync(b.p, host, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  return FEDG_OK;
}

extern "C" {

const char* fedg_last_error(void) { return g_err.c_str(); }
int fedg_version(void) { return 100; }

int fedg_create(const fedg_mesh_desc* d, fedg_ctx** out) {
  if (!d || !out) return fail(FEDG_ERR_ARG, "null argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(FEDG_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(ce) + " (this library has no CPU fallback)");
  const int np = d->polyorder + 1;
  // p = 7: tensor-core stage kernel; p = 1, 3, 5: the generic node-per-thread kernel (an even number of nodes per direction: its
  // rows are read as 128-bit pairs and an element's 8 np^3 bytes must keep the 16-byte alignment of the bulk copies)
  if (np != 8 && np != 6 && np != 4 && np != 2) return fail(FEDG_ERR_UNSUPPORTED, "polyorder must be 7, 5, 3 or 1 in this build");
  if (d->Ne != d->NeX * d->NeY * d->NeZ || d->Ne2D != d->NeX * d->NeY) return fail(FEDG_ERR_ARG, "Ne / NeX*NeY*NeZ mismatch");
  for (const void* p : {(const void*)d->D1D, (const void*)d->Lift, (const void*)d->VPOrdM1, (const void*)d->IntWeight_lgl,
                        (const void*)d->Escale, (const void*)d->Fscale, (const void*)d->normal_fn, (const void*)d->J,
                        (const void*)d->Gsqrt, (const void*)d->GI3, (const void*)d->GsqrtH, (const void*)d->zlev,
                        (const void*)d->VMapM, (const void*)d->VMapP, (const void*)d->VMapB, (const void*)d->EMap3Dto2D})
    if (!p) return fail(FEDG_ERR_ARG, "null array in fedg_mesh_desc");

  std::unique_ptr<fedg_ctx> c(new fedg_ctx);
  c->np = np; c->Nfp = np * np; c->Np = np * np * np; c->NfpTot = 6 * np * np;
  c->Ne = d->Ne; c->NeA = d->NeA; c->NeX = d->NeX; c->NeY = d->NeY; c->NeZ = d->NeZ; c->Ne2D = d->Ne2D; c->Nhalo = d->Nhalo;
  const int Np = c->Np, Nfp = c->Nfp, NfpTot = c->NfpTot, Ne = c->Ne;
  c->nint = size_t(Np) * Ne; c->nall = c->nint + size_t(c->Nhalo);
  if (size_t(Np) * d->NeA < c->nall) return fail(FEDG_ERR_ARG, "NeA too small for Nhalo");
  const int fsz[6] = {c->NeX * c->NeZ, c->NeY * c->NeZ, c->NeX * c->NeZ, c->NeY * c->NeZ, c->NeX * c->NeY, c->NeX * c->NeY};
  c->face_off[0] = 0;
  for (int f = 0; f < 6; ++f) c->face_off[f + 1] = c->face_off[f] + fsz[f] * Nfp;
  if (c->face_off[6] != c->Nhalo) return fail(FEDG_ERR_ARG, "Nhalo does not match the tile face sizes");
  c->c.GRAV = d->GRAV; c->c.Rdry = d->Rdry; c->c.CPdry = d->CPdry; c->c.CVdry = d->CVdry; c->c.PRES00 = d->PRES00;
  c->c.rP0 = 1.0 / d->PRES00; c->c.gamm = d->CPdry / d->CVdry; c->c.CPovCV = d->CPdry / d->CVdry; c->OHM = d->OHM;
  c->my_rank = d->my_rank;
  for (int f = 0; f < 6; ++f) {
    c->nbr_rank[f] = d->nbr_rank[f]; c->nbr_face[f] = d->nbr_face[f] - 1; c->vel_bc[f] = d->vel_bc[f];
    if (c->nbr_face[f] < 0 || c->nbr_face[f] > 5) return fail(FEDG_ERR_ARG, "nbr_face must be 1..6");
    if (fsz[f] != fsz[c->nbr_face[f]]) return fail(FEDG_ERR_ARG, "neighbour face size mismatch");
  }

  // ---- element tables
  ElemTables& T = c->tab;
  std::memset(&T, 0, sizeof(T));
  T.np = np;
  for (int i = 0; i < np; ++i)
    for (int l = 0; l < np; ++l) {
      T.D[i * np + l] = d->D1D[i + l * np];
      T.VP[i * np + l] = d->VPOrdM1[i + l * np];
      T.Fh[i * np + l] = T.Fv[i * np + l] = (i == l) ? 1.0 : 0.0;
    }
  auto LiftAt = [&](int n, int col) { return d->Lift[size_t(n) + size_t(col) * Np]; };
  for (int m = 0; m < np; ++m) {
    T.Lw[m * 2] = LiftAt(m, 3 * Nfp);                // x- face, face node (j=0,k=0), volume node (m,0,0)
    T.Lw[m * 2 + 1] = LiftAt(m, 1 * Nfp);            // x+ face
  }
  {  // the lifting matrix must have the tensor-product structure the TensorProd3D operator assumes
    double scale = 0.0, dev = 0.0, offmax = 0.0;
    for (int m = 0; m < 2 * np; ++m) scale = std::max(scale, std::fabs(T.Lw[m]));
    for (int k = 0; k < np; ++k) for (int j = 0; j < np; ++j) for (int i = 0; i < np; ++i) {
      int n = i + j * np + k * np * np;
      const int cols[6] = {i + k * np, Nfp + j + k * np, 2 * Nfp + i + k * np, 3 * Nfp + j + k * np, 4 * Nfp + i + j * np, 5 * Nfp + i + j * np};
      const double w[6] = {T.Lw[j * 2], T.Lw[i * 2 + 1], T.Lw[j * 2 + 1], T.Lw[i * 2], T.Lw[k * 2], T.Lw[k * 2 + 1]};
      for (int f = 0; f < 6; ++f) dev = std::max(dev, std::fabs(LiftAt(n, cols[f]) - w[f]));
      // one off-pattern probe per node
      int colx = (cols[0] + 1) % Nfp;
      if (colx != cols[0]) offmax = std::max(offmax, std::fabs(LiftAt(n, colx)));
    }
    if (dev > 1e-9 * scale || offmax > 1e-9 * scale)
      return fail(FEDG_ERR_UNSUPPORTED, "elem%Lift is not of tensor-product (I x I x invM1D e_face) form");
  }

  // ---- geometry: compress what is constant per element / per element face on MeshCubeDom3D
  std::vector<double> esc(size_t(3) * Ne), fsc(size_t(6) * Ne);
  for (int ke = 0; ke < Ne; ++ke) {
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
      const double* E = d->Escale + (size_t(a) + 3 * size_t(b)) * size_t(Np) * Ne + size_t(ke) * Np;
      double v0 = E[0];
      for (int p = 1; p < Np; ++p)
        if (std::fabs(E[p] - v0) > 1e-12 * std::fabs(v0)) return fail(FEDG_ERR_UNSUPPORTED, "Escale varies inside an element");
      if (a == b) esc[size_t(a) * Ne + ke] = v0;
      else if (v0 != 0.0) return fail(FEDG_ERR_UNSUPPORTED, "Escale has off-diagonal entries");
    }
    for (int f = 0; f < 6; ++f) {
      const double* F = d->Fscale + size_t(ke) * NfpTot + f * Nfp;
      for (int p = 1; p < Nfp; ++p)
        if (std::fabs(F[p] - F[0]) > 1e-12 * std::fabs(F[0])) return fail(FEDG_ERR_UNSUPPORTED, "Fscale varies on a face");
      fsc[size_t(f) * Ne + ke] = F[0];
      const int ax = (f == 1 || f == 3) ? 0 : (f == 0 || f == 2) ? 1 : 2;
      const double sg = (f == 1 || f == 2 || f == 5) ? 1.0 : -1.0;
      for (int p = 0; p < Nfp; ++p)
        for (int a = 0; a < 3; ++a) {
          double nv = d->normal_fn[size_t(a) * NfpTot * Ne + size_t(ke) * NfpTot + f * Nfp + p];
          if (nv != (a == ax ? sg : 0.0)) return fail(FEDG_ERR_UNSUPPORTED, "normal_fn is not axis aligned");
        }
    }
  }
  bool terrain = false;
  for (size_t n = 0; n < c->nall && !terrain; ++n)
    if (d->Gsqrt[n] != 1.0 || d->GI3[n] != 0.0 || d->GI3[size_t(Np) * d->NeA + n] != 0.0) terrain = true;
  for (size_t n = 0; n < size_t(Nfp) * c->Ne2D && !terrain; ++n) if (d->GsqrtH[n] != 1.0) terrain = true;
  // ---- cubed-sphere panel tile: the horizontal Jacobian lives in 2D tables, the kernels take the flat-geometry path
  if (d->panelID != 0) {
    if (d->panelID < 1 || d->panelID > 6 || !d->GIJ || !d->gam || !d->pos2D) return fail(FEDG_ERR_ARG, "panelID set but GIJ / gam / pos2D missing");
    if (np != 8) return fail(FEDG_ERR_UNSUPPORTED, "the global equation set is built for p = 7 only");
    const size_t n2 = size_t(Nfp) * c->Ne2D;
    for (int ke = 0; ke < Ne; ++ke)
      for (int p = 0; p < Np; ++p) {
        const size_t i = size_t(ke) * Np + p;
        const double gh = d->GsqrtH[size_t(d->EMap3Dto2D[ke] - 1) * Nfp + (p % Nfp)];
        if (d->gam[i] != 1.0 || d->GI3[i] != 0.0 || d->GI3[size_t(Np) * d->NeA + i] != 0.0 || d->Gsqrt[i] != gh)
          return fail(FEDG_ERR_UNSUPPORTED, "global panel: only the shallow-atmosphere metric without topography is available (gam = 1, GI3 = 0, Gsqrt = GsqrtH)");
      }
    for (size_t h = c->nint; h < c->nall; ++h) {   // fill_halo_metric: halo metric = own face value
      const long src = long(d->VMapB[h - c->nint]) - 1;
      if (d->Gsqrt[h] != d->Gsqrt[src]) return fail(FEDG_ERR_UNSUPPORTED, "global panel: halo Gsqrt must equal the own face value");
    }
    std::vector<double> g2(6 * n2);
    for (size_t i = 0; i < n2; ++i) {
      g2[i] = d->GsqrtH[i];
      g2[n2 + i] = d->GIJ[i];                 // (1,1)
      g2[2 * n2 + i] = d->GIJ[2 * n2 + i];    // (1,2): Fortran (Nfp,Ne2D,2,2) -> offset ((j-1)*2 + (i-1)) * n2
      g2[3 * n2 + i] = d->GIJ[3 * n2 + i];    // (2,2)
      g2[4 * n2 + i] = std::tan(d->pos2D[i]);
      g2[5 * n2 + i] = std::tan(d->pos2D[n2 + i]);
      if (std::fabs(d->GIJ[n2 + i] - d->GIJ[2 * n2 + i]) > 1e-14 * std::fabs(d->GIJ[i])) return fail(FEDG_ERR_ARG, "GIJ is not symmetric");
    }
    c->global = true; c->panel = d->panelID;
    terrain = false;
    c->g2d.n = 0;
    { cudaError_t e = c->g2d.alloc(g2.size()); if (e != cudaSuccess) return fail(FEDG_ERR_CUDA, cudaGetErrorString(e)); }
    { cudaError_t e = cudaMemcpy(c->g2d.p, g2.data(), g2.size() * sizeof(double), cudaMemcpyHostToDevice); if (e != cudaSuccess) return fail(FEDG_ERR_CUDA, cudaGetErrorString(e)); }
  }
  c->terrain = terrain;

  // ---- connectivity (to 0-based) and the same-rank halo source map
  std::vector<int> vP(size_t(NfpTot) * Ne), e2(Ne), vB(std::max(c->Nhalo, 1)), src(std::max(c->Nhalo, 1));
  for (int ke = 0; ke < Ne; ++ke) {
    e2[ke] = d->EMap3Dto2D[ke] - 1;
    if (e2[ke] < 0 || e2[ke] >= c->Ne2D) return fail(FEDG_ERR_ARG, "EMap3Dto2D out of range");
    for (int f = 0; f < 6; ++f) for (int p = 0; p < Nfp; ++p) {
      size_t q = size_t(ke) * NfpTot + f * Nfp + p;
      int nloc;
      switch (f) {
        case 0: nloc = (p % np) + (p / np) * Nfp; break;
        case 1: nloc = (np - 1) + (p % np) * np + (p / np) * Nfp; break;
        case 2: nloc = (p % np) + (np - 1) * np + (p / np) * Nfp; break;
        case 3: nloc = (p % np) * np + (p / np) * Nfp; break;
        case 4: nloc = p; break;
        default: nloc = p + (np - 1) * Nfp; break;
      }
      if (d->VMapM[q] - 1 != ke * Np + nloc) return fail(FEDG_ERR_UNSUPPORTED, "VMapM does not follow the hexahedral Fmask order");
      long vp = long(d->VMapP[q]) - 1;
      if (vp < 0 || vp >= long(c->nall)) return fail(FEDG_ERR_ARG, "VMapP out of range");
      vP[q] = int(vp);
    }
  }
  // z faces whose exterior nodes are Nfp consecutive, 16-byte aligned values (the element above / below, or a halo face):
  // stage_p7 fetches them with one bulk copy per field instead of per-node gathers
  c->zface_contig = true;
  for (int ke = 0; ke < Ne && c->zface_contig; ++ke)
    for (int f = 4; f < 6 && c->zface_contig; ++f) {
      const int* row = &vP[size_t(ke) * NfpTot + f * Nfp];
      if (row[0] % 2 != 0) c->zface_contig = false;
      for (int p = 1; p < Nfp; ++p) if (row[p] != row[0] + p) { c->zface_contig = false; break; }
    }
  for (int h = 0; h < c->Nhalo; ++h) {
    vB[h] = d->VMapB[h] - 1;
    if (vB[h] < 0 || size_t(vB[h]) >= c->nint) return fail(FEDG_ERR_ARG, "VMapB out of range");
  }
  for (int f = 0; f < 6; ++f) {
    const int fo = c->nbr_face[f];
    const bool local = c->nbr_rank[f] == c->my_rank;
    for (int m = 0; m < fsz[f] * Nfp; ++m) src[c->face_off[f] + m] = local ? vB[c->face_off[fo] + m] : -1;
  }

  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaMalloc(&c->d_vmapP, vP.size() * sizeof(int)));
  CUDA_TRY(cudaMalloc(&c->d_emap2d, e2.size() * sizeof(int)));
  CUDA_TRY(cudaMalloc(&c->d_vmapB, vB.size() * sizeof(int)));
  CUDA_TRY(cudaMalloc(&c->d_halo_src, src.size() * sizeof(int)));
  CUDA_TRY(cudaMemcpy(c->d_vmapP, vP.data(), vP.size() * sizeof(int), cudaMemcpyHostToDevice));
  {
    // VMapP is read once per stage by every element and sits at the head of a dependent load chain (index, then the gather):
    // keep it in the persisting part of L2 so that the 1.3 GB streamed per stage does not evict it (FEDG_L2_PERSIST=0: off)
    const char* e = getenv("FEDG_L2_PERSIST");
    int dev = 0, max_persist = 0, max_window = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    const size_t bytes = vP.size() * sizeof(int);
    if (!(e && e[0] == '0') && max_persist > 0 && max_window > 0) {
      const size_t set_aside = std::min<size_t>(size_t(max_persist), std::max<size_t>(bytes, size_t(1) << 20));
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, set_aside);
      cudaStreamAttrValue av{};
      av.accessPolicyWindow.base_ptr = c->d_vmapP;
      av.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, size_t(max_window));
      av.accessPolicyWindow.hitRatio = float(std::min(1.0, double(set_aside) / double(av.accessPolicyWindow.num_bytes)));
      av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
    }
  }
  CUDA_TRY(cudaMemcpy(c->d_emap2d, e2.data(), e2.size() * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(c->d_vmapB, vB.data(), vB.size() * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(c->d_halo_src, src.data(), src.size() * sizeof(int), cudaMemcpyHostToDevice));
  fedg_ctx* cc = c.get();
  int rc;
  if ((rc = upload(cc, c->escale, esc.data(), esc.size()))) return rc;
  if ((rc = upload(cc, c->fscale, fsc.data(), fsc.size()))) return rc;
  if ((rc = upload(cc, c->w3, d->IntWeight_lgl, Np))) return rc;
  if ((rc = upload(cc, c->Jac, d->J, c->nint))) return rc;
  if ((rc = upload(cc, c->zlev, d->zlev, c->nint))) return rc;
  if (c->global) { if ((rc = upload(cc, c->gsqrt, d->Gsqrt, c->nall))) return rc; }   // weights of the modal filter / monitors
  if (terrain) {
    if ((rc = upload(cc, c->gsqrt, d->Gsqrt, c->nall))) return rc;
    if ((rc = upload(cc, c->g13, d->GI3, c->nall))) return rc;
    if ((rc = upload(cc, c->g23, d->GI3 + size_t(Np) * d->NeA, c->nall))) return rc;
    if ((rc = upload(cc, c->gsqrtH, d->GsqrtH, size_t(Nfp) * c->Ne2D))) return rc;
  }
  for (auto& s : c->prog) for (auto& b : s) CUDA_TRY(b.alloc(c->nall));
  for (DevBuf* b : {&c->dens_hyd, &c->pres_hyd, &c->therm_hyd, &c->pres}) CUDA_TRY(b->alloc(c->nall));
  for (auto& b : c->dp) CUDA_TRY(b.alloc(c->nall));
  CUDA_TRY(c->mon.alloc(8));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  *out = c.release();
  return FEDG_OK;
}

void fedg_destroy(fedg_ctx* ctx) { delete ctx; }

int fedg_rk_info(const char* scheme, int* nstage, int* tend_buf_size, int* low_storage, int* imex) {
  RKTable t;
  if (!scheme || !t.init(scheme)) return fail(FEDG_ERR_ARG, std::string("unsupported RK scheme ") + (scheme ? scheme : "(null)"));
  *nstage = t.nstage; *tend_buf_size = t.tend_buf_size; *low_storage = t.low_storage; *imex = t.imex;
  return FEDG_OK;
}
int fedg_rk_coef(const char* scheme, double* a_ex, double* b_ex, double* a_im, double* b_im, double* sig, double* gam) {
  RKTable t;
  if (!scheme || !t.init(scheme)) return fail(FEDG_ERR_ARG, "unsupported RK scheme");
  std::memcpy(a_ex, t.a_ex.data(), t.a_ex.size() * 8); std::memcpy(b_ex, t.b_ex.data(), t.b_ex.size() * 8);
  std::memcpy(a_im, t.a_im.data(), t.a_im.size() * 8); std::memcpy(b_im, t.b_im.data(), t.b_im.size() * 8);
  std::memcpy(sig, t.sig.data(), t.sig.size() * 8); std::memcpy(gam, t.gam.data(), t.gam.size() * 8);
  return FEDG_OK;
}

// Stage coefficients of the fused update from the scheme tables
// (rk_advance_low_storage2D scale_timeint_rk.F90:1182-1266, rk_advance_general2D :2201-2355).
static void build_stage_coefs(const RKTable& t, double dt, std::vector<RKStage>& stages, bool& vt_used) {
  const int s = t.nstage;
  const double EPS = 2.220446e-16;
  stages.assign(s, RKStage{});
  vt_used = false;
  if (t.low_storage) {
    for (int n = 0; n < s - 1; ++n)
      if (std::fabs(t.sg(s, n)) > EPS || std::fabs(t.gm(s, n)) > EPS) vt_used = true;
    for (int n = 0; n < s; ++n) {
      RKStage& r = stages[n];
      const double sig_ss = t.sg(n + 1, n), gam_ss = dt * t.gm(n + 1, n);
      r.c_q = sig_ss; r.c_k = gam_ss;
      if (n == s - 1) { r.add_vt = vt_used; continue; }
      r.c_q0 = 1.0 - sig_ss; r.use_q0 = (r.c_q0 != 0.0);
      const double sig_Ns = t.sg(s, n), gam_Ns = dt * t.gm(s, n);
      const bool upd = std::fabs(sig_Ns) > EPS || std::fabs(t.gm(s, n)) > EPS;
      if (vt_used && (upd || n == 0)) {
        r.vt_update = 1; r.vt_init = (n == 0); r.vt_init_q = 0.0;
        r.vt_q = upd ? sig_Ns : 0.0; r.vt_k = upd ? gam_Ns : 0.0;
      }
    }
  } else {  // general explicit scheme with tend_buf_size == 1
    for (int n = 0; n < s; ++n) {
      RKStage& r = stages[n];
      if (s == 1) { r.c_q = 1.0; r.c_k = dt * t.b_ex[0]; continue; }
      vt_used = true;
      if (n == s - 1) { r.add_vt = 1; r.c_q = 0.0; r.c_k = dt * t.b_ex[n]; continue; }
      r.use_q0 = 1; r.c_q0 = 1.0; r.c_q = 0.0; r.c_k = dt * t.aex(n + 1, n);
      r.vt_update = 1; r.vt_init = (n == 0); r.vt_init_q = 1.0; r.vt_q = 0.0; r.vt_k = dt * t.b_ex[n];
    }
  }
}

static void build_stages(fedg_ctx* c) { build_stage_coefs(c->rk, c->dt, c->stages, c->vt_used); }

int fedg_dyn_init(fedg_ctx* c, const char* eqs_type, const char* tinteg_type, double dt, int modalfilter_flag,
                  const double* filter_h1D, const double* filter_v1D) {
  if (!c || !eqs_type || !tinteg_type) return fail(FEDG_ERR_ARG, "null argument");
  std::string eqs(eqs_type);
  if (eqs == "NONHYDRO3D_HEVE") c->hevi = false;
  else if (eqs == "NONHYDRO3D_HEVI") c->hevi = true;
  else if (eqs == "GLOBALNONHYDRO3D_HEVI" || eqs == "GLOBALNONHYDRO3D_HEVE") {
    if (!c->global) return fail(FEDG_ERR_ARG, eqs + " needs a cubed-sphere panel mesh (fedg_mesh_desc.panelID)");
    c->hevi = (eqs == "GLOBALNONHYDRO3D_HEVI");
  }
  else return fail(FEDG_ERR_UNSUPPORTED, "EQS_TYPE " + eqs + " is not available in this build (NONHYDRO3D_HEVE, NONHYDRO3D_HEVI, GLOBALNONHYDRO3D_HEVE, GLOBALNONHYDRO3D_HEVI)");
  if (c->global && eqs.rfind("GLOBAL", 0) != 0) return fail(FEDG_ERR_ARG, "a cubed-sphere panel mesh runs the GLOBALNONHYDRO3D equation sets only");
  if (!c->rk.init(tinteg_type)) return fail(FEDG_ERR_ARG, std::string("unsupported TINTEG_TYPE ") + tinteg_type);
  if (!(dt > 0.0)) return fail(FEDG_ERR_ARG, "dt must be positive");
  c->dt = dt;
  if (c->hevi) {
    // driver_nonhydro3d.F90:437-452: the HEVI equation sets run with an IMEX scheme
    if (!c->rk.imex) return fail(FEDG_ERR_ARG, "HEVI needs an IMEX scheme (IMEX_ARK232, IMEX_ARK324)");
    if (c->np != 8) return fail(FEDG_ERR_UNSUPPORTED, "the vertical-implicit column kernel is built for p = 7 only");
    // terrain-following HEVI (regional): the eight-lane column kernel with the metric terms, two passes per implicit stage
    if (c->terrain) { if (c->vi_pvu.n < c->nint) CUDA_TRY(c->vi_pvu.alloc(c->nint)); if (c->vi_pvv.n < c->nint) CUDA_TRY(c->vi_pvv.alloc(c->nint)); }
    if (2 * c->rk.nstage > MAXTERM) return fail(FEDG_ERR_UNSUPPORTED, "too many IMEX stages");
    c->kex.resize(size_t(c->rk.nstage) * NVAR); c->kim.resize(size_t(c->rk.nstage) * NVAR);
    for (auto& b : c->kex) if (b.n < c->nint) CUDA_TRY(b.alloc(c->nint));
    for (auto& b : c->kim) if (b.n < c->nint) CUDA_TRY(b.alloc(c->nint));
    const size_t nscr = size_t(c->NeZ) * 120 * size_t(c->Ne2D) * 64;
    if (c->vi_scratch.n < nscr) CUDA_TRY(c->vi_scratch.alloc(nscr));
  } else {
    if (c->rk.imex) return fail(FEDG_ERR_ARG, "HEVE needs an explicit RK scheme");
    if (c->rk.tend_buf_size != 1) return fail(FEDG_ERR_UNSUPPORTED, "explicit schemes with several tendency buffers are not supported");
    build_stages(c);
    if (c->vt_used) for (auto& b : c->vt) if (b.n < c->nall) CUDA_TRY(b.alloc(c->nall));
  }
  c->modalfilter = modalfilter_flag != 0;
  const int np = c->np;
  for (int i = 0; i < np; ++i)
    for (int l = 0; l < np; ++l) {
      c->tab.Fh[i * np + l] = c->modalfilter ? 0.0 : (i == l ? 1.0 : 0.0);
      c->tab.Fv[i * np + l] = c->tab.Fh[i * np + l];
    }
  if (c->modalfilter) {
    if (!filter_h1D || !filter_v1D) return fail(FEDG_ERR_ARG, "modal filter matrices missing");
    for (int i = 0; i < np; ++i)
      for (int l = 0; l < np; ++l) { c->tab.Fh[i * np + l] = filter_h1D[i + l * np]; c->tab.Fv[i * np + l] = filter_v1D[i + l * np]; }
  }
  c->tab_dirty = true;
  c->dyn_ready = true;
  return FEDG_OK;
}

int fedg_set_prog(fedg_ctx* c, const double* DDENS, const double* MOMX, const double* MOMY, const double* MOMZ, const double* DRHOT) {
  if (!c || !DDENS || !MOMX || !MOMY || !MOMZ || !DRHOT) return fail(FEDG_ERR_ARG, "null argument");
  const double* h[NVAR] = {DDENS, MOMX, MOMY, MOMZ, DRHOT};
  for (int v = 0; v < NVAR; ++v)
    CUDA_TRY(cudaMemcpyAsync(c->prog[c->cur][v].p, h[v], c->nall * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->dp_valid[c->cur] = false;
  c->xbuf = c->cur;
  return FEDG_OK;
}

int fedg_get_prog(fedg_ctx* c, double* DDENS, double* MOMX, double* MOMY, double* MOMZ, double* DRHOT) {
  if (!c || !DDENS || !MOMX || !MOMY || !MOMZ || !DRHOT) return fail(FEDG_ERR_ARG, "null argument");
  double* h[NVAR] = {DDENS, MOMX, MOMY, MOMZ, DRHOT};
  for (int v = 0; v < NVAR; ++v)
    CUDA_TRY(cudaMemcpyAsync(h[v], c->prog[c->cur][v].p, c->nall * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return FEDG_OK;
}

// gather own-face values into the halo slots of an auxiliary field (same-rank faces)
static int fill_aux_halo(fedg_ctx* c, double* field);
// halo of the auxiliary fields across ranks (AUX_VARS exchange after a restart is read, model mod_atmos_vars.F90:553-636)
static int exchange_aux_remote(fedg_ctx* c);

int fedg_set_aux(fedg_ctx* c, const double* DENS_hyd, const double* PRES_hyd, const double* THERM_hyd, const double* Rtot,
                 const double* CVtot, const double* CPtot) {
  if (!c || !DENS_hyd || !PRES_hyd || !Rtot || !CVtot || !CPtot) return fail(FEDG_ERR_ARG, "null argument");
  int rc;
  if ((rc = upload(c, c->dens_hyd, DENS_hyd, c->nint))) return rc;
  if ((rc = upload(c, c->pres_hyd, PRES_hyd, c->nint))) return rc;
  if (THERM_hyd) { if ((rc = upload(c, c->therm_hyd, THERM_hyd, c->nint))) return rc; }
  else launch_calc_rhot_hyd(c->pres_hyd.p, c->c, c->therm_hyd.p, long(c->nint), c->stream);
  // the vertical-implicit solver recomputes RHOT_hyd from PRES_hyd with dry constants (hevi_common_2.F90:211-212, 1253-1254)
  if (c->rhot_hyd_vi.n < c->nall) CUDA_TRY(c->rhot_hyd_vi.alloc(c->nall));
  launch_calc_rhot_hyd(c->pres_hyd.p, c->c, c->rhot_hyd_vi.p, long(c->nint), c->stream);
  bool moist = false;
  for (size_t n = 0; n < c->nint && !moist; ++n)
    if (Rtot[n] != c->c.Rdry || CVtot[n] != c->c.CVdry || CPtot[n] != c->c.CPdry) moist = true;
  c->moist = moist;
  if (moist) {
    for (DevBuf* b : {&c->rtot, &c->cvtot, &c->cptot}) if (b->n < c->nall) CUDA_TRY(b->alloc(c->nall));
    if ((rc = upload(c, c->rtot, Rtot, c->nint))) return rc;
    if ((rc = upload(c, c->cvtot, CVtot, c->nint))) return rc;
    if ((rc = upload(c, c->cptot, CPtot, c->nint))) return rc;
  }
  for (DevBuf* b : {&c->dens_hyd, &c->pres_hyd, &c->therm_hyd}) if ((rc = fill_aux_halo(c, b->p))) return rc;
  if (moist) for (DevBuf* b : {&c->rtot, &c->cvtot, &c->cptot}) if ((rc = fill_aux_halo(c, b->p))) return rc;
  if ((rc = exchange_aux_remote(c))) return rc;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->aux_ready = true;
  for (bool& v : c->dp_valid) v = false;
  return FEDG_OK;
}

int fedg_set_phyd_hgrad(fedg_ctx* c, const double* DPhydDx, const double* DPhydDy) {
  if (!c) return fail(FEDG_ERR_ARG, "null argument");
  c->has_phyd = false;
  if (!DPhydDx || !DPhydDy) return FEDG_OK;
  bool nz = false;
  for (size_t n = 0; n < c->nint && !nz; ++n) if (DPhydDx[n] != 0.0 || DPhydDy[n] != 0.0) nz = true;
  if (!nz) return FEDG_OK;
  int rc;
  if ((rc = upload(c, c->dphydx, DPhydDx, c->nint))) return rc;
  if ((rc = upload(c, c->dphydy, DPhydDy, c->nint))) return rc;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->has_phyd = true;
  return FEDG_OK;
}

// update_phyd_hgrad (driver_nonhydro3d.F90:1060-1095): DPhydDx / DPhydDy from the PRES_hyd on the device, whose halo holds what the
// last aux exchange left there (fedg_set_aux: own tile + NCCL neighbours; fedg_group_exchange_aux: linked local meshes)
int fedg_update_phyd_hgrad(fedg_ctx* c, const double* PRES_hyd_ref) {
  if (!c) return fail(FEDG_ERR_ARG, "null argument");
  if (!c->aux_ready) return fail(FEDG_ERR_STATE, "fedg_set_aux must be called first");
  ensure_tables(c);
  for (DevBuf* b : {&c->dphydx, &c->dphydy}) if (b->n < c->nint) CUDA_TRY(b->alloc(c->nint));
  DevBuf ref;
  struct Rel { DevBuf& b; ~Rel() { b.release(); } } rel{ref};
  if (PRES_hyd_ref) {
    CUDA_TRY(ref.alloc(c->nall));
    CUDA_TRY(cudaMemcpyAsync(ref.p, PRES_hyd_ref, c->nall * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  launch_phyd_hgrad(c->pres_hyd.p, PRES_hyd_ref ? ref.p : nullptr, c->gsqrt.p, c->g13.p, c->g23.p, c->gsqrtH.p, c->escale.p, c->fscale.p,
                    c->d_vmapP, c->d_emap2d, c->dphydx.p, c->dphydy.p, c->np, c->Ne, c->terrain, c->stream);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  c->has_phyd = true;
  return FEDG_OK;
}

int fedg_set_phy_tend(fedg_ctx* c, const double* DENS_tp, const double* MOMX_tp, const double* MOMY_tp, const double* MOMZ_tp,
                      const double* RHOT_tp, const double* RHOH_p) {
  if (!c) return fail(FEDG_ERR_ARG, "null argument");
  c->has_phyt = false;
  const double* h[6] = {DENS_tp, MOMX_tp, MOMY_tp, MOMZ_tp, RHOT_tp, RHOH_p};
  for (const double* p : h) if (!p) return FEDG_OK;           // NULL: no physics tendencies
  bool nz = false;                                            // all zero (dry dynamics-only run): skip the six extra reads
  for (int k = 0; k < 6 && !nz; ++k)
    for (size_t n = 0; n < c->nint && !nz; ++n) if (h[k][n] != 0.0) nz = true;
  if (!nz) return FEDG_OK;
  for (int k = 0; k < 6; ++k) { int rc = upload(c, c->phyt[k], h[k], c->nint); if (rc) return rc; }
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->has_phyt = true;
  return FEDG_OK;
}

int fedg_set_coriolis(fedg_ctx* c, const double* cor) {
  if (!c) return fail(FEDG_ERR_ARG, "null argument");
  c->has_cor = false;
  if (!cor) return FEDG_OK;
  const size_t n2 = size_t(c->Nfp) * c->Ne2D;
  bool nz = false;
  for (size_t n = 0; n < n2 && !nz; ++n) if (cor[n] != 0.0) nz = true;
  if (!nz) return FEDG_OK;
  int rc;
  if ((rc = upload(c, c->coriolis, cor, n2))) return rc;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->has_cor = true;
  return FEDG_OK;
}

}  // extern "C"

// ---- host-side stage sequencing ------------------------------------------------------------
namespace {

__global__ void aux_halo_kernel(double* q, const int* src, int nint, int nhalo) {
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h < nhalo && src[h] >= 0) q[size_t(nint) + h] = q[src[h]];
}

// halo slots of a face linked to another local mesh: gather the six fields, re-express (MOMX, MOMY) in the own basis
// (2x2 matrix per node = LonLat2CSVec(own panel, own face node) o CS2LonLatVec(source panel, source node))
struct LinkFields { const double* s[6]; double* d[6]; };
__global__ void halo_link_kernel(LinkFields F, const int* __restrict__ src, const double* __restrict__ rot, size_t dst0, int cnt) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= cnt) return;
  const int i = src ? src[m] : m;
  const size_t o = dst0 + m;
  F.d[V_DDENS][o] = F.s[V_DDENS][i]; F.d[V_MOMZ][o] = F.s[V_MOMZ][i]; F.d[V_DRHOT][o] = F.s[V_DRHOT][i]; F.d[5][o] = F.s[5][i];
  const double sx = F.s[V_MOMX][i], sy = F.s[V_MOMY][i];
  double r00 = 1.0, r01 = 0.0, r10 = 0.0, r11 = 1.0;
  if (rot) { r00 = rot[4 * size_t(m)]; r01 = rot[4 * size_t(m) + 1]; r10 = rot[4 * size_t(m) + 2]; r11 = rot[4 * size_t(m) + 3]; }
  F.d[V_MOMX][o] = r00 * sx + r01 * sy;
  F.d[V_MOMY][o] = r10 * sx + r11 * sy;
}

void fill_halo(fedg_ctx* c, int buf, bool apply_bc);
// the six background fields in the slots of the link kernels: slots 1 / 2 (MOMX / MOMY, rotated across panel edges) carry
// DENS_hyd twice with the identity, so that no rotation touches a scalar
void aux_link_fields(fedg_ctx* c, double* f[6]) {
  f[V_DDENS] = c->dens_hyd.p; f[V_MOMX] = c->dens_hyd.p; f[V_MOMY] = c->dens_hyd.p; f[V_MOMZ] = c->pres_hyd.p; f[V_DRHOT] = c->therm_hyd.p;
  f[5] = c->pres_hyd.p;
}
void fill_halo_links(fedg_ctx* c, int buf, bool aux = false) {
  for (int f = 0; f < 6; ++f) {
    const auto& l = c->link[f];
    if (!l.src && !l.recvbuf) continue;
    LinkFields F{};
    double* own[6];
    if (aux) aux_link_fields(c, own);
    else { for (int v = 0; v < NVAR; ++v) own[v] = c->prog[buf][v].p; own[5] = c->dp[buf].p; }
    for (int v = 0; v < 6; ++v) F.d[v] = own[v];
    if (l.src) {
      double* sf[6];
      if (aux) aux_link_fields(l.src, sf);
      else { const int sb = l.src->xbuf; for (int v = 0; v < NVAR; ++v) sf[v] = l.src->prog[sb][v].p; sf[5] = l.src->dp[sb].p; }
      for (int v = 0; v < 6; ++v) F.s[v] = sf[v];
    } else {   // received by group_exchange_remote, fields in the internal variable order + DPRES
      for (int v = 0; v < 6; ++v) F.s[v] = l.recvbuf + size_t(v) * l.cnt;
    }
    halo_link_kernel<<<(l.cnt + 255) / 256, 256, 0, c->stream>>>(F, l.src ? l.d_src : nullptr, aux ? nullptr : l.d_rot, c->nint + size_t(l.off), l.cnt);
  }
}

void fill_halo(fedg_ctx* c, int buf, bool apply_bc) {
  HaloParams H{};
  for (int v = 0; v < NVAR; ++v) H.q[v] = c->prog[buf][v].p;
  H.dp = c->dp[buf].p;
  H.src = c->d_halo_src; H.vmapB = c->d_vmapB;
  H.gsqrt = c->gsqrt.p; H.g13 = c->g13.p; H.g23 = c->g23.p; H.gsqrtH = c->gsqrtH.p; H.emap2d = c->d_emap2d;
  for (int f = 0; f < 7; ++f) H.face_off[f] = c->face_off[f];
  for (int f = 0; f < 6; ++f) {
    // a face carries the BC only when its neighbour is the tile itself with the same face (bnd_Init_lc)
    bool phys = c->nbr_rank[f] == c->my_rank && c->nbr_face[f] == f;
    H.bc[f] = (apply_bc && phys) ? c->vel_bc[f] : FEDG_BND_NOSPEC;
  }
  H.Np = c->Np; H.Ne = c->Ne; H.Nfp = c->Nfp; H.np = c->np; H.Nhalo = c->Nhalo; H.terrain = c->terrain;
  launch_halo_fill(H, c->stream);
  fill_halo_links(c, buf);
}

void fill_stage_params(fedg_ctx* c, StageParams& P, int in, int out, int q0) {
  for (int v = 0; v < NVAR; ++v) {
    P.qin[v] = c->prog[in][v].p; P.qout[v] = c->prog[out][v].p; P.q0[v] = c->prog[q0][v].p;
    P.vt[v] = c->vt[v].p; P.tend_out[v] = nullptr;
  }
  P.dens_hyd = c->dens_hyd.p; P.pres_hyd = c->pres_hyd.p; P.therm_hyd = c->therm_hyd.p;
  P.rtot = c->rtot.p; P.cvtot = c->cvtot.p; P.cptot = c->cptot.p;
  P.gsqrt = c->gsqrt.p; P.g13 = c->g13.p; P.g23 = c->g23.p; P.gsqrtH = c->gsqrtH.p;
  P.dphydx = c->dphydx.p; P.dphydy = c->dphydy.p; P.coriolis = c->coriolis.p;
  P.escale = c->escale.p; P.fscale = c->fscale.p; P.vmapP = c->d_vmapP; P.emap2d = c->d_emap2d;
  P.pres_out = c->pres.p; P.dpin = c->dp[in].p; P.dpout = c->dp[out].p; P.tab = c->d_tab;
  P.c = c->c; P.Ne = c->Ne; P.Ne2D = c->Ne2D;
  P.has_cor = c->has_cor; P.has_phyd = c->has_phyd; P.do_filter = 0; P.write_pres = 0;
  P.g2d = c->g2d.p; P.OHM = c->OHM; P.is_global = c->global; P.panel = c->panel;
  for (int k = 0; k < 6; ++k) P.phyt[k] = c->phyt[k].p;
  P.has_phyt = c->has_phyt;
  P.sponge = c->has_sponge ? c->sponge.p : nullptr; P.sponge_h = c->sponge_h;
  { const char* e = getenv("FEDG_EXACT_POW"); P.exact_pow = (e && e[0] == '1') ? 1 : 0; }   // stage_p7: pow() instead of exp(e log x)
  { const char* e = getenv("FEDG_P7_L2PF"); P.l2_prefetch = (e && e[0] == '0') ? 0 : 1; }
  { const char* e = getenv("FEDG_P7_ZEXT"); P.zface_contig = (!(e && e[0] == '0') && c->zface_contig) ? 1 : 0; }   // A/B knob
}

// DPRES of prog[buf] (interior): produced by the stage kernel that wrote prog[buf]; computed here only after the
// state was replaced from the host (calc_pressure, nonhydro3d_common.F90:350-393).
void ensure_dp(fedg_ctx* c, int buf) {
  if (c->dp_valid[buf]) return;
  launch_calc_pres(c->prog[buf][V_DRHOT].p, c->pres_hyd.p, c->therm_hyd.p, c->rtot.p, c->cvtot.p, c->cptot.p, c->moist, c->c,
                   c->pres.p, c->dp[buf].p, long(c->nint), c->stream);
  c->dp_valid[buf] = true;
}

int exchange_and_stage(fedg_ctx* c, StageParams& P, int buf, bool hevi, cudaEvent_t e0 = nullptr, cudaEvent_t e1 = nullptr);
int run_numdiff(fedg_ctx* c, int buf);

void fill_vi_params(fedg_ctx* c, VIParams& V, int in, int out, int i0, int stage, double impl_fac) {
  for (int v = 0; v < NVAR; ++v) {
    V.qcur[v] = c->prog[in][v].p; V.q0[v] = c->prog[i0][v].p; V.qout[v] = c->prog[out][v].p;
    V.kim[v] = c->kim[size_t(stage) * NVAR + v].p;
  }
  V.dpout = c->dp[out].p;
  V.dens_hyd = c->dens_hyd.p; V.pres_hyd = c->pres_hyd.p; V.therm_hyd = c->therm_hyd.p; V.rhot_hyd_vi = c->rhot_hyd_vi.p;
  V.rtot = c->rtot.p; V.cvtot = c->cvtot.p; V.cptot = c->cptot.p;
  V.escale = c->escale.p; V.fscale = c->fscale.p; V.tab = c->d_tab; V.htab = &c->tab; V.scratch = c->vi_scratch.p;
  V.c = c->c; V.impl_fac = impl_fac; V.Ne = c->Ne; V.Ne2D = c->Ne2D; V.NeZ = c->NeZ;
  { const char* e = getenv("FEDG_EXACT_POW"); V.exact_pow = (e && e[0] == '1') ? 1 : 0; }
  if (c->terrain && !c->global) {
    V.gsqrt = c->gsqrt.p; V.g13 = c->g13.p; V.g23 = c->g23.p; V.gsqrtH = c->gsqrtH.p;
    V.pvu_out = c->vi_pvu.p; V.pvv_out = c->vi_pvv.p;
  }
}

// HEVI / IMEX step (driver_nonhydro3d.F90:703-763 + 769-921): per stage  cal_vi -> StoreImplicit -> halo + BC ->
// explicit tendency -> Advance (general IMEX form, scale_timeint_rk.F90:2201-2355, accumulated from var0 in the
// reference's term order); then the modal filter.
// stage pieces (so that several local meshes on one device can be advanced stage by stage, fedg_group_update)
void hevi_begin_step(fedg_ctx* c) {
  c->hs.i0 = c->cur; c->hs.in = c->cur;
  if (c->trc.couple)      // DDENS0_TRC = tint%var0 of DDENS (driver_nonhydro3d.F90:926-937): the last stage may overwrite this buffer
    cudaMemcpyAsync(c->trc.dens0.p, c->prog[c->cur][V_DDENS].p, c->nint * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
}
int hevi_stage_vi(fedg_ctx* c, int s, cudaEvent_t e0, cudaEvent_t e1) {
  const int i0 = c->hs.i0, bA = (i0 + 1) % 3, bB = (i0 + 2) % 3, in = c->hs.in;
  const int mid = (in == i0) ? bA : in;            // the column solve may update in place except on var0
  // the last combination overwrites var0 in place (every thread reads its own base value before it writes): the step ends in the
  // buffer it started in, so that one step is a fixed launch sequence (CUDA-graph replay)
  const int nxt = (s == c->rk.nstage - 1) ? i0 : ((mid == bA) ? bB : bA);
  c->hs.mid = mid; c->hs.nxt = nxt;
  VIParams V{};
  fill_vi_params(c, V, in, mid, i0, s, c->rk.aim(s, s) * c->dt);
  if (e0) CUDA_TRY(cudaEventRecord(e0, c->stream));
  launch_vi(V, c->moist, c->stream);
  if (e1) CUDA_TRY(cudaEventRecord(e1, c->stream));
  c->dp_valid[mid] = true;
  c->xbuf = mid;
  return FEDG_OK;
}
int hevi_stage_ex(fedg_ctx* c, int s) {
  StageParams P{};
  fill_stage_params(c, P, c->hs.mid, c->hs.mid, c->hs.i0);
  for (int v = 0; v < NVAR; ++v) P.tend_out[v] = c->kex[size_t(s) * NVAR + v].p;
  c->trc.stage = s;
  return exchange_and_stage(c, P, c->hs.mid, true);
}
void hevi_stage_combine(fedg_ctx* c, int s, bool with_filter = true) {
  const RKTable& t = c->rk;
  const int ns = t.nstage, i0 = c->hs.i0, nxt = c->hs.nxt;
  const double dt = c->dt;
  LinCombParams L{};
  L.n = c->nint; L.nterm = 0;
  for (int v = 0; v < NVAR; ++v) { L.base[v] = c->prog[i0][v].p; L.out[v] = c->prog[nxt][v].p; }
  for (int ss = 0; ss <= s; ++ss) {
    const double ce = (s == ns - 1) ? dt * t.b_ex[ss] : dt * t.aex(s + 1, ss);
    const double ci = (s == ns - 1) ? dt * t.b_im[ss] : dt * t.aim(s + 1, ss);
    for (int v = 0; v < NVAR; ++v) { L.k[L.nterm][v] = c->kex[size_t(ss) * NVAR + v].p; L.k[L.nterm + 1][v] = c->kim[size_t(ss) * NVAR + v].p; }
    L.coef[L.nterm] = ce; L.coef[L.nterm + 1] = ci;
    L.nterm += 2;
  }
  // the last stage's combination carries the modal filter of the step (driver_nonhydro3d.F90:940-951); with the tracer coupling the
  // unfiltered density is needed first (DDENS_TRC), so the filter runs on its own in hevi_end_step
  if (s == ns - 1 && c->modalfilter && with_filter && !c->trc.couple) launch_lincomb_filter(L, c->d_tab, c->gsqrt.p, c->terrain || c->global, c->Ne, c->np, c->stream);
  else launch_lincomb(L, c->stream);
  c->dp_valid[nxt] = false;
  c->hs.in = nxt;
}
void hevi_end_step(fedg_ctx* c) {
  c->cur = c->hs.in; c->xbuf = c->cur;
  if (c->trc.couple) {    // DDENS_TRC = DDENS after the RK loop, before the modal filter (driver_nonhydro3d.F90:926-951)
    cudaMemcpyAsync(c->trc.dens1.p, c->prog[c->cur][V_DDENS].p, c->nint * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
    if (c->modalfilter) {
      double* q[NVAR];
      for (int v = 0; v < NVAR; ++v) q[v] = c->prog[c->cur][v].p;
      launch_modal_filter5(q, c->gsqrt.p, c->terrain || c->global, c->Ne, c->np, c->stream);
      c->dp_valid[c->cur] = false;
    }
    c->trc.have_flux = true;
  }
}

// explicit (HEVE) stage pieces: buffer choice + pressure of the stage input, then exchange + fused stage kernel
void heve_stage_prepare(fedg_ctx* c, int s) {
  const int ns = c->rk.nstage, i0 = c->hs.i0, in = c->hs.in;
  int out;
  if (s == ns - 1) out = (ns == 1) ? (i0 + 1) % 3 : i0;
  else { out = (in + 1) % 3; if (out == i0) out = (out + 1) % 3; }
  c->hs.nxt = out;
  ensure_dp(c, in);
  c->xbuf = in;
}
int heve_stage(fedg_ctx* c, int s, cudaEvent_t e0, cudaEvent_t e1) {
  const int ns = c->rk.nstage, in = c->hs.in, out = c->hs.nxt;
  StageParams P{};
  fill_stage_params(c, P, in, out, c->hs.i0);
  P.rk = c->stages[s];
  if (s == ns - 1) { P.do_filter = c->modalfilter && !c->trc.couple; P.write_pres = 1; }
  c->trc.stage = s;
  int rc = exchange_and_stage(c, P, in, false, e0, e1);
  if (rc) return rc;
  c->dp_valid[out] = true;
  c->hs.in = out;
  return FEDG_OK;
}

int run_steps_hevi(fedg_ctx* c, int nsteps, size_t& iev, long& launches) {
  const int ns = c->rk.nstage;
  for (int step = 0; step < nsteps; ++step) {
    hevi_begin_step(c);
    for (int s = 0; s < ns; ++s) {
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      if (c->profile) { e0 = c->ev[iev++]; e1 = c->ev[iev++]; }
      { int rc = hevi_stage_vi(c, s, e0, e1); if (rc) return rc; }
      { int rc = hevi_stage_ex(c, s); if (rc) return rc; }
      hevi_stage_combine(c, s);
      launches += 4;
    }
    hevi_end_step(c);
    if (c->nd.on && c->nd.in_update) { int rc = run_numdiff(c, c->cur); if (rc) return rc; }
  }
  return FEDG_OK;
}

// MeshFieldComm_Exchange + boundary condition + stage kernel.  With remote neighbours: pack and ship the tile faces on
// the communication stream, process the interior elements meanwhile, then the tile-boundary elements once the halo
// has arrived (HIDE_MPI_COMM_FLAG path of the reference, driver_nonhydro3d.F90:859-895).
int exchange_and_stage(fedg_ctx* c, StageParams& P, int buf, bool hevi, cudaEvent_t e0, cudaEvent_t e1) {
  fill_halo(c, buf, true);    // faces whose neighbour is on this rank + physical boundaries
  if (c->trc.couple) {        // atm_dyn_dgm_trcadvect3d_save_massflux on the stage state (driver_nonhydro3d.F90:900-917)
    const int st = c->trc.stage;
    const double w_h = c->rk.b_ex[st], w_v = c->rk.imex ? c->rk.b_im[st] : c->rk.b_ex[st];
    const double* q[NVAR];
    for (int v = 0; v < NVAR; ++v) q[v] = c->prog[buf][v].p;
    double* mf[3] = {c->trc.mflx[0].p, c->trc.mflx[1].p, c->trc.mflx[2].p};
    CUDA_TRY(launch_trc_save_massflux(q, c->dens_hyd.p, c->pres_hyd.p, c->dp[buf].p, c->d_vmapP, mf, c->trc.alphM_t.p, c->trc.alphP_t.p, w_h, w_v,
                                      c->c.gamm, hevi, st == 0, c->Np, c->Nfp, c->NfpTot, c->np, c->Ne, c->stream));
  }
  if (e0) cudaEventRecord(e0, c->stream);   // the timed region brackets the stage kernel(s) (+ the exchange when there is one)
  struct Rec { cudaEvent_t e; cudaStream_t s; ~Rec() { if (e) cudaEventRecord(e, s); } } rec{e1, c->stream};
  if (!c->comm.active || c->comm.nremote == 0) {
    launch_stage(P, c->np, c->terrain, c->moist, hevi, c->stream);
    return FEDG_OK;
  }
  double* q[NVAR];
  for (int v = 0; v < NVAR; ++v) q[v] = c->prog[buf][v].p;
  std::string err;
  int rc = comm_exchange_start(c->comm, q, c->dp[buf].p, c->d_vmapB, c->nint, c->stream, err);
  if (rc) return fail(rc, err);
  if (c->n_inner > 0) {
    P.elem_list = c->d_elem_inner; P.nelem = c->n_inner;
    launch_stage(P, c->np, c->terrain, c->moist, hevi, c->stream);
  }
  comm_exchange_wait(c->comm, c->stream);
  if (c->n_bnd > 0) {
    P.elem_list = c->d_elem_bnd; P.nelem = c->n_bnd;
    launch_stage(P, c->np, c->terrain, c->moist, hevi, c->stream);
  }
  P.elem_list = nullptr; P.nelem = 0;
  return FEDG_OK;
}

// halo of auxiliary work fields: same-tile faces by gather, remote faces through the six-slot NCCL exchange
int exchange_work_fields(fedg_ctx* c, double* const f[], int nf) {
  for (int k = 0; k < nf; ++k)
    if (c->Nhalo > 0) aux_halo_kernel<<<(c->Nhalo + 255) / 256, 256, 0, c->stream>>>(f[k], c->d_halo_src, int(c->nint), c->Nhalo);
  if (c->comm.active && c->comm.nremote > 0) {
    double* q[NVAR];
    for (int v = 0; v < NVAR; ++v) q[v] = f[v % nf];
    std::string err;
    int rc = comm_exchange_start(c->comm, q, f[(nf - 1)], c->d_vmapB, c->nint, c->stream, err);
    if (rc) return fail(rc, err);
    comm_exchange_wait(c->comm, c->stream);
  }
  return FEDG_OK;
}

// AtmDyn_Nonhydro3D_Numdiff%Apply on prog[buf] (numdiff.F90:214-376)
int run_numdiff(fedg_ctx* c, int buf) {
  for (auto& b : c->nd_g) if (b.n < c->nall) CUDA_TRY(b.alloc(c->nall));
  if (c->nd.lap_num > 1) for (auto& b : c->nd_lap) if (b.n < c->nall) CUDA_TRY(b.alloc(c->nall));
  ensure_tables(c);
  // PROG_VARS%MeshFieldComm_Exchange (no boundary condition: the numdiff rules are applied at the face nodes)
  ensure_dp(c, buf);
  fill_halo(c, buf, false);
  if (c->comm.active && c->comm.nremote > 0) {
    double* q[NVAR];
    for (int v = 0; v < NVAR; ++v) q[v] = c->prog[buf][v].p;
    std::string err;
    int rc = comm_exchange_start(c->comm, q, c->dp[buf].p, c->d_vmapB, c->nint, c->stream, err);
    if (rc) return fail(rc, err);
    comm_exchange_wait(c->comm, c->stream);
  }
  NumdiffParams P{};
  P.ddens = c->prog[buf][V_DDENS].p; P.dens_hyd = c->dens_hyd.p; P.tab = c->d_tab;
  P.escale = c->escale.p; P.fscale = c->fscale.p; P.vmapP = c->d_vmapP;
  for (int f = 0; f < 6; ++f) {
    const bool phys = c->nbr_rank[f] == c->my_rank && c->nbr_face[f] == f;
    P.vel_bc[f] = phys ? c->vel_bc[f] : 0; P.therm_bc[f] = phys ? c->nd.therm_bc[f] : 0;
  }
  for (int f = 0; f < 7; ++f) P.face_off[f] = c->face_off[f];
  P.Np = c->Np; P.Nfp = c->Nfp; P.NfpTot = c->NfpTot; P.np = c->np; P.Ne = c->Ne; P.nint = c->nint; P.dt = c->dt;
  const double nd_sign = ((c->nd.lap_num + 1) % 2 == 0) ? 1.0 : -1.0;
  const int order[NVAR] = {V_DRHOT, V_MOMZ, V_MOMX, V_MOMY, V_DDENS};
  // p = 7 and one Laplacian (the shipped configuration): the five variables of a half-step go through ONE launch each -- gradients of
  // all five, exchange of the fifteen gradient fields, then the divergence + update of the four density-weighted variables, and DDENS
  // last in its own launch (the others read the neighbours' un-diffused DDENS for their weights, as in the reference's order).
  // FEDG_ND_KERNEL=1 / 2 select the node-per-thread kernel / the tensor-core kernel one variable at a time (A/B runs).
  {
    const char* e = getenv("FEDG_ND_KERNEL");
    if (c->np == 8 && c->nd.lap_num == 1 && !(e && (e[0] == '1' || e[0] == '2'))) {
      for (auto& b : c->nd_g5) if (b.n < c->nall) CUDA_TRY(b.alloc(c->nall));
      NumdiffMulti M{};
      M.P = P; M.P.bc_on_v = 1; M.nvar = NVAR;
      for (int iv = 0; iv < NVAR; ++iv) {
        const int v = order[iv];
        NdVar& V = M.v[iv];
        V.in0 = V.in1 = c->prog[buf][v].p; V.in2 = nullptr;
        V.out0 = c->nd_g5[3 * iv].p; V.out1 = c->nd_g5[3 * iv + 1].p; V.out2 = c->nd_g5[3 * iv + 2].p;
        V.var = nullptr; V.varid = v; V.dens_flag = (v != V_DDENS) ? 1 : 0;
      }
      launch_numdiff_multi(0, M, c->stream);
      for (int iv = 0; iv < NVAR; ++iv) {
        double* g3[3] = {c->nd_g5[3 * iv].p, c->nd_g5[3 * iv + 1].p, c->nd_g5[3 * iv + 2].p};
        int rc = exchange_work_fields(c, g3, 3);
        if (rc) return rc;
      }
      M.P.coef_h = nd_sign * c->nd.coef_h; M.P.coef_v = nd_sign * c->nd.coef_v;
      for (int iv = 0; iv < NVAR; ++iv) {
        NdVar& V = M.v[iv];
        V.in0 = c->nd_g5[3 * iv].p; V.in1 = c->nd_g5[3 * iv + 1].p; V.in2 = c->nd_g5[3 * iv + 2].p;
        V.out0 = V.out1 = V.out2 = nullptr; V.var = c->prog[buf][order[iv]].p;
      }
      M.nvar = NVAR - 1;                        // DRHOT, MOMZ, MOMX, MOMY
      launch_numdiff_multi(2, M, c->stream);
      M.v[0] = M.v[NVAR - 1]; M.nvar = 1;       // DDENS
      launch_numdiff_multi(2, M, c->stream);
      c->dp_valid[buf] = false;
      return FEDG_OK;
    }
  }
  double* g[3] = {c->nd_g[0].p, c->nd_g[1].p, c->nd_g[2].p};
  for (int iv = 0; iv < NVAR; ++iv) {
    const int v = order[iv];
    const bool dens_weight = v != V_DDENS;
    double* var = c->prog[buf][v].p;
    P.varid = v;
    P.in0 = var; P.in1 = var; P.in2 = nullptr; P.out0 = g[0]; P.out1 = g[1]; P.out2 = g[2];
    P.dens_flag = dens_weight; P.bc_on_v = 1;
    launch_numdiff(0, P, c->stream);
    { int rc = exchange_work_fields(c, g, 3); if (rc) return rc; }
    for (int it = 1; it <= c->nd.lap_num - 1; ++it) {
      double* lp[2] = {c->nd_lap[0].p, c->nd_lap[1].p};
      P.in0 = g[0]; P.in1 = g[1]; P.in2 = g[2]; P.out0 = lp[0]; P.out1 = lp[1]; P.out2 = nullptr; P.dens_flag = 0;
      launch_numdiff(1, P, c->stream);
      { int rc = exchange_work_fields(c, lp, 2); if (rc) return rc; }
      P.in0 = lp[0]; P.in1 = lp[1]; P.in2 = nullptr; P.out0 = g[0]; P.out1 = g[1]; P.out2 = g[2]; P.dens_flag = 0; P.bc_on_v = 0;
      launch_numdiff(0, P, c->stream);
      { int rc = exchange_work_fields(c, g, 3); if (rc) return rc; }
    }
    P.in0 = g[0]; P.in1 = g[1]; P.in2 = g[2]; P.var = var; P.dens_flag = dens_weight;
    P.coef_h = nd_sign * c->nd.coef_h; P.coef_v = nd_sign * c->nd.coef_v;
    launch_numdiff(2, P, c->stream);
  }
  c->dp_valid[buf] = false;
  return FEDG_OK;
}

// wait = false: the steps are only enqueued on the context's stream (no events, no host synchronisation): the pipelined host update
int run_steps(fedg_ctx* c, int nsteps, bool wait = true) {
  if (!c->dyn_ready || !c->aux_ready) return fail(FEDG_ERR_STATE, "fedg_dyn_init and fedg_set_aux must be called before the update");
  ensure_tables(c);
  const int ns = c->rk.nstage;
  const bool saved_profile = c->profile;
  if (!wait) c->profile = false;
  struct RestoreProfile { fedg_ctx* c; bool v; ~RestoreProfile() { c->profile = v; } } restore_profile{c, saved_profile};
  const size_t need_ev = 2 + (c->profile ? size_t(2) * ns * nsteps : 0);
  while (c->ev.size() < need_ev) { cudaEvent_t e; CUDA_TRY(cudaEventCreate(&e)); c->ev.push_back(e); }
  CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
  size_t iev = 2;
  long launches = 0;
  c->last_ms_vi = 0;
  if (c->hevi) {
    int rc = run_steps_hevi(c, nsteps, iev, launches);
    if (rc) return rc;
    nsteps = 0;
  }
  for (int step = 0; step < nsteps; ++step) {
    hevi_begin_step(c);
    for (int s = 0; s < ns; ++s) {
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      if (c->profile) { e0 = c->ev[iev++]; e1 = c->ev[iev++]; }
      heve_stage_prepare(c, s);
      { int rc = heve_stage(c, s, e0, e1); if (rc) return rc; }
      launches += 2;
    }
    hevi_end_step(c);
    if (c->nd.on && c->nd.in_update) { int rc = run_numdiff(c, c->cur); if (rc) return rc; }
  }
  CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
  c->last_launches = launches;
  if (!wait) return FEDG_OK;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
  c->last_ms_total = ms; c->last_ms_stage = 0;
  if (c->profile)
    for (size_t k = 2; k + 1 < iev; k += 2) { float m2 = 0; CUDA_TRY(cudaEventElapsedTime(&m2, c->ev[k], c->ev[k + 1])); c->last_ms_stage += m2; }
  return FEDG_OK;
}
}  // namespace

namespace {
__global__ void fill_kernel(double* p, double v, size_t n) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
}  // namespace

// The slots of the six-field exchange below depend on `moist`: the ranks must agree on it (a dry tile next to a moist one would
// otherwise receive Rtot / CVtot / CPtot in the slots where it expects the hydrostatic fields).  One moist rank makes every rank take
// the moist path; a rank whose own thermodynamic fields equal the dry constants fills its arrays with them.
static int agree_moist(fedg_ctx* c) {
  if (!c->comm.active) return FEDG_OK;
  double flag = c->moist ? 1.0 : 0.0;
  CUDA_TRY(cudaMemcpyAsync(c->mon.p + 7, &flag, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  std::string err;
  int rc = comm_allreduce_max(c->comm, c->mon.p + 7, 1, c->stream, err);
  if (rc) return fail(rc, err);
  CUDA_TRY(cudaMemcpyAsync(&flag, c->mon.p + 7, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (flag != 0.0 && !c->moist) {
    const double v[3] = {c->c.Rdry, c->c.CVdry, c->c.CPdry};
    DevBuf* b[3] = {&c->rtot, &c->cvtot, &c->cptot};
    for (int k = 0; k < 3; ++k) {
      if (b[k]->n < c->nall) CUDA_TRY(b[k]->alloc(c->nall));
      fill_kernel<<<unsigned((c->nall + 255) / 256), 256, 0, c->stream>>>(b[k]->p, v[k], c->nall);
    }
    c->moist = true;
  }
  return FEDG_OK;
}

static int exchange_aux_remote(fedg_ctx* c) {
  if (!c->comm.active) return FEDG_OK;
  { int rc = agree_moist(c); if (rc) return rc; }
  if (c->comm.nremote == 0) return FEDG_OK;
  // the six-field exchange ships any six arrays: send the background fields in its slots
  double* q[NVAR] = {c->dens_hyd.p, c->pres_hyd.p, c->therm_hyd.p, c->moist ? c->rtot.p : c->dens_hyd.p, c->moist ? c->cvtot.p : c->pres_hyd.p};
  double* sixth = c->moist ? c->cptot.p : c->therm_hyd.p;
  std::string err;
  int rc = comm_exchange_start(c->comm, q, sixth, c->d_vmapB, c->nint, c->stream, err);
  if (rc) return fail(rc, err);
  comm_exchange_wait(c->comm, c->stream);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return FEDG_OK;
}

static int fill_aux_halo(fedg_ctx* c, double* field) {
  if (c->Nhalo > 0) aux_halo_kernel<<<(c->Nhalo + 255) / 256, 256, 0, c->stream>>>(field, c->d_halo_src, int(c->nint), c->Nhalo);
  CUDA_TRY(cudaGetLastError());
  return FEDG_OK;
}

extern "C" {

int fedg_dyn_update(fedg_ctx* c, int nsteps) {
  if (!c || nsteps < 0) return fail(FEDG_ERR_ARG, "bad argument");
  return run_steps(c, nsteps);
}

static int host_pipe_init(fedg_ctx* c) {
  auto& hp = c->hp;
  if (hp.ready) return FEDG_OK;
  CUDA_TRY(cudaStreamCreateWithFlags(&hp.h2d, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&hp.d2h, cudaStreamNonBlocking));
  for (int k = 0; k < fedg_ctx::NSLOT; ++k) {
    for (auto& b : hp.in[k]) CUDA_TRY(b.alloc(c->nint));
    for (auto& b : hp.out[k]) CUDA_TRY(b.alloc(c->nint));
    for (cudaEvent_t* e : {&hp.in_ready[k], &hp.in_free[k], &hp.out_ready[k], &hp.out_done[k]})
      CUDA_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  hp.ready = true;
  return FEDG_OK;
}

// One pass of the pipeline for slot `slot` (0 .. 2).  Three streams: H2D copies -> (event) -> compute stream: staging -> state, the
// steps, state -> staging -> (event) -> D2H copies.  PCIe is full duplex: while one slot is being downloaded, the next is computed and
// a third uploaded, so a caller that rotates three sets of host arrays is bound by one direction of the link (7.2 ms for 335 MB each
// way, measured with both directions busy), not by the sum of both plus the step; with two sets the chain of one slot (upload + step +
// download = 16 ms) bounds the rate at 8.9 ms per call.  Only the (Np, Ne) interior travels: the halo slots [Ne+1:NeA] are rebuilt on the device by the
// exchange of every stage before anything reads them.
int fedg_dyn_update_host_async(fedg_ctx* c, const double* DDENS, const double* MOMX, const double* MOMY, const double* MOMZ,
                               const double* DRHOT, double* DDENS_out, double* MOMX_out, double* MOMY_out, double* MOMZ_out,
                               double* DRHOT_out, int nsteps, int slot) {
  if (!c || !DDENS || !MOMX || !MOMY || !MOMZ || !DRHOT || !DDENS_out || !MOMX_out || !MOMY_out || !MOMZ_out || !DRHOT_out || nsteps < 0)
    return fail(FEDG_ERR_ARG, "null argument");
  if (slot < 0 || slot >= fedg_ctx::NSLOT) return fail(FEDG_ERR_ARG, "slot must be 0, 1 or 2");
  if (!c->dyn_ready || !c->aux_ready) return fail(FEDG_ERR_STATE, "fedg_dyn_init and fedg_set_aux must be called before the update");
  { int rc = host_pipe_init(c); if (rc) return rc; }
  auto& hp = c->hp;
  if (hp.pending[slot]) return fail(FEDG_ERR_STATE, "slot still in flight: call fedg_dyn_update_host_wait on it first");
  const double* hin[NVAR] = {DDENS, MOMX, MOMY, MOMZ, DRHOT};
  double* hout[NVAR] = {DDENS_out, MOMX_out, MOMY_out, MOMZ_out, DRHOT_out};
  const size_t bytes = c->nint * sizeof(double);
  // upload: the staging buffer of this slot is free once the compute stream has copied its previous contents into the state
  CUDA_TRY(cudaStreamWaitEvent(hp.h2d, hp.in_free[slot], 0));
  for (int v = 0; v < NVAR; ++v) CUDA_TRY(cudaMemcpyAsync(hp.in[slot][v].p, hin[v], bytes, cudaMemcpyHostToDevice, hp.h2d));
  CUDA_TRY(cudaEventRecord(hp.in_ready[slot], hp.h2d));
  // compute
  CUDA_TRY(cudaStreamWaitEvent(c->stream, hp.in_ready[slot], 0));
  for (int v = 0; v < NVAR; ++v)
    CUDA_TRY(cudaMemcpyAsync(c->prog[c->cur][v].p, hp.in[slot][v].p, bytes, cudaMemcpyDeviceToDevice, c->stream));
  CUDA_TRY(cudaEventRecord(hp.in_free[slot], c->stream));
  c->dp_valid[c->cur] = false;
  c->xbuf = c->cur;
  { int rc = run_steps(c, nsteps, false); if (rc) return rc; }
  CUDA_TRY(cudaStreamWaitEvent(c->stream, hp.out_done[slot], 0));     // the previous download of this slot has left the staging buffer
  for (int v = 0; v < NVAR; ++v)
    CUDA_TRY(cudaMemcpyAsync(hp.out[slot][v].p, c->prog[c->cur][v].p, bytes, cudaMemcpyDeviceToDevice, c->stream));
  CUDA_TRY(cudaEventRecord(hp.out_ready[slot], c->stream));
  // download
  CUDA_TRY(cudaStreamWaitEvent(hp.d2h, hp.out_ready[slot], 0));
  for (int v = 0; v < NVAR; ++v) CUDA_TRY(cudaMemcpyAsync(hout[v], hp.out[slot][v].p, bytes, cudaMemcpyDeviceToHost, hp.d2h));
  CUDA_TRY(cudaEventRecord(hp.out_done[slot], hp.d2h));
  hp.pending[slot] = true;
  return FEDG_OK;
}

int fedg_dyn_update_host_wait(fedg_ctx* c, int slot) {
  if (!c || slot < 0 || slot >= fedg_ctx::NSLOT) return fail(FEDG_ERR_ARG, "bad argument");
  auto& hp = c->hp;
  if (!hp.ready || !hp.pending[slot]) return FEDG_OK;
  CUDA_TRY(cudaEventSynchronize(hp.out_done[slot]));
  hp.pending[slot] = false;
  CUDA_TRY(cudaGetLastError());
  return FEDG_OK;
}

int fedg_dyn_update_host(fedg_ctx* c, double* DDENS, double* MOMX, double* MOMY, double* MOMZ, double* DRHOT, int nsteps) {
  if (!c || !DDENS || !MOMX || !MOMY || !MOMZ || !DRHOT) return fail(FEDG_ERR_ARG, "null argument");
  for (int k = 0; k < fedg_ctx::NSLOT; ++k) { int rc = fedg_dyn_update_host_wait(c, k); if (rc) return rc; }
  // One blocking call is a strict chain upload -> steps -> download (nothing to overlap): the copies go straight into / out of the
  // state buffers on the compute stream, no staging.  Only the (Np, Ne) interior travels (20 % fewer PCIe bytes at 32x32x16).
  double* h[NVAR] = {DDENS, MOMX, MOMY, MOMZ, DRHOT};
  for (int v = 0; v < NVAR; ++v)
    CUDA_TRY(cudaMemcpyAsync(c->prog[c->cur][v].p, h[v], c->nint * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  c->dp_valid[c->cur] = false;
  c->xbuf = c->cur;
  int rc;
  if ((rc = run_steps(c, nsteps))) return rc;
  for (int v = 0; v < NVAR; ++v)
    CUDA_TRY(cudaMemcpyAsync(h[v], c->prog[c->cur][v].p, c->nint * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return FEDG_OK;
}

int fedg_exchange_halo(fedg_ctx* c, int apply_bc) {
  if (!c) return fail(FEDG_ERR_ARG, "null argument");
  fill_halo(c, c->cur, apply_bc != 0);
  if (c->comm.active && c->comm.nremote > 0) {
    ensure_dp(c, c->cur);
    double* q[NVAR];
    for (int v = 0; v < NVAR; ++v) q[v] = c->prog[c->cur][v].p;
    std::string err;
    int rc = comm_exchange_start(c->comm, q, c->dp[c->cur].p, c->d_vmapB, c->nint, c->stream, err);
    if (rc) return fail(rc, err);
    comm_exchange_wait(c->comm, c->stream);
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  return FEDG_OK;
}

int fedg_cal_tend_ex(fedg_ctx* c, double* DENS_dt, double* MOMX_dt, double* MOMY_dt, double* MOMZ_dt, double* RHOT_dt) {
  if (!c || !DENS_dt || !MOMX_dt || !MOMY_dt || !MOMZ_dt || !RHOT_dt) return fail(FEDG_ERR_ARG, "null argument");
  if (!c->dyn_ready || !c->aux_ready) return fail(FEDG_ERR_STATE, "fedg_dyn_init and fedg_set_aux must be called first");
  ensure_tables(c);
  for (auto& b : c->tendbuf) if (b.n < c->nint) CUDA_TRY(b.alloc(c->nint));
  ensure_dp(c, c->cur);
  StageParams P{};
  fill_stage_params(c, P, c->cur, c->cur, c->cur);
  for (int v = 0; v < NVAR; ++v) P.tend_out[v] = c->tendbuf[v].p;
  { int rc = exchange_and_stage(c, P, c->cur, c->hevi); if (rc) return rc; }
  double* h[NVAR] = {DENS_dt, MOMX_dt, MOMY_dt, MOMZ_dt, RHOT_dt};
  for (int v = 0; v < NVAR; ++v)
    CUDA_TRY(cudaMemcpyAsync(h[v], c->tendbuf[v].p, c->nint * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  return FEDG_OK;
}

int fedg_cal_vi(fedg_ctx* c, double impl_fac, const double* DDENS0, const double* MOMX0, const double* MOMY0, const double* MOMZ0,
                const double* DRHOT0, double* DENS_dt, double* MOMX_dt, double* MOMY_dt, double* MOMZ_dt, double* RHOT_dt) {
  if (!c || !DDENS0 || !MOMX0 || !MOMY0 || !MOMZ0 || !DRHOT0 || !DENS_dt || !MOMX_dt || !MOMY_dt || !MOMZ_dt || !RHOT_dt)
    return fail(FEDG_ERR_ARG, "null argument");
  if (!c->dyn_ready || !c->aux_ready || !c->hevi) return fail(FEDG_ERR_STATE, "fedg_dyn_init(NONHYDRO3D_HEVI) and fedg_set_aux must be called first");
  ensure_tables(c);
  const int cur = c->cur, b0 = (cur + 1) % 3, b1 = (cur + 2) % 3;
  const double* h0[NVAR] = {DDENS0, MOMX0, MOMY0, MOMZ0, DRHOT0};
  for (int v = 0; v < NVAR; ++v)
    CUDA_TRY(cudaMemcpyAsync(c->prog[b0][v].p, h0[v], c->nint * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  VIParams V{};
  fill_vi_params(c, V, cur, b1, b0, 0, impl_fac);
  launch_vi(V, c->moist, c->stream);
  c->dp_valid[b0] = c->dp_valid[b1] = false;
  double* h[NVAR] = {DENS_dt, MOMX_dt, MOMY_dt, MOMZ_dt, RHOT_dt};
  for (int v = 0; v < NVAR; ++v)
    CUDA_TRY(cudaMemcpyAsync(h[v], c->kim[v].p, c->nint * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  return FEDG_OK;
}

int fedg_get_pres(fedg_ctx* c, double* PRES, double* DPRES) {
  if (!c || !PRES || !DPRES) return fail(FEDG_ERR_ARG, "null argument");
  if (!c->aux_ready) return fail(FEDG_ERR_STATE, "fedg_set_aux must be called first");
  launch_calc_pres(c->prog[c->cur][V_DRHOT].p, c->pres_hyd.p, c->therm_hyd.p, c->rtot.p, c->cvtot.p, c->cptot.p, c->moist, c->c,
                   c->pres.p, c->dp[c->cur].p, long(c->nint), c->stream);
  c->dp_valid[c->cur] = true;
  CUDA_TRY(cudaMemcpyAsync(PRES, c->pres.p, c->nint * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(DPRES, c->dp[c->cur].p, c->nint * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  return FEDG_OK;
}

int fedg_monitor(fedg_ctx* c, double* out) {
  if (!c || !out) return fail(FEDG_ERR_ARG, "null argument");
  if (!c->aux_ready) return fail(FEDG_ERR_STATE, "fedg_set_aux must be called first");
  launch_calc_pres(c->prog[c->cur][V_DRHOT].p, c->pres_hyd.p, c->therm_hyd.p, c->rtot.p, c->cvtot.p, c->cptot.p, c->moist, c->c,
                   c->pres.p, c->dp[c->cur].p, long(c->nint), c->stream);
  c->dp_valid[c->cur] = true;
  const double* q[NVAR];
  for (int v = 0; v < NVAR; ++v) q[v] = c->prog[c->cur][v].p;
  launch_monitor(q, c->dens_hyd.p, c->pres.p, c->rtot.p, c->moist, c->w3.p, c->Jac.p, c->gsqrt.p, c->terrain || c->global, c->zlev.p, c->c,
                 c->Np, c->Ne, c->mon.p, c->stream);
  CUDA_TRY(cudaMemcpyAsync(out, c->mon.p, 5 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  return FEDG_OK;
}

int fedg_elem_op(fedg_ctx* c, const char* name, const double* in, double* out, int nelem) {
  if (!c || !name || !in || !out || nelem <= 0) return fail(FEDG_ERR_ARG, "bad argument");
  static const char* names[6] = {"Dx", "Dy", "Dz", "Lift", "VFilterPM1", "ModalFilter"};
  int op = -1;
  for (int k = 0; k < 6; ++k) if (std::strcmp(name, names[k]) == 0) op = k;
  if (op < 0) return fail(FEDG_ERR_ARG, std::string("unknown element operation ") + name);
  ensure_tables(c);
  const size_t nin = size_t(op == 3 ? c->NfpTot : c->Np) * nelem, nout = size_t(c->Np) * nelem;
  struct Bufs { DevBuf a, b; ~Bufs() { a.release(); b.release(); } } d;     // released on every return path
  CUDA_TRY(d.a.alloc(nin)); CUDA_TRY(d.b.alloc(nout));
  CUDA_TRY(cudaMemcpyAsync(d.a.p, in, nin * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  launch_elem_op(op, d.a.p, d.b.p, nelem, c->np, c->stream);
  CUDA_TRY(cudaMemcpyAsync(out, d.b.p, nout * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  return FEDG_OK;
}

int fedg_last_timing(fedg_ctx* c, double* ms_total, double* ms_stage_kernels, long* n_launches) {
  if (!c) return fail(FEDG_ERR_ARG, "null argument");
  if (ms_total) *ms_total = c->last_ms_total;
  if (ms_stage_kernels) *ms_stage_kernels = c->last_ms_stage;
  if (n_launches) *n_launches = c->last_launches;
  return FEDG_OK;
}

int fedg_comm_unique_id(void* id128) {
  if (!id128) return fail(FEDG_ERR_ARG, "null argument");
  std::string err;
  int rc = comm_unique_id(id128, err);
  return rc ? fail(rc, err) : FEDG_OK;
}

int fedg_comm_init(fedg_ctx* c, const void* id128, int rank, int nranks) {
  if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(FEDG_ERR_ARG, "bad argument");
  if (rank != c->my_rank) return fail(FEDG_ERR_ARG, "rank differs from fedg_mesh_desc.my_rank");
  if (c->comm.active) return fail(FEDG_ERR_STATE, "communicator already initialised");
  std::string err;
  int rc = comm_init(c->comm, id128, rank, nranks, c->nbr_rank, c->nbr_face, c->face_off, err);
  if (rc) return fail(rc, err);
  // elements touching a face whose neighbour tile lives on another rank
  const int fsz[6] = {c->NeX * c->NeZ, c->NeY * c->NeZ, c->NeX * c->NeZ, c->NeY * c->NeZ, c->NeX * c->NeY, c->NeX * c->NeY};
  (void)fsz;
  std::vector<char> is_bnd(c->Ne, 0);
  for (int ke = 0; ke < c->Ne; ++ke) {
    const int ix = ke % c->NeX, iy = (ke / c->NeX) % c->NeY;
    if ((iy == 0 && c->nbr_rank[0] != rank) || (ix == c->NeX - 1 && c->nbr_rank[1] != rank) || (iy == c->NeY - 1 && c->nbr_rank[2] != rank) ||
        (ix == 0 && c->nbr_rank[3] != rank))
      is_bnd[ke] = 1;
  }
  if (c->nbr_rank[4] != rank || c->nbr_rank[5] != rank) return fail(FEDG_ERR_UNSUPPORTED, "vertical tile decomposition is not supported (NprcZ = 1 in the reference)");
  std::vector<int> inner, bnd;
  for (int ke = 0; ke < c->Ne; ++ke) (is_bnd[ke] ? bnd : inner).push_back(ke);
  c->n_inner = int(inner.size()); c->n_bnd = int(bnd.size());
  CUDA_TRY(cudaMalloc(&c->d_elem_inner, std::max<size_t>(inner.size(), 1) * sizeof(int)));
  CUDA_TRY(cudaMalloc(&c->d_elem_bnd, std::max<size_t>(bnd.size(), 1) * sizeof(int)));
  CUDA_TRY(cudaMemcpy(c->d_elem_inner, inner.data(), inner.size() * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(c->d_elem_bnd, bnd.data(), bnd.size() * sizeof(int), cudaMemcpyHostToDevice));
  if (c->aux_ready) { int rc2 = exchange_aux_remote(c); if (rc2) return rc2; }
  return FEDG_OK;
}

}  // extern "C"


// ---- stage-level seams: the reference driver's own stage loop over device-resident buffers --------------------------------
namespace {
// rk_advance_low_storage2D / rk_advance_general2D with one tendency buffer (scale_timeint_rk.F90:1182-1266, 2201-2355): the update
// the fused stage kernel applies, as a stand-alone pass for a driver that keeps its own stage loop
struct AdvParams { const double* q[NVAR]; const double* q0[NVAR]; double* vt[NVAR]; const double* k[NVAR]; double* out[NVAR]; RKStage rk; size_t n; };
__global__ void rk_advance_kernel(const __grid_constant__ AdvParams A) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
#pragma unroll
  for (int v = 0; v < NVAR; ++v) {
    const double q = A.q[v][i], tend = A.k[v][i];
    double base = 0.0;
    if (A.rk.use_q0) base = A.rk.c_q0 * A.q0[v][i];
    if (A.rk.add_vt) base = A.vt[v][i];
    const double r = base + A.rk.c_q * q + A.rk.c_k * tend;
    if (A.rk.vt_update) {
      const double vb = A.rk.vt_init ? A.rk.vt_init_q * q : A.vt[v][i];
      A.vt[v][i] = vb + A.rk.vt_q * q + A.rk.vt_k * tend;
    }
    A.out[v][i] = r;
  }
}
int seam_ready(fedg_ctx* c) {
  if (!c) return fail(FEDG_ERR_ARG, "null argument");
  if (!c->dyn_ready || !c->aux_ready) return fail(FEDG_ERR_STATE, "fedg_dyn_init and fedg_set_aux must be called first");
  ensure_tables(c);
  return FEDG_OK;
}
int seam_sync(fedg_ctx* c) {
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  return FEDG_OK;
}
}  // namespace

extern "C" {

int fedg_rk_store_var0(fedg_ctx* c) {
  { int rc = seam_ready(c); if (rc) return rc; }
  hevi_begin_step(c);
  if (!c->hevi && c->kex.size() < size_t(NVAR)) {
    c->kex.resize(NVAR);
    for (auto& b : c->kex) if (b.n < c->nint) CUDA_TRY(b.alloc(c->nint));
  }
  c->seam.in_step = true; c->seam.halo_inflight = c->seam.halo_ready = c->seam.vi_done = false;
  return FEDG_OK;
}

int fedg_halo_start(fedg_ctx* c) {
  { int rc = seam_ready(c); if (rc) return rc; }
  const int buf = c->cur;
  ensure_dp(c, buf);
  fill_halo(c, buf, true);
  if (c->comm.active && c->comm.nremote > 0) {
    double* q[NVAR];
    for (int v = 0; v < NVAR; ++v) q[v] = c->prog[buf][v].p;
    std::string err;
    int rc = comm_exchange_start(c->comm, q, c->dp[buf].p, c->d_vmapB, c->nint, c->stream, err);
    if (rc) return fail(rc, err);
    c->seam.halo_inflight = true;
  }
  c->seam.halo_buf = buf; c->seam.halo_ready = false;
  return FEDG_OK;
}

int fedg_halo_wait(fedg_ctx* c) {
  if (!c) return fail(FEDG_ERR_ARG, "null argument");
  if (c->seam.halo_inflight) { comm_exchange_wait(c->comm, c->stream); c->seam.halo_inflight = false; }
  c->seam.halo_ready = true;
  return FEDG_OK;
}

int fedg_cal_vi_dev(fedg_ctx* c, int stage) {
  { int rc = seam_ready(c); if (rc) return rc; }
  if (!c->hevi) return fail(FEDG_ERR_STATE, "cal_vi belongs to the HEVI equation sets");
  if (!c->seam.in_step) return fail(FEDG_ERR_STATE, "fedg_rk_store_var0 opens a step");
  if (stage < 1 || stage > c->rk.nstage) return fail(FEDG_ERR_ARG, "stage out of range (1-based)");
  c->hs.in = c->cur;
  { int rc = hevi_stage_vi(c, stage - 1, nullptr, nullptr); if (rc) return rc; }
  c->seam.vi_done = true;
  return seam_sync(c);
}

int fedg_rk_store_implicit(fedg_ctx* c, int stage) {
  if (!c) return fail(FEDG_ERR_ARG, "null argument");
  if (!c->seam.vi_done) return fail(FEDG_ERR_STATE, "fedg_cal_vi_dev of this stage comes first");
  (void)stage;
  c->cur = c->hs.mid; c->xbuf = c->cur;     // q + impl_fac * k_im was produced by the column kernel together with k_im
  c->seam.vi_done = false;
  return FEDG_OK;
}

int fedg_cal_tend_ex_dev(fedg_ctx* c, int stage) {
  { int rc = seam_ready(c); if (rc) return rc; }
  if (!c->seam.in_step) return fail(FEDG_ERR_STATE, "fedg_rk_store_var0 opens a step");
  if (stage < 1 || stage > c->rk.nstage) return fail(FEDG_ERR_ARG, "stage out of range (1-based)");
  const int buf = c->cur, slot = c->hevi ? stage - 1 : 0;
  ensure_dp(c, buf);
  StageParams P{};
  fill_stage_params(c, P, buf, buf, c->hs.i0);
  for (int v = 0; v < NVAR; ++v) P.tend_out[v] = c->kex[size_t(slot) * NVAR + v].p;
  if (c->seam.halo_inflight) {           // exchange started, not waited for: interior elements, then the tile-boundary ones
    if (c->n_inner > 0) { P.elem_list = c->d_elem_inner; P.nelem = c->n_inner; launch_stage(P, c->np, c->terrain, c->moist, c->hevi, c->stream); }
    comm_exchange_wait(c->comm, c->stream); c->seam.halo_inflight = false;
    if (c->n_bnd > 0) { P.elem_list = c->d_elem_bnd; P.nelem = c->n_bnd; launch_stage(P, c->np, c->terrain, c->moist, c->hevi, c->stream); }
  } else if (c->seam.halo_ready && c->seam.halo_buf == buf) {
    launch_stage(P, c->np, c->terrain, c->moist, c->hevi, c->stream);
  } else {
    int rc = exchange_and_stage(c, P, buf, c->hevi); if (rc) return rc;
  }
  c->seam.halo_ready = false;
  c->hs.mid = buf;
  return seam_sync(c);
}

int fedg_rk_advance(fedg_ctx* c, int stage) {
  { int rc = seam_ready(c); if (rc) return rc; }
  if (!c->seam.in_step) return fail(FEDG_ERR_STATE, "fedg_rk_store_var0 opens a step");
  const int ns = c->rk.nstage, s = stage - 1;
  if (s < 0 || s >= ns) return fail(FEDG_ERR_ARG, "stage out of range (1-based)");
  if (c->hevi) {
    hevi_stage_combine(c, s, false);
    c->cur = c->hs.in; c->xbuf = c->cur;
  } else {
    c->hs.in = c->cur;
    heve_stage_prepare(c, s);
    const int in = c->hs.in, out = c->hs.nxt;
    AdvParams A{};
    for (int v = 0; v < NVAR; ++v) {
      A.q[v] = c->prog[in][v].p; A.q0[v] = c->prog[c->hs.i0][v].p; A.vt[v] = c->vt[v].p; A.k[v] = c->kex[v].p; A.out[v] = c->prog[out][v].p;
    }
    A.rk = c->stages[s]; A.n = c->nint;
    rk_advance_kernel<<<unsigned((c->nint + 255) / 256), 256, 0, c->stream>>>(A);
    c->dp_valid[out] = false;
    c->hs.in = out; c->cur = out; c->xbuf = out;
  }
  if (s == ns - 1) c->seam.in_step = false;
  return seam_sync(c);
}

int fedg_modalfilter_apply(fedg_ctx* c) {
  { int rc = seam_ready(c); if (rc) return rc; }
  if (!c->modalfilter) return FEDG_OK;
  double* q[NVAR];
  for (int v = 0; v < NVAR; ++v) q[v] = c->prog[c->cur][v].p;
  launch_modal_filter5(q, c->gsqrt.p, c->terrain || c->global, c->Ne, c->np, c->stream);
  c->dp_valid[c->cur] = false;
  return seam_sync(c);
}

int fedg_rk_get_tend(fedg_ctx* c, int implicit, int stage, double* DENS_dt, double* MOMX_dt, double* MOMY_dt, double* MOMZ_dt, double* RHOT_dt) {
  if (!c || !DENS_dt || !MOMX_dt || !MOMY_dt || !MOMZ_dt || !RHOT_dt) return fail(FEDG_ERR_ARG, "null argument");
  const auto& buf = implicit ? c->kim : c->kex;
  const int slot = c->hevi ? stage - 1 : 0;
  if (slot < 0 || size_t(slot + 1) * NVAR > buf.size()) return fail(FEDG_ERR_ARG, "no such tendency buffer");
  double* h[NVAR] = {DENS_dt, MOMX_dt, MOMY_dt, MOMZ_dt, RHOT_dt};
  for (int v = 0; v < NVAR; ++v)
    CUDA_TRY(cudaMemcpyAsync(h[v], buf[size_t(slot) * NVAR + v].p, c->nint * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  return seam_sync(c);
}

// ElementOperationBase3D%Div (scale_element_operation_base.F90:106-118; tensorprod3D.F90.erb:299-329): for nelem elements
// vec_out(:,1..3) = Dx vec_in(:,1), Dy vec_in(:,2), Dz vec_in(:,3) and vec_out(:,4) = Lift vec_in_lift.  Host arrays
// vec_in (Np,3,nelem), vec_in_lift (NfpTot,nelem), vec_out (Np,4,nelem).
int fedg_elem_div(fedg_ctx* c, const double* vec_in, const double* vec_in_lift, double* vec_out, int nelem) {
  if (!c || !vec_in || !vec_in_lift || !vec_out || nelem <= 0) return fail(FEDG_ERR_ARG, "bad argument");
  ensure_tables(c);
  const size_t Np = c->Np, NfT = c->NfpTot;
  std::vector<double> tmp(Np * nelem), res(Np * nelem);
  struct Bufs { DevBuf a, b; ~Bufs() { a.release(); b.release(); } } d;
  CUDA_TRY(d.a.alloc(std::max(Np, NfT) * nelem)); CUDA_TRY(d.b.alloc(Np * nelem));
  for (int k = 0; k < 4; ++k) {
    const size_t nin = (k == 3 ? NfT : Np) * nelem;
    const double* src = vec_in_lift;
    if (k < 3) {
      for (int e = 0; e < nelem; ++e) std::memcpy(&tmp[e * Np], vec_in + (size_t(e) * 3 + k) * Np, Np * sizeof(double));
      src = tmp.data();
    }
    CUDA_TRY(cudaMemcpyAsync(d.a.p, src, nin * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    launch_elem_op(k, d.a.p, d.b.p, nelem, c->np, c->stream);
    CUDA_TRY(cudaMemcpyAsync(res.data(), d.b.p, Np * nelem * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaGetLastError());
    for (int e = 0; e < nelem; ++e) std::memcpy(vec_out + (size_t(e) * 4 + k) * Np, &res[e * Np], Np * sizeof(double));
  }
  return FEDG_OK;
}

}  // extern "C"

// ---- sample/advect3d (config 1) --------------------------------------------------------------------------
namespace {
void fill_advect_params(fedg_ctx* c, AdvectParams& P, int in, int out, int i0) {
  AdvectState& a = c->adv;
  P.q = a.q[in].p; P.u = a.u.p; P.v = a.v.p; P.w = a.w.p;
  P.qout = a.q[out].p; P.q0 = a.q[i0].p; P.vt = a.vt.p; P.tend_out = nullptr;
  for (int k = 0; k < 4; ++k) { P.ellval[k] = a.ellval[k].p; P.ellcol[k] = a.ellcol[k]; P.colsz[k] = a.colsz[k]; }
  P.escale = c->escale.p; P.fscale = c->fscale.p; P.vmapP = c->d_vmapP;
  P.Np = c->Np; P.Nfp = c->Nfp; P.NfpTot = c->NfpTot; P.np = c->np; P.Ne = c->Ne; P.epb = a.epb;
}

// one time step of test_advect3d.f90:81-126 enqueued on the context's stream; ends with adv.cur unchanged
int enqueue_advect_step(fedg_ctx* c) {
  AdvectState& a = c->adv;
  const int ns = a.rk.nstage, i0 = a.cur;
  int in = i0;
  for (int s = 0; s < ns; ++s) {
    int out;
    if (s == ns - 1) out = (ns == 1) ? (i0 + 1) % 3 : i0;
    else { out = (in + 1) % 3; if (out == i0) out = (out + 1) % 3; }
    launch_advect_halo(a.q[in].p, a.u.p, a.v.p, a.w.p, c->d_halo_src, c->nint, c->Nhalo, false, c->stream);
    AdvectParams P{};
    fill_advect_params(c, P, in, out, i0);
    P.rk = a.stages[s];
    CUDA_TRY(launch_advect_stage(P, c->stream));
    in = out;
  }
  a.cur = in;
  return FEDG_OK;
}
}  // namespace

extern "C" {

int fedg_sparsemat_matmul(const fedg_sparsemat* A, const double* b, double* c, int nvec) {
  if (!A || !A->val || !A->colIdx || !b || !c || nvec <= 0 || A->M <= 0 || A->N <= 0 || A->col_size <= 0)
    return fail(FEDG_ERR_ARG, "bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(FEDG_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
  const size_t nell = size_t(A->M) * A->col_size;
  std::vector<int> col0(nell);
  for (size_t l = 0; l < nell; ++l) {
    col0[l] = A->colIdx[l] - 1;
    if (col0[l] < 0 || col0[l] >= A->N) return fail(FEDG_ERR_ARG, "colIdx out of range (must be 1-based)");
  }
  DevBuf dv, db, dc; int* dcol = nullptr;
  CUDA_TRY(dv.alloc(nell)); CUDA_TRY(db.alloc(size_t(A->N) * nvec)); CUDA_TRY(dc.alloc(size_t(A->M) * nvec));
  CUDA_TRY(cudaMalloc(&dcol, nell * sizeof(int)));
  cudaError_t e = cudaMemcpy(dv.p, A->val, nell * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dcol, col0.data(), nell * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(db.p, b, size_t(A->N) * nvec * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_ell_spmv(A->M, A->N, A->col_size, dv.p, dcol, db.p, dc.p, nvec, nullptr);
  if (e == cudaSuccess) e = cudaMemcpy(c, dc.p, size_t(A->M) * nvec * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dcol); dv.release(); db.release(); dc.release();
  if (e != cudaSuccess) return fail(FEDG_ERR_CUDA, cudaGetErrorString(e));
  return FEDG_OK;
}

namespace {
struct IntDevBuf {   // device int array freed on scope exit
  int* p = nullptr;
  ~IntDevBuf() { if (p) cudaFree(p); }
};
struct DevBufGuard {   // DevBuf has no destructor (contexts own theirs): release on scope exit here
  DevBuf b;
  ~DevBufGuard() { b.release(); }
};
// mode 0: c = A b1; mode 1: c = A (b1 .* b2); mode 2: c(NQ,M) = A b1(NQ,N)
int sparsemat_any_product(const fedg_sparsemat_any* A, const double* b1, const double* b2, double* c, int mode, int NQ) {
  if (!A || !A->val || !A->colIdx || !b1 || !c || (mode == 1 && !b2) || NQ <= 0 || A->M <= 0 || A->N <= 0) return fail(FEDG_ERR_ARG, "bad argument");
  const bool csr = A->storage_format_id == 1;
  if (!csr && A->storage_format_id != 2) return fail(FEDG_ERR_ARG, "storage_format_id must be 1 (CSR) or 2 (ELL)");
  if (csr && (!A->rowPtr || A->rowPtrSize != A->M + 1 || A->nnz < 0)) return fail(FEDG_ERR_ARG, "CSR needs rowPtr(M + 1) and nnz");
  if (!csr && A->col_size <= 0) return fail(FEDG_ERR_ARG, "ELL needs col_size");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(FEDG_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
  const size_t nent = csr ? size_t(A->nnz) : size_t(A->M) * A->col_size;
  std::vector<int> col0(std::max<size_t>(nent, 1)), rp0(csr ? A->M + 1 : 1);
  for (size_t l = 0; l < nent; ++l) {
    col0[l] = A->colIdx[l] - 1;
    if (col0[l] < 0 || col0[l] >= A->N) return fail(FEDG_ERR_ARG, "colIdx out of range (must be 1-based)");
  }
  if (csr) {
    for (int i = 0; i <= A->M; ++i) {
      rp0[i] = A->rowPtr[i] - 1;
      if (rp0[i] < 0 || size_t(rp0[i]) > nent || (i > 0 && rp0[i] < rp0[i - 1])) return fail(FEDG_ERR_ARG, "rowPtr must be 1-based, ascending, within nnz");
    }
  }
  DevBufGuard dv, d1, d2, dc; IntDevBuf dcol, drp;
  const size_t nb = size_t(A->N) * NQ, nc = size_t(A->M) * NQ;
  CUDA_TRY(dv.b.alloc(std::max<size_t>(nent, 1))); CUDA_TRY(d1.b.alloc(nb)); CUDA_TRY(dc.b.alloc(nc));
  if (mode == 1) CUDA_TRY(d2.b.alloc(nb));
  CUDA_TRY(cudaMalloc(&dcol.p, std::max<size_t>(nent, 1) * sizeof(int)));
  if (csr) CUDA_TRY(cudaMalloc(&drp.p, size_t(A->M + 1) * sizeof(int)));
  if (nent) { CUDA_TRY(cudaMemcpy(dv.b.p, A->val, nent * sizeof(double), cudaMemcpyHostToDevice)); CUDA_TRY(cudaMemcpy(dcol.p, col0.data(), nent * sizeof(int), cudaMemcpyHostToDevice)); }
  if (csr) CUDA_TRY(cudaMemcpy(drp.p, rp0.data(), size_t(A->M + 1) * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(d1.b.p, b1, nb * sizeof(double), cudaMemcpyHostToDevice));
  if (mode == 1) CUDA_TRY(cudaMemcpy(d2.b.p, b2, nb * sizeof(double), cudaMemcpyHostToDevice));
  // matmul2 keeps the right-hand-side index fastest: b(NQ,N), c(NQ,M)
  const size_t sb_col = (mode == 2) ? size_t(NQ) : 1, sb_q = (mode == 2) ? 1 : size_t(A->N), sc_row = (mode == 2) ? size_t(NQ) : 1, sc_q = (mode == 2) ? 1 : size_t(A->M);
  CUDA_TRY(launch_sparsemat_general(A->M, NQ, csr ? 0 : A->col_size, dv.b.p, dcol.p, csr ? drp.p : nullptr, d1.b.p, mode == 1 ? d2.b.p : nullptr, dc.b.p,
                                    sb_col, sb_q, sc_row, sc_q, nullptr));
  CUDA_TRY(cudaMemcpy(c, dc.b.p, nc * sizeof(double), cudaMemcpyDeviceToHost));
  return FEDG_OK;
}
}  // namespace

int fedg_sparsemat_matmul1(const fedg_sparsemat_any* A, const double* b, double* c) { return sparsemat_any_product(A, b, nullptr, c, 0, 1); }
int fedg_sparsemat_matmul1_2(const fedg_sparsemat_any* A, const double* b1, const double* b2, double* c) { return sparsemat_any_product(A, b1, b2, c, 1, 1); }
int fedg_sparsemat_matmul2(const fedg_sparsemat_any* A, const double* b, double* c, int NQ) { return sparsemat_any_product(A, b, nullptr, c, 2, NQ); }

int fedg_advect3d_init(fedg_ctx* c, const char* tinteg_type, double dt, const fedg_sparsemat* Dx, const fedg_sparsemat* Dy,
                       const fedg_sparsemat* Dz, const fedg_sparsemat* Lift) {
  if (!c || !tinteg_type || !Dx || !Dy || !Dz || !Lift) return fail(FEDG_ERR_ARG, "null argument");
  AdvectState& a = c->adv;
  a.release();
  if (!a.rk.init(tinteg_type)) return fail(FEDG_ERR_ARG, std::string("unsupported TINTEG_SCHEME_TYPE ") + tinteg_type);
  if (a.rk.imex) return fail(FEDG_ERR_ARG, "advect3d needs an explicit RK scheme");
  if (a.rk.tend_buf_size != 1) return fail(FEDG_ERR_UNSUPPORTED, "explicit schemes with several tendency buffers are not supported");
  if (!(dt > 0.0)) return fail(FEDG_ERR_ARG, "dt must be positive");
  if (c->comm.active && c->comm.nremote > 0) return fail(FEDG_ERR_UNSUPPORTED, "advect3d runs on a single tile");
  a.dt = dt;
  build_stage_coefs(a.rk, dt, a.stages, a.vt_used);
  const fedg_sparsemat* mats[4] = {Dx, Dy, Dz, Lift};
  for (int k = 0; k < 4; ++k) {
    const fedg_sparsemat* m = mats[k];
    if (m->M != c->Np || m->N != (k == 3 ? c->NfpTot : c->Np) || m->col_size <= 0 || !m->val || !m->colIdx)
      return fail(FEDG_ERR_ARG, "sparsemat shape does not match the element (Dx,Dy,Dz: Np x Np; Lift: Np x NfpTot)");
    if (k > 0 && k < 3 && m->col_size != Dx->col_size) return fail(FEDG_ERR_ARG, "Dx, Dy, Dz must share col_size");
    const size_t nell = size_t(m->M) * m->col_size;
    std::vector<int> col0(nell);
    for (size_t l = 0; l < nell; ++l) {
      col0[l] = m->colIdx[l] - 1;
      if (col0[l] < 0 || col0[l] >= m->N) return fail(FEDG_ERR_ARG, "sparsemat colIdx out of range (must be 1-based)");
    }
    a.colsz[k] = m->col_size;
    CUDA_TRY(a.ellval[k].alloc(nell));
    CUDA_TRY(cudaMalloc(&a.ellcol[k], nell * sizeof(int)));
    CUDA_TRY(cudaMemcpy(a.ellval[k].p, m->val, nell * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(a.ellcol[k], col0.data(), nell * sizeof(int), cudaMemcpyHostToDevice));
  }
  a.epb = std::max(1, 256 / c->Np);
  for (auto& b : a.q) CUDA_TRY(b.alloc(c->nall));
  for (DevBuf* b : {&a.u, &a.v, &a.w, &a.vt}) CUDA_TRY(b->alloc(c->nall));
  CUDA_TRY(a.tend.alloc(c->nint));
  AdvectParams P{};
  fill_advect_params(c, P, 0, 1, 0);
  if (advect_smem_bytes(P) > 227 * 1024) return fail(FEDG_ERR_UNSUPPORTED, "operators do not fit in shared memory at this order");
  a.cur = 0;
  a.ready = true;
  return FEDG_OK;
}

int fedg_advect3d_set(fedg_ctx* c, const double* q, const double* u, const double* v, const double* w) {
  if (!c || !q || !u || !v || !w) return fail(FEDG_ERR_ARG, "null argument");
  AdvectState& a = c->adv;
  if (!a.ready) return fail(FEDG_ERR_STATE, "fedg_advect3d_init must be called first");
  const double* h[4] = {q, u, v, w};
  double* d[4] = {a.q[a.cur].p, a.u.p, a.v.p, a.w.p};
  for (int k = 0; k < 4; ++k) CUDA_TRY(cudaMemcpyAsync(d[k], h[k], c->nall * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  launch_advect_halo(a.q[a.cur].p, a.u.p, a.v.p, a.w.p, c->d_halo_src, c->nint, c->Nhalo, true, c->stream);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  return FEDG_OK;
}

int fedg_advect3d_get(fedg_ctx* c, double* q) {
  if (!c || !q) return fail(FEDG_ERR_ARG, "null argument");
  if (!c->adv.ready) return fail(FEDG_ERR_STATE, "fedg_advect3d_init must be called first");
  CUDA_TRY(cudaMemcpyAsync(q, c->adv.q[c->adv.cur].p, c->nall * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return FEDG_OK;
}

int fedg_advect3d_cal_tend(fedg_ctx* c, double* dqdt) {
  if (!c || !dqdt) return fail(FEDG_ERR_ARG, "null argument");
  AdvectState& a = c->adv;
  if (!a.ready) return fail(FEDG_ERR_STATE, "fedg_advect3d_init must be called first");
  launch_advect_halo(a.q[a.cur].p, a.u.p, a.v.p, a.w.p, c->d_halo_src, c->nint, c->Nhalo, true, c->stream);
  AdvectParams P{};
  fill_advect_params(c, P, a.cur, a.cur, a.cur);
  P.tend_out = a.tend.p;
  CUDA_TRY(launch_advect_stage(P, c->stream));
  CUDA_TRY(cudaMemcpyAsync(dqdt, a.tend.p, c->nint * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  return FEDG_OK;
}

int fedg_advect3d_update(fedg_ctx* c, int nsteps) {
  if (!c || nsteps < 0) return fail(FEDG_ERR_ARG, "bad argument");
  AdvectState& a = c->adv;
  if (!a.ready) return fail(FEDG_ERR_STATE, "fedg_advect3d_init must be called first");
  while (c->ev.size() < 2) { cudaEvent_t e; CUDA_TRY(cudaEventCreate(&e)); c->ev.push_back(e); }
  if (!a.step_graph && nsteps > 1 && a.rk.nstage > 1) {
    // launch-bound at the sample's size (512 elements): capture one step (2 kernels per stage) and replay it
    cudaGraph_t g = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const int cur0 = a.cur;
    int rc = enqueue_advect_step(c);
    cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(FEDG_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
    if (a.cur != cur0) { cudaGraphDestroy(g); return fail(FEDG_ERR_STATE, "step does not return to its start buffer"); }
    e = cudaGraphInstantiate(&a.step_graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(FEDG_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
  }
  CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
  for (int n = 0; n < nsteps; ++n) {
    if (a.step_graph) CUDA_TRY(cudaGraphLaunch(a.step_graph, c->stream));
    else { int rc = enqueue_advect_step(c); if (rc) return rc; }
  }
  CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
  c->last_ms_total = ms; c->last_ms_stage = 0; c->last_launches = long(nsteps) * a.rk.nstage * 2;
  return FEDG_OK;
}

}  // extern "C"

// ---- several local meshes on one device (the reference's LOCAL_MESH_NUM > 1; cubed-sphere panels) ----------------
namespace {
// sendbuf[v][m] = field_v[idx[m]] for the six travelling fields (extract_bounddata for a panel edge owned by another rank)
__global__ void pack_link_kernel(LinkFields F, const int* __restrict__ idx, double* __restrict__ buf, int cnt) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= cnt) return;
  const int i = idx[m];
#pragma unroll
  for (int v = 0; v < 6; ++v) buf[size_t(v) * cnt + m] = F.s[v][i];
}

// Panel-edge data between the local meshes of different ranks: every mesh packs what its remote neighbours gather from its
// stage-input state (prog[xbuf], dp[xbuf]), one NCCL group ships all messages of the rank; the receive buffers are scattered
// (with the basis change) by fill_halo_links of the receiving mesh.  Runs on the group's single stream.
int group_exchange_remote(fedg_ctx** ctxs, int n, bool aux = false) {
  std::vector<P2PMsg> sends, recvs;
  fedg_ctx* lead = ctxs[0];
  for (int i = 0; i < n; ++i) {
    fedg_ctx* c = ctxs[i];
    for (auto& m : c->outmsg) {
      LinkFields F{};
      if (aux) { double* f[6]; aux_link_fields(c, f); for (int v = 0; v < 6; ++v) F.s[v] = f[v]; }
      else {
        for (int v = 0; v < NVAR; ++v) F.s[v] = c->prog[c->xbuf][v].p;
        F.s[5] = c->dp[c->xbuf].p;
      }
      pack_link_kernel<<<(m.cnt + 255) / 256, 256, 0, lead->stream>>>(F, m.d_idx, m.sendbuf, m.cnt);
      sends.push_back(P2PMsg{m.peer, m.msg_id, m.sendbuf, size_t(6) * m.cnt});
    }
    for (auto& l : c->link) if (l.recvbuf) recvs.push_back(P2PMsg{l.peer, l.msg_id, l.recvbuf, size_t(6) * l.cnt});
  }
  if (sends.empty() && recvs.empty()) return FEDG_OK;
  std::string err;
  int rc = comm_p2p_group(lead->comm, sends, recvs, lead->stream, err);
  if (rc) return fail(rc, err);
  return FEDG_OK;
}
}  // namespace


extern "C" {

int fedg_link_halo(fedg_ctx* c, int face, fedg_ctx* src, const int* src_index, const double* rot) {
  if (!c || !src || !src_index || face < 1 || face > 6) return fail(FEDG_ERR_ARG, "bad argument");
  const int f = face - 1, cnt = c->face_off[f + 1] - c->face_off[f];
  if (src->Np != c->Np) return fail(FEDG_ERR_ARG, "linked meshes must share the element");
  int dev_a = -1;
  cudaGetDevice(&dev_a);
  std::vector<int> idx(cnt);
  for (int m = 0; m < cnt; ++m) {
    idx[m] = src_index[m] - 1;
    if (idx[m] < 0 || size_t(idx[m]) >= src->nint) return fail(FEDG_ERR_ARG, "src_index out of range (1-based interior index of the source mesh)");
  }
  auto& l = c->link[f];
  if (l.d_src) cudaFree(l.d_src);
  if (l.d_rot) cudaFree(l.d_rot);
  if (l.recvbuf) cudaFree(l.recvbuf);
  l = fedg_ctx::HaloLink{};
  CUDA_TRY(cudaMalloc(&l.d_src, size_t(std::max(cnt, 1)) * sizeof(int)));
  CUDA_TRY(cudaMemcpy(l.d_src, idx.data(), size_t(cnt) * sizeof(int), cudaMemcpyHostToDevice));
  if (rot) {
    CUDA_TRY(cudaMalloc(&l.d_rot, size_t(std::max(cnt, 1)) * 4 * sizeof(double)));
    CUDA_TRY(cudaMemcpy(l.d_rot, rot, size_t(cnt) * 4 * sizeof(double), cudaMemcpyHostToDevice));
  }
  l.src = src; l.off = c->face_off[f]; l.cnt = cnt;
  return FEDG_OK;
}

int fedg_link_halo_recv(fedg_ctx* c, int face, int peer_rank, int msg_id, const double* rot) {
  if (!c || face < 1 || face > 6 || peer_rank < 0) return fail(FEDG_ERR_ARG, "bad argument");
  if (peer_rank == c->my_rank) return fail(FEDG_ERR_ARG, "the source mesh is on this rank: use fedg_link_halo");
  const int f = face - 1, cnt = c->face_off[f + 1] - c->face_off[f];
  auto& l = c->link[f];
  if (l.d_src) cudaFree(l.d_src);
  if (l.d_rot) cudaFree(l.d_rot);
  if (l.recvbuf) cudaFree(l.recvbuf);
  l = fedg_ctx::HaloLink{};
  CUDA_TRY(cudaMalloc(&l.recvbuf, size_t(std::max(cnt, 1)) * 6 * sizeof(double)));
  if (rot) {
    CUDA_TRY(cudaMalloc(&l.d_rot, size_t(std::max(cnt, 1)) * 4 * sizeof(double)));
    CUDA_TRY(cudaMemcpy(l.d_rot, rot, size_t(cnt) * 4 * sizeof(double), cudaMemcpyHostToDevice));
  }
  l.off = c->face_off[f]; l.cnt = cnt; l.peer = peer_rank; l.msg_id = msg_id;
  return FEDG_OK;
}

int fedg_link_halo_send(fedg_ctx* c, int peer_rank, int msg_id, const int* src_index, int n) {
  if (!c || !src_index || n < 1 || peer_rank < 0) return fail(FEDG_ERR_ARG, "bad argument");
  if (peer_rank == c->my_rank) return fail(FEDG_ERR_ARG, "the receiving mesh is on this rank: use fedg_link_halo");
  std::vector<int> idx(n);
  for (int m = 0; m < n; ++m) {
    idx[m] = src_index[m] - 1;
    if (idx[m] < 0 || size_t(idx[m]) >= c->nint) return fail(FEDG_ERR_ARG, "src_index out of range (1-based interior index of this mesh)");
  }
  fedg_ctx::OutMsg m{};
  m.peer = peer_rank; m.msg_id = msg_id; m.cnt = n;
  CUDA_TRY(cudaMalloc(&m.d_idx, size_t(n) * sizeof(int)));
  CUDA_TRY(cudaMemcpy(m.d_idx, idx.data(), size_t(n) * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&m.sendbuf, size_t(n) * 6 * sizeof(double)));
  c->outmsg.push_back(m);
  return FEDG_OK;
}

int fedg_group_exchange_halo(fedg_ctx** ctxs, int n, int apply_bc) {
  if (!ctxs || n < 1) return fail(FEDG_ERR_ARG, "bad argument");
  for (int i = 0; i < n; ++i) if (!ctxs[i] || !ctxs[i]->aux_ready) return fail(FEDG_ERR_STATE, "fedg_set_aux must be called on every mesh of the group");
  fedg_ctx* lead = ctxs[0];
  std::vector<cudaStream_t> saved(n);
  for (int i = 0; i < n; ++i) { CUDA_TRY(cudaStreamSynchronize(ctxs[i]->stream)); saved[i] = ctxs[i]->stream; ctxs[i]->stream = lead->stream; }
  struct Restore { fedg_ctx** c; std::vector<cudaStream_t>& s; int n; ~Restore() { for (int i = 0; i < n; ++i) c[i]->stream = s[i]; } } restore{ctxs, saved, n};
  for (int i = 0; i < n; ++i) { ensure_dp(ctxs[i], ctxs[i]->cur); ctxs[i]->xbuf = ctxs[i]->cur; }
  { int rc = group_exchange_remote(ctxs, n); if (rc) return rc; }
  for (int i = 0; i < n; ++i) fill_halo(ctxs[i], ctxs[i]->cur, apply_bc != 0);
  CUDA_TRY(cudaStreamSynchronize(lead->stream));
  CUDA_TRY(cudaGetLastError());
  return FEDG_OK;
}

int fedg_group_exchange_aux(fedg_ctx** ctxs, int n) {
  if (!ctxs || n < 1) return fail(FEDG_ERR_ARG, "bad argument");
  for (int i = 0; i < n; ++i) {
    if (!ctxs[i] || !ctxs[i]->aux_ready) return fail(FEDG_ERR_STATE, "fedg_set_aux must be called on every mesh of the group");
    if (ctxs[i]->moist) return fail(FEDG_ERR_UNSUPPORTED, "linked halos of Rtot / CVtot / CPtot are not exchanged yet: dry background only");
  }
  fedg_ctx* lead = ctxs[0];
  std::vector<cudaStream_t> saved(n);
  for (int i = 0; i < n; ++i) { CUDA_TRY(cudaStreamSynchronize(ctxs[i]->stream)); saved[i] = ctxs[i]->stream; ctxs[i]->stream = lead->stream; }
  struct Restore { fedg_ctx** c; std::vector<cudaStream_t>& s; int n; ~Restore() { for (int i = 0; i < n; ++i) c[i]->stream = s[i]; } } restore{ctxs, saved, n};
  { int rc = group_exchange_remote(ctxs, n, true); if (rc) return rc; }
  for (int i = 0; i < n; ++i) fill_halo_links(ctxs[i], 0, true);
  CUDA_TRY(cudaStreamSynchronize(lead->stream));
  CUDA_TRY(cudaGetLastError());
  for (int i = 0; i < n; ++i) for (bool& v : ctxs[i]->dp_valid) v = false;
  return FEDG_OK;
}

int fedg_group_update(fedg_ctx** ctxs, int n, int nsteps) {
  if (!ctxs || n < 1 || nsteps < 0) return fail(FEDG_ERR_ARG, "bad argument");
  for (int i = 0; i < n; ++i) {
    fedg_ctx* c = ctxs[i];
    if (!c) return fail(FEDG_ERR_ARG, "null context");
    if (!c->dyn_ready || !c->aux_ready) return fail(FEDG_ERR_STATE, "fedg_dyn_init and fedg_set_aux must be called on every mesh of the group");
    if (c->hevi != ctxs[0]->hevi) return fail(FEDG_ERR_ARG, "the meshes of a group share the equation set");
    if (c->comm.active && c->comm.nremote > 0) return fail(FEDG_ERR_UNSUPPORTED, "group stepping and NCCL tiles cannot be combined yet");
    if (c->rk.nstage != ctxs[0]->rk.nstage || c->dt != ctxs[0]->dt) return fail(FEDG_ERR_ARG, "the meshes of a group share scheme and step");
    if (c->nd.on && c->nd.in_update)
      return fail(FEDG_ERR_UNSUPPORTED, "numerical diffusion inside the update is not available for a group of local meshes (its work-field exchange is per mesh)");
  }
  // one stream for the whole group: the order of the launches is the dependency order between the meshes
  fedg_ctx* lead = ctxs[0];
  std::vector<cudaStream_t> saved(n);
  for (int i = 0; i < n; ++i) { CUDA_TRY(cudaStreamSynchronize(ctxs[i]->stream)); saved[i] = ctxs[i]->stream; ctxs[i]->stream = lead->stream; }
  struct Restore { fedg_ctx** c; std::vector<cudaStream_t>& s; int n; ~Restore() { for (int i = 0; i < n; ++i) c[i]->stream = s[i]; } } restore{ctxs, saved, n};
  for (int i = 0; i < n; ++i) { ensure_tables(ctxs[i]); ensure_dp(ctxs[i], ctxs[i]->cur); }
  while (lead->ev.size() < 2) { cudaEvent_t e; CUDA_TRY(cudaEventCreate(&e)); lead->ev.push_back(e); }
  CUDA_TRY(cudaEventRecord(lead->ev[0], lead->stream));
  const int ns = lead->rk.nstage;
  long launches = 0;
  // FEDG_GROUP_TIMING=1 (diagnostic): events around every remote exchange, the sum is printed per rank on stderr
  static const bool timing = [] { const char* e = getenv("FEDG_GROUP_TIMING"); return e && e[0] == '1'; }();
  std::vector<cudaEvent_t> tev;
  auto mark = [&]() { if (timing) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, lead->stream); tev.push_back(e); } };
  for (int step = 0; step < nsteps; ++step) {
    for (int i = 0; i < n; ++i) hevi_begin_step(ctxs[i]);
    for (int s = 0; s < ns; ++s) {
      if (lead->hevi) {
        for (int i = 0; i < n; ++i) { int rc = hevi_stage_vi(ctxs[i], s, nullptr, nullptr); if (rc) return rc; }     // cal_vi + StoreImplicit
        mark();
        { int rc = group_exchange_remote(ctxs, n); if (rc) return rc; }                                              // panel edges owned by other ranks
        mark();
        for (int i = 0; i < n; ++i) { int rc = hevi_stage_ex(ctxs[i], s); if (rc) return rc; }                        // exchange + cal_tend_ex
        for (int i = 0; i < n; ++i) hevi_stage_combine(ctxs[i], s);                                                  // Advance
        launches += 4L * n;
      } else {
        for (int i = 0; i < n; ++i) heve_stage_prepare(ctxs[i], s);                                                  // pressure of every stage input
        { int rc = group_exchange_remote(ctxs, n); if (rc) return rc; }
        for (int i = 0; i < n; ++i) { int rc = heve_stage(ctxs[i], s, nullptr, nullptr); if (rc) return rc; }        // exchange + tendency + Advance
        launches += 2L * n;
      }
    }
    for (int i = 0; i < n; ++i) hevi_end_step(ctxs[i]);
  }
  CUDA_TRY(cudaEventRecord(lead->ev[1], lead->stream));
  CUDA_TRY(cudaStreamSynchronize(lead->stream));
  CUDA_TRY(cudaGetLastError());
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, lead->ev[0], lead->ev[1]));
  lead->last_ms_total = ms; lead->last_ms_stage = 0; lead->last_launches = launches;
  if (timing) {
    float ex = 0, mx = 0;
    for (size_t k = 0; k + 1 < tev.size(); k += 2) { float t = 0; cudaEventElapsedTime(&t, tev[k], tev[k + 1]); ex += t; mx = std::max(mx, t); }
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
    fprintf(stderr, "[fedg group timing] rank %d: %d steps %.2f ms total, remote exchanges %.2f ms (%zu, longest %.3f ms)\n", lead->my_rank, nsteps, ms, ex,
            tev.size() / 2, mx);
  }
  return FEDG_OK;
}

}  // extern "C"

// ---- numerical diffusion (row f1) ---------------------------------------------------------------------------
extern "C" {

int fedg_numdiff_init(fedg_ctx* c, int nd_laplacian_num, double nd_coef_h, double nd_coef_v, const int* therm_bc, int apply_in_update) {
  if (!c) return fail(FEDG_ERR_ARG, "null argument");
  if (!c->dyn_ready) return fail(FEDG_ERR_STATE, "fedg_dyn_init must be called first (the diffusion step uses TIME_DT)");
  if (nd_laplacian_num < 1) return fail(FEDG_ERR_ARG, "ND_LAPLACIAN_NUM must be >= 1");
  if (c->terrain || c->global) return fail(FEDG_ERR_UNSUPPORTED, "numerical diffusion is available on the flat regional mesh only");
  for (int f = 0; f < 6; ++f) if (c->link[f].src) return fail(FEDG_ERR_UNSUPPORTED, "numerical diffusion with linked local meshes is not available");
  c->nd.lap_num = nd_laplacian_num; c->nd.coef_h = nd_coef_h; c->nd.coef_v = nd_coef_v;
  for (int f = 0; f < 6; ++f) c->nd.therm_bc[f] = therm_bc ? therm_bc[f] : 0;
  c->nd.in_update = apply_in_update != 0;
  c->nd.on = true;
  return FEDG_OK;
}

int fedg_numdiff_apply(fedg_ctx* c) {
  if (!c) return fail(FEDG_ERR_ARG, "null argument");
  if (!c->nd.on) return fail(FEDG_ERR_STATE, "fedg_numdiff_init must be called first");
  if (!c->aux_ready) return fail(FEDG_ERR_STATE, "fedg_set_aux must be called first");
  int rc = run_numdiff(c, c->cur);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  return FEDG_OK;
}

}  // extern "C"

// ---- sponge layer (row f3) -----------------------------------------------------------------------------------
namespace {
struct DevBufGuardZ { DevBuf b; ~DevBufGuardZ() { b.release(); } };
// calc_wdampcoef (spongelayer.F90:189-218) for every node from zlev
__global__ void sponge_coef_kernel(const double* __restrict__ zlev, double* __restrict__ coef, double r_tau, double height, int Np, int Nfp,
                                   int np, int Ne, int Ne2D, int NeZ) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= size_t(Np) * Ne) return;
  const int ke = int(i / Np), p = int(i - size_t(ke) * Np);
  const int keTop = (ke % Ne2D) + (NeZ - 1) * Ne2D;
  const double z = zlev[i], zTop = zlev[size_t(keTop) * Np + (p % Nfp) + (np - 1) * Nfp];
  coef[i] = 0.25 * r_tau * (1.0 + (z - height >= 0.0 ? 1.0 : -1.0)) * (1.0 - cos(3.14159265358979323846 * (z - height) / (zTop - height)));
}
}  // namespace

namespace {
// zsrc: device array (Np, Ne) of the computational height the damping profile is evaluated in
int sponge_init_from(fedg_ctx* c, const double* zsrc, double sl_wdamp_tau, double sl_wdamp_height, int sl_wdamp_layer, int sl_horiveldamp_flag) {
  if (sl_wdamp_layer > c->NeZ) return fail(FEDG_ERR_ARG, "SL_wdamp_layer should be less than total of vertical elements (NeGZ)");
  double tau = sl_wdamp_tau, height = sl_wdamp_height;
  if (sl_wdamp_layer > 0) {   // height of the first node of that layer (spongelayer.F90:104-106)
    CUDA_TRY(cudaMemcpy(&height, zsrc + size_t(sl_wdamp_layer - 1) * c->Ne2D * c->Np, sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (tau < 0.0) tau = c->dt * 10.0;
  else if (tau < c->dt) return fail(FEDG_ERR_ARG, "SL_wdamp_tau should be larger than TIME_DT (ATMOS_DYN)");
  if (c->sponge.n < c->nint) CUDA_TRY(c->sponge.alloc(c->nint));
  sponge_coef_kernel<<<unsigned((c->nint + 255) / 256), 256, 0, c->stream>>>(zsrc, c->sponge.p, 1.0 / tau, height, c->Np, c->Nfp, c->np, c->Ne,
                                                                          c->Ne2D, c->NeZ);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  c->sponge_h = sl_horiveldamp_flag ? 1.0 : 0.0;
  c->has_sponge = true;
  return FEDG_OK;
}
}  // namespace

extern "C" int fedg_sponge_init(fedg_ctx* c, double sl_wdamp_tau, double sl_wdamp_height, int sl_wdamp_layer, int sl_horiveldamp_flag) {
  if (!c) return fail(FEDG_ERR_ARG, "null argument");
  if (!c->dyn_ready) return fail(FEDG_ERR_STATE, "fedg_dyn_init must be called first (the default SL_WDAMP_TAU is 10 TIME_DT)");
  // the reference damps in the computational height pos_en(:,:,3) (spongelayer.F90:168-172); the descriptor carries zlev, which
  // equals it on meshes without topography (regional flat, cubed sphere) only: with topography use fedg_sponge_init_pos
  if (c->terrain) return fail(FEDG_ERR_UNSUPPORTED, "mesh with topography: hand the computational height over with fedg_sponge_init_pos");
  return sponge_init_from(c, c->zlev.p, sl_wdamp_tau, sl_wdamp_height, sl_wdamp_layer, sl_horiveldamp_flag);
}

extern "C" int fedg_sponge_init_pos(fedg_ctx* c, double sl_wdamp_tau, double sl_wdamp_height, int sl_wdamp_layer, int sl_horiveldamp_flag,
                                    const double* pos_en3) {
  if (!c || !pos_en3) return fail(FEDG_ERR_ARG, "null argument");
  if (!c->dyn_ready) return fail(FEDG_ERR_STATE, "fedg_dyn_init must be called first (the default SL_WDAMP_TAU is 10 TIME_DT)");
  DevBufGuardZ z;
  CUDA_TRY(z.b.alloc(c->nint));
  CUDA_TRY(cudaMemcpy(z.b.p, pos_en3, c->nint * sizeof(double), cudaMemcpyHostToDevice));
  return sponge_init_from(c, z.b.p, sl_wdamp_tau, sl_wdamp_height, sl_wdamp_layer, sl_horiveldamp_flag);
}

// ---- tracer advection with a prescribed mass flux (row f4; NOT yet validated on hardware, see tracer.cu) ------------------
extern "C" int fedg_trcadv_init(fedg_ctx* c, const char* tinteg_type, double dt, int modalfilter_flag, const double* filter_h1D,
                                const double* filter_v1D, int disable_limiter) {
  if (!c || !tinteg_type) return fail(FEDG_ERR_ARG, "null argument");
  TracerState& t = c->trc;
  t.release();
  if (!t.rk.init(tinteg_type)) return fail(FEDG_ERR_ARG, std::string("unsupported TINTEG_TYPE ") + tinteg_type);
  if (t.rk.imex || !t.rk.low_storage) return fail(FEDG_ERR_UNSUPPORTED, "the tracer integrator takes the low-storage explicit schemes (Advance_trcvar)");
  if (!(dt > 0.0)) return fail(FEDG_ERR_ARG, "dt must be positive");
  if (c->terrain || c->global) return fail(FEDG_ERR_UNSUPPORTED, "tracer advection is available on the flat regional mesh only");
  if (c->comm.active && c->comm.nremote > 0) return fail(FEDG_ERR_UNSUPPORTED, "tracer advection runs on a single tile");
  if (c->np != 4 && c->np != 8) return fail(FEDG_ERR_UNSUPPORTED, "tracer advection: p = 3 or p = 7");
  if (!c->aux_ready) return fail(FEDG_ERR_STATE, "fedg_set_aux must be called first (DENS_hyd)");
  t.dt = dt; t.disable_limiter = disable_limiter != 0; t.do_filter = modalfilter_flag != 0;
  const int np = c->np;
  std::vector<double> f(size_t(2) * np * np, 0.0);
  for (int i = 0; i < np; ++i) f[i * np + i] = f[np * np + i * np + i] = 1.0;
  if (t.do_filter) {
    if (!filter_h1D || !filter_v1D) return fail(FEDG_ERR_ARG, "modal filter matrices missing");
    for (int i = 0; i < np; ++i)
      for (int l = 0; l < np; ++l) { f[i * np + l] = filter_h1D[i + l * np]; f[np * np + i * np + l] = filter_v1D[i + l * np]; }   // column-major in
  }
  CUDA_TRY(t.filt.alloc(f.size()));
  CUDA_TRY(cudaMemcpy(t.filt.p, f.data(), f.size() * sizeof(double), cudaMemcpyHostToDevice));
  // 1D LGL weights from IntWeight_lgl(i,1,1) = w(i) w(1)^2 and sum_i w(i) = 2
  std::vector<double> w3(np);
  CUDA_TRY(cudaMemcpy(w3.data(), c->w3.p, np * sizeof(double), cudaMemcpyDeviceToHost));
  double sum = 0.0;
  for (int i = 0; i < np; ++i) sum += w3[i];
  for (int i = 0; i < np; ++i) t.w1d[i] = w3[i] * 2.0 / sum;
  for (auto& b : t.q) CUDA_TRY(b.alloc(c->nall));
  CUDA_TRY(t.fct.alloc(c->nall));
  CUDA_TRY(t.rhoq.alloc(c->nall));
  CUDA_TRY(t.var0.alloc(c->nint));
  CUDA_TRY(t.vartmp.alloc(c->nint));
  CUDA_TRY(t.alphM.alloc(size_t(c->NfpTot) * c->Ne));
  CUDA_TRY(t.alphP.alloc(size_t(c->NfpTot) * c->Ne));
  t.ready = true;
  return FEDG_OK;
}

namespace {
void trc_common_params(fedg_ctx* c, TracerParams& P, bool with_rhoq) {
  TracerState& t = c->trc;
  P.dens_hyd = c->dens_hyd.p;
  P.fct = t.fct.p; P.rhoq_tp = with_rhoq ? t.rhoq.p : nullptr;
  P.var0 = t.var0.p; P.vartmp = t.vartmp.p;
  P.escale = c->escale.p; P.fscale = c->fscale.p; P.jac = c->Jac.p; P.w3 = c->w3.p; P.vmapP = c->d_vmapP; P.tab = c->d_tab; P.filt = t.filt.p;
  for (int i = 0; i < MAXNP; ++i) P.w1d[i] = t.w1d[i];
  P.Np = c->Np; P.Nfp = c->Nfp; P.NfpTot = c->NfpTot; P.np = c->np; P.Ne = c->Ne;
  P.disable_limiter = t.disable_limiter;
}
// the stage loop of AtmDynDGMDriver_trcadv3d_update (driver_trcadv3d.F90:426-528) on t.q[0]; returns the buffer holding the result
int trc_run_stages(fedg_ctx* c, TracerParams& P, int nsteps, bool tmar, int& in) {
  TracerState& t = c->trc;
  const RKTable& rk = t.rk;
  const int ns = rk.nstage;
  const double EPS = 2.220446e-16;
  in = 0;
  for (int step = 0; step < nsteps; ++step)
    for (int st = 0; st < ns; ++st) {
      P.q = t.q[in].p; P.qout = t.q[in ^ 1].p;
      P.stage = st; P.nstage = ns;
      P.sig_ss = rk.sg(st + 1, st); P.gam_ss = t.dt * rk.gm(st + 1, st);
      P.sig_Ns = rk.sg(ns, st); P.gam_Ns = t.dt * rk.gm(ns, st);
      P.upd_vartmp = (std::fabs(P.sig_Ns) > EPS || std::fabs(rk.gm(ns, st)) > EPS) ? 1 : 0;
      double c0 = 0.0, c1 = 0.0;
      for (int j = 0; j < ns; ++j) { c0 += rk.aex(st, j); if (st + 1 < ns) c1 += rk.aex(st + 1, j); }
      P.c_ssm1 = c0; P.c_ss = c1;
      P.dttmp = t.dt * rk.gm(st + 1, st) / rk.sg(st + 1, st);
      P.do_filter = (st == ns - 1 && t.do_filter) ? 1 : 0;
      P.do_tmar = (st == ns - 1 && tmar) ? 1 : 0;
      if (c->Nhalo > 0) aux_halo_kernel<<<(c->Nhalo + 255) / 256, 256, 0, c->stream>>>(t.q[in].p, c->d_halo_src, int(c->nint), c->Nhalo);
      { cudaError_t e = launch_trc_fct(P, c->stream); if (e != cudaSuccess) return fail(FEDG_ERR_CUDA, cudaGetErrorString(e)); }
      if (c->Nhalo > 0) aux_halo_kernel<<<(c->Nhalo + 255) / 256, 256, 0, c->stream>>>(t.fct.p, c->d_halo_src, int(c->nint), c->Nhalo);
      { cudaError_t e = launch_trc_stage(P, c->stream); if (e != cudaSuccess) return fail(FEDG_ERR_CUDA, cudaGetErrorString(e)); }
      in ^= 1;
    }
  return FEDG_OK;
}
}  // namespace

extern "C" int fedg_trcadv_update(fedg_ctx* c, double* QTRC, const double* RHOQ_tp, int nsteps) {
  if (!c || !QTRC || nsteps < 0) return fail(FEDG_ERR_ARG, "bad argument");
  TracerState& t = c->trc;
  if (!t.ready) return fail(FEDG_ERR_STATE, "fedg_trcadv_init must be called first");
  ensure_tables(c);
  const int cur = c->cur;
  ensure_dp(c, cur);
  fill_halo(c, cur, true);           // halo + boundary condition of the state whose momentum is the mass flux (driver_trcadv3d.F90:404-420)
  CUDA_TRY(cudaMemcpyAsync(t.q[0].p, QTRC, c->nint * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (RHOQ_tp) CUDA_TRY(cudaMemcpyAsync(t.rhoq.p, RHOQ_tp, c->nint * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  TracerParams P{};
  trc_common_params(c, P, RHOQ_tp != nullptr);
  P.mfx = c->prog[cur][V_MOMX].p; P.mfy = c->prog[cur][V_MOMY].p; P.mfz = c->prog[cur][V_MOMZ].p;
  P.ddens = c->prog[cur][V_DDENS].p; P.ddens0 = c->prog[cur][V_DDENS].p;
  P.alphM = t.alphM.p; P.alphP = t.alphP.p;
  P.q = t.q[0].p;
  { cudaError_t e = launch_trc_alphdens(P, c->stream); if (e != cudaSuccess) return fail(FEDG_ERR_CUDA, cudaGetErrorString(e)); }
  int in = 0;
  { int rc = trc_run_stages(c, P, nsteps, !t.disable_limiter, in); if (rc) return rc; }
  CUDA_TRY(cudaMemcpyAsync(QTRC, t.q[in].p, c->nint * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  if (in != 0) std::swap(t.q[0], t.q[1]);
  return FEDG_OK;
}

extern "C" int fedg_trcadv_couple(fedg_ctx* c, int on) {
  if (!c) return fail(FEDG_ERR_ARG, "null argument");
  TracerState& t = c->trc;
  if (!on) { t.release_coupling(); return FEDG_OK; }
  if (c->terrain || c->global) return fail(FEDG_ERR_UNSUPPORTED, "tracer advection is available on the flat regional mesh only");
  if (c->comm.active && c->comm.nremote > 0) return fail(FEDG_ERR_UNSUPPORTED, "tracer advection runs on a single tile");
  for (int f = 0; f < 6; ++f) if (c->link[f].src || c->link[f].recvbuf) return fail(FEDG_ERR_UNSUPPORTED, "tracer advection runs on a single local mesh");
  for (auto& b : t.mflx) CUDA_TRY(b.alloc(c->nall));
  CUDA_TRY(t.alphM_t.alloc(size_t(c->NfpTot) * c->Ne)); CUDA_TRY(t.alphP_t.alloc(size_t(c->NfpTot) * c->Ne));
  CUDA_TRY(t.dens0.alloc(c->nall)); CUDA_TRY(t.dens1.alloc(c->nall));
  t.couple = true; t.have_flux = false;
  return FEDG_OK;
}

extern "C" int fedg_trcadv_update_coupled(fedg_ctx* c, double* QTRC, const double* RHOQ_tp) {
  if (!c || !QTRC) return fail(FEDG_ERR_ARG, "bad argument");
  TracerState& t = c->trc;
  if (!t.ready) return fail(FEDG_ERR_STATE, "fedg_trcadv_init must be called first");
  if (!t.couple || !t.have_flux) return fail(FEDG_ERR_STATE, "no accumulated mass flux: call fedg_trcadv_couple and run a dynamics step first");
  ensure_tables(c);
  const int cur = c->cur, sp = (cur + 1) % 3;
  // MeshFieldComm_Exchange of the averaged mass flux + ApplyBC_PROGVARS_lc on it (driver_trcadv3d.F90:404-420): staged through a spare
  // state buffer so that the halo / boundary-condition kernel of the dynamics does it
  const double* src[NVAR];
  src[V_DDENS] = t.dens1.p; src[V_MOMX] = t.mflx[0].p; src[V_MOMY] = t.mflx[1].p; src[V_MOMZ] = t.mflx[2].p; src[V_DRHOT] = c->prog[cur][V_DRHOT].p;
  for (int v = 0; v < NVAR; ++v)
    CUDA_TRY(cudaMemcpyAsync(c->prog[sp][v].p, src[v], c->nint * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  c->dp_valid[sp] = false;
  ensure_dp(c, sp);
  fill_halo(c, sp, true);
  c->dp_valid[sp] = false;
  CUDA_TRY(cudaMemcpyAsync(t.q[0].p, QTRC, c->nint * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (RHOQ_tp) CUDA_TRY(cudaMemcpyAsync(t.rhoq.p, RHOQ_tp, c->nint * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  TracerParams P{};
  trc_common_params(c, P, RHOQ_tp != nullptr);
  P.mfx = c->prog[sp][V_MOMX].p; P.mfy = c->prog[sp][V_MOMY].p; P.mfz = c->prog[sp][V_MOMZ].p;
  P.ddens = t.dens1.p; P.ddens0 = t.dens0.p;
  P.alphM = t.alphM_t.p; P.alphP = t.alphP_t.p;
  int in = 0;
  { int rc = trc_run_stages(c, P, 1, false, in); if (rc) return rc; }
  CUDA_TRY(launch_trc_rescale(t.q[in].p, c->dens_hyd.p, t.dens1.p, c->prog[cur][V_DDENS].p, c->nint, c->stream));
  CUDA_TRY(cudaMemcpyAsync(QTRC, t.q[in].p, c->nint * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaGetLastError());
  if (in != 0) std::swap(t.q[0], t.q[1]);
  return FEDG_OK;
}
