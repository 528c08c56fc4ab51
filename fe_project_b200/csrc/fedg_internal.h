// Internal declarations shared by the C-ABI layer and the CUDA kernels (sm_100a, FP64).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/fedg.h"

namespace fedg {

constexpr int NVAR = 5;
// internal variable order of the per-variable pointer arrays
enum { V_DDENS = 0, V_MOMX = 1, V_MOMY = 2, V_MOMZ = 3, V_DRHOT = 4 };

constexpr int MAXNP = 8;  // nodes per direction supported by the kernels (p <= 7)

// Element operator tables, uploaded to __constant__ memory (uniform indices become constant-bank
// operands of DFMA; per-thread rows are fetched once per kernel).
struct ElemTables {
  double D[MAXNP * MAXNP];    // D1D[i][l], row-major
  double Lw[MAXNP * 2];       // lift1d[m][side]
  double VP[MAXNP * MAXNP];   // VPOrdM1[k][l]
  double Fh[MAXNP * MAXNP];   // horizontal modal filter [i][l]
  double Fv[MAXNP * MAXNP];   // vertical modal filter [k][l]
  int np;
  int pad;
};

// Coefficients of one explicit RK stage in the fused update
//   q_new  = c_q0*q0 + c_q*q + c_k*k  (+ varTmp when add_vt)
//   varTmp = (vt_init ? vt_init_q*q : varTmp) + vt_q*q + vt_k*k   (when vt_update)
struct RKStage {
  double c_q0, c_q, c_k;
  double vt_init_q, vt_q, vt_k;
  int use_q0, add_vt, vt_update, vt_init;
};

struct PhysConst {
  double GRAV, Rdry, CPdry, CVdry, PRES00, rP0, gamm, CPovCV;
};

struct StageParams {
  const double* qin[NVAR];
  double* qout[NVAR];
  const double* q0[NVAR];
  double* vt[NVAR];
  double* tend_out[NVAR];  // when non-null: write the tendency instead of the updated state
  const double* dpin;      // DPRES of the stage-input state (interior + halo)
  double* dpout;           // DPRES of the stage-output state (interior), written by the kernel
  const ElemTables* tab;   // device copy of the operator tables
  const double *dens_hyd, *pres_hyd, *therm_hyd, *rtot, *cvtot, *cptot;
  const double *gsqrt, *g13, *g23, *gsqrtH;
  const double *dphydx, *dphydy, *coriolis;
  const double* escale;  // [3][Ne]
  const double* fscale;  // [6][Ne]
  const int* vmapP;      // (NfpTot,Ne) 0-based
  const int* emap2d;     // (Ne) 0-based
  double* pres_out;
  RKStage rk;
  PhysConst c;
  int Ne, Ne2D;
  const int* elem_list;  // when non-null: process elements elem_list[0..nelem) (interior / tile-boundary split)
  int nelem;
  int has_cor, has_phyd, do_filter, write_pres, exact_pow;
  // global (cubed-sphere panel) equation set: 2D metric tables [6][Ne2D*Nfp] = GsqrtH, G11, G12, G22, X = tan(alpha),
  // Y = tan(beta); planetary rotation rate; panel id 1..6
  const double* g2d;
  double OHM;
  int is_global, panel;
  // physics tendencies DENS_tp, MOMX_tp, MOMY_tp, MOMZ_tp, RHOT_tp, RHOH_p (add_phy_tend, driver_nonhydro3d.F90:1098-1178)
  const double* phyt[6];
  int has_phyt;
  // sponge layer: Rayleigh damping coefficient per node (Np,Ne), NULL = off; sponge_h = 1 damps MOMX / MOMY too
  const double* sponge;
  double sponge_h;
  int l2_prefetch;       // stage_p7: L2 prefetch of the late-use fields at block start (A/B knob FEDG_P7_L2PF)
  int zface_contig;      // stage_p7: exterior z-face values are 64 consecutive nodes per face (bulk copies instead of gathers)
};

struct HaloParams {
  double* q[NVAR];
  double* dp;            // DPRES travels with the prognostic variables
  const int* src;        // (Nhalo) interior flat index feeding each halo slot (same-rank faces), -1 = remote
  const int* vmapB;      // (Nhalo) own face node of each halo slot
  const double *gsqrt, *g13, *g23, *gsqrtH;
  const int* emap2d;
  int face_off[7];
  int bc[6];
  int Np, Ne, Nfp, np, Nhalo, terrain;
};

// vertical-implicit column solve (vi_solver.cu)
struct VIParams {
  const double* qcur[NVAR];   // state entering the stage (DDENS_, MOMX_, ...)
  const double* q0[NVAR];     // var0: state at the start of the step (Newton linearisation point)
  double* kim[NVAR];          // out: implicit tendency of the stage
  double* qout[NVAR];         // out: qcur + impl_fac * kim (may alias qcur when qcur != q0)
  double* dpout;              // out: DPRES of qout
  const double *dens_hyd, *pres_hyd, *therm_hyd, *rhot_hyd_vi, *rtot, *cvtot, *cptot;
  const double *escale, *fscale;
  const ElemTables* tab;
  double* scratch;            // NeZ * 120 * (Ne2D*64) doubles
  PhysConst c;
  double impl_fac;
  int Ne, Ne2D, NeZ;
  int exact_pow;              // pow() instead of exp(e log x) for the equation of state
  const ElemTables* htab;     // HOST copy of the operator tables (the second kernel derives its constant-memory tables from it)
  // terrain-following mesh (gsqrt != nullptr): Gsqrt, G13, G23 per node, GsqrtH per 2D node; MOMX', MOMY' after their implicit solve
  // (written by pass 0 into pvu_out / pvv_out, read by pass 1 through pvu / pvv)
  const double *gsqrt, *g13, *g23, *gsqrtH;
  const double *pvu, *pvv;
  double *pvu_out, *pvv_out;
  int pass0;
};
constexpr int MAXTERM = 20;   // 2 * stages of the largest IMEX scheme supported
struct LinCombParams {
  const double* base[NVAR];
  double* out[NVAR];
  const double* k[MAXTERM][NVAR];
  double coef[MAXTERM];
  int nterm;
  size_t n;
};
// tracer advection with a prescribed mass flux (tracer.cu; row f4)
struct TracerParams {
  const double* q;             // stage input (Np*Ne + halo), halo filled
  double* qout;
  const double *mfx, *mfy, *mfz;          // mass flux = momentum of the dynamical state (halo + boundary condition applied)
  const double *ddens, *ddens0, *dens_hyd;
  double *alphM, *alphP;       // (NfpTot, Ne)
  double* fct;                 // FCT coefficient (Np*Ne + halo)
  const double* rhoq_tp;       // physics tendency of rho*q, may be NULL
  double *var0, *vartmp;       // work arrays of the tracer integrator (Np*Ne)
  const double *escale, *fscale, *jac, *w3;
  const int* vmapP;
  const ElemTables* tab;
  const double* filt;          // tracer modal filter [2][np*np] row-major: horizontal, vertical
  double w1d[MAXNP];           // 1D LGL weights (surface quadrature of FaceIntMat)
  double sig_ss, gam_ss, sig_Ns, gam_Ns, c_ssm1, c_ss, dttmp;
  int stage, nstage, upd_vartmp;
  int disable_limiter, do_filter, do_tmar;
  int Np, Nfp, NfpTot, np, Ne;
};
cudaError_t launch_trc_save_massflux(const double* const prog[NVAR], const double* dens_hyd, const double* pres_hyd, const double* dpres,
                                     const int* vmapP, double* const mflx[3], double* alphM, double* alphP, double w_h, double w_v, double gamm,
                                     bool hevi, bool first, int Np, int Nfp, int NfpTot, int np, int Ne, cudaStream_t s);
cudaError_t launch_trc_rescale(double* q, const double* dens_hyd, const double* dd_trc, const double* ddens, size_t n, cudaStream_t s);
cudaError_t launch_trc_alphdens(const TracerParams& P, cudaStream_t s);
cudaError_t launch_trc_fct(const TracerParams& P, cudaStream_t s);
cudaError_t launch_trc_stage(const TracerParams& P, cudaStream_t s);

// sample/advect3d stage (advect3d.cu)
struct AdvectParams {
  const double *q, *u, *v, *w;   // stage input (Np*Ne + Nhalo), halo filled
  double* qout;
  const double* q0;
  double* vt;
  double* tend_out;              // when non-null: write dqdt instead of the updated q
  const double* ellval[4];       // Dx, Dy, Dz, Lift in ELL storage (slot-major, M = Np)
  const int* ellcol[4];          // 0-based columns
  int colsz[4];
  const double* escale;          // [3][Ne]
  const double* fscale;          // [6][Ne]
  const int* vmapP;              // (NfpTot,Ne) 0-based
  RKStage rk;
  int Np, Nfp, NfpTot, np, Ne, epb;
};
size_t advect_smem_bytes(const AdvectParams& P);
cudaError_t launch_advect_stage(const AdvectParams& P, cudaStream_t s);
void launch_advect_halo(double* q, double* u, double* v, double* w, const int* src, size_t nint, int nhalo, bool with_vel,
                        cudaStream_t s);
cudaError_t launch_sparsemat_general(int M, int nq, int col_size, const double* val, const int* col, const int* rowptr, const double* b1,
                                     const double* b2, double* c, size_t sb_col, size_t sb_q, size_t sc_row, size_t sc_q, cudaStream_t s);
cudaError_t launch_ell_spmv(int M, int N, int col_size, const double* val, const int* col, const double* b, double* c, int nvec,
                            cudaStream_t s);

// numerical diffusion (numdiff.cu): one half-step of the local-DG Laplacian per launch
struct NumdiffParams {
  const double *in0, *in1, *in2;   // FLX: Varh, Varv (fields incl. halo);  LAP / TEND: Gx, Gy, Gz
  double *out0, *out1, *out2;      // FLX: Gx, Gy, Gz;  LAP: lapla_h, lapla_v
  double* var;                     // TEND: var += dt * tend
  const double *ddens, *dens_hyd;
  const ElemTables* tab;
  const double *escale, *fscale;
  const int* vmapP;
  int vel_bc[6], therm_bc[6];      // per tile face, 0 where the face is not a physical boundary
  int face_off[7];
  int varid, dens_flag, bc_on_v;
  double coef_h, coef_v, dt;
  int Np, Nfp, NfpTot, np, Ne;
  size_t nint;
};
// per-variable arguments of a half-step (the five prognostic variables go through the same kernel)
struct NdVar {
  const double *in0, *in1, *in2;
  double *out0, *out1, *out2, *var;
  int varid, dens_flag;
};
constexpr int ND_MAXVAR = 5;
struct NumdiffMulti {
  NumdiffParams P;           // everything the variables share (in0 .. var, varid, dens_flag of P are not used)
  NdVar v[ND_MAXVAR];
  int nvar;
};
void launch_numdiff(int mode, const NumdiffParams& P, cudaStream_t s);
void launch_numdiff_multi(int mode, const NumdiffMulti& M, cudaStream_t s);

void launch_vi(const VIParams& p, bool moist, cudaStream_t s);
bool launch_vi2(const VIParams& p, const ElemTables& tab, bool moist, cudaStream_t s);
void launch_lincomb(const LinCombParams& L, cudaStream_t s);
void launch_lincomb_filter(const LinCombParams& L, const ElemTables* tab, const double* gsqrt, bool weighted, int Ne, int np,
                           cudaStream_t s);
void launch_modal_filter5(double* const q[NVAR], const double* gsqrt, bool terrain, int Ne, int np, cudaStream_t s);

// halo exchange over NCCL (halo_comm.cu)
struct RemoteFace {
  int f = 0, peer = 0, peer_face = 0, off = 0, cnt = 0;   // own face id, neighbour rank, its face id, halo offset / node count
  double* sendbuf = nullptr;                               // [6][cnt]
  double* recvbuf = nullptr;                               // [6][cnt]
};
// Direct peer-memory exchange (default when the ranks can map each other's memory; FEDG_HALO=nccl keeps NCCL send/recv): every
// rank owns one receive area [flags | parity 0: faces | parity 1: faces]; a face is six fields of cnt nodes.  The neighbour's pack
// kernel stores the face values straight into that area over NVLink and then raises the face's flag to the exchange number; the
// receiver's unpack kernel spins on its own flag.  Two parities: a sender may run one exchange ahead of the receiver, never two
// (it needs the receiver's data of exchange n before it can pack n + 1, and the receiver packs n only after unpacking n - 1).
struct PeerHalo {
  bool on = false;
  unsigned char* area = nullptr;                 // own receive area (cudaMalloc, exported with cudaIpcGetMemHandle)
  size_t face_off[2][6] = {{0}};                 // byte offset of (parity, own face id) inside the area
  void* peer_base[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // mapped area of the peer of remote face i (shared between faces of one peer)
  bool peer_owner[6] = {false, false, false, false, false, false};               // this entry opened the handle (closes it)
  double* dst[2][6] = {{nullptr}};               // where the data of own remote face i lands on the peer
  unsigned long long* dst_flag[2][6] = {{nullptr}};
  unsigned long long seq = 0;                    // exchanges started so far
  double* cur_q[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};     // fields of the exchange in flight (start -> wait)
  size_t cur_nint = 0;
};
struct CommState {
  bool active = false;
  void* comm = nullptr;          // ncclComm_t
  int rank = 0, nranks = 1, nremote = 0;
  RemoteFace face[6];
  int recv_order[6] = {0, 1, 2, 3, 4, 5};
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_packed = nullptr, ev_done = nullptr;
  PeerHalo p2p;
};
int comm_unique_id(void* id128, std::string& err);
int comm_init(CommState& cs, const void* id128, int rank, int nranks, const int nbr_rank[6], const int nbr_face[6], const int face_off[7],
              std::string& err);
void comm_destroy(CommState& cs);
int comm_exchange_start(CommState& cs, double* const q[NVAR], double* dp, const int* d_vmapB, size_t nint, cudaStream_t compute,
                        std::string& err);
void comm_exchange_wait(CommState& cs, cudaStream_t compute);
int comm_allreduce_max(CommState& cs, double* d_inout, int n, cudaStream_t s, std::string& err);
struct P2PMsg { int peer; int msg_id; double* buf; size_t count; };
int comm_p2p_group(CommState& cs, std::vector<P2PMsg>& sends, std::vector<P2PMsg>& recvs, cudaStream_t s, std::string& err);
int comm_allreduce_sum(CommState& cs, double* d_inout, int n, cudaStream_t s, std::string& err);

void upload_tables(const ElemTables& t, cudaStream_t s);
void launch_stage(const StageParams& p, int np, bool terrain, bool moist, bool hevi, cudaStream_t s);
void launch_halo_fill(const HaloParams& p, cudaStream_t s);
void launch_calc_pres(const double* drhot, const double* pres_hyd, const double* therm_hyd, const double* rtot,
                      const double* cvtot, const double* cptot, bool moist, PhysConst c, double* pres, double* dpres,
                      long n, cudaStream_t s);
void launch_calc_rhot_hyd(const double* pres_hyd, PhysConst c, double* therm_hyd, long n, cudaStream_t s);
void launch_monitor(const double* const q[NVAR], const double* dens_hyd, const double* pres, const double* rtot,
                    bool moist, const double* w3, const double* Jac, const double* gsqrt, bool terrain,
                    const double* zlev, PhysConst c, int Np, int Ne, double* out5, cudaStream_t s);
void launch_elem_op(int op, const double* in, double* out, int nelem, int np, cudaStream_t s);
void launch_phyd_hgrad(const double* pres_hyd, const double* pres_ref, const double* gsqrt, const double* g13, const double* g23,
                       const double* gsqrtH, const double* escale, const double* fscale, const int* vmapP, const int* emap2d,
                       double* outx, double* outy, int np, int Ne, bool terrain, cudaStream_t s);

}  // namespace fedg
