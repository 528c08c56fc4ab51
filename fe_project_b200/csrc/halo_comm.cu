// Face-halo exchange between tiles on different GPUs: NCCL send/recv over NVLink (replaces MeshFieldCommBase
// Put / Exchange / Get over MPI, FElib/src/data/scale_meshfieldcomm_base.F90:58-139, 870-884).
//
// Per exchange: ONE pack kernel gathers, for every remote tile face, the face nodes of the six travelling fields (five
// prognostic variables + DPRES) through VMapB into that face's contiguous send buffer (extract_bounddata,
// scale_meshfieldcomm_base.F90:617-687); ONE ncclSend / ncclRecv pair per face ships the six fields as a single message
// (the first version sent every field separately and wrote straight into the halo slots: 12 point-to-point operations
// per face, whose latency -- about 60 us per face and stage at 8 GPUs -- was not hidden); ONE unpack kernel on the
// communication stream scatters the receive buffers into the halo slots (set_bounddata, :690-760).  The group runs on
// a dedicated high-priority stream while the interior elements are processed (HIDE_MPI_COMM_FLAG semantics,
// driver_nonhydro3d.F90:859-895).
//
// NCCL point-to-point operations carry no tags (the reference tags messages with 10*tileID+faceID): messages between
// one pair of ranks match in posting order.  Sends are posted in ascending order of the sender's face id, receives in
// ascending order of the SENDER's face id (= nbr_face of the receiving face), which makes both sides agree.
//
// Default data path since round 2: DIRECT PEER MEMORY (PeerHalo in fedg_internal.h) -- the pack kernel of the sender stores the
// face values into the receiver's halo staging area over NVLink (the areas are exchanged once as CUDA IPC handles, shipped with
// ncclAllGather), a flag per face carries the exchange number, the receiver's unpack kernel spins on it.  No NCCL call, no second
// stream and no host work on the hot path: pack -> interior elements -> wait + unpack -> boundary elements, all on the compute
// stream.  (Round 1 measured the NCCL version at 0.62 weak-scaling efficiency on 8 GPUs: 1 ms per step of host enqueue latency
// of the grouped point-to-point calls.)  NCCL send/recv stays as the fallback when the areas cannot be mapped, and with
// FEDG_HALO=nccl for A/B runs.
//
// libnccl is bound at run time (dlopen): under Python the process already holds torch's bundled libnccl.so.2 and a
// second copy must not be loaded; under the Fortran driver the system libnccl.so.2 is found on the loader path.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fedg_internal.h"

namespace fedg {

namespace {
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string& err) {
    if (h) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define FEDG_SYM(field, sym)                                              \
  field = reinterpret_cast<decltype(field)>(dlsym(h, sym));               \
  if (!field) { err = std::string("libnccl lacks ") + sym; return false; }
    FEDG_SYM(GetUniqueId, "ncclGetUniqueId") FEDG_SYM(CommInitRank, "ncclCommInitRank") FEDG_SYM(CommDestroy, "ncclCommDestroy")
    FEDG_SYM(GroupStart, "ncclGroupStart") FEDG_SYM(GroupEnd, "ncclGroupEnd") FEDG_SYM(Send, "ncclSend") FEDG_SYM(Recv, "ncclRecv")
    FEDG_SYM(AllReduce, "ncclAllReduce") FEDG_SYM(AllGather, "ncclAllGather") FEDG_SYM(GetErrorString, "ncclGetErrorString")
#undef FEDG_SYM
    return true;
  }
};
NcclApi g_nccl;

struct FaceSet {            // the remote faces of a tile (at most four lateral ones)
  double* buf[6];
  int off[6], cnt[6], start[7];   // halo offset / node count per face, prefix sum of cnt
  int n;
};
struct SixFields { double* f[6]; };

// gather the face nodes of six fields into sendbuf[face][field][m]
__global__ void pack_faces_kernel(SixFields q, const int* __restrict__ vmapB, FaceSet fs) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= fs.start[fs.n]) return;
  int i = 0;
  while (g >= fs.start[i + 1]) ++i;
  const int m = g - fs.start[i], cnt = fs.cnt[i];
  const int src = vmapB[fs.off[i] + m];
  double* buf = fs.buf[i];
#pragma unroll
  for (int v = 0; v < 6; ++v) buf[size_t(v) * cnt + m] = q.f[v][src];
}
// scatter recvbuf[face][field][m] into the halo slots of the six fields
__global__ void unpack_faces_kernel(SixFields q, size_t nint, FaceSet fs) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= fs.start[fs.n]) return;
  int i = 0;
  while (g >= fs.start[i + 1]) ++i;
  const int m = g - fs.start[i], cnt = fs.cnt[i];
  const double* buf = fs.buf[i];
#pragma unroll
  for (int v = 0; v < 6; ++v) q.f[v][nint + fs.off[i] + m] = buf[size_t(v) * cnt + m];
}

// ---- direct peer-memory exchange ----------------------------------------------------------------------------------------
struct PeerFaceSet {
  double* dst[6];                      // face data on the PEER (pack) / in the own area (unpack)
  unsigned long long* flag[6];         // flag on the peer (signal) / own flag (wait)
  int off[6], cnt[6], start[7];
  int n;
};
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Gather the face nodes of six fields and store them into the neighbour's receive area (remote stores over NVLink); the last block to
// finish raises the flags.  Ordering: every block synchronises, then ONE thread per block issues the system-scope fence (cumulative
// over the block's stores through the barrier) before it counts the block as done; the last block fences again and publishes the
// exchange number with release stores.
__global__ void pack_peer_kernel(SixFields q, const int* __restrict__ vmapB, PeerFaceSet fs, unsigned long long seq, unsigned int* done_ctr) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < fs.start[fs.n]) {
    int i = 0;
    while (g >= fs.start[i + 1]) ++i;
    const int m = g - fs.start[i], cnt = fs.cnt[i];
    const int src = vmapB[fs.off[i] + m];
    double* buf = fs.dst[i];
#pragma unroll
    for (int v = 0; v < 6; ++v) buf[size_t(v) * cnt + m] = q.f[v][src];
  }
  __shared__ int last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = (atomicAdd(done_ctr, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (last && threadIdx.x < fs.n) {
    __threadfence_system();
    if (threadIdx.x == 0) *done_ctr = 0u;          // next exchange (stream-ordered after this kernel)
    st_release_sys(fs.flag[threadIdx.x], seq);
  }
}
// wait until the neighbour's data of exchange `seq` has landed, then scatter it into the halo slots of the six fields
__global__ void wait_unpack_peer_kernel(SixFields q, size_t nint, PeerFaceSet fs, unsigned long long seq, long long timeout_clk) {
  __shared__ int ok;
  if (threadIdx.x == 0) {
    // the faces this block touches: [first, last] node of the block
    const int g0 = blockIdx.x * blockDim.x, g1 = min(g0 + (int)blockDim.x, fs.start[fs.n]) - 1;
    int done = 1;
    const long long t0 = clock64();
    for (int i = 0; i < fs.n; ++i) {
      if (fs.start[i + 1] <= g0 || fs.start[i] > g1) continue;
      while (ld_acquire_sys(fs.flag[i]) < seq) {
        if (clock64() - t0 > timeout_clk) { done = 0; break; }
        __nanosleep(200);
      }
    }
    ok = done;
  }
  __syncthreads();
  if (!ok) { asm volatile("trap;"); }   // a neighbour never delivered: abort the launch instead of hanging the device
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= fs.start[fs.n]) return;
  int i = 0;
  while (g >= fs.start[i + 1]) ++i;
  const int m = g - fs.start[i], cnt = fs.cnt[i];
  const double* buf = fs.dst[i];
#pragma unroll
  for (int v = 0; v < 6; ++v) q.f[v][nint + fs.off[i] + m] = __ldcv(buf + size_t(v) * cnt + m);   // written by another GPU: bypass L1
}

struct PeerTable {                     // what a rank publishes about its receive area
  cudaIpcMemHandle_t handle;
  unsigned long long face_off[2][6];   // byte offsets, ~0 = face not remote
  int valid;
  int pad;
};

// Set up the direct path; any failure leaves cs.p2p.on = false (NCCL fallback).  Collective over the communicator.
void peer_halo_init(CommState& cs, int device) {
  PeerHalo& ph = cs.p2p;
  ph = PeerHalo{};
  { const char* e = getenv("FEDG_HALO"); if (e && std::strcmp(e, "nccl") == 0) return; }
  ncclComm_t comm = static_cast<ncclComm_t>(cs.comm);
  // own area: 256 B header (12 flags of 8 B; the pack kernel's block counter at byte 192), then two parities of the remote faces
  PeerTable mine{};
  size_t bytes = 256;
  for (int par = 0; par < 2; ++par)
    for (int f = 0; f < 6; ++f) mine.face_off[par][f] = ~0ull;
  for (int par = 0; par < 2; ++par)
    for (int i = 0; i < cs.nremote; ++i) {
      const RemoteFace& rf = cs.face[i];
      mine.face_off[par][rf.f] = bytes; ph.face_off[par][rf.f] = bytes;
      bytes += (size_t(6) * rf.cnt * sizeof(double) + 255) / 256 * 256;
    }
  bool ok = cudaMalloc(&ph.area, bytes) == cudaSuccess && cudaMemset(ph.area, 0, bytes) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
  if (ok) ok = cudaIpcGetMemHandle(&mine.handle, ph.area) == cudaSuccess;
  mine.valid = ok ? 1 : 0;
  if (!ok) cudaGetLastError();
  // all ranks take part in the gather, whatever their local outcome
  PeerTable* d_tab = nullptr;
  std::vector<PeerTable> all(cs.nranks);
  bool gathered = cudaMalloc(&d_tab, sizeof(PeerTable) * (cs.nranks + 1)) == cudaSuccess &&
                  cudaMemcpy(d_tab + cs.nranks, &mine, sizeof(PeerTable), cudaMemcpyHostToDevice) == cudaSuccess;
  if (gathered) gathered = g_nccl.AllGather(d_tab + cs.nranks, d_tab, sizeof(PeerTable), ncclChar, comm, cs.stream) == ncclSuccess &&
                           cudaStreamSynchronize(cs.stream) == cudaSuccess &&
                           cudaMemcpy(all.data(), d_tab, sizeof(PeerTable) * cs.nranks, cudaMemcpyDeviceToHost) == cudaSuccess;
  if (d_tab) cudaFree(d_tab);
  bool all_ok = gathered;
  for (int r = 0; r < cs.nranks && all_ok; ++r) if (!all[r].valid) all_ok = false;
  // map the areas of the neighbours
  for (int i = 0; i < cs.nremote && all_ok; ++i) {
    const RemoteFace& rf = cs.face[i];
    for (int k = 0; k < i; ++k) if (cs.face[k].peer == rf.peer) { ph.peer_base[i] = ph.peer_base[k]; break; }
    if (!ph.peer_base[i]) {
      if (cudaIpcOpenMemHandle(&ph.peer_base[i], all[rf.peer].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError(); ph.peer_base[i] = nullptr; all_ok = false; break;
      }
      ph.peer_owner[i] = true;
    }
    for (int par = 0; par < 2; ++par) {
      const unsigned long long off = all[rf.peer].face_off[par][rf.peer_face];
      if (off == ~0ull) { all_ok = false; break; }
      ph.dst[par][i] = reinterpret_cast<double*>(static_cast<unsigned char*>(ph.peer_base[i]) + off);
      ph.dst_flag[par][i] = reinterpret_cast<unsigned long long*>(ph.peer_base[i]) + (par * 6 + rf.peer_face);
    }
  }
  // agreement: the direct path is used only if EVERY rank could map its neighbours (a mixed set-up would deadlock)
  double flag = all_ok ? 0.0 : 1.0, *d_flag = nullptr;
  bool agreed = false;
  if (cudaMalloc(&d_flag, sizeof(double)) == cudaSuccess && cudaMemcpy(d_flag, &flag, sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
      g_nccl.AllReduce(d_flag, d_flag, 1, ncclDouble, ncclMax, comm, cs.stream) == ncclSuccess && cudaStreamSynchronize(cs.stream) == cudaSuccess &&
      cudaMemcpy(&flag, d_flag, sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess)
    agreed = (flag == 0.0);
  if (d_flag) cudaFree(d_flag);
  ph.on = agreed;
  (void)device;
  if (getenv("FEDG_HALO_VERBOSE")) fprintf(stderr, "[fedg] rank %d: halo exchange over %s\n", cs.rank, ph.on ? "direct peer memory" : "NCCL send/recv");
}

void peer_halo_destroy(CommState& cs) {
  PeerHalo& ph = cs.p2p;
  for (int i = 0; i < 6; ++i) if (ph.peer_owner[i] && ph.peer_base[i]) cudaIpcCloseMemHandle(ph.peer_base[i]);
  if (ph.area) cudaFree(ph.area);
  ph = PeerHalo{};
}
}  // namespace

int comm_unique_id(void* id128, std::string& err) {
  if (!g_nccl.load(err)) return FEDG_ERR_COMM;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclResult_t r = g_nccl.GetUniqueId(static_cast<ncclUniqueId*>(id128));
  if (r != ncclSuccess) { err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return FEDG_ERR_COMM; }
  return FEDG_OK;
}

int comm_init(CommState& cs, const void* id128, int rank, int nranks, const int nbr_rank[6], const int nbr_face[6], const int face_off[7],
              std::string& err) {
  if (!g_nccl.load(err)) return FEDG_ERR_COMM;
  for (int f = 0; f < 6; ++f)
    if (nbr_rank[f] < 0 || nbr_rank[f] >= nranks) { err = "nbr_rank out of range"; return FEDG_ERR_ARG; }
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  ncclResult_t r = g_nccl.CommInitRank(&comm, nranks, id, rank);
  if (r != ncclSuccess) { err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); return FEDG_ERR_COMM; }
  cs.comm = comm; cs.rank = rank; cs.nranks = nranks;
  cs.active = true;          // from here on comm_destroy releases whatever has been created
  cs.nremote = 0;
  auto bail = [&](int code, const char* what) { err = what; comm_destroy(cs); return code; };
  for (int f = 0; f < 6; ++f) {
    if (nbr_rank[f] == rank) continue;
    RemoteFace& rf = cs.face[cs.nremote++];
    rf.f = f; rf.peer = nbr_rank[f]; rf.peer_face = nbr_face[f]; rf.off = face_off[f]; rf.cnt = face_off[f + 1] - face_off[f];
    if (cudaMalloc(&rf.sendbuf, size_t(6) * rf.cnt * sizeof(double)) != cudaSuccess) return bail(FEDG_ERR_CUDA, "cudaMalloc(sendbuf)");
    if (cudaMalloc(&rf.recvbuf, size_t(6) * rf.cnt * sizeof(double)) != cudaSuccess) return bail(FEDG_ERR_CUDA, "cudaMalloc(recvbuf)");
  }
  // receive order: ascending face id of the sender
  for (int i = 0; i < cs.nremote; ++i) cs.recv_order[i] = i;
  std::sort(cs.recv_order, cs.recv_order + cs.nremote, [&](int a, int b) { return cs.face[a].peer_face < cs.face[b].peer_face; });
  {  // highest priority: the few blocks of the exchange must not queue behind the interior-element grid
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&cs.stream, cudaStreamNonBlocking, hi) != cudaSuccess) return bail(FEDG_ERR_CUDA, "cudaStreamCreate(comm)");
  }
  if (cudaEventCreateWithFlags(&cs.ev_packed, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&cs.ev_done, cudaEventDisableTiming) != cudaSuccess)
    return bail(FEDG_ERR_CUDA, "cudaEventCreate(comm)");
  int dev = 0;
  cudaGetDevice(&dev);
  peer_halo_init(cs, dev);
  return FEDG_OK;
}

void comm_destroy(CommState& cs) {
  if (!cs.active) return;
  peer_halo_destroy(cs);
  for (int i = 0; i < cs.nremote; ++i) { if (cs.face[i].sendbuf) cudaFree(cs.face[i].sendbuf); if (cs.face[i].recvbuf) cudaFree(cs.face[i].recvbuf); }
  if (cs.comm) g_nccl.CommDestroy(static_cast<ncclComm_t>(cs.comm));
  if (cs.stream) cudaStreamDestroy(cs.stream);
  if (cs.ev_packed) cudaEventDestroy(cs.ev_packed);
  if (cs.ev_done) cudaEventDestroy(cs.ev_done);
  cs = CommState{};
}

// Put + Exchange: pack on the compute stream, ship on the communication stream.  q[5] + dp are field arrays (interior
// followed by the halo slots); nint = Np*Ne.
int comm_exchange_start(CommState& cs, double* const q[NVAR], double* dp, const int* d_vmapB, size_t nint, cudaStream_t compute,
                        std::string& err) {
  if (!cs.active || cs.nremote == 0) return FEDG_OK;
  if (cs.p2p.on) {
    PeerHalo& ph = cs.p2p;
    const unsigned long long seq = ++ph.seq;
    const int par = int(seq & 1ull);
    PeerFaceSet fs{};
    fs.n = cs.nremote;
    for (int i = 0; i < cs.nremote; ++i) {
      const RemoteFace& rf = cs.face[i];
      fs.dst[i] = ph.dst[par][i]; fs.flag[i] = ph.dst_flag[par][i];
      fs.off[i] = rf.off; fs.cnt[i] = rf.cnt; fs.start[i + 1] = fs.start[i] + rf.cnt;
    }
    SixFields six{};
    for (int v = 0; v < NVAR; ++v) { six.f[v] = q[v]; ph.cur_q[v] = q[v]; }
    six.f[5] = dp; ph.cur_q[5] = dp; ph.cur_nint = nint;
    const int ntot = fs.start[cs.nremote];
    pack_peer_kernel<<<(ntot + 255) / 256, 256, 0, compute>>>(six, d_vmapB, fs, seq, reinterpret_cast<unsigned int*>(ph.area + 192));
    return FEDG_OK;
  }
  FaceSet snd{}, rcv{};
  snd.n = rcv.n = cs.nremote;
  for (int i = 0; i < cs.nremote; ++i) {
    const RemoteFace& rf = cs.face[i];
    snd.buf[i] = rf.sendbuf; rcv.buf[i] = rf.recvbuf;
    snd.off[i] = rcv.off[i] = rf.off; snd.cnt[i] = rcv.cnt[i] = rf.cnt;
    snd.start[i + 1] = rcv.start[i + 1] = snd.start[i] + rf.cnt;
  }
  SixFields six{};
  for (int v = 0; v < NVAR; ++v) six.f[v] = q[v];
  six.f[5] = dp;
  const int ntot = snd.start[cs.nremote];
  pack_faces_kernel<<<(ntot + 255) / 256, 256, 0, compute>>>(six, d_vmapB, snd);
  cudaEventRecord(cs.ev_packed, compute);
  cudaStreamWaitEvent(cs.stream, cs.ev_packed, 0);
  ncclComm_t comm = static_cast<ncclComm_t>(cs.comm);
  ncclResult_t r = g_nccl.GroupStart();
  for (int i = 0; i < cs.nremote && r == ncclSuccess; ++i) {           // ascending own face id; one message per face
    const RemoteFace& rf = cs.face[i];
    r = g_nccl.Send(rf.sendbuf, size_t(6) * rf.cnt, ncclDouble, rf.peer, comm, cs.stream);
  }
  for (int oi = 0; oi < cs.nremote && r == ncclSuccess; ++oi) {        // ascending face id of the sender
    const RemoteFace& rf = cs.face[cs.recv_order[oi]];
    r = g_nccl.Recv(rf.recvbuf, size_t(6) * rf.cnt, ncclDouble, rf.peer, comm, cs.stream);
  }
  ncclResult_t r2 = g_nccl.GroupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess) { err = std::string("NCCL halo exchange: ") + g_nccl.GetErrorString(r); return FEDG_ERR_COMM; }
  unpack_faces_kernel<<<(ntot + 255) / 256, 256, 0, cs.stream>>>(six, nint, rcv);
  cudaEventRecord(cs.ev_done, cs.stream);
  return FEDG_OK;
}

// Get: the compute stream waits for the halo data
void comm_exchange_wait(CommState& cs, cudaStream_t compute) {
  if (!cs.active || cs.nremote == 0) return;
  if (cs.p2p.on) {
    PeerHalo& ph = cs.p2p;
    const unsigned long long seq = ph.seq;
    const int par = int(seq & 1ull);
    PeerFaceSet fs{};
    fs.n = cs.nremote;
    for (int i = 0; i < cs.nremote; ++i) {
      const RemoteFace& rf = cs.face[i];
      fs.dst[i] = reinterpret_cast<double*>(ph.area + ph.face_off[par][rf.f]);
      fs.flag[i] = reinterpret_cast<unsigned long long*>(ph.area) + (par * 6 + rf.f);
      fs.off[i] = rf.off; fs.cnt[i] = rf.cnt; fs.start[i + 1] = fs.start[i] + rf.cnt;
    }
    SixFields six{};
    for (int v = 0; v < 6; ++v) six.f[v] = ph.cur_q[v];
    const int ntot = fs.start[cs.nremote];
    static long long timeout_clk = 0;
    if (!timeout_clk) { const char* e = getenv("FEDG_HALO_TIMEOUT_S"); const double sec = e ? atof(e) : 20.0; timeout_clk = (long long)(sec * 1.9e9); }
    wait_unpack_peer_kernel<<<(ntot + 255) / 256, 256, 0, compute>>>(six, ph.cur_nint, fs, seq, timeout_clk);
    return;
  }
  cudaStreamWaitEvent(compute, cs.ev_done, 0);
}

// Messages between local meshes on different ranks (cubed-sphere panel edges, fedg_link_halo_send / _recv): one NCCL group
// on stream s.  Point-to-point operations carry no tags, so both sides post the messages of a rank pair in ascending msg_id
// (the id of the receiving (panel, face), the same number on both ranks).
int comm_p2p_group(CommState& cs, std::vector<P2PMsg>& sends, std::vector<P2PMsg>& recvs, cudaStream_t s, std::string& err) {
  if (!cs.active) { err = "no communicator: call fedg_comm_init on the first mesh of the group"; return FEDG_ERR_STATE; }
  auto by_id = [](const P2PMsg& a, const P2PMsg& b) { return a.peer != b.peer ? a.peer < b.peer : a.msg_id < b.msg_id; };
  std::sort(sends.begin(), sends.end(), by_id);
  std::sort(recvs.begin(), recvs.end(), by_id);
  ncclComm_t comm = static_cast<ncclComm_t>(cs.comm);
  ncclResult_t r = g_nccl.GroupStart();
  for (size_t i = 0; i < sends.size() && r == ncclSuccess; ++i) r = g_nccl.Send(sends[i].buf, sends[i].count, ncclDouble, sends[i].peer, comm, s);
  for (size_t i = 0; i < recvs.size() && r == ncclSuccess; ++i) r = g_nccl.Recv(recvs[i].buf, recvs[i].count, ncclDouble, recvs[i].peer, comm, s);
  ncclResult_t r2 = g_nccl.GroupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess) { err = std::string("NCCL panel-edge exchange: ") + g_nccl.GetErrorString(r); return FEDG_ERR_COMM; }
  return FEDG_OK;
}

// global sums for the monitors (MPI_Allreduce in file/scale_file_monitor_meshfield.F90:203-211)
int comm_allreduce_max(CommState& cs, double* d_inout, int n, cudaStream_t s, std::string& err) {
  if (!cs.active) return FEDG_OK;
  ncclResult_t r = g_nccl.AllReduce(d_inout, d_inout, size_t(n), ncclDouble, ncclMax, static_cast<ncclComm_t>(cs.comm), s);
  if (r != ncclSuccess) { err = std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r); return FEDG_ERR_COMM; }
  return FEDG_OK;
}
int comm_allreduce_sum(CommState& cs, double* d_inout, int n, cudaStream_t s, std::string& err) {
  if (!cs.active) return FEDG_OK;
  ncclResult_t r = g_nccl.AllReduce(d_inout, d_inout, size_t(n), ncclDouble, ncclSum, static_cast<ncclComm_t>(cs.comm), s);
  if (r != ncclSuccess) { err = std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r); return FEDG_ERR_COMM; }
  return FEDG_OK;
}

}  // namespace fedg
