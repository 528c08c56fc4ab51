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
// libnccl is bound at run time (dlopen): under Python the process already holds torch's bundled libnccl.so.2 and a
// second copy must not be loaded; under the Fortran driver the system libnccl.so.2 is found on the loader path.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
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
    FEDG_SYM(AllReduce, "ncclAllReduce") FEDG_SYM(GetErrorString, "ncclGetErrorString")
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
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  ncclResult_t r = g_nccl.CommInitRank(&comm, nranks, id, rank);
  if (r != ncclSuccess) { err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); return FEDG_ERR_COMM; }
  cs.comm = comm; cs.rank = rank; cs.nranks = nranks;
  cs.nremote = 0;
  for (int f = 0; f < 6; ++f) {
    if (nbr_rank[f] == rank) continue;
    if (nbr_rank[f] < 0 || nbr_rank[f] >= nranks) { err = "nbr_rank out of range"; return FEDG_ERR_ARG; }
    RemoteFace& rf = cs.face[cs.nremote++];
    rf.f = f; rf.peer = nbr_rank[f]; rf.peer_face = nbr_face[f]; rf.off = face_off[f]; rf.cnt = face_off[f + 1] - face_off[f];
    if (cudaMalloc(&rf.sendbuf, size_t(6) * rf.cnt * sizeof(double)) != cudaSuccess) { err = "cudaMalloc(sendbuf)"; return FEDG_ERR_CUDA; }
    if (cudaMalloc(&rf.recvbuf, size_t(6) * rf.cnt * sizeof(double)) != cudaSuccess) { err = "cudaMalloc(recvbuf)"; return FEDG_ERR_CUDA; }
  }
  // receive order: ascending face id of the sender
  for (int i = 0; i < cs.nremote; ++i) cs.recv_order[i] = i;
  std::sort(cs.recv_order, cs.recv_order + cs.nremote, [&](int a, int b) { return cs.face[a].peer_face < cs.face[b].peer_face; });
  {  // highest priority: the few blocks of the exchange must not queue behind the interior-element grid
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&cs.stream, cudaStreamNonBlocking, hi) != cudaSuccess) { err = "cudaStreamCreate(comm)"; return FEDG_ERR_CUDA; }
  }
  cudaEventCreateWithFlags(&cs.ev_packed, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&cs.ev_done, cudaEventDisableTiming);
  cs.active = true;
  return FEDG_OK;
}

void comm_destroy(CommState& cs) {
  if (!cs.active) return;
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
int comm_allreduce_sum(CommState& cs, double* d_inout, int n, cudaStream_t s, std::string& err) {
  if (!cs.active) return FEDG_OK;
  ncclResult_t r = g_nccl.AllReduce(d_inout, d_inout, size_t(n), ncclDouble, ncclSum, static_cast<ncclComm_t>(cs.comm), s);
  if (r != ncclSuccess) { err = std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r); return FEDG_ERR_COMM; }
  return FEDG_OK;
}

}  // namespace fedg
