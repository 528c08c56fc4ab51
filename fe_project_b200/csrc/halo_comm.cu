// Face-halo exchange between tiles on different GPUs: NCCL send/recv over NVLink (replaces MeshFieldCommBase
// Put / Exchange / Get over MPI, FElib/src/data/scale_meshfieldcomm_base.F90:58-139, 870-884).
//
// Per exchange and remote tile face: one pack kernel gathers the face nodes of the six travelling fields (five
// prognostic variables + DPRES) through VMapB into a contiguous device buffer (extract_bounddata,
// scale_meshfieldcomm_base.F90:617-687); ncclSend ships it; the matching ncclRecv writes straight into the halo slots
// of the receiver's field arrays (halo slots of a tile face are contiguous, scale_meshutil_3d.F90:570-602), so there
// is no unpack kernel.  All sends and receives of one exchange form one NCCL group on a dedicated stream so that the
// interior elements can be processed meanwhile (HIDE_MPI_COMM_FLAG semantics, driver_nonhydro3d.F90:859-895).
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

// gather the face nodes of six fields into buf[field][m]
__global__ void pack_face_kernel(const double* q0, const double* q1, const double* q2, const double* q3, const double* q4, const double* dp,
                                 const int* __restrict__ vmapB, int off, int cnt, double* __restrict__ buf) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= cnt) return;
  const int src = vmapB[off + m];
  buf[m] = q0[src]; buf[cnt + m] = q1[src]; buf[2 * cnt + m] = q2[src]; buf[3 * cnt + m] = q3[src]; buf[4 * cnt + m] = q4[src];
  buf[5 * size_t(cnt) + m] = dp[src];
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
  }
  // receive order: ascending face id of the sender
  for (int i = 0; i < cs.nremote; ++i) cs.recv_order[i] = i;
  std::sort(cs.recv_order, cs.recv_order + cs.nremote, [&](int a, int b) { return cs.face[a].peer_face < cs.face[b].peer_face; });
  if (cudaStreamCreateWithFlags(&cs.stream, cudaStreamNonBlocking) != cudaSuccess) { err = "cudaStreamCreate(comm)"; return FEDG_ERR_CUDA; }
  cudaEventCreateWithFlags(&cs.ev_packed, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&cs.ev_done, cudaEventDisableTiming);
  cs.active = true;
  return FEDG_OK;
}

void comm_destroy(CommState& cs) {
  if (!cs.active) return;
  for (int i = 0; i < cs.nremote; ++i) if (cs.face[i].sendbuf) cudaFree(cs.face[i].sendbuf);
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
  for (int i = 0; i < cs.nremote; ++i) {
    const RemoteFace& rf = cs.face[i];
    pack_face_kernel<<<(rf.cnt + 255) / 256, 256, 0, compute>>>(q[0], q[1], q[2], q[3], q[4], dp, d_vmapB, rf.off, rf.cnt, rf.sendbuf);
  }
  cudaEventRecord(cs.ev_packed, compute);
  cudaStreamWaitEvent(cs.stream, cs.ev_packed, 0);
  ncclComm_t comm = static_cast<ncclComm_t>(cs.comm);
  ncclResult_t r = g_nccl.GroupStart();
  for (int i = 0; i < cs.nremote && r == ncclSuccess; ++i) {           // ascending own face id; six messages per face
    const RemoteFace& rf = cs.face[i];
    for (int v = 0; v < 6 && r == ncclSuccess; ++v)
      r = g_nccl.Send(rf.sendbuf + size_t(v) * rf.cnt, size_t(rf.cnt), ncclDouble, rf.peer, comm, cs.stream);
  }
  for (int oi = 0; oi < cs.nremote && r == ncclSuccess; ++oi) {        // ascending face id of the sender
    const RemoteFace& rf = cs.face[cs.recv_order[oi]];
    for (int v = 0; v < 6 && r == ncclSuccess; ++v) {
      double* dst = (v < NVAR ? q[v] : dp) + nint + rf.off;            // halo slots of this face: contiguous
      r = g_nccl.Recv(dst, size_t(rf.cnt), ncclDouble, rf.peer, comm, cs.stream);
    }
  }
  ncclResult_t r2 = g_nccl.GroupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess) { err = std::string("NCCL halo exchange: ") + g_nccl.GetErrorString(r); return FEDG_ERR_COMM; }
  cudaEventRecord(cs.ev_done, cs.stream);
  return FEDG_OK;
}

// Get: the compute stream waits for the halo data
void comm_exchange_wait(CommState& cs, cudaStream_t compute) {
  if (!cs.active || cs.nremote == 0) return;
  cudaStreamWaitEvent(compute, cs.ev_done, 0);
}

// global sums for the monitors (MPI_Allreduce in file/scale_file_monitor_meshfield.F90:203-211)
int comm_allreduce_sum(CommState& cs, double* d_inout, int n, cudaStream_t s, std::string& err) {
  if (!cs.active) return FEDG_OK;
  ncclResult_t r = g_nccl.AllReduce(d_inout, d_inout, size_t(n), ncclDouble, ncclSum, static_cast<ncclComm_t>(cs.comm), s);
  if (r != ncclSuccess) { err = std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r); return FEDG_ERR_COMM; }
  return FEDG_OK;
}

}  // namespace fedg
