// comm.cu -- see comm.h
#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

namespace beom {
namespace {
struct Api {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
} api;
ncclComm_t g_comm = nullptr;
int g_rank = 0, g_size = 1;

int load(std::string *err) {
  if (api.h) return 0;
  api.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!api.h) api.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!api.h) {
    if (err) *err = std::string("cannot load libnccl.so.2: ") + dlerror();
    return -40;
  }
#define SYM(field, name)                                      \
  *(void **)(&api.field) = dlsym(api.h, name);                \
  if (!api.field) {                                           \
    if (err) *err = std::string("NCCL symbol missing: ") + name; \
    return -41;                                               \
  }
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(AllReduce, "ncclAllReduce")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return 0;
}
int nccl_fail(ncclResult_t r, const char *what, std::string *err) {
  if (err) *err = std::string(what) + ": " + (api.GetErrorString ? api.GetErrorString(r) : "NCCL error");
  return -42;
}
#define NC(call, what)                                   \
  do {                                                   \
    ncclResult_t r_ = (call);                            \
    if (r_ != ncclSuccess) return nccl_fail(r_, what, err); \
  } while (0)
}  // namespace

int comm_unique_id(char id[128], std::string *err) {
  int rc = load(err);
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId u;
  NC(api.GetUniqueId(&u), "ncclGetUniqueId");
  memcpy(id, &u, 128);
  return 0;
}
int comm_init(const char id[128], int rank, int nranks, int device, std::string *err) {
  int rc = load(err);
  if (rc) return rc;
  if (g_comm) comm_finalize();
  if (device >= 0) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
      if (err) *err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
      return -43;
    }
  }
  ncclUniqueId u;
  memcpy(&u, id, 128);
  NC(api.CommInitRank(&g_comm, nranks, u, rank), "ncclCommInitRank");
  g_rank = rank;
  g_size = nranks;
  return 0;
}
int comm_finalize() {
  if (g_comm && api.CommDestroy) api.CommDestroy(g_comm);
  g_comm = nullptr;
  g_rank = 0;
  g_size = 1;
  return 0;
}
bool comm_ready() { return g_comm != nullptr; }
int comm_rank() { return g_rank; }
int comm_size() { return g_size; }

int comm_exchange(const double *send_lo, double *recv_lo, int peer_lo, const double *send_hi, double *recv_hi, int peer_hi,
                  size_t count, cudaStream_t s, std::string *err) {
  if (!g_comm) { if (err) *err = "comm_exchange: communicator not initialised"; return -44; }
  // Sends in the order (lo, hi), receives in the order (hi, lo): on a two-rank ring both peers are the same rank, and
  // messages between one pair match in posting order -- what this rank sends "down" is what the peer receives "from above".
  NC(api.GroupStart(), "ncclGroupStart");
  if (peer_lo >= 0) NC(api.Send(send_lo, count, ncclDouble, peer_lo, g_comm, s), "ncclSend");
  if (peer_hi >= 0) NC(api.Send(send_hi, count, ncclDouble, peer_hi, g_comm, s), "ncclSend");
  if (peer_hi >= 0) NC(api.Recv(recv_hi, count, ncclDouble, peer_hi, g_comm, s), "ncclRecv");
  if (peer_lo >= 0) NC(api.Recv(recv_lo, count, ncclDouble, peer_lo, g_comm, s), "ncclRecv");
  NC(api.GroupEnd(), "ncclGroupEnd");
  return 0;
}
int comm_allreduce_sum(double *buf, size_t count, cudaStream_t s, std::string *err) {
  if (!g_comm) return 0;
  NC(api.AllReduce(buf, buf, count, ncclDouble, ncclSum, g_comm, s), "ncclAllReduce");
  return 0;
}
int comm_allreduce_max(double *buf, size_t count, cudaStream_t s, std::string *err) {
  if (!g_comm) return 0;
  NC(api.AllReduce(buf, buf, count, ncclDouble, ncclMax, g_comm, s), "ncclAllReduce");
  return 0;
}
}  // namespace beom
