// comm.cu -- NCCL over NVLink: halo send/recv and dot-product all-reduce.
// Replaces S/Communicator.cpp (MPI_Ssend/Isend/Irecv/Waitall :57-62,99-130 and
// MPI_Allreduce :81-141) for the traffic around and inside the Krylov loop.
// libnccl is dlopen'ed (the torch-bundled libnccl.so.2 is already mapped in a
// torchrun rank; a single-GPU process never touches it).
#include <dlfcn.h>

#include <algorithm>
#include <vector>
#include <nccl.h>

#include "comm.cuh"

namespace phb {
namespace {
struct Nccl {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
} g;

int load() {
  if (g.h) return PHB_OK;
  const char *names[] = {getenv("PHB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    if (!n) continue;
    g.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g.h) break;
  }
  if (!g.h) {
    set_error("cannot dlopen libnccl (set PHB_NCCL_LIB): %s", dlerror());
    return PHB_ERR_COMM;
  }
#define SYM(f, name)                                    \
  *(void **)(&g.f) = dlsym(g.h, name);                  \
  if (!g.f) {                                           \
    set_error("libnccl lacks %s", name);                \
    return PHB_ERR_COMM;                                \
  }
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllReduce, "ncclAllReduce")
  SYM(AllGather, "ncclAllGather")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return PHB_OK;
}
#define PHB_NCCL(call)                                                            \
  do {                                                                            \
    ncclResult_t r__ = (call);                                                    \
    if (r__ != ncclSuccess) {                                                     \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g.GetErrorString(r__)); \
      return PHB_ERR_COMM;                                                        \
    }                                                                             \
  } while (0)
}  // namespace

int comm_unique_id(void *out128) {
  PHB_REQUIRE(out128, "phb_comm_unique_id: out is NULL");
  PHB_CHECK(load());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  PHB_NCCL(g.GetUniqueId(&id));
  memcpy(out128, &id, 128);
  return PHB_OK;
}

int comm_init(phb_ctx *c, int rank, int nProcs, const void *id128) {
  PHB_REQUIRE(nProcs >= 1 && rank >= 0 && rank < nProcs, "phb_ctx_init_comm: bad rank %d/%d", rank, nProcs);
  c->rank = rank;
  c->nProcs = nProcs;
  if (nProcs == 1) return PHB_OK;
  PHB_REQUIRE(id128, "phb_ctx_init_comm: id is NULL");
  PHB_CHECK(load());
  PHB_CUDA(cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm;
  PHB_NCCL(g.CommInitRank(&comm, nProcs, id, rank));
  c->comm = (ncclComm *)comm;
  return PHB_OK;
}

void comm_destroy(phb_ctx *c) {
  if (c->comm && g.CommDestroy) g.CommDestroy((ncclComm_t)c->comm);
  c->comm = nullptr;
}

int comm_allreduce_sum(phb_ctx *c, double *dev, int n) {
  if (c->nProcs == 1) return PHB_OK;
  PHB_NCCL(g.AllReduce(dev, dev, n, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
  return PHB_OK;
}
int comm_allreduce_max(phb_ctx *c, double *dev, int n) {
  if (c->nProcs == 1) return PHB_OK;
  PHB_NCCL(g.AllReduce(dev, dev, n, ncclDouble, ncclMax, (ncclComm_t)c->comm, c->stream));
  return PHB_OK;
}

int comm_exchange(phb_ctx *c, const double *sendBuf, const int *sendOff, const int *sendCnt,
                  double *recvBuf, const int *recvOff, const int *recvCnt) {
  if (c->nProcs == 1) return PHB_OK;
  PHB_NCCL(g.GroupStart());
  for (int q = 0; q < c->nProcs; ++q) {
    if (q == c->rank) continue;
    if (sendCnt[q])
      PHB_NCCL(g.Send(sendBuf + sendOff[q], sendCnt[q], ncclDouble, q, (ncclComm_t)c->comm, c->stream));
    if (recvCnt[q])
      PHB_NCCL(g.Recv(recvBuf + recvOff[q], recvCnt[q], ncclDouble, q, (ncclComm_t)c->comm, c->stream));
  }
  PHB_NCCL(g.GroupEnd());
  return PHB_OK;
}

int comm_exchange_bytes(phb_ctx *c, const void *sendBuf, const int *sendOff, const int *sendCnt, void *recvBuf,
                        const int *recvOff, const int *recvCnt, size_t elem) {
  if (c->nProcs == 1) return PHB_OK;
  PHB_NCCL(g.GroupStart());
  for (int q = 0; q < c->nProcs; ++q) {
    if (q == c->rank) continue;
    if (sendCnt[q])
      PHB_NCCL(g.Send((const char *)sendBuf + (size_t)sendOff[q] * elem, (size_t)sendCnt[q] * elem, ncclChar, q,
                      (ncclComm_t)c->comm, c->stream));
    if (recvCnt[q])
      PHB_NCCL(g.Recv((char *)recvBuf + (size_t)recvOff[q] * elem, (size_t)recvCnt[q] * elem, ncclChar, q,
                      (ncclComm_t)c->comm, c->stream));
  }
  PHB_NCCL(g.GroupEnd());
  return PHB_OK;
}

int comm_allgatherv_host(phb_ctx *c, const std::vector<char> &mine, std::vector<std::vector<char>> &all) {
  const int P = c->nProcs;
  all.assign(P, std::vector<char>());
  if (P == 1) { all[0] = mine; return PHB_OK; }
  long long *dSizes = nullptr;
  PHB_CUDA(cudaMalloc((void **)&dSizes, P * sizeof(long long)));
  const long long my = (long long)mine.size();
  std::vector<long long> sizes(P, 0);
  PHB_CUDA(cudaMemcpyAsync(dSizes + c->rank, &my, sizeof(long long), cudaMemcpyHostToDevice, c->stream));
  PHB_NCCL(g.AllGather(dSizes + c->rank, dSizes, 1, ncclInt64, (ncclComm_t)c->comm, c->stream));
  PHB_CUDA(cudaMemcpyAsync(sizes.data(), dSizes, P * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  PHB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(dSizes);
  size_t slot = 16;
  for (long long sz : sizes) slot = std::max(slot, ((size_t)sz + 15) / 16 * 16);
  char *dBuf = nullptr;
  PHB_CUDA(cudaMalloc((void **)&dBuf, slot * P));
  if (my) PHB_CUDA(cudaMemcpyAsync(dBuf + slot * c->rank, mine.data(), (size_t)my, cudaMemcpyHostToDevice, c->stream));
  PHB_NCCL(g.AllGather(dBuf + slot * c->rank, dBuf, slot, ncclChar, (ncclComm_t)c->comm, c->stream));
  std::vector<char> host(slot * P);
  PHB_CUDA(cudaMemcpyAsync(host.data(), dBuf, slot * P, cudaMemcpyDeviceToHost, c->stream));
  PHB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(dBuf);
  for (int q = 0; q < P; ++q) all[q].assign(host.begin() + slot * q, host.begin() + slot * q + (size_t)sizes[q]);
  return PHB_OK;
}
}  // namespace phb
