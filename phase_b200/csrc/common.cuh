// common.cuh -- internal helpers shared by the translation units of libphase_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/phase_b200.h"

namespace phb {

void set_error(const char *fmt, ...);

#define PHB_CUDA(call)                                                         \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      phb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,             \
                     cudaGetErrorString(e__));                                 \
      return PHB_ERR_CUDA;                                                     \
    }                                                                          \
  } while (0)

#define PHB_CHECK(expr)                                                        \
  do {                                                                         \
    int rc__ = (expr);                                                         \
    if (rc__ != PHB_OK) return rc__;                                           \
  } while (0)

#define PHB_REQUIRE(cond, ...)                                                 \
  do {                                                                         \
    if (!(cond)) {                                                             \
      phb::set_error(__VA_ARGS__);                                             \
      return PHB_ERR_ARG;                                                      \
    }                                                                          \
  } while (0)

#define PHB_TRY_BEGIN try {
#define PHB_TRY_END                                                            \
  }                                                                            \
  catch (const std::exception &ex__) {                                         \
    phb::set_error("exception: %s", ex__.what());                              \
    return PHB_ERR_ARG;                                                        \
  }                                                                            \
  catch (...) {                                                                \
    phb::set_error("unknown exception");                                       \
    return PHB_ERR_ARG;                                                        \
  }

// device buffer with explicit lifetime (no exceptions on the hot path)
template <class T> struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  bool owned = true;  // false: points into storage owned by someone else (peer arena)
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p && owned) cudaFree(p);
    p = nullptr;
    n = 0;
    owned = true;
  }
  void attach(T *ext, size_t count) {
    release();
    p = ext;
    n = count;
    owned = false;
  }
  int alloc(size_t count) {
    if (count == n && p) return PHB_OK;
    release();
    if (count == 0) return PHB_OK;
    PHB_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
    n = count;
    return PHB_OK;
  }
  int upload(const T *h, size_t count, cudaStream_t st) {
    PHB_CHECK(alloc(count));
    if (count) PHB_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, st));
    return PHB_OK;
  }
  int upload(const std::vector<T> &h, cudaStream_t st) { return upload(h.data(), h.size(), st); }
  int zero(cudaStream_t st) {
    if (n) PHB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), st));
    return PHB_OK;
  }
};

constexpr int kSliceRows = 32;  // SELL slice height = one warp, lane <-> row

}  // namespace phb

struct ncclComm;
namespace phb { struct PinnedStage; }

// Peer-memory communication over NVLink (CUDA IPC, one arena per rank, identical
// layout on every rank): reductions and halos inside the Krylov loop are done by
// small kernels that store straight into the peers' arenas and spin on epoch flags
// -- no NCCL call, no host involvement (peer.cu).
constexpr int kMaxPeers = 8;
constexpr int kPeerRedChannels = 16, kPeerHaloChannels = 16;
constexpr size_t kPeerHeaderBytes = 32768;  // red slots (two per channel and source, by epoch parity) | halo flags | local epochs
struct PeerComm {
  bool enabled = false;
  char *arena = nullptr;
  size_t arenaBytes = 0, regionBytes = 0, vecBytes = 0;
  long long maxCols = 0;
  int maxRegions = 0;
  unsigned regionMask = 0;   // regions handed to live solvers (returned on phb_solver_destroy)
  char *peerArena[kMaxPeers] = {nullptr};
  bool opened[kMaxPeers] = {false};
};

struct phb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;       // compute stream (all kernels)
  cudaStream_t commStream = nullptr;   // halo traffic
  int numSMs = 148;
  int rank = 0, nProcs = 1;
  ncclComm *comm = nullptr;
  long long launches = 0;
  // pinned scratch for small device->host reads
  double *pinned = nullptr;
  PeerComm peer;
  // live solvers: their CUDA graphs may hold NCCL nodes, which must be gone before the communicator is
  std::vector<struct phb_solver *> solvers;
  phb::PinnedStage *stage = nullptr;   // pageable <-> device transfers of Seam 1 (hostcopy.cuh), created on first use
  // first refused kernel launch since the last status check (PHB_LAUNCH)
  cudaError_t launchError = cudaSuccess;
  char launchWhere[160] = "";
};

namespace phb {
void solver_drop_graph(struct phb_solver *s);   // solver.cu
// PHB_ERR_CUDA (with the launch site in phb_last_error) if a kernel launch was refused since the last call
inline int launch_status(phb_ctx *c) {
  if (c->launchError == cudaSuccess) return PHB_OK;
  set_error("kernel launch failed at %s: %s", c->launchWhere, cudaGetErrorString(c->launchError));
  c->launchError = cudaSuccess;
  cudaGetLastError();
  return PHB_ERR_CUDA;
}
}

// A refused launch (bad grid, too much shared memory ...) is recorded in the context -- launch sites sit in
// void helpers and inside stream capture -- and reported by the next phb_ctx_launch_status() check
// (end of every solve and time step, phb_ctx_sync).
#define PHB_LAUNCH(ctx, kernel, grid, block, smem, ...)                        \
  do {                                                                         \
    auto kfn__ = kernel;                                                       \
    kfn__<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);            \
    (ctx)->launches++;                                                         \
    cudaError_t le__ = cudaPeekAtLastError();                                  \
    if (le__ != cudaSuccess && (ctx)->launchError == cudaSuccess) {            \
      (ctx)->launchError = le__;                                               \
      snprintf((ctx)->launchWhere, sizeof((ctx)->launchWhere), "%s:%d %s",     \
               __FILE__, __LINE__, #kernel);                                   \
    }                                                                          \
  } while (0)
