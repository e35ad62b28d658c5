// peer.cu -- NVLink peer-memory collectives fused into the Krylov loop.
//
// The reference moves halos with MPI_Ssend/Irecv and reduces dot products with
// MPI_Allreduce (S/Communicator.cpp:57-141) -- inside Trilinos for the solve.  The
// message sizes here are tiny (2 doubles; a few thousand boundary values), so the
// cost is latency, not bandwidth.  Instead of one NCCL kernel per message this
// file uses ONE small kernel per exchange that
//   * stores its contribution straight into every peer's arena over NVLink
//     (peer pointers from cudaIpcOpenMemHandle),
//   * publishes an epoch flag after a system-scope fence, and
//   * spins on its own flags until the peers' contributions have landed.
// All ranks sum the contributions in rank order, so the reduced values (and hence
// every convergence decision) are bitwise identical on all ranks.
#include "peerdev.cuh"
#include "solver.cuh"

namespace phb {
namespace {

__device__ __forceinline__ bool done_test(const KrylovSums *S, int maxIters) {
  return !(S->rr > S->thresh) || S->iters >= (double)maxIters;
}

// one warp: lane q < nProcs talks to peer q
__global__ void k_peer_allreduce(PeerView pv, int ch, double *vals, int nvals, KrylovSums *S, int maxIters,
                                 int finishIter, int cur) {
  if (S && done_test(S, maxIters)) return;
  const int t = threadIdx.x;
  __shared__ double sv[kMaxPeers][6];
  unsigned long long e = 0;
  if (t == 0) {
    unsigned long long *ep = local_epoch(pv.arena, ch);
    e = *ep + 1;
    *ep = e;
  }
  e = __shfl_sync(0xffffffffu, e, 0);
  if (t < pv.nProcs) {
    RedSlot *dst = red_slot(pv.peer[t], ch, pv.rank, e);
    for (int k = 0; k < nvals; ++k) dst->v[k] = vals[k];
    __threadfence_system();
    st_release_sys(&dst->epoch, e);
    const RedSlot *src = red_slot(pv.arena, ch, t, e);
    while (ld_acquire_sys(&src->epoch) < e) {}
    for (int k = 0; k < nvals; ++k) sv[t][k] = *reinterpret_cast<const volatile double *>(&src->v[k]);
  }
  __syncwarp();
  if (t == 0) {
    for (int k = 0; k < nvals; ++k) {
      double x = 0.;
      for (int q = 0; q < pv.nProcs; ++q) x += sv[q][k];  // fixed order on every rank
      vals[k] = x;
    }
    if (finishIter) krylov_finish(S, cur);  // alpha, omega, beta, rho', ||r'||^2, iters
  }
}

// single block: pack + remote store per peer, flag, then wait for the incoming halos
__global__ void __launch_bounds__(1024)
k_peer_halo(PeerView pv, int ch, PeerHalo h, const double *x, size_t vecOff, int nComp, int ld, KrylovSums *S,
            int maxIters) {
  if (S && done_test(S, maxIters)) return;
  __shared__ unsigned long long se;
  if (threadIdx.x == 0) {
    unsigned long long *ep = local_epoch(pv.arena, kPeerRedChannels + ch);
    se = *ep + 1;
    *ep = se;
  }
  __syncthreads();
  const unsigned long long e = se;
  for (int q = 0; q < pv.nProcs; ++q) {
    const int cnt = h.sendCnt[q];
    if (q == pv.rank || cnt == 0) continue;
    double *dst = reinterpret_cast<double *>(pv.peer[q] + vecOff);
    const int *sd = h.sendDev + h.sendOff[q];
    for (int j = threadIdx.x; j < cnt * nComp; j += blockDim.x) {
      const int c = j / cnt, i = j - c * cnt;
      dst[(size_t)c * h.peerLd[q] + h.peerRecvOff[q] + i] = x[(size_t)c * ld + sd[i]];
    }
  }
  __threadfence_system();
  __syncthreads();
  const int t = threadIdx.x;
  if (t < pv.nProcs && t != pv.rank) {
    if (h.sendCnt[t] > 0) st_release_sys(halo_flag(pv.peer[t], ch, pv.rank), e);
    if (h.recvCnt[t] > 0) {
      const unsigned long long *f = halo_flag(pv.arena, ch, t);
      while (ld_acquire_sys(f) < e) {}
    }
  }
  __syncthreads();
}

}  // namespace

PeerView peer_view(const phb_ctx *c) {
  PeerView v;
  v.arena = c->peer.arena;
  for (int q = 0; q < kMaxPeers; ++q) v.peer[q] = c->peer.peerArena[q];
  v.rank = c->rank;
  v.nProcs = c->nProcs;
  return v;
}

int peer_arena_create(phb_ctx *c, long long maxCols, int maxRegions, void *handle64) {
  PHB_REQUIRE(c && handle64 && maxCols > 0 && maxRegions > 0, "phb_ctx_peer_arena_create: bad argument");
  PHB_REQUIRE(c->device >= 0, "phb_ctx_peer_arena_create: host-only context");
  PHB_REQUIRE(c->nProcs <= kMaxPeers, "peer communication supports at most %d ranks", kMaxPeers);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  PeerComm &P = c->peer;
  if (P.arena) { cudaFree(P.arena); P.arena = nullptr; }
  P.maxCols = maxCols;
  P.maxRegions = maxRegions;
  P.vecBytes = ((size_t)2 * maxCols * sizeof(double) + 255) & ~(size_t)255;  // up to 2 components
  P.regionBytes = 4 * P.vecBytes;                                              // p, s, ph, sh
  P.arenaBytes = kPeerHeaderBytes + (size_t)maxRegions * P.regionBytes;
  PHB_CUDA(cudaMalloc((void **)&P.arena, P.arenaBytes));
  PHB_CUDA(cudaMemset(P.arena, 0, P.arenaBytes));
  cudaIpcMemHandle_t h;
  PHB_CUDA(cudaIpcGetMemHandle(&h, P.arena));
  memcpy(handle64, &h, 64);
  P.regionMask = 0;
  return PHB_OK;
}

int peer_arena_open(phb_ctx *c, const void *handles) {
  PHB_REQUIRE(c && handles, "phb_ctx_peer_arena_open: bad argument");
  PeerComm &P = c->peer;
  PHB_REQUIRE(P.arena, "phb_ctx_peer_arena_open: create the local arena first");
  for (int q = 0; q < c->nProcs; ++q) {
    if (q == c->rank) { P.peerArena[q] = P.arena; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + 64 * q, 64);
    void *p = nullptr;
    PHB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    P.peerArena[q] = (char *)p;
    P.opened[q] = true;
  }
  P.enabled = true;
  return PHB_OK;
}

void peer_destroy(phb_ctx *c) {
  PeerComm &P = c->peer;
  for (int q = 0; q < kMaxPeers; ++q)
    if (P.opened[q] && P.peerArena[q]) { cudaIpcCloseMemHandle(P.peerArena[q]); P.opened[q] = false; }
  if (P.arena) cudaFree(P.arena);
  P.arena = nullptr;
  P.enabled = false;
}

double *peer_vector(phb_ctx *c, int region, int vec) {
  PeerComm &P = c->peer;
  if (!P.enabled || region < 0 || region >= P.maxRegions || vec < 0 || vec > 3) return nullptr;
  return reinterpret_cast<double *>(P.arena + kPeerHeaderBytes + (size_t)region * P.regionBytes + (size_t)vec * P.vecBytes);
}

bool peer_owns(const phb_ctx *c, const void *p) {
  const PeerComm &P = c->peer;
  return P.enabled && (const char *)p >= P.arena + kPeerHeaderBytes && (const char *)p < P.arena + P.arenaBytes;
}

int peer_allreduce(phb_ctx *c, int ch, double *vals, int nvals, void *S, int maxIters, int finishIter, int cur) {
  PHB_REQUIRE(ch >= 0 && ch < kPeerRedChannels, "peer_allreduce: channel %d out of range", ch);
  PHB_LAUNCH(c, k_peer_allreduce, 1, 32, 0, peer_view(c), ch, vals, nvals, (KrylovSums *)S, maxIters, finishIter, cur);
  return PHB_OK;
}

int peer_halo(phb_ctx *c, int ch, const PeerHalo &h, double *x, int nComp, int ld, void *S, int maxIters) {
  PHB_REQUIRE(ch >= 0 && ch < kPeerHaloChannels, "peer_halo: channel %d out of range", ch);
  PHB_REQUIRE(peer_owns(c, x), "peer_halo: vector is not inside the peer arena");
  const size_t vecOff = (size_t)((char *)x - c->peer.arena);
  PHB_LAUNCH(c, k_peer_halo, 1, 1024, 0, peer_view(c), ch, h, x, vecOff, nComp, ld, (KrylovSums *)S, maxIters);
  return PHB_OK;
}

}  // namespace phb
