// comm.cuh -- NCCL plumbing (loaded at run time from the torch-bundled libnccl so
// that a single-GPU process never needs it).
#pragma once
#include <vector>

#include "common.cuh"

namespace phb {
int comm_unique_id(void *out128);
int comm_init(phb_ctx *c, int rank, int nProcs, const void *id128);
void comm_destroy(phb_ctx *c);
// sum-allreduce of n doubles in place on the compute stream (no-op when nProcs == 1)
int comm_allreduce_sum(phb_ctx *c, double *dev, int n);
int comm_allreduce_max(phb_ctx *c, double *dev, int n);
// grouped neighbour exchange on the compute stream:
// for each peer q: send sendBuf[sendOff[q] .. +sendCnt[q]) , recv into recvBuf[recvOff[q] .. +recvCnt[q])
int comm_exchange(phb_ctx *c, const double *sendBuf, const int *sendOff, const int *sendCnt,
                  double *recvBuf, const int *recvOff, const int *recvCnt);

// the same with elements of `elem` bytes (offsets and counts in elements): single-precision level vectors
int comm_exchange_bytes(phb_ctx *c, const void *sendBuf, const int *sendOff, const int *sendCnt, void *recvBuf,
                        const int *recvOff, const int *recvCnt, size_t elem);
// all-gather of one host byte blob per rank (setup-time metadata exchange; staged through device memory)
int comm_allgatherv_host(phb_ctx *c, const std::vector<char> &mine, std::vector<std::vector<char>> &all);

// ---- peer-memory path (peer.cu)
struct PeerHalo {             // per-mesh halo description for the peer kernels
  const int *sendDev;         // device: send list (device cell ids), grouped by peer
  int sendOff[kMaxPeers], sendCnt[kMaxPeers], recvCnt[kMaxPeers];
  int peerRecvOff[kMaxPeers]; // where my values land in peer q's vectors
  int peerLd[kMaxPeers];      // peer q's vector leading dimension
};
int peer_arena_create(phb_ctx *c, long long maxCols, int maxRegions, void *handle64);
int peer_arena_open(phb_ctx *c, const void *handles);
void peer_destroy(phb_ctx *c);
// carve vector `vec` (0..3) of solver region `region` out of the arena
double *peer_vector(phb_ctx *c, int region, int vec);
bool peer_owns(const phb_ctx *c, const void *p);
// all-reduce (sum) of nvals <= 6 doubles at `vals` through channel ch; when S != NULL the kernel
// exits early once the Krylov loop is done; finishIter publishes the iteration's scalars
int peer_allreduce(phb_ctx *c, int ch, double *vals, int nvals, void *S, int maxIters, int finishIter, int cur);
// push the send-list entries of x (inside the arena) into the peers' ghost segments and wait for ours
int peer_halo(phb_ctx *c, int ch, const PeerHalo &h, double *x, int nComp, int ld, void *S, int maxIters);
}  // namespace phb
