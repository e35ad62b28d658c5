// peerdev.cuh -- device side of the NVLink peer-memory exchanges: arena header layout, epoch flags,
// and the block-/warp-level helpers that let COMPUTE kernels push their own halo values and partial
// sums to the peers (producer side, last CTA to finish) and wait for the peers' contributions in their
// prologue (consumer side) -- no separate communication kernel, no host involvement.
#pragma once
#include "comm.cuh"

namespace phb {

struct RedSlot { double v[6]; unsigned long long epoch, pad; };  // 64 B
struct PeerView {
  char *arena;
  char *peer[kMaxPeers];
  int rank, nProcs;
};
// what a compute kernel has to do for its neighbours (all fields by value: lives in the kernel parameters)
struct PeerFuse {
  int on = 0;          // 0: single GPU / NCCL / unfused peer kernels
  int waitHalo = 0;    // prologue: wait for the ghosts of the gathered vector (channel haloCh)
  int pushHalo = 0;    // epilogue (last CTA): push my send list of the produced vector (channel haloCh)
  int pushRed = 0;     // epilogue (last CTA, warp 0): push my partial sums (channel redCh)
  int waitRed = 0;     // prologue: all-reduce of the peers' partial sums (channel redCh)
  int haloCh = 0, redCh = 0;
  size_t vecOff = 0;   // offset of the exchanged vector inside every rank's arena
  PeerView pv;
  PeerHalo halo;
};

// Reduction slots are double-buffered by epoch parity: consecutive reductions on one channel may follow each
// other without a global synchronisation in between, and a fast rank must not overwrite the values of epoch e
// that a slow rank has not read yet.  With two slots, epoch e + 2 (same slot as e) can only be written by a rank
// that has finished reduction e + 1, which needed every rank's contribution to e + 1 -- sent after it read e.
__device__ __forceinline__ RedSlot *red_slot(char *arena, int ch, int src, unsigned long long epoch) {
  return reinterpret_cast<RedSlot *>(arena) + ((size_t)ch * 2 + (epoch & 1ull)) * kMaxPeers + src;
}
__device__ __forceinline__ unsigned long long *halo_flag(char *arena, int ch, int src) {
  return reinterpret_cast<unsigned long long *>(arena + 2 * kPeerRedChannels * kMaxPeers * sizeof(RedSlot)) +
         (size_t)ch * kMaxPeers + src;
}
__device__ __forceinline__ unsigned long long *local_epoch(char *arena, int idx) {
  return reinterpret_cast<unsigned long long *>(arena + 20480) + idx;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Block-uniform: true in the last CTA of the grid to get here (everything the other CTAs wrote before
// is visible to it).  `ticket` must be 0 at kernel start; it is reset for the next launch.
__device__ __forceinline__ bool last_block(unsigned *ticket) {
  __shared__ bool isLast;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    isLast = (t == gridDim.x - 1);
    if (isLast) *ticket = 0u;
  }
  __syncthreads();
  if (isLast) __threadfence();
  return isLast;
}

// producer: all threads of ONE block.  Gathers x[sendDev] and stores it into the peers' ghost segments.
__device__ __forceinline__ void halo_push_block(const PeerFuse &F, const double *x, int nComp, int ld) {
  __shared__ unsigned long long se;
  if (threadIdx.x == 0) {
    unsigned long long *ep = local_epoch(F.pv.arena, kPeerRedChannels + F.haloCh);
    se = *ep + 1;
    *ep = se;
  }
  __syncthreads();
  for (int q = 0; q < F.pv.nProcs; ++q) {
    const int cnt = F.halo.sendCnt[q];
    if (q == F.pv.rank || cnt == 0) continue;
    double *dst = reinterpret_cast<double *>(F.pv.peer[q] + F.vecOff);
    const int *sd = F.halo.sendDev + F.halo.sendOff[q];
    for (int j = threadIdx.x; j < cnt * nComp; j += blockDim.x) {
      const int c = j / cnt, i = j - c * cnt;
      dst[(size_t)c * F.halo.peerLd[q] + F.halo.peerRecvOff[q] + i] = __ldcg(x + (size_t)c * ld + sd[i]);
    }
  }
  __threadfence_system();
  __syncthreads();
  const int t = threadIdx.x;
  if (t < F.pv.nProcs && t != F.pv.rank && F.halo.sendCnt[t] > 0) st_release_sys(halo_flag(F.pv.peer[t], F.haloCh, F.pv.rank), se);
}
// consumer: every block, in its prologue, before any ghost value is read
__device__ __forceinline__ void halo_wait_block(const PeerFuse &F) {
  const int t = threadIdx.x;
  if (t < F.pv.nProcs && t != F.pv.rank && F.halo.recvCnt[t] > 0) {
    const unsigned long long e = *local_epoch(F.pv.arena, kPeerRedChannels + F.haloCh);  // set by my own push
    const unsigned long long *f = halo_flag(F.pv.arena, F.haloCh, t);
    while (ld_acquire_sys(f) < e) {}
  }
  __syncthreads();
}
// producer: warp 0 of the last CTA (all 32 lanes), after `vals` has been written by lane 0
__device__ __forceinline__ void reduce_push_warp(const PeerView &pv, int ch, const double *vals, int nvals) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  unsigned long long e = 0;
  if (lane == 0) {
    unsigned long long *ep = local_epoch(pv.arena, ch);
    e = *ep + 1;
    *ep = e;
  }
  e = __shfl_sync(0xffffffffu, e, 0);
  if (lane < pv.nProcs) {
    RedSlot *dst = red_slot(pv.peer[lane], ch, pv.rank, e);
    for (int k = 0; k < nvals; ++k) dst->v[k] = __ldcg(vals + k);
    __threadfence_system();
    st_release_sys(&dst->epoch, e);
  }
}
// consumer: every block in its prologue; out[k] = sum over ranks (rank order -> identical everywhere)
template <int NV>
__device__ __forceinline__ void reduce_wait_block(const PeerView &pv, int ch, double (&out)[NV]) {
  __shared__ double sv[kMaxPeers][NV];
  const int t = threadIdx.x;
  if (t < pv.nProcs) {
    const unsigned long long e = *local_epoch(pv.arena, ch);  // set by my own push (previous kernel)
    const RedSlot *src = red_slot(pv.arena, ch, t, e);
    while (ld_acquire_sys(&src->epoch) < e) {}
#pragma unroll
    for (int k = 0; k < NV; ++k) sv[t][k] = *reinterpret_cast<const volatile double *>(&src->v[k]);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = 0.;
    for (int q = 0; q < pv.nProcs; ++q) x += sv[q][k];
    out[k] = x;
  }
}

PeerView peer_view(const phb_ctx *c);

}  // namespace phb
