// kernels.cuh -- device helpers shared by the solver and assembly kernels:
// deterministic block + grid reductions (warp shuffles, fixed-order partials,
// last-block finish) and streaming load wrappers.
#pragma once
#include "common.cuh"

namespace phb {

struct SellView {
  const int *sliceOff;
  const int *col;
  int nRows, nSlices, nCols;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming (read-once) loads: keep them out of L1 and first in line for L2 eviction
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int *p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream(const float *p) { return __ldcs(p); }

constexpr int kMaxBlockWarps = 32;

// Sum NS per-thread values over the grid, deterministically:
//   warp shuffle -> shared -> one partial per block -> the last block to arrive
//   (ticket counter) adds the partials in fixed order and writes out[0..NS).
// `ticket` must be zero on entry and is reset for the next launch.
// Returns true in ALL lanes of warp 0 of the last CTA (warp-uniform) after lane 0 wrote `out`;
// scalar follow-ups must be guarded with lane == 0.
template <int NS, bool MAX = false>
__device__ __forceinline__ bool grid_reduce(double (&v)[NS], double *partials, unsigned *ticket,
                                            double *out) {
  __shared__ double sh[NS][kMaxBlockWarps];
  __shared__ bool isLast;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const double w = MAX ? warp_max(v[j]) : warp_sum(v[j]);
    if (lane == 0) sh[j][warp] = w;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      double w = lane < nWarps ? sh[j][lane] : (MAX ? -1e300 : 0.);
      w = MAX ? warp_max(w) : warp_sum(w);
      if (lane == 0) partials[(size_t)blockIdx.x * NS + j] = w;
    }
    if (lane == 0) {
      __threadfence();
      const unsigned t = atomicAdd(ticket, 1u);
      isLast = (t == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!isLast) return false;
  __threadfence();
  double acc[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) acc[j] = MAX ? -1e300 : 0.;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const double w = __ldcg(&partials[(size_t)b * NS + j]);
      acc[j] = MAX ? fmax(acc[j], w) : acc[j] + w;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const double w = MAX ? warp_max(acc[j]) : warp_sum(acc[j]);
    if (lane == 0) sh[j][warp] = w;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      double w = lane < nWarps ? sh[j][lane] : (MAX ? -1e300 : 0.);
      w = MAX ? warp_max(w) : warp_sum(w);
      if (lane == 0) out[j] = w;
    }
    if (lane == 0) *ticket = 0u;
    __syncwarp();
    return true;
  }
  return false;
}

// T = double for the Krylov SpMV, float inside a single-precision multigrid cycle
template <int NC, int W, typename T>
__device__ __forceinline__ void slice_dot(const int *__restrict__ col, const T *__restrict__ vals,
                                          size_t base, const T *__restrict__ x, int ld, T (&acc)[NC]) {
  int c[W];
  T a[W];
#pragma unroll
  for (int k = 0; k < W; ++k) {
    c[k] = ld_stream(col + base + (size_t)k * 32);
    a[k] = ld_stream(vals + base + (size_t)k * 32);
  }
#pragma unroll
  for (int k = 0; k < W; ++k) {
#pragma unroll
    for (int i = 0; i < NC; ++i) acc[i] = fma(a[k], __ldg(x + (size_t)i * ld + c[k]), acc[i]);
  }
}

template <int NC, typename T>
__device__ __forceinline__ void slice_dot_any(const int *__restrict__ col, const T *__restrict__ vals,
                                              size_t base, int w, const T *__restrict__ x, int ld, T (&acc)[NC]) {
  switch (w) {
    case 3: slice_dot<NC, 3, T>(col, vals, base, x, ld, acc); break;
    case 4: slice_dot<NC, 4, T>(col, vals, base, x, ld, acc); break;
    case 5: slice_dot<NC, 5, T>(col, vals, base, x, ld, acc); break;
    case 6: slice_dot<NC, 6, T>(col, vals, base, x, ld, acc); break;
    case 7: slice_dot<NC, 7, T>(col, vals, base, x, ld, acc); break;
    default: {
      int k = 0;
      for (; k + 4 <= w; k += 4) slice_dot<NC, 4, T>(col, vals, base + (size_t)k * 32, x, ld, acc);
      for (; k < w; ++k) slice_dot<NC, 1, T>(col, vals, base + (size_t)k * 32, x, ld, acc);
    }
  }
}

// Wide rows (Galerkin operators: 9-13 entries; restrictions: 20-30): the whole row -- or sixteen entries at a time --
// is in flight before the first gather, so a slice costs one or two load round trips instead of one per four entries.
// Levels of 10^5-10^6 rows hand each warp two or three slices: they are bound by that chain, not by bytes.
template <int NC, typename T>
__device__ __forceinline__ void slice_dot_wide(const int *__restrict__ col, const T *__restrict__ vals,
                                               size_t base, int w, const T *__restrict__ x, int ld, T (&acc)[NC]) {
  int k = 0;
  for (; k + 16 <= w; k += 16) slice_dot<NC, 16, T>(col, vals, base + (size_t)k * 32, x, ld, acc);
  switch (w - k) {
    case 15: slice_dot<NC, 15, T>(col, vals, base + (size_t)k * 32, x, ld, acc); break;
    case 14: slice_dot<NC, 14, T>(col, vals, base + (size_t)k * 32, x, ld, acc); break;
    case 13: slice_dot<NC, 13, T>(col, vals, base + (size_t)k * 32, x, ld, acc); break;
    case 12: slice_dot<NC, 12, T>(col, vals, base + (size_t)k * 32, x, ld, acc); break;
    case 11: slice_dot<NC, 11, T>(col, vals, base + (size_t)k * 32, x, ld, acc); break;
    case 10: slice_dot<NC, 10, T>(col, vals, base + (size_t)k * 32, x, ld, acc); break;
    case 9: slice_dot<NC, 9, T>(col, vals, base + (size_t)k * 32, x, ld, acc); break;
    case 8: slice_dot<NC, 8, T>(col, vals, base + (size_t)k * 32, x, ld, acc); break;
    case 7: slice_dot<NC, 7, T>(col, vals, base + (size_t)k * 32, x, ld, acc); break;
    case 6: slice_dot<NC, 6, T>(col, vals, base + (size_t)k * 32, x, ld, acc); break;
    case 5: slice_dot<NC, 5, T>(col, vals, base + (size_t)k * 32, x, ld, acc); break;
    case 4: slice_dot<NC, 4, T>(col, vals, base + (size_t)k * 32, x, ld, acc); break;
    case 3: slice_dot<NC, 3, T>(col, vals, base + (size_t)k * 32, x, ld, acc); break;
    case 2: slice_dot<NC, 2, T>(col, vals, base + (size_t)k * 32, x, ld, acc); break;
    case 1: slice_dot<NC, 1, T>(col, vals, base + (size_t)k * 32, x, ld, acc); break;
    default: break;
  }
}

}  // namespace phb
