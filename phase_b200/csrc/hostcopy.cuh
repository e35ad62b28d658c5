// hostcopy.cuh -- moving the caller's pageable arrays (std::vector storage behind Seam 1) to and from the device at
// link speed: multi-threaded memcpy into a double-buffered pinned stage, DMA of one chunk overlapping the copy of
// the next.  cudaMemcpy from pageable memory stages through a single thread and reaches a fifth of the PCIe rate.
#pragma once
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"

namespace phb {

inline int host_copy_threads() {
  if (const char *e = getenv("PHB_HOST_THREADS")) return std::max(1, atoi(e));
  return std::max(1, std::min(8, (int)std::thread::hardware_concurrency()));
}

template <typename F> void parallel_ranges(size_t bytes, size_t minPerThread, F f) {
  const int T = (int)std::max<size_t>(1, std::min<size_t>(host_copy_threads(), bytes / std::max<size_t>(1, minPerThread)));
  if (T == 1) { f(0, bytes); return; }
  const size_t chunk = ((bytes + T - 1) / T + 63) & ~(size_t)63;
  std::vector<std::thread> th;
  for (int t = 0; t < T; ++t) {
    const size_t a = std::min(bytes, t * chunk), b = std::min(bytes, (t + 1) * chunk);
    if (a < b) th.emplace_back(f, a, b);
  }
  for (auto &x : th) x.join();
}

inline void parallel_memcpy(void *dst, const void *src, size_t bytes) {
  parallel_ranges(bytes, 1 << 20, [&](size_t a, size_t b) { memcpy((char *)dst + a, (const char *)src + a, b - a); });
}

inline bool parallel_equal(const void *x, const void *y, size_t bytes) {
  std::atomic<bool> same(true);
  parallel_ranges(bytes, 1 << 20, [&](size_t a, size_t b) {
    if (memcmp((const char *)x + a, (const char *)y + a, b - a) != 0) same.store(false);
  });
  return same.load();
}

// two pinned slots per context, grown on demand
struct PinnedStage {
  static constexpr size_t kSlot = 32u << 20;
  char *slot[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  int ensure() {
    if (slot[0]) return PHB_OK;
    for (int i = 0; i < 2; ++i) {
      PHB_CUDA(cudaMallocHost((void **)&slot[i], kSlot));
      PHB_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    }
    return PHB_OK;
  }
  void release() {
    for (int i = 0; i < 2; ++i) {
      if (slot[i]) cudaFreeHost(slot[i]);
      if (done[i]) cudaEventDestroy(done[i]);
      slot[i] = nullptr; done[i] = nullptr;
    }
  }
  // pageable host -> device; returns with every byte read from `src` (the caller may free it), DMA possibly in flight
  int upload(void *dev, const void *src, size_t bytes, cudaStream_t st) {
    PHB_CHECK(ensure());
    int k = 0;
    for (size_t off = 0; off < bytes; off += kSlot, k ^= 1) {
      const size_t n = std::min(kSlot, bytes - off);
      PHB_CUDA(cudaEventSynchronize(done[k]));   // the previous DMA out of this slot has finished
      parallel_memcpy(slot[k], (const char *)src + off, n);
      PHB_CUDA(cudaMemcpyAsync((char *)dev + off, slot[k], n, cudaMemcpyHostToDevice, st));
      PHB_CUDA(cudaEventRecord(done[k], st));
    }
    return PHB_OK;
  }
  // device -> pageable host, complete on return
  int download(void *dst, const void *dev, size_t bytes, cudaStream_t st) {
    PHB_CHECK(ensure());
    size_t prevOff = 0, prevN = 0;
    int k = 0;
    for (size_t off = 0; off < bytes; off += kSlot, k ^= 1) {
      const size_t n = std::min(kSlot, bytes - off);
      PHB_CUDA(cudaMemcpyAsync(slot[k], (const char *)dev + off, n, cudaMemcpyDeviceToHost, st));
      PHB_CUDA(cudaEventRecord(done[k], st));
      if (prevN) {                               // drain the other slot while this DMA runs
        PHB_CUDA(cudaEventSynchronize(done[k ^ 1]));
        parallel_memcpy((char *)dst + prevOff, slot[k ^ 1], prevN);
      }
      prevOff = off; prevN = n;
    }
    if (prevN) {
      PHB_CUDA(cudaEventSynchronize(done[k ^ 1]));
      parallel_memcpy((char *)dst + prevOff, slot[k ^ 1], prevN);
    }
    return PHB_OK;
  }
};

}  // namespace phb
