// Communicator.h -- one GPU (+ NCCL communicator) per process; the interface
// subset of src/System/Communicator.h:13-140 the hot path's callers use.
#ifndef PHASE_B200_COMMUNICATOR_H
#define PHASE_B200_COMMUNICATOR_H
#include <cstdarg>
#include <cstdio>

#include "Exception.h"

class Communicator {
public:
  // device >= 0: CUDA device; rank/nProcs/id as handed out by the launcher
  explicit Communicator(int device = 0, int rank = 0, int nProcs = 1, const void *ncclId128 = nullptr) {
    phase::check(phb_ctx_create(device, &ctx_), "Communicator", "Communicator");
    if (nProcs > 1) phase::check(phb_ctx_init_comm(ctx_, rank, nProcs, ncclId128), "Communicator", "Communicator");
  }
  Communicator(const Communicator &) = delete;
  Communicator &operator=(const Communicator &) = delete;
  ~Communicator() { phb_ctx_destroy(ctx_); }
  int printf(const char *format, ...) const {
    if (!isMainProc()) return 0;
    va_list ap;
    va_start(ap, format);
    const int n = vprintf(format, ap);
    va_end(ap);
    return n;
  }
  int rank() const { return phb_ctx_rank(ctx_); }
  int nProcs() const { return phb_ctx_nprocs(ctx_); }
  int mainProcNo() const { return 0; }
  bool isMainProc() const { return rank() == mainProcNo(); }
  void barrier() const { phb_ctx_sync(ctx_); }
  phb_ctx *handle() const { return ctx_; }

private:
  phb_ctx *ctx_ = nullptr;
};
#endif
