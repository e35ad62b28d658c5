// comm.cuh -- NCCL plumbing (loaded at run time from the torch-bundled libnccl so
// that a single-GPU process never needs it).
#pragma once
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
}  // namespace phb
