// solver.cu -- fp64 sliced-ELL SpMV and the fused BiCGStab kernels (K8, K9, K10),
// the device-resident iteration loop (CUDA graph, no host sync per iteration) and
// the SparseMatrixSolver-shaped C entry points (Seam 1).
//
// What this replaces in the reference (all under /root/reference/src/Math):
//   SparseMatrixSolver.h:11-62                 the abstract backend interface
//   EigenSparseMatrixSolver.cpp:28-39,60-64    set() -> triplets -> SparseLU solve
//   TrilinosSparseMatrixSolver.cpp:23-41,71-96 Tpetra maps/CrsMatrix rebuilt per solve
//   TrilinosBelosSparseMatrixSolver.cpp:31-42,44-86  Belos BICGSTAB + Ifpack2, keys
// Algorithm: right-preconditioned BiCGStab (Belos' default "BICGSTAB"); Jacobi is
// folded into the matrix once per solve (A D^-1), so an iteration is 2 SpMV + 3
// fused vector kernels; all dot products are reduced on the device.
#include <algorithm>
#include <cmath>

#include "comm.cuh"
#include "hostcopy.cuh"
#include "kernels.cuh"
#include "peerdev.cuh"
#include "solver.cuh"

using namespace phb;

namespace {

constexpr int kThreads = 256;
constexpr int kBlocksPerSM = 8;

// ------------------------------------------------------------------ SpMV
// One warp per 32-row slice, lane <-> row: every load of col/vals is a fully
// coalesced 128 B / 256 B warp transaction; x is gathered through L1/L2 (banded
// matrices keep the window resident).  NC components share the coefficients.
// EPI 0: y = A x
// EPI 1: y = A x, sigma = (w . y)                       [v = A p, (rhat . v)]
// EPI 2: y = A x, ts,tt,rs,rt,ss = (y.w),(y.y),(w2.w),(w2.y),(w.w)   [t = A M^-1 s; w = s, w2 = rhat]
// EPI 3: y = b - A x, w2 = y, rr = rho0 = (y . y), bb = (b . b)   [initial residual]
template <int NC, int EPI>
__global__ void __launch_bounds__(kThreads)
k_spmv(SellView A, const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y,
       int ld, const double *__restrict__ w, double *w2, KrylovSums *S, int maxIters,
       double *partials, unsigned *ticket, int cur, int localFinish, PeerFuse F, TensorTerm TT) {
  if (EPI == 1 || EPI == 2) {
    if (krylov_done(S, maxIters)) return;
    if (F.on && F.waitHalo) halo_wait_block(F);  // the peers' ghost values of x have landed
  }
  const int lane = threadIdx.x & 31;
  const int warpsPerBlock = blockDim.x >> 5;
  const int warp = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5);
  const int nWarps = gridDim.x * warpsPerBlock;
  double s0 = 0., s1 = 0., s2 = 0., s3 = 0., s4 = 0.;
  for (int slice = warp; slice < A.nSlices; slice += nWarps) {
    const int off = __ldg(A.sliceOff + slice);
    const int wdt = (__ldg(A.sliceOff + slice + 1) - off) >> 5;
    const int row = slice * 32 + lane;
    double acc[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) acc[i] = 0.;
    slice_dot_any<NC>(A.col, vals, (size_t)off + lane, wdt, x, ld, acc);
    if (row < A.nRows) {
      if (NC == 2 && TT.t) {   // (cell, cell) tensor block of SYMMETRY patches
        const int o = TT.perm ? TT.perm[row] : row;
        const double sc = TT.scale ? TT.scale[row] : 1.;
        const double x0 = sc * x[row], x1 = sc * x[(size_t)ld + row];
        acc[0] += TT.t[o] * x0 + TT.t[(size_t)TT.n + o] * x1;
        acc[NC - 1] += TT.t[2 * (size_t)TT.n + o] * x0 + TT.t[3 * (size_t)TT.n + o] * x1;
      }
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const size_t idx = (size_t)i * ld + row;
        if (EPI == 3) {
          const double bi = w[idx];
          const double ri = bi - acc[i];
          y[idx] = ri;
          w2[idx] = ri;
          s0 = fma(ri, ri, s0);
          s1 = fma(bi, bi, s1);
        } else {
          y[idx] = acc[i];
          if (EPI == 1) s0 = fma(w[idx], acc[i], s0);
          if (EPI == 2) {
            const double sv = w[idx], rh = w2[idx];
            s0 = fma(acc[i], sv, s0);
            s1 = fma(acc[i], acc[i], s1);
            s2 = fma(rh, sv, s2);
            s3 = fma(rh, acc[i], s3);
            s4 = fma(sv, sv, s4);
          }
        }
      }
    }
  }
  if (EPI == 1) {
    double v[1] = {s0};
    const bool fin = grid_reduce<1>(v, partials, ticket, &S->sigma);
    if (fin && F.on && F.pushRed) reduce_push_warp(F.pv, F.redCh, &S->sigma, 1);  // my partial -> every peer
  } else if (EPI == 2) {
    double v[5] = {s0, s1, s2, s3, s4};
    if (grid_reduce<5>(v, partials, ticket, &S->ts) && localFinish && lane == 0) krylov_finish(S, cur);
  } else if (EPI == 3) {
    double v[2] = {s0, s1};
    grid_reduce<2>(v, partials, ticket, &S->rr);  // rr, bb adjacent
  }
}

// after the (all-reduced) initial sums: state such that the first fused update
// leaves x and r untouched and sets p = r  (s = r, t = p = v = 0, alpha = omega = beta = 0)
__global__ void k_init_scalars(KrylovSums *S, double tol) {
  S->rho[0] = S->rr;
  S->rho[1] = 1.;
  S->sigma = 1.;
  S->ts = S->tt = S->rs = S->rt = S->ss = 0.;
  S->alpha = S->omega = S->beta = 0.;
  S->thresh = tol * tol * S->bb;
  S->iters = 0.;
}

// fused update opening iteration i: it applies the tail of iteration i-1 and the head of i
//   x += alpha ph + omega sh ;  r = s - omega t ;  p = r + beta (p - omega v)
// (alpha, omega, beta of iteration i-1 from krylov_finish).  One pass: 5-7 reads, 3 writes.
template <int NC, bool PRECOND>
__global__ void __launch_bounds__(kThreads)
k_update_fused(int n, int ld, double *__restrict__ x, double *__restrict__ r, double *__restrict__ p,
               const double *ph, const double *sh, const double *__restrict__ s,
               const double *__restrict__ t, const double *__restrict__ v, const KrylovSums *S, int maxIters,
               PeerFuse F, unsigned *ticket) {
  if (krylov_done(S, maxIters)) return;
  const double alpha = S->alpha, omega = S->omega, beta = S->beta;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const size_t k = (size_t)c * ld + i;
      const double sk = s[k], pk = p[k];
      x[k] += alpha * (PRECOND ? ph[k] : pk) + omega * (PRECOND ? sh[k] : sk);
      const double rk = sk - omega * t[k];
      r[k] = rk;
      p[k] = rk + beta * (pk - omega * v[k]);
    }
  }
  if (F.on && F.pushHalo) {  // the last CTA ships the boundary values of the new p to the peers
    if (last_block(ticket)) halo_push_block(F, p, NC, ld);
  }
}
// x += alpha ph + omega sh of the LAST completed iteration (the loop exits before the next fused update)
template <int NC>
__global__ void k_final_x(int n, int ld, double *__restrict__ x, const double *__restrict__ ph,
                          const double *__restrict__ sh, const KrylovSums *S) {
  const double alpha = S->alpha, omega = S->omega;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const size_t k = (size_t)c * ld + i;
      x[k] += alpha * ph[k] + omega * sh[k];
    }
  }
}

// s = r - alpha v
template <int NC>
__global__ void __launch_bounds__(kThreads)
k_update_s(int n, int ld, const double *__restrict__ r, const double *__restrict__ v,
           double *__restrict__ s, KrylovSums *S, int cur, int maxIters, PeerFuse F, unsigned *ticket) {
  if (krylov_done(S, maxIters)) return;
  double sigma;
  if (F.on && F.waitRed) {  // R1 across the GPUs: sum the peers' partial (rhat . v) in rank order
    double g[1];
    reduce_wait_block<1>(F.pv, F.redCh, g);
    sigma = g[0];
    if (blockIdx.x == 0 && threadIdx.x == 0) S->sigma = sigma;  // for krylov_finish (no reader in this kernel)
  } else {
    sigma = S->sigma;
  }
  const double alpha = S->rho[cur] / sigma;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const size_t k = (size_t)c * ld + i;
      s[k] = r[k] - alpha * v[k];
    }
  }
  if (F.on && F.pushHalo) {
    if (last_block(ticket)) halo_push_block(F, s, NC, ld);
  }
}

// NCCL path: publish the iteration's scalars after the all-reduce of R2
__global__ void k_iter_scalars(KrylovSums *S, int cur, int maxIters) {
  if (krylov_done(S, maxIters)) return;
  krylov_finish(S, cur);
}

// ---------------------------------------------------------------- Jacobi fold
__global__ void k_extract_dinv(SellView A, const double *__restrict__ vals, double *__restrict__ dinv) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= A.nRows) return;
  const int off = A.sliceOff[row >> 5];
  const double d = vals[(size_t)off + (row & 31)];  // entry 0 is the diagonal
  dinv[row] = d != 0. ? 1. / d : 1.;
}
// generic diagonal search for matrices handed over through set_csr
__global__ void k_extract_dinv_search(SellView A, const double *__restrict__ vals, const int *__restrict__ rowLen,
                                      double *__restrict__ dinv) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= A.nRows) return;
  const int off = A.sliceOff[row >> 5];
  double d = 0.;
  for (int k = 0; k < rowLen[row]; ++k) {
    const size_t slot = (size_t)off + (size_t)k * 32 + (row & 31);
    if (A.col[slot] == row) d += vals[slot];
  }
  dinv[row] = d != 0. ? 1. / d : 1.;
}
__global__ void k_scale_cols(long long nSlots, const int *__restrict__ col, const double *__restrict__ vals,
                             const double *__restrict__ dinv, double *__restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nSlots;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = vals[i] * dinv[col[i]];
}
// y = x * d  (d broadcast over components), mode 1: y = x / d
template <int NC>
__global__ void k_diag_apply(int n, int ld, const double *__restrict__ x, const double *__restrict__ d,
                             double *__restrict__ y, int divide) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double di = d[i];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const size_t k = (size_t)c * ld + i;
      y[k] = divide ? x[k] / di : x[k] * di;
    }
  }
}

// sum of b over owned rows (+ row count), and b -= shift: compatibility projection for singular systems
__global__ void __launch_bounds__(kThreads)
k_sum_rows(int n, const double *__restrict__ b, double *partials, unsigned *ticket, double *out) {
  double s0 = 0.;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s0 += b[i];
  double v[2] = {s0, 0.};
  if (blockIdx.x == 0 && threadIdx.x == 0) v[1] = (double)n;
  grid_reduce<2>(v, partials, ticket, out);
}
__global__ void k_shift_rows(int n, double *__restrict__ b, const double *__restrict__ sumCount) {
  const double shift = sumCount[0] / sumCount[1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) b[i] -= shift;
}

// halo pack: sendBuf[c][j] = x[c*ld + sendDev[j]]
__global__ void k_pack(int nSend, int nComp, int ld, const int *__restrict__ sendDev,
                       const double *__restrict__ x, double *__restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nSend * nComp) return;
  const int c = i / nSend, j = i - c * nSend;
  buf[i] = x[(size_t)c * ld + sendDev[j]];
}

__global__ void k_scatter_vals(long long nnz, const int *__restrict__ csr2slot,
                               const double *__restrict__ csrVals, double *__restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nnz;
       i += (long long)gridDim.x * blockDim.x) {
    const int s = csr2slot[i];
    if (s >= 0) out[s] = csrVals[i];
  }
}

int grid_for(const phb_ctx *c, long long work) {
  long long g = (work + kThreads - 1) / kThreads;
  const long long cap = (long long)c->numSMs * kBlocksPerSM;
  return (int)std::max<long long>(1, std::min(g, cap));
}
int spmv_grid(const phb_ctx *c, const SellPattern *P) {
  const long long warps = P->nSlices;
  return grid_for(c, warps * 32);
}

template <int EPI>
void launch_spmv(phb_solver *s, const double *vals, const double *x, double *y, const double *w, double *w2,
                 int cur = 0, int localFinish = 0, const PeerFuse &F = PeerFuse()) {
  const SellPattern *P = s->runPat ? s->runPat : s->pat;
  const SellView A = view_of(P);
  const int grid = spmv_grid(s->ctx, P);
  TensorTerm TT;
  if (s->tens && s->nComp == 2) {
    TT.t = s->tens; TT.n = s->pat->nRows;
    TT.scale = s->precond == PHB_PC_JACOBI && vals == s->scaled.p ? s->dinv.p : nullptr;
    TT.perm = s->runPat == &s->ilu.pat ? s->ilu.new2old.p : nullptr;
  }
  if (s->nComp == 1)
    PHB_LAUNCH(s->ctx, (k_spmv<1, EPI>), grid, kThreads, 0, A, vals, x, y, s->ld, w, w2, s->sums.p, s->maxIters,
               s->partials.p, s->ticket.p, cur, localFinish, F, TT);
  else
    PHB_LAUNCH(s->ctx, (k_spmv<2, EPI>), grid, kThreads, 0, A, vals, x, y, s->ld, w, w2, s->sums.p, s->maxIters,
               s->partials.p, s->ticket.p, cur, localFinish, F, TT);
}

// ghost refresh of a gathered vector before an SpMV (grid_->sendMessages analogue
// inside the Krylov loop; UG/FiniteVolumeGrid2D.tpp:3-49)
bool use_peer(const phb_solver *s) {
  return s->ctx->peer.enabled && s->ctx->nProcs > 1 && s->peerRegion >= 0 && s->halo &&
         (int)s->halo->peerLd.size() == s->ctx->nProcs;
}

int halo_exchange(phb_solver *s, double *x, bool inLoop = false) {
  const phb_mesh *m = s->halo;
  if (!m || s->ctx->nProcs == 1) return PHB_OK;
  const int nSend = (int)m->hSendDev.size();
  if (use_peer(s) && peer_owns(s->ctx, x)) {
    int vec = x == s->p.p ? 0 : x == s->s.p ? 1 : x == s->ph.p ? 2 : 3;
    PeerHalo h;
    h.sendDev = s->runSendDev ? s->runSendDev : m->dSendDev.p;
    for (int q = 0; q < kMaxPeers; ++q) {
      const bool in = q < s->ctx->nProcs;
      h.sendOff[q] = in ? m->hSendOff[q] : 0; h.sendCnt[q] = in ? m->hSendCnt[q] : 0;
      h.recvCnt[q] = in ? m->hRecvCnt[q] : 0;
      h.peerRecvOff[q] = in ? m->peerRecvOff[q] : 0; h.peerLd[q] = in ? m->peerLd[q] : 0;
    }
    return peer_halo(s->ctx, s->peerRegion * 4 + vec, h, x, s->nComp, s->ld, inLoop ? s->sums.p : nullptr,
                     s->maxIters);
  }
  phb_mesh *mm = const_cast<phb_mesh *>(m);
  if (nSend)
    PHB_LAUNCH(s->ctx, k_pack, (nSend * s->nComp + 255) / 256, 256, 0, nSend, s->nComp, s->ld,
               s->runSendDev ? s->runSendDev : m->dSendDev.p, x, mm->dSendBuf.p);
  for (int c = 0; c < s->nComp; ++c)
    PHB_CHECK(comm_exchange(s->ctx, mm->dSendBuf.p + (size_t)c * nSend, m->hSendOff.data(), m->hSendCnt.data(),
                            x + (size_t)c * s->ld, m->hRecvOff.data(), m->hRecvCnt.data()));
  return PHB_OK;
}

// sum-all-reduce of the Krylov sums: peer kernel (one launch, bitwise identical on all ranks) or NCCL
int reduce_sums(phb_solver *s, int which, double *vals, int n, bool inLoop, int finishIter, int cur) {
  if (s->ctx->nProcs == 1) return PHB_OK;
  if (use_peer(s))
    return peer_allreduce(s->ctx, s->peerRegion * 4 + which, vals, n, inLoop ? s->sums.p : nullptr, s->maxIters,
                          finishIter, cur);
  return comm_allreduce_sum(s->ctx, vals, n);
}

// communication duties of one kernel when the peer exchanges are fused into the compute kernels
PeerFuse make_fuse(phb_solver *s, const double *vec, int redWhich) {
  PeerFuse F;
  const phb_mesh *m = s->halo;
  F.on = 1;
  F.pv = peer_view(s->ctx);
  F.redCh = s->peerRegion * 4 + redWhich;
  if (vec) {
    const int vi = vec == s->p.p ? 0 : vec == s->s.p ? 1 : vec == s->ph.p ? 2 : 3;
    F.haloCh = s->peerRegion * 4 + vi;
    F.vecOff = (size_t)((const char *)vec - s->ctx->peer.arena);
  }
  F.halo.sendDev = s->runSendDev ? s->runSendDev : m->dSendDev.p;
  for (int q = 0; q < kMaxPeers; ++q) {
    const bool in = q < s->ctx->nProcs;
    F.halo.sendOff[q] = in ? m->hSendOff[q] : 0; F.halo.sendCnt[q] = in ? m->hSendCnt[q] : 0;
    F.halo.recvCnt[q] = in ? m->hRecvCnt[q] : 0;
    F.halo.peerRecvOff[q] = in ? m->peerRecvOff[q] : 0; F.halo.peerLd[q] = in ? m->peerLd[q] : 0;
  }
  return F;
}

int enqueue_iteration(phb_solver *s, const double *A, int cur) {
  phb_ctx *c = s->ctx;
  const int n = s->pat->nRows, ld = s->ld;
  const int gv = grid_for(c, n);
  const bool amg = s->precond == PHB_PC_AMG;
  const bool ilu = s->precond == PHB_PC_ILU0 || amg;   // explicit preconditioner: p^ = M^-1 p, s^ = M^-1 s
  const bool multi = c->nProcs > 1;
  const bool fused = multi && use_peer(s) && s->peerFused && !amg;   // exchanges ride inside the compute kernels
  double *ph = ilu ? s->ph.p : s->p.p, *sh = ilu ? s->sh.p : s->s.p;
  PeerFuse none, fPushP, fSpmv1, fUpdS, fSpmv2, fPushPh, fPushSh;
  if (fused) {
    fPushP = make_fuse(s, s->p.p, 0); fPushP.pushHalo = 1;              // Jacobi: p leaves with the fused update
    fPushPh = make_fuse(s, ph, 0); fPushPh.pushHalo = 1;                // ILU: p^ leaves with the last sweep
    fSpmv1 = make_fuse(s, ph, 0); fSpmv1.waitHalo = 1; fSpmv1.pushRed = 1;
    fUpdS = make_fuse(s, s->s.p, 0); fUpdS.waitRed = 1; fUpdS.pushHalo = ilu ? 0 : 1;
    fPushSh = make_fuse(s, sh, 0); fPushSh.pushHalo = 1;
    fSpmv2 = make_fuse(s, sh, 1); fSpmv2.waitHalo = 1;
  }
#define FUSED(NCV, PRE)                                                                                          \
  PHB_LAUNCH(c, (k_update_fused<NCV, PRE>), gv, kThreads, 0, n, ld, s->x.p, s->r.p, s->p.p, ph, sh, s->s.p, s->t.p, \
             s->v.p, s->sums.p, s->maxIters, (fused && !ilu) ? fPushP : none, s->ticket.p)
  if (s->nComp == 1) { if (ilu) FUSED(1, true); else FUSED(1, false); }
  else { if (ilu) FUSED(2, true); else FUSED(2, false); }
#undef FUSED
  if (amg) PHB_CHECK(amg_apply(s, s->p.p, ph, true));
  else if (ilu) PHB_CHECK(ilu_apply(s, s->p.p, ph, fused ? &fPushPh : nullptr));   // ph = M^-1 p
  if (!fused) PHB_CHECK(halo_exchange(s, ph, true));
  launch_spmv<1>(s, A, ph, s->v.p, s->rhat.p, nullptr, cur, 0, fused ? fSpmv1 : none);   // v = A ph, R1
  if (!fused) PHB_CHECK(reduce_sums(s, 0, &s->sums.p->sigma, 1, true, 0, cur));
  if (s->nComp == 1)
    PHB_LAUNCH(c, k_update_s<1>, gv, kThreads, 0, n, ld, s->r.p, s->v.p, s->s.p, s->sums.p, cur, s->maxIters,
               fused ? fUpdS : none, s->ticket.p);
  else
    PHB_LAUNCH(c, k_update_s<2>, gv, kThreads, 0, n, ld, s->r.p, s->v.p, s->s.p, s->sums.p, cur, s->maxIters,
               fused ? fUpdS : none, s->ticket.p);
  if (amg) PHB_CHECK(amg_apply(s, s->s.p, sh, true));
  else if (ilu) PHB_CHECK(ilu_apply(s, s->s.p, sh, fused ? &fPushSh : nullptr));   // sh = M^-1 s
  if (!fused) PHB_CHECK(halo_exchange(s, sh, true));
  // t = A sh, R2: (t.s), (t.t), (rhat.s), (rhat.t), (s.s); single GPU: the last CTA finishes the iteration
  launch_spmv<2>(s, A, sh, s->t.p, s->s.p, s->rhat.p, cur, multi ? 0 : 1, fused ? fSpmv2 : none);
  if (multi) {
    if (use_peer(s)) {  // the all-reduce kernel also finishes the iteration
      PHB_CHECK(reduce_sums(s, 1, &s->sums.p->ts, 5, true, 1, cur));
    } else {
      PHB_CHECK(comm_allreduce_sum(c, &s->sums.p->ts, 5));
      PHB_LAUNCH(c, k_iter_scalars, 1, 1, 0, s->sums.p, cur, s->maxIters);
    }
  }
  return PHB_OK;
}

int ensure_vectors(phb_solver *s) {
  const size_t len = (size_t)s->ld * s->nComp;
  phb_ctx *c = s->ctx;
  PHB_CHECK(s->b.alloc(len)); PHB_CHECK(s->x.alloc(len)); PHB_CHECK(s->r.alloc(len));
  PHB_CHECK(s->rhat.alloc(len)); PHB_CHECK(s->v.alloc(len)); PHB_CHECK(s->t.alloc(len));
  // the gathered vectors (p, s and their preconditioned images) live in the peer arena when
  // there is one, so that peers can store their halo values straight into the ghost segments
  if (c->peer.enabled && c->nProcs > 1 && s->peerRegion == -1) {
    // first free region; solvers are created and destroyed in the same order on every rank, so the choice agrees
    s->peerRegion = -2;
    if (c->peer.maxRegions * 4 <= kPeerHaloChannels)
      for (int r = 0; r < c->peer.maxRegions; ++r)
        if (!(c->peer.regionMask & (1u << r))) { c->peer.regionMask |= 1u << r; s->peerRegion = r; break; }
  }
  if (s->peerRegion >= 0 && len * sizeof(double) <= c->peer.vecBytes) {
    s->p.attach(peer_vector(c, s->peerRegion, 0), len); s->s.attach(peer_vector(c, s->peerRegion, 1), len);
    s->ph.attach(peer_vector(c, s->peerRegion, 2), len); s->sh.attach(peer_vector(c, s->peerRegion, 3), len);
  } else {
    if (s->peerRegion >= 0) { c->peer.regionMask &= ~(1u << s->peerRegion); s->peerRegion = -2; }
    PHB_CHECK(s->p.alloc(len)); PHB_CHECK(s->s.alloc(len));
  }
  const size_t nb = (size_t)s->ctx->numSMs * kBlocksPerSM;
  if (s->partials.n != nb * 8) {
    PHB_CHECK(s->partials.alloc(nb * 8));
    PHB_CHECK(s->ticket.alloc(1));
    PHB_CHECK(s->ticket.zero(s->ctx->stream));
    PHB_CHECK(s->sums.alloc(1));
    PHB_CHECK(s->sums.zero(s->ctx->stream));
  }
  return PHB_OK;
}

}  // namespace

namespace phb {

void solver_drop_graph(phb_solver *s) {
  if (s && s->graphExec) { cudaGraphExecDestroy(s->graphExec); s->graphExec = nullptr; }
}

int solver_bind(phb_solver *s, const SellPattern *pat, const double *dVals, int nComp, const phb_mesh *halo) {
  PHB_REQUIRE(nComp == 1 || nComp == 2, "solver: nComp must be 1 or 2");
  const bool resized = (s->pat != pat) || s->nComp != nComp || s->ld != pat->nCols;
  s->pat = pat;
  s->dVals = dVals;
  s->tens = nullptr;   // set by the equation after binding when it carries a tensor block
  s->nComp = nComp;
  s->ld = pat->nCols;
  s->halo = halo;
  if (resized) {
    // vectors are zero-filled once so that ghost / padding entries are finite
    PHB_CHECK(ensure_vectors(s));
    const size_t len = (size_t)s->ld * s->nComp;
    for (DevBuf<double> *v : {&s->b, &s->x, &s->r, &s->rhat, &s->p, &s->v, &s->s, &s->t})
      PHB_CUDA(cudaMemsetAsync(v->p, 0, len * sizeof(double), s->ctx->stream));
  }
  return PHB_OK;
}

int solver_run(phb_solver *s, int *iters, double *relres) {
  phb_ctx *c = s->ctx;
  PHB_REQUIRE(s->pat && s->dVals, "solver: no matrix set");
  PHB_REQUIRE(s->method == "BICGSTAB", "solver \"%s\" is not available (BICGSTAB only)", s->method.c_str());
  const SellPattern *P = s->pat;
  const int n = P->nRows, ld = s->ld;
  const SellView A = view_of(P);
  const int gv = grid_for(c, n);
  const double *Aw = s->dVals;
  s->runPat = P;
  s->runSendDev = nullptr;
  // ---- singular (all-Neumann) scalar systems: make the right-hand side compatible, 1^T b = 0.
  // Round-off in an assembled b (e.g. a converged mass imbalance) otherwise leaves a component the
  // Krylov iteration can never remove.
  if (s->projectConstant && s->nComp == 1) {
    PHB_CHECK(s->proj.alloc(2));
    PHB_LAUNCH(c, k_sum_rows, gv, kThreads, 0, n, s->b.p, s->partials.p, s->ticket.p, s->proj.p);
    PHB_CHECK(comm_allreduce_sum(c, s->proj.p, 2));
    PHB_LAUNCH(c, k_shift_rows, gv, kThreads, 0, n, s->b.p, s->proj.p);
  }
  // ---- ILU(0): iterate on the symmetrically permuted system (rows grouped by independent set)
  if (s->precond == PHB_PC_ILU0) {
    PHB_CHECK(ilu_prepare(s, P, (s->halo && c->nProcs > 1) ? s->halo : nullptr));
    PHB_CHECK(ilu_factor(s, s->dVals));
    const size_t len = (size_t)ld * s->nComp;
    if (s->ph.owned) { PHB_CHECK(s->ph.alloc(len)); PHB_CHECK(s->sh.alloc(len)); }
    PHB_CUDA(cudaMemsetAsync(s->ph.p, 0, len * sizeof(double), c->stream));
    PHB_CUDA(cudaMemsetAsync(s->sh.p, 0, len * sizeof(double), c->stream));
    // b, x0 -> permuted numbering (t and v are free before the first iteration)
    PHB_CHECK(ilu_permute(s, s->b.p, s->t.p, 0));
    PHB_CHECK(ilu_permute(s, s->x.p, s->v.p, 0));
    PHB_CUDA(cudaMemcpyAsync(s->b.p, s->t.p, len * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    PHB_CUDA(cudaMemcpyAsync(s->x.p, s->v.p, len * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    Aw = s->ilu.vals.p;
    s->runPat = &s->ilu.pat;
    s->runSendDev = s->ilu.sendDev.p;
  }
  // ---- AMG: V-cycle on the natural ordering, hierarchy cached across solves
  if (s->precond == PHB_PC_AMG) {
    PHB_CHECK(amg_prepare(s));
    const size_t len = (size_t)ld * s->nComp;
    if (s->ph.owned) { PHB_CHECK(s->ph.alloc(len)); PHB_CHECK(s->sh.alloc(len)); }
    PHB_CUDA(cudaMemsetAsync(s->ph.p, 0, len * sizeof(double), c->stream));
    PHB_CUDA(cudaMemsetAsync(s->sh.p, 0, len * sizeof(double), c->stream));
  }
  // ---- Jacobi fold: iterate on y = D x with A D^-1
  if (s->precond == PHB_PC_JACOBI) {
    PHB_CHECK(s->dinv.alloc((size_t)ld));
    PHB_CHECK(s->scaled.alloc((size_t)P->nSlots));
    if (P == &s->own)
      PHB_LAUNCH(c, k_extract_dinv_search, (n + 255) / 256, 256, 0, A, s->dVals, P->rowLen.p, s->dinv.p);
    else
      PHB_LAUNCH(c, k_extract_dinv, (n + 255) / 256, 256, 0, A, s->dVals, s->dinv.p);
    if (s->halo && c->nProcs > 1) {
      const int keep = s->nComp;
      s->nComp = 1;
      int rc = halo_exchange(s, s->dinv.p);
      s->nComp = keep;
      PHB_CHECK(rc);
    }
    PHB_LAUNCH(c, k_scale_cols, grid_for(c, P->nSlots), kThreads, 0, P->nSlots, P->col.p, s->dVals, s->dinv.p,
               s->scaled.p);
    Aw = s->scaled.p;
    if (s->nComp == 1)
      PHB_LAUNCH(c, k_diag_apply<1>, gv, kThreads, 0, n, ld, s->x.p, s->dinv.p, s->x.p, 1);
    else
      PHB_LAUNCH(c, k_diag_apply<2>, gv, kThreads, 0, n, ld, s->x.p, s->dinv.p, s->x.p, 1);
  }
  KrylovSums *hs = reinterpret_cast<KrylovSums *>(c->pinned);
  int totalIters = 0;
  double rel = 0.;
  for (int attempt = 0; attempt < 3; ++attempt) {
    // ---- r = b - A x, rhat = s = r, p = v = t = 0
    const size_t vbytes = (size_t)ld * s->nComp * sizeof(double);
    PHB_CUDA(cudaMemsetAsync(s->p.p, 0, vbytes, c->stream));
    PHB_CUDA(cudaMemsetAsync(s->v.p, 0, vbytes, c->stream));
    PHB_CUDA(cudaMemsetAsync(s->t.p, 0, vbytes, c->stream));
    PHB_CHECK(halo_exchange(s, s->x.p));
    launch_spmv<3>(s, Aw, s->x.p, s->r.p, s->b.p, s->rhat.p);
    PHB_CUDA(cudaMemcpyAsync(s->s.p, s->r.p, vbytes, cudaMemcpyDeviceToDevice, c->stream));
    PHB_CHECK(reduce_sums(s, 3, &s->sums.p->rr, 2, false, 0, 0));
    PHB_LAUNCH(c, k_init_scalars, 1, 1, 0, s->sums.p, s->tol);
    const int budget = s->maxIters - totalIters;
    if (budget <= 0) break;
    const int saveMax = s->maxIters;
    s->maxIters = budget;  // kernels count iterations of this attempt
    // ---- iterations: graphs of K iterations, polled every `burst` graphs
    // AMG converges in tens of iterations: short graphs, polled every time
    const bool amg = s->precond == PHB_PC_AMG;
    const int K = amg ? 2 : std::max(2, s->itersPerGraph & ~1);
    bool graphOk = s->useGraph;
    const void *key[4] = {s->runPat, Aw,
                          (const void *)(intptr_t)(s->nComp * 1000003 + s->maxIters + 7919 * s->precond + (s->tens ? 104729 : 0)),
                          s->halo};
    if (graphOk && (!s->graphExec || memcmp(key, s->graphKey, sizeof(key)) != 0)) {
      if (s->graphExec) { cudaGraphExecDestroy(s->graphExec); s->graphExec = nullptr; }
      cudaGraph_t g = nullptr;
      const long long before = c->launches;
      PHB_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
      int rc = PHB_OK;
      for (int k = 0; k < K && rc == PHB_OK; ++k) rc = enqueue_iteration(s, Aw, k & 1);
      cudaError_t ce = cudaStreamEndCapture(c->stream, &g);
      c->launches = before;
      if (rc != PHB_OK) { s->maxIters = saveMax; return rc; }
      if (ce != cudaSuccess || cudaGraphInstantiate(&s->graphExec, g, 0) != cudaSuccess) {
        cudaGetLastError();
        graphOk = false;
        s->graphExec = nullptr;
      }
      if (g) cudaGraphDestroy(g);
      memcpy(s->graphKey, key, sizeof(key));
    }
    const int launchesPerIter = 4 + ((s->halo && c->nProcs > 1) ? (use_peer(s) ? (s->peerFused ? 1 : 4) : 3) : 0) +
                                (s->precond == PHB_PC_ILU0 ? 2 * ilu_launches_per_apply(s) : 0) +
                                (amg ? 2 * amg_launches_per_apply(s) : 0);
    int launched = 0, burst = 1;
    // With the multigrid preconditioner a solve is a handful of iterations and every poll leaves the GPU idle for a host
    // round trip plus a graph launch: the first poll waits until one iteration short of what this solver needed last
    // time (kernels past convergence leave on the device-side test, so overshooting costs empty launches only)
    if (amg && graphOk && s->lastIters > K) burst = std::min(8, std::max(1, (s->lastIters - 1) / K));   // at most 16 iterations blind
    bool done = false, midRefresh = false;
    while (!done && launched < budget) {
      for (int g = 0; g < burst && launched < budget; ++g) {
        if (graphOk) {
          PHB_CUDA(cudaGraphLaunch(s->graphExec, c->stream));
          c->launches += (long long)K * launchesPerIter;
        } else {
          for (int k = 0; k < K; ++k) PHB_CHECK(enqueue_iteration(s, Aw, k & 1));
        }
        launched += K;
      }
      PHB_CUDA(cudaMemcpyAsync(hs, s->sums.p, sizeof(KrylovSums), cudaMemcpyDeviceToHost, c->stream));
      PHB_CUDA(cudaStreamSynchronize(c->stream));
      done = !(hs->rr > hs->thresh) || hs->iters >= (double)budget;
      burst = amg ? 1 : std::min(burst * 2, 8);
      // a hierarchy built for other coefficients that has already cost twice its usual iterations: stop here, recompute
      // its values on the device and restart from the current x (`amgRefresh auto`)
      if (!done && amg && amg_wants_refresh(s, totalIters + (int)hs->iters)) { midRefresh = true; break; }
    }
    // the x-update of the last completed iteration is still pending (it rides in the NEXT fused update)
    {
      const bool pre = s->precond == PHB_PC_ILU0 || s->precond == PHB_PC_AMG;
      const double *ph = pre ? s->ph.p : s->p.p;
      const double *sh = pre ? s->sh.p : s->s.p;
      if (s->nComp == 1) PHB_LAUNCH(c, k_final_x<1>, gv, kThreads, 0, n, ld, s->x.p, ph, sh, s->sums.p);
      else PHB_LAUNCH(c, k_final_x<2>, gv, kThreads, 0, n, ld, s->x.p, ph, sh, s->sums.p);
    }
    s->maxIters = saveMax;
    totalIters += (int)hs->iters;
    // ---- true residual (the recursive one can drift); restart if it disagrees
    PHB_CHECK(halo_exchange(s, s->x.p));
    launch_spmv<3>(s, Aw, s->x.p, s->r.p, s->b.p, s->rhat.p);
    PHB_CHECK(reduce_sums(s, 3, &s->sums.p->rr, 2, false, 0, 0));
    PHB_CUDA(cudaMemcpyAsync(hs, s->sums.p, sizeof(KrylovSums), cudaMemcpyDeviceToHost, c->stream));
    PHB_CUDA(cudaStreamSynchronize(c->stream));
    rel = hs->bb > 0. ? std::sqrt(hs->rr / hs->bb) : std::sqrt(hs->rr);
    if (!std::isfinite(rel)) {
      s->lastIters = totalIters;
      s->lastRelres = rel;
      if (iters) *iters = totalIters;
      if (relres) *relres = rel;
      set_error("BiCGStab breakdown (non-finite residual) after %d iterations", totalIters);
      return PHB_ERR_BREAKDOWN;
    }
    if (rel <= s->tol * 1.0000001 || totalIters >= s->maxIters) break;
    if (midRefresh) { PHB_CHECK(amg_refresh_midsolve(s)); --attempt; }
  }
  if (s->precond == PHB_PC_JACOBI) {
    if (s->nComp == 1)
      PHB_LAUNCH(c, k_diag_apply<1>, gv, kThreads, 0, n, ld, s->x.p, s->dinv.p, s->x.p, 0);
    else
      PHB_LAUNCH(c, k_diag_apply<2>, gv, kThreads, 0, n, ld, s->x.p, s->dinv.p, s->x.p, 0);
  }
  if (s->precond == PHB_PC_ILU0) {  // back to the caller's numbering
    PHB_CHECK(ilu_permute(s, s->x.p, s->v.p, 1));
    PHB_CUDA(cudaMemcpyAsync(s->x.p, s->v.p, (size_t)ld * s->nComp * sizeof(double), cudaMemcpyDeviceToDevice,
                             c->stream));
  }
  s->runPat = P;
  s->runSendDev = nullptr;
  s->lastIters = totalIters;
  s->lastRelres = rel;
  if (s->precond == PHB_PC_AMG) { amg_record_iters(s, totalIters); PHB_CHECK(amg_check(s)); }
  if (iters) *iters = totalIters;
  if (relres) *relres = rel;
  return launch_status(c);
}

}  // namespace phb

// ------------------------------------------------------------------ C ABI
extern "C" {

int phb_solver_create(phb_ctx *ctx, phb_solver **out) {
  PHB_REQUIRE(ctx && out, "phb_solver_create: NULL argument");
  if (ctx->device < 0) { set_error("phb_solver_create: host-only context"); return PHB_ERR_STATE; }
  phb_solver *s = new phb_solver();
  s->ctx = ctx;
  if (getenv("PHB_NO_GRAPH")) s->useGraph = false;
  ctx->solvers.push_back(s);
  *out = s;
  return PHB_OK;
}

int phb_solver_destroy(phb_solver *s) {
  if (!s) return PHB_OK;
  auto &live = s->ctx->solvers;
  live.erase(std::remove(live.begin(), live.end(), s), live.end());
  if (s->peerRegion >= 0) s->ctx->peer.regionMask &= ~(1u << s->peerRegion);
  if (s->graphExec) cudaGraphExecDestroy(s->graphExec);
  delete s;
  return PHB_OK;
}

int phb_solver_setup(phb_solver *s, const char *key, const char *value) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(s && key && value, "phb_solver_setup: NULL argument");
  std::string k(key), v(value), lv(value);
  std::transform(lv.begin(), lv.end(), lv.begin(), ::tolower);
  if (k == "maxIters") {
    s->maxIters = std::stoi(v);
    PHB_REQUIRE(s->maxIters > 0, "maxIters must be positive");
  } else if (k == "tolerance") {
    s->tol = std::stod(v);
    PHB_REQUIRE(s->tol > 0., "tolerance must be positive");
  } else if (k == "solver") {
    std::string u(v);
    std::transform(u.begin(), u.end(), u.begin(), ::toupper);
    PHB_REQUIRE(u == "BICGSTAB", "solver \"%s\" is not available (BICGSTAB only)", value);
    s->method = u;
  } else if (k == "preconditioner" || k == "innerPreconditioner") {
    if (lv == "none") s->precond = PHB_PC_NONE;
    else if (lv == "jacobi" || lv == "diagonal") s->precond = PHB_PC_JACOBI;
    else if (lv == "ilu0" || lv == "riluk" || lv == "schwarz" || lv == "ilu") s->precond = PHB_PC_ILU0;
    else if (lv == "amg" || lv == "muelu" || lv == "sa" || lv == "multigrid") s->precond = PHB_PC_AMG;
    else PHB_REQUIRE(false, "unknown preconditioner \"%s\"", value);
  } else if (k == "ordering" || k == "iluOrdering") {
    if (lv == "multicolor" || lv == "multicolour" || lv == "colour" || lv == "color") s->iluOrdering = 0;
    else if (lv == "levels" || lv == "natural" || lv == "wavefront") s->iluOrdering = 1;
    else PHB_REQUIRE(false, "unknown ILU ordering \"%s\" (multicolor | levels)", value);
  } else if (k == "amgTheta") {
    s->amg.theta = std::stod(v); s->amg.built = false;
    PHB_REQUIRE(s->amg.theta >= 0. && s->amg.theta < 1., "amgTheta must lie in [0, 1)");
  } else if (k == "amgAggTheta") {
    s->amg.thetaAgg = std::stod(v); s->amg.built = false;
    PHB_REQUIRE(s->amg.thetaAgg >= 0. && s->amg.thetaAgg <= 1., "amgAggTheta must lie in [0, 1]");
  } else if (k == "amgCoarsest") {
    s->amg.coarsest = std::stoi(v); s->amg.built = false;
    PHB_REQUIRE(s->amg.coarsest >= 1 && s->amg.coarsest <= 1024, "amgCoarsest must lie in 1..1024 (dense coarsest solve)");
  } else if (k == "amgSweeps") {
    s->amg.nu = std::stoi(v);
    PHB_REQUIRE(s->amg.nu >= 1 && s->amg.nu <= 4, "amgSweeps must lie in 1..4");
    if (s->graphExec) { cudaGraphExecDestroy(s->graphExec); s->graphExec = nullptr; }
  } else if (k == "amgSmootherWeight") {
    s->amg.omegaS = std::stod(v); s->amg.built = false;
    PHB_REQUIRE(s->amg.omegaS > 0. && s->amg.omegaS < 2., "amgSmootherWeight must lie in (0, 2)");
  } else if (k == "amgCoarseSmootherWeight") {
    s->amg.omegaC = std::stod(v); s->amg.built = false;
    PHB_REQUIRE(s->amg.omegaC >= 0. && s->amg.omegaC < 2., "amgCoarseSmootherWeight must lie in [0, 2)");
  } else if (k == "amgPrecision") {
    PHB_REQUIRE(lv == "single" || lv == "double" || lv == "float", "amgPrecision must be \"single\" or \"double\"");
    s->amg.single = lv != "double";
  } else if (k == "amgScope") {
    PHB_REQUIRE(lv == "global" || lv == "local", "amgScope must be \"global\" or \"local\"");
    s->amg.global = lv == "global"; s->amg.built = false;
  } else if (k == "amgFuseRows") {
    s->amg.fuseRows = std::stoll(v); s->amg.built = false;
    PHB_REQUIRE(s->amg.fuseRows >= 0, "amgFuseRows must not be negative");
  } else if (k == "amgTailRows") {
    s->amg.tailRows = std::stoll(v); s->amg.built = false;
    PHB_REQUIRE(s->amg.tailRows >= 1, "amgTailRows must be positive");
  } else if (k == "amgRebuild") {
    PHB_REQUIRE(lv == "auto" || lv == "always", "amgRebuild must be \"auto\" or \"always\"");
    s->amg.rebuildAlways = lv == "always";
  } else if (k == "amgRefresh") {  // numeric re-setup on the device when the coefficients change on a fixed pattern
    PHB_REQUIRE(lv == "off" || lv == "auto" || lv == "always", "amgRefresh must be \"off\", \"auto\" or \"always\"");
    s->amg.refreshMode = lv == "off" ? 0 : lv == "auto" ? 1 : 2;
    s->amg.built = false;
  } else if (k == "peerFusion") {
    s->peerFused = std::stoi(v) != 0;
    if (s->graphExec) { cudaGraphExecDestroy(s->graphExec); s->graphExec = nullptr; }
  } else if (k == "nullSpace") {
    PHB_REQUIRE(lv == "constant" || lv == "none", "nullSpace must be \"constant\" or \"none\"");
    s->projectConstant = lv == "constant";
  } else if (k == "iluFill") {
    PHB_REQUIRE(std::stod(v) == 0., "only iluFill 0 is supported");
  } else if (k == "itersPerGraph") {
    s->itersPerGraph = std::max(2, std::stoi(v));
    if (s->graphExec) { cudaGraphExecDestroy(s->graphExec); s->graphExec = nullptr; }
  } else if (k == "useGraph") {
    s->useGraph = std::stoi(v) != 0;
  } else if (k == "lib" || k == "schwarzIters" || k == "schwarzCombineMode" || k == "schwarzOverlap") {
    // accepted for case-file compatibility; rank-local preconditioning = overlap 0
  } else {
    PHB_REQUIRE(false, "phb_solver_setup: unknown key \"%s\"", key);
  }
  return PHB_OK;
  PHB_TRY_END
}

int phb_solver_set_rank(phb_solver *s, int nRows, int nCols) {
  PHB_REQUIRE(s && nRows >= 0 && nCols >= 0, "phb_solver_set_rank: bad argument");
  s->nRows = nRows;
  s->nColsGlobal = nCols;
  return PHB_OK;
}

int phb_solver_set_halo(phb_solver *s, const phb_mesh *m, int nComp) {
  PHB_REQUIRE(s && m, "phb_solver_set_halo: NULL argument");
  PHB_REQUIRE(nComp == 1, "phb_solver_set_halo: host-assembled distributed systems support nComp = 1");
  s->halo = m;
  return PHB_OK;
}

// rows local, columns global, -1 = trailing padding (M/EigenSparseMatrixSolver.cpp:33-36)
int phb_solver_set_csr(phb_solver *s, int nRows, const int *rowPtr, const int *colInd, const double *vals) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(s && rowPtr && colInd && vals && nRows > 0, "phb_solver_set_csr: bad argument");
  phb_ctx *c = s->ctx;
  const long long nnzIn = rowPtr[nRows];
  // exact comparison with the cached pattern (value-dependent patterns are the rule behind this seam: `+=` drops
  // exact zeros), spread over the host cores: 96 MB at 4M rows
  const bool samePattern = s->haveMatrix && (int)s->cRowPtr.size() == nRows + 1 &&
                           (long long)s->cColInd.size() == nnzIn &&
                           parallel_equal(s->cRowPtr.data(), rowPtr, (nRows + 1) * sizeof(int)) &&
                           parallel_equal(s->cColInd.data(), colInd, nnzIn * sizeof(int));
  if (!c->stage) c->stage = new PinnedStage();
  if (!samePattern) {
    s->cRowPtr.assign(rowPtr, rowPtr + nRows + 1);
    s->cColInd.assign(colInd, colInd + nnzIn);
    const phb_mesh *hm = (c->nProcs > 1) ? s->halo : nullptr;
    const int rowOffset = hm ? hm->rowOffset : 0;
    int nCols = nRows;
    std::vector<std::pair<int, int>> ghostMap;  // (global row, dev col)
    if (hm) {
      PHB_REQUIRE(hm->nLocal == nRows, "phb_solver_set_csr: nRows %d != mesh nLocal %d", nRows, hm->nLocal);
      nCols = hm->nDev;
      for (int cell = 0; cell < hm->nCells; ++cell)
        if (hm->owner[cell] != hm->rank) ghostMap.push_back({hm->globalRow[cell], hm->cell2dev[cell]});
      std::sort(ghostMap.begin(), ghostMap.end());
    }
    SellPattern &S = s->own;
    S.nRows = nRows; S.nCols = nCols;
    S.nSlices = (nRows + 31) / 32;
    S.hRowLen.assign(nRows, 0);
    S.nnz = 0;
    for (int r = 0; r < nRows; ++r) {
      int len = 0;
      for (int j = rowPtr[r]; j < rowPtr[r + 1]; ++j) len += colInd[j] >= 0;
      S.hRowLen[r] = len;
      S.nnz += len;
    }
    S.hSliceOff.assign(S.nSlices + 1, 0);
    for (int sl = 0; sl < S.nSlices; ++sl) {
      int w = 1;
      for (int r = sl * 32; r < std::min(nRows, sl * 32 + 32); ++r) w = std::max(w, S.hRowLen[r]);
      S.hSliceOff[sl + 1] = S.hSliceOff[sl] + w * 32;
    }
    S.nSlots = S.hSliceOff[S.nSlices];
    S.hCol.resize(S.nSlots);
    std::vector<int> c2s(std::max<long long>(nnzIn, 1), -1);
    for (int sl = 0; sl < S.nSlices; ++sl) {
      const int w = (S.hSliceOff[sl + 1] - S.hSliceOff[sl]) / 32;
      for (int lane = 0; lane < 32; ++lane) {
        const int r = sl * 32 + lane;
        const int pad = r < nRows ? r : nRows - 1;
        int k = 0;
        if (r < nRows)
          for (int j = rowPtr[r]; j < rowPtr[r + 1]; ++j) {
            const int g = colInd[j];
            if (g < 0) continue;
            int col;
            if (g >= rowOffset && g < rowOffset + nRows) col = g - rowOffset;
            else {
              auto it = std::lower_bound(ghostMap.begin(), ghostMap.end(), std::make_pair(g, -1));
              PHB_REQUIRE(it != ghostMap.end() && it->first == g,
                          "phb_solver_set_csr: column %d of row %d is neither owned nor a ghost", g, r);
              col = it->second;
            }
            const size_t slot = (size_t)S.hSliceOff[sl] + (size_t)k * 32 + lane;
            S.hCol[slot] = col;
            c2s[j] = (int)slot;
            ++k;
          }
        for (; k < w; ++k) S.hCol[(size_t)S.hSliceOff[sl] + (size_t)k * 32 + lane] = pad;
      }
    }
    PHB_CHECK(S.sliceOff.upload(S.hSliceOff, c->stream));
    PHB_CHECK(S.rowLen.upload(S.hRowLen, c->stream));
    PHB_CHECK(S.col.upload(S.hCol, c->stream));
    PHB_CHECK(s->csr2slot.upload(c2s, c->stream));
    PHB_CHECK(s->ownVals.alloc((size_t)S.nSlots));
    PHB_CHECK(s->ownVals.zero(c->stream));
    if (s->graphExec) { cudaGraphExecDestroy(s->graphExec); s->graphExec = nullptr; }
    // `own` is rebuilt in place (same address, possibly the same slot count): everything derived from the old
    // columns -- the ILU ordering / slot map and the multigrid hierarchy -- is stale
    s->ilu.src = nullptr;
    s->amg.built = false;
  }
  PHB_CHECK(s->csrVals.alloc((size_t)nnzIn));
  PHB_CHECK(c->stage->upload(s->csrVals.p, vals, (size_t)nnzIn * sizeof(double), c->stream));
  PHB_LAUNCH(c, k_scatter_vals, grid_for(c, nnzIn), kThreads, 0, nnzIn, s->csr2slot.p, s->csrVals.p, s->ownVals.p);
  PHB_CUDA(cudaStreamSynchronize(c->stream));  // caller may free its vectors now
  s->haveMatrix = true;
  s->nRows = nRows;
  PHB_CHECK(phb::solver_bind(s, &s->own, s->ownVals.p, 1, (c->nProcs > 1) ? s->halo : nullptr));
  return PHB_OK;
  PHB_TRY_END
}

// duplicates are summed (M/SparseMatrixSolver.cpp:5-31)
int phb_solver_set_coo(phb_solver *s, int nRows, long long nEntries, const int *rows, const int *cols,
                       const double *vals) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(s && rows && cols && vals && nRows > 0 && nEntries >= 0, "phb_solver_set_coo: bad argument");
  std::vector<std::vector<std::pair<int, double>>> R(nRows);
  for (long long i = 0; i < nEntries; ++i) {
    PHB_REQUIRE(rows[i] >= 0 && rows[i] < nRows, "phb_solver_set_coo: row %d out of range", rows[i]);
    auto &row = R[rows[i]];
    bool hit = false;
    for (auto &e : row)
      if (e.first == cols[i]) { e.second += vals[i]; hit = true; break; }
    if (!hit) row.push_back({cols[i], vals[i]});
  }
  std::vector<int> rp(nRows + 1, 0), ci;
  std::vector<double> va;
  for (int r = 0; r < nRows; ++r) {
    for (auto &e : R[r]) { ci.push_back(e.first); va.push_back(e.second); }
    rp[r + 1] = (int)ci.size();
  }
  if (ci.empty()) { ci.push_back(-1); va.push_back(0.); }
  return phb_solver_set_csr(s, nRows, rp.data(), ci.data(), va.data());
  PHB_TRY_END
}

int phb_solver_set_rhs(phb_solver *s, const double *b, int n) {
  PHB_REQUIRE(s && b, "phb_solver_set_rhs: NULL argument");
  PHB_REQUIRE(s->haveMatrix && s->pat == &s->own, "phb_solver_set_rhs: call phb_solver_set_csr first");
  PHB_REQUIRE(n == s->own.nRows, "phb_solver_set_rhs: size %d != rank %d", n, s->own.nRows);
  if (!s->ctx->stage) s->ctx->stage = new PinnedStage();
  PHB_CHECK(s->ctx->stage->upload(s->b.p, b, (size_t)n * sizeof(double), s->ctx->stream));
  PHB_CUDA(cudaStreamSynchronize(s->ctx->stream));
  s->haveRhs = true;
  return PHB_OK;
}

int phb_solver_set_guess(phb_solver *s, const double *x0, int n) {
  PHB_REQUIRE(s && x0, "phb_solver_set_guess: NULL argument");
  PHB_REQUIRE(s->haveMatrix && s->pat == &s->own, "phb_solver_set_guess: call phb_solver_set_csr first");
  PHB_REQUIRE(n == s->own.nRows, "phb_solver_set_guess: size %d != rank %d", n, s->own.nRows);
  if (!s->ctx->stage) s->ctx->stage = new PinnedStage();
  PHB_CHECK(s->ctx->stage->upload(s->x.p, x0, (size_t)n * sizeof(double), s->ctx->stream));
  PHB_CUDA(cudaStreamSynchronize(s->ctx->stream));
  s->haveGuess = true;
  return PHB_OK;
}

int phb_solver_solve(phb_solver *s, int *iters, double *relres) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(s, "phb_solver_solve: solver is NULL");
  if (!s->haveMatrix || !s->haveRhs) {
    set_error("phb_solver_solve: matrix and right-hand side must be set first");
    return PHB_ERR_STATE;
  }
  PHB_CHECK(phb::solver_bind(s, &s->own, s->ownVals.p, 1, (s->ctx->nProcs > 1) ? s->halo : nullptr));
  if (!s->haveGuess)  // no guess is passed on the reference path: start from 0
    PHB_CUDA(cudaMemsetAsync(s->x.p, 0, (size_t)s->ld * sizeof(double), s->ctx->stream));
  s->haveGuess = false;
  int rc = phb::solver_run(s, iters, relres);
  if (rc != PHB_OK) return rc;
  s->hostX.resize(s->own.nRows);
  if (!s->ctx->stage) s->ctx->stage = new PinnedStage();
  PHB_CHECK(s->ctx->stage->download(s->hostX.data(), s->x.p, (size_t)s->own.nRows * sizeof(double), s->ctx->stream));
  return PHB_OK;
  PHB_TRY_END
}

int phb_solver_get_x(const phb_solver *s, double *x, int n) {
  PHB_REQUIRE(s && x, "phb_solver_get_x: NULL argument");
  PHB_REQUIRE(n == (int)s->hostX.size(), "phb_solver_get_x: size %d != rank %d", n, (int)s->hostX.size());
  parallel_memcpy(x, s->hostX.data(), (size_t)n * sizeof(double));
  return PHB_OK;
}

int phb_solver_spmv(phb_solver *s, const double *x, double *y, int n) {
  PHB_REQUIRE(s && x && y, "phb_solver_spmv: NULL argument");
  PHB_REQUIRE(s->haveMatrix, "phb_solver_spmv: no matrix set");
  PHB_CHECK(phb::solver_bind(s, &s->own, s->ownVals.p, 1, (s->ctx->nProcs > 1) ? s->halo : nullptr));
  PHB_REQUIRE(n == s->own.nRows, "phb_solver_spmv: size mismatch");
  cudaStream_t st = s->ctx->stream;
  PHB_CUDA(cudaMemcpyAsync(s->p.p, x, n * sizeof(double), cudaMemcpyHostToDevice, st));
  PHB_CHECK(halo_exchange(s, s->p.p));
  launch_spmv<0>(s, s->dVals, s->p.p, s->v.p, nullptr, nullptr);
  PHB_CUDA(cudaMemcpyAsync(y, s->v.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  PHB_CUDA(cudaStreamSynchronize(st));
  return PHB_OK;
}

// z = M^-1 r with the preconditioner of the last solve (hierarchy / factors as they stand), host vectors over the
// owned rows: lets tests compare ONE application of the device V-cycle with a transcription of the same hierarchy
int phb_solver_apply_preconditioner(phb_solver *s, const double *r, double *z, int n) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(s && r && z, "phb_solver_apply_preconditioner: NULL argument");
  PHB_REQUIRE(s->pat && s->dVals && s->lastIters >= 0 && s->ph.p && s->p.p,
              "phb_solver_apply_preconditioner: solve once first");
  PHB_REQUIRE(n == s->pat->nRows * s->nComp, "phb_solver_apply_preconditioner: size %d != %d", n, s->pat->nRows * s->nComp);
  PHB_REQUIRE(s->precond == PHB_PC_AMG && s->amg.built, "phb_solver_apply_preconditioner: multigrid preconditioner only");
  cudaStream_t st = s->ctx->stream;
  const int nr = s->pat->nRows;
  for (int c = 0; c < s->nComp; ++c)
    PHB_CUDA(cudaMemcpyAsync(s->p.p + (size_t)c * s->ld, r + (size_t)c * nr, nr * sizeof(double), cudaMemcpyHostToDevice, st));
  PHB_CHECK(amg_apply(s, s->p.p, s->ph.p, false));
  for (int c = 0; c < s->nComp; ++c)
    PHB_CUDA(cudaMemcpyAsync(z + (size_t)c * nr, s->ph.p + (size_t)c * s->ld, nr * sizeof(double), cudaMemcpyDeviceToHost, st));
  PHB_CUDA(cudaStreamSynchronize(st));
  return launch_status(s->ctx);
  PHB_TRY_END
}

int phb_solver_time_spmv(phb_solver *s, int reps, double *ms) {
  PHB_REQUIRE(s && ms && reps > 0, "phb_solver_time_spmv: bad argument");
  PHB_REQUIRE(s->pat && s->dVals, "phb_solver_time_spmv: no matrix bound");
  cudaStream_t st = s->ctx->stream;
  cudaEvent_t e0, e1;
  PHB_CUDA(cudaEventCreate(&e0));
  PHB_CUDA(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch_spmv<0>(s, s->dVals, s->p.p, s->v.p, nullptr, nullptr);
  PHB_CUDA(cudaEventRecord(e0, st));
  for (int i = 0; i < reps; ++i) launch_spmv<0>(s, s->dVals, s->p.p, s->v.p, nullptr, nullptr);
  PHB_CUDA(cudaEventRecord(e1, st));
  PHB_CUDA(cudaEventSynchronize(e1));
  float t = 0.f;
  PHB_CUDA(cudaEventElapsedTime(&t, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms = (double)t / reps;
  return PHB_OK;
}

// algorithmic bytes: SpMV 12 nnz + 4 (n+1) + 16 n per component;
// BiCGStab iteration = 2 SpMV + 2 preconditioner applies + the fused vector passes
int phb_solver_bytes(const phb_solver *s, double out[2]) {
  PHB_REQUIRE(s && out && s->pat, "phb_solver_bytes: no matrix bound");
  const double nnz = (double)s->pat->nnz, n = (double)s->pat->nRows, nc = s->nComp;
  out[0] = 12. * nnz + 4. * (n + 1.) + 16. * n * nc;
  // ILU(0) apply: L and U sweeps over the pattern once, 12 nnz + 8 (n+1) + 8 n + 32 n nc
  const double prec = s->precond == PHB_PC_ILU0 ? 12. * nnz + 8. * (n + 1.) + 8. * n + 32. * n * nc
                      : s->precond == PHB_PC_AMG ? phb::amg_cycle_bytes(s) : 0.;
  // vector passes: fused update 64 n (80 n with a separate preconditioned image) + s-update 24 n, per component
  const bool pre = s->precond == PHB_PC_ILU0 || s->precond == PHB_PC_AMG;
  out[1] = 2. * out[0] + 2. * prec + (pre ? 104. : 88.) * n * nc;
  return PHB_OK;
}

}  // extern "C"
