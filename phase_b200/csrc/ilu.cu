// ilu.cu -- ILU(0) preconditioner (K10): symbolic analysis on the host, numeric
// factorisation and the two triangular sweeps on the device, scheduled by
// independent-set blocks ("levels").
//
// Replaces what the reference gets from Ifpack2: additive Schwarz (overlap 0) with
// an inner RILUK, fill level 0 (M/TrilinosBelosSparseMatrixSolver.cpp:63-83).  The
// preconditioner is rank-local: ghost columns are ignored, which is exactly
// Schwarz overlap 0.
//
// Scheduling.  Rows are split into ordered independent sets (no two rows of a set
// are adjacent), either by greedy multicolouring ("multicolor": 2 sets for quads,
// 3-4 for triangles) or by the wavefront levels of the given ordering ("levels":
// the natural-order ILU(0), ~2 sqrt(N) sets on grid-like meshes).  The system is
// symmetrically permuted so that every set is a contiguous row range; L holds the
// entries whose column lies in an earlier set, U those in a later one.  One launch
// per set and sweep; sets without lower (upper) entries need no matrix read.
#include <algorithm>
#include <numeric>

#include "kernels.cuh"
#include "solver.cuh"

using namespace phb;

namespace {

constexpr int kThreads = 256;

struct IluView {
  const int *sliceOff, *col, *rowLen, *diagK, *blockOf;
  const signed char *kind;
  int nRows;
};

__device__ __forceinline__ size_t slot_of(const int *sliceOff, int row, int k) {
  return (size_t)sliceOff[row >> 5] + (size_t)k * 32 + (row & 31);
}

__global__ void k_permute_vals(long long nSlots, const int *__restrict__ slotMap, const double *__restrict__ src,
                               double *__restrict__ a, double *__restrict__ lu) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nSlots;
       i += (long long)gridDim.x * blockDim.x) {
    const int o = slotMap[i];
    const double v = o >= 0 ? src[o] : 0.;
    a[i] = v;
    lu[i] = v;
  }
}

// numeric ILU(0) of the rows [r0, r1) of one set: IKJ with the lower entries taken
// in ascending set order; rows of earlier sets are final.
__global__ void k_ilu_factor(IluView V, double *__restrict__ lu, int r0, int r1) {
  const int i = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r1) return;
  const int len = V.rowLen[i];
  unsigned long long done = 0ull;
  for (;;) {
    int e = -1, eb = 0x7fffffff;
    for (int k = 0; k < len && k < 64; ++k) {
      if ((done >> k) & 1ull) continue;
      const size_t sl = slot_of(V.sliceOff, i, k);
      if (V.kind[sl] != 1) continue;
      const int b = V.blockOf[V.col[sl]];
      if (b < eb) { eb = b; e = k; }
    }
    if (e < 0) break;
    done |= 1ull << e;
    const size_t se = slot_of(V.sliceOff, i, e);
    const int kr = V.col[se];
    const double lik = lu[se] / lu[slot_of(V.sliceOff, kr, V.diagK[kr])];
    lu[se] = lik;
    const int klen = V.rowLen[kr];
    for (int f = 0; f < klen; ++f) {
      const size_t sf = slot_of(V.sliceOff, kr, f);
      if (V.kind[sf] != 3) continue;
      const int j = V.col[sf];
      for (int g = 0; g < len; ++g) {
        const size_t sg = slot_of(V.sliceOff, i, g);
        if (V.kind[sg] != 0 && V.col[sg] == j) { lu[sg] -= lik * lu[sf]; break; }
      }
    }
  }
}

// The first set has no lower entries and the last no upper ones, so y_0 = r_0 is never
// materialised (readers take r for columns < n0) and the last forward launch divides by the
// diagonal straight away: 2 (nSets - 1) launches per apply, the matrix is read once.
// Same access scheme as the SpMV: warp <-> 32-row slice, lane <-> row, coalesced col/val loads
// issued four entries ahead; lower / upper membership follows from the column alone
// (rows of a set are a contiguous range [r0, r1) of the permuted numbering):
//   lower  <=> col <  r0          upper <=> r1 <= col < nRows        (col >= nRows: ghost, ignored)
//
// MODE 0: forward, y_i = r_i - sum_lower l_ik y_k            MODE 1: same, then z_i = y_i / u_ii (last set)
// MODE 2: backward, z_i = (y_i - sum_upper u_ij z_j) / u_ii  MODE 3: same with y_i = r_i (first set)
template <int NC, int MODE>
__global__ void __launch_bounds__(kThreads)
k_ilu_sweep(IluView V, const double *__restrict__ lu, const double *__restrict__ r, double *z, int ld, int r0,
            int r1, int n0, PeerFuse F, unsigned *ticket) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int s0 = r0 >> 5, s1 = (r1 - 1) >> 5;
  for (int slice = s0 + blockIdx.x * wpb + (threadIdx.x >> 5); slice <= s1; slice += gridDim.x * wpb) {
    const int off = __ldg(V.sliceOff + slice);
    const int w = (__ldg(V.sliceOff + slice + 1) - off) >> 5;
    const int row = slice * 32 + lane;
    const bool active = row >= r0 && row < r1;
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c)
      acc[c] = active ? ((MODE == 2) ? z[(size_t)c * ld + row] : r[(size_t)c * ld + row]) : 0.;
    const size_t base = (size_t)off + lane;
    for (int k0 = 0; k0 < w; k0 += 4) {
      int cj[4];
      double aj[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool in = k0 + q < w;
        cj[q] = in ? ld_stream(V.col + base + (size_t)(k0 + q) * 32) : -1;
        aj[q] = in ? ld_stream(lu + base + (size_t)(k0 + q) * 32) : 0.;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = cj[q];
        const bool use = (MODE < 2) ? (j >= 0 && j < r0) : (j >= r1 && j < V.nRows);
        if (use && active) {
          const double *src = (MODE < 2 && j < n0) ? r : z;
#pragma unroll
          for (int c = 0; c < NC; ++c) acc[c] -= aj[q] * src[(size_t)c * ld + j];
        }
      }
    }
    if (active) {
      double d = 1.;
      if (MODE != 0) d = lu[base + (size_t)V.diagK[row] * 32];
#pragma unroll
      for (int c = 0; c < NC; ++c) z[(size_t)c * ld + row] = (MODE != 0) ? acc[c] / d : acc[c];
    }
  }
  if (F.on && F.pushHalo) {  // final launch of an apply: the last CTA ships z's boundary values to the peers
    if (last_block(ticket)) halo_push_block(F, z, NC, ld);
  }
}

// y[c][new] = x[c][old] (gather, dir 0) or y[c][old] = x[c][new] (scatter, dir 1) over owned rows
__global__ void k_permute_vec(int n, int nc, int ld, const int *__restrict__ new2old,
                              const double *__restrict__ x, double *__restrict__ y, int dir) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)n * nc;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t / n), i = (int)(t - (long long)c * n);
    const int o = new2old[i];
    if (dir == 0) y[(size_t)c * ld + i] = x[(size_t)c * ld + o];
    else y[(size_t)c * ld + o] = x[(size_t)c * ld + i];
  }
}

IluView view_of(const IluData &D) {
  IluView v;
  v.sliceOff = D.pat.sliceOff.p; v.col = D.pat.col.p; v.rowLen = D.pat.rowLen.p;
  v.diagK = D.diagK.p; v.blockOf = D.blockOf.p; v.kind = D.kind.p; v.nRows = D.pat.nRows;
  return v;
}

}  // namespace

namespace phb {

// symbolic phase, cached per (pattern, ordering)
int ilu_prepare(phb_solver *s, const SellPattern *P, const phb_mesh *halo) {
  IluData &D = s->ilu;
  if (D.src == P && D.ordering == s->iluOrdering && D.srcSlots == P->nSlots) return PHB_OK;
  cudaStream_t st = s->ctx->stream;
  const int n = P->nRows;
  auto oslot = [&](int r, int k) { return (size_t)P->hSliceOff[r >> 5] + (size_t)k * 32 + (r & 31); };
  // ordered independent sets.  Host-assembled patterns (Seam 1, immersed-boundary stencils) need not be
  // structurally symmetric: the sets are built on the pattern of A + A^T (`tPtr/tInd` = for row j the rows
  // i < j that hold an entry (i, j)), so that two coupled rows never share a set and L / U stay consistent.
  std::vector<int> tPtr, tInd;
  if (P == &s->own) {
    int maxLen = 0;
    for (int i = 0; i < n; ++i) maxLen = std::max(maxLen, P->hRowLen[i]);
    if (maxLen > 64) {
      set_error("ilu0: a row holds %d entries; the factorisation kernel handles at most 64 per row", maxLen);
      return PHB_ERR_UNSUPPORTED;
    }
    tPtr.assign(n + 1, 0);
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < P->hRowLen[i]; ++k) {
        const int j = P->hCol[oslot(i, k)];
        if (j > i && j < n) tPtr[j + 1]++;
      }
    for (int i = 0; i < n; ++i) tPtr[i + 1] += tPtr[i];
    tInd.resize(tPtr[n]);
    std::vector<int> fill(tPtr.begin(), tPtr.end() - 1);
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < P->hRowLen[i]; ++k) {
        const int j = P->hCol[oslot(i, k)];
        if (j > i && j < n) tInd[fill[j]++] = i;
      }
  }
  std::vector<int> block(n, 0);
  int nBlocks = 1;
  if (s->iluOrdering == 0) {  // greedy multicolouring in the given order
    std::vector<int> mark;
    for (int i = 0; i < n; ++i) {
      mark.assign(16, 0);
      auto see = [&](int j) {
        if (block[j] >= (int)mark.size()) mark.resize(block[j] + 1, 0);
        mark[block[j]] = 1;
      };
      for (int k = 0; k < P->hRowLen[i]; ++k) {
        const int j = P->hCol[oslot(i, k)];
        if (j >= n || j >= i) continue;
        see(j);
      }
      if (!tPtr.empty())
        for (int k = tPtr[i]; k < tPtr[i + 1]; ++k) see(tInd[k]);
      int c = 0;
      while (c < (int)mark.size() && mark[c]) ++c;
      block[i] = c;
      nBlocks = std::max(nBlocks, c + 1);
    }
  } else {  // wavefront levels of the given (natural) ordering
    for (int i = 0; i < n; ++i) {
      int lv = 0;
      for (int k = 0; k < P->hRowLen[i]; ++k) {
        const int j = P->hCol[oslot(i, k)];
        if (j < i) lv = std::max(lv, block[j] + 1);
      }
      if (!tPtr.empty())
        for (int k = tPtr[i]; k < tPtr[i + 1]; ++k) lv = std::max(lv, block[tInd[k]] + 1);
      block[i] = lv;
      nBlocks = std::max(nBlocks, lv + 1);
    }
  }
  // symmetric permutation: rows sorted by set (stable)
  std::vector<int> new2old(n), old2new(n);
  std::iota(new2old.begin(), new2old.end(), 0);
  std::stable_sort(new2old.begin(), new2old.end(), [&](int a, int b) { return block[a] < block[b]; });
  for (int i = 0; i < n; ++i) old2new[new2old[i]] = i;
  D.blockPtr.assign(nBlocks + 1, 0);
  for (int i = 0; i < n; ++i) D.blockPtr[block[i] + 1]++;
  std::partial_sum(D.blockPtr.begin(), D.blockPtr.end(), D.blockPtr.begin());
  // permuted sliced-ELL pattern
  SellPattern &Q = D.pat;
  Q.nRows = n; Q.nCols = P->nCols; Q.nSlices = (n + 31) / 32; Q.nnz = P->nnz;
  Q.hRowLen.resize(n);
  for (int i = 0; i < n; ++i) Q.hRowLen[i] = P->hRowLen[new2old[i]];
  Q.hSliceOff.assign(Q.nSlices + 1, 0);
  for (int sl = 0; sl < Q.nSlices; ++sl) {
    int w = 1;
    for (int r = sl * 32; r < std::min(n, sl * 32 + 32); ++r) w = std::max(w, Q.hRowLen[r]);
    Q.hSliceOff[sl + 1] = Q.hSliceOff[sl] + w * 32;
  }
  Q.nSlots = Q.hSliceOff[Q.nSlices];
  Q.hCol.assign(Q.nSlots, 0);
  std::vector<int> slotMap(Q.nSlots, -1), diagK(n, 0), blockOf(n);
  std::vector<signed char> kind(Q.nSlots, 0);
  D.hasLower.assign(nBlocks, 0); D.hasUpper.assign(nBlocks, 0);
  for (int i = 0; i < n; ++i) blockOf[i] = block[new2old[i]];
  for (int sl = 0; sl < Q.nSlices; ++sl) {
    const int w = (Q.hSliceOff[sl + 1] - Q.hSliceOff[sl]) / 32;
    for (int lane = 0; lane < 32; ++lane) {
      const int r = sl * 32 + lane;
      for (int k = 0; k < w; ++k) {
        const size_t ns = (size_t)Q.hSliceOff[sl] + (size_t)k * 32 + lane;
        if (r >= n) { Q.hCol[ns] = n - 1; continue; }
        if (k >= Q.hRowLen[r]) { Q.hCol[ns] = r; continue; }
        const int o = new2old[r];
        const size_t os = oslot(o, k);
        const int oc = P->hCol[os];
        slotMap[ns] = (int)os;
        if (oc >= n) { Q.hCol[ns] = oc; continue; }   // ghost column: not part of the local factor
        const int nc = old2new[oc];
        Q.hCol[ns] = nc;
        if (nc == r) { kind[ns] = 2; diagK[r] = k; }
        else if (blockOf[nc] < blockOf[r]) { kind[ns] = 1; D.hasLower[blockOf[r]] = 1; }
        else { kind[ns] = 3; D.hasUpper[blockOf[r]] = 1; }
      }
    }
  }
  PHB_CHECK(Q.sliceOff.upload(Q.hSliceOff, st)); PHB_CHECK(Q.rowLen.upload(Q.hRowLen, st));
  PHB_CHECK(Q.col.upload(Q.hCol, st));
  PHB_CHECK(D.slotMap.upload(slotMap, st)); PHB_CHECK(D.diagK.upload(diagK, st));
  PHB_CHECK(D.blockOf.upload(blockOf, st)); PHB_CHECK(D.kind.upload(kind, st));
  PHB_CHECK(D.new2old.upload(new2old, st));
  PHB_CHECK(D.vals.alloc((size_t)Q.nSlots)); PHB_CHECK(D.lu.alloc((size_t)Q.nSlots));
  // halo send list in the permuted numbering
  std::vector<int> send;
  if (halo)
    for (int d : halo->hSendDev) send.push_back(old2new[d]);
  PHB_CHECK(D.sendDev.upload(send, st));
  PHB_CUDA(cudaStreamSynchronize(st));
  D.src = P; D.srcSlots = P->nSlots; D.ordering = s->iluOrdering; D.nBlocks = nBlocks;
  if (s->graphExec) { cudaGraphExecDestroy(s->graphExec); s->graphExec = nullptr; }
  return PHB_OK;
}

int ilu_factor(phb_solver *s, const double *vals) {
  IluData &D = s->ilu;
  phb_ctx *c = s->ctx;
  const long long nSlots = D.pat.nSlots;
  const int g = (int)std::max<long long>(1, std::min<long long>((nSlots + kThreads - 1) / kThreads, (long long)c->numSMs * 8));
  PHB_LAUNCH(c, k_permute_vals, g, kThreads, 0, nSlots, D.slotMap.p, vals, D.vals.p, D.lu.p);
  const IluView V = view_of(D);
  for (int b = 0; b < D.nBlocks; ++b) {
    if (!D.hasLower[b]) continue;
    const int r0 = D.blockPtr[b], r1 = D.blockPtr[b + 1];
    if (r1 > r0) PHB_LAUNCH(c, k_ilu_factor, (r1 - r0 + kThreads - 1) / kThreads, kThreads, 0, V, D.lu.p, r0, r1);
  }
  return PHB_OK;
}

// z = U^-1 L^-1 r  (vectors in the permuted numbering, leading dimension ld)
int ilu_apply(phb_solver *s, const double *r, double *z, const PeerFuse *pushHalo) {
  IluData &D = s->ilu;
  phb_ctx *c = s->ctx;
  const IluView V = view_of(D);
  const int ld = s->ld, nb = D.nBlocks, n0 = D.blockPtr[1];
  // the launch that completes z (set 0 of the backward sweep, or the only set) carries the halo push
  const int finalMode = 3;
  auto launch = [&](int mode, int r0, int r1) {
    const PeerFuse F = (pushHalo && mode == finalMode) ? *pushHalo : PeerFuse();
    const long long slices = ((r1 - 1) >> 5) - (r0 >> 5) + 1;
    const int grid = (int)std::max<long long>(1, std::min<long long>((slices * 32 + kThreads - 1) / kThreads,
                                                                    (long long)c->numSMs * 8));
#define SWEEP(NCV, M) PHB_LAUNCH(c, (k_ilu_sweep<NCV, M>), grid, kThreads, 0, V, D.lu.p, r, z, ld, r0, r1, n0, F, s->ticket.p)
    if (s->nComp == 1) {
      switch (mode) { case 0: SWEEP(1, 0); break; case 1: SWEEP(1, 1); break; case 2: SWEEP(1, 2); break; default: SWEEP(1, 3); }
    } else {
      switch (mode) { case 0: SWEEP(2, 0); break; case 1: SWEEP(2, 1); break; case 2: SWEEP(2, 2); break; default: SWEEP(2, 3); }
    }
#undef SWEEP
  };
  for (int b = 1; b < nb; ++b) {
    const int r0 = D.blockPtr[b], r1 = D.blockPtr[b + 1];
    if (r1 > r0) launch(b == nb - 1 ? 1 : 0, r0, r1);
  }
  for (int b = std::max(0, nb - 2); b >= 0; --b) {
    const int r0 = D.blockPtr[b], r1 = D.blockPtr[b + 1];
    if (r1 > r0) launch(b == 0 ? 3 : 2, r0, r1);
  }
  return PHB_OK;
}

int ilu_permute(phb_solver *s, const double *x, double *y, int dir) {
  phb_ctx *c = s->ctx;
  const int n = s->ilu.pat.nRows;
  const long long work = (long long)n * s->nComp;
  const int g = (int)std::max<long long>(1, std::min<long long>((work + kThreads - 1) / kThreads, (long long)c->numSMs * 8));
  PHB_LAUNCH(c, k_permute_vec, g, kThreads, 0, n, s->nComp, s->ld, s->ilu.new2old.p, x, y, dir);
  return PHB_OK;
}

int ilu_launches_per_apply(const phb_solver *s) {
  return 2 * std::max(1, s->ilu.nBlocks - 1) - (s->ilu.nBlocks == 1 ? 1 : 0);
}

}  // namespace phb
