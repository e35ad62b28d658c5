// solver.cuh -- internal definition of phb_solver.
#pragma once
#include "kernels.cuh"
#include "peerdev.cuh"
#include "structs.cuh"

// device-resident Krylov scalars.  One BiCGStab iteration needs TWO reductions:
//   R1: sigma = (rhat . v)                                  after v = A M^-1 p
//   R2: ts, tt, rs, rt, ss = (t.s), (t.t), (rhat.s), (rhat.t), (s.s)   after t = A M^-1 s
// from which omega = ts/tt, rho' = rs - omega rt and ||r'||^2 = ss - 2 omega ts + omega^2 tt
// follow without a third reduction ("finish" step, krylov_finish()).
struct KrylovSums {
  double rho[2];   // (rhat . r), parity-indexed by iteration
  double sigma;    // R1
  double ts, tt, rs, rt, ss;  // R2 (contiguous: one all-reduce of 5 doubles)
  double rr;       // ||r||^2 (recurrence), bb follows (contiguous pair for the initial all-reduce)
  double bb;       // ||b||^2
  double thresh;   // tol^2 * bb
  double iters;    // iterations completed
  double alpha, omega, beta;  // of the last completed iteration (consumed by the fused update)
  double pad;
};

// after R2 of iteration `cur` (single writer, no concurrent readers)
__host__ __device__ inline void krylov_finish(KrylovSums *S, int cur) {
  const double rho = S->rho[cur];
  const double alpha = rho / S->sigma;
  const double omega = S->tt != 0. ? S->ts / S->tt : 0.;
  const double rhoN = S->rs - omega * S->rt;
  const double rr = S->ss - 2. * omega * S->ts + omega * omega * S->tt;
  S->alpha = alpha;
  S->omega = omega;
  S->beta = (rhoN / rho) * (alpha / omega);
  S->rho[cur ^ 1] = rhoN;
  S->rr = rr > 0. ? rr : (rr == rr ? 0. : rr);  // clamp round-off negatives, keep NaN
  S->iters += 1.;
}

__device__ __forceinline__ bool krylov_done(const KrylovSums *S, int maxIters) {
  return !(S->rr > S->thresh) || S->iters >= (double)maxIters;
}

// tensor term of the operator as the SpMV kernels see it
struct TensorTerm {
  const double *t = nullptr;      // [4][n]
  const double *scale = nullptr;  // Jacobi fold: the iterate is D x, the term acts on dinv .* iterate
  const int *perm = nullptr;      // ILU: run row -> original row
  int n = 0;
};

inline phb::SellView view_of(const SellPattern *P) {
  phb::SellView v;
  v.sliceOff = P->sliceOff.p;
  v.col = P->col.p;
  v.nRows = P->nRows;
  v.nSlices = P->nSlices;
  v.nCols = P->nCols;
  return v;
}

// ILU(0): permuted pattern + factors (ilu.cu)
struct IluData {
  const SellPattern *src = nullptr;
  long long srcSlots = 0;
  int ordering = -1, nBlocks = 0;
  SellPattern pat;                 // rows grouped by independent set
  std::vector<int> blockPtr;       // nBlocks + 1 row offsets
  std::vector<char> hasLower, hasUpper;
  phb::DevBuf<int> slotMap, diagK, blockOf, new2old, sendDev;
  phb::DevBuf<signed char> kind;   // per slot: 0 pad/ghost, 1 lower, 2 diagonal, 3 upper
  phb::DevBuf<double> vals, lu;
};

// smoothed-aggregation AMG hierarchy (amg.cu)
struct AmgMat {
  SellPattern pat;
  phb::DevBuf<double> vals;        // bytes: float or double values per `amgPrecision`
};
namespace phb {
// `amgAggTheta` default: a coupling makes its neighbour a member of the row's aggregate when it reaches this fraction
// of the row's largest coupling (amg.cu: aggregate)
constexpr double kAmgAggTheta = 0.1;
// `amgCoarseSmootherWeight` default: Jacobi weight omegaC / lambda_max(D^-1 A) on the Galerkin levels (0: Gershgorin rule)
constexpr double kAmgCoarseWeight = 1.6;
}  // namespace phb

struct AmgLevel {
  int n = 0, ld = 0;               // rows, leading dimension of the level vectors
  AmgMat A, P, R;                  // level operator (levels >= 1), prolongator n x n_c, restriction n_c x n
  phb::DevBuf<double> w;           // smoother weight omega / a_ii
  double omegaEff = 1.8;           // w = omegaEff / rho_Gershgorin / a_ii: what the numeric re-setup rescales
  phb::DevBuf<double> x, x2, b, r;
  // distributed levels (nProcs > 1, `amgScope global`): the operator has ghost columns, refreshed by a
  // neighbour exchange before every residual / smoothing sweep
  bool dist = false;
  int nSend = 0;
  std::vector<int> sendOff, sendCnt, recvOff, recvCnt;   // per rank; recvOff = index into the level vector
  phb::DevBuf<int> sendIdx;
  phb::DevBuf<double> sendBuf;
};
constexpr int kCoarseSweeps = 8;   // Jacobi sweeps on a coarsest level too large for a dense inverse

// Ghost refreshes of the distributed multigrid levels over NVLink peer memory (amg.cu): the level vectors that
// peers write into (x, x2 of every distributed level, the gathered right-hand side of the first replicated
// level) live in ONE CUDA-IPC block per rank; one one-CTA kernel per refresh packs, stores into the peers'
// blocks, raises an epoch flag there and waits for its own.
constexpr int kAmgPeerChannels = 32;          // 2 per distributed level + 1 for the tail gather
constexpr size_t kAmgPeerHeaderBytes = 4096;  // flags[channel][source] | local epochs[channel]
struct AmgPeerLevel {                          // how to reach level l of every peer (by value into the kernels)
  unsigned long long vecOff[2][kMaxPeers];    // byte offset of x / x2 inside peer q's block
  int ld[kMaxPeers];                          // leading dimension of peer q's level vectors
  int recvOff[kMaxPeers];                     // where MY values land in peer q's vector
  int sendOff[kMaxPeers], sendCnt[kMaxPeers], recvCnt[kMaxPeers];
};
struct AmgPeer {
  char *block = nullptr;
  size_t bytes = 0;
  char *peer[kMaxPeers] = {nullptr};
  bool opened[kMaxPeers] = {false};
  std::vector<AmgPeerLevel> lev;              // distributed levels
  AmgPeerLevel tail;                          // gathered vector of the first replicated level (vecOff[0], ld)
  ~AmgPeer() {
    for (int q = 0; q < kMaxPeers; ++q)
      if (opened[q] && peer[q]) cudaIpcCloseMemHandle(peer[q]);
    if (block) cudaFree(block);
  }
};

// One phase of the fused small-level kernel (amg.cu: k_amg_tail): every level from `fuseFrom` down to the dense
// coarsest solve and back up runs in ONE launch, the phases separated by grid barriers instead of kernel boundaries.
constexpr int kMaxTailOps = 48;   // op list of the fused small-level kernel, staged in shared memory (4 per level + 1)
struct AmgTailOp {
  int kind;                 // 0 r = b - A (w.*b) | 1 y = R r | 2 x = Ainv b (dense) | 3 x = w.*b + P xc | 4 y = x + w.*(b - A x)
  int n, ld, ldIn;          // rows; leading dimension of the row-wise vectors; of the gathered vector
  int lanes;                // lanes per row (power of two <= 32)
  const int *sliceOff, *col;
  const void *vals;         // matrix values (cycle precision), or the fp64 dense inverse (kind 2)
  const void *w, *b, *x;    // smoother weights; right-hand side; gathered vector (x, r or xc)
  void *y;                  // result
};

// Symbolic part of a setup kept on the device for the numeric re-setup (amg_refresh.cuh): per level the CSR pattern
// of the operator, the strength flags and aggregates, the patterns of P, R = P^T and A P, and for every sliced-ELL
// matrix of the cycle the CSR entry each slot holds.
struct AmgRefreshLevel {
  int n = 0, nc = 0;
  long long nnzA = 0, nnzP = 0, nnzAP = 0;
  phb::DevBuf<int> aRp, aCi, aSrc, agg, pRp, pCi, rRp, rCi, rSrc, apRp, apCi, aSell, pSell, rSell;
  phb::DevBuf<unsigned char> strong;
  phb::DevBuf<double> aV, diag, df, pV, rV, apV;
};
struct AmgRefresh {
  std::vector<std::unique_ptr<AmgRefreshLevel>> lev;
  phb::DevBuf<double> scal, dense[2];   // 4 scalars per level: rho, rhoP, mean diagonal, zero-diagonal flag
  phb::DevBuf<int> flag;
  bool singular = false;
  double bytes = 0.;                    // device memory the symbolic data takes
};

struct AmgData {
  std::vector<std::unique_ptr<AmgLevel>> lev;
  std::unique_ptr<AmgRefresh> refresh;      // single-rank hierarchies (`amgRefresh` != off)
  int refreshMode = 1;                      // `amgRefresh`: 0 off | 1 auto (when the iteration count drifts) | 2 always
  int refreshes = 0, itersAfterRefresh = -1;
  double refreshMs = 0.;
  phb::DevBuf<double> coarseInv, refVals, chk;
  phb::DevBuf<float> refValsF;
  bool single = true, builtSingle = true;   // cycle precision (`amgPrecision single|double`)
  bool global = true;                       // nProcs > 1: hierarchy spans the ranks (else rank-local blocks)
  int nDist = 0;                            // leading distributed levels; level nDist is gathered on every rank
  long long tailRows = 200000;
  std::vector<int> tailOff, tailCnt, tailSendOff, tailSendCnt;
  std::unique_ptr<AmgPeer> peer;            // set when the context has peer memory enabled (else NCCL send/recv)
  // fused tail: levels fuseFrom .. L-1 in one launch (fuseFrom < 0: off)
  long long fuseRows = 200000;              // `amgFuseRows`: levels with at most this many rows are fused (0: never)
  int fuseFrom = -1, nTailOps = 0;
  phb::DevBuf<AmgTailOp> tailOps;
  phb::DevBuf<unsigned> tailBar;
  const SellPattern *src = nullptr;
  unsigned long long builtTag = 0;          // coefficient tag of the matrix the hierarchy's values belong to (0: untagged)
  bool built = false, denseCoarse = false, stale = false, rebuildAlways = false;
  int nComp = 1, nCoarse = 0, nu = 1, coarsest = 1000, setups = 0, itersAfterSetup = -1;
  // smoother weight omegaS / rho(D^-1 A): 1.8 instead of the textbook 4/3 -- inside a Krylov method the stronger damping of
  // the mid-range modes wins (scipy transcription, 1M cells: 10-12 instead of 13-15 iterations; 11-13 % fewer on the
  // variable-density and 7-point operators); |1 - 1.8 lambda / rho| <= 0.8 for every mode, so the sweep stays a contraction
  double theta = 0., thetaAgg = phb::kAmgAggTheta, omegaS = 1.8, omegaC = phb::kAmgCoarseWeight, setupMs = 0., opComplexity = 1.;
};

struct phb_solver {
  phb_ctx *ctx = nullptr;
  // configuration (keys of LinearAlgebra.<eqn>, M/TrilinosBelosSparseMatrixSolver.cpp:44-86)
  int maxIters = 500;
  double tol = 1e-8;
  int precond = PHB_PC_ILU0;       // the reference's Belos default: Schwarz(overlap 0) + RILUK(0)
  std::string method = "BICGSTAB";
  int itersPerGraph = 8;
  bool projectConstant = false;    // singular all-Neumann systems: remove the constant from b (hard part 3)
  bool useGraph = true;
  // ---- matrix as handed over by set_csr (host CSR cache for pattern reuse)
  int nRows = 0, nColsGlobal = 0;
  std::vector<int> cRowPtr, cColInd;
  SellPattern own;
  phb::DevBuf<int> csr2slot;
  phb::DevBuf<double> csrVals, ownVals;
  bool haveMatrix = false;
  // ---- active system (either `own` or borrowed from an equation)
  const SellPattern *pat = nullptr;
  const double *dVals = nullptr;
  int nComp = 1;
  int ld = 0;  // vector leading dimension (= pat->nCols)
  const phb_mesh *halo = nullptr;  // halo lists (nProcs > 1)
  // 2 x 2 tensor on the (row, row) block of a two-component system ([xx | xy | yx | yy][nRows], SYMMETRY patches): part
  // of the operator in every SpMV, left out of the preconditioners (they see the shared coefficients only)
  const double *tens = nullptr;
  phb::DevBuf<double> scaled, dinv;
  IluData ilu;
  int iluOrdering = 0;             // 0 multicolour, 1 wavefront levels of the given ordering
  AmgData amg;
  phb::DevBuf<double> ph, sh;      // M^-1 p, M^-1 s (ILU only; alias p, s otherwise)
  // run-time view of the system being iterated on (permuted when ILU is active)
  const SellPattern *runPat = nullptr;
  const int *runSendDev = nullptr;
  // peer exchanges pushed/awaited inside the compute kernels instead of by separate one-CTA kernels.
  // Measured on 2 B200s (2M rows per GPU): 0.235 vs 0.223 ms per iteration -- the waits serialise the
  // same way and every CTA pays for them, so the separate kernels stay the default (`peerFusion 1` opts in).
  bool peerFused = false;
  unsigned long long valsTag = 0;  // coefficient tag of the solve in progress (fv.cu: eqn_solve_tagged), 0 = none
  int peerRegion = -1;             // slot of this solver in the peer arena (-1 unassigned, -2 not usable)
  double *runPh = nullptr, *runSh = nullptr;
  phb::DevBuf<double> b, x, r, rhat, p, v, s, t;
  phb::DevBuf<double> partials, proj;
  phb::DevBuf<unsigned> ticket;
  phb::DevBuf<KrylovSums> sums;
  bool haveRhs = false, haveGuess = false;
  // CUDA graph of `itersPerGraph` iterations, keyed on the active buffers
  cudaGraphExec_t graphExec = nullptr;
  const void *graphKey[4] = {nullptr, nullptr, nullptr, nullptr};
  // results
  int lastIters = 0;
  double lastRelres = 0.;
  std::vector<double> hostX;
};

namespace phb {
// core entry used by phb_solver_solve and phb_eqn_solve: solves A x = b with the
// device vectors already in s->b / s->x (x = initial guess), matrix (pat,dVals).
int solver_run(phb_solver *s, int *iters, double *relres);
int solver_bind(phb_solver *s, const SellPattern *pat, const double *dVals, int nComp,
                const phb_mesh *halo);
int ilu_prepare(phb_solver *s, const SellPattern *P, const phb_mesh *halo);
int ilu_factor(phb_solver *s, const double *vals);
int ilu_apply(phb_solver *s, const double *r, double *z, const PeerFuse *pushHalo = nullptr);
int ilu_permute(phb_solver *s, const double *x, double *y, int dir);
int ilu_launches_per_apply(const phb_solver *s);
int amg_prepare(phb_solver *s);
int amg_apply(phb_solver *s, const double *in, double *out, bool inLoop);
int amg_launches_per_apply(const phb_solver *s);
void amg_record_iters(phb_solver *s, int iters);
bool amg_wants_refresh(const phb_solver *s, int itersSoFar);
int amg_refresh_midsolve(phb_solver *s);
int amg_check(phb_solver *s);
double amg_cycle_bytes(const phb_solver *s);
int amg_time(phb_solver *s, int reps, double out[8]);
}  // namespace phb
