// amg.cu -- smoothed-aggregation algebraic multigrid preconditioner (`preconditioner amg`).
//
// Role in the reference: the MueLu backend (Math/TrilinosMueluSparseMatrixSolver.cpp:27-32 builds a
// MueLu hierarchy from the Tpetra matrix and hands it to Belos as the right preconditioner, `lib muelu`);
// here it is a preconditioner choice of the one B200 backend.  One V(nu,nu) cycle with damped-Jacobi
// smoothing is a fixed linear operator, so it drops into the right-preconditioned BiCGStab unchanged.
//
// Split of the work:
//   * setup (host, once per matrix): strength graph -> greedy aggregation -> tentative prolongator T
//     (piecewise constant) -> P = (I - w D^-1 A) T -> R = P^T -> A_c = R A P, repeated until the level
//     has <= `amgCoarsest` rows; the coarsest operator is inverted densely (regularised with the
//     constant vector when the system is singular, i.e. all-Neumann pressure).  The hierarchy is kept
//     while the matrix stays a scalar multiple of the one it was built from (FractionalStep's pEqn_ is
//     `laplacian(dt, p)`: constant up to dt) and rebuilt only when the iteration count degrades.
//   * cycle (device, inside the CUDA graph of the Krylov loop): every level operator, P and R are
//     sliced-ELL matrices driven by the same warp<->slice streaming code as the Krylov SpMV.
// Rank-local on several GPUs (ghost columns dropped) = the reference's additive Schwarz with overlap 0.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <numeric>
#include <thread>

#include "comm.cuh"
#include "kernels.cuh"
#include "solver.cuh"

using namespace phb;

// ===================================================================== host setup
namespace {

struct HCsr {
  int n = 0, m = 0;  // rows, cols
  std::vector<int> rp, ci;
  std::vector<double> v;
  long long nnz() const { return (long long)ci.size(); }
};

HCsr transpose(const HCsr &A) {
  HCsr T;
  T.n = A.m; T.m = A.n;
  T.rp.assign(T.n + 1, 0);
  for (int c : A.ci) T.rp[c + 1]++;
  for (int i = 0; i < T.n; ++i) T.rp[i + 1] += T.rp[i];
  T.ci.resize(A.ci.size());
  T.v.resize(A.ci.size());
  std::vector<int> fill(T.rp.begin(), T.rp.end() - 1);
  for (int i = 0; i < A.n; ++i)
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k) {
      const int dst = fill[A.ci[k]]++;
      T.ci[dst] = i;
      T.v[dst] = A.v[k];
    }
  return T;
}

// ---- host threads for the row-parallel parts of the setup (results do not depend on the thread count)
int g_hostRanks = 1;   // ranks sharing this host (set by the distributed setup)
int host_threads() {
  if (const char *e = getenv("PHB_HOST_THREADS")) return std::max(1, atoi(e));
  const int hw = (int)std::thread::hardware_concurrency();
  return std::max(1, std::min(16, hw / std::max(1, g_hostRanks)));
}
// f(begin, end, chunk) over [0, n) in contiguous chunks, one thread each; returns the number of chunks
template <typename F> int parallel_chunks(int n, F f) {
  const int T = std::max(1, std::min(host_threads(), n / 8192));
  const int chunk = (n + T - 1) / T;
  if (T == 1) { f(0, n, 0); return 1; }
  std::vector<std::thread> th;
  for (int t = 0; t < T; ++t) th.emplace_back(f, std::min(n, t * chunk), std::min(n, (t + 1) * chunk), t);
  for (auto &x : th) x.join();
  return T;
}
int chunk_count(int n) { return std::max(1, std::min(host_threads(), n / 8192)); }

// rows computed chunk by chunk into private arrays, then stitched: C.rp holds row lengths on entry
void stitch(HCsr &C, std::vector<std::vector<int>> &ci, std::vector<std::vector<double>> &v,
            const std::vector<int> &chunkBegin) {
  for (int i = 0; i < C.n; ++i) C.rp[i + 1] += C.rp[i];
  C.ci.resize(C.rp[C.n]);
  C.v.resize(C.rp[C.n]);
  std::vector<std::thread> th;
  for (size_t t = 0; t < ci.size(); ++t)
    th.emplace_back([&, t] {
      std::copy(ci[t].begin(), ci[t].end(), C.ci.begin() + C.rp[chunkBegin[t]]);
      std::copy(v[t].begin(), v[t].end(), C.v.begin() + C.rp[chunkBegin[t]]);
    });
  for (auto &x : th) x.join();
}

// C = A B (Gustavson, marker array); rows of C sorted by column; row-parallel
HCsr spgemm(const HCsr &A, const HCsr &B) {
  HCsr C;
  C.n = A.n; C.m = B.m;
  C.rp.assign(C.n + 1, 0);
  const int T = chunk_count(A.n);
  std::vector<std::vector<int>> ci(T);
  std::vector<std::vector<double>> v(T);
  std::vector<int> chunkBegin(T, 0);
  parallel_chunks(A.n, [&](int begin, int end, int t) {
    chunkBegin[t] = begin;
    std::vector<int> marker(B.m, -1);
    std::vector<int> &cc = ci[t];
    std::vector<double> &cv = v[t];
    {  // one allocation per chunk: rows of A B hold about (entries per row of A) x (entries per row of B) / 2 entries
      const double perRowA = A.n ? (double)A.nnz() / A.n : 0., perRowB = B.n ? (double)B.nnz() / B.n : 0.;
      const size_t guess = (size_t)((end - begin) * std::max(perRowA, std::min(perRowA * perRowB, perRowA * perRowB * 0.5 + 2.)));
      cc.reserve(guess);
      cv.reserve(guess);
    }
    std::vector<std::pair<int, double>> row;
    for (int i = begin; i < end; ++i) {
      const int start = (int)cc.size();
      for (int ka = A.rp[i]; ka < A.rp[i + 1]; ++ka) {
        const int j = A.ci[ka];
        const double a = A.v[ka];
        for (int kb = B.rp[j]; kb < B.rp[j + 1]; ++kb) {
          const int c = B.ci[kb];
          if (marker[c] < start) {
            marker[c] = (int)cc.size();
            cc.push_back(c);
            cv.push_back(a * B.v[kb]);
          } else {
            cv[marker[c]] += a * B.v[kb];
          }
        }
      }
      const int stop = (int)cc.size();
      row.clear();
      for (int k = start; k < stop; ++k) row.push_back({cc[k], cv[k]});
      std::sort(row.begin(), row.end(), [](const std::pair<int, double> &x, const std::pair<int, double> &y) {
        return x.first < y.first;
      });
      for (int k = start; k < stop; ++k) { cc[k] = row[k - start].first; cv[k] = row[k - start].second; }
      for (int k = start; k < stop; ++k) marker[cc[k]] = -1;  // positions moved: forget them
      C.rp[i + 1] = stop - start;
    }
  });
  stitch(C, ci, v, chunkBegin);
  return C;
}

std::vector<double> diagonal(const HCsr &A) {
  std::vector<double> d(A.n, 0.);
  for (int i = 0; i < A.n; ++i)
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
      if (A.ci[k] == i) d[i] += A.v[k];
  return d;
}

// `amgTheta` (strength of connection, also filters the operator the prolongator is smoothed with) and `amgAggTheta`
// (membership graph of the aggregates only).
// Smoother rule: damped Jacobi with weight omegaS / rho_Gershgorin on the levels before `firstCoarse`, and
// omegaC / lambda on the others, lambda a power-iteration estimate of the largest eigenvalue of D^-1 A (omegaC = 0:
// the Gershgorin rule everywhere).  On Galerkin levels the Gershgorin bound (2 for every zero-row-sum M-matrix)
// overestimates lambda (1.35-1.5) by a third and the sweep is correspondingly weak.
struct Strength {
  double theta, agg;
  double omegaS = 1.8, omegaC = phb::kAmgCoarseWeight;
  int firstCoarse = 1;
  Strength(double t = 0., double a = phb::kAmgAggTheta) : theta(t), agg(a) {}
};

Strength strength_of(const AmgData &D) {
  Strength st(D.theta, D.thetaAgg);
  st.omegaS = D.omegaS;
  st.omegaC = D.omegaC;
  return st;
}

// Greedy aggregation on the strength graph |a_ij|^2 >= theta^2 |a_ii a_jj| (three passes: roots whose
// strong neighbourhood is free, leftovers join a neighbouring aggregate, the rest seed new aggregates).
int aggregate(const HCsr &A, const std::vector<double> &d, Strength st, std::vector<int> &agg,
              std::vector<char> &strong) {
  const int n = A.n;
  const double theta = st.theta;
  strong.assign(A.ci.size(), 0);
  for (int i = 0; i < n; ++i)
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k) {
      const int j = A.ci[k];
      const double a = A.v[k];
      // columns >= n are ghosts (rows of another rank): aggregates and prolongator smoothing stay rank-local
      strong[k] = (j != i && j < n && a != 0. && a * a >= theta * theta * std::fabs(d[i] * d[j])) ? 1 : 0;
    }
  // Membership graph of the aggregates: strong couplings that reach thetaAgg x the largest strong coupling of the row.
  // The Galerkin operators of a smoothed prolongator carry many small long-range entries; with every one of them a
  // member-maker, a root of a 17-entry row founds a 17-cell aggregate where 9 is typical, and a smooth error mode
  // localised there survives the coarse correction (uniform Poisson, 4000 x 2000 cells: 21 iterations, 4001 x 2000: 10;
  // with the filter 11 and 10).  The prolongator is still smoothed with the unfiltered strength graph, so this is not
  // `amgTheta`: that one makes the variable-density and the momentum hierarchies worse (31 vs 21, stalled coarsening).
  std::vector<char> sa(strong);
  {
    const double thetaAgg = st.agg;
    if (thetaAgg > 0.)
      for (int i = 0; i < n; ++i) {
        double mx = 0.;
        for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
          if (strong[k]) mx = std::max(mx, std::fabs(A.v[k]));
        for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
          if (strong[k] && std::fabs(A.v[k]) < thetaAgg * mx) sa[k] = 0;
      }
  }
  agg.assign(n, -1);
  int nc = 0;
  // roots in natural order: on mesh-derived matrices this tiles the graph regularly; a scrambled order was tried and
  // needs three times the iterations (irregular aggregates, 31-35 instead of 10 at 4M-8M cells)
  for (int i = 0; i < n; ++i) {
    if (agg[i] >= 0) continue;
    bool free_ = true;
    for (int k = A.rp[i]; k < A.rp[i + 1] && free_; ++k)
      if (sa[k] && agg[A.ci[k]] >= 0) free_ = false;
    if (!free_) continue;
    agg[i] = nc;
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
      if (sa[k]) agg[A.ci[k]] = nc;
    ++nc;
  }
  // leftovers join the neighbouring root aggregate they are coupled to most strongly (sum of |a_ij| over its
  // members), the smaller one on a tie: first-found joins make the aggregate shapes -- and with them the iteration
  // count -- depend on the row length of the mesh (4000 x 2000 cells: 23 iterations, 4001 x 2000: 10)
  std::vector<int> pass2(agg), size(nc, 0);
  for (int i = 0; i < n; ++i)
    if (agg[i] >= 0) size[agg[i]]++;
  for (int i = 0; i < n; ++i) {
    if (agg[i] >= 0) continue;
    int best = -1;
    double bestW = 0.;
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k) {
      if (!sa[k] || agg[A.ci[k]] < 0) continue;
      const int a = agg[A.ci[k]];
      double w = 0.;
      for (int q = A.rp[i]; q < A.rp[i + 1]; ++q)
        if (sa[q] && agg[A.ci[q]] == a) w += std::fabs(A.v[q]);
      if (best < 0 || w > bestW * (1. + 1e-12) || (w >= bestW * (1. - 1e-12) && size[a] < size[best])) { best = a; bestW = w; }
    }
    if (best >= 0) { pass2[i] = best; size[best]++; }
  }
  agg.swap(pass2);
  for (int i = 0; i < n; ++i) {
    if (agg[i] >= 0) continue;
    agg[i] = nc;
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
      if (sa[k] && agg[A.ci[k]] < 0) agg[A.ci[k]] = nc;
    ++nc;
  }
  return nc;
}

struct HostLevel {
  HCsr A, P, R;
  std::vector<double> diag;  // of the (filtered) operator the smoother uses
  double rho = 2.;           // Gershgorin bound of rho(D^-1 A)
  double wScale = 0.9;       // smoother weights = wScale / a_ii (omegaS / rho, or omegaC / lambda on Galerkin levels)
  // kept for the numeric re-setup on the device (build_hierarchy(..., keepSymbolic)): aggregates, strength flags, pattern of A P
  std::vector<int> agg, apRp, apCi;
  std::vector<char> strong;
};

struct HostHierarchy {
  std::vector<HostLevel> lev;
  std::vector<double> coarseInv;  // dense, row-major, lev.back().A.n squared (empty: Jacobi sweeps instead)
  bool singular = false;
  double setupMs = 0.;
  double opComplexity = 1.;
};

constexpr int kDenseMax = 1024;

}  // namespace
struct phb_amg_host {
  HostHierarchy H;
};
namespace {

bool dense_inverse(std::vector<double> &M, int n) {
  // LU with partial pivoting (row-major, in place), then the columns of the inverse by forward / backward substitution,
  // column blocks in parallel
  std::vector<int> perm(n);
  std::iota(perm.begin(), perm.end(), 0);
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = std::fabs(M[(size_t)k * n + k]);
    for (int r = k + 1; r < n; ++r)
      if (std::fabs(M[(size_t)r * n + k]) > best) { best = std::fabs(M[(size_t)r * n + k]); piv = r; }
    if (!(best > 0.)) return false;
    if (piv != k) {
      std::swap_ranges(M.begin() + (size_t)k * n, M.begin() + (size_t)(k + 1) * n, M.begin() + (size_t)piv * n);
      std::swap(perm[k], perm[piv]);
    }
    const double inv = 1. / M[(size_t)k * n + k];
    const double *rk = &M[(size_t)k * n];
    for (int r = k + 1; r < n; ++r) {
      double *rr = &M[(size_t)r * n];
      const double l = rr[k] * inv;
      if (l == 0.) continue;
      rr[k] = l;
      for (int c = k + 1; c < n; ++c) rr[c] -= l * rk[c];
    }
  }
  std::vector<double> I((size_t)n * n);
  const int T = std::max(1, std::min(host_threads(), n / 64));
  auto solve_cols = [&](int j0, int j1) {
    std::vector<double> y(n);
    for (int j = j0; j < j1; ++j) {
      // column j of the inverse solves A x = e_j, i.e. L U x = P e_j
      for (int i = 0; i < n; ++i) {
        double acc = perm[i] == j ? 1. : 0.;
        const double *ri = &M[(size_t)i * n];
        for (int k = 0; k < i; ++k) acc -= ri[k] * y[k];
        y[i] = acc;
      }
      for (int i = n - 1; i >= 0; --i) {
        double acc = y[i];
        const double *ri = &M[(size_t)i * n];
        for (int k = i + 1; k < n; ++k) acc -= ri[k] * y[k];
        y[i] = acc / ri[i];
      }
      for (int i = 0; i < n; ++i) I[(size_t)i * n + j] = y[i];
    }
  };
  if (T == 1) {
    solve_cols(0, n);
  } else {
    std::vector<std::thread> th;
    const int chunk = (n + T - 1) / T;
    for (int t = 0; t < T; ++t) th.emplace_back(solve_cols, std::min(n, t * chunk), std::min(n, (t + 1) * chunk));
    for (auto &x : th) x.join();
  }
  M.swap(I);
  return true;
}

double gershgorin(const HCsr &A, const std::vector<double> &d) {
  double rho = 0.;
  for (int i = 0; i < A.n; ++i) {
    double s = 0.;
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k) s += std::fabs(A.v[k]);
    rho = std::max(rho, s / std::fabs(d[i]));
  }
  return rho;
}

// largest eigenvalue of D^-1 A by power iteration from a fixed pseudo-random vector (row-parallel, independent of the
// thread count up to the rounding of the norms)
double power_lambda(const HCsr &A, const std::vector<double> &d, int iters = 20) {
  const int n = A.n;
  std::vector<double> x(n), y(n);
  for (int i = 0; i < n; ++i) {
    unsigned h = (unsigned)i * 2654435761u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    x[i] = ((h & 1u) ? 1. : -1.) * (0.5 + (double)((h >> 8) & 0xffffu) / 65536.);
  }
  double lam = 0.;
  for (int it = 0; it < iters; ++it) {
    const int T = chunk_count(n);
    std::vector<double> nx(T, 0.), ny(T, 0.);
    parallel_chunks(n, [&](int begin, int end, int t) {
      double sx = 0., sy = 0.;
      for (int i = begin; i < end; ++i) {
        double acc = 0.;
        for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
          if (A.ci[k] < n) acc += A.v[k] * x[A.ci[k]];
        y[i] = acc / d[i];
        sx += x[i] * x[i]; sy += y[i] * y[i];
      }
      nx[t] = sx; ny[t] = sy;
    });
    double sx = 0., sy = 0.;
    for (int t = 0; t < T; ++t) { sx += nx[t]; sy += ny[t]; }
    if (!(sx > 0.) || !(sy > 0.)) return 0.;
    lam = std::sqrt(sy / sx);
    const double inv = 1. / std::sqrt(sy);
    for (int i = 0; i < n; ++i) x[i] = y[i] * inv;
  }
  return lam;
}

// smoother weight scale of a level (HostLevel::wScale); L.diag and L.rho are set
void smoother_rule(const HCsr &A, const Strength &st, int level, HostLevel &L) {
  L.wScale = st.omegaS / L.rho;
  if (st.omegaC > 0. && level >= st.firstCoarse) {
    const double lam = power_lambda(A, L.diag);
    // the iteration approaches lambda from below: 5 % on top, never beyond the bound, never below half of it
    if (lam > 0.) L.wScale = st.omegaC / std::min(L.rho, std::max(0.5 * L.rho, 1.05 * lam));
  }
}

// One coarsening step on the rows of A (A.n owned rows; columns >= A.n are ghosts): fills the smoother data of
// L and the prolongator P = (I - (omegaP / rho) Df^-1 Af) T with T(i, agg[i]) = 1 and Af the operator with weak
// and ghost couplings lumped onto the diagonal (so that P reproduces the constant whenever A does).
// Returns 1 when the level cannot be coarsened any further, 0 on success, < 0 on error.
int make_prolongator(const HCsr &A, Strength theta, double omegaP, int level, HostLevel &L, HCsr &P, int &nc,
                     bool allowStall = false) {
  const int n = A.n;
  std::vector<double> d = diagonal(A);
  for (int i = 0; i < n; ++i)
    if (d[i] == 0.) { set_error("amg: zero diagonal in row %d of level %d", i, level); return PHB_ERR_BREAKDOWN; }
  L.diag = d;
  L.rho = gershgorin(A, d);
  std::vector<int> agg;
  std::vector<char> strong;
  nc = aggregate(A, d, theta, agg, strong);
  if (!allowStall && (nc >= n || (long long)nc * 10 > (long long)n * 9)) return 1;
  std::vector<double> df(d);
  for (int i = 0; i < n; ++i)
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
      if (!strong[k] && A.ci[k] != i) df[i] += A.v[k];
  for (int i = 0; i < n; ++i)
    if (df[i] == 0. || (df[i] > 0.) != (d[i] > 0.)) df[i] = d[i];
  double rho = 0.;
  for (int i = 0; i < n; ++i) {
    double s = std::fabs(df[i]);
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
      if (strong[k]) s += std::fabs(A.v[k]);
    rho = std::max(rho, s / std::fabs(df[i]));
  }
  P = HCsr();
  P.n = n; P.m = nc;
  P.rp.assign(n + 1, 0);
  const double w = omegaP / rho;
  const int T = chunk_count(n);
  std::vector<std::vector<int>> pci(T);
  std::vector<std::vector<double>> pv(T);
  std::vector<int> chunkBegin(T, 0);
  parallel_chunks(n, [&](int begin, int end, int t) {
    chunkBegin[t] = begin;
    std::vector<std::pair<int, double>> row;
    pci[t].reserve((size_t)(end - begin) * 4);
    pv[t].reserve((size_t)(end - begin) * 4);
    for (int i = begin; i < end; ++i) {
      row.clear();
      row.push_back({agg[i], 1. - w});  // diagonal term of Af: df/df = 1
      for (int k = A.rp[i]; k < A.rp[i + 1]; ++k) {
        if (!strong[k]) continue;
        const int c = agg[A.ci[k]];
        const double val = -w * A.v[k] / df[i];
        bool hit = false;
        for (auto &e : row)
          if (e.first == c) { e.second += val; hit = true; break; }
        if (!hit) row.push_back({c, val});
      }
      std::sort(row.begin(), row.end(), [](const std::pair<int, double> &x, const std::pair<int, double> &y) {
        return x.first < y.first;
      });
      for (auto &e : row) { pci[t].push_back(e.first); pv[t].push_back(e.second); }
      P.rp[i + 1] = (int)row.size();
    }
  });
  stitch(P, pci, pv, chunkBegin);
  L.agg.swap(agg);
  L.strong.swap(strong);
  return 0;
}

bool rows_sum_to_zero(const HCsr &A) {
  double maxRow = 0., maxDiag = 0.;
  for (int i = 0; i < A.n; ++i) {
    double sum = 0.;
    for (int k = A.rp[i]; k < A.rp[i + 1]; ++k) {
      sum += A.v[k];
      if (A.ci[k] == i) maxDiag = std::max(maxDiag, std::fabs(A.v[k]));
    }
    maxRow = std::max(maxRow, std::fabs(sum));
  }
  return maxRow <= 1e-10 * maxDiag;
}

int build_hierarchy(HCsr A0, Strength theta, int coarsest, double omegaP, HostHierarchy &H, bool keepSymbolic = false) {
  const auto t0 = std::chrono::steady_clock::now();
  H.lev.clear();
  if (A0.nnz() > 1200000000LL) {  // row pointers of the products are 32-bit
    set_error("amg: %lld entries are too many for the host setup (32-bit row pointers)", A0.nnz());
    return PHB_ERR_UNSUPPORTED;
  }
  // singular with the constant in the null space? (all-Neumann pressure: every row sums to zero)
  H.singular = rows_sum_to_zero(A0);
  const long long nnz0 = std::max<long long>(1, A0.nnz());
  long long nnzAll = 0;
  HCsr A = std::move(A0);
  for (int level = 0;; ++level) {
    HostLevel L;
    const int n = A.n;
    const auto tL = std::chrono::steady_clock::now();
    nnzAll += A.nnz();
    const bool last = n <= coarsest || level >= 15;
    HCsr P;
    int nc = 0, rc = 1;
    if (last) {
      L.diag = diagonal(A);
      for (int i = 0; i < n; ++i)
        if (L.diag[i] == 0.) { set_error("amg: zero diagonal in row %d of level %d", i, level); return PHB_ERR_BREAKDOWN; }
      L.rho = gershgorin(A, L.diag);
    } else {
      rc = make_prolongator(A, theta, omegaP, level, L, P, nc);
      if (rc < 0) return rc;
    }
    smoother_rule(A, theta, level, L);
    if (rc == 1) {  // coarsest level (or coarsening stalled)
      L.A = std::move(A);
      H.lev.push_back(std::move(L));
      break;
    }
    const auto t1 = std::chrono::steady_clock::now();
    L.R = transpose(P);
    const auto t2 = std::chrono::steady_clock::now();
    HCsr AP = spgemm(A, P);
    const auto t3 = std::chrono::steady_clock::now();
    HCsr Ac = spgemm(L.R, AP);
    const auto t4 = std::chrono::steady_clock::now();
    if (getenv("PHB_AMG_TIMING")) {
      auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
      };
      fprintf(stderr, "amg setup level %d (%d rows): prolongator %.0f ms, transpose %.0f, A*P %.0f, R*(AP) %.0f\n", level, n,
              ms(tL, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4));
    }
    if (keepSymbolic) { L.apRp.swap(AP.rp); L.apCi.swap(AP.ci); }
    else { std::vector<int>().swap(L.agg); std::vector<char>().swap(L.strong); }
    L.P = std::move(P);
    L.A = std::move(A);
    H.lev.push_back(std::move(L));
    A = std::move(Ac);
  }
  H.opComplexity = (double)nnzAll / (double)nnz0;
  // coarsest solve
  const HCsr &C = H.lev.back().A;
  H.coarseInv.clear();
  if (C.n <= kDenseMax) {
    const int n = C.n;
    std::vector<double> M((size_t)n * n, 0.);
    double meanDiag = 0.;
    for (int i = 0; i < n; ++i)
      for (int k = C.rp[i]; k < C.rp[i + 1]; ++k) {
        M[(size_t)i * n + C.ci[k]] += C.v[k];
        if (C.ci[k] == i) meanDiag += C.v[k] / n;
      }
    if (H.singular)
      for (size_t k = 0; k < M.size(); ++k) M[k] += meanDiag / n;  // + (mean diag / n) 1 1^T
    if (dense_inverse(M, n)) H.coarseInv.swap(M);
  }
  H.setupMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return PHB_OK;
}

// ===================================================================== distributed setup (nProcs > 1)
// Aggregates and prolongator smoothing are rank-local (P is block diagonal over the ranks), the Galerkin
// operators keep the couplings between ranks: A_c(rows of r) = P_r^T [A_rr P_r | A_rq P_q(ghost rows)], which
// needs the P rows of the ghost cells (one neighbour exchange per level) and gives every coarse level its
// own ghost columns and halo lists.  Once the global row count is small (`tailRows`) the level is gathered
// on every rank and the rest of the hierarchy is the serial one, replicated: no communication below it.
struct Exchanger {
  int rank = 0, nProcs = 1;
  virtual ~Exchanger() {}
  virtual int allgatherv(const std::vector<char> &mine, std::vector<std::vector<char>> &all) = 0;
};

struct Halo {
  std::vector<int> sendPtr, sendIdx;  // per destination rank: owned rows whose values it needs
  std::vector<int> recvPtr;           // per source rank: ghost k of rank q is column n + recvPtr[q] + k
};

struct DistLevel {
  HostLevel L;            // A: n x (n + g), P: n x nc (local coarse ids), R = P^T
  Halo halo;
  std::vector<int> gid;   // global id (within this level) of every column, owned then ghosts
  int n = 0, g = 0;
};

struct DistHierarchy {
  std::vector<DistLevel> dist;
  std::vector<int> tailOff;   // nProcs + 1: rank segments of the first replicated level
  phb_amg_host tail;
  bool singular = false;
  double setupMs = 0.;
};

template <typename T> void put(std::vector<char> &b, const T &v) {
  const char *p = reinterpret_cast<const char *>(&v);
  b.insert(b.end(), p, p + sizeof(T));
}
template <typename T> T take(const char *&p) {
  T v;
  memcpy(&v, p, sizeof(T));
  p += sizeof(T);
  return v;
}

int build_dist_hierarchy(Exchanger &ex, HCsr A, Halo halo, std::vector<int> gid, Strength theta, int coarsest,
                         long long tailRows, double omegaP, DistHierarchy &H) {
  const auto t0 = std::chrono::steady_clock::now();
  const int NP = ex.nProcs, me = ex.rank;
  H.dist.clear();
  int localSingular = rows_sum_to_zero(A) ? 1 : 0;
  std::vector<std::vector<char>> all;
  typedef std::pair<int, double> Entry;
  auto byCol = [](const Entry &x, const Entry &y) { return x.first < y.first; };
  for (int level = 0;; ++level) {
    DistLevel D;
    const int n = A.n, g = A.m - A.n;
    D.n = n; D.g = g;
    D.halo = halo;
    D.gid = gid;
    // ---- rank-local aggregation (ghost columns are never strong, so aggregates stay inside the rank)
    std::vector<double> d = diagonal(A);
    int bad = 0;
    for (int i = 0; i < n; ++i)
      if (d[i] == 0.) { set_error("amg: zero diagonal in row %d of level %d", i, level); bad = 1; break; }
    std::vector<int> agg;
    std::vector<char> strong;
    int nc = 0;
    if (!bad) {
      D.L.diag = d;
      D.L.rho = gershgorin(A, d);
      D.L.wScale = theta.omegaS / D.L.rho;   // distributed levels: Gershgorin rule (no distributed power iteration)
      nc = aggregate(A, d, theta, agg, strong);
    }
    // ---- exchange 0: coarse sizes + the aggregate of every cell a neighbour holds as a ghost
    std::vector<char> blob;
    put<int>(blob, bad ? -1 : nc);
    put<int>(blob, localSingular);
    for (int q = 0; q < NP; ++q) {
      const int cnt = bad ? 0 : halo.sendPtr[q + 1] - halo.sendPtr[q];
      put<int>(blob, cnt);
      for (int k = 0; k < cnt; ++k) put<int>(blob, agg[halo.sendIdx[halo.sendPtr[q] + k]]);
    }
    PHB_CHECK(ex.allgatherv(blob, all));
    std::vector<int> ncAll(NP), off(NP + 1, 0);
    bool failed = false, singular = true;
    for (int q = 0; q < NP; ++q) {
      const char *p = all[q].data();
      ncAll[q] = take<int>(p);
      if (ncAll[q] < 0) failed = true;
      if (!take<int>(p)) singular = false;
    }
    if (failed) {
      if (!bad) set_error("amg: the hierarchy setup failed on another rank");
      return PHB_ERR_BREAKDOWN;
    }
    if (level == 0) H.singular = singular;
    for (int q = 0; q < NP; ++q) off[q + 1] = off[q] + ncAll[q];
    const bool lastDist = (long long)off[NP] <= tailRows || level + 1 >= 10;
    std::vector<int> ghostAgg(g, -1);  // GLOBAL coarse id of the aggregate of every ghost cell
    for (int q = 0; q < NP; ++q) {
      if (q == me) continue;
      const char *p = all[q].data();
      take<int>(p); take<int>(p);
      for (int dst = 0; dst < NP; ++dst) {
        const int cnt = take<int>(p);
        if (dst == me && cnt != halo.recvPtr[q + 1] - halo.recvPtr[q]) {
          set_error("amg: halo lists of ranks %d and %d disagree on level %d", me, q, level);
          return PHB_ERR_STATE;
        }
        for (int k = 0; k < cnt; ++k) {
          const int cid = take<int>(p);
          if (dst == me) ghostAgg[halo.recvPtr[q] + k] = off[q] + cid;
        }
      }
    }
    // ---- prolongator rows of the owned cells, GLOBAL coarse ids: P = (I - w Df^-1 Af) T where Af keeps
    // the couplings to other ranks (smooth basis functions across the interfaces) and lumps weak local ones
    std::vector<double> df(d);
    for (int i = 0; i < n; ++i)
      for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
        if (!strong[k] && A.ci[k] != i && A.ci[k] < n) df[i] += A.v[k];
    for (int i = 0; i < n; ++i)
      if (df[i] == 0. || (df[i] > 0.) != (d[i] > 0.)) df[i] = d[i];
    double rho = 0.;
    for (int i = 0; i < n; ++i) {
      double sum = std::fabs(df[i]);
      for (int k = A.rp[i]; k < A.rp[i + 1]; ++k)
        if (strong[k] || A.ci[k] >= n) sum += std::fabs(A.v[k]);
      rho = std::max(rho, sum / std::fabs(df[i]));
    }
    const double w = omegaP / std::max(rho, 1e-300);
    std::vector<std::vector<Entry>> Prow(n + g);
    for (int i = 0; i < n; ++i) {
      std::vector<Entry> &row = Prow[i];
      row.push_back({off[me] + agg[i], 1. - w});
      for (int k = A.rp[i]; k < A.rp[i + 1]; ++k) {
        const int j = A.ci[k];
        if (!(strong[k] || (j >= n && A.v[k] != 0.))) continue;
        const int c = j < n ? off[me] + agg[j] : ghostAgg[j - n];
        const double val = -w * A.v[k] / df[i];
        bool hit = false;
        for (auto &e : row)
          if (e.first == c) { e.second += val; hit = true; break; }
        if (!hit) row.push_back({c, val});
      }
      std::sort(row.begin(), row.end(), byCol);
    }
    // ---- exchange 1: the P rows of the cells my neighbours hold as ghosts
    blob.clear();
    for (int q = 0; q < NP; ++q) {
      const int cnt = halo.sendPtr[q + 1] - halo.sendPtr[q];
      put<int>(blob, cnt);
      for (int k = 0; k < cnt; ++k) {
        const std::vector<Entry> &row = Prow[halo.sendIdx[halo.sendPtr[q] + k]];
        put<int>(blob, (int)row.size());
        for (auto &e : row) { put<int>(blob, e.first); put<double>(blob, e.second); }
      }
    }
    PHB_CHECK(ex.allgatherv(blob, all));
    for (int q = 0; q < NP; ++q) {
      if (q == me) continue;
      const char *p = all[q].data();
      for (int dst = 0; dst < NP; ++dst) {
        const int cnt = take<int>(p);
        for (int k = 0; k < cnt; ++k) {
          const int len = take<int>(p);
          for (int e = 0; e < len; ++e) {
            const int c = take<int>(p);
            const double val = take<double>(p);
            if (dst == me) Prow[n + halo.recvPtr[q] + k].push_back({c, val});
          }
        }
      }
    }
    // coarse ghosts: every coarse id of another rank seen in these rows, grouped by owner (ascending global id)
    std::vector<int> cg;
    for (auto &row : Prow)
      for (auto &e : row)
        if (e.first < off[me] || e.first >= off[me + 1]) cg.push_back(e.first);
    std::sort(cg.begin(), cg.end());
    cg.erase(std::unique(cg.begin(), cg.end()), cg.end());
    const int gc = (int)cg.size();
    Halo ch;
    ch.recvPtr.assign(NP + 1, 0);
    std::vector<std::vector<int>> need(NP);
    for (int c : cg) {
      const int q = (int)(std::upper_bound(off.begin(), off.end(), c) - off.begin()) - 1;
      need[q].push_back(c - off[q]);
    }
    for (int q = 0; q < NP; ++q) ch.recvPtr[q + 1] = ch.recvPtr[q] + (int)need[q].size();
    auto localCol = [&](int c) {
      if (c >= off[me] && c < off[me + 1]) return c - off[me];
      return nc + (int)(std::lower_bound(cg.begin(), cg.end(), c) - cg.begin());
    };
    // Pe: all rows, local column numbering (own coarse ids, then coarse ghosts) -> Galerkin product
    // P : owned rows; on the last distributed level its columns index the gathered (global) coarse vector
    // R : transpose of the part of P inside this rank (restriction needs no communication)
    HCsr Pe, P, Ploc;
    Pe.n = n + g; Pe.m = nc + gc; Pe.rp.assign(1, 0);
    P.n = n; P.m = lastDist ? off[NP] : nc + gc; P.rp.assign(1, 0);
    Ploc.n = n; Ploc.m = nc; Ploc.rp.assign(1, 0);
    std::vector<Entry> tmp;
    for (int i = 0; i < n + g; ++i) {
      tmp.clear();
      for (auto &e : Prow[i]) tmp.push_back({localCol(e.first), e.second});
      std::sort(tmp.begin(), tmp.end(), byCol);
      for (auto &e : tmp) { Pe.ci.push_back(e.first); Pe.v.push_back(e.second); }
      Pe.rp.push_back((int)Pe.ci.size());
      if (i >= n) continue;
      if (lastDist) {
        for (auto &e : Prow[i]) { P.ci.push_back(e.first); P.v.push_back(e.second); }  // sorted by global id
      } else {
        for (auto &e : tmp) { P.ci.push_back(e.first); P.v.push_back(e.second); }
      }
      P.rp.push_back((int)P.ci.size());
      for (auto &e : tmp)
        if (e.first < nc) { Ploc.ci.push_back(e.first); Ploc.v.push_back(e.second); }
      Ploc.rp.push_back((int)Ploc.ci.size());
    }
    D.L.R = transpose(Ploc);
    HCsr AP = spgemm(A, Pe);
    HCsr Ac = spgemm(D.L.R, AP);   // nc x (nc + gc)
    // ---- exchange 2: tell every rank which of its coarse rows I hold as ghosts
    blob.clear();
    for (int q = 0; q < NP; ++q) {
      put<int>(blob, (int)need[q].size());
      for (int cid : need[q]) put<int>(blob, cid);
    }
    PHB_CHECK(ex.allgatherv(blob, all));
    ch.sendPtr.assign(NP + 1, 0);
    for (int q = 0; q < NP; ++q) {
      const char *p = all[q].data();
      for (int dst = 0; dst < NP; ++dst) {
        const int cnt = take<int>(p);
        for (int k = 0; k < cnt; ++k) {
          const int cid = take<int>(p);
          if (dst == me && q != me) ch.sendIdx.push_back(cid);
        }
      }
      ch.sendPtr[q + 1] = (int)ch.sendIdx.size();
    }
    std::vector<int> cgid(nc + gc);
    for (int i = 0; i < nc; ++i) cgid[i] = off[me] + i;
    for (int k = 0; k < gc; ++k) cgid[nc + k] = cg[k];
    D.L.P = std::move(P);
    D.L.A = std::move(A);
    H.dist.push_back(std::move(D));
    A = std::move(Ac);
    halo = std::move(ch);
    gid = std::move(cgid);
    localSingular = 1;
    if (lastDist) { H.tailOff = off; break; }
  }
  // ---- replicated tail: gather the level on every rank (rows in rank order, global column ids)
  std::vector<char> blob;
  put<int>(blob, A.n);
  put<int>(blob, (int)A.nnz());
  for (int i = 0; i <= A.n; ++i) put<int>(blob, A.rp[i]);
  for (long long k = 0; k < A.nnz(); ++k) put<int>(blob, gid[A.ci[k]]);
  for (long long k = 0; k < A.nnz(); ++k) put<double>(blob, A.v[k]);
  PHB_CHECK(ex.allgatherv(blob, all));
  HCsr G;
  G.n = G.m = H.tailOff[NP];
  G.rp.assign(1, 0);
  std::vector<Entry> row;
  for (int q = 0; q < NP; ++q) {
    const char *p = all[q].data();
    const int nr = take<int>(p), nz = take<int>(p);
    const int *rp = reinterpret_cast<const int *>(p);
    const int *ci = rp + nr + 1;
    const char *vp = reinterpret_cast<const char *>(ci + nz);
    for (int i = 0; i < nr; ++i) {
      row.clear();
      for (int k = rp[i]; k < rp[i + 1]; ++k) {
        double v;
        memcpy(&v, vp + (size_t)k * sizeof(double), sizeof(double));
        row.push_back({ci[k], v});
      }
      std::sort(row.begin(), row.end(), byCol);
      for (auto &e : row) { G.ci.push_back(e.first); G.v.push_back(e.second); }
      G.rp.push_back((int)G.ci.size());
    }
  }
  Strength tailRule = theta;
  tailRule.firstCoarse = 0;   // the gathered level is a Galerkin level of the global hierarchy
  PHB_CHECK(build_hierarchy(std::move(G), tailRule, coarsest, omegaP, H.tail.H));
  H.setupMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return PHB_OK;
}

// in-process exchanger: the ranks are threads of one process (CPU tests of the distributed setup)
struct ThreadBoard {
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  long gen = 0;
  std::vector<std::vector<char>> slots, published;
};
struct ThreadExchanger : Exchanger {
  ThreadBoard *B = nullptr;
  int allgatherv(const std::vector<char> &mine, std::vector<std::vector<char>> &all) override {
    std::unique_lock<std::mutex> lk(B->m);
    B->slots[rank] = mine;
    if (++B->arrived == nProcs) {
      B->published = B->slots;
      B->arrived = 0;
      B->gen++;
      B->cv.notify_all();
    } else {
      const long g = B->gen;
      B->cv.wait(lk, [&] { return B->gen != g; });
    }
    all = B->published;
    return PHB_OK;
  }
};

// sliced-ELL image of a host CSR matrix (diagonal moved to entry 0 when square)
void sell_from_csr(const HCsr &A, bool diagFirst, SellPattern &S, std::vector<double> &slotVals,
                   std::vector<int> *slotSrc = nullptr) {
  S.nRows = A.n; S.nCols = A.m;
  S.nSlices = (A.n + 31) / 32;
  S.hRowLen.assign(A.n, 0);
  S.nnz = A.nnz();
  for (int r = 0; r < A.n; ++r) S.hRowLen[r] = A.rp[r + 1] - A.rp[r];
  S.hSliceOff.assign(S.nSlices + 1, 0);
  for (int sl = 0; sl < S.nSlices; ++sl) {
    int w = 1;
    for (int r = sl * 32; r < std::min(A.n, sl * 32 + 32); ++r) w = std::max(w, S.hRowLen[r]);
    S.hSliceOff[sl + 1] = S.hSliceOff[sl] + w * 32;
  }
  S.nSlots = S.hSliceOff[S.nSlices];
  S.hCol.assign(S.nSlots, 0);
  slotVals.assign(S.nSlots, 0.);
  if (slotSrc) slotSrc->assign(S.nSlots, -1);   // CSR entry each slot holds (numeric re-setup)
  for (int sl = 0; sl < S.nSlices; ++sl) {
    const int w = (S.hSliceOff[sl + 1] - S.hSliceOff[sl]) / 32;
    for (int lane = 0; lane < 32; ++lane) {
      const int r = sl * 32 + lane;
      const int pad = diagFirst ? std::min(r, A.n - 1) : 0;
      int k = 0;
      if (r < A.n) {
        if (diagFirst)
          for (int j = A.rp[r]; j < A.rp[r + 1]; ++j)
            if (A.ci[j] == r) {
              const size_t slot = (size_t)S.hSliceOff[sl] + lane;
              S.hCol[slot] = r; slotVals[slot] = A.v[j]; k = 1;
              if (slotSrc) (*slotSrc)[slot] = j;
              break;
            }
        for (int j = A.rp[r]; j < A.rp[r + 1]; ++j) {
          if (diagFirst && A.ci[j] == r) continue;
          const size_t slot = (size_t)S.hSliceOff[sl] + (size_t)k * 32 + lane;
          S.hCol[slot] = A.ci[j]; slotVals[slot] = A.v[j];
          if (slotSrc) (*slotSrc)[slot] = j;
          ++k;
        }
      }
      for (; k < w; ++k) S.hCol[(size_t)S.hSliceOff[sl] + (size_t)k * 32 + lane] = pad;
    }
  }
}

// local block (owned rows x owned columns) of a sliced-ELL matrix as host CSR
HCsr csr_from_sell(const SellPattern &S, const std::vector<double> &slotVals, bool keepGhosts = false) {
  HCsr A;
  A.n = A.m = S.nRows;
  if (keepGhosts) A.m = S.nCols;
  A.rp.assign(A.n + 1, 0);
  std::vector<std::pair<int, double>> row;
  for (int r = 0; r < A.n; ++r) {
    const int sl = r >> 5, lane = r & 31;
    row.clear();
    for (int k = 0; k < S.hRowLen[r]; ++k) {
      const size_t slot = (size_t)S.hSliceOff[sl] + (size_t)k * 32 + lane;
      const int c = S.hCol[slot];
      if (c >= A.m) continue;  // ghost column dropped: rank-local preconditioner
      bool hit = false;
      for (auto &e : row)
        if (e.first == c) { e.second += slotVals[slot]; hit = true; break; }
      if (!hit) row.push_back({c, slotVals[slot]});
    }
    std::sort(row.begin(), row.end(), [](const std::pair<int, double> &x, const std::pair<int, double> &y) {
      return x.first < y.first;
    });
    for (auto &e : row) { A.ci.push_back(e.first); A.v.push_back(e.second); }
    A.rp[r + 1] = (int)A.ci.size();
  }
  return A;
}

// ===================================================================== device cycle
// The cycle runs in single precision by default (`amgPrecision single`): it is a preconditioner, the
// Krylov recurrences around it stay fp64, and every kernel here is bound by HBM bytes -- float matrices
// and vectors halve them.  T = float | double is the type of the level matrices and level vectors;
// TB / TY are the types of the right-hand side read and of the vector written (double where the cycle
// touches the Krylov vectors: b on level 0 and the final result).
constexpr int kThreads = 256;
constexpr int kBlocksPerSM = 8;
constexpr int kSubLanes = 8;        // lanes per row in the small-level kernel
constexpr int kSubRows = 200000;    // levels with at most this many rows use it

template <int MODE, int NC, typename T, typename TB, typename TY>
__device__ __forceinline__ void amg_epilogue(int row, const T (&acc)[NC], const T *__restrict__ x, int ldx,
                                             TY *__restrict__ y, int ldy, const TB *__restrict__ b,
                                             const T *__restrict__ w) {
  const T wr = MODE == 2 ? w[row] : T(0);
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const size_t iy = (size_t)i * ldy + row;
    if (MODE == 0) y[iy] = (TY)acc[i];
    if (MODE == 1) y[iy] = (TY)((T)b[iy] - acc[i]);
    if (MODE == 2) y[iy] = (TY)(x[(size_t)i * ldx + row] + wr * ((T)b[iy] - acc[i]));
    if (MODE == 3) y[iy] += (TY)acc[i];
  }
}

// MODE 0: y = M x          MODE 1: y = b - M x
// MODE 2: y = x + w (b - M x)   (damped Jacobi, out of place; w = omega / a_ii)
// MODE 3: y += M x
// NC components share the coefficients; x has leading dimension ldx, y and b have ldy.
// Streaming variant: warp <-> slice, lane <-> row (large levels, HBM-bound).
// WIDE: rows of more than seven entries (slice_dot_wide, kernels.cuh) -- an instantiation of its own, so that its
// registers do not cost the 5-entry level-0 sweeps their occupancy
template <int MODE, int NC, typename T, typename TB, typename TY, bool WIDE = false>
__global__ void __launch_bounds__(kThreads)
k_amg_spmv(SellView M, const T *__restrict__ vals, const T *__restrict__ x, int ldx, TY *__restrict__ y, int ldy,
           const TB *__restrict__ b, const T *__restrict__ w, const KrylovSums *S, int maxIters) {
  if (S && krylov_done(S, maxIters)) return;
  const int lane = threadIdx.x & 31;
  const int warpsPerBlock = blockDim.x >> 5;
  const int warp = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5);
  const int nWarps = gridDim.x * warpsPerBlock;
  for (int slice = warp; slice < M.nSlices; slice += nWarps) {
    const int off = __ldg(M.sliceOff + slice);
    const int wdt = (__ldg(M.sliceOff + slice + 1) - off) >> 5;
    const int row = slice * 32 + lane;
    T acc[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) acc[i] = T(0);
    if (WIDE) slice_dot_wide<NC>(M.col, vals, (size_t)off + lane, wdt, x, ldx, acc);
    else slice_dot_any<NC>(M.col, vals, (size_t)off + lane, wdt, x, ldx, acc);
    if (row < M.nRows) amg_epilogue<MODE, NC, T, TB, TY>(row, acc, x, ldx, y, ldy, b, w);
  }
}

// Small levels are latency-bound, not bandwidth-bound: kSubLanes lanes share a row, every lane issues its
// (up to 4) entries at once, partial sums meet in shuffles -- the dependent-load chain no longer grows
// with the row length (restriction rows hold 20-30 entries).
template <int MODE, int NC, typename T, typename TB, typename TY>
__global__ void __launch_bounds__(kThreads)
k_amg_spmv_sub(SellView M, const T *__restrict__ vals, const T *__restrict__ x, int ldx, TY *__restrict__ y, int ldy,
               const TB *__restrict__ b, const T *__restrict__ w, const KrylovSums *S, int maxIters) {
  if (S && krylov_done(S, maxIters)) return;
  const int g = threadIdx.x & (kSubLanes - 1);
  const int row = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) / kSubLanes);
  const bool live = row < M.nRows;
  int off = 0, wdt = 0;
  if (live) {
    off = __ldg(M.sliceOff + (row >> 5));
    wdt = (__ldg(M.sliceOff + (row >> 5) + 1) - off) >> 5;
  }
  const size_t base = (size_t)off + (row & 31);
  T acc[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) acc[i] = T(0);
  for (int k0 = 0; k0 < wdt; k0 += 4 * kSubLanes) {
    int c[4];
    T a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + j * kSubLanes + g;
      c[j] = k < wdt ? __ldg(M.col + base + (size_t)k * 32) : -1;
      a[j] = k < wdt ? __ldg(vals + base + (size_t)k * 32) : T(0);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c[j] >= 0) {
#pragma unroll
        for (int i = 0; i < NC; ++i) acc[i] = fma(a[j], __ldg(x + (size_t)i * ldx + c[j]), acc[i]);
      }
  }
#pragma unroll
  for (int i = 0; i < NC; ++i)
#pragma unroll
    for (int o = kSubLanes / 2; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  if (live && g == 0) amg_epilogue<MODE, NC, T, TB, TY>(row, acc, x, ldx, y, ldy, b, w);
}

// x = w .* b  (first pre-smoothing sweep from a zero guess)
template <typename T, typename TB>
__global__ void k_amg_scale(int n, int nc, int ld, const T *__restrict__ w, const TB *__restrict__ b,
                            T *__restrict__ x, const KrylovSums *S, int maxIters) {
  if (S && krylov_done(S, maxIters)) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const T wi = w[i];
    for (int c = 0; c < nc; ++c) x[(size_t)c * ld + i] = wi * (T)b[(size_t)c * ld + i];
  }
}

// coarsest level: x = Ainv b, one warp per row of the dense inverse (kept in fp64: it is tiny)
template <typename TB, typename TY>
__global__ void k_amg_dense(int n, int nc, int ld, const double *__restrict__ Ainv, const TB *__restrict__ b,
                            TY *__restrict__ x, const KrylovSums *S, int maxIters) {
  if (S && krylov_done(S, maxIters)) return;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  for (int c = 0; c < nc; ++c) {
    double acc = 0.;
    for (int k = lane; k < n; k += 32) acc = fma(Ainv[(size_t)row * n + k], (double)b[(size_t)c * ld + k], acc);
    acc = warp_sum(acc);
    if (lane == 0) x[(size_t)c * ld + row] = (TY)acc;
  }
}

__global__ void k_amg_to_float(long long n, const double *__restrict__ a, float *__restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (float)a[i];
}

// send list of a distributed level -> contiguous buffer [comp][nSend]
template <typename T>
__global__ void k_amg_pack(int nSend, int nc, int ld, const int *__restrict__ sendIdx, const T *__restrict__ x,
                           T *__restrict__ buf, const KrylovSums *S, int maxIters) {
  if (S && krylov_done(S, maxIters)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nSend) return;
  const int src = sendIdx[i];
  for (int c = 0; c < nc; ++c) buf[(size_t)c * nSend + i] = x[(size_t)c * ld + src];
}

// ---- ghost refresh over NVLink peer memory (one CTA): gather my send list of x, store it straight into the
// ghost segments of the peers' copy of the same level vector, publish an epoch at each peer, wait for the epochs
// of the peers I receive from.  sendIdx == nullptr: contiguous segment [sendOff, sendOff + sendCnt) (tail gather).
// Safety of re-using a ghost segment: between two refreshes of the same vector every rank passes a refresh or
// reduction involving all its neighbours (other levels, the tail gather, the Krylov all-reduces), which it can
// only complete after those neighbours have issued -- in stream order -- the kernels that read the old ghosts.
struct AmgPeerView {
  char *block;
  char *peer[kMaxPeers];
  int rank, nProcs;
};
__device__ __forceinline__ unsigned long long *amg_flag(char *block, int ch, int src) {
  return reinterpret_cast<unsigned long long *>(block) + (size_t)ch * kMaxPeers + src;
}
__device__ __forceinline__ unsigned long long *amg_epoch(char *block, int ch) {
  return reinterpret_cast<unsigned long long *>(block + kAmgPeerChannels * kMaxPeers * sizeof(unsigned long long)) + ch;
}
template <typename T>
__global__ void __launch_bounds__(1024)
k_amg_peer_halo(AmgPeerView pv, int ch, AmgPeerLevel L, int vec, const T *__restrict__ x, int nc, int ld,
                const int *__restrict__ sendIdx, const KrylovSums *S, int maxIters) {
  if (S && krylov_done(S, maxIters)) return;
  __shared__ unsigned long long se;
  if (threadIdx.x == 0) {
    unsigned long long *ep = amg_epoch(pv.block, ch);
    se = *ep + 1;
    *ep = se;
  }
  __syncthreads();
  const unsigned long long e = se;
  for (int q = 0; q < pv.nProcs; ++q) {
    const int cnt = L.sendCnt[q];
    if (q == pv.rank || cnt == 0) continue;
    T *dst = reinterpret_cast<T *>(pv.peer[q] + L.vecOff[vec][q]);
    const int off = L.sendOff[q];
    for (int j = threadIdx.x; j < cnt * nc; j += blockDim.x) {
      const int c = j / cnt, i = j - c * cnt;
      const int src = sendIdx ? sendIdx[off + i] : off + i;
      dst[(size_t)c * L.ld[q] + L.recvOff[q] + i] = x[(size_t)c * ld + src];
    }
  }
  __threadfence_system();
  __syncthreads();
  const int t = threadIdx.x;
  if (t < pv.nProcs && t != pv.rank) {
    if (L.sendCnt[t] > 0) st_release_sys(amg_flag(pv.peer[t], ch, pv.rank), e);
    if (L.recvCnt[t] > 0) {
      const unsigned long long *f = amg_flag(pv.block, ch, t);
      while (ld_acquire_sys(f) < e) {}
    }
  }
  __syncthreads();
}

// ---- fused small levels: one launch runs levels fuseFrom .. L-1 down, the dense coarsest solve and the way back up.
// These levels hold 2 % of the cycle's data but, as separate launches of 5-10 us each, 30 % of its time (and all of
// it when the mesh is split over many GPUs).  The pre-smoothing sweep from a zero guess is folded away: its result
// x = w.*b is recomputed where it is needed (gather side of the residual, row side of the prolongation), so a level
// is four phases: residual, restriction | prolongation, post-smoothing.  Phases are separated by a sense-reversing
// grid barrier; the grid is one CTA per SM, all resident.  A spin that runs far too long raises `bar[2]` and gives up
// (a hung kernel would take the whole process with it).
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void grid_barrier(unsigned *bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned gen = ld_acquire_gpu(bar + 1);
    __threadfence();
    if (atomicAdd(bar, 1u) == gridDim.x - 1) {
      bar[0] = 0u;
      __threadfence();
      st_release_gpu(bar + 1, gen + 1u);
    } else {
      long long spins = 0;
      while (ld_acquire_gpu(bar + 1) == gen)
        if (++spins > (1ll << 31)) { bar[2] = 1u; break; }
    }
  }
  __syncthreads();
}

template <typename T, int NC>
__global__ void __launch_bounds__(1024)
k_amg_tail(const AmgTailOp *__restrict__ ops, int nOps, unsigned *bar, const KrylovSums *S, int maxIters) {
  if (S && krylov_done(S, maxIters)) return;   // the same answer in every CTA: nobody waits for a CTA that left
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nThreads = gridDim.x * blockDim.x;
  // the op list once into shared memory: one global round trip per launch instead of one per phase
  __shared__ AmgTailOp sOps[kMaxTailOps];
  {
    const int words = nOps * (int)(sizeof(AmgTailOp) / sizeof(int));
    const int *src = reinterpret_cast<const int *>(ops);
    int *dst = reinterpret_cast<int *>(sOps);
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
  }
  for (int o = 0; o < nOps; ++o) {
    const AmgTailOp op = sOps[o];
    if (op.kind == 2) {   // dense coarsest solve: one warp per row of the fp64 inverse
      const double *Ainv = static_cast<const double *>(op.vals);
      const T *b = static_cast<const T *>(op.b);
      T *x = static_cast<T *>(op.y);
      const int lane = threadIdx.x & 31;
      for (int row = tid >> 5; row < op.n; row += nThreads >> 5) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          double acc = 0.;
          for (int k = lane; k < op.n; k += 32) acc = fma(Ainv[(size_t)row * op.n + k], (double)b[(size_t)c * op.ld + k], acc);
          acc = warp_sum(acc);
          if (lane == 0) x[(size_t)c * op.ld + row] = (T)acc;
        }
      }
    } else {
      const T *vals = static_cast<const T *>(op.vals), *w = static_cast<const T *>(op.w);
      const T *b = static_cast<const T *>(op.b), *xin = static_cast<const T *>(op.x);
      T *y = static_cast<T *>(op.y);
      const int L = op.lanes, g = tid & (L - 1);
      const int rowsPerPass = nThreads / L;
      for (int row0 = 0; row0 < op.n; row0 += rowsPerPass) {   // uniform trip count: the shuffles below stay converged
        const int row = row0 + tid / L;
        const bool live = row < op.n;
        T acc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = T(0);
        if (live) {
          const int off = op.sliceOff[row >> 5];
          const int wdt = (op.sliceOff[(row >> 5) + 1] - off) >> 5;
          const size_t base = (size_t)off + (row & 31);
          for (int k0 = g; k0 < wdt; k0 += 4 * L) {   // four entries in flight per lane: the phase is latency-bound
            int col[4];
            T a[4], wc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int k = k0 + j * L;
              col[j] = k < wdt ? op.col[base + (size_t)k * 32] : -1;
              a[j] = k < wdt ? vals[base + (size_t)k * 32] : T(0);
            }
            if (op.kind == 0) {          // gathered vector = w .* b of this level
#pragma unroll
              for (int j = 0; j < 4; ++j) wc[j] = col[j] >= 0 ? w[col[j]] : T(0);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (col[j] >= 0) {
#pragma unroll
                  for (int c = 0; c < NC; ++c) acc[c] = fma(a[j], wc[j] * b[(size_t)c * op.ldIn + col[j]], acc[c]);
                }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (col[j] >= 0) {
#pragma unroll
                  for (int c = 0; c < NC; ++c) acc[c] = fma(a[j], xin[(size_t)c * op.ldIn + col[j]], acc[c]);
                }
            }
          }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c)
          for (int s = L >> 1; s > 0; s >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], s);
        if (live && g == 0) {
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const size_t i = (size_t)c * op.ld + row;
            if (op.kind == 0) y[i] = b[i] - acc[c];
            else if (op.kind == 1) y[i] = acc[c];
            else if (op.kind == 3) y[i] = w[row] * b[i] + acc[c];
            else y[i] = xin[i] + w[row] * (b[i] - acc[c]);
          }
        }
      }
    }
    if (o + 1 < nOps) grid_barrier(bar);
  }
}

// max relative deviation of `a` from ratio * ref over the slots, ratio = a[first] / ref[first]
__global__ void __launch_bounds__(kThreads)
k_amg_changed(long long nSlots, const double *__restrict__ a, const double *__restrict__ ref, int first,
              double *partials, unsigned *ticket, double *out) {
  const double ratio = ref[first] != 0. ? a[first] / ref[first] : 1.;
  double dev = 0.;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nSlots;
       i += (long long)gridDim.x * blockDim.x) {
    const double r = ratio * ref[i];
    const double d = fabs(a[i] - r);
    if (d > 0.) dev = fmax(dev, d / fmax(fabs(r), 1e-300));
  }
  double v[1] = {dev};
  if (grid_reduce<1, true>(v, partials, ticket, out) && (threadIdx.x & 31) == 0) out[1] = ratio;
}

int grid_rows(const phb_ctx *c, long long rows) {
  const long long g = (rows + kThreads - 1) / kThreads;
  const long long cap = (long long)c->numSMs * kBlocksPerSM;
  return (int)std::max<long long>(1, std::min(g, cap));
}

template <int MODE, typename T, typename TB, typename TY>
void launch(phb_solver *s, const SellPattern &P, const T *vals, const T *x, int ldx, TY *y, int ldy, const TB *b,
            const T *w, bool inLoop) {
  const SellView V = view_of(&P);
  const KrylovSums *S = inLoop ? s->sums.p : nullptr;
  if (P.nRows <= kSubRows) {
    const int grid = (int)(((long long)P.nRows * kSubLanes + kThreads - 1) / kThreads);
    if (s->nComp == 1)
      PHB_LAUNCH(s->ctx, (k_amg_spmv_sub<MODE, 1, T, TB, TY>), grid, kThreads, 0, V, vals, x, ldx, y, ldy, b, w, S,
                 s->maxIters);
    else
      PHB_LAUNCH(s->ctx, (k_amg_spmv_sub<MODE, 2, T, TB, TY>), grid, kThreads, 0, V, vals, x, ldx, y, ldy, b, w, S,
                 s->maxIters);
    return;
  }
  const int grid = grid_rows(s->ctx, (long long)P.nSlices * 32);
  // mean slice width beyond seven entries: Galerkin operators and restrictions
  const bool wide = P.nSlices > 0 && P.nSlots > (long long)P.nSlices * 32 * 7;
  if (wide) {
    if (s->nComp == 1)
      PHB_LAUNCH(s->ctx, (k_amg_spmv<MODE, 1, T, TB, TY, true>), grid, kThreads, 0, V, vals, x, ldx, y, ldy, b, w, S,
                 s->maxIters);
    else
      PHB_LAUNCH(s->ctx, (k_amg_spmv<MODE, 2, T, TB, TY, true>), grid, kThreads, 0, V, vals, x, ldx, y, ldy, b, w, S,
                 s->maxIters);
    return;
  }
  if (s->nComp == 1)
    PHB_LAUNCH(s->ctx, (k_amg_spmv<MODE, 1, T, TB, TY>), grid, kThreads, 0, V, vals, x, ldx, y, ldy, b, w, S,
               s->maxIters);
  else
    PHB_LAUNCH(s->ctx, (k_amg_spmv<MODE, 2, T, TB, TY>), grid, kThreads, 0, V, vals, x, ldx, y, ldy, b, w, S,
               s->maxIters);
}

// typed views of the byte buffers of a level
template <typename T> T *as(const phb::DevBuf<double> &b) { return reinterpret_cast<T *>(b.p); }
template <typename T> int alloc_as(phb::DevBuf<double> &b, size_t count, cudaStream_t st) {
  PHB_CHECK(b.alloc((count * sizeof(T) + 7) / 8));
  return b.zero(st);
}
template <typename T> int upload_as(phb::DevBuf<double> &b, const std::vector<double> &h, cudaStream_t st) {
  std::vector<T> t(h.begin(), h.end());
  PHB_CHECK(b.alloc((t.size() * sizeof(T) + 7) / 8));
  if (!t.empty()) PHB_CUDA(cudaMemcpyAsync(b.p, t.data(), t.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  PHB_CUDA(cudaStreamSynchronize(st));  // t goes out of scope
  return PHB_OK;
}

template <typename T>
int upload_mat(phb_ctx *c, const HCsr &H, bool diagFirst, AmgMat &M, phb::DevBuf<int> *slotSrc = nullptr) {
  std::vector<double> slotVals;
  std::vector<int> src;
  sell_from_csr(H, diagFirst, M.pat, slotVals, slotSrc ? &src : nullptr);
  if (slotSrc) PHB_CHECK(slotSrc->upload(src, c->stream));
  PHB_CHECK(M.pat.sliceOff.upload(M.pat.hSliceOff, c->stream));
  PHB_CHECK(M.pat.rowLen.upload(M.pat.hRowLen, c->stream));
  PHB_CHECK(M.pat.col.upload(M.pat.hCol, c->stream));
  PHB_CHECK(upload_as<T>(M.vals, slotVals, c->stream));
  return PHB_OK;
}


// ===================================================================== numeric re-setup on the device
#include "amg_refresh.cuh"

// level-0 CSR entry -> sliced-ELL slot of the Krylov matrix; false when two slots of a row share a column (the host
// setup sums them, the gather kernel cannot)
bool csr_slot_sources(const SellPattern &S, const HCsr &A, std::vector<int> &src) {
  src.assign(A.ci.size(), -1);
  for (int r = 0; r < A.n; ++r) {
    const int sl = r >> 5, lane = r & 31;
    for (int k = 0; k < S.hRowLen[r]; ++k) {
      const size_t slot = (size_t)S.hSliceOff[sl] + (size_t)k * 32 + lane;
      const int c = S.hCol[slot];
      if (c >= A.m) continue;
      const auto b = A.ci.begin() + A.rp[r], e = A.ci.begin() + A.rp[r + 1];
      const auto it = std::lower_bound(b, e, c);
      if (it == e || *it != c) return false;
      int &dst = src[it - A.ci.begin()];
      if (dst >= 0) return false;
      dst = (int)slot;
    }
  }
  for (int v : src)
    if (v < 0) return false;
  return true;
}

// uploads the symbolic data of level l (h = its host level; `last`: no coarsening below it)
int capture_level(phb_solver *s, const HostLevel &h, int l, bool last, AmgRefreshLevel &R, double &bytes) {
  cudaStream_t st = s->ctx->stream;
  R.n = h.A.n;
  R.nnzA = h.A.nnz();
  PHB_CHECK(R.aRp.upload(h.A.rp, st)); PHB_CHECK(R.aCi.upload(h.A.ci, st));
  PHB_CHECK(R.aV.alloc((size_t)R.nnzA)); PHB_CHECK(R.diag.alloc((size_t)R.n));
  bytes += 4. * (R.n + 1) + 12. * R.nnzA + 8. * R.n;
  std::vector<int> src, rsrc;
  std::vector<unsigned char> strong;
  if (l == 0) {
    if (!csr_slot_sources(*s->pat, h.A, src)) return 1;
    PHB_CHECK(R.aSrc.upload(src, st));
    bytes += 4. * R.nnzA;
  }
  if (!last) {
    R.nc = h.P.m;
    R.nnzP = h.P.nnz();
    R.nnzAP = (long long)h.apCi.size();
    strong.assign(h.strong.begin(), h.strong.end());
    PHB_CHECK(R.strong.upload(strong, st)); PHB_CHECK(R.agg.upload(h.agg, st)); PHB_CHECK(R.df.alloc((size_t)R.n));
    PHB_CHECK(R.pRp.upload(h.P.rp, st)); PHB_CHECK(R.pCi.upload(h.P.ci, st)); PHB_CHECK(R.pV.alloc((size_t)R.nnzP));
    // R = P^T holds the entries of P in the order transpose() emits them
    rsrc.resize((size_t)R.nnzP);
    std::vector<int> fill(h.R.rp.begin(), h.R.rp.end() - 1);
    for (int i = 0; i < h.P.n; ++i)
      for (int k = h.P.rp[i]; k < h.P.rp[i + 1]; ++k) rsrc[fill[h.P.ci[k]]++] = k;
    PHB_CHECK(R.rRp.upload(h.R.rp, st)); PHB_CHECK(R.rCi.upload(h.R.ci, st)); PHB_CHECK(R.rSrc.upload(rsrc, st));
    PHB_CHECK(R.rV.alloc((size_t)R.nnzP));
    PHB_CHECK(R.apRp.upload(h.apRp, st)); PHB_CHECK(R.apCi.upload(h.apCi, st)); PHB_CHECK(R.apV.alloc((size_t)R.nnzAP));
    bytes += 1. * R.nnzA + 12. * R.n + 8. * (R.n + 1) + 4. * (R.nc + 1) + 32. * R.nnzP + 12. * R.nnzAP;
  }
  PHB_CUDA(cudaStreamSynchronize(st));   // the staging vectors go out of scope
  return PHB_OK;
}

inline RfCsr rf_view(const phb::DevBuf<int> &rp, const phb::DevBuf<int> &ci, const phb::DevBuf<double> &v) {
  return RfCsr{rp.p, ci.p, v.p};
}

// values of every level from the current level-0 matrix (s->dVals); the patterns, the cycle's buffers and the captured
// graphs stay as they are
template <typename T>
int refresh_numeric(phb_solver *s) {
  phb_ctx *c = s->ctx;
  AmgData &D = s->amg;
  AmgRefresh &F = *D.refresh;
  const int nLev = (int)D.lev.size();
  const double omegaP = 4. / 3.;
  cudaEvent_t e0, e1;
  PHB_CUDA(cudaEventCreate(&e0)); PHB_CUDA(cudaEventCreate(&e1));
  PHB_CUDA(cudaEventRecord(e0, c->stream));
  PHB_CHECK(F.scal.zero(c->stream));
  PHB_CHECK(F.flag.zero(c->stream));
  {
    AmgRefreshLevel &R0 = *F.lev[0];
    PHB_LAUNCH(c, k_rf_gather, grid_rows(c, R0.nnzA), kThreads, 0, R0.nnzA, R0.aSrc.p, s->dVals, R0.aV.p);
  }
  for (int l = 0; l < nLev; ++l) {
    AmgRefreshLevel &R = *F.lev[l];
    AmgLevel &L = *D.lev[l];
    const bool last = l + 1 == nLev;
    double *scal = F.scal.p + 4 * l;
    const RfCsr A = rf_view(R.aRp, R.aCi, R.aV);
    PHB_LAUNCH(c, k_rf_diag, grid_rows(c, R.n), kThreads, 0, R.n, A, last ? (const unsigned char *)nullptr : R.strong.p,
               R.diag.p, last ? (double *)nullptr : R.df.p, scal);
    PHB_LAUNCH(c, k_rf_weights<T>, grid_rows(c, R.n), kThreads, 0, R.n, R.diag.p, scal, L.omegaEff, as<T>(L.w));
    if (l > 0)
      PHB_LAUNCH(c, k_rf_fill<T>, grid_rows(c, L.A.pat.nSlots), kThreads, 0, (long long)L.A.pat.nSlots, R.aSell.p, R.aV.p,
                 as<T>(L.A.vals));
    if (last) break;
    AmgRefreshLevel &N = *F.lev[l + 1];
    PHB_LAUNCH(c, k_rf_prolong<8>, grid_rows(c, 8LL * R.n), kThreads, 0, R.n, A, R.strong.p, R.agg.p, R.df.p, scal, omegaP,
               R.pRp.p, R.pCi.p, R.pV.p);
    PHB_LAUNCH(c, k_rf_permute, grid_rows(c, R.nnzP), kThreads, 0, R.nnzP, R.rSrc.p, R.pV.p, R.rV.p);
    PHB_LAUNCH(c, k_rf_product<16>, grid_rows(c, 16LL * R.n), kThreads, 0, R.n, A, rf_view(R.pRp, R.pCi, R.pV), R.apRp.p,
               R.apCi.p, R.apV.p);
    PHB_LAUNCH(c, k_rf_product<16>, grid_rows(c, 16LL * R.nc), kThreads, 0, R.nc, rf_view(R.rRp, R.rCi, R.rV),
               rf_view(R.apRp, R.apCi, R.apV), N.aRp.p, N.aCi.p, N.aV.p);
    PHB_LAUNCH(c, k_rf_fill<T>, grid_rows(c, L.P.pat.nSlots), kThreads, 0, (long long)L.P.pat.nSlots, R.pSell.p, R.pV.p,
               as<T>(L.P.vals));
    PHB_LAUNCH(c, k_rf_fill<T>, grid_rows(c, L.R.pat.nSlots), kThreads, 0, (long long)L.R.pat.nSlots, R.rSell.p, R.rV.p,
               as<T>(L.R.vals));
  }
  if (D.denseCoarse) {
    AmgRefreshLevel &R = *F.lev[nLev - 1];
    const int n = R.n;
    double *scal = F.scal.p + 4 * (nLev - 1);
    PHB_LAUNCH(c, k_rf_diag_mean, 1, 1024, 0, n, R.diag.p, scal);
    PHB_LAUNCH(c, k_rf_dense_fill, grid_rows(c, (long long)n * n), kThreads, 0, n, scal, F.singular ? 1 : 0, F.dense[0].p);
    PHB_LAUNCH(c, k_rf_dense_scatter, grid_rows(c, n), kThreads, 0, n, rf_view(R.aRp, R.aCi, R.aV), F.dense[0].p);
    const int tiles = (n + kGjTile - 1) / kGjTile;
    int cur = 0;
    for (int k0 = 0; k0 < n; k0 += kGjBlock, cur ^= 1)
      PHB_LAUNCH(c, k_rf_gj_step, dim3(tiles, tiles), 256, 0, n, k0, F.dense[cur].p, F.dense[cur ^ 1].p, F.flag.p);
    PHB_CUDA(cudaMemcpyAsync(D.coarseInv.p, F.dense[cur].p, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToDevice,
                             c->stream));
  }
  // level 0 of the cycle = the current matrix
  PHB_CUDA(cudaMemcpyAsync(D.refVals.p, s->dVals, (size_t)s->pat->nSlots * sizeof(double), cudaMemcpyDeviceToDevice,
                           c->stream));
  if (sizeof(T) == 4)
    PHB_LAUNCH(c, k_amg_to_float, grid_rows(c, s->pat->nSlots), kThreads, 0, s->pat->nSlots, D.refVals.p, D.refValsF.p);
  PHB_CUDA(cudaEventRecord(e1, c->stream));
  // zero diagonal on some level / vanishing pivot in the dense inverse: the host setup decides what to do
  int flag = 0;
  std::vector<double> scalH((size_t)4 * nLev);
  PHB_CUDA(cudaMemcpyAsync(&flag, F.flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  PHB_CUDA(cudaMemcpyAsync(scalH.data(), F.scal.p, scalH.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PHB_CUDA(cudaStreamSynchronize(c->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  for (int l = 0; l < nLev; ++l)
    if (scalH[4 * l + 3] != 0.) flag = 1;
  if (flag) return 1;
  D.refreshMs = ms;
  D.refreshes++;
  D.itersAfterRefresh = -1;
  D.stale = false;
  return PHB_OK;
}

template <typename T> int build_tail_ops(phb_solver *s);

template <typename T>
int rebuild_t(phb_solver *s) {
  phb_ctx *c = s->ctx;
  AmgData &D = s->amg;
  const SellPattern *P = s->pat;
  std::vector<double> slotVals((size_t)P->nSlots);
  PHB_CUDA(cudaMemcpyAsync(slotVals.data(), s->dVals, slotVals.size() * sizeof(double), cudaMemcpyDeviceToHost,
                           c->stream));
  PHB_CUDA(cudaStreamSynchronize(c->stream));
  HostHierarchy H;
  const bool keep = D.refreshMode != 0;
  PHB_CHECK(build_hierarchy(csr_from_sell(*P, slotVals), strength_of(D), D.coarsest, 4. / 3., H, keep));
  D.lev.clear();
  D.nDist = 0;
  D.refresh.reset();
  std::unique_ptr<AmgRefresh> F(keep ? new AmgRefresh() : nullptr);
  const int nLev = (int)H.lev.size();
  for (int l = 0; l < nLev; ++l) {
    std::unique_ptr<AmgLevel> L(new AmgLevel());
    HostLevel &h = H.lev[l];
    L->n = h.A.n;
    AmgRefreshLevel *RL = nullptr;
    if (F) { F->lev.emplace_back(new AmgRefreshLevel()); RL = F->lev.back().get(); }
    if (l > 0) PHB_CHECK(upload_mat<T>(c, h.A, true, L->A, RL ? &RL->aSell : nullptr));
    if (l + 1 < nLev) {
      PHB_CHECK(upload_mat<T>(c, h.P, false, L->P, RL ? &RL->pSell : nullptr));
      PHB_CHECK(upload_mat<T>(c, h.R, false, L->R, RL ? &RL->rSell : nullptr));
    }
    if (F) {
      const int rc = capture_level(s, h, l, l + 1 == nLev, *RL, F->bytes);
      if (rc < 0) return rc;
      if (rc > 0) F.reset();   // pattern the gather cannot follow: host setups only
    }
    std::vector<double> w(L->n);
    for (int i = 0; i < L->n; ++i) w[i] = h.wScale / h.diag[i];
    L->omegaEff = h.wScale * h.rho;
    PHB_CHECK(upload_as<T>(L->w, w, c->stream));
    L->ld = l == 0 ? P->nCols : L->n;                      // level 0 vectors carry (zero) ghost entries
    const size_t len = (size_t)L->ld * s->nComp;
    PHB_CHECK(alloc_as<T>(L->x, len, c->stream)); PHB_CHECK(alloc_as<T>(L->x2, len, c->stream));
    PHB_CHECK(alloc_as<T>(L->r, len, c->stream));
    if (l > 0) PHB_CHECK(alloc_as<T>(L->b, len, c->stream));
    D.lev.push_back(std::move(L));
  }
  D.nCoarse = H.lev.back().A.n;
  D.denseCoarse = !H.coarseInv.empty();
  if (D.denseCoarse) PHB_CHECK(D.coarseInv.upload(H.coarseInv, c->stream));
  if (F) {
    F->singular = H.singular;
    PHB_CHECK(F->scal.alloc((size_t)4 * nLev));
    PHB_CHECK(F->flag.alloc(1));
    if (D.denseCoarse)
      for (int k = 0; k < 2; ++k) PHB_CHECK(F->dense[k].alloc((size_t)D.nCoarse * D.nCoarse));
    D.refresh = std::move(F);
  }
  D.refreshes = 0;
  D.itersAfterRefresh = -1;
  // level 0 operator of the cycle = the matrix the hierarchy was built from (kept in fp64 for the
  // change test, plus the float image the single-precision cycle streams)
  PHB_CHECK(D.refVals.alloc((size_t)P->nSlots));
  PHB_CUDA(cudaMemcpyAsync(D.refVals.p, s->dVals, (size_t)P->nSlots * sizeof(double), cudaMemcpyDeviceToDevice,
                           c->stream));
  if (sizeof(T) == 4) {
    PHB_CHECK(D.refValsF.alloc((size_t)P->nSlots));
    PHB_LAUNCH(c, k_amg_to_float, grid_rows(c, P->nSlots), kThreads, 0, P->nSlots, D.refVals.p, D.refValsF.p);
  }
  PHB_CUDA(cudaStreamSynchronize(c->stream));
  PHB_CHECK(build_tail_ops<T>(s));
  D.src = P;
  D.nComp = s->nComp;
  D.builtSingle = sizeof(T) == 4;
  D.built = true;
  D.setups++;
  D.setupMs = H.setupMs;
  D.opComplexity = H.opComplexity;
  D.itersAfterSetup = -1;
  D.stale = false;
  if (s->graphExec) { cudaGraphExecDestroy(s->graphExec); s->graphExec = nullptr; }
  return PHB_OK;
}

struct NcclExchanger : Exchanger {
  phb_ctx *c = nullptr;
  int allgatherv(const std::vector<char> &mine, std::vector<std::vector<char>> &all) override {
    return comm_allgatherv_host(c, mine, all);
  }
};

template <typename T>
int finish_level(phb_solver *s, AmgLevel &L, const HostLevel &h, int ld, bool hasCoarse, bool uploadA) {
  phb_ctx *c = s->ctx;
  AmgData &D = s->amg;
  L.n = h.A.n;
  if (uploadA) PHB_CHECK(upload_mat<T>(c, h.A, true, L.A));
  if (hasCoarse) {
    PHB_CHECK(upload_mat<T>(c, h.P, false, L.P));
    PHB_CHECK(upload_mat<T>(c, h.R, false, L.R));
  }
  std::vector<double> w(L.n);
  for (int i = 0; i < L.n; ++i) w[i] = h.wScale / h.diag[i];
  L.omegaEff = h.wScale * h.rho;
  PHB_CHECK(upload_as<T>(L.w, w, c->stream));
  L.ld = ld;
  const size_t len = (size_t)ld * s->nComp;
  PHB_CHECK(alloc_as<T>(L.x, len, c->stream)); PHB_CHECK(alloc_as<T>(L.x2, len, c->stream));
  PHB_CHECK(alloc_as<T>(L.r, len, c->stream)); PHB_CHECK(alloc_as<T>(L.b, len, c->stream));
  return PHB_OK;
}

// Peer-memory ghost refreshes of the distributed levels: move the vectors peers write into (x, x2 per distributed
// level, the gathered right-hand side of the first replicated level) into one CUDA-IPC block, exchange handles
// and layouts, open the peers' blocks.  Collective: every rank calls it with the same level structure.
template <typename T>
int setup_peer(phb_solver *s, Exchanger &ex) {
  phb_ctx *c = s->ctx;
  AmgData &D = s->amg;
  const int NP = c->nProcs, me = c->rank, nD = D.nDist;
  auto align = [](size_t b) { return (b + 255) & ~(size_t)255; };
  std::vector<size_t> off(2 * nD + 1);
  size_t total = kAmgPeerHeaderBytes;
  for (int l = 0; l < nD; ++l) {
    const size_t vb = align((size_t)D.lev[l]->ld * s->nComp * sizeof(T));
    off[2 * l] = total; total += vb;
    off[2 * l + 1] = total; total += vb;
  }
  AmgLevel &TL = *D.lev[nD];
  off[2 * nD] = total;
  total += align((size_t)TL.ld * s->nComp * sizeof(T));
  std::unique_ptr<AmgPeer> P(new AmgPeer());
  P->bytes = total;
  PHB_CUDA(cudaMalloc((void **)&P->block, total));
  PHB_CUDA(cudaMemset(P->block, 0, total));
  cudaIpcMemHandle_t h;
  PHB_CUDA(cudaIpcGetMemHandle(&h, P->block));
  std::vector<char> blob;
  blob.insert(blob.end(), (const char *)&h, (const char *)&h + sizeof(h));
  put<int>(blob, nD);
  for (int l = 0; l < nD; ++l) {
    put<unsigned long long>(blob, off[2 * l]);
    put<unsigned long long>(blob, off[2 * l + 1]);
    put<int>(blob, D.lev[l]->ld);
    for (int q = 0; q < NP; ++q) put<int>(blob, D.lev[l]->recvOff[q]);
  }
  put<unsigned long long>(blob, off[2 * nD]);
  put<int>(blob, TL.ld);
  std::vector<std::vector<char>> all;
  PHB_CHECK(ex.allgatherv(blob, all));
  P->lev.assign(nD, AmgPeerLevel());
  memset(P->lev.data(), 0, nD * sizeof(AmgPeerLevel));
  memset(&P->tail, 0, sizeof(AmgPeerLevel));
  for (int q = 0; q < NP; ++q) {
    const char *p = all[q].data();
    cudaIpcMemHandle_t hq;
    memcpy(&hq, p, sizeof(hq));
    p += sizeof(hq);
    if (q == me) P->peer[q] = P->block;
    else {
      void *m = nullptr;
      PHB_CUDA(cudaIpcOpenMemHandle(&m, hq, cudaIpcMemLazyEnablePeerAccess));
      P->peer[q] = (char *)m;
      P->opened[q] = true;
    }
    const int nq = take<int>(p);
    PHB_REQUIRE(nq == nD, "amg: rank %d holds %d distributed levels, rank %d holds %d", q, nq, me, nD);
    for (int l = 0; l < nD; ++l) {
      AmgPeerLevel &L = P->lev[l];
      L.vecOff[0][q] = take<unsigned long long>(p);
      L.vecOff[1][q] = take<unsigned long long>(p);
      L.ld[q] = take<int>(p);
      for (int r = 0; r < NP; ++r) {
        const int ro = take<int>(p);
        if (r == me) L.recvOff[q] = ro;      // where my values land in rank q's vector
      }
      L.sendOff[q] = D.lev[l]->sendOff[q]; L.sendCnt[q] = D.lev[l]->sendCnt[q]; L.recvCnt[q] = D.lev[l]->recvCnt[q];
    }
    P->tail.vecOff[0][q] = take<unsigned long long>(p);
    P->tail.ld[q] = take<int>(p);
    P->tail.recvOff[q] = D.tailOff[me];
    P->tail.sendOff[q] = D.tailOff[me];
    P->tail.sendCnt[q] = q == me ? 0 : D.tailOff[me + 1] - D.tailOff[me];
    P->tail.recvCnt[q] = q == me ? 0 : D.tailCnt[q];
  }
  // the level vectors now live inside the block (zero-filled: ghost and padding entries are finite)
  auto words = [](size_t bytes) { return (bytes + 7) / 8; };
  for (int l = 0; l < nD; ++l) {
    const size_t vb = (size_t)D.lev[l]->ld * s->nComp * sizeof(T);
    D.lev[l]->x.attach(reinterpret_cast<double *>(P->block + off[2 * l]), words(vb));
    D.lev[l]->x2.attach(reinterpret_cast<double *>(P->block + off[2 * l + 1]), words(vb));
  }
  TL.b.attach(reinterpret_cast<double *>(P->block + off[2 * nD]), words((size_t)TL.ld * s->nComp * sizeof(T)));
  D.peer = std::move(P);
  return PHB_OK;
}

// hierarchy spanning the ranks: level 0 = the solver's matrix with its ghost columns and the mesh's halo lists
template <typename T>
int rebuild_dist_t(phb_solver *s) {
  phb_ctx *c = s->ctx;
  AmgData &D = s->amg;
  const SellPattern *P = s->pat;
  const phb_mesh *m = s->halo;
  const int NP = c->nProcs, me = c->rank;
  std::vector<double> slotVals((size_t)P->nSlots);
  PHB_CUDA(cudaMemcpyAsync(slotVals.data(), s->dVals, slotVals.size() * sizeof(double), cudaMemcpyDeviceToHost,
                           c->stream));
  PHB_CUDA(cudaStreamSynchronize(c->stream));
  HCsr A0 = csr_from_sell(*P, slotVals, true);
  Halo h0;
  h0.sendPtr.assign(NP + 1, 0);
  h0.recvPtr.assign(NP + 1, 0);
  for (int q = 0; q < NP; ++q) {
    h0.sendPtr[q + 1] = h0.sendPtr[q] + m->hSendCnt[q];
    h0.recvPtr[q + 1] = h0.recvPtr[q] + m->hRecvCnt[q];
    PHB_REQUIRE(m->hSendOff[q] == h0.sendPtr[q] && (m->hRecvCnt[q] == 0 || m->hRecvOff[q] == P->nRows + h0.recvPtr[q]),
                "amg: unexpected halo layout of the mesh (peer %d)", q);
  }
  h0.sendIdx = m->hSendDev;
  std::vector<int> gid(A0.m);
  std::iota(gid.begin(), gid.end(), 0);
  NcclExchanger ex;
  ex.rank = me; ex.nProcs = NP; ex.c = c;
  g_hostRanks = NP;   // the ranks of one node share its cores
  DistHierarchy H;
  PHB_CHECK(build_dist_hierarchy(ex, std::move(A0), std::move(h0), std::move(gid), strength_of(D), D.coarsest, D.tailRows,
                                 4. / 3., H));
  D.lev.clear();
  D.peer.reset();
  D.nDist = (int)H.dist.size();
  for (int l = 0; l < D.nDist; ++l) {
    std::unique_ptr<AmgLevel> L(new AmgLevel());
    DistLevel &d = H.dist[l];
    PHB_CHECK(finish_level<T>(s, *L, d.L, d.n + d.g, true, l > 0));
    L->dist = true;
    L->sendOff.assign(NP, 0); L->sendCnt.assign(NP, 0); L->recvOff.assign(NP, 0); L->recvCnt.assign(NP, 0);
    for (int q = 0; q < NP; ++q) {
      L->sendOff[q] = d.halo.sendPtr[q]; L->sendCnt[q] = d.halo.sendPtr[q + 1] - d.halo.sendPtr[q];
      L->recvOff[q] = d.n + d.halo.recvPtr[q]; L->recvCnt[q] = d.halo.recvPtr[q + 1] - d.halo.recvPtr[q];
    }
    L->nSend = (int)d.halo.sendIdx.size();
    if (L->nSend) PHB_CHECK(L->sendIdx.upload(d.halo.sendIdx, c->stream));
    PHB_CHECK(alloc_as<T>(L->sendBuf, (size_t)std::max(1, L->nSend) * s->nComp, c->stream));
    PHB_CUDA(cudaStreamSynchronize(c->stream));
    D.lev.push_back(std::move(L));
  }
  HostHierarchy &TH = H.tail.H;
  const int nTail = (int)TH.lev.size();
  for (int l = 0; l < nTail; ++l) {
    std::unique_ptr<AmgLevel> L(new AmgLevel());
    PHB_CHECK(finish_level<T>(s, *L, TH.lev[l], TH.lev[l].A.n, l + 1 < nTail, true));
    D.lev.push_back(std::move(L));
  }
  D.tailOff = H.tailOff;
  D.tailCnt.assign(NP, 0); D.tailSendOff.assign(NP, 0); D.tailSendCnt.assign(NP, 0);
  for (int q = 0; q < NP; ++q) {
    D.tailCnt[q] = H.tailOff[q + 1] - H.tailOff[q];
    D.tailSendOff[q] = H.tailOff[me];
    D.tailSendCnt[q] = H.tailOff[me + 1] - H.tailOff[me];
  }
  D.nCoarse = TH.lev.back().A.n;
  D.denseCoarse = !TH.coarseInv.empty();
  if (D.denseCoarse) PHB_CHECK(D.coarseInv.upload(TH.coarseInv, c->stream));
  if (c->peer.enabled && NP <= kMaxPeers && 2 * D.nDist + 1 <= kAmgPeerChannels) PHB_CHECK(setup_peer<T>(s, ex));
  PHB_CHECK(D.refVals.alloc((size_t)P->nSlots));
  PHB_CUDA(cudaMemcpyAsync(D.refVals.p, s->dVals, (size_t)P->nSlots * sizeof(double), cudaMemcpyDeviceToDevice,
                           c->stream));
  if (sizeof(T) == 4) {
    PHB_CHECK(D.refValsF.alloc((size_t)P->nSlots));
    PHB_LAUNCH(c, k_amg_to_float, grid_rows(c, P->nSlots), kThreads, 0, P->nSlots, D.refVals.p, D.refValsF.p);
  }
  PHB_CUDA(cudaStreamSynchronize(c->stream));
  PHB_CHECK(build_tail_ops<T>(s));
  D.src = P;
  D.nComp = s->nComp;
  D.builtSingle = sizeof(T) == 4;
  D.built = true;
  D.setups++;
  D.setupMs = H.setupMs;
  {  // work per rank relative to its level-0 rows: own part of the distributed levels + the replicated tail
    double nnzAll = 0.;
    for (auto &d : H.dist) nnzAll += (double)d.L.A.nnz();
    for (auto &t : TH.lev) nnzAll += (double)t.A.nnz();
    D.opComplexity = nnzAll / std::max(1., (double)H.dist[0].L.A.nnz());
  }
  D.itersAfterSetup = -1;
  D.stale = false;
  if (s->graphExec) { cudaGraphExecDestroy(s->graphExec); s->graphExec = nullptr; }
  return PHB_OK;
}

template <typename T> const T *level0_vals(const AmgData &D);

// op list of the fused tail (levels fuseFrom .. L-1); called at the end of every (re)build
template <typename T>
int build_tail_ops(phb_solver *s) {
  AmgData &D = s->amg;
  phb_ctx *c = s->ctx;
  const int L = (int)D.lev.size();
  D.fuseFrom = -1;
  D.nTailOps = 0;
  if (D.fuseRows <= 0 || D.nu != 1 || !D.denseCoarse || L < 3) return PHB_OK;
  int F = -1;
  for (int l = std::max(1, D.nDist); l <= L - 2; ++l)
    if (D.lev[l]->n <= D.fuseRows && !D.lev[l]->dist) { F = l; break; }
  if (F < 0) return PHB_OK;
  const int nThreads = c->numSMs * 1024;
  auto lanes_for = [&](int rows) {
    int l = 1;
    while (l < 8 && (long long)rows * (2 * l) <= nThreads) l *= 2;
    return l;
  };
  std::vector<AmgTailOp> ops;
  auto sparse = [&](int kind, const AmgMat &M, int n, int ld, int ldIn, const void *w, const void *b, const void *x, void *y) {
    AmgTailOp o;
    o.kind = kind; o.n = n; o.ld = ld; o.ldIn = ldIn; o.lanes = lanes_for(n);
    o.sliceOff = M.pat.sliceOff.p; o.col = M.pat.col.p; o.vals = M.vals.p;
    o.w = w; o.b = b; o.x = x; o.y = y;
    ops.push_back(o);
  };
  for (int l = F; l <= L - 2; ++l) {
    AmgLevel &V = *D.lev[l], &C = *D.lev[l + 1];
    sparse(0, V.A, V.n, V.ld, V.ld, V.w.p, V.b.p, nullptr, V.r.p);
    sparse(1, V.R, C.n, C.ld, V.ld, nullptr, nullptr, V.r.p, C.b.p);
  }
  {
    AmgLevel &V = *D.lev[L - 1];
    AmgTailOp o;
    o.kind = 2; o.n = V.n; o.ld = V.ld; o.ldIn = V.ld; o.lanes = 32;
    o.sliceOff = nullptr; o.col = nullptr; o.vals = D.coarseInv.p; o.w = nullptr; o.b = V.b.p; o.x = nullptr; o.y = V.x.p;
    ops.push_back(o);
  }
  for (int l = L - 2; l >= F; --l) {
    AmgLevel &V = *D.lev[l], &C = *D.lev[l + 1];
    const void *xc = l + 1 == L - 1 ? C.x.p : C.x2.p;
    sparse(3, V.P, V.n, V.ld, C.ld, V.w.p, V.b.p, xc, V.x.p);
    sparse(4, V.A, V.n, V.ld, V.ld, V.w.p, V.b.p, V.x.p, V.x2.p);
  }
  if ((int)ops.size() > kMaxTailOps) return PHB_OK;   // deeper than the kernel's shared op list: separate launches
  PHB_CHECK(D.tailOps.upload(ops, c->stream));
  PHB_CHECK(D.tailBar.alloc(4));
  PHB_CHECK(D.tailBar.zero(c->stream));
  PHB_CUDA(cudaStreamSynchronize(c->stream));
  D.fuseFrom = F;
  D.nTailOps = (int)ops.size();
  return PHB_OK;
}
template <> const float *level0_vals<float>(const AmgData &D) { return D.refValsF.p; }
template <> const double *level0_vals<double>(const AmgData &D) { return D.refVals.p; }

template <typename T> struct Cycle {
  phb_solver *s;
  AmgData &D;
  bool inLoop;
  const KrylovSums *S;
  int nc;
  bool failed = false;
  const SellPattern &pat(int l) const { return l == 0 ? *s->pat : D.lev[l]->A.pat; }
  const T *val(int l) const { return l == 0 ? level0_vals<T>(D) : as<T>(D.lev[l]->A.vals); }

  // ghost refresh of a distributed level's vector (no-op on replicated / single-rank levels)
  int halo(int l, T *x) {
    AmgLevel &V = *D.lev[l];
    if (!V.dist) return PHB_OK;
    phb_ctx *c = s->ctx;
    if (D.peer) {
      const int vec = x == as<T>(V.x) ? 0 : 1;
      PHB_LAUNCH(c, (k_amg_peer_halo<T>), 1, 1024, 0, peer_view(), 2 * l + vec, D.peer->lev[l], vec, (const T *)x, nc, V.ld,
                 (const int *)V.sendIdx.p, S, s->maxIters);
      return PHB_OK;
    }
    T *buf = as<T>(V.sendBuf);
    if (V.nSend)
      PHB_LAUNCH(c, (k_amg_pack<T>), (V.nSend + 255) / 256, 256, 0, V.nSend, nc, V.ld, V.sendIdx.p, (const T *)x, buf, S,
                 s->maxIters);
    for (int k = 0; k < nc; ++k)
      PHB_CHECK(comm_exchange_bytes(c, buf + (size_t)k * V.nSend, V.sendOff.data(), V.sendCnt.data(),
                                    x + (size_t)k * V.ld, V.recvOff.data(), V.recvCnt.data(), sizeof(T)));
    return PHB_OK;
  }
  // first replicated level: every rank contributes its segment of the right-hand side
  AmgPeerView peer_view() const {
    AmgPeerView v;
    v.block = D.peer->block;
    for (int q = 0; q < kMaxPeers; ++q) v.peer[q] = D.peer->peer[q];
    v.rank = s->ctx->rank;
    v.nProcs = s->ctx->nProcs;
    return v;
  }
  int gather_tail(T *b) {
    AmgLevel &V = *D.lev[D.nDist];
    if (D.peer) {
      PHB_LAUNCH(s->ctx, (k_amg_peer_halo<T>), 1, 1024, 0, peer_view(), 2 * D.nDist, D.peer->tail, 0, (const T *)b, nc, V.ld,
                 (const int *)nullptr, S, s->maxIters);
      return PHB_OK;
    }
    for (int k = 0; k < nc; ++k)
      PHB_CHECK(comm_exchange_bytes(s->ctx, b + (size_t)k * V.ld, D.tailSendOff.data(), D.tailSendCnt.data(),
                                    b + (size_t)k * V.ld, D.tailOff.data(), D.tailCnt.data(), sizeof(T)));
    return PHB_OK;
  }
  int mySeg(int l) const { return (D.nDist > 0 && l == D.nDist) ? D.tailOff[s->ctx->rank] : 0; }

  // pre-smoothing from a zero guess, residual, restriction; returns the level's iterate
  template <typename TB> T *down(int l, const TB *b) {
    AmgLevel &V = *D.lev[l];
    const int ld = V.ld;
    T *x = as<T>(V.x), *x2 = as<T>(V.x2);
    PHB_LAUNCH(s->ctx, (k_amg_scale<T, TB>), grid_rows(s->ctx, V.n), kThreads, 0, V.n, nc, ld, as<T>(V.w), b, x, S,
               s->maxIters);
    for (int k = 1; k < D.nu; ++k) {
      if (halo(l, x) != PHB_OK) { failed = true; return x; }
      launch<2>(s, pat(l), val(l), (const T *)x, ld, x2, ld, b, (const T *)as<T>(V.w), inLoop);
      std::swap(x, x2);
    }
    if (halo(l, x) != PHB_OK) { failed = true; return x; }
    launch<1>(s, pat(l), val(l), (const T *)x, ld, as<T>(V.r), ld, b, (const T *)nullptr, inLoop);
    AmgLevel &C = *D.lev[l + 1];
    launch<0>(s, V.R.pat, (const T *)as<T>(V.R.vals), (const T *)as<T>(V.r), ld, as<T>(C.b) + mySeg(l + 1), C.ld,
              (const T *)nullptr, (const T *)nullptr, inLoop);
    if (D.nDist > 0 && l + 1 == D.nDist && gather_tail(as<T>(C.b)) != PHB_OK) failed = true;
    return x;
  }
  // coarse correction + post-smoothing; the last sweep of level 0 writes the fp64 result
  template <typename TB> T *up(int l, const TB *b, T *x, const T *xc, double *out) {
    AmgLevel &V = *D.lev[l];
    const int ld = V.ld;
    T *x2 = x == as<T>(V.x) ? as<T>(V.x2) : as<T>(V.x);
    // the smoothed prolongator reaches into the neighbours' aggregates: refresh the ghosts of x_c first (on the
    // last distributed level its columns index the gathered coarse vector, which every rank holds in full)
    if (halo(l + 1, const_cast<T *>(xc)) != PHB_OK) { failed = true; return x; }
    launch<3>(s, V.P.pat, (const T *)as<T>(V.P.vals), xc, D.lev[l + 1]->ld, x, ld, (const T *)nullptr,
              (const T *)nullptr, inLoop);
    for (int k = 0; k < D.nu; ++k) {
      if (halo(l, x) != PHB_OK) { failed = true; return x; }
      if (out && k == D.nu - 1) {
        launch<2>(s, pat(l), val(l), (const T *)x, ld, out, ld, b, (const T *)as<T>(V.w), inLoop);
        return nullptr;
      }
      launch<2>(s, pat(l), val(l), (const T *)x, ld, x2, ld, b, (const T *)as<T>(V.w), inLoop);
      std::swap(x, x2);
    }
    return x;
  }
  // coarsest level: dense inverse, or a fixed number of Jacobi sweeps when it is too large for one
  template <typename TB> T *coarse(int l, const TB *b, double *out) {
    AmgLevel &V = *D.lev[l];
    const int ld = V.ld;
    T *x = as<T>(V.x), *x2 = as<T>(V.x2);
    if (D.denseCoarse) {
      if (out)
        PHB_LAUNCH(s->ctx, (k_amg_dense<TB, double>), (V.n + 7) / 8, 256, 0, V.n, nc, ld, D.coarseInv.p, b, out, S,
                   s->maxIters);
      else
        PHB_LAUNCH(s->ctx, (k_amg_dense<TB, T>), (V.n + 7) / 8, 256, 0, V.n, nc, ld, D.coarseInv.p, b, x, S,
                   s->maxIters);
      return x;
    }
    PHB_LAUNCH(s->ctx, (k_amg_scale<T, TB>), grid_rows(s->ctx, V.n), kThreads, 0, V.n, nc, ld, as<T>(V.w), b, x, S,
               s->maxIters);
    for (int k = 0; k < kCoarseSweeps; ++k) {
      if (out && k == kCoarseSweeps - 1) {
        launch<2>(s, pat(l), val(l), (const T *)x, ld, out, ld, b, (const T *)as<T>(V.w), inLoop);
        return nullptr;
      }
      launch<2>(s, pat(l), val(l), (const T *)x, ld, x2, ld, b, (const T *)as<T>(V.w), inLoop);
      std::swap(x, x2);
    }
    return x;
  }

  int run(const double *in, double *out) {
    const int L = (int)D.lev.size();
    if (L == 1) { coarse(0, in, out); return PHB_OK; }
    std::vector<T *> xOf(L);
    xOf[0] = down(0, in);
    const int F = D.fuseFrom;
    if (F >= 1 && D.nu == 1) {   // levels F .. L-1 in one launch (k_amg_tail); its result is level F's post-smoothed iterate
      for (int l = 1; l < F; ++l) xOf[l] = down(l, (const T *)as<T>(D.lev[l]->b));
      if (nc == 1)
        PHB_LAUNCH(s->ctx, (k_amg_tail<T, 1>), s->ctx->numSMs, 1024, 0, (const AmgTailOp *)D.tailOps.p, D.nTailOps, D.tailBar.p, S,
                   s->maxIters);
      else
        PHB_LAUNCH(s->ctx, (k_amg_tail<T, 2>), s->ctx->numSMs, 1024, 0, (const AmgTailOp *)D.tailOps.p, D.nTailOps, D.tailBar.p, S,
                   s->maxIters);
      xOf[F] = as<T>(D.lev[F]->x2);
      for (int l = F - 1; l >= 1; --l)
        xOf[l] = up(l, (const T *)as<T>(D.lev[l]->b), xOf[l], (const T *)xOf[l + 1], nullptr);
      up(0, in, xOf[0], (const T *)xOf[1], out);
      return failed ? PHB_ERR_COMM : PHB_OK;
    }
    for (int l = 1; l + 1 < L; ++l) xOf[l] = down(l, (const T *)as<T>(D.lev[l]->b));
    xOf[L - 1] = coarse(L - 1, (const T *)as<T>(D.lev[L - 1]->b), nullptr);
    for (int l = L - 2; l >= 1; --l)
      xOf[l] = up(l, (const T *)as<T>(D.lev[l]->b), xOf[l], (const T *)xOf[l + 1], nullptr);
    up(0, in, xOf[0], (const T *)xOf[1], out);
    return failed ? PHB_ERR_COMM : PHB_OK;
  }
};

}  // namespace

namespace phb {

// Called once per solve, before the Krylov loop: (re)build the hierarchy when there is none, the pattern
// changed, or the matrix drifted from the one it was built from AND the last solve needed markedly more
// iterations than the first solve after the setup did.
int amg_prepare(phb_solver *s) {
  phb_ctx *c = s->ctx;
  AmgData &D = s->amg;
  bool need = !D.built || D.src != s->pat || D.refVals.n != (size_t)s->pat->nSlots || D.nComp != s->nComp ||
              D.builtSingle != D.single;
  // the caller vouches that this matrix is the one the hierarchy's values were computed from (same tag on every rank):
  // nothing to compare, nothing to agree on
  if (!need && s->valsTag != 0ull && s->valsTag == D.builtTag) return PHB_OK;
  if (!need) {
    PHB_CHECK(D.chk.alloc(2));
    int first = 0;
    PHB_LAUNCH(c, k_amg_changed, grid_rows(c, s->pat->nSlots), kThreads, 0, s->pat->nSlots, s->dVals, D.refVals.p,
               first, s->partials.p, s->ticket.p, D.chk.p);
    PHB_CUDA(cudaMemcpyAsync(c->pinned, D.chk.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PHB_CUDA(cudaStreamSynchronize(c->stream));
    const double dev = c->pinned[0];
    D.stale = !(dev <= 1e-9);
    if (D.stale) {
      // coefficients changed on the same pattern.  Cheap answer: recompute the values of every level on the device with
      // the aggregates and patterns as they are (once the iteration count has drifted 20 % from the count after the
      // setup, or every time with `amgRefresh always`); a refresh that does not bring the count back means the
      // aggregates no longer fit the coefficients, and the host setup runs again.
      const int base = D.itersAfterSetup;
      const bool doubled = base >= 0 && s->lastIters > std::max(2 * base, base + 10);
      if (D.refresh && D.refreshMode != 0 && c->nProcs == 1) {
        const bool inVain = D.refreshes > 0 && base >= 0 && D.itersAfterRefresh > std::max(2 * base, base + 10);
        if (inVain) {
          need = true;
        } else if (D.refreshMode == 2 || (base >= 0 && s->lastIters > std::max((6 * base + 4) / 5, base + 2))) {
          const int rc = D.single ? refresh_numeric<float>(s) : refresh_numeric<double>(s);
          if (rc < 0) return rc;
          if (rc > 0) need = true;
        }
      } else if (doubled) {
        need = true;
      }
    }
    if (D.rebuildAlways && D.stale) need = true;
  }
  const bool dist = c->nProcs > 1 && s->halo && D.global;
  bool staleAnywhere = D.stale;
  if (c->nProcs > 1) {  // the setup talks to the other ranks: everybody rebuilds or nobody does
    // ... and everybody holds the same opinion on whether the hierarchy still fits (the coefficient tags below must
    // agree across the ranks, or the next solve would leave some of them alone in this exchange)
    PHB_CHECK(D.chk.alloc(2));
    const double flag[2] = {need ? 1. : 0., D.stale ? 1. : 0.};
    PHB_CUDA(cudaMemcpyAsync(D.chk.p, flag, sizeof(flag), cudaMemcpyHostToDevice, c->stream));
    PHB_CHECK(comm_allreduce_max(c, D.chk.p, 2));
    PHB_CUDA(cudaMemcpyAsync(c->pinned, D.chk.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PHB_CUDA(cudaStreamSynchronize(c->stream));
    need = c->pinned[0] > 0.5;
    staleAnywhere = c->pinned[1] > 0.5;
  }
  if (need) {
    if (dist) PHB_CHECK(D.single ? rebuild_dist_t<float>(s) : rebuild_dist_t<double>(s));
    else PHB_CHECK(D.single ? rebuild_t<float>(s) : rebuild_t<double>(s));
  }
  // values of the hierarchy == current coefficients (fresh setup, refresh, or the comparison found no drift)?
  D.builtTag = (need || !staleAnywhere) ? s->valsTag : 0ull;
  return PHB_OK;
}

void amg_record_iters(phb_solver *s, int iters) {
  if (s->amg.built && s->amg.itersAfterSetup < 0) s->amg.itersAfterSetup = iters;
  if (s->amg.built && s->amg.refreshes > 0 && s->amg.itersAfterRefresh == -1) s->amg.itersAfterRefresh = iters;
}

// `amgRefresh auto`, inside a solve: the hierarchy belongs to other coefficients and this solve has already used twice
// the iterations the hierarchy needed when it was new
bool amg_wants_refresh(const phb_solver *s, int itersSoFar) {
  const AmgData &D = s->amg;
  if (!D.stale || !D.refresh || D.refreshMode != 1 || s->ctx->nProcs != 1 || D.itersAfterSetup < 0) return false;
  return itersSoFar >= std::max(2 * D.itersAfterSetup, D.itersAfterSetup + 10);
}
int amg_refresh_midsolve(phb_solver *s) {
  AmgData &D = s->amg;
  const int rc = D.single ? refresh_numeric<float>(s) : refresh_numeric<double>(s);
  if (rc < 0) return rc;
  D.stale = false;               // rc > 0 (vanishing pivot): keep the old values, do not try again in this solve
  D.itersAfterRefresh = -2;      // the count of this solve includes the iterations spent before the refresh: not a measure
  return PHB_OK;
}

// PHB_ERR_STATE when a grid barrier of the fused tail gave up (co-residency of the CTAs lost: should not happen
// with one CTA per SM on a stream of our own)
int amg_check(phb_solver *s) {
  AmgData &D = s->amg;
  if (D.fuseFrom < 1 || !D.tailBar.p) return PHB_OK;
  unsigned flag = 0;
  PHB_CUDA(cudaMemcpyAsync(&flag, D.tailBar.p + 2, sizeof(unsigned), cudaMemcpyDeviceToHost, s->ctx->stream));
  PHB_CUDA(cudaStreamSynchronize(s->ctx->stream));
  if (flag) {
    set_error("amg: a grid barrier of the fused small-level kernel timed out");
    return PHB_ERR_STATE;
  }
  return PHB_OK;
}

int refresh_numeric_f(phb_solver *s) { return refresh_numeric<float>(s); }
int refresh_numeric_d(phb_solver *s) { return refresh_numeric<double>(s); }

int amg_launches_per_apply(const phb_solver *s) {
  const AmgData &D = s->amg;
  const int L = (int)D.lev.size();
  if (L == 0) return 0;
  const int perLevel = 1 + (D.nu - 1) + 2 + 1 + D.nu;  // scale, extra pre, residual + restrict, prolong, post
  int packs = 0;                                       // one pack kernel per ghost refresh of a distributed level
  for (int l = 0; l < D.nDist; ++l)
    if (D.lev[l]->nSend) packs += 2 * D.nu + (l > 0 ? 1 : 0);
  if (D.fuseFrom >= 1 && D.nu == 1) return D.fuseFrom * perLevel + 1 + packs;
  return (L - 1) * perLevel + (D.denseCoarse ? 1 : 1 + kCoarseSweeps) + packs;
}

// algorithmic bytes of one cycle: every matrix streamed once per use (index + value per entry, 4 B per slice
// offset -- the sliced-ELL kernels read no row pointers), every vector read or written ONCE per kernel that
// touches it (the gathered x and the row's own x are the same array); v = bytes of a cycle scalar (4 single,
// 8 double), the Krylov vectors touched on level 0 (b, result) are fp64
double amg_cycle_bytes(const phb_solver *s) {
  const AmgData &D = s->amg;
  const int L = (int)D.lev.size();
  if (L == 0) return 0.;
  const double v = D.builtSingle ? 4. : 8.;
  auto mat = [v](const SellPattern &P) { return (4. + v) * (double)P.nnz + 4. * (P.nSlices + 1.); };
  double total = 0.;
  for (int l = 0; l < L; ++l) {
    const AmgLevel &V = *D.lev[l];
    const double k = s->nComp, n = V.n, a = mat(l == 0 ? *s->pat : V.A.pat);
    const double vb = l == 0 ? 8. : v;                            // right-hand side of this level
    const double jac = a + (v + (2. * v + vb) * k) * n;           // w | x, y, b per component
    const double res = a + (2. * v + vb) * k * n;                 // x, r, b
    if (l + 1 < L) {
      const double nc = D.lev[l + 1]->n;
      total += (v + (v + vb) * k) * n + (D.nu - 1) * jac + res + (mat(V.R.pat) + v * k * (n + nc)) +
               (mat(V.P.pat) + v * k * nc + 2. * v * k * n) + D.nu * jac + (l == 0 ? (8. - v) * k * n : 0.);
    } else {
      total += D.denseCoarse ? 8. * n * n + 2. * v * k * n : (v + (v + vb) * k) * n + kCoarseSweeps * jac;
    }
  }
  return total;
}

// z = M^-1 r : one V(nu, nu) cycle.  All kernels test the device-side convergence flag first, so the
// tail of a graph after convergence costs launches only.
int amg_apply(phb_solver *s, const double *in, double *out, bool inLoop) {
  AmgData &D = s->amg;
  const KrylovSums *S = inLoop ? s->sums.p : nullptr;
  if (D.builtSingle) {
    Cycle<float> cy{s, D, inLoop, S, s->nComp};
    return cy.run(in, out);
  }
  Cycle<double> cy{s, D, inLoop, S, s->nComp};
  return cy.run(in, out);
}

// Live timing of the cycle's dominant launches (bench.py roofline leg): CUDA events on the context stream around
// `reps` launches of each level-0 kernel and of the whole cycle, on the solver's own vectors.
template <typename T>
int amg_time_t(phb_solver *s, int reps, double out[8]) {
  phb_ctx *c = s->ctx;
  AmgData &D = s->amg;
  AmgLevel &V = *D.lev[0], &C = *D.lev[1];
  const int ld = V.ld;
  const SellPattern &pat0 = *s->pat;
  const T *val0 = level0_vals<T>(D), *w = as<T>(V.w);
  T *x = as<T>(V.x), *r = as<T>(V.r);
  const double *in = s->p.p;
  double *res = s->ph.p;
  cudaEvent_t e0, e1;
  PHB_CUDA(cudaEventCreate(&e0));
  PHB_CUDA(cudaEventCreate(&e1));
  Cycle<T> cy{s, D, false, nullptr, s->nComp};
  auto timed = [&](int which, double *ms) -> int {
    for (int k = -2; k < reps; ++k) {
      if (k == 0) PHB_CUDA(cudaEventRecord(e0, c->stream));
      if (which == 0) launch<1>(s, pat0, val0, (const T *)x, ld, r, ld, in, (const T *)nullptr, false);
      if (which == 1) launch<0>(s, V.R.pat, (const T *)as<T>(V.R.vals), (const T *)r, ld, as<T>(C.b) + cy.mySeg(1), C.ld,
                                (const T *)nullptr, (const T *)nullptr, false);
      if (which == 2) launch<3>(s, V.P.pat, (const T *)as<T>(V.P.vals), (const T *)as<T>(C.x), C.ld, x, ld, (const T *)nullptr,
                                (const T *)nullptr, false);
      if (which == 3) launch<2>(s, pat0, val0, (const T *)x, ld, res, ld, in, w, false);
      if (which == 4) PHB_CHECK(cy.run(in, res));
    }
    PHB_CUDA(cudaEventRecord(e1, c->stream));
    PHB_CUDA(cudaEventSynchronize(e1));
    float t = 0.f;
    PHB_CUDA(cudaEventElapsedTime(&t, e0, e1));
    *ms = (double)t / reps;
    return PHB_OK;
  };
  PHB_CUDA(cudaMemsetAsync(x, 0, (size_t)ld * s->nComp * sizeof(T), c->stream));
  for (int k = 0; k < 5; ++k) PHB_CHECK(timed(k, &out[k]));
  PHB_CUDA(cudaMemsetAsync(x, 0, (size_t)ld * s->nComp * sizeof(T), c->stream));
  PHB_CUDA(cudaStreamSynchronize(c->stream));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double v = sizeof(T), n = V.n, k = s->nComp;
  // Jacobi sweep of level 0: matrix (index + value per entry, slice offsets), w, then per component x (cycle
  // precision, read once), b read and result written (both fp64: Krylov vectors)
  out[5] = (4. + v) * (double)pat0.nnz + 4. * (pat0.nSlices + 1.) + v * n + (v + 16.) * k * n;
  out[6] = amg_cycle_bytes(s);
  out[7] = amg_launches_per_apply(s);
  return PHB_OK;
}

int amg_time(phb_solver *s, int reps, double out[8]) {
  AmgData &D = s->amg;
  PHB_REQUIRE(D.built && D.lev.size() >= 2 && s->pat && s->ph.p && s->p.p,
              "phb_solver_time_amg: no multigrid hierarchy with at least two levels has been used yet");
  PHB_REQUIRE(reps > 0, "phb_solver_time_amg: reps must be positive");
  return D.builtSingle ? amg_time_t<float>(s, reps, out) : amg_time_t<double>(s, reps, out);
}

}  // namespace phb

// ===================================================================== C ABI (inspection / tests)
struct phb_amg_dist {
  int nRanks = 0;
  std::vector<DistHierarchy> H;   // one per (virtual) rank
};

extern "C" {

// Distributed setup with the ranks as threads of this process: `part[i]` = owner of global row i.
// Local numbering as on the device: owned rows in ascending global order, ghosts grouped by owner.
int phb_amg_dist_build(int nRanks, int n, const int *rowPtr, const int *colInd, const double *vals, const int *part,
                       double theta, int coarsest, long long tailRows, phb_amg_dist **out) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(nRanks >= 1 && nRanks <= 64 && n > 0 && rowPtr && colInd && vals && part && out,
              "phb_amg_dist_build: bad argument");
  std::vector<std::vector<int>> owned(nRanks);
  std::vector<int> local(n);
  for (int i = 0; i < n; ++i) {
    PHB_REQUIRE(part[i] >= 0 && part[i] < nRanks, "phb_amg_dist_build: part[%d] out of range", i);
    local[i] = (int)owned[part[i]].size();
    owned[part[i]].push_back(i);
  }
  // ghosts of every rank: columns of its rows owned elsewhere, grouped by owner, ascending global id
  std::vector<std::vector<std::vector<int>>> ghosts(nRanks, std::vector<std::vector<int>>(nRanks));
  for (int r = 0; r < nRanks; ++r) {
    for (int i : owned[r])
      for (int k = rowPtr[i]; k < rowPtr[i + 1]; ++k)
        if (colInd[k] >= 0 && part[colInd[k]] != r) ghosts[r][part[colInd[k]]].push_back(colInd[k]);
    for (auto &g : ghosts[r]) {
      std::sort(g.begin(), g.end());
      g.erase(std::unique(g.begin(), g.end()), g.end());
    }
  }
  std::vector<HCsr> A(nRanks);
  std::vector<Halo> halo(nRanks);
  std::vector<std::vector<int>> gid(nRanks);
  for (int r = 0; r < nRanks; ++r) {
    const int nl = (int)owned[r].size();
    Halo &h = halo[r];
    h.recvPtr.assign(nRanks + 1, 0);
    h.sendPtr.assign(nRanks + 1, 0);
    gid[r] = owned[r];
    for (int q = 0; q < nRanks; ++q) {
      h.recvPtr[q + 1] = h.recvPtr[q] + (int)ghosts[r][q].size();
      for (int g : ghosts[r][q]) gid[r].push_back(g);
      for (int g : ghosts[q][r]) h.sendIdx.push_back(local[g]);
      h.sendPtr[q + 1] = (int)h.sendIdx.size();
    }
    HCsr &M = A[r];
    M.n = nl; M.m = nl + h.recvPtr[nRanks];
    M.rp.assign(1, 0);
    std::vector<std::pair<int, double>> row;
    for (int i : owned[r]) {
      row.clear();
      for (int k = rowPtr[i]; k < rowPtr[i + 1]; ++k) {
        const int c = colInd[k];
        if (c < 0) continue;
        int lc;
        if (part[c] == r) lc = local[c];
        else {
          const auto &g = ghosts[r][part[c]];
          lc = nl + h.recvPtr[part[c]] + (int)(std::lower_bound(g.begin(), g.end(), c) - g.begin());
        }
        row.push_back({lc, vals[k]});
      }
      std::sort(row.begin(), row.end(), [](const std::pair<int, double> &x, const std::pair<int, double> &y) {
        return x.first < y.first;
      });
      for (auto &e : row) { M.ci.push_back(e.first); M.v.push_back(e.second); }
      M.rp.push_back((int)M.ci.size());
    }
  }
  std::unique_ptr<phb_amg_dist> h(new phb_amg_dist());
  h->nRanks = nRanks;
  h->H.resize(nRanks);
  ThreadBoard board;
  board.slots.resize(nRanks);
  std::vector<int> rc(nRanks, PHB_OK);
  std::vector<std::string> err(nRanks);
  std::vector<std::thread> th;
  for (int r = 0; r < nRanks; ++r)
    th.emplace_back([&, r] {
      ThreadExchanger ex;
      ex.rank = r; ex.nProcs = nRanks; ex.B = &board;
      rc[r] = build_dist_hierarchy(ex, std::move(A[r]), std::move(halo[r]), std::move(gid[r]), theta,
                                   coarsest > 0 ? coarsest : 1000, tailRows, 4. / 3., h->H[r]);
      if (rc[r] != PHB_OK) err[r] = phb_last_error();
    });
  for (auto &t : th) t.join();
  for (int r = 0; r < nRanks; ++r)
    if (rc[r] != PHB_OK) { set_error("rank %d: %s", r, err[r].c_str()); return rc[r]; }
  *out = h.release();
  return PHB_OK;
  PHB_TRY_END
}

int phb_amg_dist_info(const phb_amg_dist *h, int *nDistLevels, int *nTailLevels, int *singular) {
  PHB_REQUIRE(h && nDistLevels && nTailLevels, "phb_amg_dist_info: NULL argument");
  *nDistLevels = (int)h->H[0].dist.size();
  *nTailLevels = (int)h->H[0].tail.H.lev.size();
  if (singular) *singular = h->H[0].singular ? 1 : 0;
  return PHB_OK;
}

// the replicated part of rank `rank` as a serial hierarchy (borrowed: do not destroy)
const phb_amg_host *phb_amg_dist_tail(const phb_amg_dist *h, int rank) {
  if (!h || rank < 0 || rank >= h->nRanks) return nullptr;
  return &h->H[rank].tail;
}

// which: 0 = A_l (rows = the rank's cells of level l, global column ids of level l),
//        1 = P_l (same rows, global column ids of level l + 1),
//        2 = R_l (rows = the rank's cells of level l + 1, global column ids of level l)
// which + 10: the same matrices with the rank's LOCAL column numbering (owned, then ghosts), i.e. exactly the
// data the device cycle works on
static const HCsr &dist_pick(const DistLevel &D, int which) {
  which %= 10;
  return which == 0 ? D.L.A : which == 1 ? D.L.P : D.L.R;
}

int phb_amg_dist_matrix_size(const phb_amg_dist *h, int rank, int level, int which, int *nRows, long long *nnz) {
  PHB_REQUIRE(h && nRows && nnz && rank >= 0 && rank < h->nRanks && level >= 0 &&
              level < (int)h->H[rank].dist.size() && which >= 0 && which % 10 <= 2 && which < 20,
              "phb_amg_dist_matrix_size: bad argument");
  const HCsr &M = dist_pick(h->H[rank].dist[level], which);
  *nRows = M.n; *nnz = M.nnz();
  return PHB_OK;
}

int phb_amg_dist_matrix(const phb_amg_dist *h, int rank, int level, int which, int *rowPtr, int *colGid,
                        double *vals, int *rowGid) {
  PHB_REQUIRE(h && rowPtr && colGid && vals && rowGid && rank >= 0 && rank < h->nRanks && level >= 0 &&
              level < (int)h->H[rank].dist.size() && which >= 0 && which % 10 <= 2 && which < 20,
              "phb_amg_dist_matrix: bad argument");
  const DistHierarchy &H = h->H[rank];
  const DistLevel &D = H.dist[level];
  const HCsr &M = dist_pick(D, which);
  const bool last = level + 1 == (int)H.dist.size();
  std::copy(M.rp.begin(), M.rp.end(), rowPtr);
  std::copy(M.v.begin(), M.v.end(), vals);
  if (which >= 10) {
    std::copy(M.ci.begin(), M.ci.end(), colGid);
    for (int i = 0; i < M.n; ++i) rowGid[i] = i;
    return PHB_OK;
  }
  auto coarseGid = [&](int c) { return last ? H.tailOff[rank] + c : H.dist[level + 1].gid[c]; };
  if (which == 2) {
    for (int i = 0; i < M.n; ++i) rowGid[i] = coarseGid(i);
    for (long long k = 0; k < M.nnz(); ++k) colGid[k] = D.gid[M.ci[k]];
    return PHB_OK;
  }
  for (int i = 0; i < D.n; ++i) rowGid[i] = D.gid[i];
  if (which == 0) {
    for (long long k = 0; k < M.nnz(); ++k) colGid[k] = D.gid[M.ci[k]];
  } else if (last) {   // columns already index the gathered coarse vector
    for (long long k = 0; k < M.nnz(); ++k) colGid[k] = M.ci[k];
  } else {
    for (long long k = 0; k < M.nnz(); ++k) colGid[k] = H.dist[level + 1].gid[M.ci[k]];
  }
  return PHB_OK;
}

// halo lists of a distributed level: sendPtr/recvPtr hold nRanks + 1 entries
int phb_amg_dist_halo(const phb_amg_dist *h, int rank, int level, int *sendPtr, int *sendIdx, int *recvPtr) {
  PHB_REQUIRE(h && sendPtr && sendIdx && recvPtr && rank >= 0 && rank < h->nRanks && level >= 0 &&
              level < (int)h->H[rank].dist.size(), "phb_amg_dist_halo: bad argument");
  const Halo &a = h->H[rank].dist[level].halo;
  std::copy(a.sendPtr.begin(), a.sendPtr.end(), sendPtr);
  std::copy(a.sendIdx.begin(), a.sendIdx.end(), sendIdx);
  std::copy(a.recvPtr.begin(), a.recvPtr.end(), recvPtr);
  return PHB_OK;
}

// rank segments of the first replicated level: nRanks + 1 offsets
int phb_amg_dist_tail_offsets(const phb_amg_dist *h, int *out) {
  PHB_REQUIRE(h && out, "phb_amg_dist_tail_offsets: NULL argument");
  std::copy(h->H[0].tailOff.begin(), h->H[0].tailOff.end(), out);
  return PHB_OK;
}

// global ids (within the level) of the rank's ghost columns, in ghost order: recvPtr[nRanks] entries
int phb_amg_dist_ghost_gids(const phb_amg_dist *h, int rank, int level, int *out) {
  PHB_REQUIRE(h && out && rank >= 0 && rank < h->nRanks && level >= 0 && level < (int)h->H[rank].dist.size(),
              "phb_amg_dist_ghost_gids: bad argument");
  const DistLevel &D = h->H[rank].dist[level];
  std::copy(D.gid.begin() + D.n, D.gid.end(), out);
  return PHB_OK;
}

int phb_amg_dist_destroy(phb_amg_dist *h) {
  delete h;
  return PHB_OK;
}


int phb_amg_host_build(int n, const int *rowPtr, const int *colInd, const double *vals, double theta,
                       int coarsest, phb_amg_host **out) {
  return phb_amg_host_build_ex(n, rowPtr, colInd, vals, theta, -1., -1., coarsest, out);
}

int phb_amg_host_build_ex(int n, const int *rowPtr, const int *colInd, const double *vals, double theta,
                          double aggTheta, double coarseWeight, int coarsest, phb_amg_host **out) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(n > 0 && rowPtr && colInd && vals && out, "phb_amg_host_build: bad argument");
  Strength rule(theta);
  if (aggTheta >= 0.) rule.agg = aggTheta;
  if (coarseWeight >= 0.) rule.omegaC = coarseWeight;
  HCsr A;
  A.n = A.m = n;
  A.rp.assign(n + 1, 0);
  std::vector<std::pair<int, double>> row;
  for (int r = 0; r < n; ++r) {
    row.clear();
    for (int k = rowPtr[r]; k < rowPtr[r + 1]; ++k) {
      if (colInd[k] < 0) continue;
      PHB_REQUIRE(colInd[k] < n, "phb_amg_host_build: column %d out of range", colInd[k]);
      row.push_back({colInd[k], vals[k]});
    }
    std::sort(row.begin(), row.end(), [](const std::pair<int, double> &x, const std::pair<int, double> &y) {
      return x.first < y.first;
    });
    for (auto &e : row) { A.ci.push_back(e.first); A.v.push_back(e.second); }
    A.rp[r + 1] = (int)A.ci.size();
  }
  std::unique_ptr<phb_amg_host> h(new phb_amg_host());
  PHB_CHECK(build_hierarchy(std::move(A), rule, coarsest > 0 ? coarsest : 1000, 4. / 3., h->H));
  *out = h.release();
  return PHB_OK;
  PHB_TRY_END
}

int phb_amg_host_levels(const phb_amg_host *h, int *nLevels, int *singular, int *denseCoarse) {
  PHB_REQUIRE(h && nLevels, "phb_amg_host_levels: NULL argument");
  *nLevels = (int)h->H.lev.size();
  if (singular) *singular = h->H.singular ? 1 : 0;
  if (denseCoarse) *denseCoarse = h->H.coarseInv.empty() ? 0 : 1;
  return PHB_OK;
}

static const HCsr *pick(const phb_amg_host *h, int level, int which) {
  if (level < 0 || level >= (int)h->H.lev.size()) return nullptr;
  const HostLevel &L = h->H.lev[level];
  const HCsr *M = which == 0 ? &L.A : which == 1 ? &L.P : which == 2 ? &L.R : nullptr;
  if (M && which != 0 && level + 1 == (int)h->H.lev.size()) return nullptr;
  return M;
}

int phb_amg_host_level_size(const phb_amg_host *h, int level, int which, int *nRows, int *nCols, long long *nnz,
                            double *rho) {
  PHB_REQUIRE(h && nRows && nCols && nnz, "phb_amg_host_level_size: NULL argument");
  const HCsr *M = pick(h, level, which);
  PHB_REQUIRE(M, "phb_amg_host_level_size: no matrix %d on level %d", which, level);
  *nRows = M->n; *nCols = M->m; *nnz = M->nnz();
  if (rho) *rho = h->H.lev[level].rho;
  return PHB_OK;
}

int phb_amg_host_level_weight(const phb_amg_host *h, int level, double *wScale) {
  PHB_REQUIRE(h && wScale && level >= 0 && level < (int)h->H.lev.size(), "phb_amg_host_level_weight: bad argument");
  *wScale = h->H.lev[level].wScale;
  return PHB_OK;
}

int phb_amg_host_level_csr(const phb_amg_host *h, int level, int which, int *rowPtr, int *colInd, double *vals) {
  PHB_REQUIRE(h && rowPtr && colInd && vals, "phb_amg_host_level_csr: NULL argument");
  const HCsr *M = pick(h, level, which);
  PHB_REQUIRE(M, "phb_amg_host_level_csr: no matrix %d on level %d", which, level);
  std::copy(M->rp.begin(), M->rp.end(), rowPtr);
  std::copy(M->ci.begin(), M->ci.end(), colInd);
  std::copy(M->v.begin(), M->v.end(), vals);
  return PHB_OK;
}

int phb_amg_host_coarse_inverse(const phb_amg_host *h, double *inv) {
  PHB_REQUIRE(h && inv, "phb_amg_host_coarse_inverse: NULL argument");
  PHB_REQUIRE(!h->H.coarseInv.empty(), "phb_amg_host_coarse_inverse: the coarsest level is not dense");
  std::copy(h->H.coarseInv.begin(), h->H.coarseInv.end(), inv);
  return PHB_OK;
}

int phb_amg_host_destroy(phb_amg_host *h) {
  delete h;
  return PHB_OK;
}

// out = ms per launch of the level-0 [residual, restriction, prolongation, Jacobi sweep], ms per whole cycle,
// algorithmic bytes of the Jacobi launch, of the cycle, launches per cycle (single rank: no exchanges are timed)
int phb_solver_time_amg(phb_solver *s, int reps, double out[8]) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(s && out, "phb_solver_time_amg: NULL argument");
  PHB_REQUIRE(s->ctx->nProcs == 1 || s->amg.nDist == 0, "phb_solver_time_amg: single-rank hierarchies only");
  return phb::amg_time(s, reps, out);
  PHB_TRY_END
}

// [levels, operator complexity, setup ms (host), setups so far, coarsest rows, kernel launches per cycle,
//  iterations of the first solve after the last setup, hierarchy currently stale (0/1)]
// out: [0] refreshes since the last host setup, [1] ms of the last one (device time), [2] iterations of the first solve
// after it, [3] 1 when the symbolic data for the numeric re-setup is resident, [4] its bytes
int phb_solver_amg_refresh_info(const phb_solver *s, double out[8]) {
  PHB_REQUIRE(s && out, "phb_solver_amg_refresh_info: NULL argument");
  const AmgData &D = s->amg;
  for (int k = 0; k < 8; ++k) out[k] = 0.;
  out[0] = D.refreshes; out[1] = D.refreshMs; out[2] = D.itersAfterRefresh; out[3] = D.refresh ? 1. : 0.;
  out[4] = D.refresh ? D.refresh->bytes : 0.;
  return PHB_OK;
}

// Numeric re-setup now, from the matrix values resident in the solver (what `amgRefresh` does on its own when the
// iteration count drifts).  PHB_ERR_STATE when there is no hierarchy or no symbolic data to refresh.
int phb_solver_amg_refresh(phb_solver *s) {
  PHB_REQUIRE(s, "phb_solver_amg_refresh: NULL solver");
  AmgData &D = s->amg;
  if (!D.built || !D.refresh || s->ctx->nProcs != 1 || !s->dVals) {
    set_error("phb_solver_amg_refresh: no single-rank hierarchy with symbolic data (solve once with preconditioner amg, amgRefresh auto|always)");
    return PHB_ERR_STATE;
  }
  const int rc = D.single ? phb::refresh_numeric_f(s) : phb::refresh_numeric_d(s);
  if (rc > 0) {
    set_error("phb_solver_amg_refresh: zero diagonal or vanishing pivot on a coarse level");
    return PHB_ERR_BREAKDOWN;
  }
  return rc;
}

// Values of one matrix of the hierarchy as the cycle streams them (sliced-ELL slots, converted to double):
// which = 0 operator (levels >= 1), 1 prolongator, 2 restriction, 3 smoother weights, 4 dense coarse inverse (level ignored).
// Returns the number of values; with out == NULL only the count.
long long phb_solver_amg_values(const phb_solver *s, int level, int which, double *out, long long cap) {
  if (!s || !s->amg.built) { set_error("phb_solver_amg_values: no hierarchy"); return PHB_ERR_STATE; }
  const AmgData &D = s->amg;
  if (level < 0 || level >= (int)D.lev.size()) { set_error("phb_solver_amg_values: bad level"); return PHB_ERR_ARG; }
  const AmgLevel &L = *D.lev[level];
  const void *p = nullptr;
  long long n = 0;
  bool cyc = true;
  if (which == 0 && level > 0) { p = L.A.vals.p; n = L.A.pat.nSlots; }
  else if (which == 1 && level + 1 < (int)D.lev.size()) { p = L.P.vals.p; n = L.P.pat.nSlots; }
  else if (which == 2 && level + 1 < (int)D.lev.size()) { p = L.R.vals.p; n = L.R.pat.nSlots; }
  else if (which == 3) { p = L.w.p; n = L.n; }
  else if (which == 4 && D.denseCoarse) { p = D.coarseInv.p; n = (long long)D.nCoarse * D.nCoarse; cyc = false; }
  else { set_error("phb_solver_amg_values: level %d has no matrix %d", level, which); return PHB_ERR_ARG; }
  if (!out) return n;
  if (cap < n) { set_error("phb_solver_amg_values: buffer too small"); return PHB_ERR_ARG; }
  cudaStream_t st = s->ctx->stream;
  if (cyc && D.builtSingle) {
    std::vector<float> t((size_t)n);
    PHB_CUDA(cudaMemcpyAsync(t.data(), p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
    PHB_CUDA(cudaStreamSynchronize(st));
    for (long long k = 0; k < n; ++k) out[k] = t[k];
  } else {
    PHB_CUDA(cudaMemcpyAsync(out, p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    PHB_CUDA(cudaStreamSynchronize(st));
  }
  return n;
}

int phb_solver_amg_info(const phb_solver *s, double out[8]) {
  PHB_REQUIRE(s && out, "phb_solver_amg_info: NULL argument");
  const AmgData &D = s->amg;
  out[0] = (double)D.lev.size(); out[1] = D.opComplexity; out[2] = D.setupMs; out[3] = D.setups;
  out[4] = D.nCoarse; out[5] = phb::amg_launches_per_apply(s); out[6] = D.itersAfterSetup; out[7] = D.stale ? 1. : 0.;
  return PHB_OK;
}

}  // extern "C"
