// B200SparseMatrixSolver.h -- the drop-in backend: a SparseMatrixSolver whose
// virtuals forward to libphase_b200.so (phb_solver_*).  This is the class a
// Phase maintainer adds next to EigenSparseMatrixSolver / Trilinos*SparseMatrixSolver
// (see INTEGRATION.md); `lib b200` in LinearAlgebra.<eqn> selects it.
//
// Contract honoured: rows local / columns global / -1 trailing
// padding skipped; setRhs receives -rhs_ and copies it; x(i) is a host getter valid
// after solve(); setup() reads solver / maxIters / tolerance / preconditioner (jacobi | ilu0 | amg) /
// iluFill and the amg* keys; one instance lives as long as its equation, so the device copy of the
// pattern, the CUDA graph of the iteration and all work vectors are cached in
// the C-side solver across time steps.
#ifndef PHASE_B200_B200_SPARSE_MATRIX_SOLVER_H
#define PHASE_B200_B200_SPARSE_MATRIX_SOLVER_H
#include "Communicator.h"
#include "SparseMatrixSolver.h"

class B200SparseMatrixSolver : public SparseMatrixSolver {
public:
  explicit B200SparseMatrixSolver(const Communicator &comm) : comm_(comm) {
    phase::check(phb_solver_create(comm.handle(), &s_), "B200SparseMatrixSolver", "B200SparseMatrixSolver");
  }
  B200SparseMatrixSolver(const B200SparseMatrixSolver &) = delete;
  ~B200SparseMatrixSolver() { phb_solver_destroy(s_); }

  Type type() const override { return B200; }
  void setRank(int rank) override { setRank(rank, rank); }
  void setRank(int rowRank, int colRank) override {
    phase::check(phb_solver_set_rank(s_, rowRank, colRank), "B200SparseMatrixSolver", "setRank");
    rank_ = rowRank;
  }
  void set(const CoefficientList &eqn) override {
    std::vector<Index> rows, cols;
    std::vector<Scalar> vals;
    for (size_t r = 0; r < eqn.size(); ++r)
      for (const Entry &e : eqn[r]) { rows.push_back((Index)r); cols.push_back(e.first); vals.push_back(e.second); }
    const int n = std::max<int>(rank_, (int)eqn.size());
    phase::check(phb_solver_set_coo(s_, n, (long long)rows.size(), rows.data(), cols.data(), vals.data()),
                 "B200SparseMatrixSolver", "set");
    rank_ = n;
  }
  void set(const std::vector<Index> &rowPtr, const std::vector<Index> &colInds,
           const std::vector<Scalar> &vals) override {
    phase::check(phb_solver_set_csr(s_, (int)rowPtr.size() - 1, rowPtr.data(), colInds.data(), vals.data()),
                 "B200SparseMatrixSolver", "set");
    rank_ = (int)rowPtr.size() - 1;
  }
  void set(const std::vector<SparseEntry> &entries) override {
    std::vector<Index> rows, cols;
    std::vector<Scalar> vals;
    int n = rank_;
    for (const SparseEntry &e : entries) {
      rows.push_back(e.row); cols.push_back(e.col); vals.push_back(e.val);
      n = std::max(n, e.row + 1);
    }
    phase::check(phb_solver_set_coo(s_, n, (long long)rows.size(), rows.data(), cols.data(), vals.data()),
                 "B200SparseMatrixSolver", "set");
    rank_ = n;
  }
  void setGuess(const Vector &x0) override {
    phase::check(phb_solver_set_guess(s_, x0.data().data(), (int)x0.size()), "B200SparseMatrixSolver", "setGuess");
  }
  void setRhs(const Vector &rhs) override {
    phase::check(phb_solver_set_rhs(s_, rhs.data().data(), (int)rhs.size()), "B200SparseMatrixSolver", "setRhs");
  }
  Scalar solve() override {
    phase::check(phb_solver_solve(s_, &iters_, &error_), "B200SparseMatrixSolver", "solve");
    x_.resize(rank_);
    phase::check(phb_solver_get_x(s_, x_.data(), rank_), "B200SparseMatrixSolver", "solve");
    return error_;
  }
  Scalar x(Index idx) const override { return x_[idx]; }
  // keys of LinearAlgebra.<eqn> (reference: M/TrilinosBelosSparseMatrixSolver.cpp:44-86)
  void setup(const boost::property_tree::ptree &p) override {
    static const char *keys[] = {"solver", "maxIters", "tolerance", "preconditioner", "iluFill",
                                 "innerPreconditioner", "schwarzIters", "schwarzCombineMode", "schwarzOverlap",
                                 "ordering", "nullSpace", "itersPerGraph",
                                 // preconditioner amg (the reference's `lib muelu` role)
                                 "amgTheta", "amgCoarsest", "amgSweeps", "amgSmootherWeight", "amgPrecision",
                                 "amgRebuild", "amgScope", "amgTailRows", "amgFuseRows", "amgRefresh", "amgAggTheta",
                                 "amgCoarseSmootherWeight"};
    for (const char *k : keys) {
      const boost::property_tree::ptree *c = p.get_child_optional(k);
      if (c) phase::check(phb_solver_setup(s_, k, c->data().c_str()), "B200SparseMatrixSolver", "setup");
    }
  }
  int nIters() const override { return iters_; }
  Scalar error() const override { return error_; }
  bool supportsMPI() const override { return true; }
  void printStatus(const std::string &msg) const override {
    comm_.printf("%s %s iterations = %d, error = %lf.\n", msg.c_str(), "Krylov", nIters(), error());
  }
  // device-resident path used by FiniteVolumeEquation<T>::solve (no host round trip)
  phb_solver *handle() const { return s_; }
  void setLastResult(int iters, Scalar err) { iters_ = iters; error_ = err; }

private:
  const Communicator &comm_;
  phb_solver *s_ = nullptr;
  int rank_ = 0, iters_ = 0;
  Scalar error_ = 0.;
  std::vector<Scalar> x_;
};
#endif
