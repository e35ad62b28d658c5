// ref_driver.cpp -- ORACLE SIDE (test infrastructure only).
//
// C-ABI shim around the reference's OWN CrsEquation / Vector /
// SparseMatrixSolver translation units, which oracle/build_ref.py compiles
// where they lie under /root/reference/src/Math (no reference source is copied
// into this repo).  It gives the tests the exact reference behaviour of
//   * CrsEquation(nRows,nnz) ELL padding, addCoeff/setCoeff slot search and
//     insert fallback (M/CrsEquation.cpp:9-14,113-151),
//   * operator+=/-=/== compaction (M/CrsEquation.cpp:185-311),
//   * solve(): set(rowPtr,colInd,vals) + setRhs(-rhs_) hand-off
//     (M/CrsEquation.cpp:169-175) seen through a recording SparseMatrixSolver.
#include <cstring>
#include <memory>
#include <vector>

#include "Math/CrsEquation.h"

namespace {
class RecordingSolver : public SparseMatrixSolver {
public:
  Type type() const override { return EIGEN; }
  void setRank(int rank) override { rank_ = rank; x_.assign(rank, 0.); }
  void setRank(int r, int) override { setRank(r); }
  void set(const CoefficientList &) override {}
  void set(const std::vector<Index> &rowPtr, const std::vector<Index> &colInds,
           const std::vector<Scalar> &vals) override {
    rowPtr_ = rowPtr; colInd_ = colInds; vals_ = vals;
  }
  void set(const std::vector<SparseEntry> &) override {}
  void setGuess(const Vector &) override {}
  void setRhs(const Vector &rhs) override { b_ = rhs.data(); }
  Scalar solve() override { return 0.; }
  Scalar x(Index i) const override { return x_[i]; }
  int nIters() const override { return 1; }
  Scalar error() const override { return 0.; }
  bool supportsMPI() const override { return false; }
  int rank_ = 0;
  std::vector<Index> rowPtr_, colInd_;
  std::vector<Scalar> vals_, b_, x_;
};
struct Eq {
  CrsEquation e;
  std::shared_ptr<RecordingSolver> rec;
};
}  // namespace

extern "C" {
void *ref_crs_create(int nRows, int nnz) {
  Eq *q = new Eq{CrsEquation((Size)nRows, (Size)nnz), std::make_shared<RecordingSolver>()};
  q->e.setSparseSolver(q->rec);
  return q;
}
void *ref_crs_clone(void *h) {
  Eq *s = (Eq *)h;
  Eq *q = new Eq{s->e, std::make_shared<RecordingSolver>()};
  q->e.setSparseSolver(q->rec);
  return q;
}
void ref_crs_destroy(void *h) { delete (Eq *)h; }
void ref_crs_add_coeff(void *h, int r, int c, double v) { ((Eq *)h)->e.addCoeff(r, c, v); }
void ref_crs_set_coeff(void *h, int r, int c, double v) { ((Eq *)h)->e.setCoeff(r, c, v); }
void ref_crs_add_rhs(void *h, int r, double v) { ((Eq *)h)->e.addRhs(r, v); }
void ref_crs_scale_row(void *h, int r, double v) { ((Eq *)h)->e.scaleRow(r, v); }
void ref_crs_add_eq(void *l, void *r) { ((Eq *)l)->e += ((Eq *)r)->e; }
void ref_crs_sub_eq(void *l, void *r) { ((Eq *)l)->e == ((Eq *)r)->e; }
void ref_crs_sub_vec(void *l, const double *v, int n) {
  Vector vec(n);
  for (int i = 0; i < n; ++i) vec(i) = v[i];
  ((Eq *)l)->e == vec;
}
void ref_crs_scale(void *l, double s) { ((Eq *)l)->e *= s; }
int ref_crs_rank(void *h) { return (int)((Eq *)h)->e.rank(); }
int ref_crs_nnz(void *h) { return (int)((Eq *)h)->e.colInd().size(); }
void ref_crs_export(void *h, int *rowPtr, int *colInd, double *vals, double *rhs) {
  const CrsEquation &e = ((Eq *)h)->e;
  std::memcpy(rowPtr, e.rowPtr().data(), e.rowPtr().size() * sizeof(int));
  std::memcpy(colInd, e.colInd().data(), e.colInd().size() * sizeof(int));
  std::memcpy(vals, e.vals().data(), e.vals().size() * sizeof(double));
  for (Size i = 0; i < e.rank(); ++i) rhs[i] = e.b((Index)i);
}
// CrsEquation::solve() hand-off as the backend sees it
int ref_crs_solve_handoff(void *h, int *rowPtr, int *colInd, double *vals, double *b) {
  Eq *q = (Eq *)h;
  q->e.solve();
  std::memcpy(rowPtr, q->rec->rowPtr_.data(), q->rec->rowPtr_.size() * sizeof(int));
  std::memcpy(colInd, q->rec->colInd_.data(), q->rec->colInd_.size() * sizeof(int));
  std::memcpy(vals, q->rec->vals_.data(), q->rec->vals_.size() * sizeof(double));
  std::memcpy(b, q->rec->b_.data(), q->rec->b_.size() * sizeof(double));
  return q->rec->rank_;
}
}
