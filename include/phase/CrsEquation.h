// CrsEquation.h -- host CSR equation with the reference's exact insertion and
// compaction semantics (src/Math/CrsEquation.{h,cpp}): ELL-nnz padded rows with
// colInd = -1, first-free-slot addCoeff with insert fallback (:113-131), and
// operator+=/-= that rebuild compact rows dropping exact zeros (:185-275).
// It is the fall-through path for per-entry add()/set() callers (immersed-boundary
// modules build arbitrary (cell, nb) stencils this way); equations
// built from fv:: operators never materialise it (see FiniteVolumeEquation.h).
#ifndef PHASE_B200_CRS_EQUATION_H
#define PHASE_B200_CRS_EQUATION_H
#include <algorithm>
#include <numeric>
#include <ostream>

#include "SparseMatrixSolver.h"

class CrsEquation {
public:
  CrsEquation() : rowPtr_(1, 0) {}
  CrsEquation(Size nRows, Size nnz)
      : rowPtr_(nRows + 1, (Index)nnz), colInd_(nRows * nnz, -1), vals_(nRows * nnz), rhs_(nRows, 0.) {
    rowPtr_[0] = 0;
    std::partial_sum(rowPtr_.begin(), rowPtr_.end(), rowPtr_.begin());
  }
  CrsEquation(const CrsEquation &) = default;
  CrsEquation(CrsEquation &&) = default;
  virtual ~CrsEquation() {}
  // assignment deliberately keeps the destination's solver (M/CrsEquation.cpp:16-26)
  CrsEquation &operator=(const CrsEquation &eqn) {
    if (this != &eqn) {
      solver_ = eqn.solver_ ? eqn.solver_ : solver_;
      rowPtr_ = eqn.rowPtr_; colInd_ = eqn.colInd_; vals_ = eqn.vals_; rhs_ = eqn.rhs_;
    }
    return *this;
  }
  CrsEquation &operator=(CrsEquation &&eqn) {
    solver_ = eqn.solver_ ? eqn.solver_ : solver_;
    rowPtr_ = std::move(eqn.rowPtr_); colInd_ = std::move(eqn.colInd_);
    vals_ = std::move(eqn.vals_); rhs_ = std::move(eqn.rhs_);
    return *this;
  }
  void setRank(Size rank) {
    rowPtr_.resize(rank + 1, rowPtr_.back());
    rhs_.resize(rank, 0.);
    if (solver_) solver_->setRank((int)rank);
  }
  void clear() { rowPtr_.assign(1, 0); colInd_.clear(); vals_.clear(); rhs_.clear(); }
  Size rank() const { return rowPtr_.size() - 1; }
  Size capacity(Size row) const { return rowPtr_[row + 1] - rowPtr_[row]; }
  void addRow(Size nnz) {
    rowPtr_.push_back(rowPtr_.back() + (Index)nnz);
    colInd_.resize(colInd_.size() + nnz, -1);
    vals_.resize(vals_.size() + nnz);
    rhs_.resize(rhs_.size() + 1, 0.);
  }
  void addCoeff(Index localRow, Index globalCol, Scalar val) { put(localRow, globalCol, val, true); }
  void setCoeff(Index localRow, Index globalCol, Scalar val) { put(localRow, globalCol, val, false); }
  void scaleRow(Index localRow, Scalar val) {
    for (Index j = rowPtr_[localRow]; j < rowPtr_[localRow + 1]; ++j) vals_[j] *= val;
    rhs_(localRow) *= val;
  }
  void addRhs(Index localRow, Scalar val) { rhs_(localRow) += val; }
  void setRhs(Index localRow, Scalar val) { rhs_(localRow) = val; }
  const std::vector<Index> &rowPtr() const { return rowPtr_; }
  const std::vector<Index> &colInd() const { return colInd_; }
  const std::vector<Scalar> &vals() const { return vals_; }
  Scalar coeff(Index localRow, Index globalCol) const {
    for (Index j = rowPtr_[localRow]; j < rowPtr_[localRow + 1]; ++j)
      if (colInd_[j] == globalCol) return vals_[j];
    return 0.;
  }
  Scalar x(Index idx) const { return solver_->x(idx); }
  Scalar b(Index idx) const { return rhs_(idx); }
  void setSparseSolver(const std::shared_ptr<SparseMatrixSolver> &solver) { solver_ = solver; }
  const std::shared_ptr<SparseMatrixSolver> &sparseSolver() const { return solver_; }
  // hand-off: set(rowPtr,colInd,vals); setRhs(-rhs_) (M/CrsEquation.cpp:169-175)
  virtual Scalar solve() {
    solver_->setRank((int)rank());
    solver_->set(rowPtr_, colInd_, vals_);
    solver_->setRhs(-rhs_);
    solver_->solve();
    return solver_->error();
  }
  CrsEquation &operator+=(const CrsEquation &rhs) { merge(rhs, +1.); rhs_ += rhs.rhs_; return *this; }
  CrsEquation &operator-=(const CrsEquation &rhs) { merge(rhs, -1.); rhs_ -= rhs.rhs_; return *this; }
  CrsEquation &operator+=(const Vector &rhs) { rhs_ += rhs; return *this; }
  CrsEquation &operator-=(const Vector &rhs) { rhs_ -= rhs; return *this; }
  CrsEquation &operator*=(Scalar rhs) { for (Scalar &v : vals_) v *= rhs; rhs_ *= rhs; return *this; }
  CrsEquation &operator/=(Scalar rhs) { for (Scalar &v : vals_) v /= rhs; rhs_ /= rhs; return *this; }
  CrsEquation &operator==(Scalar rhs) { if (rhs != 0.) rhs_ -= rhs; return *this; }
  CrsEquation &operator==(const CrsEquation &rhs) { return operator-=(rhs); }
  CrsEquation &operator==(const Vector &rhs) { return operator-=(rhs); }

protected:
  void put(Index row, Index col, Scalar val, bool add) {
    for (Index j = rowPtr_[row]; j < rowPtr_[row + 1]; ++j) {
      if (colInd_[j] == col) { if (add) vals_[j] += val; else vals_[j] = val; return; }
      if (colInd_[j] == -1) { colInd_[j] = col; vals_[j] = val; return; }
    }
    colInd_.insert(colInd_.begin() + rowPtr_[row + 1], col);
    vals_.insert(vals_.begin() + rowPtr_[row + 1], val);
    for (size_t r = row + 1; r < rowPtr_.size(); ++r) ++rowPtr_[r];
  }
  void merge(const CrsEquation &rhs, Scalar sign) {
    std::vector<Index> tp(1, 0), tc;
    std::vector<Scalar> tv;
    for (Size row = 0; row < rank(); ++row) {
      const size_t first = tc.size();
      for (Index j = rowPtr_[row]; j < rowPtr_[row + 1]; ++j) {
        if (vals_[j] == 0. || colInd_[j] < 0) continue;
        tc.push_back(colInd_[j]); tv.push_back(vals_[j]);
      }
      for (Index j = rhs.rowPtr_[row]; j < rhs.rowPtr_[row + 1]; ++j) {
        if (rhs.vals_[j] == 0. || rhs.colInd_[j] < 0) continue;
        auto it = std::find(tc.begin() + first, tc.end(), rhs.colInd_[j]);
        if (it != tc.end()) tv[it - tc.begin()] += sign * rhs.vals_[j];
        else { tc.push_back(rhs.colInd_[j]); tv.push_back(sign * rhs.vals_[j]); }
      }
      tp.push_back((Index)tc.size());
    }
    rowPtr_ = tp; colInd_ = tc; vals_ = tv;
  }
  std::vector<Index> rowPtr_, colInd_;
  std::vector<Scalar> vals_;
  Vector rhs_;
  std::shared_ptr<SparseMatrixSolver> solver_;
};
inline std::ostream &operator<<(std::ostream &os, const CrsEquation &eqn) {
  for (size_t row = 0; row + 1 < eqn.rowPtr().size(); ++row)
    for (Index j = eqn.rowPtr()[row]; j < eqn.rowPtr()[row + 1]; ++j)
      os << "(" << row << "," << eqn.colInd()[j] << "," << eqn.vals()[j] << ")\n";
  return os;
}
inline CrsEquation operator+(CrsEquation lhs, const CrsEquation &rhs) { lhs += rhs; return lhs; }
inline CrsEquation operator-(CrsEquation lhs, const CrsEquation &rhs) { lhs -= rhs; return lhs; }
inline CrsEquation operator+(CrsEquation lhs, const Vector &rhs) { lhs += rhs; return lhs; }
inline CrsEquation operator-(CrsEquation lhs, const Vector &rhs) { lhs -= rhs; return lhs; }
inline CrsEquation operator*(CrsEquation lhs, Scalar rhs) { lhs *= rhs; return lhs; }
inline CrsEquation operator/(CrsEquation lhs, Scalar rhs) { lhs /= rhs; return lhs; }
#endif
