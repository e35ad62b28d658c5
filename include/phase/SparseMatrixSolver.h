// SparseMatrixSolver.h -- Seam 1: the abstract backend interface, virtual for
// virtual as in the reference (src/Math/SparseMatrixSolver.h:11-62,
// SparseMatrixSolver.cpp:5-46) plus the enum value of the new backend.
#ifndef PHASE_B200_SPARSE_MATRIX_SOLVER_H
#define PHASE_B200_SPARSE_MATRIX_SOLVER_H
#include <cstdio>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "Exception.h"
#include "PropertyTree.h"
#include "SparseEntry.h"
#include "Vector.h"

class SparseMatrixSolver {
public:
  enum Type { EIGEN, TRILINOS_BELOS, TRILINOS_AMESOS2, TRILINOS_MUELU, B200 };
  typedef std::pair<Index, Scalar> Entry;
  typedef std::vector<Entry> Row;
  typedef std::vector<Row> CoefficientList;

  virtual ~SparseMatrixSolver() {}
  virtual Type type() const = 0;
  virtual void setRank(int rank) = 0;
  virtual void setRank(int rowRank, int colRank) = 0;
  // duplicates are summed (SparseMatrixSolver.cpp:5-31)
  virtual void set(const std::vector<std::tuple<Index, Index, Scalar>> &entries) {
    CoefficientList coeffs;
    for (const auto &e : entries) {
      const Index row = std::get<0>(e), col = std::get<1>(e);
      const Scalar val = std::get<2>(e);
      if (row >= (Index)coeffs.size()) coeffs.resize(row + 1);
      bool isNew = true;
      for (Entry &x : coeffs[row])
        if (x.first == col) { x.second += val; isNew = false; break; }
      if (isNew) coeffs[row].push_back(Entry(col, val));
    }
    set(coeffs);
  }
  virtual void set(const CoefficientList &eqn) = 0;
  virtual void set(const std::vector<Index> &rowPtr, const std::vector<Index> &colInds,
                   const std::vector<Scalar> &vals) = 0;
  virtual void set(const std::vector<SparseEntry> &entries) = 0;
  virtual void setGuess(const Vector &x0) = 0;
  virtual void setRhs(const Vector &rhs) = 0;
  virtual Scalar solve() = 0;
  virtual Scalar solve(const Vector &x0) { setGuess(x0); return solve(); }
  virtual Scalar solveLeastSquares() {
    throw Exception("SparseMatrixSolver", "solveLeastSquares",
                    "least squares solver is not available for this sparse matrix solver type.");
  }
  virtual Scalar x(Index idx) const = 0;
  virtual void setup(const boost::property_tree::ptree &) {}
  virtual int nIters() const = 0;
  virtual Scalar error() const = 0;
  virtual bool supportsMPI() const = 0;
  virtual void printStatus(const std::string &msg) const {
    printf("%s iterations = %d, error = %lf.\n", msg.c_str(), nIters(), error());
  }

protected:
  int nPreconUses_ = 1, maxPreconUses_ = 1;
};
#endif
