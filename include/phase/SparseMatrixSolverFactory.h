// SparseMatrixSolverFactory.h -- string -> backend (reference:
// src/Math/SparseMatrixSolverFactory.{h,cpp}:10-40).  This build ships the b200
// backend only; the reference's library names are refused with the reference's
// own "bad solver type" error rather than silently mapped to something else.
#ifndef PHASE_B200_SPARSE_MATRIX_SOLVER_FACTORY_H
#define PHASE_B200_SPARSE_MATRIX_SOLVER_FACTORY_H
#include "B200SparseMatrixSolver.h"

class SparseMatrixSolverFactory {
public:
  enum Type { EIGEN, TRILINOS_BELOS, TRILINOS_AMESOS2, TRILINOS_MUELU, B200 };
  std::shared_ptr<SparseMatrixSolver> create(Type type, const Communicator &comm) const {
    switch (type) {
    case B200: return std::make_shared<B200SparseMatrixSolver>(comm);
    default: return nullptr;
    }
  }
  std::shared_ptr<SparseMatrixSolver> create(const std::string &type, const Communicator &comm) const {
    if (type == "b200" || type == "phase_b200") return create(B200, comm);
    throw Exception("SparseMatrixSolverFactory", "create", "bad solver type \"" + type + "\".");
  }
};
#endif
