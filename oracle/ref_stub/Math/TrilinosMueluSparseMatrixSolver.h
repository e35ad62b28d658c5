// Shadows the reference's Math/TrilinosMueluSparseMatrixSolver.h (needs Trilinos) in the oracle build:
// FiniteVolumeEquation<T>::solve only names the type for its MueLu special case, which the oracle's
// recording backend never takes.  TEST INFRASTRUCTURE.
#ifndef PHASE_ORACLE_MUELU_STUB
#define PHASE_ORACLE_MUELU_STUB
#include <vector>
#include "2D/Geometry/Point2D.h"
#include "Math/SparseMatrixSolver.h"
class TrilinosMueluSparseMatrixSolver : public SparseMatrixSolver {
public:
  void setCoordinates(const std::vector<Point2D> &) {}
};
#endif
