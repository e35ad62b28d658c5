// SparseEntry.h -- reference: src/Math/SparseEntry.h:6-16.
#ifndef PHASE_B200_SPARSE_ENTRY_H
#define PHASE_B200_SPARSE_ENTRY_H
#include "Types.h"
class SparseEntry {
public:
  SparseEntry() {}
  SparseEntry(Index row, Index col, Scalar val) : row(row), col(col), val(val) {}
  Index row, col;
  Scalar val;
};
#endif
