/* Stand-in for <metis.h>: the single-rank oracle build never partitions (FiniteVolumeGrid2D::partition is only
 * reached with nProcs > 1); the symbol exists so that FiniteVolumeGrid2D.cpp links. */
#ifndef PHASE_ORACLE_METIS_STUB
#define PHASE_ORACLE_METIS_STUB
#include <cstdlib>
typedef int idx_t;
typedef float real_t;
#define METIS_OK 1
inline int METIS_PartMeshDual(idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, real_t *,
                              idx_t *, idx_t *, idx_t *, idx_t *) { abort(); return 0; }
inline int METIS_PartGraphRecursive(idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, real_t *,
                                    real_t *, idx_t *, idx_t *, idx_t *) { abort(); return 0; }
#endif
