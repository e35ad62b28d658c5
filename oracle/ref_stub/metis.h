/* Stand-in for <metis.h> (TEST INFRASTRUCTURE): METIS_PartMeshDual hands back the partition vector the test
 * installed with phase_metis_set_partition() -- the partition VECTOR is an input of every parity check (the
 * reference's METIS version and options are unpinned), everything the reference derives from it is its own code. */
#ifndef PHASE_ORACLE_METIS_STUB
#define PHASE_ORACLE_METIS_STUB
#include <cstdlib>
typedef int idx_t;
typedef float real_t;
#define METIS_OK 1
void phase_metis_set_partition(const int *part, int n);
int METIS_PartMeshDual(idx_t *ne, idx_t *nn, idx_t *eptr, idx_t *eind, idx_t *vwgt, idx_t *vsize, idx_t *ncommon, idx_t *nparts,
                       real_t *tpwgts, idx_t *options, idx_t *objval, idx_t *epart, idx_t *npart);
inline int METIS_PartGraphRecursive(idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, real_t *,
                                    real_t *, idx_t *, idx_t *, idx_t *) { abort(); return 0; }
#endif
