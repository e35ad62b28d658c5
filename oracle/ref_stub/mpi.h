/* Stand-in for <mpi.h> (no MPI in this image): TEST INFRASTRUCTURE used to compile the reference's own sources in
 * place for oracle/_ref.  One rank: collectives copy, point-to-point calls must not happen. */
#ifndef PHASE_ORACLE_MPI_STUB
#define PHASE_ORACLE_MPI_STUB
#include <cstdlib>
#include <cstring>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR, count_; };
#define MPI_COMM_WORLD 0
#define MPI_ANY_TAG (-1)
#define MPI_ANY_SOURCE (-1)
#define MPI_SUCCESS 0
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
enum { MPI_BYTE = 1, MPI_CHAR = 1, MPI_INT = 4, MPI_LONG = 8, MPI_UNSIGNED_LONG = 9, MPI_DOUBLE = 10, MPI_DERIVED_BASE_ = 100 };
enum { MPI_SUM = 1, MPI_MIN = 2, MPI_MAX = 3 };
inline int phase_mpi_size_(MPI_Datatype t) {
  if (t >= MPI_DERIVED_BASE_) return t - MPI_DERIVED_BASE_;
  return t == MPI_BYTE ? 1 : t == MPI_INT ? 4 : 8;
}
inline int MPI_Init(int *, char ***) { return 0; }
inline int MPI_Finalize() { return 0; }
inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
inline int MPI_Comm_size(MPI_Comm, int *n) { *n = 1; return 0; }
inline int MPI_Barrier(MPI_Comm) { return 0; }
inline int MPI_Type_vector(int count, int blocklen, int, MPI_Datatype t, MPI_Datatype *out) {
  *out = MPI_DERIVED_BASE_ + count * blocklen * phase_mpi_size_(t);
  return 0;
}
inline int MPI_Type_commit(MPI_Datatype *) { return 0; }
inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
inline int MPI_Allreduce(const void *in, void *out, int n, MPI_Datatype t, MPI_Op, MPI_Comm) {
  if (in != out) memcpy(out, in, (size_t)n * phase_mpi_size_(t));
  return 0;
}
inline int MPI_Gather(const void *in, int n, MPI_Datatype t, void *out, int, MPI_Datatype, int, MPI_Comm) {
  memcpy(out, in, (size_t)n * phase_mpi_size_(t));
  return 0;
}
inline int MPI_Allgather(const void *in, int n, MPI_Datatype t, void *out, int, MPI_Datatype, MPI_Comm) {
  memcpy(out, in, (size_t)n * phase_mpi_size_(t));
  return 0;
}
inline int MPI_Gatherv(const void *in, int n, MPI_Datatype t, void *out, const int *, const int *displs, MPI_Datatype,
                       int, MPI_Comm) {
  memcpy((char *)out + (displs ? displs[0] : 0) * phase_mpi_size_(t), in, (size_t)n * phase_mpi_size_(t));
  return 0;
}
inline int MPI_Allgatherv(const void *in, int n, MPI_Datatype t, void *out, const int *, const int *displs,
                          MPI_Datatype, MPI_Comm) {
  memcpy((char *)out + (displs ? displs[0] : 0) * phase_mpi_size_(t), in, (size_t)n * phase_mpi_size_(t));
  return 0;
}
inline int phase_mpi_no_p2p_() { abort(); return 1; }   /* a single rank has nobody to talk to */
inline int MPI_Ssend(const void *, int, MPI_Datatype, int, int, MPI_Comm) { return phase_mpi_no_p2p_(); }
inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { return phase_mpi_no_p2p_(); }
inline int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { return phase_mpi_no_p2p_(); }
inline int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { return phase_mpi_no_p2p_(); }
inline int MPI_Waitall(int n, MPI_Request *, MPI_Status *) { return n == 0 ? 0 : phase_mpi_no_p2p_(); }
inline int MPI_Probe(int, int, MPI_Comm, MPI_Status *) { return phase_mpi_no_p2p_(); }
inline int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *) { return phase_mpi_no_p2p_(); }
#endif
