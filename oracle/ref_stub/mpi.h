/* Stand-in for <mpi.h> (no MPI in this image): TEST INFRASTRUCTURE used to compile the reference's own sources in
 * place for oracle/_ref.  The ranks are THREADS of one process (oracle/ref_mpi_threads.cpp): collectives meet on a
 * shared board, point-to-point messages go through mailboxes with MPI's matching rules (source, tag), synchronous
 * sends complete when the matching receive is posted.  Outside phase_mpi_run() there is one rank. */
#ifndef PHASE_ORACLE_MPI_STUB
#define PHASE_ORACLE_MPI_STUB
#include <cstdlib>
#include <cstring>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR, bytes_; };
#define MPI_COMM_WORLD 0
#define MPI_ANY_TAG (-1)
#define MPI_ANY_SOURCE (-1)
#define MPI_SUCCESS 0
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
enum { MPI_BYTE = 1, MPI_CHAR = 1, MPI_INT = 4, MPI_LONG = 8, MPI_UNSIGNED_LONG = 9, MPI_DOUBLE = 10, MPI_DERIVED_BASE_ = 100 };
enum { MPI_SUM = 1, MPI_MIN = 2, MPI_MAX = 3 };
int MPI_Init(int *, char ***);
int MPI_Finalize();
int MPI_Comm_rank(MPI_Comm, int *);
int MPI_Comm_size(MPI_Comm, int *);
int MPI_Barrier(MPI_Comm);
int MPI_Type_vector(int count, int blocklen, int stride, MPI_Datatype t, MPI_Datatype *out);
int MPI_Type_commit(MPI_Datatype *);
int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm);
int MPI_Allreduce(const void *in, void *out, int n, MPI_Datatype t, MPI_Op op, MPI_Comm);
int MPI_Gather(const void *in, int n, MPI_Datatype t, void *out, int nr, MPI_Datatype tr, int root, MPI_Comm);
int MPI_Allgather(const void *in, int n, MPI_Datatype t, void *out, int nr, MPI_Datatype tr, MPI_Comm);
int MPI_Gatherv(const void *in, int n, MPI_Datatype t, void *out, const int *counts, const int *displs, MPI_Datatype tr, int root,
                MPI_Comm);
int MPI_Allgatherv(const void *in, int n, MPI_Datatype t, void *out, const int *counts, const int *displs, MPI_Datatype tr,
                   MPI_Comm);
int MPI_Ssend(const void *buf, int n, MPI_Datatype t, int dest, int tag, MPI_Comm);
int MPI_Recv(void *buf, int n, MPI_Datatype t, int source, int tag, MPI_Comm, MPI_Status *);
int MPI_Isend(const void *buf, int n, MPI_Datatype t, int dest, int tag, MPI_Comm, MPI_Request *);
int MPI_Irecv(void *buf, int n, MPI_Datatype t, int source, int tag, MPI_Comm, MPI_Request *);
int MPI_Waitall(int n, MPI_Request *, MPI_Status *);
int MPI_Probe(int source, int tag, MPI_Comm, MPI_Status *);
int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *);
/* run f(rank, user) on nRanks threads, each one MPI rank; returns when all are done */
void phase_mpi_run(int nRanks, void (*f)(int rank, void *user), void *user);
#endif
