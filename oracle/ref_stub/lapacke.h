/* Stand-in for <lapacke.h>: see cblas.h. */
#ifndef PHASE_ORACLE_LAPACKE_STUB
#define PHASE_ORACLE_LAPACKE_STUB
#ifdef __cplusplus
extern "C" {
#endif
typedef int lapack_int;
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
lapack_int scipy_LAPACKE_dgetrf(int, lapack_int m, lapack_int n, double *a, lapack_int lda, lapack_int *ipiv);
lapack_int scipy_LAPACKE_dgetri(int, lapack_int n, double *a, lapack_int lda, const lapack_int *ipiv);
lapack_int scipy_LAPACKE_dgesv(int, lapack_int n, lapack_int nrhs, double *a, lapack_int lda, lapack_int *ipiv, double *b,
                               lapack_int ldb);
lapack_int scipy_LAPACKE_dgels(int, char trans, lapack_int m, lapack_int n, lapack_int nrhs, double *a, lapack_int lda,
                               double *b, lapack_int ldb);
double scipy_LAPACKE_dlange(int, char norm, lapack_int m, lapack_int n, const double *a, lapack_int lda);
lapack_int scipy_LAPACKE_dgecon(int, char norm, lapack_int n, const double *a, lapack_int lda, double anorm, double *rcond);
#define LAPACKE_dgetrf scipy_LAPACKE_dgetrf
#define LAPACKE_dgetri scipy_LAPACKE_dgetri
#define LAPACKE_dgesv scipy_LAPACKE_dgesv
#define LAPACKE_dgels scipy_LAPACKE_dgels
#define LAPACKE_dlange scipy_LAPACKE_dlange
#define LAPACKE_dgecon scipy_LAPACKE_dgecon
#ifdef __cplusplus
}
#endif
#endif
