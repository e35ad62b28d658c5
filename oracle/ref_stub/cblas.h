/* Stand-in for <cblas.h>: the calls resolve to the OpenBLAS that ships inside scipy (symbols prefixed scipy_,
 * LP64).  TEST INFRASTRUCTURE, see oracle/build.py. */
#ifndef PHASE_ORACLE_CBLAS_STUB
#define PHASE_ORACLE_CBLAS_STUB
#ifdef __cplusplus
extern "C" {
#endif
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
void scipy_cblas_dgemm(enum CBLAS_ORDER, enum CBLAS_TRANSPOSE, enum CBLAS_TRANSPOSE, int M, int N, int K, double alpha,
                       const double *A, int lda, const double *B, int ldb, double beta, double *C, int ldc);
void scipy_cblas_dgemv(enum CBLAS_ORDER, enum CBLAS_TRANSPOSE, int M, int N, double alpha, const double *A, int lda,
                       const double *X, int incX, double beta, double *Y, int incY);
#define cblas_dgemm scipy_cblas_dgemm
#define cblas_dgemv scipy_cblas_dgemv
#ifdef __cplusplus
}
#endif
#endif
