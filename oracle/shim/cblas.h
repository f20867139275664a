/* TEST INFRASTRUCTURE ONLY (oracle build shim; never linked into the product).
 *
 * Minimal <cblas.h> that lets the UNMODIFIED reference sources under
 * /root/reference/src compile against SciPy's bundled OpenBLAS, whose exported
 * symbols all carry a `scipy_` prefix.
 * Only the routines the reference actually calls are declared.
 */
#ifndef ORACLE_SHIM_CBLAS_H
#define ORACLE_SHIM_CBLAS_H
#include <stdlib.h>
#include <stdint.h>
#include <stddef.h>
#include <complex.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum CBLAS_ORDER     {CblasRowMajor=101, CblasColMajor=102} CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE {CblasNoTrans=111, CblasTrans=112, CblasConjTrans=113, CblasConjNoTrans=114} CBLAS_TRANSPOSE;
typedef enum CBLAS_UPLO      {CblasUpper=121, CblasLower=122} CBLAS_UPLO;
typedef enum CBLAS_DIAG      {CblasNonUnit=131, CblasUnit=132} CBLAS_DIAG;
typedef enum CBLAS_SIDE      {CblasLeft=141, CblasRight=142} CBLAS_SIDE;
typedef CBLAS_ORDER CBLAS_LAYOUT;

#define SHIM_BLAS(name) scipy_##name

#define cblas_dcopy SHIM_BLAS(cblas_dcopy)
#define cblas_dscal SHIM_BLAS(cblas_dscal)
#define cblas_dgemm SHIM_BLAS(cblas_dgemm)
#define cblas_dgemv SHIM_BLAS(cblas_dgemv)
#define cblas_dnrm2 SHIM_BLAS(cblas_dnrm2)
#define cblas_dsyrk SHIM_BLAS(cblas_dsyrk)
#define cblas_dtrmm SHIM_BLAS(cblas_dtrmm)
#define cblas_dtrsm SHIM_BLAS(cblas_dtrsm)
#define cblas_scopy SHIM_BLAS(cblas_scopy)
#define cblas_sscal SHIM_BLAS(cblas_sscal)
#define cblas_sgemm SHIM_BLAS(cblas_sgemm)
#define cblas_strmm SHIM_BLAS(cblas_strmm)
#define cblas_strsm SHIM_BLAS(cblas_strsm)
#define cblas_zcopy SHIM_BLAS(cblas_zcopy)
#define cblas_zscal SHIM_BLAS(cblas_zscal)
#define cblas_zgemm SHIM_BLAS(cblas_zgemm)
#define cblas_ztrmm SHIM_BLAS(cblas_ztrmm)

void   cblas_dcopy(int n, const double *x, int incx, double *y, int incy);
void   cblas_dscal(int n, double alpha, double *x, int incx);
double cblas_dnrm2(int n, const double *x, int incx);
void   cblas_dgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int m, int n,
                   double alpha, const double *a, int lda, const double *x, int incx,
                   double beta, double *y, int incy);
void   cblas_dgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb,
                   int m, int n, int k, double alpha, const double *a, int lda,
                   const double *b, int ldb, double beta, double *c, int ldc);
void   cblas_dsyrk(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans,
                   int n, int k, double alpha, const double *a, int lda, double beta,
                   double *c, int ldc);
void   cblas_dtrmm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo,
                   enum CBLAS_TRANSPOSE ta, enum CBLAS_DIAG diag, int m, int n, double alpha,
                   const double *a, int lda, double *b, int ldb);
void   cblas_dtrsm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo,
                   enum CBLAS_TRANSPOSE ta, enum CBLAS_DIAG diag, int m, int n, double alpha,
                   const double *a, int lda, double *b, int ldb);

void   cblas_scopy(int n, const float *x, int incx, float *y, int incy);
void   cblas_sscal(int n, float alpha, float *x, int incx);
void   cblas_sgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb,
                   int m, int n, int k, float alpha, const float *a, int lda,
                   const float *b, int ldb, float beta, float *c, int ldc);
void   cblas_strmm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo,
                   enum CBLAS_TRANSPOSE ta, enum CBLAS_DIAG diag, int m, int n, float alpha,
                   const float *a, int lda, float *b, int ldb);
void   cblas_strsm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo,
                   enum CBLAS_TRANSPOSE ta, enum CBLAS_DIAG diag, int m, int n, float alpha,
                   const float *a, int lda, float *b, int ldb);

void   cblas_zcopy(int n, const void *x, int incx, void *y, int incy);
void   cblas_zscal(int n, const void *alpha, void *x, int incx);
void   cblas_zgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb,
                   int m, int n, int k, const void *alpha, const void *a, int lda,
                   const void *b, int ldb, const void *beta, void *c, int ldc);
void   cblas_ztrmm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo,
                   enum CBLAS_TRANSPOSE ta, enum CBLAS_DIAG diag, int m, int n, const void *alpha,
                   const void *a, int lda, void *b, int ldb);

/* OpenBLAS control entry points (used by the oracle driver, not the reference). */
void  scipy_openblas_set_num_threads(int n);
char *scipy_openblas_get_config(void);

#ifdef __cplusplus
}
#endif
#endif
