/* TEST INFRASTRUCTURE ONLY (oracle build shim; never linked into the product).
 *
 * Minimal <lapacke.h> for compiling the UNMODIFIED reference against SciPy's
 * OpenBLAS (`scipy_`-prefixed LP64 symbols).  Declares just what the reference's
 * src/ and testing/ call (SURVEY.md section 8(c) lists them).
 */
#ifndef ORACLE_SHIM_LAPACKE_H
#define ORACLE_SHIM_LAPACKE_H
#include <stdlib.h>
#include <stdint.h>
#include <complex.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
typedef int lapack_int;
typedef double _Complex lapack_complex_double;
typedef float  _Complex lapack_complex_float;

#define SHIM_LAPACK(name) scipy_##name

#define LAPACKE_dgeqrf      SHIM_LAPACK(LAPACKE_dgeqrf)
#define LAPACKE_dgesvd      SHIM_LAPACK(LAPACKE_dgesvd)
#define LAPACKE_dgesvd_work SHIM_LAPACK(LAPACKE_dgesvd_work)
#define LAPACKE_dormqr      SHIM_LAPACK(LAPACKE_dormqr)
#define LAPACKE_dorgqr      SHIM_LAPACK(LAPACKE_dorgqr)
#define LAPACKE_dlacpy      SHIM_LAPACK(LAPACKE_dlacpy)
#define LAPACKE_dlaset      SHIM_LAPACK(LAPACKE_dlaset)
#define LAPACKE_dlange      SHIM_LAPACK(LAPACKE_dlange)
#define LAPACKE_dge_trans   SHIM_LAPACK(LAPACKE_dge_trans)
#define LAPACKE_dlatms      SHIM_LAPACK(LAPACKE_dlatms)
#define LAPACKE_dpotrf_work SHIM_LAPACK(LAPACKE_dpotrf_work)
#define LAPACKE_sgeqrf      SHIM_LAPACK(LAPACKE_sgeqrf)
#define LAPACKE_sgesvd      SHIM_LAPACK(LAPACKE_sgesvd)
#define LAPACKE_sormqr      SHIM_LAPACK(LAPACKE_sormqr)
#define LAPACKE_slacpy      SHIM_LAPACK(LAPACKE_slacpy)
#define LAPACKE_sge_trans   SHIM_LAPACK(LAPACKE_sge_trans)
#define LAPACKE_slatms      SHIM_LAPACK(LAPACKE_slatms)
#define LAPACKE_zgeqrf      SHIM_LAPACK(LAPACKE_zgeqrf)
#define LAPACKE_zgesvd      SHIM_LAPACK(LAPACKE_zgesvd)
#define LAPACKE_zunmqr      SHIM_LAPACK(LAPACKE_zunmqr)
#define LAPACKE_zlacpy      SHIM_LAPACK(LAPACKE_zlacpy)
#define LAPACKE_zlaset      SHIM_LAPACK(LAPACKE_zlaset)
#define LAPACKE_zge_trans   SHIM_LAPACK(LAPACKE_zge_trans)

lapack_int LAPACKE_dgeqrf(int layout, lapack_int m, lapack_int n, double *a, lapack_int lda, double *tau);
lapack_int LAPACKE_dgesvd(int layout, char jobu, char jobvt, lapack_int m, lapack_int n, double *a,
                          lapack_int lda, double *s, double *u, lapack_int ldu, double *vt,
                          lapack_int ldvt, double *superb);
lapack_int LAPACKE_dgesvd_work(int layout, char jobu, char jobvt, lapack_int m, lapack_int n, double *a,
                          lapack_int lda, double *s, double *u, lapack_int ldu, double *vt,
                          lapack_int ldvt, double *work, lapack_int lwork);
lapack_int LAPACKE_dormqr(int layout, char side, char trans, lapack_int m, lapack_int n, lapack_int k,
                          const double *a, lapack_int lda, const double *tau, double *c, lapack_int ldc);
lapack_int LAPACKE_dorgqr(int layout, lapack_int m, lapack_int n, lapack_int k, double *a,
                          lapack_int lda, const double *tau);
lapack_int LAPACKE_dlacpy(int layout, char uplo, lapack_int m, lapack_int n, const double *a,
                          lapack_int lda, double *b, lapack_int ldb);
lapack_int LAPACKE_dlaset(int layout, char uplo, lapack_int m, lapack_int n, double alpha,
                          double beta, double *a, lapack_int lda);
double     LAPACKE_dlange(int layout, char norm, lapack_int m, lapack_int n, const double *a, lapack_int lda);
void       LAPACKE_dge_trans(int layout, lapack_int m, lapack_int n, const double *in, lapack_int ldin,
                          double *out, lapack_int ldout);
lapack_int LAPACKE_dlatms(int layout, lapack_int m, lapack_int n, char dist, lapack_int *iseed, char sym,
                          double *d, lapack_int mode, double cond, double dmax, lapack_int kl,
                          lapack_int ku, char pack, double *a, lapack_int lda);
lapack_int LAPACKE_dpotrf_work(int layout, char uplo, lapack_int n, double *a, lapack_int lda);

lapack_int LAPACKE_sgeqrf(int layout, lapack_int m, lapack_int n, float *a, lapack_int lda, float *tau);
lapack_int LAPACKE_sgesvd(int layout, char jobu, char jobvt, lapack_int m, lapack_int n, float *a,
                          lapack_int lda, float *s, float *u, lapack_int ldu, float *vt,
                          lapack_int ldvt, float *superb);
lapack_int LAPACKE_sormqr(int layout, char side, char trans, lapack_int m, lapack_int n, lapack_int k,
                          const float *a, lapack_int lda, const float *tau, float *c, lapack_int ldc);
lapack_int LAPACKE_slacpy(int layout, char uplo, lapack_int m, lapack_int n, const float *a,
                          lapack_int lda, float *b, lapack_int ldb);
void       LAPACKE_sge_trans(int layout, lapack_int m, lapack_int n, const float *in, lapack_int ldin,
                          float *out, lapack_int ldout);
lapack_int LAPACKE_slatms(int layout, lapack_int m, lapack_int n, char dist, lapack_int *iseed, char sym,
                          float *d, lapack_int mode, float cond, float dmax, lapack_int kl,
                          lapack_int ku, char pack, float *a, lapack_int lda);

lapack_int LAPACKE_zgeqrf(int layout, lapack_int m, lapack_int n, lapack_complex_double *a,
                          lapack_int lda, lapack_complex_double *tau);
lapack_int LAPACKE_zgesvd(int layout, char jobu, char jobvt, lapack_int m, lapack_int n,
                          lapack_complex_double *a, lapack_int lda, double *s,
                          lapack_complex_double *u, lapack_int ldu, lapack_complex_double *vt,
                          lapack_int ldvt, double *superb);
lapack_int LAPACKE_zunmqr(int layout, char side, char trans, lapack_int m, lapack_int n, lapack_int k,
                          const lapack_complex_double *a, lapack_int lda,
                          const lapack_complex_double *tau, lapack_complex_double *c, lapack_int ldc);
lapack_int LAPACKE_zlacpy(int layout, char uplo, lapack_int m, lapack_int n,
                          const lapack_complex_double *a, lapack_int lda,
                          lapack_complex_double *b, lapack_int ldb);
lapack_int LAPACKE_zlaset(int layout, char uplo, lapack_int m, lapack_int n,
                          lapack_complex_double alpha, lapack_complex_double beta,
                          lapack_complex_double *a, lapack_int lda);
void       LAPACKE_zge_trans(int layout, lapack_int m, lapack_int n, const lapack_complex_double *in,
                          lapack_int ldin, lapack_complex_double *out, lapack_int ldout);

/* Fortran-interface routines the reference calls directly. */
#define dlaset_ SHIM_LAPACK(dlaset_)
#define dlacpy_ SHIM_LAPACK(dlacpy_)
#define dlange_ SHIM_LAPACK(dlange_)
#define dlamch_ SHIM_LAPACK(dlamch_)
#define slaset_ SHIM_LAPACK(slaset_)
#define slacpy_ SHIM_LAPACK(slacpy_)
#define slange_ SHIM_LAPACK(slange_)
#define zlaset_ SHIM_LAPACK(zlaset_)
#define zlacpy_ SHIM_LAPACK(zlacpy_)
void   dlaset_(const char *uplo, const int *m, const int *n, const double *alpha, const double *beta, double *a, const int *lda);
void   dlacpy_(const char *uplo, const int *m, const int *n, const double *a, const int *lda, double *b, const int *ldb);
double dlange_(const char *norm, const int *m, const int *n, const double *a, const int *lda, double *work);
double dlamch_(const char *cmach);
void   slaset_(const char *uplo, const int *m, const int *n, const float *alpha, const float *beta, float *a, const int *lda);
void   slacpy_(const char *uplo, const int *m, const int *n, const float *a, const int *lda, float *b, const int *ldb);
float  slange_(const char *norm, const int *m, const int *n, const float *a, const int *lda, float *work);
void   zlaset_(const char *uplo, const int *m, const int *n, const void *alpha, const void *beta, void *a, const int *lda);
void   zlacpy_(const char *uplo, const int *m, const int *n, const void *a, const int *lda, void *b, const int *ldb);

/* The reference's tests use the MKL-style un-suffixed names; they need a real
 * prototype (an implicit int return would corrupt the norm). */
#define dlange dlange_
#define dlamch dlamch_
#define slange slange_

#ifdef __cplusplus
}
#endif
#endif
