/*
 * Drop-in replacement for the hot-path part of the reference's src/include/hcore_z.h:
 *     HCORE_zgemm       reference src/include/hcore_z.h:27-48
 *                       (definition src/compute/hcore_zgemm.c:663-975)
 *     HCORE_zuncompress reference src/include/hcore_z.h:76-88
 * Complex double tiles U * V^T with a PLAIN transpose (reference hcore_zgemm.c:165, :189, :309),
 * real ranks (double*), real accuracy; rk > 0 selects fixed-rank truncation (:376).
 * HOST pointers, identical argument list.  C callers see `double _Complex`; for C++ (and the CUDA
 * host code) the scalar is a layout- and ABI-compatible struct of two doubles (x86-64 SysV passes
 * both in two SSE registers).
 */
#ifndef HCORE_B200_DROPIN_HCORE_Z_H
#define HCORE_B200_DROPIN_HCORE_Z_H
#include "flop_util_structs.h"
#include "hcore.h"
#include "hcore_b200.h"

#ifdef __cplusplus
extern "C" {
typedef hcore_b200_dcomplex HCORE_Complex64_t;
typedef void HCORE_zarray_t;
#else
#include <complex.h>
typedef double _Complex HCORE_Complex64_t;
typedef double _Complex HCORE_zarray_t;
#endif

void HCORE_zgemm(HCORE_enum transA, int transB, int M, int N, HCORE_Complex64_t alpha,
                 HCORE_zarray_t *AU, HCORE_zarray_t *AV, double *Ark, int LDA,
                 HCORE_zarray_t *BU, HCORE_zarray_t *BV, double *Brk, int LDB,
                 HCORE_Complex64_t beta,
                 HCORE_zarray_t *CU, HCORE_zarray_t *CV, double *Crk, int LDC,
                 int rk, int maxrk, double acc, HCORE_zarray_t *work, flop_counter *flops);

/* CD (M x N) <- AU BV^T (plain transpose); Ark[0] must equal Brk[0]; alpha and beta are ignored as in the reference
 * (reference src/include/hcore_z.h:76-88, definition src/compute/hcore_zuncompress.c:35-95). */
void HCORE_zuncompress(HCORE_enum transA, HCORE_enum transB, int M, int N, HCORE_Complex64_t alpha,
                       HCORE_zarray_t *AU, double *Ark, int LDA,
                       HCORE_zarray_t *BV, double *Brk, int LDB,
                       HCORE_Complex64_t beta, HCORE_zarray_t *CD, int LDC);

#ifdef __cplusplus
}
#endif
#endif
