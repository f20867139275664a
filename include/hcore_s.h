/*
 * Drop-in replacement for the hot-path part of the reference's src/include/hcore_s.h:
 *     HCORE_sgemm    reference src/include/hcore_s.h:25-46
 *                    (definition src/compute/hcore_sgemm.c:703-1012, a float clone of HCORE_dgemm)
 * Identical prototype, HOST pointers, ranks travel as float.  The B200 path reads the float
 * factors, carries out the whole update + recompression in FP64 (DMMA) and rounds the new
 * factors to float on the way out: results are at least as accurate as the reference's
 * sgeqrf/sgesvd path (see tests/test_gpu_parity.py::test_sgemm_*).
 * The other prototypes of the reference header (HCORE_ssyrk, HCORE_sgemmcd, ...) have no
 * definition in the reference either and are not declared.
 */
#ifndef HCORE_B200_DROPIN_HCORE_S_H
#define HCORE_B200_DROPIN_HCORE_S_H
#include "flop_util_structs.h"
#include "hcore.h"
#ifdef __cplusplus
extern "C" {
#endif

void HCORE_sgemm(HCORE_enum transA, HCORE_enum transB, int M, int N, float alpha,
                 float *AU, float *AV, float *Ark, int LDA,
                 float *BU, float *BV, float *Brk, int LDB,
                 float beta,
                 float *CU, float *CV, float *Crk, int LDC,
                 int rk, int maxrk, float acc, float *work, flop_counter *flops);

/* link-compatible tunables of reference src/compute/hcore_sgemm.c:37-43 (accepted, ignored) */
extern int sgemm_use_trmm, sgemm_print_index, sgemm_print_mat;

#ifdef __cplusplus
}
#endif
#endif
