/*
 * hcore_b200.h -- C ABI of the B200-native tile-low-rank (TLR) update path.
 *
 * This is the drop-in boundary: plain C, pointers and sizes
 * only.  It has two layers:
 *
 *  1. The LEGACY single-tile symbols of ecrc/hcore with byte-for-byte identical
 *     prototypes (host pointers, column-major, ranks passed as floating point):
 *       HCORE_dgemm    replaces reference src/include/hcore_d.h:25-46
 *                      (definition src/compute/hcore_dgemm.c:703-1012)
 *       HCORE_dgemmcd  replaces reference src/include/hcore_d.h:68-86
 *                      (definition src/compute/hcore_dgemmcd.c:36-107)
 *       HCORE_dsyrk    replaces reference src/include/hcore_d.h:107-116
 *                      (definition src/compute/hcore_dsyrk.c:47-189)
 *       HCORE_sgemm    replaces reference src/include/hcore_s.h:25-46
 *                      (definition src/compute/hcore_sgemm.c:703-1012)
 *       HCORE_zgemm    replaces reference src/include/hcore_z.h:27-48
 *                      (definition src/compute/hcore_zgemm.c:663-975)
 *     They are declared in include/hcore_d.h / hcore_s.h / hcore_z.h (same file names the reference installs)
 *     and run as a batch of one on the GPU.  There is NO CPU fallback: without a
 *     CUDA device they print an error and exit(-1), the reference's own error
 *     convention (src/compute/hcore_dgemm.c:743-746).
 *
 *  2. The BATCHED entry points this library adds (north star: "the existing C entry
 *     points ... gain a batched variant over thousands of independent tiles").
 *     All array arguments are DEVICE pointers; tiles are independent; errors that
 *     make the reference exit(-1) become a per-tile int32 status.
 *
 * Matrix conventions are the reference's: column-major, a tile is U * V^T with
 * U, V of size nb x rank and leading dimension ld (effectively nb, reference
 * src/compute/hcore_dgemm.c:96-113), the operation is always
 *     C <- alpha * A * B^T + beta * C
 * whatever transA/transB say (they are never read by the reference,
 * src/compute/hcore_dgemm.c:703-755).
 */
#ifndef HCORE_B200_H
#define HCORE_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (per tile, and as function return values) -------------------------- */
#define HCORE_B200_OK                0
#define HCORE_B200_ERR_ZERO_RANK     1  /* Ark, Brk or Crk == 0   (ref hcore_dgemm.c:743-746: exit)      */
#define HCORE_B200_ERR_MAXRK         2  /* maxrk < Ark + Crk      (ref hcore_dgemm.c:924-927: exit)      */
#define HCORE_B200_ERR_UNSUPPORTED   3  /* Ark+Crk above the compiled stack width (256) or > nb              */
#define HCORE_B200_ERR_NEWRANK       4  /* rank after truncation > maxrk (ref hcore_dgemm.c:465-468)     */
#define HCORE_B200_ERR_NO_CONVERGE   5  /* Jacobi SVD hit its sweep limit (ref: dgesvd info != 0, :407)  */
#define HCORE_B200_ERR_RANK_MISMATCH 6  /* uncompress: Ark != Brk (ref hcore_duncompress.c:69: assert)    */
#define HCORE_B200_ERR_CUDA        100  /* a CUDA runtime call failed (function return value only)       */
#define HCORE_B200_ERR_ARG         101  /* invalid argument           (function return value only)       */

/* Largest stacked width Ark + Crk the kernels are compiled for. */
#define HCORE_B200_MAX_STACK       128   /* default X-workspace rank of the dense-C kernels when no stride tells it */
#define HCORE_B200_MAX_STACK_Z     256   /* HCORE_zgemm: Ark + Crk up to 256 (ranks up to 128) */
#define HCORE_B200_MAX_STACK_D     256   /* HCORE_dgemm / HCORE_sgemm: Ark + Crk up to 256 (ranks up to 128) */

typedef struct hcore_b200_context *hcore_b200_handle_t;

/* complex double scalar passed by value (layout of C99 `double _Complex`) */
typedef struct { double re, im; } hcore_b200_dcomplex;

/* ---- context ------------------------------------------------------------------------- */
/* Creates a context on CUDA device `device` (-1: current device).  The context owns the
 * workspace (block-reflector records, G, R factors, U_r / W) and grows it on demand.    */
int  hcore_b200_create(hcore_b200_handle_t *handle, int device);
int  hcore_b200_destroy(hcore_b200_handle_t handle);
/* Upper bound on the number of tiles processed per kernel launch (workspace = chunk *
 * about 3.9 MiB at nb=1024, stack 128).  0 selects the default (16 x #SM).               */
int  hcore_b200_set_chunk(hcore_b200_handle_t handle, int tiles_per_chunk);
const char *hcore_b200_last_error(hcore_b200_handle_t handle);
const char *hcore_b200_version(void);
/* Number of kernels launched by this context so far (bench.py reports it). */
int64_t hcore_b200_launch_count(hcore_b200_handle_t handle);

/* ---- tile scheduler ------------------------------------------------------------------ */
/* Contiguous block partition of `ntiles` over `world` ranks by tile index
 * (SURVEY.md section 8(e)); rank r owns [*first, *first + *count).                       */
void hcore_b200_shard_range(int64_t ntiles, int world, int rank, int64_t *first, int64_t *count);

/* After a sharded batch every rank holds the new ranks / status of ITS tiles; this gathers both arrays (int32, tile
 * order of the unsharded batch) on every rank with ONE ncclAllGather each -- the only collective of the path (SURVEY.md
 * section 8(e)).  nccl_comm is the caller's ncclComm_t (world ranks, one per GPU); NCCL is looked up at run time
 * (symbols already in the process, else libnccl.so.2).  d_*_local hold `count` entries (hcore_b200_shard_range),
 * d_*_all `ntiles`; either pair may be NULL.  world == 1 with a NULL communicator degenerates to a copy. */
int hcore_b200_gather_status(hcore_b200_handle_t handle, void *nccl_comm, int world, int rank, int64_t ntiles,
                             const int32_t *d_rank_local, const int32_t *d_status_local,
                             int32_t *d_rank_all, int32_t *d_status_all, void *stream);

/* ---- batched TLR update + recompression (device pointers) ----------------------------- */
/* Batched HCORE_dgemm (reference src/compute/hcore_dgemm.c:703-1012) over `ntiles`
 * independent tiles.
 *   d_AU[t], d_AV[t] : nb x Ark[t] (ld lda)      d_BU[t], d_BV[t] : nb x Brk[t] (ld ldb)
 *   d_CU[t], d_CV[t] : buffers of nb x maxrk (ld ldc); first Crk[t] columns valid on
 *                      entry, first (new) Crk[t] columns valid on exit
 *   d_Ark, d_Brk     : int32 ranks (in)           d_Crk : int32 ranks (in/out)
 *   rk   != 0 : fixed-rank truncation min(rk, Ark+Crk)      (ref hcore_dgemm.c:448-452)
 *   rk   == 0 : first i >= 1 with sigma_i < acc (absolute)  (ref hcore_dgemm.c:453-464)
 *   d_status  : int32 per tile, HCORE_B200_OK or an error code; a failed tile is left
 *               untouched.  May be NULL.
 * `stream` is a cudaStream_t (NULL: default stream).  Asynchronous w.r.t. the host.     */
int HCORE_dgemm_batched(hcore_b200_handle_t handle, int ntiles, int nb,
                        double alpha,
                        const double *const *d_AU, const double *const *d_AV, const int32_t *d_Ark, int lda,
                        const double *const *d_BU, const double *const *d_BV, const int32_t *d_Brk, int ldb,
                        double beta,
                        double *const *d_CU, double *const *d_CV, int32_t *d_Crk, int ldc,
                        int rk, int maxrk, double acc,
                        int32_t *d_status, void *stream);

/* Same, for factors stored at a fixed element stride between consecutive tiles. */
int HCORE_dgemm_batched_strided(hcore_b200_handle_t handle, int ntiles, int nb,
                        double alpha,
                        const double *d_AU, const double *d_AV, const int32_t *d_Ark, int lda, int64_t strideA,
                        const double *d_BU, const double *d_BV, const int32_t *d_Brk, int ldb, int64_t strideB,
                        double beta,
                        double *d_CU, double *d_CV, int32_t *d_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, double acc,
                        int32_t *d_status, void *stream);

/* Batched HCORE_sgemm (reference src/compute/hcore_sgemm.c:703): same contract with float
 * factor arrays; the arithmetic is carried out in FP64 and the new factors are rounded to float. */
int HCORE_sgemm_batched(hcore_b200_handle_t handle, int ntiles, int nb,
                        float alpha,
                        const float *const *d_AU, const float *const *d_AV, const int32_t *d_Ark, int lda,
                        const float *const *d_BU, const float *const *d_BV, const int32_t *d_Brk, int ldb,
                        float beta,
                        float *const *d_CU, float *const *d_CV, int32_t *d_Crk, int ldc,
                        int rk, int maxrk, float acc,
                        int32_t *d_status, void *stream);
int HCORE_sgemm_batched_strided(hcore_b200_handle_t handle, int ntiles, int nb,
                        float alpha,
                        const float *d_AU, const float *d_AV, const int32_t *d_Ark, int lda, int64_t strideA,
                        const float *d_BU, const float *d_BV, const int32_t *d_Brk, int ldb, int64_t strideB,
                        float beta,
                        float *d_CU, float *d_CV, int32_t *d_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, float acc,
                        int32_t *d_status, void *stream);

/* Batched HCORE_zgemm (reference src/compute/hcore_zgemm.c:663-975): complex double factors
 * (interleaved re/im, i.e. arrays of C99 double _Complex), tiles are U * V^T with a PLAIN
 * transpose, ranks int32, rk > 0 selects fixed-rank truncation (reference :376).              */
int HCORE_zgemm_batched_strided(hcore_b200_handle_t handle, int ntiles, int nb,
                        hcore_b200_dcomplex alpha,
                        const void *d_AU, const void *d_AV, const int32_t *d_Ark, int lda, int64_t strideA,
                        const void *d_BU, const void *d_BV, const int32_t *d_Brk, int ldb, int64_t strideB,
                        hcore_b200_dcomplex beta,
                        void *d_CU, void *d_CV, int32_t *d_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, double acc,
                        int32_t *d_status, void *stream);

/* Pointer-array form of the same: six device arrays of per-tile device pointers. */
int HCORE_zgemm_batched(hcore_b200_handle_t handle, int ntiles, int nb,
                        hcore_b200_dcomplex alpha,
                        const void *const *d_AU, const void *const *d_AV, const int32_t *d_Ark, int lda,
                        const void *const *d_BU, const void *const *d_BV, const int32_t *d_Brk, int ldb,
                        hcore_b200_dcomplex beta,
                        void *const *d_CU, void *const *d_CV, int32_t *d_Crk, int ldc,
                        int rk, int maxrk, double acc,
                        int32_t *d_status, void *stream);

/* Host-buffer variants: every pointer is a HOST pointer (pinned memory gives full overlap).  Slices of the batch are
 * staged through the device on three internal streams (host-to-device copies, kernels, device-to-host copies, chained
 * by events: both PCIe directions and the SMs are busy at the same time) and the call returns when the results are
 * back in host memory.  This is the "reference-facing" end-to-end path bench.py reports as `e2e`.
 * Columns of CU / CV: the first Crk[t] columns go in; on return the first (new) Crk[t] columns hold the factors and the
 * columns from there up to min(maxrk, max over the slice of Ark + Crk) are ZERO (the copy size is fixed before the
 * ranks are known, so nothing has to wait for them).  A failed tile keeps its first Crk[t] columns and its rank. */
int HCORE_dgemm_batched_host(hcore_b200_handle_t handle, int ntiles, int nb,
                        double alpha,
                        const double *h_AU, const double *h_AV, const int32_t *h_Ark, int lda, int64_t strideA,
                        const double *h_BU, const double *h_BV, const int32_t *h_Brk, int ldb, int64_t strideB,
                        double beta,
                        double *h_CU, double *h_CV, int32_t *h_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, double acc,
                        int32_t *h_status);
int HCORE_sgemm_batched_host(hcore_b200_handle_t handle, int ntiles, int nb,
                        float alpha,
                        const float *h_AU, const float *h_AV, const int32_t *h_Ark, int lda, int64_t strideA,
                        const float *h_BU, const float *h_BV, const int32_t *h_Brk, int ldb, int64_t strideB,
                        float beta,
                        float *h_CU, float *h_CV, int32_t *h_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, float acc,
                        int32_t *h_status);
int HCORE_zgemm_batched_host(hcore_b200_handle_t handle, int ntiles, int nb,
                        hcore_b200_dcomplex alpha,
                        const void *h_AU, const void *h_AV, const int32_t *h_Ark, int lda, int64_t strideA,
                        const void *h_BU, const void *h_BV, const int32_t *h_Brk, int ldb, int64_t strideB,
                        hcore_b200_dcomplex beta,
                        void *h_CU, void *h_CV, int32_t *h_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, double acc,
                        int32_t *h_status);

/* ---- split-phase batched update (device pointers) --------------------------------------
 * Batched form of the reference's HCORE_dgemm_qr_svd / HCORE_dgemm_ormqr pair
 * (src/compute/hcore_dgemm.c:1312-1597, :1602-1699): phase 1 stacks, factors and truncates every
 * tile and writes the revealed ranks to d_newrank WITHOUT touching CU, CV or Crk; the caller may
 * then allocate exact-size factor arrays; phase 2 forms CU = Q_u [U_r; 0], CV = Q_v [W; 0] in the
 * arrays it is given (any ldc >= nb) and sets Crk.  The block reflectors stay in the handle's
 * workspace between the two calls, so ntiles must fit one workspace chunk (hcore_b200_set_chunk;
 * default 16 x #SM tiles) and no other update may run on the handle in between
 * (HCORE_B200_ERR_ARG otherwise).  The fused call and the pair give bit-identical factors. */
int HCORE_dgemm_batched_qr_svd(hcore_b200_handle_t handle, int ntiles, int nb,
                        double alpha,
                        const double *const *d_AU, const double *const *d_AV, const int32_t *d_Ark, int lda,
                        const double *const *d_BU, const double *const *d_BV, const int32_t *d_Brk, int ldb,
                        double beta,
                        const double *const *d_CU, const double *const *d_CV, const int32_t *d_Crk, int ldc,
                        int rk, int maxrk, double acc,
                        int32_t *d_newrank, int32_t *d_status, void *stream);
int HCORE_dgemm_batched_ormqr(hcore_b200_handle_t handle, int ntiles, int nb,
                        double *const *d_CU, double *const *d_CV, int32_t *d_Crk, int ldc, void *stream);
int HCORE_dgemm_batched_qr_svd_strided(hcore_b200_handle_t handle, int ntiles, int nb,
                        double alpha,
                        const double *d_AU, const double *d_AV, const int32_t *d_Ark, int lda, int64_t strideA,
                        const double *d_BU, const double *d_BV, const int32_t *d_Brk, int ldb, int64_t strideB,
                        double beta,
                        const double *d_CU, const double *d_CV, const int32_t *d_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, double acc,
                        int32_t *d_newrank, int32_t *d_status, void *stream);
int HCORE_dgemm_batched_ormqr_strided(hcore_b200_handle_t handle, int ntiles, int nb,
                        double *d_CU, double *d_CV, int32_t *d_Crk, int ldc, int64_t strideC, void *stream);

/* ---- B-dense variant (device pointers) --------------------------------------------------
 * Reference HCORE_dgemm_qr_svd_b_dense (src/compute/hcore_dgemm.c:1022-1303, __qrb_b_dense :236-288):
 * the operand B is a DENSE nb x nb tile (leading dimension ldb), C <- alpha (AU AV^T) B^T + beta C; the
 * stacked V panel is [CV | B AV].  Fused call, its first half (ranks to d_newrank; finish with
 * HCORE_dgemm_batched_ormqr) and a strided convenience form. */
int HCORE_dgemm_batched_b_dense(hcore_b200_handle_t handle, int ntiles, int nb,
                        double alpha,
                        const double *const *d_AU, const double *const *d_AV, const int32_t *d_Ark, int lda,
                        const double *const *d_B, int ldb,
                        double beta,
                        double *const *d_CU, double *const *d_CV, int32_t *d_Crk, int ldc,
                        int rk, int maxrk, double acc,
                        int32_t *d_status, void *stream);
int HCORE_dgemm_batched_qr_svd_b_dense(hcore_b200_handle_t handle, int ntiles, int nb,
                        double alpha,
                        const double *const *d_AU, const double *const *d_AV, const int32_t *d_Ark, int lda,
                        const double *const *d_B, int ldb,
                        double beta,
                        const double *const *d_CU, const double *const *d_CV, const int32_t *d_Crk, int ldc,
                        int rk, int maxrk, double acc,
                        int32_t *d_newrank, int32_t *d_status, void *stream);
int HCORE_dgemm_batched_b_dense_strided(hcore_b200_handle_t handle, int ntiles, int nb,
                        double alpha,
                        const double *d_AU, const double *d_AV, const int32_t *d_Ark, int lda, int64_t strideA,
                        const double *d_B, int ldb, int64_t strideB,
                        double beta,
                        double *d_CU, double *d_CV, int32_t *d_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, double acc,
                        int32_t *d_status, void *stream);

/* ---- batched dense-C updates (device pointers) ---------------------------------------- */
/* Batched HCORE_dgemmcd (reference src/compute/hcore_dgemmcd.c:36-107):
 *   CD[t] <- CD[t] - (Au Av^T)(Bu Bv^T)^T ; alpha/beta are hard-wired to -1/+1 there. */
int HCORE_dgemmcd_batched_strided(hcore_b200_handle_t handle, int ntiles, int nb,
                        const double *d_AU, const double *d_AV, const int32_t *d_Ark, int64_t strideA,
                        const double *d_BU, const double *d_BV, const int32_t *d_Brk, int64_t strideB,
                        double *d_CD, int64_t strideC, void *stream);

/* Batched HCORE_dsyrk (reference src/compute/hcore_dsyrk.c:47-189):
 *   CD[t] <- CD[t] + alpha * AU (AV^T AV) AU^T, strict upper triangle set to 0 (:174-181). */
int HCORE_dsyrk_batched_strided(hcore_b200_handle_t handle, int ntiles, int nb,
                        double alpha,
                        const double *d_AU, const double *d_AV, const int32_t *d_Ark, int64_t strideA,
                        double *d_CD, int64_t strideC, void *stream);

/* Pointer-array forms (device arrays of per-tile device pointers, every leading dimension = nb as in the reference);
 * kmax = the largest rank in the batch (sizes the workspace). */
int HCORE_dgemmcd_batched(hcore_b200_handle_t handle, int ntiles, int nb, int kmax,
                        const double *const *d_AU, const double *const *d_AV, const int32_t *d_Ark,
                        const double *const *d_BU, const double *const *d_BV, const int32_t *d_Brk,
                        double *const *d_CD, void *stream);
int HCORE_dsyrk_batched(hcore_b200_handle_t handle, int ntiles, int nb, int kmax, double alpha,
                        const double *const *d_AU, const double *const *d_AV, const int32_t *d_Ark,
                        double *const *d_CD, void *stream);

/* ---- the remaining tile kernels of one TLR-Cholesky step (device pointers) ------------ */
/* Batched HCORE_dgemmbdcd (reference src/compute/hcore_dgemmbdcd.c:35-122):
 *   CD[t] (M x N, ld ldc) <- alpha AU[t] (AV[t]^T BD[t]) + beta CD[t];  AU, AV: M x Ark[t] (ld lda), BD dense M x N (ld ldb).
 *   beta == 0: CD is not read. */
int HCORE_dgemmbdcd_batched_strided(hcore_b200_handle_t handle, int ntiles, int M, int N, double alpha,
                        const double *d_AU, const double *d_AV, const int32_t *d_Ark, int lda, int64_t strideA,
                        const double *d_BD, int ldb, int64_t strideB, double beta,
                        double *d_CD, int ldc, int64_t strideC, void *stream);

/* Batched HCORE_duncompress / HCORE_zuncompress (reference src/compute/hcore_duncompress.c:33-85,
 * hcore_zuncompress.c:35-95): CD[t] (M x N) <- AU[t] BV[t]^T (plain transpose), AU: M x Ark[t], BV: N x Brk[t].
 * A tile with Ark[t] != Brk[t] (the reference asserts) gets HCORE_B200_ERR_RANK_MISMATCH in d_status (may be NULL)
 * and its CD is left untouched.  Complex arrays are interleaved re/im (C99 double _Complex), strides in elements. */
int HCORE_duncompress_batched_strided(hcore_b200_handle_t handle, int ntiles, int M, int N,
                        const double *d_AU, const int32_t *d_Ark, int lda, int64_t strideA,
                        const double *d_BV, const int32_t *d_Brk, int ldb, int64_t strideB,
                        double *d_CD, int ldc, int64_t strideC, int32_t *d_status, void *stream);
int HCORE_zuncompress_batched_strided(hcore_b200_handle_t handle, int ntiles, int M, int N,
                        const void *d_AU, const int32_t *d_Ark, int lda, int64_t strideA,
                        const void *d_BV, const int32_t *d_Brk, int ldb, int64_t strideB,
                        void *d_CD, int ldc, int64_t strideC, int32_t *d_status, void *stream);

/* Batched HCORE_dtrsm (reference src/compute/hcore_dtrsm.c:24-48, cblas_dtrsm semantics, HCORE_enum values):
 *   side = HCORE_Left : BV[t] (M x Brk[t]) <- alpha op(A[t])^-1 BV[t],  A[t] is M x M
 *   side = HCORE_Right: BV[t] (M x Brk[t]) <- alpha BV[t] op(A[t])^-1,  A[t] is Brk[t] x Brk[t]
 * strideA == 0: ONE triangular tile for the whole batch (the diagonal tile of a Cholesky step).
 * d_Brk may be NULL (every tile has Brank columns); otherwise Brank is the largest rank. */
int HCORE_dtrsm_batched_strided(hcore_b200_handle_t handle, int ntiles, int side, int uplo, int transA, int diag,
                        int M, int Brank, double alpha,
                        const double *d_A, int lda, int64_t strideA,
                        double *d_BV, const int32_t *d_Brk, int ldbv, int64_t strideB, void *stream);

/* Batched HCORE_dpotrf (reference src/compute/hcore_dpotrf.c:23-35, LAPACK dpotrf semantics): A[t] = L L^T
 * (uplo = HCORE_Lower) or U^T U (HCORE_Upper) in place; d_info[t] = 0 or the order of the first leading minor that is
 * not positive definite.  Three launches per 64-column panel: a single tile uses the whole GPU. */
int HCORE_dpotrf_batched_strided(hcore_b200_handle_t handle, int ntiles, int uplo, int N,
                        double *d_A, int lda, int64_t strideA, int32_t *d_info, void *stream);

/* ---- tile compression: dense -> U V^T (device pointers) -------------------------------- */
/* The step before the update path.  The reference only DECLARES a generate-and-compress kernel (HCORE_zgytlr,
 * src/include/hcore_z.h:63-75, never defined); its tests compress with LAPACKE_dgesvd and keep
 * rank = #{sigma >= acc}, capped at the storage rank (testing/testing_dgemm.c:119-236).  This entry point gives the
 * same factors on the device: A[t] (nb x nb, ld lda) ~ U[t] V[t]^T with U orthonormal (nb x rank) and V carrying the
 * singular values, truncated by the update path's rule (first i >= 1 with sigma_i < acc, absolute), rank capped at
 * maxrk.  Randomised range finder (p = maxrk + 16 columns, `power_iters` power iterations: 1 is the safe default)
 * whose orthonormalisation and truncation steps are passes of the update pipeline itself.
 * d_U / d_V: buffers of ldu x maxrk per tile (element stride strideUV); d_rank: int32 out; d_status may be NULL. */
int HCORE_dcompress_batched_strided(hcore_b200_handle_t handle, int ntiles, int nb,
                        const double *d_A, int lda, int64_t strideA,
                        double acc, int maxrk, int power_iters,
                        double *d_U, double *d_V, int ldu, int64_t strideUV,
                        int32_t *d_rank, int32_t *d_status, void *stream);

/* ---- flop model ------------------------------------------------------------------------ */
/* The reference's own accounting of one HCORE_dgemm call (flops->update; reference
 * src/misc/descutil/flop_counts.c:3-70 with the call sites of hcore_dgemm.c), used for
 * "effective GFLOP/s".  m = nb, r = rank after truncation.                               */
uint64_t hcore_b200_gemm_flops(int m, int a, int b, int c, int r);

/* ---- stage-level entry points (tests / profiling) -------------------------------------- */
/* Run only the first `nstages` kernels of the pipeline on a strided batch and leave the
 * intermediates in the context workspace; hcore_b200_debug_fetch copies one of them to
 * host memory.  what: 0 = U panel after QR (reflectors), 1 = V panel, 2 = R_u, 3 = R_v,
 * 4 = Yu (U_r), 5 = Yv (W), 6 = sigma, 7 = T factors of the U panel, 8 = U panel before QR */
int hcore_b200_debug_fetch(hcore_b200_handle_t handle, int what, int tile, double *h_out, size_t nelem);
int hcore_b200_debug_set_stages(hcore_b200_handle_t handle, int nstages);
/* Per-stage device time (ms) of the last batched call, measured with CUDA events when
 * enabled (adds synchronisation; off by default).  stage: 0 stack, 1 tsqr, 2 svd, 3 apply */
int hcore_b200_set_stage_timing(hcore_b200_handle_t handle, int enable);
double hcore_b200_stage_ms(hcore_b200_handle_t handle, int stage);

#ifdef __cplusplus
}
#endif
#endif /* HCORE_B200_H */
