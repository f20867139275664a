// Register-resident block-reflector kernels of the TLR update (FP64, sm_100a).
//
//   tlr_pg_kernel     P = AV^T BV, G = BU P^T (the only stacked-panel columns that are not a
//                     scaled copy of an input factor; reference __qrb,
//                     src/compute/hcore_dgemm.c:175-205).  Operands staged in shared memory with
//                     cp.async, DMMA.
//   tlr_tsqr_kernel   flat-tree tall-skinny Householder QR of X_u = [beta CU | alpha AU] and
//                     X_v = [CV | G] (replaces the dcopy/dscal stacking and LAPACKE_dgeqrf,
//                     src/compute/hcore_dgemm.c:96-123, :216).  The stacked panels are never
//                     materialised: every column is read from its source array.
//   tlr_apply_kernel  CU = Q_u [U_r; 0], CV = Q_v [W; 0] (replaces LAPACKE_dormqr and the
//                     copies/transposes around it, src/compute/hcore_dgemm.c:531-598, :645-674,
//                     :972-982).
//
// Common machinery.  A "record" is the block reflector of one (128-row block, 8-column
// sub-panel): the explicit 128 x 8 reflector panel V in an XOR-swizzled layout plus its 8 x 8
// compact-WY factor T (and T^T), 9 216 contiguous bytes.  Records are produced once by the
// QR kernel and consumed by ONE TMA bulk copy (cp.async.bulk) each, double buffered, both by
// the left-looking QR itself and by the apply kernel.  The matrix the reflectors act on lives
// in REGISTERS: a warp owns 8 columns, held transposed as DMMA accumulator fragments
// (lane (g, tg) holds column g, rows 8i + 2tg, 8i + 2tg + 1 of the 128-row block).  With a
// consistent permutation of the contraction index an accumulator fragment is directly usable as
// the A operand of the next DMMA, so the whole chain
//      S^T = [S0^T +] B^T V ,  S2^T = S^T T(^T) ,  B^T -= S2^T V^T
// runs with no shared-memory round trip and no CTA barrier; only V and T are read from shared
// memory, conflict-free thanks to the swizzle.  The sub-panel width equals the 8 columns a warp
// owns, so a sub-panel is FACTORED by its owner warp alone, in registers (warp_factor below).
#pragma once
#include "common.cuh"

namespace hcore_b200 {

constexpr int kSubW = 8;                           // sub-panel width = columns per warp
constexpr int kRecV = kBlockH * kSubW;             // 1024 doubles: swizzled reflector panel
constexpr int kLdT = 8;                            // leading dim of the T images (conflict-free LDS.128)
constexpr int kRecT = kSubW * kLdT;                // 64
constexpr int kRecDoubles = kRecV + 2 * kRecT;     // V | T | T^T  = 1152
constexpr int kRecBytes = kRecDoubles * 8;         // 9 216

// Swizzled position of element (row r, column c) of a 128 x 8 reflector panel.  16-byte slots
// hold row pairs; slot = c*64 + ((r/2) ^ h(c)).  h is chosen so that both access patterns of
// the chain are bank-conflict free:
//   GEMM1 (LDS.128): lanes (g, tg) read rows (8i+2tg, 8i+2tg+1) of column g
//   GEMM3 (LDS.64) : lanes (g, tg) read row 8i+g of columns 2tg and 2tg+1
__host__ __device__ __forceinline__ int vswz(int r, int c) {
    const int h = ((c & 1) << 2) ^ (((c >> 1) & 3) << 1);
    return (((c << 6) + ((r >> 1) ^ h)) << 1) + (r & 1);
}

// ---- mbarrier / TMA bulk copy -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- the register-resident block-reflector chain ------------------------------------------------
// bx[i], by[i]: B^T fragments of 8-row tile i (i = 0..15): column g of the warp's 8 columns,
// rows 8i + 2tg and 8i + 2tg + 1.
// sx, sy     : in  S0^T fragment (R strip / Y rows / zero): column g, reflectors 2tg, 2tg+1 (added AFTER
//              the B^T V product, so the load that produced it is off the critical path)
//              out S2^T fragment in the same layout.
// FORWARD = true : S2 = T^T S (factorisation, Q^T applied);  false: S2 = T S (Q applied).
template <bool FORWARD>
__device__ __forceinline__ void wy_chain(double (&bx)[16], double (&by)[16], double &sx, double &sy,
                                         const double *rec, int lane) {
    const int g = lane >> 2, tg = lane & 3;
    // ---- GEMM1: S^T += B^T V (four independent accumulator sets) ----------------------------------------
    {
        const int h = ((g & 1) << 2) ^ (((g >> 1) & 3) << 1);
        const int hlo = h & 3, hhi = (h >> 2) & 1;
        const double *base = rec + 2 * ((g << 6) + (tg ^ hlo));
        const double *pe = base + 8 * hhi, *po = base - 8 * hhi;
        double ax[4] = {0.0, 0.0, 0.0, 0.0}, ay[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const double *q = ((i & 1) ? po : pe) + 8 * i;
            const double2 v = *reinterpret_cast<const double2 *>(q);
            dmma_884(ax[i & 3], ay[i & 3], bx[i], v.x);
            dmma_884(ax[i & 3], ay[i & 3], by[i], v.y);
        }
        sx += (ax[0] + ax[1]) + (ax[2] + ax[3]);
        sy += (ay[0] + ay[1]) + (ay[2] + ay[3]);
    }
    // ---- GEMM2: S2^T = S^T Tm,  Tm = T (forward) or T^T ----------------------------------------------
    double px = 0.0, py = 0.0;
    {
        const double *Tm = rec + kRecV + (FORWARD ? 0 : kRecT);
        const double2 tv = *reinterpret_cast<const double2 *>(Tm + 2 * tg + kLdT * g);
        dmma_884(px, py, sx, tv.x);
        dmma_884(px, py, sy, tv.y);
    }
    // ---- GEMM3: B^T -= S2^T V^T ----------------------------------------------------------------------
    {
        const int gh = g >> 1;
        const int hb = tg >> 1;
        const double *base = rec + 2 * (((2 * tg) << 6) + (gh ^ ((tg & 1) << 1))) + (g & 1);
        const double *pe_e = base + 8 * hb, *po_e = base - 8 * hb;                  // even column 2tg
        const double *pe_o = base + 128 + 8 * (hb ^ 1), *po_o = base + 128 - 8 * (hb ^ 1);   // odd column 2tg+1
        const double nx = -px, ny = -py;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const double *qe = ((i & 1) ? po_e : pe_e) + 8 * i;
            const double *qo = ((i & 1) ? po_o : pe_o) + 8 * i;
            dmma_884(bx[i], by[i], nx, qe[0]);
            dmma_884(bx[i], by[i], ny, qo[0]);
        }
    }
    sx = px; sy = py;
}

// ========================================================================================
// argument blocks
// ========================================================================================
struct TlrArgs {
    const double *const *AU, *const *AV, *const *BU, *const *BV;
    double *const *CU, *const *CV;
    const int *Ark, *Brk;
    int *Crk;
    int *status;          // user status array (may be null)
    int nb, lda, ldb, ldc;
    double alpha, beta;
    int rk, maxrk;
    double acc;
    int tile0;            // first tile of this chunk in the user arrays
    int ntiles;           // tiles in this chunk
    int f32;              // 1: the user arrays hold float (HCORE_sgemm); arithmetic stays FP64
    const double *const *Bd;   // dense B tiles (nb x nb, ld ldbd) of the B-dense variant, or null
    int ldbd;
    int *rank_out;        // where the new rank goes: Crk (fused call, phase 2) or the caller's new-rank array (phase 1)
    int scan_from;        // first index the accuracy truncation scan looks at: 1 (HCORE_dgemm, reference hcore_dgemm.c:454) or
                          //    2 (HCORE_dgemm_fast, reference hcore_dgemm_fast.c:314)
    int phase;            // 0: whole update; 1: stack + QR + SVD only (HCORE_dgemm_qr_svd); 2: reflector application
                          //    only, into possibly re-allocated CU/CV (HCORE_dgemm_ormqr)
};

// User-array element access: the factor arrays are double, or float when ar.f32 is set (the
// pointers are then float arrays behind a double* type).  `idx` is an ELEMENT index.
__device__ __forceinline__ double ld_user(const double *base, size_t idx, int f32) {
    return f32 ? (double)__ldg(reinterpret_cast<const float *>(base) + idx) : __ldg(base + idx);
}
__device__ __forceinline__ void st_user(double *base, size_t idx, double v, int f32) {
    if (f32) reinterpret_cast<float *>(base)[idx] = (float)v;
    else base[idx] = v;
}

// Per-chunk scratch owned by the context.  Index [panel][tile] with panel 0 = U, 1 = V.
struct TlrWorkspace {
    double *rec;     // [2][chunk][nblk][N/8][kRecDoubles]  block-reflector records
    double *G;       // [chunk][m_pad * N]                   G = BU (AV^T BV)^T, ld m_pad
    double *R;       // [2][chunk][N * N]
    double *Y;       // [2][chunk][N * N]            U_r (panel 0) and W (panel 1)
    double *L;       // [chunk][N * (N + 4)]         core matrix of the SVD kernel, N = 256 only (else in shared memory)
    double *sigma;   // [chunk][N]                   sorted singular values
    int *ncols;      // [chunk]  a + c
    int *newrank;    // [chunk]
    int *st;         // [chunk]  internal status
    int *sweeps;     // [chunk]  Jacobi sweeps used (diagnostics)
    int chunk, m_pad, nblk;
};

__device__ __forceinline__ size_t nn_offset(const TlrWorkspace &ws, int panel, int t, int n) {
    return ((size_t)panel * ws.chunk + t) * (size_t)n * n;
}
__device__ __forceinline__ size_t rec_offset(const TlrWorkspace &ws, int panel, int t, int nsp_full) {
    return ((size_t)panel * ws.chunk + t) * (size_t)ws.nblk * nsp_full * kRecDoubles;
}

// ========================================================================================
// 1. P / G kernel
// ========================================================================================
// One CTA (256 threads) per tile.  Row chunks of 32 rows of AV, BV (then BU) are staged in shared
// memory with cp.async (16-byte, coalesced along the rows), double buffered.  Ranks above 64 are
// handled in 64-column slabs (one extra sweep per slab pair).
constexpr int kPgRows = 32;
constexpr int kPgLd = kPgRows + 4;      // 36: conflict-free fragment loads
// slab width KC: 64, or 32 for the narrow stacks (N <= 64: 46 KB instead of 108 KB of shared memory, four
// CTAs per SM instead of two -- the kernel is latency-bound there)

template <int KC>
struct PgSmemT {
    static constexpr int kLdP = KC + 4;
    static constexpr int kStage = 2 * KC * kPgLd;            // one stage: two 32 x KC operand chunks
    static constexpr int kP = 2 * kStage;                    // P slab, ld KC + 4
    static constexpr int kTotal = kP + KC * kLdP;
    static constexpr size_t kBytes = (size_t)kTotal * sizeof(double);
};
using PgSmem = PgSmemT<64>;

// rows [r0, r0+32) x ncol columns (padded with zeros to ncol8) of a column-major matrix -> dst (ld kPgLd)
__device__ __forceinline__ void pg_stage(double *dst, const double *src, size_t col0, int ld, int m, int r0, int ncol,
                                         int ncol8, bool vec_ok, int f32, int tid) {
    if (!f32) src += col0 * ld;
    if (vec_ok && r0 + kPgRows <= m) {
        for (int idx = tid; idx < (kPgRows / 2) * ncol; idx += 256) {
            const int rp = idx & 15, c = idx >> 4;
            cp_async16(dst + 2 * rp + c * kPgLd, src + (size_t)c * ld + r0 + 2 * rp);
        }
    } else {
        for (int idx = tid; idx < kPgRows * ncol; idx += 256) {
            const int r = idx & 31, c = idx >> 5;
            dst[r + c * kPgLd] = (r0 + r >= m) ? 0.0
                                 : (f32 ? ld_user(src, (col0 + c) * ld + r0 + r, 1) : __ldg(src + (size_t)c * ld + r0 + r));
        }
    }
    for (int idx = tid; idx < kPgRows * (ncol8 - ncol); idx += 256) {
        const int r = idx & 31, c = ncol + (idx >> 5);
        dst[r + c * kPgLd] = 0.0;
    }
}

// G(m_pad x a, ld m_pad) = pscale * BU (AV^T BV)^T for one tile, by the whole CTA (256 threads).
// AV: m x a (ld lda), BU, BV: m x b (ld ldb); rows m .. m_pad-1 of G are written as zeros.
// mc: contraction length of the P product when it differs from the row count m of BU / G (HCORE_dsyrk contracts over
// its leading dimension, reference src/compute/hcore_dsyrk.c:57,:103); -1 = m.
template <int KC = 64>
__device__ __forceinline__ void pg_body(double *smem, const double *AV, const double *BU, const double *BV, int a, int b,
                                        int m, int lda, int ldb, int m_pad, int f32, double pscale, double *Gt, int mc = -1) {
    using SM = PgSmemT<KC>;
    constexpr int kPgKC = KC, kPgLdP = SM::kLdP;
    double *P = smem + SM::kP;
    const int tid = threadIdx.x, warp = warp_id(), lane = lane_id();
    const int g = lane >> 2, tg = lane & 3;
    const int mp = mc < 0 ? m : mc;                                  // rows contracted by the P product
    const int nchunks = (mp + kPgRows - 1) / kPgRows, nchunks_pad = m_pad / kPgRows;
    const bool va = !f32 && ((reinterpret_cast<uintptr_t>(AV) & 15) == 0) && ((lda & 1) == 0);
    const bool vb = !f32 && ((reinterpret_cast<uintptr_t>(BV) & 15) == 0) && ((reinterpret_cast<uintptr_t>(BU) & 15) == 0) &&
                    ((ldb & 1) == 0);

    for (int ia = 0; ia < a; ia += kPgKC) {
        const int as = (a - ia < kPgKC) ? (a - ia) : kPgKC, as8 = round_up(as, 8);
        for (int jb = 0; jb < b; jb += kPgKC) {
            const int bs = (b - jb < kPgKC) ? (b - jb) : kPgKC, bs8 = round_up(bs, 8);
            const double *AVs = AV, *BVs = BV, *BUs = BU;      // slab offsets are passed as column indices
            // ---- P slab = AV[:, ia..]^T BV[:, jb..]: warp -> 8-row strip of P, all column tiles -----
            {
                const bool active = 8 * warp < as8;
                const int i0 = 8 * warp;
                double acc[KC / 8][2];
#pragma unroll
                for (int q = 0; q < KC / 8; ++q) acc[q][0] = acc[q][1] = 0.0;
                pg_stage(smem, AVs, ia, lda, mp, 0, as, as8, va, f32, tid);
                pg_stage(smem + kPgKC * kPgLd, BVs, jb, ldb, mp, 0, bs, bs8, vb, f32, tid);
                cp_async_commit();
                for (int ch = 0; ch < nchunks; ++ch) {
                    const double *cur = smem + (ch & 1) * SM::kStage;
                    if (ch + 1 < nchunks) {
                        double *nxt = smem + ((ch + 1) & 1) * SM::kStage;
                        pg_stage(nxt, AVs, ia, lda, mp, (ch + 1) * kPgRows, as, as8, va, f32, tid);
                        pg_stage(nxt + kPgKC * kPgLd, BVs, jb, ldb, mp, (ch + 1) * kPgRows, bs, bs8, vb, f32, tid);
                        cp_async_commit();
                        cp_async_wait<1>();
                    } else {
                        cp_async_wait<0>();
                    }
                    __syncthreads();
                    if (active) {
                        const double *As = cur, *Bs = cur + kPgKC * kPgLd;
#pragma unroll
                        for (int k = 0; k < kPgRows; k += 4) {
                            const double av = As[(k + tg) + (i0 + g) * kPgLd];
#pragma unroll
                            for (int q = 0; q < KC / 8; ++q) {
                                if (8 * q >= bs8) break;   // a real exit: predicated-off DMMAs still occupy the FP64 pipe
                                const double bv = Bs[(k + tg) + (8 * q + g) * kPgLd];
                                dmma_884(acc[q][0], acc[q][1], av, bv);
                            }
                        }
                    }
                    __syncthreads();
                }
                if (active) {
#pragma unroll
                    for (int q = 0; q < KC / 8; ++q) {
                        if (8 * q < bs8) {
                            P[(i0 + g) + (8 * q + 2 * tg) * kPgLdP] = pscale * acc[q][0];
                            P[(i0 + g) + (8 * q + 2 * tg + 1) * kPgLdP] = pscale * acc[q][1];
                        }
                    }
                }
            }
            __syncthreads();
            // ---- G[:, ia..] (+)= BU[:, jb..] P^T: warp -> (8-row strip of the chunk, 4 column tiles) ------
            {
                constexpr int QW = KC / 16;                   // column tiles per warp: 8 warps = 4 strips x 2 halves
                const int strip = warp & 3, q0 = QW * (warp >> 2);
                pg_stage(smem, BUs, jb, ldb, m, 0, bs, bs8, vb, f32, tid);
                cp_async_commit();
                for (int ch = 0; ch < nchunks_pad; ++ch) {
                    const double *cur = smem + (ch & 1) * SM::kStage;
                    if (ch + 1 < nchunks_pad) {
                        pg_stage(smem + ((ch + 1) & 1) * SM::kStage, BUs, jb, ldb, m, (ch + 1) * kPgRows, bs, bs8, vb, f32, tid);
                        cp_async_commit();
                        cp_async_wait<1>();
                    } else {
                        cp_async_wait<0>();
                    }
                    __syncthreads();
                    double acc[QW][2];
#pragma unroll
                    for (int q = 0; q < QW; ++q) acc[q][0] = acc[q][1] = 0.0;
#pragma unroll 4
                    for (int k = 0; k < bs8; k += 4) {
                        const double av = cur[(8 * strip + g) + (k + tg) * kPgLd];
#pragma unroll
                        for (int q = 0; q < QW; ++q) {
                            if (8 * (q0 + q) >= as8) break;
                            const double bv = P[(8 * (q0 + q) + g) + (k + tg) * kPgLdP];
                            dmma_884(acc[q][0], acc[q][1], av, bv);
                        }
                    }
                    const int row = ch * kPgRows + 8 * strip + g;
#pragma unroll
                    for (int q = 0; q < QW; ++q) {
                        const int col = 8 * (q0 + q) + 2 * tg;
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            if (col + e < as) {
                                double *dst = Gt + row + (size_t)(ia + col + e) * m_pad;
                                *dst = (jb == 0) ? acc[q][e] : *dst + acc[q][e];
                            }
                        }
                    }
                    __syncthreads();
                }
            }
        }
    }
}


template <int N>
struct PgWidth { static constexpr int value = (N <= 64) ? 32 : 64; };

template <int N>
__global__ void __launch_bounds__(256, (N <= 64 ? 4 : 2)) tlr_pg_kernel(TlrArgs ar, TlrWorkspace ws) {
    extern __shared__ __align__(16) double smem[];
    const int t = blockIdx.x, gt = ar.tile0 + t;
    const int a = ar.Ark[gt], b = ar.Brk[gt], c = ar.Crk[gt];
    const int n = a + c, m = ar.nb;

    int st = HCORE_B200_OK;
    if (a <= 0 || b <= 0 || c <= 0) st = HCORE_B200_ERR_ZERO_RANK;
    else if (ar.maxrk < n) st = HCORE_B200_ERR_MAXRK;
    else if (n > N || n > m || b > N) st = HCORE_B200_ERR_UNSUPPORTED;
    if (threadIdx.x == 0) {
        ws.st[t] = st;
        ws.ncols[t] = n;
        ws.newrank[t] = c;
        if (ar.status) ar.status[gt] = st;
    }
    if (st != HCORE_B200_OK) return;
    pg_body<PgWidth<N>::value>(smem, ar.AV[gt], ar.BU[gt], ar.BV[gt], a, b, m, ar.lda, ar.ldb, ws.m_pad, ar.f32, 1.0,
                               ws.G + (size_t)t * ws.m_pad * N);
}

// ---- 1b. G for a DENSE B operand -----------------------------------------------------------------
// B-dense variant of the update (reference __qrb_b_dense, src/compute/hcore_dgemm.c:236-288): the tile B is a
// dense nb x nb matrix, C <- alpha (AU AV^T) B^T + beta C, and the stacked V panel is [CV | G] with
// G = B AV (nb x a).  grid = (128-row blocks, tiles), 256 threads: warp w owns rows 16 w .. 16 w + 15 of the
// block and all a columns; 32-wide contraction chunks of B (128 x 32) and AV (32 x a) are staged with
// cp.async, double buffered.
template <int N>
struct GdSmem {
    static constexpr int kLdB = kBlockH + 4;                     // 132
    static constexpr int kLdA = kPgRows + 4;                     // 36
    static constexpr int kStage = kPgRows * kLdB + N * kLdA;     // B chunk | AV chunk
    static constexpr size_t kBytes = 2ull * kStage * sizeof(double);
};

template <int N>
__global__ void __launch_bounds__(256, 1) tlr_gdense_kernel(TlrArgs ar, TlrWorkspace ws) {
    using SM = GdSmem<N>;
    extern __shared__ __align__(16) double smem[];
    const int t = blockIdx.y, gt = ar.tile0 + t, i0 = blockIdx.x * kBlockH;
    const int tid = threadIdx.x, warp = warp_id(), lane = lane_id();
    const int g = lane >> 2, tg = lane & 3;
    const int a = ar.Ark[gt], c = ar.Crk[gt];
    const int n = a + c, m = ar.nb;
    int st = HCORE_B200_OK;
    if (a <= 0 || c <= 0) st = HCORE_B200_ERR_ZERO_RANK;
    else if (ar.maxrk < n) st = HCORE_B200_ERR_MAXRK;
    else if (n > N || n > m) st = HCORE_B200_ERR_UNSUPPORTED;
    if (blockIdx.x == 0 && tid == 0) {
        ws.st[t] = st;
        ws.ncols[t] = n;
        ws.newrank[t] = c;
        if (ar.status) ar.status[gt] = st;
    }
    if (st != HCORE_B200_OK) return;
    const double *B = ar.Bd[gt], *AV = ar.AV[gt];
    const int ldb = ar.ldbd, lda = ar.lda, m_pad = ws.m_pad;
    double *Gt = ws.G + (size_t)t * m_pad * N;
    const int a8 = round_up(a, 8);
    const bool vec = ((reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(AV)) & 15) == 0 && (ldb & 1) == 0 &&
                     (lda & 1) == 0 && !ar.f32;
    auto stage = [&](int k0, int buf) {
        double *Bs = smem + buf * SM::kStage, *As = Bs + kPgRows * SM::kLdB;
        const bool full = vec && i0 + kBlockH <= m && k0 + kPgRows <= m;
        if (full) {
            for (int idx = tid; idx < (kBlockH / 2) * kPgRows; idx += 256) {        // B(i0 + r, k0 + kk): rows contiguous
                const int rp = idx & 63, kk = idx >> 6;
                cp_async16(Bs + 2 * rp + kk * SM::kLdB, B + (size_t)(k0 + kk) * ldb + i0 + 2 * rp);
            }
            for (int idx = tid; idx < (kPgRows / 2) * a; idx += 256) {              // AV(k0 + kk, j): k contiguous
                const int kp = idx & 15, j = idx >> 4;
                cp_async16(As + 2 * kp + j * SM::kLdA, AV + (size_t)j * lda + k0 + 2 * kp);
            }
        } else {
            for (int idx = tid; idx < kBlockH * kPgRows; idx += 256) {
                const int r = idx & 127, kk = idx >> 7;
                Bs[r + kk * SM::kLdB] = (i0 + r < m && k0 + kk < m) ? ld_user(B, (size_t)(k0 + kk) * ldb + i0 + r, ar.f32) : 0.0;
            }
            for (int idx = tid; idx < kPgRows * a; idx += 256) {
                const int kk = idx & 31, j = idx >> 5;
                As[kk + j * SM::kLdA] = (k0 + kk < m) ? ld_user(AV, (size_t)j * lda + k0 + kk, ar.f32) : 0.0;
            }
        }
        for (int idx = tid; idx < kPgRows * (a8 - a); idx += 256) As[(idx & 31) + (a + (idx >> 5)) * SM::kLdA] = 0.0;
        cp_async_commit();
    };
    // column tiles are processed in passes of at most 16 (128 columns): N = 256 re-streams B once when a > 128
    constexpr int CT = (N / 8 > 16) ? 16 : N / 8;
    for (int ct0 = 0; 8 * ct0 < a8; ct0 += CT) {
    double acc[2][CT][2];
#pragma unroll
    for (int sidx = 0; sidx < 2; ++sidx)
#pragma unroll
        for (int tt = 0; tt < CT; ++tt) acc[sidx][tt][0] = acc[sidx][tt][1] = 0.0;
    const int nk = (m + kPgRows - 1) / kPgRows;
    __syncthreads();
    stage(0, 0);
    for (int ch = 0; ch < nk; ++ch) {
        if (ch + 1 < nk) {
            stage((ch + 1) * kPgRows, (ch + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const double *Bs = smem + (ch & 1) * SM::kStage, *As = Bs + kPgRows * SM::kLdB;
        const double *bp = Bs + 16 * warp + g + tg * SM::kLdB;
        const double *ap = As + tg + g * SM::kLdA;
#pragma unroll
        for (int k = 0; k < kPgRows; k += 4) {
            const double b0 = bp[k * SM::kLdB], b1 = bp[k * SM::kLdB + 8];
#pragma unroll
            for (int tt = 0; tt < CT; ++tt) {
                if (8 * (ct0 + tt) >= a8) break;
                const double av = ap[k + 8 * (ct0 + tt) * SM::kLdA];
                dmma_884(acc[0][tt][0], acc[0][tt][1], b0, av);
                dmma_884(acc[1][tt][0], acc[1][tt][1], b1, av);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int sidx = 0; sidx < 2; ++sidx) {
        const int row = i0 + 16 * warp + 8 * sidx + g;       // < m_pad (i0 + 127 < m_pad)
#pragma unroll
        for (int tt = 0; tt < CT; ++tt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = 8 * (ct0 + tt) + 2 * tg + e;
                if (col < a) Gt[row + (size_t)col * m_pad] = (row < m) ? acc[sidx][tt][e] : 0.0;
            }
        }
    }
    }   // column passes
}

// ========================================================================================
// 2. tall-skinny QR kernel
// ========================================================================================
// Flat-tree TSQR over 128-row blocks: block 0 is a plain blocked Householder QR (reflectors below
// the diagonal, R on top); every further block B_j is eliminated against the running R
// ("triangle on top of a dense block": reflector i = [e_i ; v_i]).  8-column sub-panels with
// compact-WY T factors.  Columns are processed LEFT-LOOKING in groups of GW = 8 * NW columns
// (NW warps per CTA, 8 columns per warp, register resident).  For every row block
//   (1) the CTA streams the records of all earlier groups (TMA, double buffered) through wy_chain;
//   (2) inside the group the warps form a pipeline without CTA barriers: warp w applies the
//       records of warps 0..w-1 (read from per-warp shared-memory slots, handed over with named
//       barriers), then factors its OWN 128 x 8 sub-panel alone, in registers (warp_factor), and
//       publishes the record to its slot and to global memory.
constexpr int kYbuf = kBlockH + 16;    // per warp, per parity: y[128] | ydiag[8] | z[8]

template <int NW>
struct TsqrSmem {
    static constexpr int kRec0 = 0;
    static constexpr int kRec1 = kRec0 + kRecDoubles;
    static constexpr int kOwn = kRec1 + kRecDoubles;                 // [NW-1] records of the group's own sub-panels
    static constexpr int kY = kOwn + (NW - 1) * kRecDoubles;         // [NW][2][kYbuf]
    static constexpr int kBar = kY + NW * 2 * kYbuf;                 // 2 mbarriers
    static constexpr int kTotal = kBar + 2;
    static constexpr size_t kBytes = (size_t)kTotal * sizeof(double);
};

// Source of one stacked-panel column: array, element offset of its first row, the scalar applied
// on load and whether the array holds floats.
struct ColSrc {
    const double *p;
    size_t off;
    double scale;
    int f32;
};
__device__ __forceinline__ ColSrc panel_column(const TlrArgs &ar, const double *Gt, int m_pad, int gt, int pnl,
                                               int col, int c, int n) {
    ColSrc s{nullptr, 0, 0.0, 0};
    if (col < c) {
        s.p = (pnl == 0 ? ar.CU[gt] : ar.CV[gt]);
        s.off = (size_t)col * ar.ldc;
        s.scale = pnl == 0 ? ar.beta : 1.0;
        s.f32 = ar.f32;
    } else if (col < n) {
        if (pnl == 0) { s.p = ar.AU[gt]; s.off = (size_t)(col - c) * ar.lda; s.scale = ar.alpha; s.f32 = ar.f32; }
        else { s.p = Gt; s.off = (size_t)(col - c) * m_pad; s.scale = 1.0; }
    }
    return s;
}

// Warp all-reduce (sum) of one double per lane through the FP64 tensor core: two dependent
// DMMAs with a matrix of ones (~70 cycles) instead of five shuffle rounds (~175-250 cycles).
//   stage 1: B[k = tg][n = g] = v(lane)  ->  D[.][n] = column sums cs(n) = sum_tg v(4n + tg);
//            lane (g, tg) receives cs(2tg), cs(2tg+1)
//   stage 2: B[k = tg][n = g] = cs(2tg) + cs(2tg+1)  ->  D = total, in every lane.
template <int NV>
__device__ __forceinline__ void warp_allsum_dmma(double (&v)[NV]) {
    double d0[NV], d1[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) { d0[q] = 0.0; d1[q] = 0.0; dmma_884(d0[q], d1[q], 1.0, v[q]); }
#pragma unroll
    for (int q = 0; q < NV; ++q) { const double p = d0[q] + d1[q]; d0[q] = 0.0; d1[q] = 0.0; dmma_884(d0[q], d1[q], 1.0, p); }
#pragma unroll
    for (int q = 0; q < NV; ++q) v[q] = d0[q];
}

// Named barriers with LITERAL ids: ptxas reserves (highest id + 1) barriers per CTA out of the 32 an
// SM has, so a register-held id (which makes it reserve all 16) would cap residency at 2 CTAs/SM.
// Call these from fully unrolled code so that `id` folds to a constant.
#define HCORE_BAR_CASE(op, k) case k: asm volatile(op " " #k ", %0;" ::"r"(nthreads) : "memory"); break;
#define HCORE_BAR_SWITCH(op)                                                                          \
    switch (id) {                                                                                     \
        HCORE_BAR_CASE(op, 1) HCORE_BAR_CASE(op, 2) HCORE_BAR_CASE(op, 3) HCORE_BAR_CASE(op, 4)       \
        HCORE_BAR_CASE(op, 5) HCORE_BAR_CASE(op, 6) HCORE_BAR_CASE(op, 7) HCORE_BAR_CASE(op, 8)       \
        HCORE_BAR_CASE(op, 9) HCORE_BAR_CASE(op, 10) HCORE_BAR_CASE(op, 11) HCORE_BAR_CASE(op, 12)    \
        HCORE_BAR_CASE(op, 13) HCORE_BAR_CASE(op, 14) HCORE_BAR_CASE(op, 15)                          \
        default: break;                                                                               \
    }
__device__ __forceinline__ void bar_sync_lit(int id, int nthreads) { HCORE_BAR_SWITCH("bar.sync") }
__device__ __forceinline__ void bar_arrive_lit(int id, int nthreads) { HCORE_BAR_SWITCH("bar.arrive") }   // does not wait

// Householder factorisation of the warp's own 128 x 8 sub-panel (columns c0 .. c0+7 of the row
// block), entirely inside the warp: no CTA barrier, one __syncwarp per Householder step.
//
//   bx/by  the 8 columns in fragment layout (lane (g, tg): column g, rows 8i+2tg, 8i+2tg+1)
//   rx/ry  the 8 x 8 "diagonal tile" in the same layout (rows c0+2tg, c0+2tg+1 of column c0+g):
//          blocks > 0: the diagonal block of the running R (reflector j = [e_j ; v_j]);
//          block 0   : rows c0..c0+7 of the block itself (tile i* = c0/8, taken out of bx/by, whose
//                      copy of that tile is dead afterwards); the Householder vectors extend into it.
// Step j:  the owner lanes (g == j) publish the raw tail y of column j (warp-private shared memory);
//          every lane forms its partial d = y^T x_g, one DMMA sums the four partials of a column
//          (A = partials, B = ones), a second one broadcasts r_g = row c0+j of column g; d_j and
//          alpha = r_j reach all lanes by one shuffle each; all lanes derive (beta, tau, scal);
//          lanes g > j update x_g -= y * scal * w_g,  w_g = tau (r_g + scal d_g).
// The owner's column is NOT rescaled in registers (v_j = scal_j * y_j is only formed when the record
// is written), so z_cj = v_c^T v_j = scal_c (r_c + scal_j d_c) for c < j; these feed the forward
// recurrence of the compact-WY factor, T[r][j] = -tau_j sum_{q=r}^{j-1} T[r][q] z_qj, one column
// behind the Householder steps (lane group r holds row r of T).
template <int N, bool BLK0>
__device__ __forceinline__ void warp_factor(double (&bx)[16], double (&by)[16], double *yb, double *slot,
                                            double *recg, double *Rg, int c0, int lane) {
    constexpr unsigned kFull = 0xffffffffu;
    const int g = lane >> 2, tg = lane & 3;
    if (BLK0 && N > kBlockH && c0 >= kBlockH) {
        // Stack wider than the row block (N = 256): these columns have no diagonal inside row block 0.  Rows
        // kBlockH .. n-1 of R are "virtual" zero rows appended to the panel (they change neither R nor the
        // physical rows of Q); the sub-panel gets the identity reflector here, its 128 entries ARE rows
        // 0 .. 127 of R, and its reflectors are born in the next row blocks (triangle-on-top-of-block with a
        // zero diagonal to start from).
        const int hsw0 = ((g & 1) << 2) ^ (((g >> 1) & 3) << 1);
        const int col0 = c0 + g;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int off = 2 * ((g << 6) + ((4 * i + tg) ^ hsw0));
            if (slot) *reinterpret_cast<double2 *>(slot + off) = make_double2(0.0, 0.0);
            *reinterpret_cast<double2 *>(recg + off) = make_double2(0.0, 0.0);
            *reinterpret_cast<double2 *>(Rg + (size_t)(8 * i + 2 * tg) + (size_t)col0 * N) = make_double2(bx[i], by[i]);
        }
        const int o0 = kRecV + g + kLdT * (2 * tg), o1 = kRecV + kRecT + 2 * tg + kLdT * g;
        if (slot) { slot[o0] = 0.0; slot[o0 + kLdT] = 0.0; *reinterpret_cast<double2 *>(slot + o1) = make_double2(0.0, 0.0); }
        recg[o0] = 0.0; recg[o0 + kLdT] = 0.0;
        *reinterpret_cast<double2 *>(recg + o1) = make_double2(0.0, 0.0);
        return;
    }
    const int istar = c0 >> 3;
    double rx, ry;
    if (BLK0) {
        rx = bx[0]; ry = by[0];
#pragma unroll
        for (int i = 1; i < 16; ++i) { rx = (i == istar) ? bx[i] : rx; ry = (i == istar) ? by[i] : ry; }
    } else {
        const double2 rd = *reinterpret_cast<const double2 *>(Rg + (size_t)(c0 + 2 * tg) + (size_t)(c0 + g) * N);
        rx = rd.x; ry = rd.y;
    }
    double trow[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) trow[q] = 0.0;
    double myscal = 0.0, zpub = 0.0, tau_prev = 0.0;
#ifdef HCORE_B200_WFPROBE
    PROF_DECL(threadIdx.x == 0)
#define WF_ADD(slot) PROF_ADD(slot)
#else
#define WF_ADD(slot)
#endif

#pragma unroll
    for (int j = 0; j <= 8; ++j) {
        double *yc = yb + (j & 1) * kYbuf;
        WF_ADD(40);
        // ---- publish: raw tail of column j, z values of step j-1 ----------------------------------
        if (j < 8 && g == j) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                double y0 = bx[i], y1 = by[i];
                if (BLK0) { y0 = (i > istar) ? y0 : 0.0; y1 = (i > istar) ? y1 : 0.0; }
                *reinterpret_cast<double2 *>(yc + 8 * i + 2 * tg) = make_double2(y0, y1);
            }
            if (BLK0)
                *reinterpret_cast<double2 *>(yc + kBlockH + 2 * tg) =
                    make_double2((2 * tg > j) ? rx : 0.0, (2 * tg + 1 > j) ? ry : 0.0);
        }
        if (j >= 2 && g < j - 1 && tg == 0) yc[kBlockH + 8 + g] = zpub;
        __syncwarp();
        WF_ADD(41);
        // ---- column j-1 of T --------------------------------------------------------------------------
        if (j >= 1) {
            const int jc = j - 1;
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int q = 0; q < jc; ++q) {
                const double zq = yc[kBlockH + 8 + q];
                if (q & 1) t1 = fma(trow[q], zq, t1);
                else t0 = fma(trow[q], zq, t0);
            }
            trow[jc] = (g < jc) ? -tau_prev * (t0 + t1) : (g == jc ? tau_prev : 0.0);
        }
        if (j == 8) break;
        WF_ADD(42);
        // ---- d_g = y^T x_g ----------------------------------------------------------------------------
        double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const double2 y = *reinterpret_cast<const double2 *>(yc + 8 * i + 2 * tg);
            if (i & 1) { p2 = fma(bx[i], y.x, p2); p3 = fma(by[i], y.y, p3); }
            else { p0 = fma(bx[i], y.x, p0); p1 = fma(by[i], y.y, p1); }
        }
        if (BLK0) {
            const double2 yd = *reinterpret_cast<const double2 *>(yc + kBlockH + 2 * tg);
            p0 = fma(rx, yd.x, p0);
            p1 = fma(ry, yd.y, p1);
        }
        WF_ADD(43);
        double d0 = 0.0, d1 = 0.0, r0 = 0.0, r1 = 0.0;
        dmma_884(d0, d1, (p0 + p1) + (p2 + p3), 1.0);                            // d_g in all four lanes of the group
        dmma_884(r0, r1, (tg == (j >> 1)) ? ((j & 1) ? ry : rx) : 0.0, 1.0);     // r_g = tile row j of column g
        const double dj = __shfl_sync(kFull, d0, 4 * j), alpha = __shfl_sync(kFull, r0, 4 * j);
        const Householder h = make_householder(alpha, dj);
        const double zc = fma(h.scal, d0, r0);
        WF_ADD(44);
        if (g > j) {
            const double wv = h.tau * zc, f = h.scal * wv;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const double2 y = *reinterpret_cast<const double2 *>(yc + 8 * i + 2 * tg);
                bx[i] = fma(-y.x, f, bx[i]);
                by[i] = fma(-y.y, f, by[i]);
            }
            if (BLK0) {
                const double2 yd = *reinterpret_cast<const double2 *>(yc + kBlockH + 2 * tg);
                rx = fma(-yd.x, f, rx);
                ry = fma(-yd.y, f, ry);
            }
            if (tg == (j >> 1)) { if (j & 1) ry -= wv; else rx -= wv; }
        } else if (g == j) {
            myscal = h.scal;
            if (tg == (j >> 1)) { if (j & 1) ry = h.beta; else rx = h.beta; }
        } else {
            zpub = myscal * zc;
        }
        tau_prev = h.tau;
        WF_ADD(45);
    }
    // ---- outputs: explicit reflector panel (swizzled) + T images -> slot and global record; R entries ---
    {
        const int hsw = ((g & 1) << 2) ^ (((g >> 1) & 3) << 1);
        const int col = c0 + g;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            double v0, v1;
            if (BLK0) {
                const int row0 = 8 * i + 2 * tg;
                const double s0 = (i == istar) ? rx : bx[i], s1 = (i == istar) ? ry : by[i];
                v0 = (row0 > col) ? s0 * myscal : (row0 == col ? 1.0 : 0.0);
                v1 = (row0 + 1 > col) ? s1 * myscal : (row0 + 1 == col ? 1.0 : 0.0);
                if (row0 < N)
                    *reinterpret_cast<double2 *>(Rg + (size_t)row0 + (size_t)col * N) =
                        make_double2((row0 <= col) ? s0 : 0.0, (row0 + 1 <= col) ? s1 : 0.0);
            } else {
                v0 = bx[i] * myscal;
                v1 = by[i] * myscal;
            }
            const int off = 2 * ((g << 6) + ((4 * i + tg) ^ hsw));
            if (slot) *reinterpret_cast<double2 *>(slot + off) = make_double2(v0, v1);
            *reinterpret_cast<double2 *>(recg + off) = make_double2(v0, v1);
        }
        if (!BLK0)
            *reinterpret_cast<double2 *>(Rg + (size_t)(c0 + 2 * tg) + (size_t)col * N) = make_double2(rx, ry);
        // T[g][2tg], T[g][2tg+1]
        const double ta = (tg == 0) ? trow[0] : (tg == 1) ? trow[2] : (tg == 2) ? trow[4] : trow[6];
        const double tb = (tg == 0) ? trow[1] : (tg == 1) ? trow[3] : (tg == 2) ? trow[5] : trow[7];
        const int o0 = kRecV + g + kLdT * (2 * tg);              // T   : B[k][n] = T[k][n] at k + 8 n
        const int o1 = kRecV + kRecT + 2 * tg + kLdT * g;        // T^T : B[k][n] = T[n][k] at k + 8 n
        if (slot) {
            slot[o0] = ta;
            slot[o0 + kLdT] = tb;
            *reinterpret_cast<double2 *>(slot + o1) = make_double2(ta, tb);
        }
        recg[o0] = ta;
        recg[o0 + kLdT] = tb;
        *reinterpret_cast<double2 *>(recg + o1) = make_double2(ta, tb);
    }
    WF_ADD(46);
}

// One CTA (NW warps) per panel: blockIdx.x = 2 * tile + panel.
// NW = 2: shared memory admits six CTAs per SM, so the register budget is 170 per thread rather than the 128 that
// eight CTAs would leave (measured: 25.9 -> 23.7 ms per 4736 tiles; the 128-register build serialised the chain).
template <int N, int NW>
__global__ void __launch_bounds__(32 * NW, (NW == 2 ? 6 : 16 / NW)) tlr_tsqr_kernel(TlrArgs ar, TlrWorkspace ws) {
    using SM = TsqrSmem<NW>;
    constexpr int GW = 8 * NW, NSP = N / kSubW, NT = 32 * NW;
    constexpr int kBarFull = 1, kBarFree = NW;         // named barriers: record of warp s published / consumed
    static_assert(2 * NW - 2 <= 15, "named barrier ids 1 .. 2 NW - 2");
    extern __shared__ __align__(16) double smem[];
    double *recbuf[2] = {smem + SM::kRec0, smem + SM::kRec1};
    double *own = smem + SM::kOwn;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM::kBar);

    const int t = blockIdx.x >> 1, pnl = blockIdx.x & 1;
    if (ws.st[t] != HCORE_B200_OK) return;
    const int gt = ar.tile0 + t;
    const int tid = threadIdx.x, warp = warp_id(), lane = lane_id();
    const int g = lane >> 2, tg = lane & 3;
    const int n = ws.ncols[t], c = ar.Crk[gt];
    const int n16 = round_up(n, kPanelW);
    const int m = ar.nb, m_pad = ws.m_pad;
    const double *Gt = ws.G + (size_t)t * m_pad * N;
    double *Rg = ws.R + nn_offset(ws, pnl, t, N);
    double *recg = ws.rec + rec_offset(ws, pnl, t, NSP);
    double *yb = smem + SM::kY + warp * 2 * kYbuf;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase[2] = {0, 0};
    if constexpr (N > kBlockH) {          // virtual rows kBlockH .. N-1 of R start as zeros (see warp_factor)
        for (int idx = tid; idx < (N - kBlockH) * N; idx += NT) Rg[(size_t)kBlockH + idx % (N - kBlockH) + (size_t)(idx / (N - kBlockH)) * N] = 0.0;
        __syncthreads();
    }
    PROF_DECL(tid == NT - 32)
    PROF_COUNT(9, 1);

    // Loop order: ROW BLOCK outer, column group inner.  The records a group streams (those of the earlier groups of
    // the SAME row block) were written by this CTA a moment ago, so the TMA reads hit L2 instead of re-streaming
    // 8 MB per tile from HBM as the group-outer order did (the running R lives in L2 either way).
    for (int j = 0; j < ws.nblk; ++j) {
        const bool blk0 = (j == 0);
        double *recj = recg + (size_t)j * NSP * kRecDoubles;
        for (int gc0 = 0; gc0 < n16; gc0 += GW) {
            const int gcols = (n16 - gc0 < GW) ? (n16 - gc0) : GW;      // multiple of 16
            const int nsub = gcols / kSubW;                             // sub-panels (= active warps) of the group
            const int lc = 8 * warp;                                    // my first slice column
            const bool have_cols = warp < nsub;
            const int mycol = gc0 + lc + g;
            const ColSrc src = panel_column(ar, Gt, m_pad, gt, pnl, mycol, c, n);
            const int nprev = gc0 / kSubW;                              // records of earlier groups
            // prefetch the first earlier record while the slice is loaded
            if (nprev > 0 && tid == 0) {
                mbar_expect_tx(&bars[0], kRecBytes);
                tma_load_1d(recbuf[0], recj, kRecBytes, &bars[0]);
            }
            // ---- my 8 columns of the row block -> registers -----------------------------------------
            double bx[16], by[16];
            const bool v16 = src.p && !src.f32 && ((reinterpret_cast<uintptr_t>(src.p + src.off) & 15) == 0);
            if (v16 && (j + 1) * kBlockH <= m) {
                // whole block in range: sixteen independent 16-byte loads in flight before the first use
                // (with the bounds checks inside the loop every load was followed by its consumer and
                // the sixteen DRAM/L2 latencies added up)
                const double2 *p2 = reinterpret_cast<const double2 *>(src.p + src.off + j * kBlockH + 2 * tg);
                double2 v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __ldg(p2 + 4 * i);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    bx[i] = src.scale * v[i].x;
                    by[i] = src.scale * v[i].y;
                }
            } else if (src.p && src.f32 && (j + 1) * kBlockH <= m &&
                       ((reinterpret_cast<uintptr_t>(reinterpret_cast<const float *>(src.p) + src.off) & 7) == 0)) {
                const float2 *p2 = reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(src.p) + src.off +
                                                                    j * kBlockH + 2 * tg);
                float2 v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __ldg(p2 + 4 * i);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    bx[i] = src.scale * (double)v[i].x;
                    by[i] = src.scale * (double)v[i].y;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int row = j * kBlockH + 8 * i + 2 * tg;
                    bx[i] = (src.p && row < m) ? src.scale * ld_user(src.p, src.off + row, src.f32) : 0.0;
                    by[i] = (src.p && row + 1 < m) ? src.scale * ld_user(src.p, src.off + row + 1, src.f32) : 0.0;
                }
            }
            PROF_ADD(0);
            // ---- (1) reflectors of the earlier groups ------------------------------------------------
            for (int sp = 0; sp < nprev; ++sp) {
                const int b = sp & 1;
                if (sp + 1 < nprev) {
                    __syncthreads();                 // everybody is done with buffer b^1 (record sp-1)
                    if (tid == 0) {
                        mbar_expect_tx(&bars[b ^ 1], kRecBytes);
                        tma_load_1d(recbuf[b ^ 1], recj + (size_t)(sp + 1) * kRecDoubles, kRecBytes, &bars[b ^ 1]);
                    }
                }
                mbar_wait(&bars[b], phase[b]);
                phase[b] ^= 1;
                if (have_cols) {
                    double sx = 0.0, sy = 0.0;
                    double *rs = Rg + (size_t)(sp * kSubW + 2 * tg) + (size_t)mycol * N;
                    double2 r0 = make_double2(0.0, 0.0);
                    if (!blk0) { r0 = *reinterpret_cast<const double2 *>(rs); sx = r0.x; sy = r0.y; }
                    wy_chain<true>(bx, by, sx, sy, recbuf[b], lane);
                    if (!blk0) *reinterpret_cast<double2 *>(rs) = make_double2(r0.x - sx, r0.y - sy);
                }
            }
            PROF_ADD(1);
            if (nprev > 0) {
                // the next async-proxy write (TMA) into the record buffers follows these generic reads
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
            }
            PROF_ADD(2);

            // ---- (2) the group's own sub-panels: warp pipeline -------------------------------------------
            if (have_cols) {
#pragma unroll
                for (int s = 0; s < NW - 1; ++s) {
                    if (s < warp) {
                        const int cnt = 32 * (nsub - s);
                        bar_sync_lit(kBarFull + s, cnt);                    // record of warp s is in its slot
                        double sx = 0.0, sy = 0.0;
                        double *rs = Rg + (size_t)(gc0 + s * kSubW + 2 * tg) + (size_t)mycol * N;
                        double2 r0 = make_double2(0.0, 0.0);
                        if (!blk0) { r0 = *reinterpret_cast<const double2 *>(rs); sx = r0.x; sy = r0.y; }
                        wy_chain<true>(bx, by, sx, sy, own + s * kRecDoubles, lane);
                        if (!blk0) *reinterpret_cast<double2 *>(rs) = make_double2(r0.x - sx, r0.y - sy);
                        bar_arrive_lit(kBarFree + s, cnt);                  // done with slot s
                    }
                }
                PROF_ADD(3);
                const bool publish = warp + 1 < nsub;
                double *slot = publish ? own + warp * kRecDoubles : nullptr;
                double *rg = recj + (size_t)(nprev + warp) * kRecDoubles;
                if (blk0) warp_factor<N, true>(bx, by, yb, slot, rg, Rg, gc0 + lc, lane);
                else warp_factor<N, false>(bx, by, yb, slot, rg, Rg, gc0 + lc, lane);
                PROF_ADD(4);
                if (publish) {
                    const int cnt = 32 * (nsub - warp);
                    __threadfence_block();
#pragma unroll
                    for (int w = 0; w < NW - 1; ++w) {
                        if (w == warp) {
                            bar_arrive_lit(kBarFull + w, cnt);
                            bar_sync_lit(kBarFree + w, cnt);                // slot reusable in the next row block
                        }
                    }
                }
                PROF_ADD(5);
            }
            // the records of this group (generic-proxy stores) are streamed by TMA (async proxy) in the next group
            if (gc0 + GW < n16) {
                __threadfence();
                asm volatile("fence.proxy.async.global;" ::: "memory");
                __syncthreads();
                if (tid == 0) asm volatile("fence.proxy.async.global;" ::: "memory");
            }
        }  // column groups
    }  // row blocks
    // padding columns n16..N-1 of R
    for (int idx = tid; idx < N * (N - n16); idx += NT) Rg[(size_t)n16 * N + idx] = 0.0;
}

// ========================================================================================
// 4. reflector application kernel
// ========================================================================================
// grid = (column chunks of 8*NW, 2 panels, tiles).  Every warp owns 8 output columns (register
// resident), the CTA streams all records of the panel once (TMA, double buffered).
template <int NW>
struct ApplySmem {
    static constexpr int kRec0 = 0;
    static constexpr int kRec1 = kRecDoubles;
    static constexpr int kBar = 2 * kRecDoubles;
    static constexpr int kTotal = kBar + 2;
    static constexpr size_t kBytes = (size_t)kTotal * sizeof(double);
};

template <int N, int NW>
__global__ void __launch_bounds__(32 * NW, 16 / NW) tlr_apply_kernel(TlrArgs ar, TlrWorkspace ws) {
    using SM = ApplySmem<NW>;
    constexpr int NSP = N / kSubW;
    extern __shared__ __align__(16) double smem[];
    double *recbuf[2] = {smem + SM::kRec0, smem + SM::kRec1};
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM::kBar);

    const int t = blockIdx.z, pnl = blockIdx.y, col0 = blockIdx.x * 8 * NW;
    if (ws.st[t] != HCORE_B200_OK) return;
    const int rnew = ws.newrank[t];
    const int gt = ar.tile0 + t;
    // split-phase call: the rank revealed by phase 1 becomes the tile's rank now (reference
    // HCORE_dgemm_ormqr, src/compute/hcore_dgemm.c:1688, Crk[0] = new_Crk)
    if (ar.phase == 2 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && ar.rank_out) ar.rank_out[gt] = rnew;
    if (col0 >= rnew) return;
    const int tid = threadIdx.x, warp = warp_id(), lane = lane_id();
    const int g = lane >> 2, tg = lane & 3;
    const int n16 = round_up(ws.ncols[t], kPanelW);
    const int nsp = n16 / kSubW;
    const int m = ar.nb;
    const int mycol = col0 + 8 * warp + g;                 // output column of this lane
    const bool warp_on = col0 + 8 * warp < rnew;           // warp has at least one valid column
    const bool col_on = mycol < rnew;
    const double *recg = ws.rec + rec_offset(ws, pnl, t, NSP);
    double *Yg = ws.Y + nn_offset(ws, pnl, t, N) + (size_t)mycol * N;     // my column of U_r / W
    double *outb = (pnl == 0 ? ar.CU[gt] : ar.CV[gt]);
    const size_t outo = (size_t)mycol * ar.ldc;          // element offset of my output column

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase[2] = {0, 0};

    // records in application order: blocks nblk-1 .. 0, sub-panels nsp-1 .. 0
    const int nrec = ws.nblk * nsp;
    auto rec_ptr = [&](int k) {
        const int j = ws.nblk - 1 - k / nsp, s = nsp - 1 - k % nsp;
        return recg + ((size_t)j * NSP + s) * kRecDoubles;
    };
    if (tid == 0) {
        mbar_expect_tx(&bars[0], kRecBytes);
        tma_load_1d(recbuf[0], rec_ptr(0), kRecBytes, &bars[0]);
    }
    double bx[16], by[16];
    for (int k = 0; k < nrec; ++k) {
        const int j = ws.nblk - 1 - k / nsp, s = nsp - 1 - k % nsp;
        const int b = k & 1;
        if (k + 1 < nrec) {
            __syncthreads();
            if (tid == 0) {
                mbar_expect_tx(&bars[b ^ 1], kRecBytes);
                tma_load_1d(recbuf[b ^ 1], rec_ptr(k + 1), kRecBytes, &bars[b ^ 1]);
            }
        }
        if (s == nsp - 1) {          // new row block: Z = 0, or [Y; 0] for block 0
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int row = 8 * i + 2 * tg;
                bx[i] = (j == 0 && col_on && row < n16) ? Yg[row] : 0.0;
                by[i] = (j == 0 && col_on && row + 1 < n16) ? Yg[row + 1] : 0.0;
            }
        }
        mbar_wait(&bars[b], phase[b]);
        phase[b] ^= 1;
        if (warp_on) {
            double sx = 0.0, sy = 0.0;
            double *ys = Yg + s * kSubW + 2 * tg;
            if (j > 0 && col_on) { sx = ys[0]; sy = ys[1]; }
            wy_chain<false>(bx, by, sx, sy, recbuf[b], lane);
            if (j > 0 && col_on) { ys[0] -= sx; ys[1] -= sy; }
            if (s == 0 && col_on) {      // row block finished -> output
                const bool v16 = !ar.f32 && (reinterpret_cast<uintptr_t>(outb + outo) & 15) == 0;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int row = j * kBlockH + 8 * i + 2 * tg;
                    if (v16 && row + 1 < m) {
                        *reinterpret_cast<double2 *>(outb + outo + row) = make_double2(bx[i], by[i]);
                    } else {
                        if (row < m) st_user(outb, outo + row, bx[i], ar.f32);
                        if (row + 1 < m) st_user(outb, outo + row + 1, by[i], ar.f32);
                    }
                }
            }
        }
    }
}

}  // namespace hcore_b200
