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
// Common machinery.  A "record" is the block reflector of one (128-row block, 16-column
// sub-panel): the explicit 128 x 16 reflector panel V in an XOR-swizzled layout plus its 16 x 16
// compact-WY factor T (and T^T), 22 528 contiguous bytes.  Records are produced once by the
// QR kernel and consumed by ONE TMA bulk copy (cp.async.bulk) each, double buffered, both by
// the left-looking QR itself and by the apply kernel.  The matrix the reflectors act on lives
// in REGISTERS: a warp owns 8 columns, held transposed as DMMA accumulator fragments
// (lane (g, tg) holds column g, rows 8i + 2tg, 8i + 2tg + 1 of the 128-row block).  With a
// consistent permutation of the contraction index an accumulator fragment is directly usable as
// the A operand of the next DMMA, so the whole chain
//      S^T = [S0^T +] B^T V ,  S2^T = S^T T(^T) ,  B^T -= S2^T V^T
// runs with no shared-memory round trip and no CTA barrier; only V and T are read from shared
// memory, conflict-free thanks to the swizzle.
#pragma once
#include "common.cuh"

namespace hcore_b200 {

constexpr int kRecV = kBlockH * kPanelW;           // 2048 doubles: swizzled reflector panel
constexpr int kLdT = 24;                           // leading dim of the T images (conflict-free LDS.128)
constexpr int kRecT = kPanelW * kLdT;              // 384
constexpr int kRecDoubles = kRecV + 2 * kRecT;     // V | T | T^T  = 2816
constexpr int kRecBytes = kRecDoubles * 8;         // 22 528

// Swizzled position of element (row r, column c) of a 128 x 16 reflector panel.  16-byte slots
// hold row pairs; slot = c*64 + ((r/2) ^ h(c)).  h is chosen so that both access patterns of
// the chain are bank-conflict free:
//   GEMM1 (LDS.128): lanes (g, tg) read rows (8i+2tg, 8i+2tg+1) of column 8t+g
//   GEMM3 (LDS.64) : lanes (g, tg) read row 8i+g of columns 8u+2tg and 8u+2tg+1
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
// bx[i], by[i]: B^T fragments of 128-row tile i (i = 0..15): column g of the warp's 8 columns,
// rows 8i + 2tg and 8i + 2tg + 1.
// sx/sy[t]   : in  S0^T fragments (R strip / Y rows / zero): column g, reflectors 8t+2tg, 8t+2tg+1
//              out S2^T fragments in the same layout.
// FORWARD = true : S2 = T^T S (factorisation, Q^T applied);  false: S2 = T S (Q applied).
template <bool FORWARD>
__device__ __forceinline__ void wy_chain(double (&bx)[16], double (&by)[16], double (&sx)[2], double (&sy)[2],
                                         const double *rec, int lane) {
    const int g = lane >> 2, tg = lane & 3;
    // ---- GEMM1: S^T += B^T V --------------------------------------------------------------------
    {
        const int h = ((g & 1) << 2) ^ (((g >> 1) & 3) << 1);
        const int hlo = h & 3, hhi = (h >> 2) & 1;
        const double *base = rec + 2 * ((g << 6) + (tg ^ hlo));
        const double *pe = base + 8 * hhi, *po = base - 8 * hhi;
        double e0x = 0.0, e0y = 0.0, e1x = 0.0, e1y = 0.0;   // second accumulator set (odd tiles)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const double *q = ((i & 1) ? po : pe) + 8 * i;
            const double2 v0 = *reinterpret_cast<const double2 *>(q);
            const double2 v1 = *reinterpret_cast<const double2 *>(q + 1024);
            if (i & 1) {
                dmma_884(e0x, e0y, bx[i], v0.x);
                dmma_884(e1x, e1y, bx[i], v1.x);
                dmma_884(e0x, e0y, by[i], v0.y);
                dmma_884(e1x, e1y, by[i], v1.y);
            } else {
                dmma_884(sx[0], sy[0], bx[i], v0.x);
                dmma_884(sx[1], sy[1], bx[i], v1.x);
                dmma_884(sx[0], sy[0], by[i], v0.y);
                dmma_884(sx[1], sy[1], by[i], v1.y);
            }
        }
        sx[0] += e0x; sy[0] += e0y; sx[1] += e1x; sy[1] += e1y;
    }
    // ---- GEMM2: S2^T = S^T Tm,  Tm = T (forward) or T^T ----------------------------------------------
    double px[2], py[2];
    {
        const double *Tm = rec + kRecV + (FORWARD ? 0 : kRecT);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            double ax = 0.0, ay = 0.0;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const double2 tv = *reinterpret_cast<const double2 *>(Tm + (8 * t + 2 * tg) + kLdT * (8 * u + g));
                dmma_884(ax, ay, sx[t], tv.x);
                dmma_884(ax, ay, sy[t], tv.y);
            }
            px[u] = ax; py[u] = ay;
        }
    }
    // ---- GEMM3: B^T -= S2^T V^T ----------------------------------------------------------------------
    {
        const int gh = g >> 1;
        const int hb = tg >> 1;
        const double *base = rec + 2 * (((2 * tg) << 6) + (gh ^ ((tg & 1) << 1))) + (g & 1);
        const double *pe_e = base + 8 * hb, *po_e = base - 8 * hb;                  // even column 2tg
        const double *pe_o = base + 128 + 8 * (hb ^ 1), *po_o = base + 128 - 8 * (hb ^ 1);   // odd column 2tg+1
        const double n0x = -px[0], n0y = -py[0], n1x = -px[1], n1y = -py[1];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const double *qe = ((i & 1) ? po_e : pe_e) + 8 * i;
            const double *qo = ((i & 1) ? po_o : pe_o) + 8 * i;
            dmma_884(bx[i], by[i], n0x, qe[0]);
            dmma_884(bx[i], by[i], n0y, qo[0]);
            dmma_884(bx[i], by[i], n1x, qe[1024]);     // columns 8 + 2tg, 8 + 2tg + 1
            dmma_884(bx[i], by[i], n1y, qo[1024]);
        }
    }
    sx[0] = px[0]; sy[0] = py[0]; sx[1] = px[1]; sy[1] = py[1];
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
    double *rec;     // [2][chunk][nblk][N/16][kRecDoubles]  block-reflector records
    double *G;       // [chunk][m_pad * N]                   G = BU (AV^T BV)^T, ld m_pad
    double *R;       // [2][chunk][N * N]
    double *Y;       // [2][chunk][N * N]            U_r (panel 0) and W (panel 1)
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
constexpr int kPgKC = 64;               // slab width
constexpr int kPgLdP = kPgKC + 4;       // 68

struct PgSmem {
    static constexpr int kStage = 2 * kPgKC * kPgLd;         // one stage: two 32 x 64 operand chunks
    static constexpr int kP = 2 * kStage;                    // P slab, ld 68
    static constexpr int kTotal = kP + kPgKC * kPgLdP;
    static constexpr size_t kBytes = (size_t)kTotal * sizeof(double);
};

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

template <int N>
__global__ void __launch_bounds__(256, 2) tlr_pg_kernel(TlrArgs ar, TlrWorkspace ws) {
    using SM = PgSmem;
    extern __shared__ __align__(16) double smem[];
    double *P = smem + SM::kP;
    const int t = blockIdx.x, gt = ar.tile0 + t;
    const int tid = threadIdx.x, warp = warp_id(), lane = lane_id();
    const int g = lane >> 2, tg = lane & 3;
    const int a = ar.Ark[gt], b = ar.Brk[gt], c = ar.Crk[gt];
    const int n = a + c, m = ar.nb;

    int st = HCORE_B200_OK;
    if (a <= 0 || b <= 0 || c <= 0) st = HCORE_B200_ERR_ZERO_RANK;
    else if (ar.maxrk < n) st = HCORE_B200_ERR_MAXRK;
    else if (n > N || n > m || b > N) st = HCORE_B200_ERR_UNSUPPORTED;
    if (tid == 0) {
        ws.st[t] = st;
        ws.ncols[t] = n;
        ws.newrank[t] = c;
        if (ar.status) ar.status[gt] = st;
    }
    if (st != HCORE_B200_OK) return;

    const double *AV = ar.AV[gt], *BU = ar.BU[gt], *BV = ar.BV[gt];
    const int m_pad = ws.m_pad;
    const int nchunks = (m + kPgRows - 1) / kPgRows, nchunks_pad = m_pad / kPgRows;
    const int f32 = ar.f32;
    const bool va = !f32 && ((reinterpret_cast<uintptr_t>(AV) & 15) == 0) && ((ar.lda & 1) == 0);
    const bool vb = !f32 && ((reinterpret_cast<uintptr_t>(BV) & 15) == 0) && ((reinterpret_cast<uintptr_t>(BU) & 15) == 0) &&
                    ((ar.ldb & 1) == 0);
    double *Gt = ws.G + (size_t)t * m_pad * N;

    for (int ia = 0; ia < a; ia += kPgKC) {
        const int as = (a - ia < kPgKC) ? (a - ia) : kPgKC, as8 = round_up(as, 8);
        for (int jb = 0; jb < b; jb += kPgKC) {
            const int bs = (b - jb < kPgKC) ? (b - jb) : kPgKC, bs8 = round_up(bs, 8);
            const double *AVs = AV, *BVs = BV, *BUs = BU;      // slab offsets are passed as column indices
            // ---- P slab = AV[:, ia..]^T BV[:, jb..]: warp -> 8-row strip of P, all column tiles -----
            {
                const bool active = 8 * warp < as8;
                const int i0 = 8 * warp;
                double acc[8][2];
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q][0] = acc[q][1] = 0.0;
                pg_stage(smem, AVs, ia, ar.lda, m, 0, as, as8, va, f32, tid);
                pg_stage(smem + kPgKC * kPgLd, BVs, jb, ar.ldb, m, 0, bs, bs8, vb, f32, tid);
                cp_async_commit();
                for (int ch = 0; ch < nchunks; ++ch) {
                    const double *cur = smem + (ch & 1) * SM::kStage;
                    if (ch + 1 < nchunks) {
                        double *nxt = smem + ((ch + 1) & 1) * SM::kStage;
                        pg_stage(nxt, AVs, ia, ar.lda, m, (ch + 1) * kPgRows, as, as8, va, f32, tid);
                        pg_stage(nxt + kPgKC * kPgLd, BVs, jb, ar.ldb, m, (ch + 1) * kPgRows, bs, bs8, vb, f32, tid);
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
                            for (int q = 0; q < 8; ++q) {
                                if (8 * q < bs8) {
                                    const double bv = Bs[(k + tg) + (8 * q + g) * kPgLd];
                                    dmma_884(acc[q][0], acc[q][1], av, bv);
                                }
                            }
                        }
                    }
                    __syncthreads();
                }
                if (active) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (8 * q < bs8) {
                            P[(i0 + g) + (8 * q + 2 * tg) * kPgLdP] = acc[q][0];
                            P[(i0 + g) + (8 * q + 2 * tg + 1) * kPgLdP] = acc[q][1];
                        }
                    }
                }
            }
            __syncthreads();
            // ---- G[:, ia..] (+)= BU[:, jb..] P^T: warp -> (8-row strip of the chunk, 4 column tiles) ------
            {
                const int strip = warp & 3, q0 = 4 * (warp >> 2);
                pg_stage(smem, BUs, jb, ar.ldb, m, 0, bs, bs8, vb, f32, tid);
                cp_async_commit();
                for (int ch = 0; ch < nchunks_pad; ++ch) {
                    const double *cur = smem + (ch & 1) * SM::kStage;
                    if (ch + 1 < nchunks_pad) {
                        pg_stage(smem + ((ch + 1) & 1) * SM::kStage, BUs, jb, ar.ldb, m, (ch + 1) * kPgRows, bs, bs8, vb, f32, tid);
                        cp_async_commit();
                        cp_async_wait<1>();
                    } else {
                        cp_async_wait<0>();
                    }
                    __syncthreads();
                    double acc[4][2];
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[q][0] = acc[q][1] = 0.0;
#pragma unroll 4
                    for (int k = 0; k < bs8; k += 4) {
                        const double av = cur[(8 * strip + g) + (k + tg) * kPgLd];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (8 * (q0 + q) < as8) {
                                const double bv = P[(8 * (q0 + q) + g) + (k + tg) * kPgLdP];
                                dmma_884(acc[q][0], acc[q][1], av, bv);
                            }
                        }
                    }
                    const int row = ch * kPgRows + 8 * strip + g;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
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

// ========================================================================================
// 2. tall-skinny QR kernel
// ========================================================================================
// Flat-tree TSQR over 128-row blocks: block 0 is a plain blocked Householder QR (reflectors below
// the diagonal, R on top); every further block B_j is eliminated against the running R
// ("triangle on top of a dense block": reflector i = [e_i ; v_i]).  16-column sub-panels with
// compact-WY T factors.  Columns are processed LEFT-LOOKING in groups of GW = 8 * NW columns
// (NW warps per CTA, 8 columns per warp, register resident): for every row block the CTA
//   (1) streams the records of all earlier sub-panels (TMA, double buffered) through wy_chain,
//   (2) factors its own sub-panels: the owning warps drop the 16 columns into shared memory, all
//       warps run the Householder steps (one column per 2*NW lanes, one barrier per step), the
//       resulting record is applied to the group's trailing columns from shared memory and
//       written to global memory.
constexpr int kLdPf = kBlockH + 8;     // 136: panel buffer of the factorisation

template <int NW>
struct TsqrSmem {
    static constexpr int kRec0 = 0;
    static constexpr int kRec1 = kRec0 + kRecDoubles;
    static constexpr int kPf = kRec1 + kRecDoubles;
    static constexpr int kVbuf = kPf + kPanelW * kLdPf;   // [2][128]
    static constexpr int kRd = kVbuf + 2 * kBlockH;       // 16 x 16 diagonal block of R, Rd[ii + 16 c]
    static constexpr int kTs = kRd + 256;                 // T, Ts[k + 16 c]
    static constexpr int kTau = kTs + 256;                // [16]
    static constexpr int kZ = kTau + 16;                  // [2][16]
    static constexpr int kHs = kZ + 32;                   // tau, scal of the current step
    static constexpr int kBar = kHs + 2;                  // 2 mbarriers
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

// Column jc of the compact-WY factor (forward columnwise larft), computed by one warp:
//   T[r][jc] = -tau_jc * sum_{q=r}^{jc-1} T[r][q] z_q ,  z_q = v_q^T v_jc,  r < jc.
// Fully unrolled and predicated so that all shared-memory loads are issued up front (this sits
// on the path to the next CTA barrier).
__device__ __forceinline__ void t_column(double *Ts, const double *zb, const double *taus, int jc, int lane) {
    const double *z = zb + (jc & 1) * 16;
    const double mtau = -taus[jc];
    const int r = lane;
    if (r < jc) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int q = 0; q < kPanelW - 1; ++q) {
            const bool on = (q >= r) && (q < jc);
            const double tv = on ? Ts[r + 16 * q] : 0.0, zv = z[q];
            if (q & 1) a1 = fma(tv, zv, a1);
            else a0 = fma(tv, zv, a0);
        }
        Ts[r + 16 * jc] = mtau * (a0 + a1);
    }
}

// Householder factorisation of the 128 x 16 sub-panel held in Pf (plain column-major, ld kLdPf).
// Produces the record (V swizzled, T, T^T) in `rec`, the R entries in Rg and (blocks > 0) uses /
// updates the 16 x 16 diagonal block of the running R.  All NW warps participate: warp w owns the
// columns w, w + NW, ... (CPW = 16 / NW of them), lane l owns rows l, l+32, l+64, l+96; every role
// decision is warp-uniform.
//
// Step ii is split so that the latency-critical scalar work overlaps the dot products:
//   phase A  the RAW (unscaled) tail y of column ii is already in shared memory; every warp forms
//            d_c = y^T x_c for its columns (for the owner's own column this IS the squared norm)
//            and, in block 0, extracts its row-i elements; one DMMA all-reduce; the owner derives
//            (beta, tau, scal) and publishes them                                   -> barrier
//   phase B  w_c = tau (r_c + scal d_c),  x_c -= y (scal w_c)  [v = scal y is never formed for the
//            update], the owner scales its column to v, the NEXT owner publishes its raw updated
//            column                                                                 -> barrier
template <int N, int NW, bool BLK0>
__device__ __forceinline__ void panel_factor(double *smem, double *rec, double *Rg, int c0, int tid) {
    using SM = TsqrSmem<NW>;
    constexpr int CPW = kPanelW / NW, RPT = kBlockH / 32;     // columns per warp, rows per lane
    constexpr int NV = BLK0 ? 2 * CPW : CPW;
    double *Pf = smem + SM::kPf, *vbuf = smem + SM::kVbuf, *Rd = smem + SM::kRd, *Ts = smem + SM::kTs,
           *taus = smem + SM::kTau, *zb = smem + SM::kZ, *hs = smem + SM::kHs;
    const int w = tid >> 5, l = tid & 31;
    double x[CPW][RPT];
#pragma unroll
    for (int q = 0; q < CPW; ++q)
#pragma unroll
        for (int k = 0; k < RPT; ++k) x[q][k] = Pf[(l + 32 * k) + (w + NW * q) * kLdPf];
    if (!BLK0) {
        for (int idx = tid; idx < 256; idx += 32 * NW)
            Rd[idx] = Rg[(size_t)(c0 + (idx & 15)) + (size_t)(c0 + (idx >> 4)) * N];
    }
    for (int idx = tid; idx < 256; idx += 32 * NW) Ts[idx] = 0.0;
    if (w == 0) {       // raw tail of column 0
#pragma unroll
        for (int k = 0; k < RPT; ++k) vbuf[l + 32 * k] = (BLK0 && l + 32 * k <= c0) ? 0.0 : x[0][k];
    }
    __syncthreads();
    PROF_DECL(tid == 32 * NW - 32)

#pragma unroll 1
    for (int ii = 0; ii < kPanelW; ++ii) {
        const int i = c0 + ii;                 // global column = diagonal row in block 0
        const int so = ii / NW;                // slot of column ii in its owner warp
        const int ki = i >> 5, li = i & 31;    // block 0: row i lives in lane li, register ki
        const double *vb = vbuf + (ii & 1) * kBlockH;
        // ---------------- phase A -------------------------------------------------------------------
        double yb[RPT], vals[NV], rd[CPW];
#pragma unroll
        for (int k = 0; k < RPT; ++k) yb[k] = vb[l + 32 * k];
        if (!BLK0) {
#pragma unroll
            for (int q = 0; q < CPW; ++q) rd[q] = Rd[ii + 16 * (w + NW * q)];
        }
#pragma unroll
        for (int q = 0; q < CPW; ++q) {
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                if (k & 1) d1 = fma(yb[k], x[q][k], d1);
                else d0 = fma(yb[k], x[q][k], d0);
            }
            vals[q] = d0 + d1;
            if (BLK0) {
                double xs = x[q][0];
#pragma unroll
                for (int k = 1; k < RPT; ++k) xs = (k == ki) ? x[q][k] : xs;
                vals[CPW + q] = (l == li) ? xs : 0.0;
            }
        }
        warp_allsum_dmma<NV>(vals);
        if (BLK0) {
#pragma unroll
            for (int q = 0; q < CPW; ++q) rd[q] = vals[CPW + q];
        }
        double beta = 0.0;
        if (w == ii % NW) {                    // owner warp: reflector scalars
            double s2 = vals[0], alpha = rd[0];
#pragma unroll
            for (int q = 1; q < CPW; ++q) { s2 = (q == so) ? vals[q] : s2; alpha = (q == so) ? rd[q] : alpha; }
            const Householder h = make_householder(alpha, s2);
            beta = h.beta;
            if (l == 0) {
                hs[0] = h.tau;
                hs[1] = h.scal;
                taus[ii] = h.tau;
                Ts[ii + 16 * ii] = h.tau;
                if (!BLK0) Rd[ii + 16 * ii] = h.beta;
            }
        }
        // T column ii-1 by the warp that owned it (its z values were published in step ii-1)
        if (ii > 1 && w == (ii - 1) % NW) t_column(Ts, zb, taus, ii - 1, l);
        __syncthreads();
        PROF_ADD(30);
        // ---------------- phase B -------------------------------------------------------------------
        const double tau = hs[0], scal = hs[1];
#pragma unroll
        for (int q = 0; q < CPW; ++q) {
            const int cc = w + NW * q;
            const double zc = fma(scal, vals[q], rd[q]);          // r_c + scal * d_c
            if (cc > ii) {
                const double wv = tau * zc, f = scal * wv;
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    x[q][k] = fma(-yb[k], f, x[q][k]);
                    if (BLK0) x[q][k] = (l == li && k == ki) ? x[q][k] - wv : x[q][k];
                }
                if (!BLK0 && l == 0) Rd[ii + 16 * cc] = rd[q] - wv;
            } else if (cc < ii) {
                if (l == 0) zb[(ii & 1) * 16 + cc] = zc;          // z_c = v_c^T v_ii for the T factor
            } else {                                              // my column becomes the reflector
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    const double v = yb[k] * scal;
                    if (BLK0) x[q][k] = (l + 32 * k > i) ? v : ((l == li && k == ki) ? beta : x[q][k]);
                    else x[q][k] = v;
                }
            }
        }
        if (ii + 1 < kPanelW && w == (ii + 1) % NW) {             // next owner: publish the raw tail
            const int so2 = (ii + 1) / NW;
            double *vn = vbuf + ((ii + 1) & 1) * kBlockH;
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                double v = x[0][k];
#pragma unroll
                for (int q = 1; q < CPW; ++q) v = (q == so2) ? x[q][k] : v;
                vn[l + 32 * k] = (BLK0 && l + 32 * k <= i + 1) ? 0.0 : v;
            }
        }
        __syncthreads();
        PROF_ADD(31);
    }
    if (w == (kPanelW - 1) % NW) t_column(Ts, zb, taus, kPanelW - 1, l);          // last T column
    // ---- outputs: explicit reflector panel (swizzled), R entries ------------------------------------
#pragma unroll
    for (int q = 0; q < CPW; ++q) {
        const int cc = w + NW * q, col = c0 + cc;
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int row = l + 32 * k;
            double v = x[q][k];
            if (BLK0) {
                v = (row > col) ? x[q][k] : (row == col ? 1.0 : 0.0);
                if (row < N) Rg[(size_t)row + (size_t)col * N] = (row <= col) ? x[q][k] : 0.0;
            }
            rec[vswz(row, cc)] = v;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < 256; idx += 32 * NW) {
        const int k = idx & 15, q = idx >> 4;            // T[k][q]
        const double tv = Ts[idx];
        rec[kRecV + k + kLdT * q] = tv;                  // T   : B[k][n] = T[k][n]
        rec[kRecV + kRecT + q + kLdT * k] = tv;          // T^T : B[k][n] = T[n][k]
        if (!BLK0) Rg[(size_t)(c0 + k) + (size_t)(c0 + q) * N] = (q >= k) ? Rd[k + 16 * q] : 0.0;
    }
    for (int idx = tid; idx < 2 * (kLdT - 16) * 16; idx += 32 * NW) {   // padding of the T images
        const int im = idx / ((kLdT - 16) * 16), r = idx % ((kLdT - 16) * 16);
        rec[kRecV + im * kRecT + 16 + (r % (kLdT - 16)) + kLdT * (r / (kLdT - 16))] = 0.0;
    }
    __syncthreads();
}

// One CTA (NW warps) per panel: blockIdx.x = 2 * tile + panel.
template <int N, int NW>
__global__ void __launch_bounds__(32 * NW, 16 / NW) tlr_tsqr_kernel(TlrArgs ar, TlrWorkspace ws) {
    using SM = TsqrSmem<NW>;
    constexpr int GW = 8 * NW, NSP = N / kPanelW, NT = 32 * NW;
    extern __shared__ __align__(16) double smem[];
    double *recbuf[2] = {smem + SM::kRec0, smem + SM::kRec1};
    double *Pf = smem + SM::kPf;
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

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase[2] = {0, 0};
    PROF_DECL(tid == NT - 32)
    PROF_COUNT(9, 1);

    for (int gc0 = 0; gc0 < n16; gc0 += GW) {
        const int gcols = (n16 - gc0 < GW) ? (n16 - gc0) : GW;      // multiple of 16
        const int lc = 8 * warp;                                    // my first slice column
        const bool have_cols = lc < gcols;
        const ColSrc src = panel_column(ar, Gt, m_pad, gt, pnl, gc0 + lc + g, c, n);
        const int nprev = gc0 / kPanelW;                            // records of earlier sub-panels
        for (int j = 0; j < ws.nblk; ++j) {
            const bool blk0 = (j == 0);
            const double *recj = recg + (size_t)j * NSP * kRecDoubles;
            // prefetch the first earlier record while the slice is loaded
            if (nprev > 0 && tid == 0) {
                mbar_expect_tx(&bars[0], kRecBytes);
                tma_load_1d(recbuf[0], recj, kRecBytes, &bars[0]);
            }
            // ---- my 8 columns of the row block -> registers -----------------------------------------
            double bx[16], by[16];
            const bool v16 = src.p && !src.f32 && ((reinterpret_cast<uintptr_t>(src.p + src.off) & 15) == 0);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int row = j * kBlockH + 8 * i + 2 * tg;
                if (v16 && row + 1 < m) {
                    const double2 v = __ldg(reinterpret_cast<const double2 *>(src.p + src.off + row));
                    bx[i] = src.scale * v.x;
                    by[i] = src.scale * v.y;
                } else {
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
                    const int pc0 = sp * kPanelW;
                    double sx[2] = {0.0, 0.0}, sy[2] = {0.0, 0.0};
                    double *rs = Rg + (size_t)pc0 + 2 * tg + (size_t)(gc0 + lc + g) * N;
                    if (!blk0) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) { sx[q] = rs[8 * q]; sy[q] = rs[8 * q + 1]; }
                    }
                    wy_chain<true>(bx, by, sx, sy, recbuf[b], lane);
                    if (!blk0) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) { rs[8 * q] -= sx[q]; rs[8 * q + 1] -= sy[q]; }
                    }
                }
            }
            PROF_ADD(1);
            __syncthreads();     // record buffers free (generic reads done), R strips visible
            PROF_ADD(2);

            // ---- (2) the group's own sub-panels ---------------------------------------------------------
            for (int lc0 = 0; lc0 < gcols; lc0 += kPanelW) {
                const int c0 = gc0 + lc0, s = c0 / kPanelW;
                double *rec = recbuf[s & 1];
                if (lc >= lc0 && lc < lc0 + kPanelW) {          // owners: registers -> panel buffer
                    const int pc = lc - lc0 + g;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        *reinterpret_cast<double2 *>(Pf + (8 * i + 2 * tg) + pc * kLdPf) = make_double2(bx[i], by[i]);
                }
                __syncthreads();
                PROF_ADD(3);
                if (blk0) panel_factor<N, NW, true>(smem, rec, Rg, c0, tid);
                else panel_factor<N, NW, false>(smem, rec, Rg, c0, tid);
                PROF_ADD(4);
                // record -> global (all threads, coalesced)
                {
                    double *dst = recg + ((size_t)j * NSP + s) * kRecDoubles;
                    for (int idx = tid; idx < kRecDoubles / 2; idx += NT)
                        reinterpret_cast<double2 *>(dst)[idx] = reinterpret_cast<const double2 *>(rec)[idx];
                }
                PROF_ADD(5);
                // trailing columns of the group
                if (have_cols && lc >= lc0 + kPanelW) {
                    double sx[2] = {0.0, 0.0}, sy[2] = {0.0, 0.0};
                    double *rs = Rg + (size_t)c0 + 2 * tg + (size_t)(gc0 + lc + g) * N;
                    if (!blk0) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) { sx[q] = rs[8 * q]; sy[q] = rs[8 * q + 1]; }
                    }
                    wy_chain<true>(bx, by, sx, sy, rec, lane);
                    if (!blk0) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) { rs[8 * q] -= sx[q]; rs[8 * q + 1] -= sy[q]; }
                    }
                }
                PROF_ADD(6);
            }
            // async-proxy writes (TMA) into the record buffers follow generic-proxy accesses
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            PROF_ADD(7);
        }  // row blocks
    }  // column groups
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
    constexpr int NSP = N / kPanelW;
    extern __shared__ __align__(16) double smem[];
    double *recbuf[2] = {smem + SM::kRec0, smem + SM::kRec1};
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM::kBar);

    const int t = blockIdx.z, pnl = blockIdx.y, col0 = blockIdx.x * 8 * NW;
    if (ws.st[t] != HCORE_B200_OK) return;
    const int rnew = ws.newrank[t];
    if (col0 >= rnew) return;
    const int gt = ar.tile0 + t;
    const int tid = threadIdx.x, warp = warp_id(), lane = lane_id();
    const int g = lane >> 2, tg = lane & 3;
    const int n16 = round_up(ws.ncols[t], kPanelW);
    const int nsp = n16 / kPanelW;
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
            double sx[2] = {0.0, 0.0}, sy[2] = {0.0, 0.0};
            double *ys = Yg + s * kPanelW + 2 * tg;
            if (j > 0 && col_on) {
#pragma unroll
                for (int q = 0; q < 2; ++q) { sx[q] = ys[8 * q]; sy[q] = ys[8 * q + 1]; }
            }
            wy_chain<false>(bx, by, sx, sy, recbuf[b], lane);
            if (j > 0 && col_on) {
#pragma unroll
                for (int q = 0; q < 2; ++q) { ys[8 * q] -= sx[q]; ys[8 * q + 1] -= sy[q]; }
            }
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
