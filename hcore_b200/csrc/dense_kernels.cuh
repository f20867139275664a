// Dense-C kernels of the TLR Cholesky step (FP64, sm_100a):
//   HCORE_dgemmcd     CD <- CD - (Au Av^T)(Bu Bv^T)^T      reference src/compute/hcore_dgemmcd.c:36-107
//   HCORE_dsyrk       CD <- CD + alpha (AU AV^T)(AU AV^T)^T, strict upper triangle zeroed
//                                                            reference src/compute/hcore_dsyrk.c:47-189
//   HCORE_dgemmbdcd   CD <- alpha AU (AV^T BD) + beta CD    reference src/compute/hcore_dgemmbdcd.c:35-122
//   HCORE_duncompress CD <- AU BV^T                         reference src/compute/hcore_duncompress.c:33-85
//   HCORE_zuncompress the same in complex double            reference src/compute/hcore_zuncompress.c:35-95
// All of them are a rank-k update  CD <- beta CD + s X Y^T  that streams the dense tile once: HBM-bound
// (8 * m * n bytes read unless beta = 0, and written).  Kernels:
//   dense_prep_kernel    X = s * Au * (Av^T Bv)   (m x b, tiny) into the context workspace   (gemmcd, syrk)
//   dense_bdtav_kernel   Y = BD^T AV              (n x a)       into the context workspace   (gemmbdcd)
//   dense_update_kernel  CD[I, J] = beta CD[I, J] + s X[I, :] Y[J, :]^T   per 128 x 64 block of CD (TMA in / out)
//   zdense_update_kernel the complex product (four real DMMAs per tile product), write only
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "wy_kernels.cuh"

namespace hcore_b200 {

struct DenseArgs {
    // ---- prep: X = scale * AU (AV^T BV) --------------------------------------------------------------------
    const double *AU, *AV, *BV;          // strided factor arrays (BV == AV for syrk)
    const int *Ark, *Brk;
    long strideA, strideB;
    int lda, ldb;                        // leading dimensions of the A and B factor arrays
    int mc;                              // contraction length of AV^T BV (rows); dsyrk's reference uses LDAU (hcore_dsyrk.c:57)
    double *X;                           // workspace [ntiles][m_pad * kcap]
    double scale;                        // -1 for gemmcd, alpha for syrk
    int m_pad, kcap;
    // ---- update: CD = beta CD + su * Xop Yop^T -------------------------------------------------------------
    const double *Xop, *Yop;             // m x k (ld ldx) and n x k (ld ldy) per tile
    long strideX, strideY, strideC;
    int ldx, ldy, ldc;
    const int *Krk;                      // contraction length per tile
    double *CD;
    int nb, ncols;                       // rows and columns of CD
    double su, beta;                     // beta == 0: CD is not read
    int syrk;                            // 1: lower triangle only + zero strict upper
    // pointer-array form (HCORE_dgemmcd_batched / HCORE_dsyrk_batched): per-tile pointers replace base + t * stride
    const double *const *pAU, *const *pAV, *const *pBU, *const *pBV;
    double *const *pCD;
};

// X = scale * Au (Av^T Bv)  (m x b, ld m_pad)      reference hcore_dgemmcd.c:66-78, hcore_dsyrk.c:103-144.
// Same shape of work as the P/G products of the low-rank path (pg_body, wy_kernels.cuh) with the roles
// swapped: P = Bv^T Av, X = Au P^T.  One CTA (256 threads) per tile, operands staged by cp.async.
__global__ void __launch_bounds__(256, 2) dense_prep_kernel(DenseArgs ar) {
    extern __shared__ __align__(16) double smem[];
    const int t = blockIdx.x;
    const int a = ar.Ark[t], b = ar.Brk[t], m = ar.nb;
    if (b <= 0) return;                    // the update kernel sees a zero contraction length
    if (a <= 0) {                          // empty A: X = 0 (never stale workspace contents)
        double *X = ar.X + (size_t)t * ar.m_pad * ar.kcap;
        for (int idx = threadIdx.x; idx < ar.m_pad * b; idx += blockDim.x) X[idx] = 0.0;
        return;
    }
    const double *AU = ar.pAU ? ar.pAU[t] : ar.AU + (size_t)t * ar.strideA, *AV = ar.pAV ? ar.pAV[t] : ar.AV + (size_t)t * ar.strideA;
    const double *BV = ar.pBV ? ar.pBV[t] : ar.BV + (size_t)t * ar.strideB;
    pg_body(smem, BV, AU, AV, b, a, m, ar.ldb, ar.lda, ar.m_pad, 0, ar.scale, ar.X + (size_t)t * ar.m_pad * ar.kcap, ar.mc);
}

// Y = BD^T AV (n x a, ld n_pad) for HCORE_dgemmbdcd (reference hcore_dgemmbdcd.c:78-84 / :100-106 computes its
// transpose AV^T BD).  grid = (64-row blocks of Y, tiles), 256 threads; warp w owns rows 8 w .. 8 w + 7 of the block.
// BD is read once (k-contiguous 32-byte segments per lane quad), AV comes out of L1/L2.
struct BdtAvArgs {
    const double *AV, *BD;
    const int *Ark;
    long strideA, strideB;
    int lda, ldb, mc, ncols;             // contraction length (rows of AV and BD), columns of BD
    double *Y;                           // per tile: ncols x a, leading dimension ldy, element stride strideY
    long strideY;
    int ldy, n_pad;                      // n_pad = ncols rounded up to the 64-row blocks of the grid
};
__global__ void __launch_bounds__(256) dense_bdtav_kernel(BdtAvArgs ar) {
    const int t = blockIdx.y, a = ar.Ark[t];
    if (a <= 0) return;
    const int warp = warp_id(), lane = lane_id(), g = lane >> 2, tg = lane & 3;
    const int i0 = blockIdx.x * 64 + 8 * warp;                     // row of Y = column of BD
    if (i0 >= ar.n_pad) return;
    const double *AV = ar.AV + (size_t)t * ar.strideA, *BD = ar.BD + (size_t)t * ar.strideB;
    double *Y = ar.Y + (size_t)t * ar.strideY;
    const bool ron = i0 + g < ar.ncols;
    const double *bp = BD + (size_t)(i0 + g) * ar.ldb + tg;
    for (int q0 = 0; q0 < a; q0 += 64) {                            // 8 column tiles of Y at a time
        double acc[8][2];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q][0] = acc[q][1] = 0.0;
        // whole 32-deep contraction chunks: the eight BD loads of a chunk are in flight together (BD is streamed from HBM
        // exactly once), AV fragments come out of L1 / L2; columns >= a of Y are never stored, so their AV operand is read
        // from a clamped (valid) column instead of being masked
        const double *avp[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int col = q0 + 8 * q + g;
            avp[q] = AV + (size_t)(col < a ? col : a - 1) * ar.lda + tg;
        }
        int k = 0;
        for (; k + 32 <= ar.mc; k += 32) {
            double bv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) bv[u] = ron ? __ldg(bp + k + 4 * u) : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (q0 + 8 * q >= a) break;        // a real exit: predicated-off DMMAs still occupy the FP64 pipe
                    dmma_884(acc[q][0], acc[q][1], bv[u], __ldg(avp[q] + k + 4 * u));
                }
            }
        }
        for (; k < ar.mc; k += 4) {
            const double bv = (ron && k + tg < ar.mc) ? __ldg(bp + k) : 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (q0 + 8 * q >= a) break;
                const double av = (k + tg < ar.mc) ? __ldg(avp[q] + k) : 0.0;
                dmma_884(acc[q][0], acc[q][1], bv, av);
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = q0 + 8 * q + 2 * tg + e;
                if (col < a && (ron || ar.ldy >= ar.n_pad)) Y[(i0 + g) + (size_t)col * ar.ldy] = acc[q][e];   // padding rows only if they exist
            }
        }
    }
}

// Y = A X for a dense tile A (m x n, ld lda) and a skinny X (n x p, ld ldx; strideX = 0: one X for the whole batch):
// the range-finder / projection products of the tile compression (HCORE_dcompress_batched_strided).
// grid = (64-row blocks of Y, tiles), 256 threads; warp w owns rows 8 w .. 8 w + 7 of the block.
struct AxArgs {
    const double *A, *X;
    const int *Prk;                      // columns of X per tile (may be null: p for every tile)
    long strideA, strideX, strideY;
    int lda, ldx, ldy, m, n, p;
    double *Y;
};
__global__ void __launch_bounds__(256) dense_ax_kernel(AxArgs ar) {
    const int t = blockIdx.y, p = ar.Prk ? ar.Prk[t] : ar.p;
    if (p <= 0) return;
    const int warp = warp_id(), lane = lane_id(), g = lane >> 2, tg = lane & 3;
    const int i0 = blockIdx.x * 64 + 8 * warp;
    if (i0 >= ar.m) return;
    const double *A = ar.A + (size_t)t * ar.strideA, *X = ar.X + (size_t)t * ar.strideX;
    double *Y = ar.Y + (size_t)t * ar.strideY;
    const bool ron = i0 + g < ar.m;
    const double *ap = A + (i0 + g) + (size_t)tg * ar.lda;
    for (int q0 = 0; q0 < p; q0 += 64) {
        double acc[8][2];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q][0] = acc[q][1] = 0.0;
        for (int k = 0; k < ar.n; k += 4) {
            const double av = (ron && k + tg < ar.n) ? __ldg(ap + (size_t)k * ar.lda) : 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int col = q0 + 8 * q + g;
                if (q0 + 8 * q < p) {
                    const double xv = (col < p && k + tg < ar.n) ? __ldg(X + (size_t)col * ar.ldx + k + tg) : 0.0;
                    dmma_884(acc[q][0], acc[q][1], av, xv);
                }
            }
        }
        if (ron) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = q0 + 8 * q + 2 * tg + e;
                    if (col < p) Y[(i0 + g) + (size_t)col * ar.ldy] = acc[q][e];
                }
            }
        }
    }
}

// deterministic test matrix of the range finder: entries in (-1, 1) from a counter-based hash (splitmix64)
__global__ void fill_omega_kernel(double *X, int n, int p, int ldx, unsigned long long seed) {
    const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= (long)n * p) return;
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    X[(idx % n) + (idx / n) * (long)ldx] = (double)(long long)(z >> 11) * (1.0 / 4503599627370496.0) - 1.0;
}
// E = [I_p; 0] (n x p, ld lde) per tile, and the per-tile rank words of a compression pass
__global__ void fill_identity_kernel(double *E, long strideE, int lde, int n, int p) {
    const int t = blockIdx.y;
    const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= (long)n * p) return;
    const int r = idx % n, c = idx / n;
    E[(size_t)t * strideE + r + (size_t)c * lde] = (r == c) ? 1.0 : 0.0;
}
// the first min(rank, maxrk) columns of the internal factors -> the caller's arrays; rank capped at maxrk
// (reference testing/testing_dgemm.c:202-205: rank > storage_maxrank -> storage_maxrank)
__global__ void compress_copy_out_kernel(const double *CU, const double *CV, long strideI, int ldi, double *U, double *V, long strideO,
                                         int ldo, int nb, const int *rank_in, const int *status_in, int maxrk, int *rank_out,
                                         int *status_out) {
    const int t = blockIdx.y;
    const int st = status_in[t];
    const int r = st == HCORE_B200_OK ? (rank_in[t] < maxrk ? rank_in[t] : maxrk) : 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        rank_out[t] = r;
        if (status_out) status_out[t] = st;
    }
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < (long)nb * r; idx += (long)gridDim.x * blockDim.x) {
        const int row = idx % nb, c = idx / nb;
        U[(size_t)t * strideO + row + (size_t)c * ldo] = CU[(size_t)t * strideI + row + (size_t)c * ldi];
        V[(size_t)t * strideO + row + (size_t)c * ldo] = CV[(size_t)t * strideI + row + (size_t)c * ldi];
    }
}

// CD[I, J] += X[I, :] * Bu[J, :]^T per 128 x 64 block of CD (reference hcore_dgemmcd.c:81-85,
// hcore_dsyrk.c:165-181).  HBM-bound: the block of CD is fetched by TMA bulk copies (one per column,
// 1 KB each) while the product is formed with DMMA, updated in shared memory and written back by TMA
// bulk stores; two CTAs per SM overlap one block's load/store with the other's arithmetic.
// The product is formed TRANSPOSED (DMMA rows = columns of CD, from Bu in shared memory; DMMA columns =
// rows of CD, X fragments straight from global memory into registers: every X element is used by
// exactly one warp), so a lane ends up with two consecutive rows of one column: 16-byte updates.
constexpr int kDuRows = 128, kDuCols = 64, kDuKc = 32;
constexpr int kDuLdC = kDuRows + 8;     // 136: conflict-free 16-byte fragment updates, 16-byte aligned columns
constexpr int kDuLdB = kDuCols + 4;     // 68: conflict-free DMMA fragment loads
struct DenseUpdSmem {
    static constexpr int kC = 0;
    static constexpr int kB = kC + kDuCols * kDuLdC;
    static constexpr int kBar = kB + kDuKc * kDuLdB;
    static constexpr int kTotal = kBar + 2;
    static constexpr size_t kBytes = (size_t)kTotal * sizeof(double);
};

__device__ __forceinline__ void tma_store_1d(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}

// grid = (row blocks of 128, column blocks of 64, tiles), 256 threads.
__global__ void __launch_bounds__(256, 2) dense_update_kernel(DenseArgs ar) {
    using SM = DenseUpdSmem;
    extern __shared__ __align__(16) double smem[];
    double *Cs = smem + SM::kC, *Bs = smem + SM::kB;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM::kBar);      // [0] block of CD, [1] Y chunk
    const int t = blockIdx.z, i0 = blockIdx.x * kDuRows, j0 = blockIdx.y * kDuCols;
    const int m = ar.nb, nn = ar.ncols, ldc = ar.ldc, ldy = ar.ldy;
    const int rows = (m - i0 < kDuRows) ? (m - i0) : kDuRows, cols = (nn - j0 < kDuCols) ? (nn - j0) : kDuCols;
    double *CD = (ar.pCD ? ar.pCD[t] : ar.CD + (size_t)t * ar.strideC) + i0 + (size_t)j0 * ldc;   // block origin
    const double *BU = (ar.pBU ? ar.pBU[t] : ar.Yop + (size_t)t * ar.strideY) + j0;
    const int tid = threadIdx.x, warp = warp_id(), lane = lane_id();
    const int g = lane >> 2, tg = lane & 3;
    const bool cd_bulk = ((m & 1) == 0) && ((ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(CD) & 15) == 0);
    const bool bulk_ok = cd_bulk && ((nn & 1) == 0) && ((ldy & 1) == 0) && ((reinterpret_cast<uintptr_t>(BU) & 15) == 0);
    // A tile flagged by the caller (contraction length < 0: uncompress with mismatched ranks) must come out untouched:
    // it is processed as CD <- 1 * CD + 0.  (Short-circuit: with beta != 0 the rank word is not needed before the loads.)
    const bool readc = ar.beta != 0.0 || ar.Krk[t] < 0;

    if (ar.syrk && j0 > i0 + rows - 1) {      // every entry is strictly above the diagonal: the reference zeroes it
        if (cd_bulk) {
            const int rp = rows >> 1;
            for (int idx = tid; idx < rp * cols; idx += 256)
                *reinterpret_cast<double2 *>(CD + 2 * (idx % rp) + (size_t)(idx / rp) * ldc) = make_double2(0.0, 0.0);
        } else {
            for (int idx = tid; idx < rows * cols; idx += 256) CD[(idx % rows) + (size_t)(idx / rows) * ldc] = 0.0;
        }
        return;
    }
    // ---- the block of CD starts moving towards shared memory ---------------------------------------
    if (bulk_ok) {
        if (tid == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (warp == 0 && readc) {
            if (lane == 0) mbar_expect_tx(&bars[0], (uint32_t)(rows * cols * 8));
            __syncwarp();
            for (int c = lane; c < cols; c += 32) tma_load_1d(Cs + c * kDuLdC, CD + (size_t)c * ldc, (uint32_t)(rows * 8), &bars[0]);
        }
    } else if (readc) {
        for (int idx = tid; idx < rows * cols; idx += 256) {
            const int rr = idx % rows, cc = idx / rows;
            Cs[rr + cc * kDuLdC] = CD[rr + (size_t)cc * ldc];
        }
    }
    // ---- product^T: warp w owns rows 16 w .. 16 w + 15 of the block (two row tiles) x 64 columns ----------
    const int kraw = ar.Krk[t];
    const int b = kraw > 0 ? kraw : 0, b8 = round_up(b, 8);
    const double *X = ar.Xop + (size_t)t * ar.strideX + i0 + 16 * warp + g;
    const bool r0on = 16 * warp + g < rows, r1on = 16 * warp + 8 + g < rows;
    double acc[8][2][2];      // [column tile][row tile][row 2 tg + h]
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) acc[tt][0][0] = acc[tt][0][1] = acc[tt][1][0] = acc[tt][1][1] = 0.0;
    uint32_t bphase = 0;
    for (int k0 = 0; k0 < b8; k0 += kDuKc) {
        const int kc = (b8 - k0 < kDuKc) ? (b8 - k0) : kDuKc;          // multiple of 8
        const int kv = (b - k0 < kc) ? (b - k0) : kc;                  // columns that exist in Y
        if (k0 > 0) __syncthreads();                                   // every warp is done with the previous chunk
        if (bulk_ok) {
            if (warp == 0) {
                if (lane == 0) mbar_expect_tx(&bars[1], (uint32_t)(kv * cols * 8));
                __syncwarp();
                for (int kk = lane; kk < kv; kk += 32)
                    tma_load_1d(Bs + kk * kDuLdB, BU + (size_t)(k0 + kk) * ldy, (uint32_t)(cols * 8), &bars[1]);
            }
            for (int idx = tid; idx < kDuCols * (kc - kv); idx += 256) Bs[(idx & 63) + (kv + (idx >> 6)) * kDuLdB] = 0.0;
            if (cols < kDuCols)                                        // ragged last column block: rows of Y that do not exist
                for (int idx = tid; idx < (kDuCols - cols) * kv; idx += 256)
                    Bs[cols + idx % (kDuCols - cols) + (idx / (kDuCols - cols)) * kDuLdB] = 0.0;
        } else {
            for (int idx = tid; idx < kDuCols * kc; idx += 256) {
                const int rr = idx & (kDuCols - 1), kk = idx >> 6;
                Bs[rr + kk * kDuLdB] = (rr < cols && kk < kv) ? __ldg(BU + rr + (size_t)(k0 + kk) * ldy) : 0.0;
            }
        }
        // X fragments of this chunk: rows 16 w + g and 16 w + 8 + g, columns k0 + 4 ks + tg
        double xa[kDuKc / 4], xb[kDuKc / 4];
#pragma unroll
        for (int ks = 0; ks < kDuKc / 4; ++ks) {
            const bool kon = k0 + 4 * ks + tg < b;
            const double *xp = X + (size_t)(k0 + 4 * ks + tg) * ar.ldx;
            xa[ks] = (kon && r0on) ? __ldg(xp) : 0.0;
            xb[ks] = (kon && r1on) ? __ldg(xp + 8) : 0.0;
        }
        if (bulk_ok) {
            if (kc > kv || cols < kDuCols) __syncthreads();
            mbar_wait(&bars[1], bphase);
            bphase ^= 1;
        } else {
            __syncthreads();
        }
        const double *ba = Bs + g + tg * kDuLdB;
#pragma unroll
        for (int ks = 0; ks < kDuKc / 4; ++ks) {
            if (4 * ks >= kc) break;       // a real exit: predicated-off DMMAs still occupy the FP64 pipe
#pragma unroll
            for (int tt = 0; tt < 8; ++tt) {
                const double bv = ba[4 * ks * kDuLdB + 8 * tt];
                dmma_884(acc[tt][0][0], acc[tt][0][1], bv, xa[ks]);
                dmma_884(acc[tt][1][0], acc[tt][1][1], bv, xb[ks]);
            }
        }
    }
    // ---- update in shared memory, write back -----------------------------------------------------------
    if (bulk_ok && readc) mbar_wait(&bars[0], 0);
    else __syncthreads();
    const double su = ar.su, beta = kraw < 0 ? 1.0 : ar.beta;
    const bool plain = (rows == kDuRows) && (cols == kDuCols) && !(ar.syrk && j0 + cols - 1 > i0);
    if (plain) {
#pragma unroll
        for (int tt = 0; tt < 8; ++tt) {
#pragma unroll
            for (int rt = 0; rt < 2; ++rt) {
                double2 *p = reinterpret_cast<double2 *>(Cs + 16 * warp + 8 * rt + 2 * tg + (8 * tt + g) * kDuLdC);
                double2 v = readc ? *p : make_double2(0.0, 0.0);
                v.x = fma(su, acc[tt][rt][0], beta * v.x);
                v.y = fma(su, acc[tt][rt][1], beta * v.y);
                *p = v;
            }
        }
    } else {
#pragma unroll
        for (int tt = 0; tt < 8; ++tt) {
#pragma unroll
            for (int rt = 0; rt < 2; ++rt) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int rr = 16 * warp + 8 * rt + 2 * tg + h, cc = 8 * tt + g;
                    if (rr < rows && cc < cols) {
                        double *p = Cs + rr + cc * kDuLdC;
                        const double old = readc ? *p : 0.0;
                        *p = (ar.syrk && j0 + cc > i0 + rr) ? 0.0 : fma(su, acc[tt][rt][h], beta * old);
                    }
                }
            }
        }
    }
    if (bulk_ok) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic writes -> visible to the bulk stores
        __syncthreads();
        if (warp == 0) {
            for (int c = lane; c < cols; c += 32) tma_store_1d(CD + (size_t)c * ldc, Cs + c * kDuLdC, (uint32_t)(rows * 8));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the reads
        }
    } else {
        __syncthreads();
        for (int idx = tid; idx < rows * cols; idx += 256) {
            const int rr = idx % rows, cc = idx / rows;
            CD[rr + (size_t)cc * ldc] = Cs[rr + cc * kDuLdC];
        }
    }
}

// ---- persistent streaming form of the same update -----------------------------------------------------------------
// dense_update_kernel gives every 128 x 64 block its own CTA and moves it with one 1 KB bulk copy per column: load,
// product, update and store of a block follow each other, and the copy engine is limited by the NUMBER of requests
// (measured: ~10 k cycles per block, the sum of its HBM time and its FP64 time).  Here persistent CTAs walk
// over (tile, 128-row strip) pairs and streams the strip's 128 x 32 blocks through a ring of shared-memory slots, each
// block moved by ONE tiled TMA request in each direction (cp.async.bulk.tensor.3d on a tensor map of the whole batch:
// rows x columns x tiles):
//     slot q+A  <- tensor load of block q+A            (A = kDsAhead blocks in flight towards the SM)
//     slot q    :  product X Y^T by DMMA (the strip's X fragments stay in registers, the 32 x k piece of Y is staged by
//                  the threads themselves one block ahead), update in shared memory
//     slot q-1  -> tensor store of the finished block
// Every warp owns a band of rows of the strip; the product is formed transposed (DMMA rows = columns of CD) so that a lane
// ends up with two consecutive rows of one column and updates the slot with 16-byte accesses.  The strictly upper
// blocks of HCORE_dsyrk are zeroed with plain stores when a strip starts and never enter the ring.
// Preconditions (checked on the host, dense_stream_ok): strided batch, m % 128 == 0, n % 32 == 0, contraction
// length <= 32, 16-byte aligned bases, even leading dimensions and strides.
// Shape used (hcore_api.cu): two 256-thread CTAs per SM, three slots each, a warp owns 16 rows, contraction lengths up
// to kDsKmax = 32: while one CTA updates and stores, the other one keeps the FP64 pipe busy.
constexpr int kDsRows = 128, kDsCols = 32, kDsKmax = 32;
constexpr int kDsLdY = kDsCols + 4;     // 36
constexpr int kDsBlock = kDsRows * kDsCols;            // doubles per slot: the TMA box, dense (ld = 128)
template <int THREADS, int KMAX, int SLOTS, int NYB>
struct DenseStreamSmem {
    static constexpr int kC = 0;                                            // [slots][32 x 128], 128-byte aligned
    static constexpr int kY = kC + SLOTS * kDsBlock;                        // [NYB][KMAX x 36]  Y piece: [k][column]
    static constexpr int kBar = kY + NYB * KMAX * kDsLdY;                   // full[slots]
    static constexpr int kTotal = kBar + SLOTS;
    static constexpr size_t kBytes = (size_t)kTotal * sizeof(double) + 128; // + slack to align the slots to 128 bytes
};

__device__ __forceinline__ void tma_load_3d(void *dst, const void *tmap, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void *tmap, int c0, int c1, int c2, const void *src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src)) : "memory");
}

// walks over the ring items (blocks that are read-modified-written) of this CTA, in the order every role sees them
struct DenseCursor {
    int t, ib;         // current strip: tile and row block (advanced without divisions)
    int j, n, k;       // next block of the strip, blocks of the strip that enter the ring, contraction length of the tile
};

template <int THREADS, int KMAX, int SLOTS, int AHEAD, int NYB>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 2 : 1)
dense_stream_kernel(DenseArgs ar, int ntiles, const __grid_constant__ CUtensorMap tmap) {
    using SM = DenseStreamSmem<THREADS, KMAX, SLOTS, NYB>;
    constexpr int NWARP = THREADS / 32, RT = 16 / NWARP;             // 8-row tiles per warp
    constexpr int YPT = 16 * KMAX / THREADS;                         // 16-byte pieces of Y per thread
    static_assert(RT * NWARP == 16 && YPT * THREADS == 16 * KMAX && SLOTS > AHEAD, "shape");
    extern __shared__ __align__(16) double smem_raw[];
    double *smem = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + SM::kBar);
    const int tid = threadIdx.x, warp = warp_id(), lane = lane_id();
    const int g = lane >> 2, tg = lane & 3;
    const int m = ar.nb, n = ar.ncols, ldc = ar.ldc, ldy = ar.ldy;
    const int nrb = m / kDsRows, ncb = n / kDsCols;
    const bool readc = ar.beta != 0.0;

    if (tid == 0) {
        for (int sl = 0; sl < SLOTS; ++sl) mbar_init(&full[sl], 1);
        fence_mbar_init();
    }
    // stale rows of a Y buffer are multiplied by zero X entries: they only have to be finite
    for (int idx = tid; idx < NYB * KMAX * kDsLdY; idx += THREADS) smem[SM::kY + idx] = 0.0;
    __syncthreads();

    const int dt = gridDim.x / nrb, dib = gridDim.x % nrb;           // strip stride of the CTA as (tiles, row blocks)
    auto start = [&]() { return DenseCursor{(int)blockIdx.x / nrb, (int)blockIdx.x % nrb, 0, 0, 0}; };
    auto advance = [&](DenseCursor &c) {            // next strip of this CTA
        c.t += dt; c.ib += dib;
        if (c.ib >= nrb) { c.ib -= nrb; ++c.t; }
    };
    auto enter = [&](DenseCursor &c) {              // first strip at or after the current one with blocks for the ring
        for (; c.t < ntiles; advance(c)) {
            c.k = ar.Krk[c.t];
            c.j = 0;
            // flagged tile (k < 0) or nothing to add: CD untouched (dsyrk still zeroes its upper triangle)
            c.n = (c.k < 0 || (c.k == 0 && ar.beta == 1.0 && !ar.syrk)) ? 0 : (ar.syrk ? min(ncb, 4 * (c.ib + 1)) : ncb);
            if (c.n > 0) return;
        }
    };
    auto step = [&](DenseCursor &c) {               // to the next ring item
        if (++c.j >= c.n) { advance(c); enter(c); }
    };
    // ---- producer (thread 0): tensor loads AHEAD items ahead ---------------------------------------------------------------
    DenseCursor pc = start();
    int pq = 0;
    auto produce = [&]() {
        if (pc.t >= ntiles) return;
        if (readc) {
            const int sl = pq % SLOTS;
            mbar_expect_tx(&full[sl], kDsBlock * 8);
            tma_load_3d(smem + SM::kC + sl * kDsBlock, &tmap, pc.ib * kDsRows, pc.j * kDsCols, pc.t, &full[sl]);
        }
        ++pq;
        step(pc);
    };
    if (tid == 0) {
        enter(pc);
        for (int a = 0; a < AHEAD; ++a) produce();
    }
    // ---- Y piece of a ring item: 32 columns x k, YPT 16-byte values per thread, fetched one item ahead -----------------------
    DenseCursor yc = start();
    enter(yc);
    double2 yv[YPT];
    auto y_fetch = [&]() {
        if (yc.t >= ntiles || yc.k <= 0) return;
        const double *Y = ar.Yop + (size_t)yc.t * ar.strideY + yc.j * kDsCols;
#pragma unroll
        for (int r = 0; r < YPT; ++r) {
            // rows >= k are not staged (y_stage): read a valid row instead of selecting after the load, so that nothing
            // depends on the loaded value before the product has been issued
            const int idx = tid + THREADS * r, kk = min(idx >> 4, yc.k - 1), cp = idx & 15;
            yv[r] = __ldg(reinterpret_cast<const double2 *>(Y + (size_t)kk * ldy) + cp);
        }
    };
    auto y_stage = [&](int buf) {
        if (yc.t >= ntiles) return;
        double *Ys = smem + SM::kY + buf * KMAX * kDsLdY;
#pragma unroll
        for (int r = 0; r < YPT; ++r) {
            const int idx = tid + THREADS * r, kk = idx >> 4, cp = idx & 15;
            if (kk < yc.k) *reinterpret_cast<double2 *>(Ys + kk * kDsLdY + 2 * cp) = yv[r];
        }
        step(yc);
    };
    y_fetch();
    y_stage(0);
    __syncthreads();

    // ---- consumer ------------------------------------------------------------------------------------------------
    int cq = 0;
    PROF_DECL(tid == 0)
    for (DenseCursor cc0 = start(); cc0.t < ntiles; advance(cc0)) {
        const int t = cc0.t, ib = cc0.ib, i0 = ib * kDsRows;
        const int k = ar.Krk[t];
        const int nblk = (k < 0 || (k == 0 && ar.beta == 1.0 && !ar.syrk)) ? 0 : (ar.syrk ? min(ncb, 4 * (ib + 1)) : ncb);
        double *CDt = ar.CD + (size_t)t * ar.strideC;
        if (ar.syrk && k >= 0) {                    // strictly upper blocks of the strip: zero (reference hcore_dsyrk.c:174-181)
            const int jz = kDsRows * (ib + 1);      // first column that is above the diagonal for every row of the strip
            for (long idx = tid; idx < (long)(kDsRows / 2) * (n - jz); idx += THREADS) {
                const int rp = (int)(idx % (kDsRows / 2)), c = jz + (int)(idx / (kDsRows / 2));
                *reinterpret_cast<double2 *>(CDt + i0 + 2 * rp + (size_t)c * ldc) = make_double2(0.0, 0.0);
            }
        }
        if (nblk == 0) continue;
        const int nks = (k + 3) >> 2;
        // X fragments of the strip: rows i0 + 8 (RT warp + rt) + g, columns 4 ks + tg
        double xr[RT][KMAX / 4];
        {
            const double *X = ar.Xop + (size_t)t * ar.strideX + i0 + 8 * RT * warp + g;
#pragma unroll
            for (int rt = 0; rt < RT; ++rt)
#pragma unroll
                for (int ks = 0; ks < KMAX / 4; ++ks)
                    xr[rt][ks] = (4 * ks + tg < k) ? __ldg(X + 8 * rt + (size_t)(4 * ks + tg) * ar.ldx) : 0.0;
        }
        for (int jb = 0; jb < nblk; ++jb, ++cq) {
            const int sl = cq % SLOTS, j0 = jb * kDsCols;
            double *Cs = smem + SM::kC + sl * kDsBlock;
            const double *Ys = smem + SM::kY + (NYB == 2 ? (cq & 1) : 0) * KMAX * kDsLdY;
            PROF_ADD(56);
            y_fetch();                              // Y piece of the next item: in flight during the product
            double acc[4][RT][2];
#pragma unroll
            for (int tt = 0; tt < 4; ++tt)
#pragma unroll
                for (int rt = 0; rt < RT; ++rt) acc[tt][rt][0] = acc[tt][rt][1] = 0.0;
            const double *ya = Ys + g + tg * kDsLdY;
            // a real (warp-uniform) exit, not predication: predicated-off DMMAs still occupy the FP64 pipe
#pragma unroll
            for (int ks = 0; ks < KMAX / 4; ++ks) {
                if (ks >= nks) break;
#pragma unroll
                for (int tt = 0; tt < 4; ++tt) {
                    const double a = ya[4 * ks * kDsLdY + 8 * tt];
#pragma unroll
                    for (int rt = 0; rt < RT; ++rt) dmma_884(acc[tt][rt][0], acc[tt][rt][1], a, xr[rt][ks]);
                }
            }
            PROF_ADD(57);
            if (NYB == 1) __syncthreads();          // every warp has read the Y piece: the next one may replace it
            if (readc) mbar_wait(&full[sl], (uint32_t)((cq / SLOTS) & 1));
            PROF_ADD(58);
            const double su = ar.su, beta = ar.beta;
            const bool diag = ar.syrk && j0 + kDsCols - 1 > i0;        // the block reaches above the diagonal
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) {
#pragma unroll
                for (int rt = 0; rt < RT; ++rt) {
                    const int cc = 8 * tt + g, rr = 8 * (RT * warp + rt) + 2 * tg;
                    double2 *p = reinterpret_cast<double2 *>(Cs + rr + cc * kDsRows);
                    double2 v = readc ? *p : make_double2(0.0, 0.0);
                    v.x = fma(su, acc[tt][rt][0], beta * v.x);
                    v.y = fma(su, acc[tt][rt][1], beta * v.y);
                    if (diag) {
                        if (j0 + cc > i0 + rr) v.x = 0.0;
                        if (j0 + cc > i0 + rr + 1) v.y = 0.0;
                    }
                    *p = v;
                }
            }
            PROF_ADD(59);
            y_stage(NYB == 2 ? ((cq + 1) & 1) : 0);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic accesses of this slot -> async proxy
            __syncthreads();
            PROF_ADD(60);
            PROF_COUNT(63, 1);
            if (tid == 0) {
                tma_store_3d(&tmap, i0, j0, t, Cs);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                // the slot refilled now is the one of item q + AHEAD - SLOTS: its store (all groups but the SLOTS - AHEAD
                // most recent ones) must have finished reading shared memory
                asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(SLOTS - AHEAD) : "memory");
                PROF_ADD(61);
                produce();
                PROF_ADD(62);
            }
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- complex uncompress: CD = AU BV^T (plain transpose), m x n, contraction Ark -------------------------------
// reference src/compute/hcore_zuncompress.c:35-95 (cblas_zgemm NoTrans / Trans with alpha = 1, beta = 0: the caller's
// alpha and beta are ignored).  grid = (64-row blocks, 64-column blocks, tiles), 256 threads; warp w owns rows
// 8 w .. 8 w + 7 of the block and all 64 columns; the real and imaginary planes of the BV block sit in shared memory.
struct ZUncArgs {
    const double2 *AU, *BV;
    const int *Ark;
    long strideA, strideB, strideC;
    int lda, ldb, ldc, m, n;
    double2 *CD;
};
constexpr int kZuK = 32, kZuLd = 64 + 4;
__global__ void __launch_bounds__(256, 2) zdense_update_kernel(ZUncArgs ar) {
    __shared__ __align__(16) double Br[kZuK * kZuLd], Bi[kZuK * kZuLd];
    const int t = blockIdx.z, i0 = blockIdx.x * 64, j0 = blockIdx.y * 64;
    const int tid = threadIdx.x, warp = warp_id(), lane = lane_id(), g = lane >> 2, tg = lane & 3;
    if (ar.Ark[t] < 0) return;                 // mismatched ranks: tile left untouched
    const int k = ar.Ark[t];
    const double2 *AU = ar.AU + (size_t)t * ar.strideA, *BV = ar.BV + (size_t)t * ar.strideB;
    double2 *CD = ar.CD + (size_t)t * ar.strideC;
    const int row = i0 + 8 * warp + g;
    double cr[8][2], ci[8][2];
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) cr[tt][0] = cr[tt][1] = ci[tt][0] = ci[tt][1] = 0.0;
    for (int k0 = 0; k0 < k; k0 += kZuK) {
        __syncthreads();
        for (int idx = tid; idx < 64 * kZuK; idx += 256) {
            const int c = idx & 63, kk = idx >> 6;
            double2 v = make_double2(0.0, 0.0);
            if (j0 + c < ar.n && k0 + kk < k) v = __ldg(BV + (j0 + c) + (size_t)(k0 + kk) * ar.ldb);
            Br[c + kk * kZuLd] = v.x;
            Bi[c + kk * kZuLd] = v.y;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < kZuK / 4; ++ks) {
            double2 a = make_double2(0.0, 0.0);
            if (row < ar.m && k0 + 4 * ks + tg < k) a = __ldg(AU + row + (size_t)(k0 + 4 * ks + tg) * ar.lda);
            const double nai = -a.y;
#pragma unroll
            for (int tt = 0; tt < 8; ++tt) {
                const double br = Br[8 * tt + g + (4 * ks + tg) * kZuLd], bi = Bi[8 * tt + g + (4 * ks + tg) * kZuLd];
                dmma_884(cr[tt][0], cr[tt][1], a.x, br);
                dmma_884(cr[tt][0], cr[tt][1], nai, bi);
                dmma_884(ci[tt][0], ci[tt][1], a.x, bi);
                dmma_884(ci[tt][0], ci[tt][1], a.y, br);
            }
        }
    }
    if (row < ar.m) {
#pragma unroll
        for (int tt = 0; tt < 8; ++tt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = j0 + 8 * tt + 2 * tg + e;
                if (col < ar.n) CD[row + (size_t)col * ar.ldc] = make_double2(cr[tt][e], ci[tt][e]);
            }
        }
    }
}

}  // namespace hcore_b200
