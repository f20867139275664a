// Dense-C updates of the TLR Cholesky trailing update (FP64, sm_100a):
//   HCORE_dgemmcd  CD <- CD - (Au Av^T)(Bu Bv^T)^T      reference src/compute/hcore_dgemmcd.c:36-107
//   HCORE_dsyrk    CD <- CD + alpha (AU AV^T)(AU AV^T)^T, strict upper triangle zeroed
//                                                         reference src/compute/hcore_dsyrk.c:47-189
// Both are a rank-k update that streams the dense nb x nb tile once: HBM-bound
// (2 * 8 * nb^2 bytes per tile).  Two kernels:
//   dense_prep_kernel    X = s * Au * (Av^T Bv)   (nb x b, tiny) into the context workspace
//   dense_update_kernel  CD[I, J] += X[I, :] * Bu[J, :]^T   per 128 x 64 block of CD (TMA in / out)
#pragma once
#include "common.cuh"
#include "wy_kernels.cuh"

namespace hcore_b200 {

struct DenseArgs {
    const double *AU, *AV, *BU, *BV;   // strided factor arrays (BU/BV == AU/AV for syrk)
    const int *Ark, *Brk;
    long strideA, strideB, strideC;
    double *CD;
    double *X;                           // workspace [ntiles][m_pad * kmax8]
    double scale;                        // -1 for gemmcd, alpha for syrk
    int nb, m_pad, kmax8;
    int syrk;                            // 1: lower triangle only + zero strict upper
};

// X = scale * Au (Av^T Bv)  (m x b, ld m_pad)      reference hcore_dgemmcd.c:66-78, hcore_dsyrk.c:103-144.
// Same shape of work as the P/G products of the low-rank path (pg_body, wy_kernels.cuh) with the roles
// swapped: P = Bv^T Av, X = Au P^T.  One CTA (256 threads) per tile, operands staged by cp.async.
__global__ void __launch_bounds__(256, 2) dense_prep_kernel(DenseArgs ar) {
    extern __shared__ __align__(16) double smem[];
    const int t = blockIdx.x;
    const int a = ar.Ark[t], b = ar.Brk[t], m = ar.nb;
    if (a <= 0 || b <= 0) return;
    const double *AU = ar.AU + (size_t)t * ar.strideA, *AV = ar.AV + (size_t)t * ar.strideA;
    const double *BV = ar.BV + (size_t)t * ar.strideB;
    pg_body(smem, BV, AU, AV, b, a, m, m, m, ar.m_pad, 0, ar.scale, ar.X + (size_t)t * ar.m_pad * ar.kmax8);
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
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM::kBar);      // [0] block of CD, [1] Bu chunk
    const int t = blockIdx.z, i0 = blockIdx.x * kDuRows, j0 = blockIdx.y * kDuCols;
    const int m = ar.nb;
    const int rows = (m - i0 < kDuRows) ? (m - i0) : kDuRows, cols = (m - j0 < kDuCols) ? (m - j0) : kDuCols;
    double *CD = ar.CD + (size_t)t * ar.strideC + i0 + (size_t)j0 * m;     // block origin, ld m
    const double *BU = ar.BU + (size_t)t * ar.strideB + j0;
    const int tid = threadIdx.x, warp = warp_id(), lane = lane_id();
    const int g = lane >> 2, tg = lane & 3;
    const bool bulk_ok = ((m & 1) == 0) && ((reinterpret_cast<uintptr_t>(CD) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(BU) & 15) == 0);

    if (ar.syrk && j0 > i0 + rows - 1) {      // every entry is strictly above the diagonal: the reference zeroes it
        if (((m & 1) == 0) && ((reinterpret_cast<uintptr_t>(CD) & 15) == 0)) {
            const int rp = rows >> 1;
            for (int idx = tid; idx < rp * cols; idx += 256)
                *reinterpret_cast<double2 *>(CD + 2 * (idx % rp) + (size_t)(idx / rp) * m) = make_double2(0.0, 0.0);
        } else {
            for (int idx = tid; idx < rows * cols; idx += 256) CD[(idx % rows) + (size_t)(idx / rows) * m] = 0.0;
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
        if (warp == 0) {
            if (lane == 0) mbar_expect_tx(&bars[0], (uint32_t)(rows * cols * 8));
            __syncwarp();
            for (int c = lane; c < cols; c += 32) tma_load_1d(Cs + c * kDuLdC, CD + (size_t)c * m, (uint32_t)(rows * 8), &bars[0]);
        }
    } else {
        for (int idx = tid; idx < rows * cols; idx += 256) {
            const int rr = idx % rows, cc = idx / rows;
            Cs[rr + cc * kDuLdC] = CD[rr + (size_t)cc * m];
        }
    }
    // ---- product^T: warp w owns rows 16 w .. 16 w + 15 of the block (two row tiles) x 64 columns ----------
    const int b = ar.Brk[t], b8 = round_up(b, 8);
    const double *X = ar.X + (size_t)t * ar.m_pad * ar.kmax8 + i0 + 16 * warp + g;
    const bool r0on = 16 * warp + g < rows, r1on = 16 * warp + 8 + g < rows;
    double acc[8][2][2];      // [column tile][row tile][row 2 tg + h]
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) acc[tt][0][0] = acc[tt][0][1] = acc[tt][1][0] = acc[tt][1][1] = 0.0;
    uint32_t bphase = 0;
    for (int k0 = 0; k0 < b8; k0 += kDuKc) {
        const int kc = (b8 - k0 < kDuKc) ? (b8 - k0) : kDuKc;          // multiple of 8
        const int kv = (b - k0 < kc) ? (b - k0) : kc;                  // columns that exist in Bu
        if (k0 > 0) __syncthreads();                                   // every warp is done with the previous chunk
        if (bulk_ok) {
            if (warp == 0) {
                if (lane == 0) mbar_expect_tx(&bars[1], (uint32_t)(kv * cols * 8));
                __syncwarp();
                for (int kk = lane; kk < kv; kk += 32)
                    tma_load_1d(Bs + kk * kDuLdB, BU + (size_t)(k0 + kk) * m, (uint32_t)(cols * 8), &bars[1]);
            }
            for (int idx = tid; idx < kDuCols * (kc - kv); idx += 256) Bs[(idx & 63) + (kv + (idx >> 6)) * kDuLdB] = 0.0;
        } else {
            for (int idx = tid; idx < kDuCols * kc; idx += 256) {
                const int rr = idx & (kDuCols - 1), kk = idx >> 6;
                Bs[rr + kk * kDuLdB] = (rr < cols && kk < kv) ? __ldg(BU + rr + (size_t)(k0 + kk) * m) : 0.0;
            }
        }
        // X fragments of this chunk: rows 16 w + g and 16 w + 8 + g, columns k0 + 4 ks + tg
        double xa[kDuKc / 4], xb[kDuKc / 4];
#pragma unroll
        for (int ks = 0; ks < kDuKc / 4; ++ks) {
            const bool kon = 4 * ks < kc;
            const double *xp = X + (size_t)(k0 + 4 * ks + tg) * ar.m_pad;
            xa[ks] = (kon && r0on) ? __ldg(xp) : 0.0;
            xb[ks] = (kon && r1on) ? __ldg(xp + 8) : 0.0;
        }
        if (bulk_ok) {
            if (kc > kv) __syncthreads();
            mbar_wait(&bars[1], bphase);
            bphase ^= 1;
        } else {
            __syncthreads();
        }
        const double *ba = Bs + g + tg * kDuLdB;
#pragma unroll
        for (int ks = 0; ks < kDuKc / 4; ++ks) {
            if (4 * ks < kc) {
#pragma unroll
                for (int tt = 0; tt < 8; ++tt) {
                    const double bv = ba[4 * ks * kDuLdB + 8 * tt];
                    dmma_884(acc[tt][0][0], acc[tt][0][1], bv, xa[ks]);
                    dmma_884(acc[tt][1][0], acc[tt][1][1], bv, xb[ks]);
                }
            }
        }
    }
    // ---- update in shared memory, write back -----------------------------------------------------------
    if (bulk_ok) mbar_wait(&bars[0], 0);
    else __syncthreads();
    const bool plain = (rows == kDuRows) && (cols == kDuCols) && !(ar.syrk && j0 + cols - 1 > i0);
    if (plain) {
#pragma unroll
        for (int tt = 0; tt < 8; ++tt) {
#pragma unroll
            for (int rt = 0; rt < 2; ++rt) {
                double2 *p = reinterpret_cast<double2 *>(Cs + 16 * warp + 8 * rt + 2 * tg + (8 * tt + g) * kDuLdC);
                double2 v = *p;
                v.x += acc[tt][rt][0];
                v.y += acc[tt][rt][1];
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
                        *p = (ar.syrk && j0 + cc > i0 + rr) ? 0.0 : *p + acc[tt][rt][h];
                    }
                }
            }
        }
    }
    if (bulk_ok) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic writes -> visible to the bulk stores
        __syncthreads();
        if (warp == 0) {
            for (int c = lane; c < cols; c += 32) tma_store_1d(CD + (size_t)c * m, Cs + c * kDuLdC, (uint32_t)(rows * 8));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the reads
        }
    } else {
        __syncthreads();
        for (int idx = tid; idx < rows * cols; idx += 256) {
            const int rr = idx % rows, cc = idx / rows;
            CD[rr + (size_t)cc * m] = Cs[rr + cc * kDuLdC];
        }
    }
}

}  // namespace hcore_b200
