// Dense-C updates of the TLR Cholesky trailing update (FP64, sm_100a):
//   HCORE_dgemmcd  CD <- CD - (Au Av^T)(Bu Bv^T)^T      reference src/compute/hcore_dgemmcd.c:36-107
//   HCORE_dsyrk    CD <- CD + alpha (AU AV^T)(AU AV^T)^T, strict upper triangle zeroed
//                                                         reference src/compute/hcore_dsyrk.c:47-189
// Both are a rank-k update that streams the dense nb x nb tile once: HBM-bound
// (2 * 8 * nb^2 bytes per tile).  Two kernels:
//   dense_prep_kernel    X = s * Au * (Av^T Bv)   (nb x b, tiny) into the context workspace
//   dense_update_kernel  CD[I, J] += X[I, :] * Bu[J, :]^T   per 128 x 128 block of CD
#pragma once
#include "common.cuh"

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

constexpr int kDenseLdP = HCORE_B200_MAX_STACK + 4;

// One CTA (256 threads) per tile.  Shared: W (a8 x b8, ld kDenseLdP).
__global__ void __launch_bounds__(256) dense_prep_kernel(DenseArgs ar) {
    extern __shared__ double smem[];
    double *W = smem;
    const int t = blockIdx.x;
    const int a = ar.Ark[t], b = ar.Brk[t], m = ar.nb;
    const double *AU = ar.AU + (size_t)t * ar.strideA, *AV = ar.AV + (size_t)t * ar.strideA;
    const double *BV = ar.BV + (size_t)t * ar.strideB;
    double *X = ar.X + (size_t)t * ar.m_pad * ar.kmax8;
    const int a8 = round_up(a, 8), b8 = round_up(b, 8), m4 = round_up(m, 4), m8 = round_up(m, 8);
    const int warp = warp_id(), nwarps = blockDim.x >> 5;
    // W = Av^T * Bv  (reference hcore_dgemmcd.c:66-70, hcore_dsyrk.c:103)
    GuardedView Av{AV, (long)m, 1, a, m};
    GuardedView Bv{BV, 1, (long)m, m, b};
    block_gemm<2, kGemmSet>(mview(W, 1, kDenseLdP), Av, Bv, a8, b8, m4, warp, nwarps);
    __syncthreads();
    if (ar.scale != 1.0) {
        for (int idx = threadIdx.x; idx < a8 * b8; idx += blockDim.x)
            W[(idx % a8) + (idx / a8) * kDenseLdP] *= ar.scale;
        __syncthreads();
    }
    // X = Au * W  (m x b)   (reference hcore_dgemmcd.c:74-78, hcore_dsyrk.c:144)
    GuardedView Au{AU, 1, (long)m, m, a};
    block_gemm<2, kGemmSet>(mview(X, 1, ar.m_pad), Au, cview(W, 1, kDenseLdP), m8, b8, a8, warp, nwarps);
}

constexpr int kDenseKc = 64;   // contraction chunk staged in shared memory

// grid = (row blocks, col blocks, tiles), 256 threads; 128 x 128 block of CD per CTA.
// Shared: Xs and Bs (128 x kDenseKc each, ld 132), refilled per contraction chunk.
__global__ void __launch_bounds__(256) dense_update_kernel(DenseArgs ar) {
    constexpr int LD = kBlockH + 4;
    extern __shared__ double smem[];
    const int t = blockIdx.z, bi = blockIdx.x, bj = blockIdx.y;
    const int m = ar.nb;
    double *CD = ar.CD + (size_t)t * ar.strideC;
    const int i0 = bi * kBlockH, j0 = bj * kBlockH;
    const int tid = threadIdx.x;
    if (ar.syrk && bj > bi) {   // block strictly above the diagonal: reference sets it to zero
        for (int idx = tid; idx < kBlockH * kBlockH; idx += blockDim.x) {
            const int rr = i0 + (idx & 127), cc = j0 + (idx >> 7);
            if (rr < m && cc < m) CD[rr + (size_t)cc * m] = 0.0;
        }
        return;
    }
    const int b = ar.Brk[t];
    const int b8 = round_up(b, 8);
    double *Xs = smem, *Bs = smem + (size_t)kDenseKc * LD;
    const double *X = ar.X + (size_t)t * ar.m_pad * ar.kmax8;
    const double *BU = ar.BU + (size_t)t * ar.strideB;
    const int lane = lane_id(), g = lane >> 2, tg = lane & 3;
    const int warp = warp_id();
    // CD[I, J] += Xs * Bs^T  (reference hcore_dgemmcd.c:81-85, hcore_dsyrk.c:165-171)
    // 8 warps x 2 strips: each warp keeps two 8 x 128 accumulator strips in registers.
    double acc[2][16][2];
#pragma unroll
    for (int sidx = 0; sidx < 2; ++sidx)
#pragma unroll
        for (int tt = 0; tt < 16; ++tt) acc[sidx][tt][0] = acc[sidx][tt][1] = 0.0;
    for (int k0 = 0; k0 < b8; k0 += kDenseKc) {
        const int kc = (b8 - k0 < kDenseKc) ? (b8 - k0) : kDenseKc;
        __syncthreads();
        for (int idx = tid; idx < kBlockH * kc; idx += blockDim.x) {
            const int rr = idx & 127, kk = idx >> 7;
            Xs[rr + kk * LD] = (i0 + rr < m) ? X[i0 + rr + (size_t)(k0 + kk) * ar.m_pad] : 0.0;
            Bs[rr + kk * LD] = (j0 + rr < m && k0 + kk < b) ? BU[j0 + rr + (size_t)(k0 + kk) * m] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int sidx = 0; sidx < 2; ++sidx) {
            const int r0 = (warp + 8 * sidx) * 8;
            for (int k = 0; k < kc; k += 4) {
                const double av = Xs[r0 + g + (k + tg) * LD];
#pragma unroll
                for (int tt = 0; tt < 16; ++tt) {
                    const double bv = Bs[tt * 8 + g + (k + tg) * LD];
                    dmma_884(acc[sidx][tt][0], acc[sidx][tt][1], av, bv);
                }
            }
        }
    }
#pragma unroll
    for (int sidx = 0; sidx < 2; ++sidx) {
        const int rr = i0 + (warp + 8 * sidx) * 8 + g;
        if (rr < m) {
#pragma unroll
            for (int tt = 0; tt < 16; ++tt) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int cc = j0 + tt * 8 + 2 * tg + h;
                    if (cc < m) {
                        double *p = CD + rr + (size_t)cc * m;
                        if (ar.syrk && cc > rr) *p = 0.0;
                        else *p += acc[sidx][tt][h];
                    }
                }
            }
        }
    }
}

}  // namespace hcore_b200
