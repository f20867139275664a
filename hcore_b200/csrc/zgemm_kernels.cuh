// Complex-double tile-low-rank update + recompression (HCORE_zgemm, reference
// src/compute/hcore_zgemm.c:663-975).
//
// Same pipeline as the FP64 path in complex arithmetic.  Every product runs on the FP64 tensor pipe (four real DMMAs per
// complex 8 x 8 x 4 step); operands stay in global memory / L2 (the complex 1024 x 128 panel is 2 MiB, the 128 x 128
// core 256 KiB: neither fits one CTA's shared memory):
//   ztlr_stack_kernel      X_u = [beta CU | alpha AU], P = AV^T BV, X_v = [CV | BU P^T]   (__zqra/__zqrb,
//                          hcore_zgemm.c:65-137, :138-215; PLAIN transposes, :165, :189)
//   ztlr_qr_blocked_kernel Householder QR of both panels in 8-column blocks (zgeqr2 order inside a block, zlarft factor T
//                          per block, trailing columns updated eight at a time on DMMA); ztlr_qr_kernel (unblocked) for
//                          nb > 1024 (replaces LAPACKE_zgeqrf, hcore_zgemm.c:123, :205)
//   ztlr_svd_kernel        T = R_u R_v^T (plain transpose, :309), truncated pivoted QR of T^H as preconditioner, one-sided
//                          Jacobi on the columns of R^H with complex rotations, sort, truncation (rk > 0 = fixed rank, :376),
//                          U_r, W = R_v (R_u^T conj(U_r))  (replaces LAPACKE_zgesvd + the conj/zunmqr/conj dance
//                          of __znewv, :593-632: CV = Q_v W)
//   ztlr_apply_wy_kernel   CU = Q_u [U_r; 0], CV = Q_v [W; 0]: the reflectors applied eight at a time through the T factors
//                          of the QR kernel, 8 output columns per CTA in DMMA accumulator fragments; ztlr_apply_kernel
//                          (one column per warp in shared memory) after the unblocked QR
//                          (replaces LAPACKE_zunmqr, :470-530, :600-620)
#pragma once
#include "common.cuh"
#include "wy_kernels.cuh"

namespace hcore_b200 {

typedef double2 zc;   // x = re, y = im

__device__ __forceinline__ zc zmake(double r, double i) { return make_double2(r, i); }
__device__ __forceinline__ zc zadd(zc a, zc b) { return zmake(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ zc zsub(zc a, zc b) { return zmake(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ zc zmul(zc a, zc b) { return zmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ zc zconj(zc a) { return zmake(a.x, -a.y); }
__device__ __forceinline__ zc zscale(zc a, double s) { return zmake(a.x * s, a.y * s); }
__device__ __forceinline__ zc zfma(zc a, zc b, zc c) {   // a*b + c
    return zmake(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ zc zcfma(zc a, zc b, zc c) {  // conj(a)*b + c
    return zmake(fma(a.x, b.x, fma(a.y, b.y, c.x)), fma(a.x, b.y, fma(-a.y, b.x, c.y)));
}
__device__ __forceinline__ double zabs2(zc a) { return fma(a.x, a.x, a.y * a.y); }
__device__ __forceinline__ zc zwarp_sum(zc v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}

struct ZArgs {
    const zc *AU, *AV, *BU, *BV;     // strided factor arrays
    zc *CU, *CV;
    long strideA, strideB, strideC;
    const zc *const *ptrs;           // pointer-array form: [6][ntiles_total] tables (AU, AV, BU, BV, CU, CV), else null
    long ptr_stride;                 // entries per table
    const int *Ark, *Brk;
    int *Crk;
    int *status;
    int nb, lda, ldb, ldc;
    zc alpha, beta;
    int rk, maxrk;
    double acc;
    int tile0, ntiles;
};

struct ZWorkspace {
    zc *panel;     // [2][chunk][m * N]     stacked panels -> reflectors + R
    zc *tau;       // [2][chunk][N]
    zc *tfac;      // [2][chunk][N / 8][8 x 8]  compact-WY triangular factors of the 8-reflector blocks (blocked QR)
    zc *T;         // [chunk][N * N]        P, then the core matrix, then a GEMM temporary
    zc *Y;         // [2][chunk][N * N]     U_r and W
    double *sigma; // [chunk][N]
    int *ncols, *newrank, *st, *sweeps;
    int chunk, N;
};

__device__ __forceinline__ zc *zpanel(const ZWorkspace &ws, int pnl, int t, int m) {
    return ws.panel + ((size_t)pnl * ws.chunk + t) * (size_t)m * ws.N;
}

// 256-bit global accesses (two consecutive complex numbers, 32-byte aligned: sm_100 LDG/STG.256)
__device__ __forceinline__ void ldg_z2(const zc *p, zc &a, zc &b) {
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
}
__device__ __forceinline__ void stg_z2(zc *p, zc a, zc b) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}

// Block-level complex product C(i, j) = sum_{k in [klo, khi)} A(i, k) B(k, j) on the FP64 tensor pipe (four real DMMAs
// per complex 8 x 8 x 4 step).  fa(i, k), fb(k, j) return the operand entries (zero outside their extent), put(i, j, v)
// stores a result entry, klim(i0, j0, klo, khi) may shrink the contraction range of an 8 x 8 output tile (triangular
// operands; multiples of 4).  Output tiles are dealt round-robin to the warps; four contraction steps of loads in flight.
template <typename FA, typename FB, typename PUT, typename KLIM>
__device__ __forceinline__ void zblock_gemm(int M, int Nc, int K, FA fa, FB fb, PUT put, KLIM klim, int warp, int nwarps, int lane) {
    const int g = lane >> 2, tg = lane & 3;
    const int tm = (M + 7) >> 3, tn = (Nc + 7) >> 3;
    for (int tile = warp; tile < tm * tn; tile += nwarps) {
        const int i0 = (tile % tm) << 3, j0 = (tile / tm) << 3;
        int klo = 0, khi = (K + 3) & ~3;
        klim(i0, j0, klo, khi);
        double cr[2] = {0.0, 0.0}, ci[2] = {0.0, 0.0};
        for (int k = klo; k < khi; k += 16) {
            zc av[4], bv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {              // four contraction steps of operand loads in flight
                const bool kon = k + 4 * u < khi;
                av[u] = kon ? fa(i0 + g, k + 4 * u + tg) : zmake(0.0, 0.0);
                bv[u] = kon ? fb(k + 4 * u + tg, j0 + g) : zmake(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (k + 4 * u >= khi) break;           // warp-uniform: a real exit, not predicated DMMAs
                dmma_884(cr[0], cr[1], av[u].x, bv[u].x);
                dmma_884(cr[0], cr[1], -av[u].y, bv[u].y);
                dmma_884(ci[0], ci[1], av[u].x, bv[u].y);
                dmma_884(ci[0], ci[1], av[u].y, bv[u].x);
            }
        }
        put(i0 + g, j0 + 2 * tg, zmake(cr[0], ci[0]));
        put(i0 + g, j0 + 2 * tg + 1, zmake(cr[1], ci[1]));
    }
}

// ---- 1. stacking ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) ztlr_stack_kernel(ZArgs ar, ZWorkspace ws) {      // 128 registers: two CTAs per SM (the kernel waits on global loads)
    const int t = blockIdx.x, gt = ar.tile0 + t, tid = threadIdx.x;
    const int a = ar.Ark[gt], b = ar.Brk[gt], c = ar.Crk[gt];
    const int n = a + c, m = ar.nb, N = ws.N;
    int st = HCORE_B200_OK;
    if (a <= 0 || b <= 0 || c <= 0) st = HCORE_B200_ERR_ZERO_RANK;
    else if (ar.maxrk < n) st = HCORE_B200_ERR_MAXRK;
    else if (n > N || n > m || b > N) st = HCORE_B200_ERR_UNSUPPORTED;
    if (tid == 0) {
        ws.st[t] = st; ws.ncols[t] = n; ws.newrank[t] = c;
        if (ar.status) ar.status[gt] = st;
    }
    if (st != HCORE_B200_OK) return;
    const zc *AU = ar.ptrs ? ar.ptrs[gt] : ar.AU + (size_t)gt * ar.strideA;
    const zc *AV = ar.ptrs ? ar.ptrs[ar.ptr_stride + gt] : ar.AV + (size_t)gt * ar.strideA;
    const zc *BU = ar.ptrs ? ar.ptrs[2 * ar.ptr_stride + gt] : ar.BU + (size_t)gt * ar.strideB;
    const zc *BV = ar.ptrs ? ar.ptrs[3 * ar.ptr_stride + gt] : ar.BV + (size_t)gt * ar.strideB;
    const zc *CU = ar.ptrs ? ar.ptrs[4 * ar.ptr_stride + gt] : ar.CU + (size_t)gt * ar.strideC;
    const zc *CV = ar.ptrs ? ar.ptrs[5 * ar.ptr_stride + gt] : ar.CV + (size_t)gt * ar.strideC;
    zc *XU = zpanel(ws, 0, t, m), *XV = zpanel(ws, 1, t, m);
    zc *P = ws.T + (size_t)t * N * N;                     // P[i + j*N], i < a, j < b

    PROF_DECL(tid == 0)
    PROF_COUNT(49, 1);
    for (int idx = tid; idx < m * n; idx += blockDim.x) {
        const int r = idx % m, col = idx / m;
        if (col < c) {
            XU[r + (size_t)col * m] = zmul(ar.beta, CU[r + (size_t)col * ar.ldc]);
            XV[r + (size_t)col * m] = CV[r + (size_t)col * ar.ldc];
        } else {
            XU[r + (size_t)col * m] = zmul(ar.alpha, AU[r + (size_t)(col - c) * ar.lda]);
        }
    }
    PROF_ADD(46);
    // P = AV^T BV (plain transpose) on the FP64 tensor pipe: four real DMMAs per complex 8 x 8 x 4 product.  A warp owns
    // an 8-row strip of P and walks over its column tiles two at a time (the A fragment is reused twice); four
    // contraction steps (16 rows) of operand loads are in flight together -- the first version, one 4 x 4 tile per warp
    // with the lanes along the contraction index and no loads in flight across iterations, spent its time waiting for L2.
    {
        const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
        const int g = lane >> 2, tg = lane & 3;
        const int ta = (a + 7) / 8, tb = (b + 7) / 8;
        const zc zero = zmake(0.0, 0.0);
        for (int ti = warp; ti < ta; ti += nwarps) {
            const int i = 8 * ti + g;
            const zc *ap = AV + (size_t)(i < a ? i : 0) * ar.lda + tg;
            for (int tj0 = 0; tj0 < tb; tj0 += 2) {
                double pr[2][2], pi[2][2];
                const zc *bp[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    pr[q][0] = pr[q][1] = pi[q][0] = pi[q][1] = 0.0;
                    const int j = 8 * (tj0 + q) + g;
                    bp[q] = BV + (size_t)(j < b ? j : 0) * ar.ldb + tg;
                }
                for (int k0 = 0; k0 < m; k0 += 16) {
                    zc av[4], bv[2][4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = k0 + 4 * u;
                        const bool kon = k + tg < m;
                        av[u] = (kon && i < a) ? ap[k] : zero;
#pragma unroll
                        for (int q = 0; q < 2; ++q) bv[q][u] = (kon && 8 * (tj0 + q) + g < b) ? bp[q][k] : zero;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (k0 + 4 * u >= m) break;
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            if (tj0 + q >= tb) break;      // a real exit: predicated-off DMMAs still occupy the FP64 pipe
                            dmma_884(pr[q][0], pr[q][1], av[u].x, bv[q][u].x);
                            dmma_884(pr[q][0], pr[q][1], -av[u].y, bv[q][u].y);
                            dmma_884(pi[q][0], pi[q][1], av[u].x, bv[q][u].y);
                            dmma_884(pi[q][0], pi[q][1], av[u].y, bv[q][u].x);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = 8 * (tj0 + q) + 2 * tg + e;
                        if (i < a && j < b) P[i + (size_t)j * N] = zmake(pr[q][e], pi[q][e]);
                    }
                }
            }
        }
    }
    __syncthreads();
    PROF_ADD(47);
    // G = BU P^T -> X_v[:, c : c+a] on the FP64 tensor pipe (zblock_gemm below; the first version, one thread per row with
    // scalar complex FMAs, was 60 % of this kernel)
    {
        const zc zero = zmake(0.0, 0.0);
        zblock_gemm(m, a, b,
                    [&](int r, int j) { return (r < m && j < b) ? BU[r + (size_t)j * ar.ldb] : zero; },
                    [&](int j, int i) { return (i < a && j < b) ? P[i + (size_t)j * N] : zero; },
                    [&](int r, int i, zc v) { if (r < m && i < a) XV[r + (size_t)(c + i) * m] = v; },
                    [&](int, int, int &, int &) {},
                    tid >> 5, blockDim.x >> 5, tid & 31);
    }
    PROF_ADD(48);
}

// ---- 2. Householder QR (zgeqr2) --------------------------------------------------------------------------
// grid = 2 * tiles, 512 threads; dynamic shared memory: m complex (current reflector) + 40 doubles.
__global__ void __launch_bounds__(512) ztlr_qr_kernel(ZArgs ar, ZWorkspace ws) {
    extern __shared__ __align__(16) double zsm[];
    const int t = blockIdx.x >> 1, pnl = blockIdx.x & 1;
    if (ws.st[t] != HCORE_B200_OK) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    const int n = ws.ncols[t], m = ar.nb;
    zc *X = zpanel(ws, pnl, t, m);
    zc *tau_g = ws.tau + ((size_t)pnl * ws.chunk + t) * ws.N;
    zc *v = reinterpret_cast<zc *>(zsm);
    double *red = zsm + 2 * (size_t)m;          // [32] partial sums, [32..35] scalars
    for (int j = 0; j < n; ++j) {
        zc *xj = X + (size_t)j * m;
        double s = 0.0;
        for (int r = j + 1 + tid; r < m; r += blockDim.x) { const zc x = xj[r]; v[r] = x; s += zabs2(x); }
        s = warp_sum(s);
        if (lane == 0) red[warp] = s;
        __syncthreads();
        if (tid == 0) {
            double xn2 = 0.0;
            for (int w = 0; w < nwarps; ++w) xn2 += red[w];
            const zc alpha = xj[j];
            zc tau = zmake(0.0, 0.0), scale = zmake(0.0, 0.0);
            double beta = alpha.x;
            if (xn2 != 0.0 || alpha.y != 0.0) {                       // LAPACK zlarfg
                beta = -copysign(sqrt(zabs2(alpha) + xn2), alpha.x);
                tau = zmake((beta - alpha.x) / beta, -alpha.y / beta);
                const zc d = zmake(alpha.x - beta, alpha.y);
                const double dn = zabs2(d);
                scale = zmake(d.x / dn, -d.y / dn);                   // 1 / (alpha - beta)
                xj[j] = zmake(beta, 0.0);
            }
            tau_g[j] = tau;
            red[32] = tau.x; red[33] = tau.y; red[34] = scale.x; red[35] = scale.y;
        }
        __syncthreads();
        const zc tau = zmake(red[32], red[33]), scale = zmake(red[34], red[35]);
        for (int r = j + 1 + tid; r < m; r += blockDim.x) { const zc x = zmul(v[r], scale); v[r] = x; xj[r] = x; }
        __syncthreads();
        if (tau.x != 0.0 || tau.y != 0.0) {
            const zc ctau = zconj(tau);                                // H^H = I - conj(tau) v v^H
            for (int c = j + 1 + warp; c < n; c += nwarps) {
                zc *xc = X + (size_t)c * m;
                zc dot = zmake(0.0, 0.0);
                for (int r = j + 1 + lane; r < m; r += 32) dot = zcfma(v[r], xc[r], dot);
                dot = zwarp_sum(dot);
                dot = zadd(dot, xc[j]);
                const zc tw = zmul(ctau, dot);
                if (lane == 0) xc[j] = zsub(xc[j], tw);
                const zc ntw = zmake(-tw.x, -tw.y);
                for (int r = j + 1 + lane; r < m; r += 32) xc[r] = zfma(ntw, v[r], xc[r]);
            }
        }
        __syncthreads();
    }
}

// ---- 2b. blocked Householder QR (nb <= 1024) ----------------------------------------------------------------
// Same reflectors as ztlr_qr_kernel (zgeqr2 order of operations per column inside a block), but the panel is streamed
// from HBM once per 8-column block instead of once per column: the 8 block columns are factored in shared memory, their
// compact-WY factor T is formed (zlarft), and the trailing columns are updated EIGHT AT A TIME on the FP64 tensor pipe,
//   X <- (I - V T^H V^H) X:   W = V^H X,   X <- X - V (T^H W),
// with the scheme of ztlr_apply_wy_kernel below: the 8 trailing columns live transposed in DMMA accumulator fragments
// (16 warps x 8 row blocks), V comes from shared memory (leading dimension m + 1: both fragment layouts conflict-free),
// the per-warp partial W tiles are summed and multiplied by T^H by 64 threads.  The first blocked version updated one
// trailing column per warp with scalar complex FMAs (22 ms per 592 tiles); the unblocked kernel moved 268 MB per panel
// through HBM (66 ms per 592 tiles, HBM-bound), this one moves 33 MB.
constexpr int kZB = 8;          // block width
constexpr int kZRows = 32;      // nb <= 32 * kZRows = 1024: 16 warps x 8 row blocks x 8 rows
constexpr int kZQrRb = 8;       // row blocks per warp

__global__ void __launch_bounds__(512, 1) ztlr_qr_blocked_kernel(ZArgs ar, ZWorkspace ws) {
    extern __shared__ __align__(16) double zsm[];
    const int t = blockIdx.x >> 1, pnl = blockIdx.x & 1;
    if (ws.st[t] != HCORE_B200_OK) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, tg = lane & 3;
    const int n = ws.ncols[t], m = ar.nb, ldp = m + 1;
    zc *X = zpanel(ws, pnl, t, m);
    zc *tau_g = ws.tau + ((size_t)pnl * ws.chunk + t) * ws.N;
    zc *tf_g = ws.tfac + ((size_t)pnl * ws.chunk + t) * ws.N * kZB;
    zc *P = reinterpret_cast<zc *>(zsm);                       // [kZB][ldp] block columns (row index absolute)
    double *red = zsm + 2 * (size_t)kZB * ldp;                 // [16] partial sums, [16..19] tau, scale
    zc *taus = reinterpret_cast<zc *>(red + 32);               // [kZB]
    zc *Sg = taus + kZB;                                       // [kZB][kZB] v_i^H v_k, i < k
    zc *Tsm = Sg + kZB * kZB;                                  // T(i, k) at i + 8 k
    zc *ntw = Tsm + kZB * kZB;                                 // -(T^H W)^T [c][i]
    zc *part = ntw + kZB * kZB;                                // [16][64] per-warp partial W^T tiles
    const zc zero = zmake(0.0, 0.0);
    const bool wide = (m & 1) == 0;                            // columns 32-byte aligned
    // The panels of a chunk do not fit the L2 (2 MB each, one per SM in flight): the eight trailing columns of the NEXT update
    // step are pulled into the L2 while the current ones are in the tensor pipe (rows r0.., 128-byte lines dealt to the threads)
    auto prefetch_cols = [&](int c0, int r0) {
        const int lpc = ((m - r0) * (int)sizeof(zc) + 127) >> 7;          // lines per column
        for (int id = tid; id < 8 * lpc; id += blockDim.x) {
            const int c = c0 + id / lpc;
            if (c < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(X + (size_t)c * m + r0) + ((size_t)(id % lpc) << 7)));
        }
    };
    PROF_DECL(tid == 0)
    PROF_COUNT(55, 1);
    for (int j0 = 0; j0 < n; j0 += kZB) {
        const int nbk = (n - j0 < kZB) ? (n - j0) : kZB, bk = j0 >> 3;
        {                                                      // all sixteen loads of a thread in flight, then the stores
            zc tmp[kZB][2];                                    // (a loop over the columns waited for every load: 20 k cycles per block)
#pragma unroll
            for (int c = 0; c < kZB; ++c)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int r = j0 + tid + 512 * u;
                    tmp[c][u] = (c < nbk && r < m) ? X[(size_t)(j0 + c) * m + r] : zero;
                }
#pragma unroll
            for (int c = 0; c < kZB; ++c)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int r = j0 + tid + 512 * u;
                    if (c < nbk && r < m) P[(size_t)c * ldp + r] = tmp[c][u];
                }
        }
        if (j0 + kZB < n) prefetch_cols(j0 + kZB, j0);
        __syncthreads();
        PROF_ADD(50);
        // ---- the block's columns, in shared memory -------------------------------------------------
        for (int jj = 0; jj < nbk; ++jj) {
            const int j = j0 + jj;
            zc *pj = P + (size_t)jj * ldp;
            double s = 0.0;
            for (int r = j + 1 + tid; r < m; r += blockDim.x) s += zabs2(pj[r]);
            s = warp_sum(s);
            if (lane == 0) red[warp] = s;
            __syncthreads();
            if (tid == 0) {
                double xn2 = 0.0;
                for (int w = 0; w < nwarps; ++w) xn2 += red[w];
                const zc alpha = pj[j];
                zc tau = zmake(0.0, 0.0), scale = zmake(0.0, 0.0);
                if (xn2 != 0.0 || alpha.y != 0.0) {                       // LAPACK zlarfg
                    const double beta = -copysign(sqrt(zabs2(alpha) + xn2), alpha.x);
                    tau = zmake((beta - alpha.x) / beta, -alpha.y / beta);
                    const zc d = zmake(alpha.x - beta, alpha.y);
                    const double dn = zabs2(d);
                    scale = zmake(d.x / dn, -d.y / dn);                   // 1 / (alpha - beta)
                    pj[j] = zmake(beta, 0.0);
                }
                tau_g[j] = tau;
                taus[jj] = tau;
                red[16] = tau.x; red[17] = tau.y; red[18] = scale.x; red[19] = scale.y;
            }
            __syncthreads();
            const zc tau = zmake(red[16], red[17]), scale = zmake(red[18], red[19]);
            for (int r = j + 1 + tid; r < m; r += blockDim.x) pj[r] = zmul(pj[r], scale);
            __syncthreads();
            if (warp >= kZB && warp - kZB < jj) {                      // warps without a column: S(i, jj) = v_i^H v_jj, i < jj
                const int i = warp - kZB;
                const zc *pi = P + (size_t)i * ldp;
                zc d0 = zmake(0.0, 0.0), d1 = zmake(0.0, 0.0);
                int r = j + 1 + lane;
                for (; r + 32 < m; r += 64) { d0 = zcfma(pi[r], pj[r], d0); d1 = zcfma(pi[r + 32], pj[r + 32], d1); }
                if (r < m) d0 = zcfma(pi[r], pj[r], d0);
                const zc dot = zadd(zwarp_sum(zadd(d0, d1)), zconj(pi[j]));   // unit diagonal of v_jj
                if (lane == 0) Sg[i + kZB * jj] = dot;
            }
            if (tau.x != 0.0 || tau.y != 0.0) {
                const zc ctau = zconj(tau);                                // H^H = I - conj(tau) v v^H
                for (int c = jj + 1 + warp; c < nbk; c += nwarps) {
                    zc *pc = P + (size_t)c * ldp;
                    zc d0 = zmake(0.0, 0.0), d1 = zmake(0.0, 0.0);
                    int r = j + 1 + lane;
                    for (; r + 32 < m; r += 64) { d0 = zcfma(pj[r], pc[r], d0); d1 = zcfma(pj[r + 32], pc[r + 32], d1); }
                    if (r < m) d0 = zcfma(pj[r], pc[r], d0);
                    zc dot = zwarp_sum(zadd(d0, d1));
                    dot = zadd(dot, pc[j]);
                    const zc tw = zmul(ctau, dot), ntw1 = zmake(-tw.x, -tw.y);
                    __syncwarp();
                    if (lane == 0) pc[j] = zsub(pc[j], tw);
                    int q = j + 1 + lane;                                  // four rows per lane in flight (pj and pc never alias)
                    for (; q + 96 < m; q += 128) {
                        const zc v0 = pj[q], v1 = pj[q + 32], v2 = pj[q + 64], v3 = pj[q + 96];
                        const zc x0 = pc[q], x1 = pc[q + 32], x2 = pc[q + 64], x3 = pc[q + 96];
                        pc[q] = zfma(ntw1, v0, x0); pc[q + 32] = zfma(ntw1, v1, x1);
                        pc[q + 64] = zfma(ntw1, v2, x2); pc[q + 96] = zfma(ntw1, v3, x3);
                    }
                    for (; q < m; q += 32) pc[q] = zfma(ntw1, pj[q], pc[q]);
                }
            }
            __syncthreads();
        }
        PROF_ADD(51);
        // ---- compact-WY factor of the block (zlarft, forward / columnwise): H_j0 ... H_{j0+7} = I - V T V^H with
        //      T(k,k) = tau_k, T(0:k, k) = -tau_k T(0:k, 0:k) (V(:, 0:k)^H v_k); also consumed by ztlr_apply_wy_kernel ----
        {
            if (tid < kZB) {                                   // thread i builds row i of T (rows are independent)
                const int i = tid;
                zc Ti[kZB];
#pragma unroll
                for (int k = 0; k < kZB; ++k) {
                    zc v = zmake(0.0, 0.0);
                    if (i < nbk && k < nbk) {
                        if (k == i) v = taus[i];
                        else if (k > i) {
                            zc acc = zmake(0.0, 0.0);
#pragma unroll
                            for (int l = 0; l < kZB; ++l)
                                if (l >= i && l < k) acc = zfma(Ti[l], Sg[l + kZB * k], acc);
                            v = zmul(zmake(-taus[k].x, -taus[k].y), acc);
                        }
                    }
                    Ti[k] = v;
                    Tsm[i + kZB * k] = v;
                    tf_g[((size_t)j0 << 3) + i + kZB * k] = v;
                }
            }
        }
        PROF_ADD(52);
        // ---- trailing columns, eight at a time (a full block whenever there are any) --------------------------
        // V(r, jj) of this block from shared memory: below the diagonal as stored, unit diagonal, zero above
        auto vsm = [&](int r, int jj, bool diag) {
            zc v = P[(size_t)jj * ldp + (r < m ? r : m - 1)];
            if (r >= m) v = zero;
            else if (diag && r <= j0 + jj) v = zmake(r == j0 + jj ? 1.0 : 0.0, 0.0);
            return v;
        };
        for (int cg0 = j0 + kZB; cg0 < n; cg0 += 8) {
            const int col = cg0 + g;
            zc *xc = X + (size_t)(col < n ? col : n - 1) * m;
            if (cg0 + 8 < n) prefetch_cols(cg0 + 8, j0);
            double zr[kZQrRb][2], zi[kZQrRb][2];               // X^T fragments: column g, rows 8 rb + 2 tg + e
#pragma unroll
            for (int i = 0; i < kZQrRb; ++i) {
                const int rb = warp + 16 * i, r = 8 * rb + 2 * tg;
                zc v0 = zero, v1 = zero;
                if (rb >= bk && col < n) {
                    if (wide) { if (r < m) ldg_z2(xc + r, v0, v1); }          // one 32-byte access per lane instead of two half sectors
                    else { if (r < m) v0 = xc[r]; if (r + 1 < m) v1 = xc[r + 1]; }
                }
                zr[i][0] = v0.x; zi[i][0] = v0.y; zr[i][1] = v1.x; zi[i][1] = v1.y;
            }
            PROF_ADD(56);
            // W^T partial: row c = g, columns i = 2 tg + {0, 1}; one accumulator pair per contraction chunk e, so that four
            // independent DMMAs separate every dependent pair (a warp issues in order: back-to-back dependent DMMAs wait 26 cycles)
            double wr[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, wi[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
            for (int i = 0; i < kZQrRb; ++i) {
                const int rb = warp + 16 * i;
                if (rb >= bk && 8 * rb < m) {                  // x conj(v)
                    const zc v0 = vsm(8 * rb + 2 * tg, g, rb == bk), v1 = vsm(8 * rb + 2 * tg + 1, g, rb == bk);
                    dmma_884(wr[0][0], wr[0][1], zr[i][0], v0.x);
                    dmma_884(wi[0][0], wi[0][1], zi[i][0], v0.x);
                    dmma_884(wr[1][0], wr[1][1], zr[i][1], v1.x);
                    dmma_884(wi[1][0], wi[1][1], zi[i][1], v1.x);
                    dmma_884(wr[0][0], wr[0][1], zi[i][0], v0.y);
                    dmma_884(wi[0][0], wi[0][1], zr[i][0], -v0.y);
                    dmma_884(wr[1][0], wr[1][1], zi[i][1], v1.y);
                    dmma_884(wi[1][0], wi[1][1], zr[i][1], -v1.y);
                }
            }
            PROF_ADD(57);
            part[warp * 64 + g * 8 + 2 * tg] = zmake(wr[0][0] + wr[1][0], wi[0][0] + wi[1][0]);
            part[warp * 64 + g * 8 + 2 * tg + 1] = zmake(wr[0][1] + wr[1][1], wi[0][1] + wi[1][1]);
            __syncthreads();
            if (tid < 64) {                                    // entry (c, k): sum the partials, then row i = k of T^H W
                const int i = tid & 7;
                zc s0 = zero, s1 = zero, s2 = zero, s3 = zero;     // sixteen loads in flight, four short chains
#pragma unroll
                for (int w = 0; w < 16; w += 4) {
                    s0 = zadd(s0, part[w * 64 + tid]); s1 = zadd(s1, part[(w + 1) * 64 + tid]);
                    s2 = zadd(s2, part[(w + 2) * 64 + tid]); s3 = zadd(s3, part[(w + 3) * 64 + tid]);
                }
                const zc s = zadd(zadd(s0, s1), zadd(s2, s3));
                zc acc = zero;
#pragma unroll
                for (int k = 0; k < 8; ++k) {                  // (T^H W)(i, c) = sum_k conj(T(k, i)) W(k, c)
                    const zc wk = zmake(__shfl_sync(0xffffffffu, s.x, (lane & ~7) + k), __shfl_sync(0xffffffffu, s.y, (lane & ~7) + k));
                    acc = zcfma(Tsm[k + 8 * i], wk, acc);
                }
                ntw[tid] = zmake(-acc.x, -acc.y);
            }
            __syncthreads();
            PROF_ADD(58);
            const zc a[2] = {ntw[g * 8 + 2 * tg], ntw[g * 8 + 2 * tg + 1]};   // contraction chunks: reflectors 2 tg + h
#pragma unroll
            for (int i = 0; i < kZQrRb; ++i) {
                const int rb = warp + 16 * i;
                if (rb >= bk && 8 * rb < m) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {              // a v, real and imaginary accumulators alternating
                        const zc v = vsm(8 * rb + g, 2 * tg + h, rb == bk);
                        dmma_884(zr[i][0], zr[i][1], a[h].x, v.x);
                        dmma_884(zi[i][0], zi[i][1], a[h].x, v.y);
                        dmma_884(zr[i][0], zr[i][1], a[h].y, -v.y);
                        dmma_884(zi[i][0], zi[i][1], a[h].y, v.x);
                    }
                }
            }
            PROF_ADD(59);
            if (col < n) {
#pragma unroll
                for (int i = 0; i < kZQrRb; ++i) {
                    const int rb = warp + 16 * i, r = 8 * rb + 2 * tg;
                    if (rb >= bk) {
                        if (wide) { if (r < m) stg_z2(xc + r, zmake(zr[i][0], zi[i][0]), zmake(zr[i][1], zi[i][1])); }
                        else {
                            if (r < m) xc[r] = zmake(zr[i][0], zi[i][0]);
                            if (r + 1 < m) xc[r + 1] = zmake(zr[i][1], zi[i][1]);
                        }
                    }
                }
            }
        }
        PROF_ADD(53);
        // ---- the factored block goes back (reflectors below the diagonal, R on and above) -----------------
        for (int c = 0; c < nbk; ++c)
            for (int r = j0 + tid; r < m; r += blockDim.x) X[(size_t)(j0 + c) * m + r] = P[(size_t)c * ldp + r];
        __syncthreads();
        PROF_ADD(54);
    }
}

// ---- 3. core SVD ------------------------------------------------------------------------------------------
constexpr double kZTolRot = 1e-15, kZTolStop = 1e-7;
constexpr int kZMaxSweeps = 60;

// RPL = rows per lane of a 16-lane column pair: 8 (stacks up to 128, 512 threads) or 16 (stacks up to 256: ranks up to
// 128, 256 threads so that the two columns of a pair -- 64 doubles per lane -- stay in registers)
template <int RPL>
__global__ void __launch_bounds__(RPL == 8 ? 512 : 256) ztlr_svd_kernel(ZArgs ar, ZWorkspace ws) {
    __shared__ double sig[HCORE_B200_MAX_STACK_Z], ssort[HCORE_B200_MAX_STACK_Z], vn[HCORE_B200_MAX_STACK_Z], red[40];
    __shared__ zc vref[HCORE_B200_MAX_STACK_Z];
    __shared__ int posof[HCORE_B200_MAX_STACK_Z], perm[HCORE_B200_MAX_STACK_Z], misc[4];
    const int t = blockIdx.x;
    if (ws.st[t] != HCORE_B200_OK) return;
    const int gt = ar.tile0 + t;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    const int n = ws.ncols[t], m = ar.nb, N = ws.N;
    const zc *Ru = zpanel(ws, 0, t, m), *Rv = zpanel(ws, 1, t, m);   // upper triangles, ld m
    zc *T = ws.T + (size_t)t * N * N;                                 // ld N
    zc *Yu = ws.Y + ((size_t)0 * ws.chunk + t) * N * N, *Yv = ws.Y + ((size_t)1 * ws.chunk + t) * N * N;

    PROF_DECL(tid == 0)
    PROF_COUNT(39, 1);
    // A = T^H with T = R_u R_v^T (plain transpose): T(i,j) = sum_{k >= max(i,j)} Ru(i,k) Rv(j,k), A(j,i) = conj(T(i,j)).
    // The left singular vectors of T are wanted: with A P = Q R (pivoted QR), T = P R^H Q^H, so they are P times the
    // left singular vectors of X = R^H, which one-sided Jacobi on the columns of X delivers without accumulating
    // anything -- the preconditioning of the real path (tlr_kernels.cuh), in complex arithmetic.
    {
        const zc zero = zmake(0.0, 0.0);
        // T(i, j) = sum_{k >= max(i, j)} Ru(i, k) Rv(j, k): A = Ru, B = Rv^T; stored as A^H: T[j + i N] = conj(T(i, j))
        zblock_gemm(n, n, n,
                    [&](int i, int k) { return (i < n && k < n && k >= i) ? Ru[i + (size_t)k * m] : zero; },
                    [&](int k, int j) { return (j < n && k < n && k >= j) ? Rv[j + (size_t)k * m] : zero; },
                    [&](int i, int j, zc v) { if (i < n && j < n) T[j + (size_t)i * N] = zconj(v); },
                    [&](int i0, int j0, int &klo, int &) { klo = (i0 > j0 ? i0 : j0) & ~3; },
                    warp, nwarps, lane);
    }
    if (tid < n) perm[tid] = tid;
    __syncthreads();

    PROF_ADD(30);
    // ---- truncated QR with column pivoting of A (in place, in global memory / L2) -----------------------------------------
    // Exact squared norms of the remaining columns in every step (they fall out of the update), pivot = the largest,
    // stop when the Frobenius norm of the remaining block is below 1 % of the truncation threshold (or, fixed-rank mode,
    // rounding level of the largest column).  zlarfg / H^H as in ztlr_qr_kernel.
    for (int cc = warp; cc < n; cc += nwarps) {
        double s2 = 0.0;
        for (int r = lane; r < n; r += 32) s2 += zabs2(T[r + (size_t)cc * N]);
        s2 = warp_sum(s2);
        if (lane == 0) vn[cc] = s2;
    }
    __syncthreads();
    int kq = n;
    double tol2 = 0.0;
    for (int j = 0; j < n; ++j) {
        // pivot and stop test (every warp, redundantly)
        double best = -1.0, fro2 = 0.0; int bi = 0x7fffffff;
        for (int cc = j + lane; cc < n; cc += 32) {
            const double v = vn[cc];
            fro2 += v;
            if (v > best) { best = v; bi = cc; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            fro2 += __shfl_xor_sync(0xffffffffu, fro2, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (j == 0) tol2 = (ar.rk > 0) ? 1e-30 * best : 1e-4 * ar.acc * ar.acc;
        if (fro2 < tol2) { kq = j; break; }
        const int piv = bi;
        __syncthreads();                                   // everybody has read vn[] before it is swapped / rewritten
        zc *xj = T + (size_t)j * N, *xp = T + (size_t)piv * N;
        if (piv != j) {
            for (int r = tid; r < n; r += blockDim.x) { const zc a = xj[r]; xj[r] = xp[r]; xp[r] = a; }
            if (tid == 0) { const int pi = perm[j]; perm[j] = perm[piv]; perm[piv] = pi; vn[piv] = vn[j]; }
        }
        __syncthreads();
        // reflector of column j (rows j .. n-1)
        double sloc = 0.0;
        for (int r = j + 1 + tid; r < n; r += blockDim.x) { const zc x = xj[r]; vref[r] = x; sloc += zabs2(x); }
        sloc = warp_sum(sloc);
        if (lane == 0) red[warp] = sloc;
        __syncthreads();
        if (tid == 0) {
            double xn2 = 0.0;
            for (int w = 0; w < nwarps; ++w) xn2 += red[w];
            const zc alpha = xj[j];
            zc tau = zmake(0.0, 0.0), scale = zmake(0.0, 0.0);
            if (xn2 != 0.0 || alpha.y != 0.0) {                           // LAPACK zlarfg
                const double beta = -copysign(sqrt(zabs2(alpha) + xn2), alpha.x);
                tau = zmake((beta - alpha.x) / beta, -alpha.y / beta);
                const zc d = zmake(alpha.x - beta, alpha.y);
                const double dn = zabs2(d);
                scale = zmake(d.x / dn, -d.y / dn);
                xj[j] = zmake(beta, 0.0);
            }
            red[32] = tau.x; red[33] = tau.y; red[34] = scale.x; red[35] = scale.y;
        }
        __syncthreads();
        const zc tau = zmake(red[32], red[33]), scale = zmake(red[34], red[35]);
        for (int r = j + 1 + tid; r < n; r += blockDim.x) vref[r] = zmul(vref[r], scale);
        __syncthreads();
        // H^H = I - conj(tau) v v^H on the remaining columns (a warp per column), new squared norms of rows > j
        const zc ctau = zconj(tau);
        for (int c = j + 1 + warp; c < n; c += nwarps) {
            zc *xc = T + (size_t)c * N;
            zc dot = zmake(0.0, 0.0);
            for (int r = j + 1 + lane; r < n; r += 32) dot = zcfma(vref[r], xc[r], dot);
            dot = zwarp_sum(dot);
            dot = zadd(dot, xc[j]);
            const zc tw = zmul(ctau, dot), ntw = zmake(-tw.x, -tw.y);
            __syncwarp();
            if (lane == 0) xc[j] = zsub(xc[j], tw);
            double s2 = 0.0;
            for (int r = j + 1 + lane; r < n; r += 32) { const zc x = zfma(ntw, vref[r], xc[r]); xc[r] = x; s2 += zabs2(x); }
            s2 = warp_sum(s2);
            if (lane == 0) vn[c] = s2;
        }
        __syncthreads();
    }
    PROF_ADD(31);
    // ---- X = R^H restricted to the kq factored rows (the reflectors below the diagonal are overwritten) -------------------
    for (int idx = tid; idx < n * n; idx += blockDim.x) {
        const int i = idx % n, j = idx / n;                // entry (i, j) of R, i <= j
        if (i <= j) {
            const zc x = (i < kq) ? zconj(T[i + (size_t)j * N]) : zmake(0.0, 0.0);
            if (i < j) T[i + (size_t)j * N] = zmake(0.0, 0.0);
            T[j + (size_t)i * N] = x;
        }
    }
    __syncthreads();

    PROF_ADD(32);
    // one-sided Jacobi on the first kq columns of X; a half-warp per pair, lane l16 owns rows l16, l16+16, ...
    const int n2 = kq + (kq & 1);
    int sweeps = 0;
    bool converged = (n2 < 2);
    for (int sweep = 0; sweep < kZMaxSweeps && !converged; ++sweep) {
        int big = 0;
        for (int step = 0; step < n2 - 1; ++step) {
            // 16 lanes per column pair (lane l16 owns rows l16, l16 + 16, ...): two pairs per warp, the
            // pairs of a step in rounds of the CTA's half-warps
            const int l16 = lane & 15, hw = (warp << 1) | (lane >> 4), nhw = nwarps << 1;
            for (int k = hw; k < n2 / 2; k += nhw) {
                int p, q;
                rr_pair(n2, step, k, p, q);
                if (q >= kq) continue;
                zc *cp = T + (size_t)p * N, *cq = T + (size_t)q * N;
                zc xp[RPL], xq[RPL];
                double app = 0.0, aqq = 0.0;
                zc apq = zmake(0.0, 0.0);
#pragma unroll
                for (int e = 0; e < RPL; ++e) {
                    const int r = l16 + 16 * e;
                    xp[e] = (r < n) ? cp[r] : zmake(0.0, 0.0);
                    xq[e] = (r < n) ? cq[r] : zmake(0.0, 0.0);
                }
#pragma unroll
                for (int e = 0; e < RPL; ++e) {
                    app += zabs2(xp[e]); aqq += zabs2(xq[e]);
                    apq = zcfma(xp[e], xq[e], apq);                  // p^H q
                }
                const unsigned hmask = (lane & 16) ? 0xffff0000u : 0x0000ffffu;   // the other half-warp may be idle
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {                    // reductions stay inside the half-warp
                    app += __shfl_xor_sync(hmask, app, o);
                    aqq += __shfl_xor_sync(hmask, aqq, o);
                    apq.x += __shfl_xor_sync(hmask, apq.x, o);
                    apq.y += __shfl_xor_sync(hmask, apq.y, o);
                }
                const double den2 = app * aqq, g2 = zabs2(apq);
                if (den2 > 0.0 && g2 > (kZTolRot * kZTolRot) * den2) {
                    if (g2 > (kZTolStop * kZTolStop) * den2) big = 1;
                    // three reciprocal square roots instead of two square roots and four divisions (the scalar chain of a
                    // rotation is on the critical path of every round): |apq| = g2 / sqrt(g2); with d = aqq - app and
                    // hyp = sqrt(d^2 + 4 g^2): cos^2 = (1 + |d| / hyp) / 2, sin = sign(d) g / (hyp cos)
                    const double rg = rsqrt(g2);
                    const zc ph = zmake(apq.x * rg, -apq.y * rg);   // conj(theta): q <- q conj(theta) makes p^H q real
                    const double dd = aqq - app, h2 = 2.0 * (g2 * rg);
                    const double rh = rsqrt(fma(dd, dd, h2 * h2));
                    const double u = 0.5 * fma(fabs(dd), rh, 1.0);
                    const double ru = rsqrt(u);
                    const double cs = u * ru, sn = (dd >= 0.0 ? h2 : -h2) * rh * (0.5 * ru);
#pragma unroll
                    for (int e = 0; e < RPL; ++e) {
                        const int r = l16 + 16 * e;
                        if (r < n) {
                            const zc qt = zmul(xq[e], ph);
                            cp[r] = zmake(cs * xp[e].x - sn * qt.x, cs * xp[e].y - sn * qt.y);
                            cq[r] = zmake(sn * xp[e].x + cs * qt.x, sn * xp[e].y + cs * qt.y);
                        }
                    }
                }
            }
            __syncthreads();
        }
        ++sweeps;
        converged = (__syncthreads_or(big) == 0);
    }

    PROF_ADD(33);
    PROF_COUNT(38, sweeps);
    PROF_COUNT(37, kq);
    // singular values, sort (descending, stable), truncation
    for (int cc = warp; cc < n; cc += nwarps) {
        double s2 = 0.0;
        for (int r = lane; r < n; r += 32) s2 += zabs2(T[r + (size_t)cc * N]);
        s2 = warp_sum(s2);
        if (lane == 0) sig[cc] = sqrt(s2);
    }
    if (tid == 0) misc[1] = n;
    __syncthreads();
    if (tid < n) {
        const double mine = sig[tid];
        int pos = 0;
        for (int jj = 0; jj < n; ++jj) { const double o = sig[jj]; pos += (o > mine || (o == mine && jj < tid)) ? 1 : 0; }
        posof[tid] = pos;
        ssort[pos] = mine;
    }
    __syncthreads();
    // reference src/compute/hcore_zgemm.c:376-395: fixed rank iff rk > 0, else the first index
    // i >= 1 with sigma_i < acc (absolute)
    if (ar.rk > 0) { if (tid == 0) misc[1] = ar.rk < n ? ar.rk : n; }
    else if (tid >= 1 && tid < n && ssort[tid] < ar.acc) atomicMin(&misc[1], tid);
    __syncthreads();
    const int rnew = misc[1];
    if (tid < n) ws.sigma[(size_t)t * N + tid] = ssort[tid];
    int st = converged ? HCORE_B200_OK : HCORE_B200_ERR_NO_CONVERGE;
    if (rnew > ar.maxrk) st = HCORE_B200_ERR_NEWRANK;
    if (tid == 0) {
        ws.newrank[t] = rnew; ws.sweeps[t] = sweeps; ws.st[t] = st;
        if (ar.status) ar.status[gt] = st;
        if (st == HCORE_B200_OK) ar.Crk[gt] = rnew;
    }
    if (st != HCORE_B200_OK) return;
    // U_r: normalised columns in sorted order
    for (int idx = tid; idx < n * n; idx += blockDim.x) {
        const int i = idx % n, cc = idx / n, pos = posof[cc];
        if (pos < rnew) {
            const double sg = sig[cc];
            Yu[perm[i] + (size_t)pos * N] = sg > 0.0 ? zscale(T[i + (size_t)cc * N], 1.0 / sg) : zmake(0.0, 0.0);
        }
    }
    __syncthreads();
    PROF_ADD(34);
    // tmp = R_u^T conj(U_r)  (n x r) into T;  W = R_v tmp
    {
        const zc zero = zmake(0.0, 0.0);
        // tmp(k, c) = sum_{i <= k} Ru(i, k) conj(Yu(i, c)): A = Ru^T (rows k, contraction i), B = conj(Yu)
        zblock_gemm(n, rnew, n,
                    [&](int k, int i) { return (k < n && i <= k) ? Ru[i + (size_t)k * m] : zero; },
                    [&](int i, int c) { return (i < n && c < rnew) ? zconj(Yu[i + (size_t)c * N]) : zero; },
                    [&](int k, int c, zc v) { if (k < n && c < rnew) T[k + (size_t)c * N] = v; },
                    [&](int k0, int, int &, int &khi) { const int lim = (k0 + 8 + 3) & ~3; khi = lim < khi ? lim : khi; },
                    warp, nwarps, lane);
        __syncthreads();
        // W(j, c) = sum_{k >= j} Rv(j, k) tmp(k, c)
        zblock_gemm(n, rnew, n,
                    [&](int j, int k) { return (j < n && k < n && k >= j) ? Rv[j + (size_t)k * m] : zero; },
                    [&](int k, int c) { return (k < n && c < rnew) ? T[k + (size_t)c * N] : zero; },
                    [&](int j, int c, zc v) { if (j < n && c < rnew) Yv[j + (size_t)c * N] = v; },
                    [&](int j0, int, int &klo, int &) { klo = j0 & ~3; },
                    warp, nwarps, lane);
    }
    PROF_ADD(35);
}

// ---- 4. reflector application --------------------------------------------------------------------------------
// grid = (ceil(N / W), 2, tiles), W warps: every warp owns one output column, kept in shared memory
// (m complex per warp), and applies H_{n-1}, ..., H_0 to [y; 0].  The reflector vectors are staged ONCE
// per CTA through a double-buffered cp.async pipeline (reflector j-1 travels while j is applied) and
// shared by the W columns; the first version had every warp stream every reflector from L2 with
// dependent loads (86 ms per 592 tiles at nb = 1024, k = 64).
// Reflectors are applied TWO AT A TIME (the kernel is bound by shared-memory bandwidth: a column and a reflector are
// 16 KB each): with d = v_j^H z, e = v_{j-1}^H z and g = v_{j-1}^H v_j (all three accumulated in the same pass, by every
// warp),  H_{j-1} H_j z = z - t_j v_j - t_{j-1} v_{j-1},  t_j = tau_j d,  t_{j-1} = tau_{j-1} (e - t_j g):
// two passes over z and the two reflectors per PAIR instead of two passes per reflector.
__global__ void __launch_bounds__(256) ztlr_apply_kernel(ZArgs ar, ZWorkspace ws) {
    extern __shared__ __align__(16) double zsm[];
    const int t = blockIdx.z, pnl = blockIdx.y;
    if (ws.st[t] != HCORE_B200_OK) return;
    const int W = blockDim.x >> 5;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int col = blockIdx.x * W + warp;
    const int rnew = ws.newrank[t];
    if ((int)blockIdx.x * W >= rnew) return;                 // whole CTA idle
    const bool on = col < rnew;
    const int gt = ar.tile0 + t;
    const int n = ws.ncols[t], m = ar.nb, N = ws.N;
    const zc *X = zpanel(ws, pnl, t, m);
    const zc *tau_g = ws.tau + ((size_t)pnl * ws.chunk + t) * N;
    const zc *y = ws.Y + ((size_t)pnl * ws.chunk + t) * N * N + (size_t)col * N;
    zc *z = reinterpret_cast<zc *>(zsm) + (size_t)warp * m;
    zc *vbuf = reinterpret_cast<zc *>(zsm) + (size_t)W * m;   // [2 pairs][2 reflectors][m]
    zc *out = (ar.ptrs ? const_cast<zc *>(ar.ptrs[(4 + pnl) * ar.ptr_stride + gt]) : (pnl == 0 ? ar.CU : ar.CV) + (size_t)gt * ar.strideC) +
              (size_t)col * ar.ldc;
    if (on)
        for (int r = lane; r < m; r += 32) z[r] = (r < n) ? y[r] : zmake(0.0, 0.0);
    // pair p = reflectors (jh, jh - 1), jh = n - 1 - 2p; the last pair of an odd n has only its upper reflector
    const int npairs = (n + 1) >> 1;
    auto stage = [&](int p, int b) {                          // rows below the diagonal of both reflectors -> vbuf[b]
        const int jh = n - 1 - 2 * p;
        zc *dst = vbuf + (size_t)b * 2 * m;
        for (int r = jh + 1 + tid; r < m; r += blockDim.x) cp_async16(dst + r, X + (size_t)jh * m + r);
        if (jh >= 1)
            for (int r = jh + tid; r < m; r += blockDim.x) cp_async16(dst + m + r, X + (size_t)(jh - 1) * m + r);
        cp_async_commit();
    };
    stage(0, 0);
    for (int p = 0; p < npairs; ++p) {
        const int b = p & 1, jh = n - 1 - 2 * p, jl = jh - 1;
        if (p + 1 < npairs) {
            stage(p + 1, b ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();                                      // the pair is in vbuf[b] for every warp
        const zc *vh = vbuf + (size_t)b * 2 * m, *vl = vh + m;    // vh: rows > jh valid, vl: rows >= jh valid (jl >= 0)
        const bool two = jl >= 0;
        // g = v_jl^H v_jh = conj(v_jl[jh]) + sum_{r > jh} conj(v_jl[r]) v_jh[r]: every warp, redundantly (no extra barrier)
        zc g = zmake(0.0, 0.0);
        const zc tauh = tau_g[jh], taul = two ? tau_g[jl] : zmake(0.0, 0.0);
        if (on) {
            zc d0 = zmake(0.0, 0.0), d1 = zmake(0.0, 0.0), e0 = zmake(0.0, 0.0), e1 = zmake(0.0, 0.0);
            zc g0 = zmake(0.0, 0.0), g1 = zmake(0.0, 0.0);
            int r = jh + 1 + lane;
            if (two) {
                for (; r + 32 < m; r += 64) {
                    const zc z0 = z[r], z1 = z[r + 32], h0 = vh[r], h1 = vh[r + 32], l0 = vl[r], l1 = vl[r + 32];
                    d0 = zcfma(h0, z0, d0); d1 = zcfma(h1, z1, d1);
                    e0 = zcfma(l0, z0, e0); e1 = zcfma(l1, z1, e1);
                    g0 = zcfma(l0, h0, g0); g1 = zcfma(l1, h1, g1);
                }
                if (r < m) { const zc z0 = z[r], h0 = vh[r], l0 = vl[r]; d0 = zcfma(h0, z0, d0); e0 = zcfma(l0, z0, e0); g0 = zcfma(l0, h0, g0); }
            } else {
                for (; r + 32 < m; r += 64) { d0 = zcfma(vh[r], z[r], d0); d1 = zcfma(vh[r + 32], z[r + 32], d1); }
                if (r < m) d0 = zcfma(vh[r], z[r], d0);
            }
            zc d = zwarp_sum(zadd(d0, d1)), e = zwarp_sum(zadd(e0, e1));
            g = zwarp_sum(zadd(g0, g1));
            const zc zjh = z[jh];
            d = zadd(d, zjh);
            zc th = zmul(tauh, d), tl = zmake(0.0, 0.0);
            if (two) {
                const zc cl = zconj(vl[jh]);
                g = zadd(g, cl);
                e = zadd(e, zadd(z[jl], zmul(cl, zjh)));
                tl = zmul(taul, zsub(e, zmul(th, g)));
            }
            const zc nth = zmake(-th.x, -th.y), ntl = zmake(-tl.x, -tl.y);
            __syncwarp();
            if (lane == 0) {
                z[jh] = two ? zsub(zsub(zjh, th), zmul(tl, vl[jh])) : zsub(zjh, th);
                if (two) z[jl] = zsub(z[jl], tl);
            }
            if (two) {
                for (int q = jh + 1 + lane; q < m; q += 32) z[q] = zfma(ntl, vl[q], zfma(nth, vh[q], z[q]));
            } else {
                for (int q = jh + 1 + lane; q < m; q += 32) z[q] = zfma(nth, vh[q], z[q]);
            }
        }
        __syncthreads();                                      // vbuf[b] may be overwritten by the next-but-one stage
    }
    if (on)
        for (int r = lane; r < m; r += 32) out[r] = z[r];
}

// ---- 4b. blocked reflector application on the FP64 tensor pipe (panels factored by ztlr_qr_blocked_kernel) --------
// grid = (ceil(N / 8), 2, tiles), 512 threads: a CTA owns 8 output columns, Z = [Y; 0] (m x 8), held TRANSPOSED in
// DMMA accumulator fragments (warp w owns the row blocks w, w + 16, ..., 8 rows each: 32 doubles per lane; 16 columns
// per CTA would be the whole register file).  The reflectors are applied eight at a time through their compact-WY
// factors (T from the QR kernel), last block first:   W = V^H Z,   Z <- Z - V (T W).
//   step 1  W^T(c, i) = sum_r Z(r, c) conj(V(r, i)): the accumulator fragment of Z^T (lane (g, t): column g, rows 2t,
//           2t + 1 of the block) IS an A fragment when the contraction chunks are taken as the even / the odd rows, so
//           Z never leaves its registers; the B fragment is conj(V(8 rb + 2t + e, j0 + g)), loaded from L2.
//           Per-warp partial tiles are summed through shared memory by 64 threads, which also multiply by T
//           (eight-lane shuffles: the eight entries of a column of W sit in eight consecutive lanes).
//   step 2  Z^T(c, r) -= sum_i (T W)(i, c) V(r, i): A = -(T W)^T from shared memory, B = V(8 rb + g, j0 + t + 4h)
//           (the same 128 KB of V again, in the other fragment layout: L1 hits where the block survived there).
// 8 real DMMAs per row block and step against two 16-byte loads; the one-column-per-warp kernel above (scalar complex
// FMAs on operands in shared memory) was bound by shared-memory bandwidth: 42 ms per 592 tiles.
constexpr int kZApRb = 8;          // row blocks per warp: 16 warps x 8 blocks x 8 rows = 1024 rows

__global__ void __launch_bounds__(512, 1) ztlr_apply_wy_kernel(ZArgs ar, ZWorkspace ws) {
    __shared__ __align__(16) zc part[16][64];          // per-warp partial W^T tiles [c][i]
    __shared__ __align__(16) zc ntw[64];               // -(T W)^T [c][i]
    __shared__ __align__(16) zc Tb[64];                // T(i, k) at i + 8 k
    const int t = blockIdx.z, pnl = blockIdx.y;
    if (ws.st[t] != HCORE_B200_OK) return;
    const int rnew = ws.newrank[t];
    const int c0 = blockIdx.x * 8;
    if (c0 >= rnew) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
    const int gt = ar.tile0 + t;
    const int n = ws.ncols[t], m = ar.nb, N = ws.N;
    const zc *X = zpanel(ws, pnl, t, m);
    const zc *tf_g = ws.tfac + ((size_t)pnl * ws.chunk + t) * N * kZB;
    const zc *Y = ws.Y + ((size_t)pnl * ws.chunk + t) * N * N;
    zc *out = ar.ptrs ? const_cast<zc *>(ar.ptrs[(4 + pnl) * ar.ptr_stride + gt]) : (pnl == 0 ? ar.CU : ar.CV) + (size_t)gt * ar.strideC;
    const zc zero = zmake(0.0, 0.0);
    const int col = c0 + g;

    double zr[kZApRb][2], zi[kZApRb][2];                // Z^T fragments [row block][e]: column g, row 2 tg + e
#pragma unroll
    for (int i = 0; i < kZApRb; ++i)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int r = 8 * (warp + 16 * i) + 2 * tg + e;
            const zc v = (r < n && col < rnew) ? Y[r + (size_t)col * N] : zero;
            zr[i][e] = v.x; zi[i][e] = v.y;
        }
    // V(r, j): reflector j below its diagonal, unit diagonal, zero above (R lives there) and outside the panel.  The load is
    // a volatile asm so that it keeps its place in the instruction stream: ONE row block ahead of the DMMAs that consume it.
    auto vload = [&](int r, int j, bool diag) {
        const int rc = r < m ? r : m - 1, jc = j < n ? j : n - 1;
        zc v;
        asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(X + rc + (size_t)jc * m));
        if (r >= m || j >= n) v = zero;
        else if (diag && r <= j) v = zmake(r == j ? 1.0 : 0.0, 0.0);
        return v;
    };

    const int nblk = (n + 7) >> 3;
    for (int bk = nblk - 1; bk >= 0; --bk) {
        const int j0 = 8 * bk;
        if (tid < 64) Tb[tid] = tf_g[((size_t)j0 << 3) + tid];
        // ---- step 1: partial W^T over this warp's row blocks ----
        // row c = g, columns i = 2 tg + {0, 1}; one accumulator pair per contraction chunk (four independent DMMAs between dependent ones)
        double wr[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, wi[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        auto ld1 = [&](int rb, zc *v) {
            const int r = 8 * rb + 2 * tg;
            v[0] = vload(r, j0 + g, rb == bk); v[1] = vload(r + 1, j0 + g, rb == bk);
        };
        zc vn[2] = {zero, zero};
        if (warp >= bk && 8 * warp < m) ld1(warp, vn);
#pragma unroll
        for (int i = 0; i < kZApRb; ++i) {
            const int rb = warp + 16 * i;
            const zc v[2] = {vn[0], vn[1]};
            if (i + 1 < kZApRb && rb + 16 >= bk && 8 * (rb + 16) < m) ld1(rb + 16, vn);
            if (rb >= bk && 8 * rb < m) {               // rows above the block: V = 0
                // z conj(v) = (zr vr + zi vi) + i (zi vr - zr vi)
                dmma_884(wr[0][0], wr[0][1], zr[i][0], v[0].x);
                dmma_884(wi[0][0], wi[0][1], zi[i][0], v[0].x);
                dmma_884(wr[1][0], wr[1][1], zr[i][1], v[1].x);
                dmma_884(wi[1][0], wi[1][1], zi[i][1], v[1].x);
                dmma_884(wr[0][0], wr[0][1], zi[i][0], v[0].y);
                dmma_884(wi[0][0], wi[0][1], zr[i][0], -v[0].y);
                dmma_884(wr[1][0], wr[1][1], zi[i][1], v[1].y);
                dmma_884(wi[1][0], wi[1][1], zr[i][1], -v[1].y);
            }
        }
        part[warp][g * 8 + 2 * tg] = zmake(wr[0][0] + wr[1][0], wi[0][0] + wi[1][0]);
        part[warp][g * 8 + 2 * tg + 1] = zmake(wr[0][1] + wr[1][1], wi[0][1] + wi[1][1]);
        __syncthreads();
        if (tid < 64) {                                 // entry (c, k): sum the partials, then row i = k of T W
            const int i = tid & 7;
            zc s = zero;
#pragma unroll
            for (int w = 0; w < 16; ++w) s = zadd(s, part[w][tid]);
            zc acc = zero;
#pragma unroll
            for (int k = 0; k < 8; ++k) {               // (T W)(i, c) = sum_k T(i, k) W(k, c); W^T(c, k) sits in lane (base + k)
                const zc wk = zmake(__shfl_sync(0xffffffffu, s.x, (lane & ~7) + k), __shfl_sync(0xffffffffu, s.y, (lane & ~7) + k));
                acc = zfma(Tb[i + 8 * k], wk, acc);
            }
            ntw[tid] = zmake(-acc.x, -acc.y);
        }
        __syncthreads();
        // ---- step 2: Z^T += -(T W)^T V^T ----
        const zc a[2] = {ntw[g * 8 + tg], ntw[g * 8 + tg + 4]};
        auto ld2 = [&](int rb, zc *v) {
            const int r = 8 * rb + g;
            v[0] = vload(r, j0 + tg, rb == bk); v[1] = vload(r, j0 + tg + 4, rb == bk);
        };
        if (warp >= bk && 8 * warp < m) ld2(warp, vn);
#pragma unroll
        for (int i = 0; i < kZApRb; ++i) {
            const int rb = warp + 16 * i;
            const zc v[2] = {vn[0], vn[1]};
            if (i + 1 < kZApRb && rb + 16 >= bk && 8 * (rb + 16) < m) ld2(rb + 16, vn);
            if (rb >= bk && 8 * rb < m) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {           // a v = (ar vr - ai vi) + i (ar vi + ai vr)
                    dmma_884(zr[i][0], zr[i][1], a[h].x, v[h].x);
                    dmma_884(zi[i][0], zi[i][1], a[h].x, v[h].y);
                    dmma_884(zr[i][0], zr[i][1], a[h].y, -v[h].y);
                    dmma_884(zi[i][0], zi[i][1], a[h].y, v[h].x);
                }
            }
        }
    }
    if (col < rnew) {
#pragma unroll
        for (int i = 0; i < kZApRb; ++i) {
            const int r = 8 * (warp + 16 * i) + 2 * tg;
            if (r < m) out[r + (size_t)col * ar.ldc] = zmake(zr[i][0], zi[i][0]);
            if (r + 1 < m) out[r + 1 + (size_t)col * ar.ldc] = zmake(zr[i][1], zi[i][1]);
        }
    }
}

}  // namespace hcore_b200
