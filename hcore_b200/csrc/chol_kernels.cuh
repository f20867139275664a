// The remaining tile kernels of one TLR-Cholesky step (FP64, sm_100a) -- SURVEY.md section 8(f3):
//
//   HCORE_dtrsm    triangular solve on the V factor of a low-rank tile (or any m x n block)
//                  reference src/compute/hcore_dtrsm.c:24-48  (cblas_dtrsm, all side/uplo/trans/diag cases)
//   HCORE_dpotrf   Cholesky factorisation of a dense diagonal tile
//                  reference src/compute/hcore_dpotrf.c:23-35 (LAPACKE_dpotrf_work)
//
// Both are blocked over 64 x 64 diagonal blocks.  The diagonal blocks are inverted once in shared memory
// (8 x 8 blocks by substitution, then three levels of DMMA block products), so everything else is GEMM-shaped
// and runs on DMMA:
//   trsm   one CTA per (tile, 16 right-hand-side columns): the 16 columns live in shared memory for the whole
//          solve; per 64-row block  X_j <- inv(L_jj) X_j, then X_i -= L_ij X_j for all i > j (right looking, L
//          fragments straight from global / L2 into registers, 16 loads in flight per lane).
//   potrf  right looking over 64-column panels, three launches per panel so that a SINGLE tile (the one diagonal
//          tile of a Cholesky step) uses the whole GPU: diagonal block (Cholesky + inverse), panel (A_ij inv(L_jj)^T),
//          trailing update (A_il -= A_ij A_lj^T over the lower block pairs).
// Every side / uplo / trans combination of trsm is reduced to "forward substitution with a lower triangular
// matrix" by strided views (transposition = swapped strides, backward substitution = negative strides).
#pragma once
#include "common.cuh"

namespace hcore_b200 {

constexpr int kTb = 64;            // diagonal block size
constexpr int kTbLd = kTb + 4;     // 68: shared-memory leading dimension (conflict-free DMMA fragment loads)
constexpr int kTbBlk = kTb * kTbLd;                              // one 64 x 64 block in shared memory
constexpr size_t kCholSmemBytes = 3ull * kTbBlk * sizeof(double);   // three blocks (dynamic shared memory of every kernel below)

// element (i, j) at p[i * si + j * sj]; strides may be negative
struct SView {
    double *p;
    long si, sj;
    __device__ __forceinline__ double &at(int i, int j) const { return p[i * si + j * sj]; }
};

// ---- inverse of a 64 x 64 lower triangular block held in shared memory ------------------------------------------
// S: the block (ld kTbLd, strict upper part ignored), Inv: result (lower triangular, upper part zero), Tm: scratch.
// Whole CTA (256 threads).  8 x 8 diagonal blocks by forward substitution (one thread per column), then
//   inv([L11 0; L21 L22]) = [inv11 0; -inv22 (L21 inv11) inv22]   for block sizes 8, 16, 32.
__device__ __forceinline__ void invert_lower_64(const double *S, double *Inv, double *Tm) {
    const int tid = threadIdx.x, warp = warp_id(), nwarps = blockDim.x >> 5;
    for (int idx = tid; idx < kTb * kTbLd; idx += blockDim.x) Inv[idx] = 0.0;
    __syncthreads();
    if (tid < kTb) {
        const int o = tid & ~7, c = tid & 7;
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < i && k >= c) s -= S[(o + i) + (o + k) * kTbLd] * x[k];
            x[i] = (i >= c) ? s / S[(o + i) + (o + i) * kTbLd] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) Inv[(o + i) + (o + c) * kTbLd] = x[i];
    }
    __syncthreads();
    for (int h = 8; h < kTb; h <<= 1) {
        const int nblk = kTb / (2 * h);
        for (int b = 0; b < nblk; ++b) {             // T = L21 * inv11
            const int o = 2 * h * b;
            const int w = (warp + nwarps - (b % nwarps)) % nwarps;
            block_gemm<2, kGemmSet>(mview(Tm + (o + h) + o * kTbLd, 1, kTbLd), cview(S + (o + h) + o * kTbLd, 1, kTbLd),
                                    cview(Inv + o + o * kTbLd, 1, kTbLd), h, h, h, w, nwarps);
        }
        __syncthreads();
        for (int b = 0; b < nblk; ++b) {             // inv21 = -inv22 * T   (inv21 is still zero)
            const int o = 2 * h * b;
            const int w = (warp + nwarps - (b % nwarps)) % nwarps;
            block_gemm<2, kGemmSub>(mview(Inv + (o + h) + o * kTbLd, 1, kTbLd), cview(Inv + (o + h) + (o + h) * kTbLd, 1, kTbLd),
                                    cview(Tm + (o + h) + o * kTbLd, 1, kTbLd), h, h, h, w, nwarps);
        }
        __syncthreads();
    }
}

// =====================================================================================================
// trsm
// =====================================================================================================
struct TrsmArgs {
    const double *A;          // triangular tiles
    long strideA;             // 0: one tile shared by the whole batch (the diagonal tile of a Cholesky step)
    long a_si, a_sj;          // strides of the LOWER-triangular forward view of op(A) (may be negative)
    long a_off;               // element offset of its (0, 0)
    double *B;                // right-hand sides, overwritten by the solution
    long strideB;
    long b_si, b_sj, b_off;   // view of B as (msz x nrhs) in the forward orientation
    const int *Brk;           // per-tile rank (may be null: `rank` for every tile)
    int rank;
    int M;                    // the other dimension
    int side_right;           // 0: msz = M, nrhs = rank ; 1: msz = rank, nrhs = M
    int unit_diag;
    double alpha;
    double *Winv;             // [tilesA][nblk_max][kTb * kTb] inverses of the diagonal blocks
    int nblk_max;
};

__device__ __forceinline__ void trsm_dims(const TrsmArgs &ar, int t, int &msz, int &nrhs) {
    const int r = ar.Brk ? ar.Brk[t] : ar.rank;
    msz = ar.side_right ? r : ar.M;
    nrhs = ar.side_right ? ar.M : r;
}

// grid = (diagonal blocks, distinct A tiles), 256 threads
__global__ void __launch_bounds__(256) trsm_inv_kernel(TrsmArgs ar) {
    extern __shared__ __align__(16) double chol_sm[];
    double *S = chol_sm, *Inv = chol_sm + kTbBlk, *Tm = chol_sm + 2 * kTbBlk;
    const int t = blockIdx.y, d = blockIdx.x;
    int msz, nrhs;
    trsm_dims(ar, t, msz, nrhs);
    // the backward cases view the matrix through negative strides: offsets depend on the tile's own size
    const int j0 = d * kTb;
    if (j0 >= msz) return;
    const double *A = ar.A + (size_t)t * ar.strideA + (ar.a_si < 0 || ar.a_sj < 0 ? (long)(msz - 1) * (-(ar.a_si + ar.a_sj)) : 0) + ar.a_off;
    for (int idx = threadIdx.x; idx < kTb * kTb; idx += blockDim.x) {
        const int r = idx & (kTb - 1), c = idx >> 6;
        double v = (r == c) ? 1.0 : 0.0;
        if (j0 + r < msz && j0 + c < msz && r >= c && !(r == c && ar.unit_diag)) v = __ldg(A + (j0 + r) * ar.a_si + (j0 + c) * ar.a_sj);
        S[r + c * kTbLd] = v;
    }
    __syncthreads();
    invert_lower_64(S, Inv, Tm);
    double *W = ar.Winv + ((size_t)t * ar.nblk_max + d) * kTb * kTb;
    for (int idx = threadIdx.x; idx < kTb * kTb; idx += blockDim.x) W[idx] = Inv[(idx & (kTb - 1)) + (idx >> 6) * kTbLd];
}

__host__ __device__ __forceinline__ int trsm_ldx(int msz_pad) { return msz_pad + 4; }

// grid = (column chunks, tiles), 256 threads, dynamic shared memory = trsm_ldx(msz_pad) * NC doubles.
// NC = right-hand-side columns per CTA: 16, or 8 when 16 columns of the tallest tile do not fit shared memory.
template <int kTrsmNc>
__global__ void __launch_bounds__(256) trsm_solve_kernel(TrsmArgs ar, int msz_pad_max) {
    extern __shared__ __align__(16) double Xs[];
    constexpr int JT = kTrsmNc / 8;
    const int t = blockIdx.y, c0 = blockIdx.x * kTrsmNc;
    int msz, nrhs;
    trsm_dims(ar, t, msz, nrhs);
    if (c0 >= nrhs || msz <= 0) return;
    const int nc = (nrhs - c0 < kTrsmNc) ? (nrhs - c0) : kTrsmNc;
    const int nblk = (msz + kTb - 1) / kTb, msz_pad = nblk * kTb, ldx = trsm_ldx(msz_pad_max);
    const int tid = threadIdx.x, warp = warp_id(), lane = lane_id(), g = lane >> 2, tg = lane & 3;
    const bool rev = ar.a_si < 0 || ar.a_sj < 0;
    const int ta = ar.strideA ? t : 0;
    const double *A = ar.A + (size_t)ta * ar.strideA + (rev ? (long)(msz - 1) * (-(ar.a_si + ar.a_sj)) : 0) + ar.a_off;
    double *B = ar.B + (size_t)t * ar.strideB + (rev ? (long)(msz - 1) * (-ar.b_si) : 0) + ar.b_off + c0 * ar.b_sj;
    const double *Winv = ar.Winv + (size_t)ta * ar.nblk_max * kTb * kTb;

    // ---- alpha * B chunk -> shared memory (zero padded to msz_pad x 16) ------------------------------------
    if (ar.b_si == 1 || ar.b_si == -1) {
        for (int idx = tid; idx < msz_pad * kTrsmNc; idx += 256) {
            const int r = idx % msz_pad, c = idx / msz_pad;
            Xs[r + c * ldx] = (r < msz && c < nc) ? ar.alpha * B[r * ar.b_si + c * ar.b_sj] : 0.0;
        }
    } else {
        for (int idx = tid; idx < msz_pad * kTrsmNc; idx += 256) {
            const int c = idx % kTrsmNc, r = idx / kTrsmNc;
            Xs[r + c * ldx] = (r < msz && c < nc) ? ar.alpha * B[r * ar.b_si + c * ar.b_sj] : 0.0;
        }
    }
    __syncthreads();

    for (int jb = 0; jb < nblk; ++jb) {
        const int j0 = jb * kTb;
        // ---- X_j <- inv(L_jj) X_j: warp w owns rows 8 w .. 8 w + 7 of the block, both column tiles ----------
        {
            const double *W = Winv + (size_t)jb * kTb * kTb + (8 * warp + g);
            double la[16];
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) la[ks] = (4 * ks <= 8 * warp + 7) ? __ldg(W + (4 * ks + tg) * kTb) : 0.0;   // lower triangular
            double acc[JT][2];
#pragma unroll
            for (int jt = 0; jt < JT; ++jt) acc[jt][0] = acc[jt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
                if (4 * ks <= 8 * warp + 7) {
                    const double *xp = Xs + j0 + 4 * ks + tg + g * ldx;
#pragma unroll
                    for (int jt = 0; jt < JT; ++jt) dmma_884(acc[jt][0], acc[jt][1], la[ks], xp[8 * jt * ldx]);
                }
            }
            __syncthreads();
#pragma unroll
            for (int jt = 0; jt < JT; ++jt) {
                Xs[j0 + 8 * warp + g + (8 * jt + 2 * tg) * ldx] = acc[jt][0];
                Xs[j0 + 8 * warp + g + (8 * jt + 2 * tg + 1) * ldx] = acc[jt][1];
            }
            __syncthreads();
        }
        // ---- X_i -= L_ij X_j for every 8-row strip below the block ------------------------------------------
        const int nstrips = (msz_pad - j0 - kTb) >> 3;
        for (int s = warp; s < nstrips; s += 8) {
            const int i0 = j0 + kTb + 8 * s;
            const bool ron = i0 + g < msz;
            const double *ap = A + (long)(i0 + g) * ar.a_si + (long)(j0 + tg) * ar.a_sj;
            double la[16];
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) la[ks] = (ron && j0 + 4 * ks + tg < msz) ? __ldg(ap + (long)(4 * ks) * ar.a_sj) : 0.0;
            double acc[JT][2];
#pragma unroll
            for (int jt = 0; jt < JT; ++jt) acc[jt][0] = acc[jt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
                const double *xp = Xs + j0 + 4 * ks + tg + g * ldx;
#pragma unroll
                for (int jt = 0; jt < JT; ++jt) dmma_884(acc[jt][0], acc[jt][1], la[ks], xp[8 * jt * ldx]);
            }
#pragma unroll
            for (int jt = 0; jt < JT; ++jt) {
                Xs[i0 + g + (8 * jt + 2 * tg) * ldx] -= acc[jt][0];
                Xs[i0 + g + (8 * jt + 2 * tg + 1) * ldx] -= acc[jt][1];
            }
        }
        __syncthreads();
    }
    // ---- solution back to B ------------------------------------------------------------------------------------
    if (ar.b_si == 1 || ar.b_si == -1) {
        for (int idx = tid; idx < msz * nc; idx += 256) {
            const int r = idx % msz, c = idx / msz;
            B[r * ar.b_si + c * ar.b_sj] = Xs[r + c * ldx];
        }
    } else {
        for (int idx = tid; idx < msz * nc; idx += 256) {
            const int c = idx % nc, r = idx / nc;
            B[r * ar.b_si + c * ar.b_sj] = Xs[r + c * ldx];
        }
    }
}

// =====================================================================================================
// potrf
// =====================================================================================================
struct PotrfArgs {
    double *A;
    long strideA;
    long si, sj;              // lower-triangular view: (i, j) at A[i * si + j * sj]  (uplo = Upper: swapped strides)
    int n;
    int *info;                // per tile, 0 or the order of the first non-positive leading minor (LAPACK convention)
    double *Winv;             // [ntiles][kTb * kTb]: inverse of the current diagonal block
};

// grid = tiles.  Cholesky of the diagonal block A_jj (in shared memory, column by column, LAPACK dpotf2's
// arithmetic: sqrt of the pivot, column scaled by its reciprocal, rank-1 update), then its inverse.
__global__ void __launch_bounds__(256) potrf_diag_kernel(PotrfArgs ar, int j0) {
    extern __shared__ __align__(16) double chol_sm[];
    double *S = chol_sm, *Inv = chol_sm + kTbBlk, *Tm = chol_sm + 2 * kTbBlk;
    const int t = blockIdx.x, tid = threadIdx.x;
    if (ar.info[t] != 0) return;
    double *A = ar.A + (size_t)t * ar.strideA;
    const int jb = (ar.n - j0 < kTb) ? (ar.n - j0) : kTb;
    for (int idx = tid; idx < kTb * kTb; idx += 256) {
        const int r = idx & (kTb - 1), c = idx >> 6;
        S[r + c * kTbLd] = (r < jb && c < jb && r >= c) ? A[(j0 + r) * ar.si + (j0 + c) * ar.sj] : ((r == c) ? 1.0 : 0.0);
    }
    __syncthreads();
    // One barrier per column (three in the first version: pivot, scaling, update).  Every thread reads the pivot d
    // itself and derives 1/sqrt(d) and 1/d redundantly; the scaled column L(:, c) = S(:, c) / sqrt(d) goes to a SEPARATE
    // block (Tm), so the rank-1 update S(r, cc) -= S(r, c) S(cc, c) / d can read the unscaled column while it is
    // being written out.
    int bad = 0;
    for (int c = 0; c < jb; ++c) {
        const double d = S[c + c * kTbLd];
        if (!(d > 0.0)) { bad = c + 1; break; }       // also catches NaN (LAPACK: ajj <= 0 or disnan); uniform over the CTA
        const double dinv = rsqrt(d), dinv2 = dinv * dinv;
        if (tid >= c && tid < kTb) Tm[tid + c * kTbLd] = (tid == c) ? d * dinv : S[tid + c * kTbLd] * dinv;
        const int nc = kTb - 1 - c;
        for (int idx = tid; idx < nc * kTb; idx += 256) {
            const int r = idx & (kTb - 1), cc = c + 1 + (idx >> 6);
            if (r >= cc) S[r + cc * kTbLd] -= (S[r + c * kTbLd] * S[cc + c * kTbLd]) * dinv2;
        }
        __syncthreads();
    }
    __syncthreads();
    // the factored columns come from Tm; on failure LAPACK leaves the columns factored so far and the failing column
    // (and what lies to its right) keeps its updated values
    const int ngood = bad ? bad - 1 : jb;
    for (int idx = tid; idx < kTb * kTb; idx += 256) {
        const int r = idx & (kTb - 1), c = idx >> 6;
        if (r >= c && c < ngood) S[r + c * kTbLd] = Tm[r + c * kTbLd];
    }
    __syncthreads();
    for (int idx = tid; idx < kTb * kTb; idx += 256) {
        const int r = idx & (kTb - 1), c = idx >> 6;
        if (r < jb && c < jb && r >= c) A[(j0 + r) * ar.si + (j0 + c) * ar.sj] = S[r + c * kTbLd];
    }
    if (bad) {
        if (tid == 0) ar.info[t] = j0 + bad;
        return;
    }
    invert_lower_64(S, Inv, Tm);
    double *W = ar.Winv + (size_t)t * kTb * kTb;
    for (int idx = tid; idx < kTb * kTb; idx += 256) W[idx] = Inv[(idx & (kTb - 1)) + (idx >> 6) * kTbLd];
}

// grid = (row blocks below the diagonal block, tiles): A_ij <- A_ij inv(L_jj)^T
__global__ void __launch_bounds__(256) potrf_panel_kernel(PotrfArgs ar, int j0) {
    extern __shared__ __align__(16) double chol_sm[];
    double *As = chol_sm, *Li = chol_sm + kTbBlk, *Out = chol_sm + 2 * kTbBlk;
    const int t = blockIdx.y, tid = threadIdx.x;
    if (ar.info[t] != 0) return;
    double *A = ar.A + (size_t)t * ar.strideA;
    const int i0 = j0 + kTb * (1 + blockIdx.x);
    const int rb = (ar.n - i0 < kTb) ? (ar.n - i0) : kTb, jb = (ar.n - j0 < kTb) ? (ar.n - j0) : kTb;
    const double *W = ar.Winv + (size_t)t * kTb * kTb;
    if (ar.si == 1) {
        for (int idx = tid; idx < kTb * kTb; idx += 256) {
            const int r = idx & (kTb - 1), c = idx >> 6;
            As[r + c * kTbLd] = (r < rb && c < jb) ? A[(i0 + r) * ar.si + (j0 + c) * ar.sj] : 0.0;
            Li[r + c * kTbLd] = W[idx];
        }
    } else {
        for (int idx = tid; idx < kTb * kTb; idx += 256) {
            const int c = idx & (kTb - 1), r = idx >> 6;
            As[r + c * kTbLd] = (r < rb && c < jb) ? A[(i0 + r) * ar.si + (j0 + c) * ar.sj] : 0.0;
            Li[(idx & (kTb - 1)) + (idx >> 6) * kTbLd] = W[idx];
        }
    }
    __syncthreads();
    // Out(r, c) = sum_k As(r, k) Li(c, k)
    block_gemm<2, kGemmSet>(mview(Out, 1, kTbLd), cview(As, 1, kTbLd), cview(Li, kTbLd, 1), kTb, kTb, kTb, warp_id(), 8);
    __syncthreads();
    if (ar.si == 1) {
        for (int idx = tid; idx < kTb * kTb; idx += 256) {
            const int r = idx & (kTb - 1), c = idx >> 6;
            if (r < rb && c < jb) A[(i0 + r) * ar.si + (j0 + c) * ar.sj] = Out[r + c * kTbLd];
        }
    } else {
        for (int idx = tid; idx < kTb * kTb; idx += 256) {
            const int c = idx & (kTb - 1), r = idx >> 6;
            if (r < rb && c < jb) A[(i0 + r) * ar.si + (j0 + c) * ar.sj] = Out[r + c * kTbLd];
        }
    }
}

// grid = (lower block pairs (i >= l) of the trailing matrix, tiles): A_il -= A_ij A_lj^T
__global__ void __launch_bounds__(256) potrf_trail_kernel(PotrfArgs ar, int j0) {
    extern __shared__ __align__(16) double chol_sm[];
    double *Ai = chol_sm, *Al = chol_sm + kTbBlk, *Out = chol_sm + 2 * kTbBlk;
    const int t = blockIdx.y, tid = threadIdx.x;
    if (ar.info[t] != 0) return;
    double *A = ar.A + (size_t)t * ar.strideA;
    const int p = blockIdx.x;
    int bi = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
    while ((bi + 1) * (bi + 2) / 2 <= p) ++bi;
    while (bi * (bi + 1) / 2 > p) --bi;
    const int bl = p - bi * (bi + 1) / 2;
    const int i0 = j0 + kTb * (1 + bi), l0 = j0 + kTb * (1 + bl);
    const int rb = (ar.n - i0 < kTb) ? (ar.n - i0) : kTb, cb = (ar.n - l0 < kTb) ? (ar.n - l0) : kTb;
    const int jb = (ar.n - j0 < kTb) ? (ar.n - j0) : kTb;
    const bool fast = ar.si == 1;
    for (int idx = tid; idx < kTb * kTb; idx += 256) {
        const int r = fast ? (idx & (kTb - 1)) : (idx >> 6), c = fast ? (idx >> 6) : (idx & (kTb - 1));
        Ai[r + c * kTbLd] = (r < rb && c < jb) ? A[(i0 + r) * ar.si + (j0 + c) * ar.sj] : 0.0;
        Al[r + c * kTbLd] = (r < cb && c < jb) ? A[(l0 + r) * ar.si + (j0 + c) * ar.sj] : 0.0;
    }
    __syncthreads();
    block_gemm<2, kGemmSet>(mview(Out, 1, kTbLd), cview(Ai, 1, kTbLd), cview(Al, kTbLd, 1), kTb, kTb, kTb, warp_id(), 8);
    __syncthreads();
    const bool diag = (bi == bl);
    for (int idx = tid; idx < kTb * kTb; idx += 256) {
        const int r = fast ? (idx & (kTb - 1)) : (idx >> 6), c = fast ? (idx >> 6) : (idx & (kTb - 1));
        if (r < rb && c < cb && !(diag && r < c)) A[(i0 + r) * ar.si + (l0 + c) * ar.sj] -= Out[r + c * kTbLd];   // the strict upper triangle is never touched
    }
}

}  // namespace hcore_b200
