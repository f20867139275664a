// Batched tile-low-rank update + recompression kernels (FP64, sm_100a).
//
// Pipeline for one batch chunk (one launch each, every CTA works on one tile or panel):
//
//   tlr_stack_kernel   P = AV^T BV, G = BU P^T, panels X_u = [beta CU | alpha AU],
//                      X_v = [CV | G]                (reference __qra/__qrb up to the QR,
//                      src/compute/hcore_dgemm.c:96-114, :175-205)
//   tlr_tsqr_kernel    flat-tree tall-skinny Householder QR of both panels, 128-row
//                      blocks resident in shared memory  (replaces LAPACKE_dgeqrf,
//                      src/compute/hcore_dgemm.c:123, :216)
//   tlr_svd_kernel     T' = R_v R_u^T, QR with column pivoting as preconditioner,
//                      one-sided Jacobi on the columns of R^T, sort, rank truncation,
//                      U_r and W = R_v (R_u^T U_r)         (replaces the core product,
//                      LAPACKE_dgesvd, the truncation loop and the sigma scaling,
//                      src/compute/hcore_dgemm.c:330-489)
//   tlr_apply_kernel   CU = Q_u [U_r; 0], CV = Q_v [W; 0] by applying the stored block
//                      reflectors  (replaces LAPACKE_dormqr + copies/transposes,
//                      src/compute/hcore_dgemm.c:531-598, :645-674, :972-982)
//
// tests/device_model.py restates exactly this arithmetic in numpy.
#pragma once
#include "common.cuh"

namespace hcore_b200 {

// ---- argument blocks -------------------------------------------------------------------

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
};

// Per-chunk scratch owned by the context.  Index [panel][tile] with panel 0 = U, 1 = V.
struct TlrWorkspace {
    double *panel;   // [2][chunk][m_pad * (N + 8)]  stacked panels, then reflectors (+R in block 0)
    double *R;       // [2][chunk][N * N]
    double *T;       // [2][chunk][nblk][N/16][256]  compact-WY T factors
    double *Y;       // [2][chunk][N * N]            U_r (panel 0) and W (panel 1)
    double *sigma;   // [chunk][N]                   sorted singular values
    int *ncols;      // [chunk]  a + c
    int *newrank;    // [chunk]
    int *st;         // [chunk]  internal status
    int *sweeps;     // [chunk]  Jacobi sweeps used (diagnostics)
    int chunk, m_pad, nblk;
};

template <int N>
struct Geometry {
    static constexpr int kPanelCols = N + 8;      // allocation width of a panel (see stack kernel)
    static constexpr int kNsp = N / kPanelW;      // sub-panels at full width
    static constexpr int kLdB = kBlockH + 4;      // shared-memory leading dim of a 128-row block
    static constexpr int kLdS = kPanelW + 4;      // leading dim of the 16-row strips
    static constexpr int kLdL = N + 4;            // leading dim of the N x N core matrix
};

__device__ __forceinline__ size_t panel_offset(const TlrWorkspace &ws, int panel, int t, int panel_cols) {
    return ((size_t)panel * ws.chunk + t) * (size_t)ws.m_pad * panel_cols;
}
__device__ __forceinline__ size_t nn_offset(const TlrWorkspace &ws, int panel, int t, int n) {
    return ((size_t)panel * ws.chunk + t) * (size_t)n * n;
}
__device__ __forceinline__ size_t t_offset(const TlrWorkspace &ws, int panel, int t, int nsp_full) {
    return ((size_t)panel * ws.chunk + t) * (size_t)ws.nblk * nsp_full * 256;
}

// ========================================================================================
// 1. stacking kernel
// ========================================================================================
// One CTA (512 threads) per tile.  Dynamic shared memory: P, padded (N x (N+4)) doubles.
template <int N>
__global__ void __launch_bounds__(512) tlr_stack_kernel(TlrArgs ar, TlrWorkspace ws) {
    using G = Geometry<N>;
    extern __shared__ double smem[];
    double *P = smem;                 // P[i + j * kLdL], i < a (row), j < b (col)
    const int t = blockIdx.x;
    const int gt = ar.tile0 + t;
    const int tid = threadIdx.x;
    const int a = ar.Ark[gt], b = ar.Brk[gt], c = ar.Crk[gt];
    const int n = a + c;
    const int m = ar.nb;

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

    const int m_pad = ws.m_pad;
    double *XU = ws.panel + panel_offset(ws, 0, t, G::kPanelCols);
    double *XV = ws.panel + panel_offset(ws, 1, t, G::kPanelCols);
    const double *AU = ar.AU[gt], *AV = ar.AV[gt], *BU = ar.BU[gt], *BV = ar.BV[gt];
    const double *CU = ar.CU[gt], *CV = ar.CV[gt];

    // X_u = [beta*CU | alpha*AU | 0], X_v[:, :c] = CV, everything else of X_v zeroed
    // (G is written by the GEMM below); rows m..m_pad-1 are zero padding.
    // One warp per column, lanes along the rows (coalesced), four loads in flight per lane.
    {
        const int warp = warp_id(), lane = lane_id(), nwarps = blockDim.x >> 5;
        for (int col = warp; col < N; col += nwarps) {
            const double *su = nullptr, *sv = nullptr;
            double scale = 0.0;
            if (col < c) { su = CU + (size_t)col * ar.ldc; sv = CV + (size_t)col * ar.ldc; scale = ar.beta; }
            else if (col < n) { su = AU + (size_t)(col - c) * ar.lda; scale = ar.alpha; }
            double *du = XU + (size_t)col * m_pad, *dv = XV + (size_t)col * m_pad;
            for (int r0 = 0; r0 < m_pad; r0 += 128) {
                double xu[4], xv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = r0 + lane + 32 * u;
                    xu[u] = (su && r < m) ? __ldg(su + r) : 0.0;
                    xv[u] = (sv && r < m) ? __ldg(sv + r) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = r0 + lane + 32 * u;      // m_pad is a multiple of 128
                    du[r] = scale * xu[u];
                    dv[r] = xv[u];
                }
            }
        }
    }

    // P = AV^T * BV   (a x b, contraction over the m rows)
    const int a8 = round_up(a, 8), b8 = round_up(b, 8), m4 = round_up(m, 4);
    {
        GuardedView Av{AV, (long)ar.lda, 1, a, m};   // A(i, k) = AV[k + i*lda]
        GuardedView Bv{BV, 1, (long)ar.ldb, m, b};   // B(k, j) = BV[k + j*ldb]
        block_gemm<4, kGemmSet>(mview(P, 1, G::kLdL), Av, Bv, a8, b8, m4, warp_id(), blockDim.x >> 5);
    }
    __syncthreads();
    // G = BU * P^T  -> X_v[:, c : c+a]
    {
        const int m8 = round_up(m, 8);
        GuardedView Bu{BU, 1, (long)ar.ldb, m, b};   // A(i, k) = BU[i + k*ldb]
        CMatView Pt = cview(P, G::kLdL, 1);          // B(k, j) = P[j + k*ldl]
        block_gemm<4, kGemmSet>(mview(XV + (size_t)c * m_pad, 1, m_pad), Bu, Pt, m8, a8, b8, warp_id(),
                                blockDim.x >> 5);
    }
}

// ========================================================================================
// 2. tall-skinny QR kernel
// ========================================================================================
// Flat-tree TSQR: the panel is processed in 128-row blocks that live in shared memory.
// Block 0 is a plain blocked Householder QR (reflectors below the diagonal, R on top);
// every further block B_j is eliminated against the running R ("triangle on top of a dense
// block"): reflector i = [e_i ; v_i] touches row i of R and all of B_j.  Column sub-panels of
// width 16 with compact-WY T factors.
//
// The columns are processed LEFT-LOOKING in groups of kGroupW = 64: for group g the CTA walks
// over all row blocks, first applies the stored reflectors of the earlier groups to the
// group's 64 columns, then factors the group's own four sub-panels.  Only a 128 x 64 slice of
// a block is resident (110 KiB of shared memory), so two CTAs share an SM and one CTA's
// latency-bound panel steps overlap the other's GEMM phases; the arithmetic (and its order
// per matrix element) is the same as in the right-looking formulation.
//
// Per sub-panel
//   (a) panel factorisation by warps 0-3: one row per thread, the 16 panel columns in
//       registers; per Householder step ONE batched 16-value reduction (the dots of the
//       current column with every panel column give the norm, the in-panel update AND the
//       V^T v products for the compact-WY T factor), one 128-thread named barrier;
//   (b) block-reflector update of the trailing columns with three DMMA GEMMs
//       S = R_strip + V^T B,  S2 = T^T S,  B -= V S2,  R_strip -= S2   (all warps).
template <int N>
struct TsqrSmem {
    using G = Geometry<N>;
    static constexpr int kGroupW = (N < 64) ? N : 64;
    static constexpr int kB = 0;
    static constexpr int kVp = kB + kGroupW * G::kLdB;
    static constexpr int kS = kVp + kPanelW * G::kLdB;
    static constexpr int kS2 = kS + kGroupW * G::kLdS;
    static constexpr int kTs = kS2 + kGroupW * G::kLdS;
    static constexpr int kRd = kTs + 256;       // diagonal 16x16 block of the R strip, row-major
    static constexpr int kRfin = kRd + 256;     // its final value, row-major
    static constexpr int kRed = kRfin + 256;    // [2][4 warps][16] reduction partials
    static constexpr int kTau = kRed + 128;
    static constexpr int kTotal = kTau + 16;
    static constexpr size_t kBytes = (size_t)kTotal * sizeof(double);
};

constexpr int kTsqrThreads = 256;
constexpr int kPanelThreads = 128;   // warps 0-3 factor the sub-panel
constexpr int kPanelBarrier = 1;

// Block-reflector update of `ncols` columns of the resident block slice (B + col_off*LDB):
//   S (pre-loaded with the R strip or zero) += Vp^T B ; S2 = T^T S ; R strip -= S2 ; B -= Vp S2
// rrow0 / rcol0 locate the strip in the global R.  Ends with a __syncthreads().
template <int N>
__device__ __forceinline__ void tsqr_block_update(double *Bc, const double *Vp, const double *Ts, double *S,
                                                  double *S2, double *Rg, bool blk0, int rrow0, int rcol0,
                                                  int ncols, int tid, int warp, int nwarps) {
    constexpr int LDB = Geometry<N>::kLdB, LDS = Geometry<N>::kLdS;
    block_gemm<1, kGemmAdd>(mview(S, 1, LDS), cview(Vp, LDB, 1), cview(Bc, 1, LDB), kPanelW, ncols, kBlockH, warp,
                            nwarps);
    __syncthreads();
    block_gemm<1, kGemmSet>(mview(S2, 1, LDS), cview(Ts, 16, 1), cview(S, 1, LDS), kPanelW, ncols, kPanelW, warp,
                            nwarps);
    __syncthreads();
    if (!blk0) {
        for (int idx = tid; idx < kPanelW * ncols; idx += kTsqrThreads) {
            const int ii = idx & 15, jj = idx >> 4;
            Rg[(size_t)(rrow0 + ii) + (size_t)(rcol0 + jj) * N] -= S2[ii + jj * LDS];
        }
    }
    block_gemm<2, kGemmSub>(mview(Bc, 1, LDB), cview(Vp, 1, LDB), cview(S2, 1, LDS), kBlockH, ncols, kPanelW, warp,
                            nwarps);
    __syncthreads();
}

// One CTA (256 threads) per panel: blockIdx.x = 2 * tile + panel.
template <int N>
__global__ void __launch_bounds__(kTsqrThreads, 2) tlr_tsqr_kernel(TlrWorkspace ws) {
    using G = Geometry<N>;
    using SM = TsqrSmem<N>;
    constexpr int LDB = G::kLdB, LDS = G::kLdS, GW = SM::kGroupW;
    extern __shared__ double smem[];
    double *B = smem + SM::kB, *Vp = smem + SM::kVp, *S = smem + SM::kS, *S2 = smem + SM::kS2,
           *Ts = smem + SM::kTs, *Rd = smem + SM::kRd, *Rfin = smem + SM::kRfin, *red = smem + SM::kRed,
           *tauv = smem + SM::kTau;

    const int t = blockIdx.x >> 1, pnl = blockIdx.x & 1;
    if (ws.st[t] != HCORE_B200_OK) return;
    const int tid = threadIdx.x, warp = warp_id(), lane = lane_id();
    const int nwarps = kTsqrThreads / 32;
    const int n16 = round_up(ws.ncols[t], kPanelW);
    const int m_pad = ws.m_pad;
    double *panel = ws.panel + panel_offset(ws, pnl, t, G::kPanelCols);
    double *Rg = ws.R + nn_offset(ws, pnl, t, N);
    double *Tg = ws.T + t_offset(ws, pnl, t, G::kNsp);

    for (int gc0 = 0; gc0 < n16; gc0 += GW) {                     // column groups, left to right
        const int gcols = (n16 - gc0 < GW) ? (n16 - gc0) : GW;    // multiple of 16
        for (int j = 0; j < ws.nblk; ++j) {
            const bool blk0 = (j == 0);
            // ---- load the group's slice of the row block --------------------------------
            for (int idx = tid; idx < kBlockH * gcols; idx += kTsqrThreads) {
                const int rr = idx & (kBlockH - 1), cc = idx >> 7;
                B[rr + cc * LDB] = panel[(size_t)j * kBlockH + rr + (size_t)(gc0 + cc) * m_pad];
            }
            // ---- (1) reflectors of the earlier groups (left-looking) ------------------------
            for (int sp = 0; sp < gc0 / kPanelW; ++sp) {
                const int pc0 = sp * kPanelW;
                for (int idx = tid; idx < kBlockH * kPanelW; idx += kTsqrThreads) {
                    const int rr = idx & 127, cc = idx >> 7, col = pc0 + cc;
                    double x = panel[(size_t)j * kBlockH + rr + (size_t)col * m_pad];
                    if (blk0) x = (rr > col) ? x : (rr == col ? 1.0 : 0.0);
                    Vp[rr + cc * LDB] = x;
                }
                Ts[tid] = Tg[((size_t)j * G::kNsp + sp) * 256 + tid];
                for (int idx = tid; idx < kPanelW * gcols; idx += kTsqrThreads) {
                    const int ii = idx & 15, jj = idx >> 4;
                    S[ii + jj * LDS] = blk0 ? 0.0 : Rg[(size_t)(pc0 + ii) + (size_t)(gc0 + jj) * N];
                }
                __syncthreads();
                tsqr_block_update<N>(B, Vp, Ts, S, S2, Rg, blk0, pc0, gc0, gcols, tid, warp, nwarps);
            }
            __syncthreads();

            // ---- (2) the group's own sub-panels ----------------------------------------------
            for (int c0 = gc0; c0 < gc0 + gcols; c0 += kPanelW) {
                const int s = c0 / kPanelW;
                const int c1 = c0 + kPanelW;
                const int lc0 = c0 - gc0, lc1 = c1 - gc0;          // local column offsets in the slice
                const int ntrail = gcols - lc1;

                if (warp < kPanelThreads / 32) {
                    // ================= (a) panel factorisation, warps 0-3 =====================
                    const int r = tid;                   // row of the block owned by this thread
                    double v[kPanelW];
#pragma unroll
                    for (int c = 0; c < kPanelW; ++c) v[c] = B[r + (lc0 + c) * LDB];
                    if (!blk0) {
                        // diagonal block of the running R (rows/cols c0..c0+15), row-major
                        for (int idx = tid; idx < 256; idx += kPanelThreads) {
                            const int ii = idx >> 4, c = idx & 15;
                            Rd[idx] = Rg[(size_t)(c0 + ii) + (size_t)(c0 + c) * N];
                        }
                    }
                    for (int idx = tid; idx < 256; idx += kPanelThreads) { Ts[idx] = 0.0; Rfin[idx] = 0.0; }
                    named_barrier(kPanelBarrier, kPanelThreads);

#pragma unroll
                    for (int ii = 0; ii < kPanelW; ++ii) {
                        const int i = c0 + ii;           // global column (= diagonal row in block 0)
                        double *redb = red + (ii & 1) * 64;
                        if (blk0 && r == i) {            // this row becomes row ii of the R strip
#pragma unroll
                            for (int c = 0; c < kPanelW; ++c) Rd[ii * 16 + c] = v[c];
                        }
                        double xi = v[ii];
                        if (blk0 && r <= i) xi = 0.0;    // block 0: only rows below the diagonal take part
                        double pr[kPanelW];
#pragma unroll
                        for (int c = 0; c < kPanelW; ++c) pr[c] = xi * v[c];
                        const double wsum = reduce16(pr, lane);
                        if ((lane & 1) == 0) redb[warp * 16 + reduce16_col(lane)] = wsum;
                        named_barrier(kPanelBarrier, kPanelThreads);

                        // totals d[c] = x_i . column c  (all of them only in warp 0, which builds T)
                        double d[kPanelW];
#pragma unroll
                        for (int c = 0; c < kPanelW; c += 2) {
                            if (c + 1 >= ii || warp == 0) {
                                const double2 p0 = *reinterpret_cast<const double2 *>(redb + c);
                                const double2 p1 = *reinterpret_cast<const double2 *>(redb + 16 + c);
                                const double2 p2 = *reinterpret_cast<const double2 *>(redb + 32 + c);
                                const double2 p3 = *reinterpret_cast<const double2 *>(redb + 48 + c);
                                d[c] = (p0.x + p1.x) + (p2.x + p3.x);
                                d[c + 1] = (p0.y + p1.y) + (p2.y + p3.y);
                            } else {
                                d[c] = d[c + 1] = 0.0;
                            }
                        }
                        const double alpha = Rd[ii * 16 + ii];
                        const Householder h = make_householder(alpha, d[ii]);
                        const double vr = xi * h.scal;
                        const bool writer = (r == kPanelThreads - 1);
#pragma unroll
                        for (int c = ii + 1; c < kPanelW; ++c) {
                            const double rd = Rd[ii * 16 + c];
                            const double w = h.tau * fma(d[c], h.scal, rd);
                            v[c] = fma(-vr, w, v[c]);
                            if (writer) Rfin[ii * 16 + c] = rd - w;
                        }
                        // column ii now holds the reflector tail; in block 0 the rows on and above
                        // the diagonal keep their R entries
                        if (!(blk0 && r <= i)) v[ii] = vr;
                        if (writer) { Rfin[ii * 16 + ii] = h.beta; tauv[ii] = h.tau; }
                        // compact-WY T, column ii (forward/columnwise larft): lane k of warp 0 owns row k
                        //   T[k][ii] = -tau * sum_{c=k}^{ii-1} T[k][c] * z_c,  z_c = V[:,c]^T v_ii
                        if (warp == 0 && lane < kPanelW) {
                            const int k = lane;
                            double accv = 0.0;
#pragma unroll
                            for (int c = 0; c < ii; ++c) {
                                const double z = fma(d[c], h.scal, Rd[ii * 16 + c]);
                                if (c >= k) accv = fma(Ts[k + c * 16], z, accv);
                            }
                            Ts[k + ii * 16] = (k < ii) ? -h.tau * accv : (k == ii ? h.tau : 0.0);
                        }
                    }
                    named_barrier(kPanelBarrier, kPanelThreads);   // Rfin complete

                    // explicit reflector panel + write-back of the block columns
#pragma unroll
                    for (int c = 0; c < kPanelW; ++c) {
                        const int col = c0 + c;
                        double vv = v[c], bb = v[c];
                        if (blk0) {
                            if (r < col) vv = 0.0;
                            else if (r == col) vv = 1.0;
                            if (r >= c0 && r <= col) bb = Rfin[(r - c0) * 16 + c];
                        }
                        Vp[r + c * LDB] = vv;
                        B[r + (lc0 + c) * LDB] = bb;
                    }
                    if (!blk0) {   // finished diagonal block of the running R
                        for (int idx = tid; idx < 256; idx += kPanelThreads) {
                            const int ii = idx >> 4, c = idx & 15;
                            Rg[(size_t)(c0 + ii) + (size_t)(c0 + c) * N] = (c >= ii) ? Rfin[idx] : 0.0;
                        }
                    }
                } else {
                    // warps 4-7 meanwhile stage S <- trailing part of the R strip (or zero)
                    for (int idx = tid - kPanelThreads; idx < kPanelW * ntrail; idx += kTsqrThreads - kPanelThreads) {
                        const int ii = idx & 15, jj = idx >> 4;
                        S[ii + jj * LDS] = blk0 ? 0.0 : Rg[(size_t)(c0 + ii) + (size_t)(c1 + jj) * N];
                    }
                }
                __syncthreads();

                // T factor of this (block, sub-panel) to global
                Tg[((size_t)j * G::kNsp + s) * 256 + tid] = Ts[tid];
                // ================= (b) trailing update inside the group, all warps ================
                if (ntrail > 0)
                    tsqr_block_update<N>(B + lc1 * LDB, Vp, Ts, S, S2, Rg, blk0, c0, c1, ntrail, tid, warp, nwarps);
                else
                    __syncthreads();
            }  // sub-panels

            // ---- store the slice (reflectors; block 0 also carries R) ---------------------
            for (int idx = tid; idx < kBlockH * gcols; idx += kTsqrThreads) {
                const int rr = idx & (kBlockH - 1), cc = idx >> 7;
                panel[(size_t)j * kBlockH + rr + (size_t)(gc0 + cc) * m_pad] = B[rr + cc * LDB];
            }
            if (blk0) {
                for (int idx = tid; idx < N * gcols; idx += kTsqrThreads) {
                    const int i = idx % N, cc = idx / N, col = gc0 + cc;
                    Rg[(size_t)i + (size_t)col * N] = (i <= col) ? B[i + cc * LDB] : 0.0;
                }
                if (gc0 + gcols >= n16) {     // zero the padding columns n16..N-1 once
                    for (int idx = tid; idx < N * (N - n16); idx += kTsqrThreads) Rg[(size_t)n16 * N + idx] = 0.0;
                }
            }
            __syncthreads();   // R / panel writes visible to this CTA's later loads, B free for reuse
        }  // row blocks
    }  // column groups
}

// ========================================================================================
// 3. core SVD kernel
// ========================================================================================
struct UpperTimesUpperT {   // T'(i,j) = sum_k Rv(i,k) Ru(j,k): only k >= max(i,j) contributes
    __device__ __forceinline__ void operator()(int i0, int j0, int &kb, int &) const {
        kb = (i0 > j0 ? i0 : j0) & ~3;
    }
};
struct UpperTransK {        // sum_k Ru(k,i) Y(k,j): Ru(k,i) = 0 for k > i
    int K;
    __device__ __forceinline__ void operator()(int i0, int, int &, int &ke) const {
        const int lim = i0 + 8;
        ke = lim < K ? lim : K;
    }
};
struct UpperK {             // sum_k Rv(i,k) M(k,j): Rv(i,k) = 0 for k < i
    __device__ __forceinline__ void operator()(int i0, int, int &kb, int &) const { kb = i0 & ~3; }
};

constexpr double kTolRot = 1e-15;
constexpr double kTolStop = 1e-7;
constexpr int kMaxSweeps = 30;

template <int N>
struct SvdSmem {
    using G = Geometry<N>;
    static constexpr int kL = 0;
    static constexpr int kVn = kL + N * G::kLdL;
    static constexpr int kVvec = kVn + N;
    static constexpr int kSig = kVvec + N;
    static constexpr int kSsort = kSig + N;
    static constexpr int kInts = kSsort + N;          // perm[N], posof[N], misc[8] as ints
    static constexpr int kTotal = kInts + N + 8;      // 2N+16 ints fit in N+8 doubles
    static constexpr size_t kBytes = (size_t)kTotal * sizeof(double);
};

// One CTA (512 threads) per tile.
template <int N>
__global__ void __launch_bounds__(512, 1) tlr_svd_kernel(TlrArgs ar, TlrWorkspace ws) {
    using G = Geometry<N>;
    using SM = SvdSmem<N>;
    constexpr int LDL = G::kLdL;
    extern __shared__ double smem[];
    double *L = smem + SM::kL, *vn = smem + SM::kVn, *vvec = smem + SM::kVvec, *sig = smem + SM::kSig,
           *ssort = smem + SM::kSsort;
    int *perm = reinterpret_cast<int *>(smem + SM::kInts);
    int *posof = perm + N;
    int *misc = posof + N;    // [0] pivot, [1] new rank
    double *dmisc = smem + SM::kTotal - 2;  // [0] tau of the current QRCP step (inside the int area tail)

    const int t = blockIdx.x;
    if (ws.st[t] != HCORE_B200_OK) return;
    const int gt = ar.tile0 + t;
    const int tid = threadIdx.x, warp = warp_id(), lane = lane_id();
    const int nwarps = blockDim.x >> 5;
    const int n = ws.ncols[t];
    const int n16 = round_up(n, kPanelW);
    const double *Ru = ws.R + nn_offset(ws, 0, t, N);
    const double *Rv = ws.R + nn_offset(ws, 1, t, N);
    double *Yu = ws.Y + nn_offset(ws, 0, t, N);
    double *Yv = ws.Y + nn_offset(ws, 1, t, N);

    // ---- T' = R_v * R_u^T --------------------------------------------------------------
    block_gemm<2, kGemmSet>(mview(L, 1, LDL), cview(Rv, 1, N), cview(Ru, N, 1), n16, n16, n16, warp, nwarps,
                            UpperTimesUpperT());
    if (tid < n16) perm[tid] = tid;
    __syncthreads();

    // ---- QR with column pivoting (preconditioner) ----------------------------------------
    for (int cc = warp; cc < n16; cc += nwarps) {
        double s2 = 0.0;
        for (int rr = lane; rr < n16; rr += 32) { const double x = L[rr + cc * LDL]; s2 = fma(x, x, s2); }
        s2 = warp_sum(s2);
        if (lane == 0) vn[cc] = s2;
    }
    __syncthreads();
    for (int i = 0; i < n16; ++i) {
        if (warp == 0) {      // pivot = argmax of the down-dated squared column norms
            double best = -1.0; int bi = i;
            for (int cc = i + lane; cc < n16; cc += 32) { const double x = vn[cc]; if (x > best) { best = x; bi = cc; } }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (lane == 0) misc[0] = bi;
        }
        __syncthreads();
        const int piv = misc[0];
        if (piv != i) {
            if (tid < n16) {
                const double x = L[tid + i * LDL];
                L[tid + i * LDL] = L[tid + piv * LDL];
                L[tid + piv * LDL] = x;
            }
            if (tid == 0) {
                const int pi = perm[i]; perm[i] = perm[piv]; perm[piv] = pi;
                const double x = vn[i]; vn[i] = vn[piv]; vn[piv] = x;
            }
        }
        __syncthreads();
        if (warp == 0) {      // Householder vector of column i (rows i..n16-1)
            double s2 = 0.0;
            for (int rr = i + 1 + lane; rr < n16; rr += 32) { const double x = L[rr + i * LDL]; s2 = fma(x, x, s2); }
            s2 = warp_sum(s2);
            const double alpha = L[i + i * LDL];
            const Householder h = make_householder(alpha, s2);
            for (int rr = i + lane; rr < n16; rr += 32) {
                const double x = L[rr + i * LDL];
                vvec[rr] = (rr == i) ? 1.0 : x * h.scal;
                L[rr + i * LDL] = (rr == i) ? h.beta : 0.0;
            }
            if (lane == 0) dmisc[0] = h.tau;
        }
        __syncthreads();
        {
            const double tau = dmisc[0];
            for (int cc = i + 1 + warp; cc < n16; cc += nwarps) {
                double x[N / 32], vv[N / 32];
                double dot = 0.0;
#pragma unroll
                for (int e = 0; e < N / 32; ++e) {
                    const int rr = lane + 32 * e;
                    const bool on = (rr >= i && rr < n16);
                    x[e] = on ? L[rr + cc * LDL] : 0.0;
                    vv[e] = on ? vvec[rr] : 0.0;
                    dot = fma(x[e], vv[e], dot);
                }
                dot = warp_sum(dot) * tau;
                if (tau != 0.0) {
#pragma unroll
                    for (int e = 0; e < N / 32; ++e) {
                        const int rr = lane + 32 * e;
                        if (rr >= i && rr < n16) {
                            x[e] = fma(-dot, vv[e], x[e]);
                            L[rr + cc * LDL] = x[e];
                        }
                    }
                }
                // down-date the squared norm with the (updated) entry in row i
                double xsel = 0.0;
#pragma unroll
                for (int e = 0; e < N / 32; ++e) xsel = (e == (i >> 5)) ? x[e] : xsel;
                const double top = __shfl_sync(0xffffffffu, xsel, i & 31);
                if (lane == 0) { const double d = vn[cc] - top * top; vn[cc] = d > 0.0 ? d : 0.0; }
            }
        }
        __syncthreads();
    }

    // ---- L <- R^T (R upper triangular, its strict lower part is already zero) ------------
    for (int idx = tid; idx < n16 * n16; idx += blockDim.x) {
        const int i = idx % n16, j = idx / n16;
        if (i < j) {
            const double x = L[i + j * LDL];
            L[j + i * LDL] = x;
            L[i + j * LDL] = 0.0;
        }
    }
    __syncthreads();

    // ---- one-sided Jacobi on the columns of L ---------------------------------------------
    // 8 lanes per column pair; lane l8 owns rows l8, l8+8, ...  The four pairs of a warp are
    // taken in the order k, k+2, k+1, k+3 so that the two 8-lane groups of a half-warp read
    // columns two apart (64-byte segments 16 banks apart: conflict-free with ld = N+4).
    const int l8 = tid & 7;
    const int grp = ((tid >> 5) << 2) | ((tid >> 4) & 1) | (((tid >> 3) & 1) << 1);
    const int npairs = n16 >> 1;
    int sweeps = 0;
    bool converged = false;
    for (int sweep = 0; sweep < kMaxSweeps && !converged; ++sweep) {
        int big = 0;
        for (int step = 0; step < n16 - 1; ++step) {
            for (int k = grp; k < npairs; k += (int)(blockDim.x >> 3)) {
                int p, qq;
                rr_pair(n16, step, k, p, qq);
                double xp[N / 8], xq[N / 8];
                double app = 0.0, aqq = 0.0, apq = 0.0;
#pragma unroll
                for (int e = 0; e < N / 8; ++e) {
                    const int rr = l8 + 8 * e;
                    const bool on = rr < n16;
                    xp[e] = on ? L[rr + p * LDL] : 0.0;
                    xq[e] = on ? L[rr + qq * LDL] : 0.0;
                    app = fma(xp[e], xp[e], app);
                    aqq = fma(xq[e], xq[e], aqq);
                    apq = fma(xp[e], xq[e], apq);
                }
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) {
                    app += __shfl_xor_sync(0xffffffffu, app, o);
                    aqq += __shfl_xor_sync(0xffffffffu, aqq, o);
                    apq += __shfl_xor_sync(0xffffffffu, apq, o);
                }
                // rotate when |cos| = |apq| / sqrt(app*aqq) > kTolRot (compared squared, no sqrt)
                const double den2 = app * aqq, apq2 = apq * apq;
                if (den2 > 0.0 && apq2 > (kTolRot * kTolRot) * den2) {
                    if (apq2 > (kTolStop * kTolStop) * den2) big = 1;
                    // tan(theta) = sign(d) * 2apq / (|d| + sqrt(d^2 + 4 apq^2)),  d = aqq - app
                    const double dd = aqq - app, h2 = 2.0 * apq;
                    const double s2 = fma(dd, dd, h2 * h2);
                    const double hyp = s2 * rsqrt(s2);
                    const double tt = (dd >= 0.0 ? h2 : -h2) * __drcp_rn(fabs(dd) + hyp);
                    const double cs = rsqrt(fma(tt, tt, 1.0));
                    const double sn = cs * tt;
#pragma unroll
                    for (int e = 0; e < N / 8; ++e) {
                        const int rr = l8 + 8 * e;
                        if (rr < n16) {
                            L[rr + p * LDL] = cs * xp[e] - sn * xq[e];
                            L[rr + qq * LDL] = sn * xp[e] + cs * xq[e];
                        }
                    }
                }
            }
            __syncthreads();
        }
        ++sweeps;
        converged = (__syncthreads_or(big) == 0);
    }

    // ---- singular values, sort, truncation -------------------------------------------------
    for (int cc = warp; cc < n16; cc += nwarps) {
        double s2 = 0.0;
        for (int rr = lane; rr < n16; rr += 32) { const double x = L[rr + cc * LDL]; s2 = fma(x, x, s2); }
        s2 = warp_sum(s2);
        if (lane == 0) sig[cc] = sqrt(s2);
    }
    if (tid == 0) misc[1] = n;
    __syncthreads();
    if (tid < n16) {
        const double mine = sig[tid];
        int pos = 0;
        for (int jj = 0; jj < n16; ++jj) {
            const double o = sig[jj];
            pos += (o > mine || (o == mine && jj < tid)) ? 1 : 0;
        }
        posof[tid] = pos;
        ssort[pos] = mine;
    }
    __syncthreads();
    // reference src/compute/hcore_dgemm.c:444-464: fixed rank if rk != 0, else the first
    // index i >= 1 with sigma_i < acc (absolute); sigma has n = a + c entries.
    if (ar.rk == 0) {
        if (tid >= 1 && tid < n && ssort[tid] < ar.acc) atomicMin(&misc[1], tid);
    } else if (tid == 0) {
        misc[1] = ar.rk < n ? ar.rk : n;
    }
    __syncthreads();
    const int rnew = misc[1];
    const int r8 = round_up(rnew, 8);
    if (tid < n16) ws.sigma[(size_t)t * N + tid] = ssort[tid];
    int st = converged ? HCORE_B200_OK : HCORE_B200_ERR_NO_CONVERGE;
    if (rnew > ar.maxrk) st = HCORE_B200_ERR_NEWRANK;
    if (tid == 0) {
        ws.newrank[t] = rnew;
        ws.sweeps[t] = sweeps;
        ws.st[t] = st;
        if (ar.status) ar.status[gt] = st;
        if (st == HCORE_B200_OK) ar.Crk[gt] = rnew;
    }
    if (st != HCORE_B200_OK) return;

    // ---- U_r = P * (normalised, sorted columns) -> Yu (n16 x r8, zero padded) -----------------
    for (int idx = tid; idx < n16 * n16; idx += blockDim.x) {
        const int i = idx % n16, cc = idx / n16;
        const int pos = posof[cc];
        if (pos < r8) {
            const double sg = sig[cc];
            const double val = (pos < rnew && sg > 0.0) ? L[i + cc * LDL] / sg : 0.0;
            Yu[perm[i] + (size_t)pos * N] = val;
        }
    }
    __syncthreads();   // Yu (global) visible to the whole CTA; L is free from here on

    // ---- W = R_v * (R_u^T * U_r) -> Yv ----------------------------------------------------------
    block_gemm<2, kGemmSet>(mview(L, 1, LDL), cview(Ru, N, 1), cview(Yu, 1, N), n16, r8, n16, warp, nwarps,
                            UpperTransK{n16});
    __syncthreads();
    block_gemm<2, kGemmSet>(mview(Yv, 1, N), cview(Rv, 1, N), cview(L, 1, LDL), n16, r8, n16, warp, nwarps,
                            UpperK());
}

// ========================================================================================
// 4. reflector application kernel
// ========================================================================================
constexpr int kApplyCols = 32;   // output columns per CTA

template <int N>
struct ApplySmem {
    using G = Geometry<N>;
    static constexpr int kLdY = N + 4;
    static constexpr int kY = 0;
    static constexpr int kZ = kY + kApplyCols * kLdY;
    static constexpr int kVs = kZ + kApplyCols * G::kLdB;
    static constexpr int kTs = kVs + kPanelW * G::kLdB;
    static constexpr int kS = kTs + 256;
    static constexpr int kS2 = kS + kApplyCols * G::kLdS;
    static constexpr int kTotal = kS2 + kApplyCols * G::kLdS;
    static constexpr size_t kBytes = (size_t)kTotal * sizeof(double);
};

// grid = (N / 32 column chunks, 2 panels, tiles); 256 threads.
template <int N>
__global__ void __launch_bounds__(256) tlr_apply_kernel(TlrArgs ar, TlrWorkspace ws) {
    using G = Geometry<N>;
    using SM = ApplySmem<N>;
    constexpr int LDB = G::kLdB, LDS = G::kLdS, LDY = SM::kLdY, NC = kApplyCols;
    extern __shared__ double smem[];
    double *Y = smem + SM::kY, *Z = smem + SM::kZ, *Vs = smem + SM::kVs, *Ts = smem + SM::kTs,
           *S = smem + SM::kS, *S2 = smem + SM::kS2;

    const int t = blockIdx.z, pnl = blockIdx.y, col0 = blockIdx.x * NC;
    if (ws.st[t] != HCORE_B200_OK) return;
    const int rnew = ws.newrank[t];
    if (col0 >= rnew) return;
    const int gt = ar.tile0 + t;
    const int tid = threadIdx.x, warp = warp_id();
    const int nwarps = blockDim.x >> 5;
    const int n16 = round_up(ws.ncols[t], kPanelW);
    const int nsp = n16 / kPanelW;
    const int m = ar.nb, m_pad = ws.m_pad;
    const int ncol = (rnew - col0 < NC) ? (rnew - col0) : NC;   // valid output columns
    const double *panel = ws.panel + panel_offset(ws, pnl, t, G::kPanelCols);
    const double *Tg = ws.T + t_offset(ws, pnl, t, G::kNsp);
    const double *Yg = ws.Y + nn_offset(ws, pnl, t, N) + (size_t)col0 * N;
    double *out = (pnl == 0 ? ar.CU[gt] : ar.CV[gt]) + (size_t)col0 * ar.ldc;

    // Y <- rows 0..n16-1 of this column chunk (columns beyond the rank are zero)
    for (int idx = tid; idx < n16 * NC; idx += blockDim.x) {
        const int i = idx % n16, jj = idx / n16;
        Y[i + jj * LDY] = (jj < ncol) ? Yg[i + (size_t)jj * N] : 0.0;
    }
    __syncthreads();

    for (int j = ws.nblk - 1; j >= 1; --j) {
        for (int idx = tid; idx < kBlockH * NC; idx += blockDim.x) Z[(idx & 127) + (idx >> 7) * LDB] = 0.0;
        for (int s = nsp - 1; s >= 0; --s) {
            const int c0 = s * kPanelW;
            for (int idx = tid; idx < kBlockH * kPanelW; idx += blockDim.x) {
                const int rr = idx & 127, cc = idx >> 7;
                Vs[rr + cc * LDB] = panel[(size_t)j * kBlockH + rr + (size_t)(c0 + cc) * m_pad];
            }
            Ts[tid] = Tg[((size_t)j * G::kNsp + s) * 256 + tid];
            for (int idx = tid; idx < kPanelW * NC; idx += blockDim.x) {
                const int ii = idx & 15, jj = idx >> 4;
                S[ii + jj * LDS] = Y[c0 + ii + jj * LDY];
            }
            __syncthreads();
            // S += Vs^T * Z
            block_gemm<1, kGemmAdd>(mview(S, 1, LDS), cview(Vs, LDB, 1), cview(Z, 1, LDB), kPanelW, NC, kBlockH,
                                    warp, nwarps);
            __syncthreads();
            // S2 = T * S
            block_gemm<1, kGemmSet>(mview(S2, 1, LDS), cview(Ts, 1, 16), cview(S, 1, LDS), kPanelW, NC, kPanelW,
                                    warp, nwarps);
            __syncthreads();
            for (int idx = tid; idx < kPanelW * NC; idx += blockDim.x) {
                const int ii = idx & 15, jj = idx >> 4;
                Y[c0 + ii + jj * LDY] -= S2[ii + jj * LDS];
            }
            block_gemm<2, kGemmSub>(mview(Z, 1, LDB), cview(Vs, 1, LDB), cview(S2, 1, LDS), kBlockH, NC, kPanelW,
                                    warp, nwarps);
            __syncthreads();
        }
        for (int idx = tid; idx < kBlockH * ncol; idx += blockDim.x) {
            const int rr = idx & 127, jj = idx >> 7;
            const int grow = j * kBlockH + rr;
            if (grow < m) out[grow + (size_t)jj * ar.ldc] = Z[rr + jj * LDB];
        }
        __syncthreads();
    }

    // ---- block 0: plain Householder block, reflectors below the diagonal -------------------
    for (int idx = tid; idx < kBlockH * NC; idx += blockDim.x) {
        const int rr = idx & 127, jj = idx >> 7;
        Z[rr + jj * LDB] = (rr < n16) ? Y[rr + jj * LDY] : 0.0;
    }
    for (int s = nsp - 1; s >= 0; --s) {
        const int c0 = s * kPanelW;
        for (int idx = tid; idx < kBlockH * kPanelW; idx += blockDim.x) {
            const int rr = idx & 127, cc = idx >> 7, col = c0 + cc;
            double x = 0.0;
            if (rr > col) x = panel[rr + (size_t)col * m_pad];
            else if (rr == col) x = 1.0;
            Vs[rr + cc * LDB] = x;
        }
        Ts[tid] = Tg[((size_t)s) * 256 + tid];
        __syncthreads();
        block_gemm<1, kGemmSet>(mview(S, 1, LDS), cview(Vs, LDB, 1), cview(Z, 1, LDB), kPanelW, NC, kBlockH, warp,
                                nwarps);
        __syncthreads();
        block_gemm<1, kGemmSet>(mview(S2, 1, LDS), cview(Ts, 1, 16), cview(S, 1, LDS), kPanelW, NC, kPanelW, warp,
                                nwarps);
        __syncthreads();
        block_gemm<2, kGemmSub>(mview(Z, 1, LDB), cview(Vs, 1, LDB), cview(S2, 1, LDS), kBlockH, NC, kPanelW, warp,
                                nwarps);
        __syncthreads();
    }
    for (int idx = tid; idx < kBlockH * ncol; idx += blockDim.x) {
        const int rr = idx & 127, jj = idx >> 7;
        if (rr < m) out[rr + (size_t)jj * ar.ldc] = Z[rr + jj * LDB];
    }
}

}  // namespace hcore_b200
