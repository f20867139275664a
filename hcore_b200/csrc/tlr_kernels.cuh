// Core SVD kernel of the batched tile-low-rank update (FP64, sm_100a).
//
//   tlr_svd_kernel     T' = R_v R_u^T, truncated QR with column pivoting as preconditioner
//                      (register resident, one barrier per Householder step: qrcp_registers),
//                      one-sided BLOCK Jacobi on the columns of R^T (4-column blocks, rotations
//                      derived from the Gram matrix of a block pair and applied with DMMAs:
//                      jacobi_pair_step), sort, rank truncation, U_r and W = R_v (R_u^T U_r)
//                      (replaces the core product, LAPACKE_dgesvd, the truncation loop and the
//                      sigma scaling, reference src/compute/hcore_dgemm.c:330-489)
//
// The other pipeline stages (P/G products, tall-skinny QR, reflector application) are in
// wy_kernels.cuh.  tests/device_model.py restates this arithmetic in numpy.
#pragma once
#include "common.cuh"
#include "wy_kernels.cuh"

namespace hcore_b200 {

#ifndef HCORE_SVD_NT
#define HCORE_SVD_NT 4
#endif
#ifndef HCORE_SVD_UNR
#define HCORE_SVD_UNR 4
#endif

template <int N>
struct Geometry {
    static constexpr int kLdL = N + 4;            // leading dim of the N x N core matrix
};

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

// N <= 128: the core matrix lives in shared memory.  N = 256 (ranks up to 128): 256 x 260 doubles do not fit
// one SM, the same code then works on a per-tile slab of the workspace that stays L2 resident.
template <int N>
struct SvdSmem {
    using G = Geometry<N>;
    static constexpr bool kGlobalL = (N > 128);
    static constexpr int kL = 0;
    static constexpr int kVn = kL + (kGlobalL ? 0 : N * G::kLdL);
    static constexpr int kVvec = kVn + N;
    static constexpr int kSig = kVvec + N;
    static constexpr int kSsort = kSig + N;
    static constexpr int kInts = kSsort + N;          // perm[N], posof[N], misc[8] as ints
    static constexpr int kTotal = kInts + N + 8;      // 2N+16 ints fit in N+8 doubles
    static constexpr size_t kBytes = (size_t)kTotal * sizeof(double);
};

// ---- register-resident QR with column pivoting (stacks up to 128 columns) ---------------------------------------
// The whole core matrix lives in registers: warp w owns columns 8w .. 8w+7, lane (g, t) holds the row pairs
// 8e + 2t, 8e + 2t + 1 (e < N/8) of column 8w + g.  Columns never move: a pivot only marks its column as factored at
// position `pos`.  One CTA barrier per Householder step:
//   after the update of step i every warp determines its best remaining column (exact squared norms of rows > i
//   fall out of the update), and SPECULATIVELY publishes that column (raw), its leading entry and the squared norm
//   of the rest to a per-warp slot; after the barrier all warps pick the same winner among the per-warp candidates
//   (redundantly: two REDUX instructions on the bit pattern of the norm with the column index in its low bits),
//   derive the reflector scalars from the winner's slot and apply the reflector to their own columns.  Slots are
//   double buffered by step parity and live in the shared-memory area of the core matrix, which is free while the
//   matrix is in registers.
// The step body is instantiated per 8-row block E = i / 8, so the rows already factored cost nothing and every
// register index is static.  Sums over the four lanes of a column are one DMMA (B = ones), warp totals two.
template <int N>
struct QrcpScratch {                          // offsets (doubles) inside the L area
    static constexpr int NW = N / 8;
    static constexpr int kY = 0;              // [2][NW][N] published columns
    static constexpr int kV = kY + 2 * NW * N;   // [2][NW] candidate key: squared norm of rows >= i + 1, column in the low bits
    static constexpr int kS = kV + 2 * NW;    // [2][NW] sum of the squared norms of the warp's remaining columns
    static constexpr int kA = kS + 2 * NW;    // [2][NW] leading entry (row i + 1) of the published column
    static constexpr int kT = kA + 2 * NW;    // [2][NW] squared norm of its rows > i + 1
    static_assert(kT + 2 * NW <= N * (N + 4), "QRCP scratch must fit the core-matrix area");
};

__device__ __forceinline__ double quad_sum_dmma(double v) {      // sum over the four lanes of a column group
    double d0 = 0.0, d1 = 0.0;
    dmma_884(d0, d1, v, 1.0);
    return d0;
}
__device__ __forceinline__ double warp_sum_dmma(double v) {      // sum over the 32 lanes, in every lane
    double d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0;
    dmma_884(d0, d1, 1.0, v);
    dmma_884(e0, e1, 1.0, d0 + d1);
    return e0;
}
// Warp-wide maximum of non-negative doubles by bit pattern (the order of the patterns is the order of the values),
// or 0 when every lane passes 0.  Two REDUX instructions.
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long k) {
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return ((unsigned long long)mh << 32) | ml;
}
// candidate key of a remaining column: its squared norm with the low 7 bits replaced by 127 - column (ties and
// near-ties go to the smaller column); 0 = no candidate
__device__ __forceinline__ unsigned long long qrcp_key(double nrm2, int col) {
    return ((unsigned long long)__double_as_longlong(nrm2) & ~0x7full) | (unsigned long long)(127 - col);
}

// candidates of the step that follows row `inext - 1`: norms of rows >= inext of every remaining column, warp argmax,
// publication.  inext / 8 is E or E + 1; both are handled by masks on the two row blocks.
template <int N, int E>
__device__ __forceinline__ void qrcp_publish(const double (&x)[N / 4], int inext, bool active, int col, int par,
                                             double *scr, int warp, int lane) {
    using S = QrcpScratch<N>;
    constexpr int RB = N / 8, NW = N / 8;
    constexpr unsigned kFull = 0xffffffffu;
    const int t = lane & 3;
    const int en = inext >> 3, rn = inext & 7;            // row inext: row block en, lane rn / 2, element rn & 1
    // leading entry and squared norm of the rows below it
    double xl = (rn & 1) ? x[2 * E + 1] : x[2 * E];
    if constexpr (E + 1 < RB) xl = (en == E) ? xl : x[2 * E + 2];
    const double alpha = __shfl_sync(kFull, xl, (lane & ~3) | (rn >> 1));
    double p[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int e = E; e < RB; ++e) {
        double v0 = x[2 * e], v1 = x[2 * e + 1];
        if (e <= E + 1) {
            v0 = (8 * e + 2 * t > inext) ? v0 : 0.0;
            v1 = (8 * e + 2 * t + 1 > inext) ? v1 : 0.0;
        }
        p[(2 * e) & 3] = fma(v0, v0, p[(2 * e) & 3]);
        p[(2 * e + 1) & 3] = fma(v1, v1, p[(2 * e + 1) & 3]);
    }
    const double tail2 = quad_sum_dmma((p[0] + p[1]) + (p[2] + p[3]));
    const double nrm2 = fma(alpha, alpha, tail2);
    const unsigned long long key = active ? qrcp_key(nrm2, col) : 0ull;
    const unsigned long long best = warp_max_u64(key);
    const double wsum = 0.25 * warp_sum_dmma(active ? nrm2 : 0.0);      // every column is held by four lanes
    const int slot = par * NW + warp;
    if (active && key == best) {
        double *ys = scr + S::kY + (size_t)slot * N + 2 * t;
#pragma unroll
        for (int e = E; e < RB; ++e) *reinterpret_cast<double2 *>(ys + 8 * e) = make_double2(x[2 * e], x[2 * e + 1]);
        if (t == 0) { scr[S::kA + slot] = alpha; scr[S::kT + slot] = tail2; }
    }
    if (lane == 0) {
        reinterpret_cast<unsigned long long *>(scr + S::kV)[slot] = best;
        scr[S::kS + slot] = wsum;
    }
}

// Householder step i (row block E = i / 8).  Returns true when the factorisation stops BEFORE this step (the
// Frobenius norm of the remaining block is below the tolerance): kq = i.
template <int N, int E>
__device__ __forceinline__ bool qrcp_step(double (&x)[N / 4], int &pos, int i, int n16, int col, bool in_range,
                                          double &tol2, double *scr, double *betas, int warp, int lane) {
    using S = QrcpScratch<N>;
    constexpr int RB = N / 8, NW = N / 8;
    constexpr unsigned kFull = 0xffffffffu;
    const int t = lane & 3, ri = i & 7, par = i & 1;
    PROF_DECL(threadIdx.x == 0)
    __syncthreads();
    PROF_ADD(40);
    // ---- winner among the per-warp candidates (every warp, redundantly) ---------------------------------------
    const unsigned long long ck = (lane < NW) ? reinterpret_cast<const unsigned long long *>(scr + S::kV)[par * NW + lane] : 0ull;
    const double fs = warp_sum_dmma((lane < NW) ? scr[S::kS + par * NW + lane] : 0.0);
    const unsigned long long best = warp_max_u64(ck);
    if (i == 0 && tol2 < 0.0) tol2 = 1e-30 * __longlong_as_double((long long)best);      // fixed-rank mode
    if (fs < tol2 || best == 0ull) return true;
    const int piv = 127 - (int)(best & 0x7full);
    const int wslot = par * NW + (piv >> 3);
    const double *ys = scr + S::kY + (size_t)wslot * N + 2 * t;
    PROF_ADD(41);
    const Householder h = make_householder_inline(scr[S::kA + wslot], scr[S::kT + wslot]);
    if (col == piv) {
        pos = i;
        if (t == 0) betas[col] = h.beta;
    }
    // ---- apply H = I - tau v v^T, v = [1; scal y], to my column: d = y^T x (rows > i) ----------------------------
    double p[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int e = E; e < RB; ++e) {
        double2 y = *reinterpret_cast<const double2 *>(ys + 8 * e);
        if (e == E) { y.x = (2 * t > ri) ? y.x : 0.0; y.y = (2 * t + 1 > ri) ? y.y : 0.0; }
        p[(2 * e) & 3] = fma(x[2 * e], y.x, p[(2 * e) & 3]);
        p[(2 * e + 1) & 3] = fma(x[2 * e + 1], y.y, p[(2 * e + 1) & 3]);
        if (((e - E) & 3) == 3) asm volatile("" ::: "memory");      // at most four loads in flight: registers
    }
    const double d = quad_sum_dmma((p[0] + p[1]) + (p[2] + p[3]));
    PROF_ADD(42);
    const double xi = __shfl_sync(kFull, (ri & 1) ? x[2 * E + 1] : x[2 * E], (lane & ~3) | (ri >> 1));
    const double wv = h.tau * fma(h.scal, d, xi), f = h.scal * wv;
    // second pass over the published column: the address is made to depend on d so that these loads are not hoisted
    // above the reduction (sixteen 16-byte values in flight on top of the register-resident column would spill)
    const double *ys2 = ys + (__double2hiint(d) & 0);
#pragma unroll
    for (int e = E; e < RB; ++e) {
        const double2 y = *reinterpret_cast<const double2 *>(ys2 + 8 * e);
        if (e == E) {
            const int r0 = 2 * t, r1 = 2 * t + 1;
            x[2 * e] = (r0 > ri) ? fma(-y.x, f, x[2 * e]) : ((r0 == ri) ? x[2 * e] - wv : x[2 * e]);
            x[2 * e + 1] = (r1 > ri) ? fma(-y.y, f, x[2 * e + 1]) : ((r1 == ri) ? x[2 * e + 1] - wv : x[2 * e + 1]);
        } else {
            x[2 * e] = fma(-y.x, f, x[2 * e]);
            x[2 * e + 1] = fma(-y.y, f, x[2 * e + 1]);
        }
        if (((e - E) & 3) == 3) asm volatile("" ::: "memory");
    }
    PROF_ADD(43);
    // ---- candidates of the next step ---------------------------------------------------------------------------------
    if (i + 1 < n16) qrcp_publish<N, E>(x, i + 1, in_range && pos < 0, col, par ^ 1, scr, warp, lane);
    PROF_ADD(44);
    PROF_COUNT(45, 1);
    return false;
}

template <int N, int E>
__device__ __forceinline__ int qrcp_blocks(double (&x)[N / 4], int &pos, int n16, int col, bool in_range, double &tol2,
                                           double *scr, double *betas, int warp, int lane) {
    if constexpr (E < N / 8) {
#pragma unroll 1
        for (int r = 0; r < 8; ++r) {
            const int i = 8 * E + r;
            if (i >= n16) return n16;
            if (qrcp_step<N, E>(x, pos, i, n16, col, in_range, tol2, scr, betas, warp, lane)) return i;
        }
        return qrcp_blocks<N, E + 1>(x, pos, n16, col, in_range, tol2, scr, betas, warp, lane);
    } else {
        return n16;
    }
}

// Truncated QRCP of the n16 x n16 matrix in L (column-major, ld N + 4); on return L holds R^T restricted to the kq
// factored rows (L(a, b) = R(b, a): row a = pivot position of a column, column b = Householder step; zero for b >= kq
// and above the diagonal), perm[a] = original column at position a.  Returns kq.  betas[N], cnt[N / 8]: scratch.
template <int N>
__device__ __forceinline__ int qrcp_registers(double *L, int n16, int rk, double acc, int *perm, int *cnt, double *betas) {
    constexpr int LDL = N + 4, RB = N / 8;
    constexpr unsigned kFull = 0xffffffffu;
    static_assert(N <= 128, "the column index travels in seven bits of the candidate key");
    const int warp = warp_id(), lane = lane_id();
    const int g = lane >> 2, t = lane & 3;
    const int col = 8 * warp + g;
    const bool in_range = col < n16;
    double x[2 * RB];
#pragma unroll
    for (int e = 0; e < RB; ++e) {
        double2 v = make_double2(0.0, 0.0);
        if (in_range && 8 * e < n16) v = *reinterpret_cast<const double2 *>(L + (size_t)col * LDL + 8 * e + 2 * t);
        x[2 * e] = v.x; x[2 * e + 1] = v.y;
    }
    __syncthreads();                          // the matrix is in registers: its shared-memory area becomes scratch
    int pos = -1;
    // stop when the squared Frobenius norm of the remaining block is below tol2: 1 % of the truncation threshold, or
    // (fixed-rank mode, set in step 0) rounding level of the largest column
    double tol2 = (rk == 0) ? 1e-4 * acc * acc : -1.0;
    qrcp_publish<N, 0>(x, 0, in_range, col, 0, L, warp, lane);
    const int kq = qrcp_blocks<N, 0>(x, pos, n16, col, in_range, tol2, L, betas, warp, lane);
    // ---- positions of the columns that were never selected: kq + rank among them (column order) -------------------
    const unsigned rest = __ballot_sync(kFull, in_range && pos < 0 && t == 0);
    if (lane == 0) cnt[warp] = __popc(rest);
    __syncthreads();                          // also: every warp is done with the scratch, betas[] is complete
    if (in_range && pos < 0) {
        int a = kq;
        for (int w = 0; w < warp; ++w) a += cnt[w];
        pos = a + __popc(rest & ((1u << (lane & ~3)) - 1u));      // remaining columns of my warp left of mine
    }
    if (in_range) {
        const bool sel = pos < kq;
        const double beta = sel ? betas[col] : 0.0;
#pragma unroll
        for (int e = 0; e < 2 * RB; ++e) {
            const int b = 8 * (e >> 1) + 2 * t + (e & 1);
            if (b < n16) {
                double v = 0.0;
                if (b < kq) v = sel ? (b < pos ? x[e] : (b == pos ? beta : 0.0)) : x[e];
                L[pos + (size_t)b * LDL] = v;
            }
        }
        if (t == 0) perm[pos] = col;
    }
    __syncthreads();
    return kq;
}

// hardware seed of 1 / sqrt(w) (MUFU.RSQ64H: relative error ~2^-22), no special cases: callers select the result
// away when w is 0 or not finite
__device__ __forceinline__ double rsqrt_seed(double w) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(w));
    return y;
}

// Two-warp named barrier k + 1 (k = 0 .. 7) with a literal id (ptxas reserves highest id + 1 barriers per CTA):
// ARRIVE = true signals without waiting, false waits for the partner's signal.
template <bool ARRIVE>
__device__ __forceinline__ void pair_barrier(int k) {
#define HCORE_PAIR_CASE(i) case i: if (ARRIVE) asm volatile("bar.arrive " #i " + 1, 64;" ::: "memory"); else asm volatile("bar.sync " #i " + 1, 64;" ::: "memory"); break;
    switch (k) {
        HCORE_PAIR_CASE(0) HCORE_PAIR_CASE(1) HCORE_PAIR_CASE(2) HCORE_PAIR_CASE(3)
        HCORE_PAIR_CASE(4) HCORE_PAIR_CASE(5) HCORE_PAIR_CASE(6) HCORE_PAIR_CASE(7)
        default: break;
    }
#undef HCORE_PAIR_CASE
}

// X <- X J for the eight columns of blocks p, q (in place): b0 / b1 = B operands J(t, n(g)), J(4 + t, n(g)) with
// n(g) = 4 (g & 1) + g / 2, so that accumulator column 2t + s is solver column 4 s + t and the results land exactly
// where the A operands came from.  4 DMMAs per 16 rows.
template <int N>
__device__ __forceinline__ void jacobi_pair_apply(double *L, int n16, int p, int q, double b0, double b1, int lane) {
    constexpr int LDL = N + 4, RB = N / 8;
    const int g = lane >> 2, t = lane & 3;
    const int mu = ((t & 1) << 1) | (t >> 1);
    double *cp = L + (size_t)(4 * p + mu) * LDL + 2 * g;
    double *cq = L + (size_t)(4 * q + mu) * LDL + 2 * g;
    if constexpr (N > 128) {               // core matrix in L2: eight loads in flight per group of 64 rows
#pragma unroll
        for (int e0 = 0; e0 < RB / 2; e0 += 4) {
            if (16 * e0 >= n16) break;
            double2 vp[4], vq[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = (16 * (e0 + u) < n16) ? e0 + u : e0;
                vp[u] = *reinterpret_cast<const double2 *>(cp + 16 * e);
                vq[u] = *reinterpret_cast<const double2 *>(cq + 16 * e);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (16 * (e0 + u) >= n16) break;
                double d0x = 0.0, d0y = 0.0, d1x = 0.0, d1y = 0.0;
                dmma_884(d0x, d0y, vp[u].x, b0);
                dmma_884(d0x, d0y, vq[u].x, b1);
                dmma_884(d1x, d1y, vp[u].y, b0);
                dmma_884(d1x, d1y, vq[u].y, b1);
                *reinterpret_cast<double2 *>(cp + 16 * (e0 + u)) = make_double2(d0x, d1x);
                *reinterpret_cast<double2 *>(cq + 16 * (e0 + u)) = make_double2(d0y, d1y);
            }
        }
    } else {
#pragma unroll
    for (int e = 0; e < RB / 2; ++e) {
        if (16 * e >= n16) break;          // a real (warp-uniform) exit: predicated-off DMMAs still occupy the FP64 pipe
        const double2 vp = *reinterpret_cast<const double2 *>(cp + 16 * e);
        const double2 vq = *reinterpret_cast<const double2 *>(cq + 16 * e);
        double d0x = 0.0, d0y = 0.0, d1x = 0.0, d1y = 0.0;
        dmma_884(d0x, d0y, vp.x, b0);
        dmma_884(d0x, d0y, vq.x, b1);
        dmma_884(d1x, d1y, vp.y, b0);
        dmma_884(d1x, d1y, vq.y, b1);
        *reinterpret_cast<double2 *>(cp + 16 * e) = make_double2(d0x, d1x);
        *reinterpret_cast<double2 *>(cq + 16 * e) = make_double2(d0y, d1y);
    }
    }
}

// ---- one-sided BLOCK Jacobi step on a pair of 4-column blocks, everything GEMM-shaped on the FP64 tensor pipe ----
// One warp owns the pair (blocks p, q of the column-major core matrix L, n16 rows):
//   1. G = X^T X (8 x 8 Gram matrix of the pair's eight columns): 2 DMMAs per 8 rows, the same register is A and B
//      operand (lane (g, t) holds rows 2t, 2t+1 of column g of the 8-row slice);
//   2. the cross rotations of the pair (4 rounds of 4 disjoint rotations; in the first step of a sweep also the
//      3 rounds of in-block rotations) are derived from G ALONE: each rotation is computed redundantly by the lanes
//      of its two rows; G and J stay in accumulator-fragment layout (an accumulator fragment is the A operand of the
//      next product): G J_r and J J_r are two DMMAs each, the row mix J_r^T (G J_r) one shuffle per entry (the lane of
//      the partner row), so a round needs five shuffles;
//   3. X <- X J: 4 DMMAs per 16 rows, in place.
// Compared with rotating the columns one pair at a time this trades the shuffle reductions, the scalar updates and
// three quarters of the barriers for DMMAs.  Column index inside the solver: m = 0..3 -> column mu4(m) of block p,
// m = 4..7 -> column mu4(m - 4) of block q, mu4 = (0, 2, 1, 3): this relabelling makes both shared-memory access
// patterns (16-byte loads) bank-conflict free with the leading dimension N + 4.
//   Accuracy: G is formed from the current columns in every step (entry-wise error eps ||x_i|| ||x_j||, i.e. the
// cosines are accurate to rounding even between columns of very different norms); the two-sided updates keep that
// scaled accuracy unless a rotation cancels a column almost completely, which only costs an inexact ANGLE in the
// remaining rounds of the step (J is a product of plane rotations whatever the angles: orthogonal to rounding).
template <int N>
__device__ __forceinline__ void jacobi_pair_step(double *L, int n16, int p, int q, bool inblock, int lane, int &big,
                                                 double *jb, int jbar) {
    constexpr int LDL = N + 4, RB = N / 8;
    constexpr unsigned kFull = 0xffffffffu;
    const int g = lane >> 2, t = lane & 3;
    // ---- 1. Gram matrix ----------------------------------------------------------------------------------------
    double gx, gy;                                     // G(g, 2t), G(g, 2t + 1)
    {
        const int mu = ((g & 1) << 1) | ((g >> 1) & 1);
        const double *gc = L + (size_t)(4 * ((g & 4) ? q : p) + mu) * LDL + 2 * t;
        double ax[4] = {0.0, 0.0, 0.0, 0.0}, ay[4] = {0.0, 0.0, 0.0, 0.0};
        if constexpr (N > 128) {
            // core matrix in L2: eight 16-byte loads in flight before the first use (a load behind every loop exit would
            // expose the L2 latency 32 times per Gram matrix); rows beyond n16 are read from a valid row and not used
#pragma unroll
            for (int j0 = 0; j0 < RB; j0 += 8) {
                if (8 * j0 >= n16) break;
                double2 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const double2 *>(gc + 8 * ((8 * (j0 + u) < n16) ? j0 + u : j0));
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (8 * (j0 + u) >= n16) break;
                    dmma_884(ax[u & 3], ay[u & 3], v[u].x, v[u].x);
                    dmma_884(ax[u & 3], ay[u & 3], v[u].y, v[u].y);
                }
            }
        } else {
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            if (8 * j >= n16) break;       // a real (warp-uniform) exit: predicated-off DMMAs still occupy the FP64 pipe
            const double2 v = *reinterpret_cast<const double2 *>(gc + 8 * j);
            dmma_884(ax[j & 3], ay[j & 3], v.x, v.x);
            dmma_884(ax[j & 3], ay[j & 3], v.y, v.y);
        }
        }
        gx = (ax[0] + ax[1]) + (ax[2] + ax[3]);
        gy = (ay[0] + ay[1]) + (ay[2] + ay[3]);
    }
    // ---- nothing to rotate?  In the last sweeps most pairs (in the final, checking sweep: all of them) have every
    // cosine below the rotation threshold: all rounds would produce the identity, G would stay as it is and X J = X
    // exactly, so the step ends here -- bit-identical to going through the motions, at a third of the cost.
    {
        const double dsel = (g & 1) ? gy : gx;         // lane (g, g / 2) holds G(g, g) here
        const double dg = __shfl_sync(kFull, dsel, 4 * g + (g >> 1));
        const double d0 = __shfl_sync(kFull, gx, 9 * t), d1 = __shfl_sync(kFull, gy, 9 * t + 4);   // G(2t, 2t), G(2t+1, 2t+1)
        const bool rel0 = inblock ? (g != 2 * t) : ((g < 4) != (t < 2));       // entries some round looks at
        const bool rel1 = inblock ? (g != 2 * t + 1) : ((g < 4) != (t < 2));
        const bool any = (rel0 && gx * gx > (kTolRot * kTolRot) * (dg * d0)) || (rel1 && gy * gy > (kTolRot * kTolRot) * (dg * d1));
        if (!__any_sync(kFull, any)) {
            if (jb) {
                if (lane == 0) jb[64] = 0.0;
                pair_barrier<true>(jbar);
            }
            return;
        }
    }
    // ---- 2. rotations ----------------------------------------------------------------------------------------
    double jx = 0.0, jy = 0.0;                         // accumulated J(g, 2t), J(g, 2t + 1)
    bool first = true;
    auto round = [&](const int pg, const bool last) {  // pg: partner of row / column g in this round
        const double dsel = (g & 1) ? gy : gx;         // lane (g, g / 2) holds G(g, g) here
        const double dg = __shfl_sync(kFull, dsel, 4 * g + (g >> 1));
        const double dp = __shfl_sync(kFull, dsel, 4 * pg + (pg >> 1));
        const double osel = (pg & 1) ? gy : gx;        // lane (g, pg / 2) holds G(g, pg) here
        const double apq = __shfl_sync(kFull, osel, 4 * g + (pg >> 1));
        const bool lo = g < pg;
        const double app = lo ? dg : dp, aqq = lo ? dp : dg;
        const double den2 = app * aqq, apq2 = apq * apq;
        // Branch-free (the lanes of a warp hold four different rotations).  cos^2 = (1 + |d| / hyp) / 2 with
        // d = aqq - app, hyp = sqrt(d^2 + 4 apq^2); sin = sign(d) apq / (hyp cos).  The two reciprocal square roots
        // are the bare hardware seeds (~2^-22): the ANGLE only has to be good to a few digits (a rotation by a slightly
        // wrong angle is still a rotation and the off-diagonal entry still shrinks by 2^-22), what must hold to rounding
        // is cos^2 + sin^2 = 1, restored by the second-order correction 1 - delta/2 + 3 delta^2/8.  This keeps the
        // dependent FP64 chain of a round short: it shares the FP64 pipe with the DMMAs of the other warps.
        const bool rot = den2 > 0.0 && apq2 > (kTolRot * kTolRot) * den2;
        if (rot && apq2 > (kTolStop * kTolStop) * den2) big = 1;
        const double dd = aqq - app;
        const double rh = rsqrt_seed(fma(dd, dd, 4.0 * apq2));
        const double u = fma(0.5 * fabs(dd), rh, 0.5);
        const double ru = rsqrt_seed(u);
        double cs = u * ru;
        double sn = ((dd >= 0.0 ? apq : -apq) * rh) * ru;
        const double dl = fma(cs, cs, fma(sn, sn, -1.0));
        const double rn = fma(dl, fma(0.375, dl, -0.5), 1.0);
        cs = rot ? cs * rn : 1.0;
        sn = rot ? sn * rn : 0.0;
        // x_lo' = cs x_lo - sn x_hi, x_hi' = sn x_lo + cs x_hi: column g of J_r is cs at row g and sg at row pg
        const double sg = lo ? -sn : sn;
        const double j0 = (2 * t == g) ? cs : ((2 * t == pg) ? sg : 0.0);          // J_r(2t, g)
        const double j1 = (2 * t + 1 == g) ? cs : ((2 * t + 1 == pg) ? sg : 0.0);  // J_r(2t + 1, g)
        if (!last) {
            // G' = J_r^T (G J_r): the column mix M = G J_r by two DMMAs (G as A operand), the row mix
            // G'(i, j) = cs_i M(i, j) + sg_i M(P(i), j) with this lane's own rotation and the partner row's lane (pg, t)
            double mx = 0.0, my = 0.0;
            dmma_884(mx, my, gx, j0);
            dmma_884(mx, my, gy, j1);
            const double px = __shfl_sync(kFull, mx, 4 * pg + t), py = __shfl_sync(kFull, my, 4 * pg + t);
            gx = fma(sg, px, cs * mx);
            gy = fma(sg, py, cs * my);
        }
        if (first) {                                   // J = J_r: row g of J_r is cs at column g and -sg at column pg
            jx = (2 * t == g) ? cs : ((2 * t == pg) ? -sg : 0.0);
            jy = (2 * t + 1 == g) ? cs : ((2 * t + 1 == pg) ? -sg : 0.0);
            first = false;
        } else {
            double nx = 0.0, ny = 0.0;
            dmma_884(nx, ny, jx, j0);                  // J <- J J_r
            dmma_884(nx, ny, jy, j1);
            jx = nx; jy = ny;
        }
    };
    if (inblock) {
#pragma unroll
        for (int b = 1; b < 4; ++b) round(g ^ b, false);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) round((g < 4) ? 4 + ((g + r) & 3) : ((g - r) & 3), r == 3);
    // ---- 3. X <- X J ------------------------------------------------------------------------------------------
    // B operand of k-chunk h: J(4h + t, n(g)) with n(g) = 4 (g & 1) + g / 2, so that accumulator column 2t + s is
    // solver column 4 s + t: the results land exactly where the A operands came from.
    double bj[2];
    {
        const int csrc = 2 * (g & 1) + (g >> 2);
        const bool odd = ((g >> 1) & 1) != 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int src = 4 * (4 * h + t) + csrc;
            const double vx = __shfl_sync(kFull, jx, src), vy = __shfl_sync(kFull, jy, src);
            bj[h] = odd ? vy : vx;
        }
    }
    if (jb) {                                          // a helper warp on a less loaded scheduler applies J (see the kernel)
        jb[lane] = bj[0];
        jb[32 + lane] = bj[1];
        if (lane == 0) jb[64] = 1.0;
        pair_barrier<true>(jbar);
        return;
    }
    jacobi_pair_apply<N>(L, n16, p, q, bj[0], bj[1], lane);
}

// One CTA per tile.  Threads = 4 N (one 16-lane group per Jacobi block pair), capped at 512: the narrow stacks
// (N = 32, 64: ranks up to 16 / 32) run four / two tiles per SM at a time instead of one mostly idle 512-thread
// CTA -- their SVD is pure latency (QRCP steps, Jacobi rounds), so co-resident tiles overlap it.
template <int N>
struct SvdThreads { static constexpr int value = (4 * N < 512) ? 4 * N : 512; };

template <int N>
__global__ void __launch_bounds__(SvdThreads<N>::value, 512 / SvdThreads<N>::value) tlr_svd_kernel(TlrArgs ar, TlrWorkspace ws) {
    using G = Geometry<N>;
    using SM = SvdSmem<N>;
    constexpr int LDL = G::kLdL;
    extern __shared__ __align__(16) double smem[];
    double *L = SM::kGlobalL ? ws.L + (size_t)blockIdx.x * N * LDL : smem + SM::kL;
    double *vn = smem + SM::kVn, *vvec = smem + SM::kVvec, *sig = smem + SM::kSig,
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

    PROF_DECL(tid == 0)
    PROF_COUNT(29, 1);
    // ---- T' = R_v * R_u^T --------------------------------------------------------------
    block_gemm<HCORE_SVD_NT, kGemmSet, HCORE_SVD_UNR>(mview(L, 1, LDL), cview(Rv, 1, N), cview(Ru, N, 1), n16, n16, n16, warp, nwarps,
                            UpperTimesUpperT());
    if (tid < n16) perm[tid] = tid;
    __syncthreads();
    PROF_ADD(10);

    int kq;
    if constexpr (N <= 128) {
        kq = qrcp_registers<N>(L, n16, ar.rk, ar.acc, perm, posof, vn);
        PROF_ADD(12);
        PROF_COUNT(20, kq);
    } else {
    // ---- QR with column pivoting (preconditioner), truncated -----------------------------------
    // Two barriers per step:
    //   phase C (warp 0)  pivot = argmax of the 16 per-warp candidates, swap, Householder vector
    //   phase D (all)     apply the reflector to the trailing columns (a warp owns columns
    //                     i+1+warp, +16, ...; batches of 4 columns share one DMMA all-reduce),
    //                     down-date their squared norms (recomputed exactly every 8th step) and
    //                     leave the warp's best candidate for the next step
    // The squared norms of the trailing columns are recomputed EXACTLY in every step (they fall out of the update
    // loop: one more DMMA all-reduce per batch of four columns, no down-dating and none of its cancellation).
    // The factorisation stops at the first step kq at which the Frobenius norm of the remaining block -- the sum of
    // those norms -- is below 1% of the truncation threshold (singular values move by less than that, Weyl), or
    // at rounding level of the largest column in fixed-rank mode.
    constexpr int RPL = N / 32;                        // rows per lane
    double *candv = sig;                               // [16] per-warp best squared norm (sig is free until later)
    double *cands = sig + 16;                          // [16] per-warp sum of the squared norms of its columns
    int *candi = reinterpret_cast<int *>(ssort);       // [16] and its column
    {
        // exact norms of all columns; warp w owns columns w, w+16, ... like phase D of "step -1"
        double best = -1.0, wsum = 0.0; int bi = 0;
        for (int cc = warp; cc < n16; cc += nwarps) {
            double s2 = 0.0;
            for (int rr = lane; rr < n16; rr += 32) { const double x = L[rr + cc * LDL]; s2 = fma(x, x, s2); }
            s2 = warp_sum(s2);
            wsum += s2;
            if (s2 > best) { best = s2; bi = cc; }
        }
        if (lane == 0) { candv[warp] = best; candi[warp] = bi; cands[warp] = wsum; }
    }
    __syncthreads();
    double tol2;
    {
        double mx = 0.0;
        for (int w = 0; w < nwarps; ++w) mx = fmax(mx, candv[w]);
        tol2 = (ar.rk == 0) ? 1e-4 * ar.acc * ar.acc : 1e-30 * mx;
    }
    kq = n16;
    PROF_ADD(11);
    for (int i = 0; i < n16; ++i) {
        if (warp == 0) {
            // -- pivot --
            double best = (lane < nwarps) ? candv[lane] : -1.0;
            int bi = (lane < nwarps) ? candi[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            best = __shfl_sync(0xffffffffu, best, 0);
            const int piv = __shfl_sync(0xffffffffu, bi, 0);
            double fro2 = (lane < nwarps) ? cands[lane] : 0.0;       // squared Frobenius norm of the remaining block
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) fro2 += __shfl_xor_sync(0xffffffffu, fro2, o);
            fro2 = __shfl_sync(0xffffffffu, fro2, 0);
            const bool stop = fro2 < tol2;
            if (lane == 0) misc[2] = stop ? 1 : 0;
            if (!stop) {
                // -- swap columns i <-> piv, build the reflector of the pivot column --
                double xi[RPL], xp[RPL];
#pragma unroll
                for (int e = 0; e < RPL; ++e) {
                    const int rr = lane + 32 * e;
                    xi[e] = (rr < n16) ? L[rr + i * LDL] : 0.0;
                    xp[e] = (rr < n16) ? L[rr + piv * LDL] : 0.0;
                }
                double rv[2] = {0.0, 0.0};             // squared norm below the diagonal, diagonal entry
#pragma unroll
                for (int e = 0; e < RPL; ++e) {
                    const int rr = lane + 32 * e;
                    if (rr > i) rv[0] = fma(xp[e], xp[e], rv[0]);
                    if (rr == i) rv[1] = xp[e];
                }
                warp_allsum_dmma<2>(rv);
                const Householder h = make_householder(rv[1], rv[0]);
#pragma unroll
                for (int e = 0; e < RPL; ++e) {
                    const int rr = lane + 32 * e;
                    if (rr < n16) {
                        vvec[rr] = (rr == i) ? 1.0 : (rr > i ? xp[e] * h.scal : 0.0);
                        L[rr + i * LDL] = (rr < i) ? xp[e] : (rr == i ? h.beta : 0.0);
                        if (piv != i) L[rr + piv * LDL] = xi[e];
                    }
                }
                if (lane == 0) {
                    dmisc[0] = h.tau;
                    if (piv != i) { const int pi = perm[i]; perm[i] = perm[piv]; perm[piv] = pi; }
                }
            }
        }
        __syncthreads();
        if (misc[2]) { kq = i; break; }
        {
            const double tau = dmisc[0];
            double vv[RPL];
#pragma unroll
            for (int e = 0; e < RPL; ++e) { const int rr = lane + 32 * e; vv[e] = (rr >= i && rr < n16) ? vvec[rr] : 0.0; }
            double best = -1.0, wsum = 0.0; int bi = 0x7fffffff;
            for (int c0 = i + 1 + warp; c0 < n16; c0 += 4 * nwarps) {
                double x[4][RPL], dot[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int cc = c0 + q * nwarps;
                    double d = 0.0;
#pragma unroll
                    for (int e = 0; e < RPL; ++e) {
                        const int rr = lane + 32 * e;
                        x[q][e] = (cc < n16 && rr < n16) ? L[rr + cc * LDL] : 0.0;
                        d = fma(x[q][e], vv[e], d);
                    }
                    dot[q] = d;
                }
                warp_allsum_dmma<4>(dot);
                double nrm[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int cc = c0 + q * nwarps;
                    const double w = dot[q] * tau;
                    double s2 = 0.0;
#pragma unroll
                    for (int e = 0; e < RPL; ++e) {
                        const int rr = lane + 32 * e;
                        x[q][e] = fma(-w, vv[e], x[q][e]);
                        if (cc < n16 && rr >= i && rr < n16) L[rr + cc * LDL] = x[q][e];
                        if (rr > i) s2 = fma(x[q][e], x[q][e], s2);
                    }
                    nrm[q] = s2;
                }
                warp_allsum_dmma<4>(nrm);                      // exact squared norms of rows > i, in every lane
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int cc = c0 + q * nwarps;
                    if (cc < n16) {
                        const double v = nrm[q];
                        wsum += v;
                        if (v > best || (v == best && cc < bi)) { best = v; bi = cc; }
                    }
                }
            }
            if (lane == 0) { candv[warp] = best; candi[warp] = bi; cands[warp] = wsum; }
        }
        __syncthreads();
    }

    PROF_ADD(12);
    PROF_COUNT(20, kq);
    // ---- L <- R^T restricted to the kq factored rows (the remaining block is dropped) -------------
    for (int idx = tid; idx < n16 * n16; idx += blockDim.x) {
        const int i = idx % n16, j = idx / n16;
        if (i >= kq) L[i + j * LDL] = 0.0;
    }
    __syncthreads();
    for (int idx = tid; idx < n16 * n16; idx += blockDim.x) {
        const int i = idx % n16, j = idx / n16;
        if (i < j) {
            const double x = L[i + j * LDL];
            L[j + i * LDL] = x;
            L[i + j * LDL] = 0.0;
        }
    }
    __syncthreads();
    PROF_ADD(13);
    }   // N > 128

    int sweeps = 0;
    bool converged;
    {
    // ---- one-sided block Jacobi on the columns of L: blocks of four columns, a warp per block pair, rotations
    //      derived from the pair's Gram matrix and applied with DMMAs (jacobi_pair_step).  Blocks meet in
    //      round-robin order: kq8/4 - 1 barrier steps per sweep, every column pair is rotated once per sweep.
    const int kq8 = (kq + 7) & ~7;                    // columns kq .. kq8-1 of L are zero
    const int nb4 = kq8 >> 2, npair = nb4 >> 1;       // blocks (even), block pairs per step (N = 256: up to two per warp)
    converged = (kq8 < 8);
    // Helpers (one 512-thread CTA per SM).  Warp w runs on scheduler w % 4, so with npair pairs the first npair % 4
    // schedulers carry one pair more than the others, and every step would wait for them while the idle warps
    // npair .. 15 sit on the lighter schedulers.  Idle warp npair + i therefore takes over the update phase X <- X J of
    // pair own(i) = (i % r) + 4 (i / r), r = npair % 4 (the pairs of the overloaded schedulers, in turn): the owner
    // hands J over through shared memory and a two-warp named barrier and is done after its rotations.
    constexpr bool kHelp = (N == 128);
    constexpr int kJbLd = 72, kMaxHelp = 5;            // per helper: 64 operand values + flag; vvec | sig | ssort hold 5 x 72
    int help_own = -1, help_idx = -1;                  // helper warp: the pair it helps / owner warp: its helper's index
    if constexpr (kHelp) {
        const int r = npair & 3, nidle = nwarps - npair;
        if (r != 0) {
            const int nh = nidle < kMaxHelp ? nidle : kMaxHelp;
            if (warp >= npair && warp - npair < nh) {
                const int i = warp - npair, o = (i % r) + 4 * (i / r);
                if (o < npair) { help_own = o; help_idx = i; }
            } else if (warp < npair && (warp & 3) < r) {
                const int i = (warp & 3) + r * (warp >> 2);      // inverse of own(i)
                if (i < nh) help_idx = i;
            }
        }
    }
    for (int sweep = 0; sweep < kMaxSweeps && !converged; ++sweep) {
        int big = 0;
        for (int step = 0; step < nb4 - 1; ++step) {
            if (warp < npair) {
                for (int k = warp; k < npair; k += nwarps) {       // one pair per warp unless N = 256 (core matrix in L2)
                    int p, qq;
                    rr_pair(nb4, step, k, p, qq);
                    jacobi_pair_step<N>(L, n16, p, qq, step == 0, lane, big, help_idx >= 0 ? vvec + kJbLd * help_idx : nullptr, help_idx);
                }
            } else if (help_own >= 0) {
                int p, qq;
                rr_pair(nb4, step, help_own, p, qq);
                const double *jb = vvec + kJbLd * help_idx;
                pair_barrier<false>(help_idx);
                if (jb[64] != 0.0) jacobi_pair_apply<N>(L, n16, p, qq, jb[lane], jb[32 + lane], lane);
            }
            __syncthreads();
        }
        ++sweeps;
        converged = (__syncthreads_or(big) == 0);
    }
    }

    PROF_ADD(14);
    PROF_COUNT(21, sweeps);
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
        if (tid >= ar.scan_from && tid < n && ssort[tid] < ar.acc) atomicMin(&misc[1], tid);
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
        if (st == HCORE_B200_OK) ar.rank_out[gt] = rnew;
    }
    if (st != HCORE_B200_OK) return;
    PROF_ADD(15);

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
    PROF_ADD(16);

    // ---- W = R_v * (R_u^T * U_r) -> Yv ----------------------------------------------------------
    block_gemm<HCORE_SVD_NT, kGemmSet, HCORE_SVD_UNR>(mview(L, 1, LDL), cview(Ru, N, 1), cview(Yu, 1, N), n16, r8, n16, warp, nwarps,
                            UpperTransK{n16});
    __syncthreads();
    block_gemm<HCORE_SVD_NT, kGemmSet, HCORE_SVD_UNR>(mview(Yv, 1, N), cview(Rv, 1, N), cview(L, 1, LDL), n16, r8, n16, warp, nwarps,
                            UpperK());
    PROF_ADD(17);
}

}  // namespace hcore_b200
