// Shared device primitives for the TLR update kernels (sm_100a).
//
// FP64 tensor work on Blackwell goes through the warp-level DMMA instruction
// (mma.sync.aligned.m8n8k4.f64): tcgen05.mma has no f64 kind.  Everything GEMM-shaped in
// this library (stacking products, block-reflector updates, core products, reconstruction)
// is expressed with block_gemm() below, which tiles an arbitrary strided C = A*B over the
// warps of a CTA in 8x8 DMMA output tiles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/hcore_b200.h"

namespace hcore_b200 {

constexpr int kWarp = 32;

// Optional in-kernel phase timing (make prof -> lib/libhcore_b200_prof.so, tools/phase_prof.py):
// one designated thread per CTA accumulates clock64() deltas between checkpoints.
#ifdef HCORE_B200_PROF
__device__ unsigned long long g_prof[64];
#define PROF_DECL(who) const bool _pw = (who); long long _pt = clock64();
#define PROF_ADD(slot)                                                              \
    do {                                                                            \
        if (_pw) {                                                                  \
            const long long _n = clock64();                                         \
            atomicAdd(&g_prof[slot], (unsigned long long)(_n - _pt));               \
            _pt = clock64();                                                        \
        }                                                                           \
    } while (0)
#define PROF_COUNT(slot, v) do { if (_pw) atomicAdd(&g_prof[slot], (unsigned long long)(v)); } while (0)
#else
#define PROF_DECL(who)
#define PROF_ADD(slot)
#define PROF_COUNT(slot, v)
#endif

// Sub-panel width of the blocked Householder factorisations and row-block height of the
// flat-tree tall-skinny QR.  tests/device_model.py uses the same constants.
constexpr int kPanelW = 16;
constexpr int kBlockH = 128;

__host__ __device__ __forceinline__ int round_up(int x, int m) { return (x + m - 1) / m * m; }

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// D(8x8) += A(8x4) * B(4x8), FP64 tensor core.  Fragment ownership (PTX ISA, mma.m8n8k4 f64):
//   a : A[lane/4][lane%4]        b : B[lane%4][lane/4]
//   d0, d1 : D[lane/4][2*(lane%4) + {0,1}]
__device__ __forceinline__ void dmma_884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Strided matrix view: element (i, j) at p[i * si + j * sj].
struct MatView {
    double *p;
    int si, sj;
    __device__ __forceinline__ double &at(int i, int j) const { return p[(long)i * si + (long)j * sj]; }
};
struct CMatView {
    const double *p;
    int si, sj;
    __device__ __forceinline__ double at(int i, int j) const { return p[(long)i * si + (long)j * sj]; }
};
__device__ __forceinline__ CMatView cview(const double *p, int si, int sj) { return CMatView{p, si, sj}; }
__device__ __forceinline__ MatView mview(double *p, int si, int sj) { return MatView{p, si, sj}; }
__device__ __forceinline__ CMatView cview(MatView m) { return CMatView{m.p, m.si, m.sj}; }

enum GemmMode {
    kGemmSet = 0,   // C  = A*B
    kGemmAdd = 1,   // C += A*B
    kGemmSub = 2    // C -= A*B
};

// Block-level C(MxN) {=,+=,-=} A(MxK) * B(KxN) with DMMA.
//   M % 8 == 0, N % 8 == 0, K % 4 == 0 (callers pad with zeros).
// Work item = one 8-row strip x NT consecutive 8-column tiles (the A fragment is reused NT
// times); items are dealt round-robin to the `nwarps` warps starting at `warp`.
// KLO/KHI restrict the contraction range per output tile (triangular operands): the
// functor klim(i0, j0, &k_begin, &k_end) may shrink [0, K) to a multiple-of-4 sub-range.
struct FullK {
    __device__ __forceinline__ void operator()(int, int, int &, int &) const {}
};

// Bounds-guarded read-only view for operands in global memory whose logical extent is
// smaller than the padded GEMM shape: reads outside [0,ni) x [0,nj) return 0.
struct GuardedView {
    const double *p;
    long si, sj;
    int ni, nj;
    __device__ __forceinline__ double at(int i, int j) const {
        return (i < ni && j < nj) ? __ldg(p + i * si + j * sj) : 0.0;
    }
};

template <int NT, GemmMode MODE, int UNR = 4, typename AView, typename BView, typename KLim = FullK>
__device__ __forceinline__ void block_gemm(MatView C, AView A, BView B, int M, int N, int K,
                                           int warp, int nwarps, KLim klim = KLim()) {
    const int lane = lane_id();
    const int g = lane >> 2, tg = lane & 3;
    const int ntile_n = N >> 3;
    const int ngroups_n = (ntile_n + NT - 1) / NT;
    const int nitems = (M >> 3) * ngroups_n;
    for (int item = warp; item < nitems; item += nwarps) {
        const int i0 = (item / ngroups_n) << 3;
        const int jt0 = (item % ngroups_n) * NT;
        double acc[NT][2];
#pragma unroll
        for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = 0.0;
        int kb = 0, ke = K;
        klim(i0, jt0 << 3, kb, ke);
        // UNR contraction steps of operand loads in flight (operands in global memory / L2: the loop is latency-bound)
#pragma unroll UNR
        for (int k = kb; k < ke; k += 4) {
            const double a = A.at(i0 + g, k + tg);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                if (jt0 + t < ntile_n) {
                    const double b = B.at(k + tg, ((jt0 + t) << 3) + g);
                    dmma_884(acc[t][0], acc[t][1], a, b);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (jt0 + t < ntile_n) {
                const int j = ((jt0 + t) << 3) + 2 * tg;
                double *c0 = &C.at(i0 + g, j);
                double *c1 = &C.at(i0 + g, j + 1);
                if (MODE == kGemmSet) {
                    *c0 = acc[t][0];
                    *c1 = acc[t][1];
                } else if (MODE == kGemmAdd) {
                    *c0 += acc[t][0];
                    *c1 += acc[t][1];
                } else {
                    *c0 -= acc[t][0];
                    *c1 -= acc[t][1];
                }
            }
        }
    }
}

// LAPACK dlarfg on (alpha, squared norm of the tail): returns beta, tau and the scale
// 1/(alpha-beta) to apply to the tail.  sigma2 == 0 -> H = I (tau = 0), as LAPACK does.
// Written with one rsqrt and one reciprocal instead of sqrt + two divisions (the scalar
// chain sits on the critical path of every Householder step):
//   nrm = sqrt(alpha^2 + sigma2),  beta = -sign(alpha) nrm,
//   tau = (beta - alpha)/beta = 1 + |alpha|/nrm,   1/(alpha - beta) = sign(alpha)/(|alpha| + nrm)
struct Householder {
    double beta, tau, scal;
};
__device__ __forceinline__ Householder make_householder(double alpha, double sigma2) {
    Householder h;
    if (sigma2 == 0.0) {
        h.beta = alpha;
        h.tau = 0.0;
        h.scal = 0.0;
    } else {
        const double s2 = fma(alpha, alpha, sigma2);
        const double rinv = rsqrt(s2);
        const double nrm = s2 * rinv;
        const double aabs = fabs(alpha);
        h.beta = -copysign(nrm, alpha);
        h.tau = fma(aabs, rinv, 1.0);
        h.scal = copysign(__drcp_rn(aabs + nrm), alpha);
    }
    return h;
}

// make_householder without library calls: rsqrt() and __drcp_rn() carry out-of-line slow paths, and a call inside the
// step body makes ptxas spill the register-resident column around it.  Hardware seeds (~2^-22) + two Newton steps:
// a few ulps, arguments are positive and normal here (sigma2 > 0).
__device__ __forceinline__ Householder make_householder_inline(double alpha, double sigma2) {
    Householder h;
    const double s2 = fma(alpha, alpha, sigma2);
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s2));
    y = fma(0.5 * y, fma(-s2 * y, y, 1.0), y);
    y = fma(0.5 * y, fma(-s2 * y, y, 1.0), y);
    const double nrm = s2 * y, aabs = fabs(alpha), den = aabs + nrm;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    r = fma(r, fma(-den, r, 1.0), r);
    r = fma(r, fma(-den, r, 1.0), r);
    const bool nz = sigma2 != 0.0;
    h.beta = nz ? -copysign(nrm, alpha) : alpha;
    h.tau = nz ? fma(aabs, y, 1.0) : 0.0;
    h.scal = nz ? copysign(r, alpha) : 0.0;
    return h;
}

// Named barrier for a subset of the CTA's warps (ids 1..15; 0 is __syncthreads).
__device__ __forceinline__ void named_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Sum 16 per-lane values over the 32 lanes of a warp by recursive halving: after the call
// every lane holds in the return value the warp total of column reduce16_col(lane).
// 16 double exchanges instead of the 80 of 16 separate butterflies.
__device__ __forceinline__ int reduce16_col(int lane) { return (lane >> 1) & 15; }
__device__ __forceinline__ double reduce16(const double (&p)[16], int lane) {
    double a8[8], a4[4], a2[2];
    {
        const bool up = lane & 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double keep = up ? p[j + 8] : p[j];
            const double send = up ? p[j] : p[j + 8];
            a8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double keep = up ? a8[j + 4] : a8[j];
            const double send = up ? a8[j] : a8[j + 4];
            a4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool up = lane & 4;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double keep = up ? a4[j + 2] : a4[j];
            const double send = up ? a4[j] : a4[j + 2];
            a2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    const bool up = lane & 2;
    const double keep = up ? a2[1] : a2[0];
    const double send = up ? a2[0] : a2[1];
    double v = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// Round-robin (chess tournament) pairing, closed form: pair k of step `step`, n even.
__device__ __forceinline__ void rr_pair(int n, int step, int k, int &p, int &q) {
    int a, b;
    if (k == 0) {
        a = n - 1;
        b = step;
    } else {
        a = step + k;                       // 0 <= step < n - 1, 0 < k < n / 2: one conditional
        a = a >= n - 1 ? a - (n - 1) : a;   // subtraction replaces the modulo
        b = step - k + n - 1;
        b = b >= n - 1 ? b - (n - 1) : b;
    }
    p = a < b ? a : b;
    q = a < b ? b : a;
}

}  // namespace hcore_b200
