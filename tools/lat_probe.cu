// Latency probe (GPU box): dependent-issue latencies of the FP64 building blocks of the
// Householder panel step on this chip.  One warp (or one CTA for the barrier), clock64 around
// a dependent chain; prints cycles per operation.
#include <cstdio>
#include <cuda_runtime.h>
#include "../hcore_b200/csrc/common.cuh"
using namespace hcore_b200;

#define REP 256
__global__ void probe(double *out, long long *cyc, double x, double y) {
    __shared__ double sm[512];
    const int tid = threadIdx.x;
    sm[tid & 511] = x + tid;
    __syncthreads();
    double a = x + tid * 1e-9, b = y;
    long long t0, t1;
    int k = 0;
    // 0: DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) a = fma(a, b, y);
    t1 = clock64(); if (tid == 0) cyc[k] = t1 - t0; ++k;
    // 1: DADD chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) a = a + b;
    t1 = clock64(); if (tid == 0) cyc[k] = t1 - t0; ++k;
    // 2: shuffle(64-bit) + DADD round
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) a += __shfl_xor_sync(0xffffffffu, a, 1 + (i & 7));
    t1 = clock64(); if (tid == 0) cyc[k] = t1 - t0; ++k;
    // 3: rsqrt(double)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; ++i) a = rsqrt(a * a + 1.0);
    t1 = clock64(); if (tid == 0) cyc[k] = t1 - t0; ++k;
    // 4: __drcp_rn
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; ++i) a = __drcp_rn(a + 2.0);
    t1 = clock64(); if (tid == 0) cyc[k] = t1 - t0; ++k;
    // 5: make_householder
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < REP; ++i) { Householder h = make_householder(a, b + 1.0); a = h.scal + h.tau; }
    t1 = clock64(); if (tid == 0) cyc[k] = t1 - t0; ++k;
    // 6: LDS dependent (pointer chase through values)
    int idx = tid & 31;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; ++i) { double v = sm[idx]; idx = (idx + (int)(v * 0.0) + 32) & 511; }
    t1 = clock64(); if (tid == 0) cyc[k] = t1 - t0; ++k;
    a += idx;
    // 7: __syncthreads
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; ++i) __syncthreads();
    t1 = clock64(); if (tid == 0) cyc[k] = t1 - t0; ++k;
    // 8: STS -> bar -> LDS round trip
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < REP; ++i) { sm[tid & 255] = a; __syncthreads(); a += sm[(tid + 1) & 255]; }
    t1 = clock64(); if (tid == 0) cyc[k] = t1 - t0; ++k;
    // 9: DMMA dependent chain
    double d0 = a, d1 = b;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP; ++i) dmma_884(d0, d1, a, b);
    t1 = clock64(); if (tid == 0) cyc[k] = t1 - t0; ++k;
    // 10: DMMA 8 independent chains
    double e[8][2];
#pragma unroll
    for (int q = 0; q < 8; ++q) e[q][0] = e[q][1] = a + q;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < REP / 8; ++i)
#pragma unroll
        for (int q = 0; q < 8; ++q) dmma_884(e[q][0], e[q][1], a, b);
    t1 = clock64(); if (tid == 0) cyc[k] = t1 - t0; ++k;
#pragma unroll
    for (int q = 0; q < 8; ++q) d0 += e[q][0] + e[q][1];
    // 11: sqrt(double)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; ++i) a = sqrt(a * a + 1.0);
    t1 = clock64(); if (tid == 0) cyc[k] = t1 - t0; ++k;
    // 12: double division
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; ++i) a = 1.0 / (a + 2.0);
    t1 = clock64(); if (tid == 0) cyc[k] = t1 - t0; ++k;
    // 13: float rsqrt + 2 Newton steps in double
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; ++i) {
        double s = a * a + 1.0;
        double r = (double)rsqrtf((float)s);
        r = r * fma(-0.5 * s * r, r, 1.5);
        r = r * fma(-0.5 * s * r, r, 1.5);
        a = r;
    }
    t1 = clock64(); if (tid == 0) cyc[k] = t1 - t0; ++k;
    out[blockIdx.x * blockDim.x + tid] = a + d0 + d1;
}

int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 1024 * 8 * 16); cudaMalloc(&cyc, 64 * 8);
    const char *names[] = {"DFMA dep", "DADD dep", "SHFL64+DADD round", "rsqrt(double)+fma", "__drcp_rn+add", "make_householder(+add)",
                           "LDS dep", "__syncthreads", "STS+bar+LDS+add", "DMMA dep", "DMMA 8 indep (per op)", "sqrt(double)+fma",
                           "1.0/x + add", "rsqrtf+2 Newton (double)"};
    for (int threads : {32, 256}) {
        probe<<<1, threads>>>(out, cyc, 1.000001, 0.999999);
        cudaDeviceSynchronize();
        probe<<<1, threads>>>(out, cyc, 1.000001, 0.999999);
        cudaDeviceSynchronize();
        long long h[64]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads=%d\n", threads);
        for (int k = 0; k < 14; ++k) printf("  %-28s %8.1f cycles/op\n", names[k], (double)h[k] / (k == 10 ? REP : REP));
    }
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
}
