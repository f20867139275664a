// Probe (GPU box): issue rate of independent FP64 instructions / LDS.128 / DMMA from ONE warp and from
// several warps of one SM sub-partition (cycles per warp-instruction).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void probe(double *out, long long *cyc, int reps) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    double a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = 1.0 + k + lane;
    const double m = 1.0000001, c = 1e-9;
    const double *p = sm + 2 * (lane & 3);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (MODE == 0) {            // 16 independent DFMA chains
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < 16; ++k) a[k] = fma(a[k], m, c);
        } else if (MODE == 1) {     // 16 x LDS.128 then 32 DFMA into 4 accumulators
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const double2 y = *reinterpret_cast<const double2 *>(p + 8 * k + ((r & 1) << 7));
                a[k & 3] = fma(y.x, m, a[k & 3]);
                a[4 + (k & 3)] = fma(y.y, m, a[4 + (k & 3)]);
            }
        } else if (MODE == 2) {     // 16 independent DMMA accumulator pairs
#pragma unroll
            for (int k = 0; k < 8; ++k) dmma(a[2 * k], a[2 * k + 1], m, c);
        } else if (MODE == 3) {     // 16 x LDS.128 + 32 independent DFMA (update pattern)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double2 y = *reinterpret_cast<const double2 *>(p + 8 * k + ((r & 1) << 7));
                a[2 * k] = fma(y.x, m, a[2 * k]);
                a[2 * k + 1] = fma(y.y, m, a[2 * k + 1]);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (lane == 0) cyc[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
}

template <int MODE>
void run(const char *name, int threads, int ninstr) {
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 32 * 8);
    const int reps = 200;
    probe<MODE><<<148, threads, 2048 * 8>>>(out, cyc, reps);
    cudaDeviceSynchronize();
    long long h[32]; cudaMemcpy(h, cyc, 32 * 8, cudaMemcpyDeviceToHost);
    printf("%-44s warps/SM=%2d : %7.2f cycles per warp-instruction (%s)\n", name, threads / 32, (double)h[0] / reps / ninstr,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int th : {32, 128, 256, 512, 1024}) {
        run<0>("64 independent DFMA (16 chains)", th, 64);
        run<1>("16 LDS.128 + 32 DFMA (8 chains), per instr", th, 48);
        run<3>("8 LDS.128 + 16 indep DFMA, per instr", th, 24);
        run<2>("8 independent DMMA", th, 8);
    }
}
