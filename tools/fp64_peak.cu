// Calibration micro-benchmark (GPU box): peak FP64 throughput of the DFMA pipe and of the
// DMMA (mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 f64) tensor path on this chip.
// MEASURED_PEAKS.json has no FP64 entry; bench.py's roofline uses the number measured here
// (profiles/fp64_peak_r1.txt) next to a cuBLAS DGEMM calibration taken in bench.py itself.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double *out, int iters, double x, double y) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma884_kernel(double *out, int iters, double x, double y) {
    double d[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) d[i][0] = d[i][1] = threadIdx.x * 1e-9;
    double a = x + threadIdx.x * 1e-12, b = y;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma16816_kernel(double *out, int iters, double x, double y) {
    double d[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = threadIdx.x * 1e-9;
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = x + i * 1e-12;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = y + i * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(d[i][0]), "+d"(d[i][1]), "+d"(d[i][2]), "+d"(d[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma1684_kernel(double *out, int iters, double x, double y) {
    double d[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = threadIdx.x * 1e-9;
    double a[2] = {x, x + 1e-12}, b = y;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                         : "+d"(d[i][0]), "+d"(d[i][1]), "+d"(d[i][2]), "+d"(d[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA and DFMA interleaved in one instruction stream (independent chains): if the tensor path and the scalar
// FP64 pipe were separate units the combined rate would exceed either alone.
template <int ILP>
__global__ void mix_kernel(double *out, int iters, double x, double y) {
    double d[ILP][2], acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { d[i][0] = d[i][1] = threadIdx.x * 1e-9; acc[i] = threadIdx.x * 1e-9 + i; }
    double a = x + threadIdx.x * 1e-12, b = y;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a), "d"(b));
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[(i + j) % ILP] = fma(acc[(i + j) % ILP], x, y);   // 8 DFMA = the flops of one DMMA
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// half of the CTA's warps issue DMMA only, the other half DFMA only
template <int ILP>
__global__ void split_kernel(double *out, int iters, double x, double y) {
    double d[ILP][2], acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { d[i][0] = d[i][1] = threadIdx.x * 1e-9; acc[i] = threadIdx.x * 1e-9 + i; }
    double a = x + threadIdx.x * 1e-12, b = y;
    if ((threadIdx.x >> 5) & 1) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < ILP; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a), "d"(b));
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], x, y);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %.0f MHz\n", p.name, sms, p.clockRate / 1000.0);
    double *out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 20000;
    for (int threads : {256, 512, 1024}) {
        for (int bps : {1, 2}) {
            if (threads * bps > 2048) continue;
            const int grid = sms * bps;
            double ms = time_ms([&] { dfma_kernel<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
            double fl = 2.0 * 8 * iters * (double)threads * grid;
            printf("DFMA      threads=%4d ctas/sm=%d : %8.2f TFLOP/s\n", threads, bps, fl / ms * 1e-9);
            ms = time_ms([&] { dmma884_kernel<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
            fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)(threads / 32) * grid;
            printf("DMMA.884  threads=%4d ctas/sm=%d : %8.2f TFLOP/s\n", threads, bps, fl / ms * 1e-9);
            ms = time_ms([&] { mix_kernel<8><<<grid, threads>>>(out, iters / 2, 1.0000001, 1e-9); });
            fl = (2.0 * 8 * 8 * 4 + 2.0 * 8 * 32) * 8 * (iters / 2) * (double)(threads / 32) * grid;
            printf("DMMA+DFMA interleaved threads=%4d ctas/sm=%d : %8.2f TFLOP/s (sum of both)\n", threads, bps, fl / ms * 1e-9);
            ms = time_ms([&] { split_kernel<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
            fl = (2.0 * 8 * 8 * 4 + 2.0 * 8 * 32) * 8 * iters * (double)(threads / 64) * grid;
            printf("DMMA | DFMA by warp parity threads=%4d ctas/sm=%d : %8.2f TFLOP/s (sum of both)\n", threads, bps, fl / ms * 1e-9);
            ms = time_ms([&] { dmma1684_kernel<4><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); });
            fl = 2.0 * 16 * 8 * 4 * 4 * iters * (double)(threads / 32) * grid;
            printf("DMMA.1684 threads=%4d ctas/sm=%d : %8.2f TFLOP/s\n", threads, bps, fl / ms * 1e-9);
            ms = time_ms([&] { dmma16816_kernel<4><<<grid, threads>>>(out, iters / 4, 1.0000001, 1e-9); });
            fl = 2.0 * 16 * 8 * 16 * 4 * (iters / 4) * (double)(threads / 32) * grid;
            printf("DMMA.16816 threads=%4d ctas/sm=%d : %8.2f TFLOP/s\n", threads, bps, fl / ms * 1e-9);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
