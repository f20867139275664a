// Probe (GPU box): cycles of one warp-local 128 x 8 sub-panel factorisation (warp_factor) in isolation,
// for 1..8 warps per CTA and 1..2 co-resident CTAs per SM, block 0 and TPQRT-type blocks.
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "../hcore_b200/csrc/wy_kernels.cuh"
using namespace hcore_b200;

template <int NW>
__global__ void __launch_bounds__(32 * NW, 16 / NW) probe(double *Rg_all, double *rec_all, long long *cyc, int reps, int blk0) {
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *yb = smem + warp * 2 * kYbuf;
    double *slot = smem + NW * 2 * kYbuf + warp * kRecDoubles;
    double *Rg = Rg_all + ((size_t)blockIdx.x * NW + warp) * 128 * 128;
    double *recg = rec_all + ((size_t)blockIdx.x * NW + warp) * kRecDoubles;
    long long total = 0;
    for (int r = 0; r < reps; ++r) {
        double bx[16], by[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            unsigned h = ((threadIdx.x * 16 + i) * 2654435761u) ^ (r * 40503u) ^ (blockIdx.x * 977u);
            bx[i] = ((h >> 8) & 0xffff) / 65536.0 - 0.5;
            by[i] = ((h >> 4) & 0xffff) / 65536.0 - 0.5;
        }
        __syncwarp();
        long long t0 = clock64();
        if (blk0) warp_factor<128, true>(bx, by, yb, slot, recg, Rg, 8 * (r & 7), lane);
        else warp_factor<128, false>(bx, by, yb, slot, recg, Rg, 8 * (r & 7), lane);
        long long t1 = clock64();
        total += t1 - t0;
    }
    if (lane == 0) cyc[blockIdx.x * NW + warp] = total / reps;
}

template <int NW>
void run(int ctas, int blk0) {
    double *Rg, *rec; long long *cyc;
    cudaMalloc(&Rg, (size_t)ctas * NW * 128 * 128 * 8); cudaMemset(Rg, 0, (size_t)ctas * NW * 128 * 128 * 8);
    cudaMalloc(&rec, (size_t)ctas * NW * kRecBytes);
    cudaMalloc(&cyc, ctas * NW * 8);
    const size_t sm = (size_t)NW * (2 * kYbuf + kRecDoubles) * 8;
    cudaFuncSetAttribute(probe<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    probe<NW><<<ctas, 32 * NW, sm>>>(Rg, rec, cyc, 20, blk0);
    cudaDeviceSynchronize();
    static long long h[8192]; cudaMemcpy(h, cyc, ctas * NW * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < ctas * NW; ++i) avg += h[i]; avg /= ctas * NW;
    printf("NW=%d ctas=%4d blk0=%d : %9.0f cycles per warp_factor (%6.0f per Householder step)  [%s]\n", NW, ctas, blk0, avg, avg / 8,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(Rg); cudaFree(rec); cudaFree(cyc);
}

void phases(const char *tag) {
#ifdef HCORE_B200_WFPROBE
    unsigned long long h[64];
    cudaMemcpyFromSymbol(h, g_prof, sizeof(h));
    const char *nm[7] = {"loop top", "publish+syncwarp", "T column", "dots", "reduce+householder", "update", "outputs"};
    unsigned long long tot = 0; for (int i = 40; i < 47; ++i) tot += h[i];
    printf("  phases (%s):", tag);
    for (int i = 40; i < 47; ++i) printf(" %s %.1f%%", nm[i - 40], 100.0 * h[i] / tot);
    printf("  total %llu\n", tot);
    memset(h, 0, sizeof(h)); cudaMemcpyToSymbol(g_prof, h, sizeof(h));
#endif
}

int main() {
    for (int blk0 = 0; blk0 < 2; ++blk0) {
        run<1>(1, blk0); phases("1 warp"); run<1>(148, blk0); run<2>(148, blk0); run<4>(148, blk0); run<8>(148, blk0); run<8>(296, blk0);
    }
}
