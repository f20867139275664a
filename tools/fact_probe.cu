// Probe (GPU box): cycles of one 128 x 16 panel factorisation (panel_factor) in isolation,
// for 1..N co-resident CTAs per SM, block 0 and TPQRT-type blocks.
#include <cstdio>
#include <cuda_runtime.h>
#include "../hcore_b200/csrc/wy_kernels.cuh"
using namespace hcore_b200;

template <int NW>
__global__ void __launch_bounds__(32 * NW) probe(double *Rg_all, long long *cyc, int reps, int blk0) {
    using SM = TsqrSmem<NW>;
    extern __shared__ __align__(16) double smem[];
    double *Pf = smem + SM::kPf;
    double *rec = smem + SM::kRec0;
    double *Rg = Rg_all + (size_t)blockIdx.x * 128 * 128;
    const int tid = threadIdx.x;
    long long total = 0;
    for (int r = 0; r < reps; ++r) {
        for (int idx = tid; idx < 16 * kLdPf; idx += 32 * NW) {
            unsigned h = (idx * 2654435761u) ^ (r * 40503u) ^ (blockIdx.x * 977u);
            Pf[idx] = ((h >> 8) & 0xffff) / 65536.0 - 0.5;
        }
        __syncthreads();
        long long t0 = clock64();
        if (blk0) panel_factor<128, NW, true>(smem, rec, Rg, 16 * (r & 3), tid); else panel_factor<128, NW, false>(smem, rec, Rg, 16 * (r & 3), tid);
        long long t1 = clock64();
        total += t1 - t0;
    }
    if (tid == 0) cyc[blockIdx.x] = total / reps;
}

template <int NW>
void run(int ctas, int blk0) {
    double *Rg; long long *cyc;
    cudaMalloc(&Rg, (size_t)ctas * 128 * 128 * 8); cudaMemset(Rg, 0, (size_t)ctas * 128 * 128 * 8);
    cudaMalloc(&cyc, ctas * 8);
    cudaFuncSetAttribute(probe<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TsqrSmem<NW>::kBytes);
    probe<NW><<<ctas, 32 * NW, TsqrSmem<NW>::kBytes>>>(Rg, cyc, 20, blk0);
    cudaDeviceSynchronize();
    long long h[2048]; cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < ctas; ++i) avg += h[i]; avg /= ctas;
    printf("NW=%d ctas=%4d blk0=%d : %9.0f cycles per panel_factor (%6.0f per Householder step)  [%s]\n", NW, ctas, blk0, avg, avg / 16,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(Rg); cudaFree(cyc);
}

int main() {
    for (int blk0 = 0; blk0 < 2; ++blk0) {
        run<8>(1, blk0); run<8>(148, blk0); run<8>(296, blk0);
        run<4>(1, blk0); run<4>(148, blk0); run<4>(296, blk0); run<4>(444, blk0);
    }
}
