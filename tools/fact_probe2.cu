// Probe (GPU box): per-phase clock stamps of one panel_factor call (instrumented copy).
#include <cstdio>
#include <cuda_runtime.h>
#include "../hcore_b200/csrc/wy_kernels.cuh"
namespace hcore_b200 {
#include "fact_probe_body.inc"
}
using namespace hcore_b200;
__global__ void __launch_bounds__(256) probe(double *Rg, long long *stamp, int blk0) {
    using SM = TsqrSmem<8>;
    extern __shared__ __align__(16) double smem[];
    double *Pf = smem + SM::kPf;
    const int tid = threadIdx.x;
    for (int r = 0; r < 2; ++r) {
        for (int idx = tid; idx < 16 * kLdPf; idx += 256) { unsigned h = (idx * 2654435761u) ^ (r * 40503u); Pf[idx] = ((h >> 8) & 0xffff) / 65536.0 - 0.5; }
        __syncthreads();
        panel_factor_t<128, 8>(smem, smem + SM::kRec0, Rg, blk0 != 0, 16, tid, stamp);
    }
}
int main() {
    double *Rg; long long *st; cudaMalloc(&Rg, 128 * 128 * 8); cudaMemset(Rg, 0, 128 * 128 * 8); cudaMalloc(&st, 16 * 8 * 12 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TsqrSmem<8>::kBytes);
    for (int blk0 = 0; blk0 < 2; ++blk0) {
        cudaMemset(st, 0, 16 * 8 * 12 * 8);
        probe<<<1, 256, TsqrSmem<8>::kBytes>>>(Rg, st, blk0);
        cudaDeviceSynchronize();
        static long long h[16 * 8 * 12]; cudaMemcpy(h, st, sizeof(h), cudaMemcpyDeviceToHost);
        printf("blk0=%d  [%s]\n", blk0, cudaGetErrorString(cudaGetLastError()));
        for (int ii = 2; ii < 8; ++ii) {
            int wo = ii % 8, wn = (ii + 3) % 8, wt = (ii - 1) % 8;
            long long *o = h + (ii * 8 + wo) * 12, *n = h + (ii * 8 + wn) * 12, *tt = h + (ii * 8 + wt) * 12, *nx = h + ((ii + 1) * 8 + wo) * 12;
            printf(" step %2d owner: top->norm %4lld  alpha %4lld  house %4lld  store %4lld  bar %4lld | vvdot %4lld  reduce %4lld  update %4lld  tail %4lld | step total %5lld\n", ii,
                   o[1] - o[0], o[2] - o[1], o[3] - o[2], o[4] - o[3], o[5] - o[4], o[6] - o[5], o[7] - o[6], o[8] - o[7], o[9] - o[8], nx[0] - o[0]);
            printf("         other: top->bar %4lld  bar %4lld | vvdot %4lld  reduce %4lld  update %4lld  tail %4lld ; Twarp tail %4lld\n",
                   n[4] - n[0], n[5] - n[4], n[6] - n[5], n[7] - n[6], n[8] - n[7], n[9] - n[8], tt[9] - tt[8]);
        }
    }
}
