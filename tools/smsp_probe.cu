// Which SM sub-partition do the warps of small CTAs land on?  (%warpid % 4 = SMSP of the warp slot.)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/smsp_probe.cu -o tools/smsp_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(int *hist, int spin) {
    unsigned wid, sm;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    extern __shared__ double sh[];
    if ((threadIdx.x & 31) == 0) atomicAdd(&hist[(threadIdx.x >> 5) * 4 + (wid & 3)], 1);
    // stay resident for a while so that several CTAs share the SM
    long long t0 = clock64();
    while (clock64() - t0 < spin) { sh[threadIdx.x] += 1.0; }
}
int main() {
    int *hist;
    cudaMalloc(&hist, 64 * sizeof(int));
    for (int nthr : {32, 64, 128}) {
        for (int smem : {30736, 48160}) {
            cudaMemset(hist, 0, 64 * sizeof(int));
            cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            probe<<<148 * 16, nthr, smem>>>(hist, 200000);
            cudaDeviceSynchronize();
            int h[64];
            cudaMemcpy(h, hist, sizeof(h), cudaMemcpyDeviceToHost);
            printf("threads/CTA=%3d smem=%5d :", nthr, smem);
            for (int w = 0; w < nthr / 32; ++w) printf("  warp%d -> SMSP[%d %d %d %d]", w, h[4 * w], h[4 * w + 1], h[4 * w + 2], h[4 * w + 3]);
            printf("  (%s)\n", cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
