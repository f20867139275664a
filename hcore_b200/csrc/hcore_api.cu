// Host side of libhcore_b200.so: context, workspace, kernel launches, the batched C ABI
// (include/hcore_b200.h) and the legacy single-tile drop-in symbols (include/hcore_d.h).
// There is no CPU fallback anywhere in this file: every entry point needs a CUDA device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <mutex>
#include <dlfcn.h>

#include "tlr_kernels.cuh"
#include "dense_kernels.cuh"
#include "chol_kernels.cuh"
#include "zgemm_kernels.cuh"
#include "../../include/hcore_d.h"
#include "../../include/hcore_s.h"
#include "../../include/hcore_z.h"

using namespace hcore_b200;

// ---- link-compatible tunables of the reference library (accepted, ignored) ---------------
extern "C" {
int use_trmm = 0, gemm_print_index = 0, gemm_print_mat = 0, hc_nelm_limit = 512;
int syrk_print_index = 0, gemmcd_print_index = 0, gemmcd_print_mat = 0, gemmcd_use_scratch = 0;
int sgemm_use_trmm = 0, sgemm_print_index = 0, sgemm_print_mat = 0;
int gemmbdcd_print_index = 0, gemmbdcd_print_mat = 0, gemmfrk_print_index = 0, gemmfrk_print_mat = 0;
// `use_scratch` is DECLARED by the reference library and DEFINED by the application (reference
// src/compute/hcore_dgemm.c:38, testing/testing_dgemm.c:26): a weak reference keeps both kinds of callers linking.
extern int use_scratch __attribute__((weak));
}

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

constexpr int kHostBufs = 3;
struct HostPipe {
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
    cudaEvent_t in_done[kHostBufs], comp_done[kHostBufs], out_done[kHostBufs];
    bool ready = false;
};

}  // namespace

struct hcore_b200_context {
    int device = 0;
    int chunk_limit = 0;            // user override, 0 = default
    int host_slice = 0;             // tiles per slice of the host-buffer pipeline (0 = default; HCORE_B200_HOST_SLICE)
    int num_sms = 148;
    int64_t launches = 0;
    std::string err;
    // TLR workspace
    DevBuf rec, G, R, Y, Lcore, sigma, ints, ptrs;
    // dense workspace
    DevBuf X, Ywork;
    // trsm / potrf workspace (inverses of the 64 x 64 diagonal blocks, per-tile info)
    DevBuf Winv, info;
    // tile compression workspace (test matrix, internal factor buffers, a zero vector, rank words)
    DevBuf cmpOm, cmpCU, cmpCV, cmpZero, cmpInts;
    bool attr_done[8] = {false, false, false, false, false, false, false, false};
    // complex (zgemm) workspace
    DevBuf zpanel, ztau, zT, zY, zsigma, zints;
    // host-staging (HCORE_dgemm_batched_host and the legacy wrappers)
    DevBuf stage[2];
    DevBuf stage_ints[2];
    // the three-stream host pipeline (HCORE_*_batched_host): staging buffers, rank / status words, streams, events
    DevBuf hstage[3], hstage_ints[3];
    HostPipe hpipe;
    // pinned host mirror of the legacy wrappers' staging block
    void *pin = nullptr;
    size_t pin_bytes = 0;
    cudaError_t pin_reserve(size_t need) {
        if (need <= pin_bytes) return cudaSuccess;
        if (pin) cudaFreeHost(pin);
        pin = nullptr; pin_bytes = 0;
        cudaError_t e = cudaHostAlloc(&pin, need, cudaHostAllocDefault);
        if (e == cudaSuccess) pin_bytes = need;
        return e;
    }
    cudaStream_t streams[2] = {nullptr, nullptr};
    // debug / profiling
    int debug_stages = 4;
    int tsqr_warps = 2;
    bool stage_timing = false;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    double stage_ms[4] = {0, 0, 0, 0};
    // geometry of the last chunk (for debug_fetch)
    TlrWorkspace last_ws{};
    int last_N = 0;
    // split-phase state (HCORE_dgemm_batched_qr_svd -> HCORE_dgemm_batched_ormqr)
    struct { bool pending = false; int ntiles = 0, nb = 0, f32 = 0, maxrk = 0; } split;
};

namespace {

#define CUDA_TRY(ctx, expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                 \
            return HCORE_B200_ERR_CUDA;                                                      \
        }                                                                                    \
    } while (0)

int pick_width(int maxrk, int cap = HCORE_B200_MAX_STACK) {
    const int w = std::min(maxrk, cap);
    if (w <= 32) return 32;
    if (w <= 64) return 64;
    if (w <= 128) return 128;
    return 256;
}

size_t per_tile_ws_bytes(int N, int m_pad, int nblk) {
    size_t b = 0;
    b += 2ull * nblk * (N / kSubW) * kRecBytes;    // block-reflector records
    b += (size_t)m_pad * N * 8;                 // G
    b += 2ull * N * N * 8;                      // R
    b += 2ull * N * N * 8;                      // Y
    if (N > 128) b += (size_t)N * (N + 4) * 8;  // L (core matrix outside shared memory)
    b += (size_t)N * 8 + 16;                    // sigma + ints
    return b;
}

int default_chunk(const hcore_b200_context *ctx, int N, int m_pad, int nblk) {
    if (ctx->chunk_limit > 0) return std::min(ctx->chunk_limit, 65535);
    // a few waves of CTAs per launch, bounded to ~6 GiB of workspace
    size_t per = per_tile_ws_bytes(N, m_pad, nblk);
    size_t cap = (6ull << 30) / per;
    // 16 x #SM tiles = 2368 on B200 (whole waves for the one-CTA-per-SM SVD kernel).  15 x #SM gives whole waves
    // for the QR kernel too and is 0.8 % faster on batches that are a multiple of it, but loses 1 % on the
    // benchmark's 16 384 tiles (a short last chunk): tools/exp_chunk.py, profiles/bench_r2j.json
    int c = ctx->num_sms * 16;
    if ((size_t)c > cap) c = (int)cap;
    return std::min(std::max(c, 1), 65535);          // the apply kernel's tile index is gridDim.z
}

template <int N, int TW, int AW>
int launch_chunk(hcore_b200_context *ctx, const TlrArgs &ar, const TlrWorkspace &ws, cudaStream_t st) {
    constexpr int kSlot = (N == 32) ? 0 : (N == 64) ? 1 : (N == 128) ? 2 : 3;
    if (!ctx->attr_done[kSlot]) {
        ctx->attr_done[kSlot] = true;
        CUDA_TRY(ctx, cudaFuncSetAttribute(tlr_pg_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)PgSmemT<PgWidth<N>::value>::kBytes));
        CUDA_TRY(ctx, cudaFuncSetAttribute(tlr_tsqr_kernel<N, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)TsqrSmem<TW>::kBytes));
        CUDA_TRY(ctx, cudaFuncSetAttribute(tlr_svd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SvdSmem<N>::kBytes));
        CUDA_TRY(ctx, cudaFuncSetAttribute(tlr_apply_kernel<N, AW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)ApplySmem<AW>::kBytes));
    }
    if (ar.phase == 2) {           // split-phase second half: reflector application only
        dim3 grid((N + 8 * AW - 1) / (8 * AW), 2, ar.ntiles);
        tlr_apply_kernel<N, AW><<<grid, 32 * AW, ApplySmem<AW>::kBytes, st>>>(ar, ws);
        ctx->launches++;
        CUDA_TRY(ctx, cudaGetLastError());
        return HCORE_B200_OK;
    }
    const bool tm = ctx->stage_timing;
    if (tm) cudaEventRecord(ctx->ev[0], st);
    if (ctx->debug_stages >= 1) {
        if (ar.Bd) {
            CUDA_TRY(ctx, cudaFuncSetAttribute(tlr_gdense_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)GdSmem<N>::kBytes));
            tlr_gdense_kernel<N><<<dim3(ws.nblk, ar.ntiles), 256, GdSmem<N>::kBytes, st>>>(ar, ws);
        } else {
            tlr_pg_kernel<N><<<ar.ntiles, 256, PgSmemT<PgWidth<N>::value>::kBytes, st>>>(ar, ws);
        }
        ctx->launches++;
    }
    if (tm) cudaEventRecord(ctx->ev[1], st);
    if (ctx->debug_stages >= 2) {
        tlr_tsqr_kernel<N, TW><<<2 * ar.ntiles, 32 * TW, TsqrSmem<TW>::kBytes, st>>>(ar, ws);
        ctx->launches++;
    }
    if (tm) cudaEventRecord(ctx->ev[2], st);
    if (ctx->debug_stages >= 3) {
        tlr_svd_kernel<N><<<ar.ntiles, SvdThreads<N>::value, SvdSmem<N>::kBytes, st>>>(ar, ws);
        ctx->launches++;
    }
    if (tm) cudaEventRecord(ctx->ev[3], st);
    if (ctx->debug_stages >= 4 && ar.phase == 0) {
        dim3 grid((N + 8 * AW - 1) / (8 * AW), 2, ar.ntiles);
        tlr_apply_kernel<N, AW><<<grid, 32 * AW, ApplySmem<AW>::kBytes, st>>>(ar, ws);
        ctx->launches++;
    }
    if (tm) {
        cudaEventRecord(ctx->ev[4], st);
        cudaEventSynchronize(ctx->ev[4]);
        for (int i = 0; i < 4; ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]);
            ctx->stage_ms[i] += ms;
        }
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return HCORE_B200_OK;
}

template <int N>
int launch_chunk_n(hcore_b200_context *ctx, const TlrArgs &ar, const TlrWorkspace &ws, cudaStream_t st) {
    // kernel shapes: warps per QR CTA (8 columns each) and per apply CTA; tunable for experiments
    if (ctx->tsqr_warps == 8) return launch_chunk<N, 8, 4>(ctx, ar, ws, st);
    if (ctx->tsqr_warps == 2) return launch_chunk<N, 2, 4>(ctx, ar, ws, st);
    if (ctx->tsqr_warps == 4) return launch_chunk<N, 4, 4>(ctx, ar, ws, st);
    return launch_chunk<N, 2, 4>(ctx, ar, ws, st);
}

// out[i] = base + i * stride_bytes (factor arrays at a fixed stride; element size 8 or 4)
__global__ void fill_ptrs_kernel(const double **out, const char *base, int64_t stride_bytes, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = reinterpret_cast<const double *>(base + (size_t)i * stride_bytes);
}

}  // namespace

extern "C" {

const char *hcore_b200_version(void) { return "hcore_b200 0.1 (sm_100a, FP64 DMMA)"; }

int hcore_b200_create(hcore_b200_handle_t *handle, int device) {
    if (!handle) return HCORE_B200_ERR_ARG;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        fprintf(stderr, "hcore_b200: no CUDA device available -- this library has no CPU fallback\n");
        return HCORE_B200_ERR_CUDA;
    }
    auto *ctx = new hcore_b200_context();
    if (device < 0) cudaGetDevice(&device);
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return HCORE_B200_ERR_CUDA; }
    cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (const char *e = getenv("HCORE_B200_TSQR_WARPS")) ctx->tsqr_warps = atoi(e);
    if (const char *e = getenv("HCORE_B200_HOST_SLICE")) ctx->host_slice = atoi(e);
    for (auto &e : ctx->ev) cudaEventCreate(&e);
    for (auto &s : ctx->streams) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    *handle = ctx;
    return HCORE_B200_OK;
}

int hcore_b200_destroy(hcore_b200_handle_t ctx) {
    if (!ctx) return HCORE_B200_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (DevBuf *b : {&ctx->rec, &ctx->G, &ctx->R, &ctx->Y, &ctx->Lcore, &ctx->sigma, &ctx->ints, &ctx->ptrs, &ctx->X, &ctx->Ywork, &ctx->Winv, &ctx->info, &ctx->cmpOm, &ctx->cmpCU, &ctx->cmpCV, &ctx->cmpZero, &ctx->cmpInts, &ctx->zpanel, &ctx->ztau, &ctx->zT, &ctx->zY, &ctx->zsigma, &ctx->zints,
                      &ctx->stage[0], &ctx->stage[1], &ctx->stage_ints[0], &ctx->stage_ints[1]})
        b->release();
    for (int b = 0; b < kHostBufs; ++b) { ctx->hstage[b].release(); ctx->hstage_ints[b].release(); }
    if (ctx->pin) cudaFreeHost(ctx->pin);
    if (ctx->hpipe.ready) {
        for (cudaStream_t st : {ctx->hpipe.s_in, ctx->hpipe.s_comp, ctx->hpipe.s_out}) cudaStreamDestroy(st);
        for (int b = 0; b < kHostBufs; ++b) { cudaEventDestroy(ctx->hpipe.in_done[b]); cudaEventDestroy(ctx->hpipe.comp_done[b]); cudaEventDestroy(ctx->hpipe.out_done[b]); }
    }
    for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
    for (auto &s : ctx->streams) if (s) cudaStreamDestroy(s);
    delete ctx;
    return HCORE_B200_OK;
}

int hcore_b200_set_chunk(hcore_b200_handle_t ctx, int tiles_per_chunk) {
    if (!ctx || tiles_per_chunk < 0) return HCORE_B200_ERR_ARG;
    ctx->chunk_limit = tiles_per_chunk;
    return HCORE_B200_OK;
}

const char *hcore_b200_last_error(hcore_b200_handle_t ctx) { return ctx ? ctx->err.c_str() : "null handle"; }
int64_t hcore_b200_launch_count(hcore_b200_handle_t ctx) { return ctx ? ctx->launches : 0; }

void hcore_b200_shard_range(int64_t ntiles, int world, int rank, int64_t *first, int64_t *count) {
    // contiguous blocks, the first (ntiles % world) ranks get one extra tile
    const int64_t base = ntiles / world, rem = ntiles % world;
    const int64_t f = rank * base + std::min<int64_t>(rank, rem);
    if (first) *first = f;
    if (count) *count = base + (rank < rem ? 1 : 0);
}

// ---- multi-GPU: the final {rank, status} gather over NCCL (north star: "NCCL only for the final rank/status gather") ----
// NCCL is resolved at run time (the symbols of an already loaded libnccl -- the application's or torch's -- else
// dlopen("libnccl.so.2")): the library itself has no link-time dependency on it.
namespace {
typedef int (*nccl_allgather_fn)(const void *, void *, size_t, int, void *, cudaStream_t);
nccl_allgather_fn resolve_nccl_allgather() {
    static nccl_allgather_fn fn = nullptr;
    if (!fn) {
        void *sym = dlsym(RTLD_DEFAULT, "ncclAllGather");
        if (!sym) {
            void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (h) sym = dlsym(h, "ncclAllGather");
        }
        fn = reinterpret_cast<nccl_allgather_fn>(sym);
    }
    return fn;
}
// dst[i * cmax + j] blocks -> contiguous tile order (the first `rem` ranks own base + 1 tiles, the others base)
__global__ void gather_compact_kernel(const int32_t *padded, int32_t *out, int64_t ntiles, int world, int64_t cmax) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= ntiles) return;
    const int64_t base = ntiles / world, rem = ntiles % world;
    int r; int64_t off;
    if (i < rem * (base + 1)) { r = (int)(i / (base + 1)); off = i - r * (base + 1); }
    else { r = (int)(rem + (i - rem * (base + 1)) / (base > 0 ? base : 1)); off = i - rem * (base + 1) - (r - rem) * base; }
    out[i] = padded[r * cmax + off];
}
}  // namespace

int hcore_b200_gather_status(hcore_b200_handle_t ctx, void *nccl_comm, int world, int rank, int64_t ntiles,
                             const int32_t *d_rank_local, const int32_t *d_status_local,
                             int32_t *d_rank_all, int32_t *d_status_all, void *stream) {
    if (!ctx || world <= 0 || rank < 0 || rank >= world || ntiles < 0) return HCORE_B200_ERR_ARG;
    if (ntiles == 0) return HCORE_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int64_t first, count;
    hcore_b200_shard_range(ntiles, world, rank, &first, &count);
    if (world == 1 && !nccl_comm) {                      // single rank: plain copies
        if (d_rank_all && d_rank_local) CUDA_TRY(ctx, cudaMemcpyAsync(d_rank_all, d_rank_local, ntiles * 4, cudaMemcpyDeviceToDevice, st));
        if (d_status_all && d_status_local) CUDA_TRY(ctx, cudaMemcpyAsync(d_status_all, d_status_local, ntiles * 4, cudaMemcpyDeviceToDevice, st));
        return HCORE_B200_OK;
    }
    nccl_allgather_fn allgather = resolve_nccl_allgather();
    if (!allgather || !nccl_comm) { ctx->err = "hcore_b200_gather_status: NCCL (libnccl.so.2) not available or null communicator"; return HCORE_B200_ERR_ARG; }
    const int64_t cmax = (ntiles + world - 1) / world;
    CUDA_TRY(ctx, ctx->info.reserve((size_t)(world + 1) * cmax * sizeof(int32_t)));
    int32_t *send = (int32_t *)ctx->info.p, *recv = send + cmax;
    const int32_t *src[2] = {d_rank_local, d_status_local};
    int32_t *dst[2] = {d_rank_all, d_status_all};
    for (int q = 0; q < 2; ++q) {
        if (!src[q] || !dst[q]) continue;
        CUDA_TRY(ctx, cudaMemsetAsync(send, 0, cmax * 4, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(send, src[q], count * 4, cudaMemcpyDeviceToDevice, st));
        const int rc = allgather(send, recv, (size_t)cmax, /*ncclInt32*/ 2, nccl_comm, st);
        if (rc != 0) { ctx->err = "ncclAllGather failed (" + std::to_string(rc) + ")"; return HCORE_B200_ERR_CUDA; }
        gather_compact_kernel<<<(unsigned)((ntiles + 255) / 256), 256, 0, st>>>(recv, dst[q], ntiles, world, cmax);
        ctx->launches++;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return HCORE_B200_OK;
}

uint64_t hcore_b200_gemm_flops(int m_, int a_, int b_, int c_, int r_) {
    // reference src/misc/descutil/flop_counts.c:3-70 (float divisions and truncations included)
    const uint64_t m = m_, a = a_, b = b_, r = r_, n = (uint64_t)a_ + c_;
    auto geqrf = [](uint64_t mm, uint64_t nn) {
        return 2 * mm * nn * nn - (uint64_t)(2 * nn * nn * nn / 3.0f) + 2 * mm * nn + (uint64_t)(17 * nn / 3.0f);
    };
    uint64_t f = 0;
    f += geqrf(m, n);                                   // __qra  (hcore_dgemm.c:125)
    f += a * b * m * 2 + m * a * b * 2;                 // __qrb GEMMs (:179, :208)
    f += geqrf(m, n);                                   // __qrb QR (:224)
    f += 22 * n * n * n;                                // gesvd model (:434)
    f += r * n;                                         // sigma scaling (:493)
    f += 4 * r * m * n - 2 * r * n * n + 3 * r * n;     // ormqr 'L' (:595)  (m,n,k) = (m,r,n)
    f += 4 * m * r * n - 2 * r * n * n + 2 * r * n + m * n - n * n / 2 + n / 2;  // ormqr 'R' (:666) (m,n,k)=(r,m,n)
    return f;
}

}  // extern "C"

// Batched update on arrays of device pointers; f32 = 1: the arrays hold float (HCORE_sgemm).
static int tlr_batched(hcore_b200_handle_t ctx, int ntiles, int nb, double alpha,
                        const double *const *d_AU, const double *const *d_AV, const int32_t *d_Ark, int lda,
                        const double *const *d_BU, const double *const *d_BV, const int32_t *d_Brk, int ldb,
                        double beta, double *const *d_CU, double *const *d_CV, int32_t *d_Crk, int ldc,
                        int rk, int maxrk, double acc, int32_t *d_status, void *stream, int f32,
                        int phase = 0, int32_t *d_newrank = nullptr, const double *const *d_Bdense = nullptr,
                        int scan_from = 1) {
    // d_Bdense != null: B-dense variant, d_BU / d_BV / d_Brk are unused and ldb is the leading dimension of B
    if (!ctx || ntiles < 0 || nb <= 0 || maxrk <= 0 || lda < nb || ldb < nb || ldc < nb) return HCORE_B200_ERR_ARG;
    if (phase == 1 && !d_newrank) return HCORE_B200_ERR_ARG;
    if (ntiles == 0) return HCORE_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int N = pick_width(maxrk, HCORE_B200_MAX_STACK_D);
    const int m_pad = round_up(std::max(nb, kBlockH), kBlockH);
    const int nblk = m_pad / kBlockH;
    const int chunk = std::min(ntiles, default_chunk(ctx, N, m_pad, nblk));
    ctx->split.pending = false;
    if (phase == 1 && chunk < ntiles) {
        // the reflectors of every tile have to survive until the second half: one workspace chunk
        ctx->err = "split-phase batch exceeds the workspace chunk (" + std::to_string(chunk) +
                   " tiles): split the batch or raise hcore_b200_set_chunk()";
        return HCORE_B200_ERR_ARG;
    }

    CUDA_TRY(ctx, ctx->rec.reserve(2ull * chunk * nblk * (N / kSubW) * kRecBytes));
    CUDA_TRY(ctx, ctx->G.reserve((size_t)chunk * m_pad * N * 8));
    CUDA_TRY(ctx, ctx->R.reserve(2ull * chunk * N * N * 8));
    CUDA_TRY(ctx, ctx->Y.reserve(2ull * chunk * N * N * 8));
    if (N > 128) CUDA_TRY(ctx, ctx->Lcore.reserve((size_t)chunk * N * (N + 4) * 8));
    CUDA_TRY(ctx, ctx->sigma.reserve((size_t)chunk * N * 8));
    CUDA_TRY(ctx, ctx->ints.reserve((size_t)chunk * 4 * sizeof(int)));

    TlrWorkspace ws{};
    ws.rec = (double *)ctx->rec.p;
    ws.G = (double *)ctx->G.p;
    ws.R = (double *)ctx->R.p;
    ws.Y = (double *)ctx->Y.p;
    ws.L = (double *)ctx->Lcore.p;
    ws.sigma = (double *)ctx->sigma.p;
    ws.ncols = (int *)ctx->ints.p;
    ws.newrank = ws.ncols + chunk;
    ws.st = ws.newrank + chunk;
    ws.sweeps = ws.st + chunk;
    ws.chunk = chunk;
    ws.m_pad = m_pad;
    ws.nblk = nblk;
    ctx->last_ws = ws;
    ctx->last_N = N;
    if (ctx->stage_timing) for (double &x : ctx->stage_ms) x = 0;

    for (int t0 = 0; t0 < ntiles; t0 += chunk) {
        TlrArgs ar{};
        ar.AU = d_AU; ar.AV = d_AV; ar.BU = d_BU; ar.BV = d_BV; ar.CU = d_CU; ar.CV = d_CV;
        ar.Ark = d_Ark; ar.Brk = d_Brk; ar.Crk = d_Crk; ar.status = d_status;
        ar.nb = nb; ar.lda = lda; ar.ldb = ldb; ar.ldc = ldc;
        ar.alpha = alpha; ar.beta = beta; ar.rk = rk; ar.maxrk = maxrk; ar.acc = acc;
        ar.tile0 = t0; ar.ntiles = std::min(chunk, ntiles - t0); ar.f32 = f32;
        ar.phase = phase; ar.rank_out = (phase == 1) ? d_newrank : d_Crk;
        ar.Bd = d_Bdense; ar.ldbd = ldb; ar.scan_from = scan_from;
        int rc;
        switch (N) {
            case 32: rc = launch_chunk_n<32>(ctx, ar, ws, st); break;
            case 64: rc = launch_chunk_n<64>(ctx, ar, ws, st); break;
            case 128: rc = launch_chunk_n<128>(ctx, ar, ws, st); break;
            default: rc = launch_chunk_n<256>(ctx, ar, ws, st); break;
        }
        if (rc != HCORE_B200_OK) return rc;
    }
    if (phase == 1) {
        ctx->split.pending = true; ctx->split.ntiles = ntiles; ctx->split.nb = nb; ctx->split.f32 = f32;
        ctx->split.maxrk = maxrk;
    }
    return HCORE_B200_OK;
}

// Second half of a split-phase update: CU = Q_u [U_r; 0], CV = Q_v [W; 0] written to the caller's (possibly
// re-allocated) factor arrays, Crk <- new rank.  The reflectors are the ones phase 1 left in the workspace.
static int tlr_batched_apply(hcore_b200_handle_t ctx, int ntiles, int nb, double *const *d_CU, double *const *d_CV,
                             int32_t *d_Crk, int ldc, void *stream) {
    if (!ctx || ldc < nb) return HCORE_B200_ERR_ARG;
    if (!ctx->split.pending || ctx->split.ntiles != ntiles || ctx->split.nb != nb) {
        ctx->err = "HCORE_dgemm_batched_ormqr without a matching HCORE_dgemm_batched_qr_svd on this handle";
        return HCORE_B200_ERR_ARG;
    }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TlrArgs ar{};
    ar.CU = d_CU; ar.CV = d_CV; ar.Crk = d_Crk; ar.rank_out = d_Crk; ar.status = nullptr;
    ar.nb = nb; ar.ldc = ldc; ar.tile0 = 0; ar.ntiles = ntiles; ar.f32 = ctx->split.f32; ar.phase = 2;
    ar.maxrk = ctx->split.maxrk;
    int rc;
    switch (ctx->last_N) {
        case 32: rc = launch_chunk_n<32>(ctx, ar, ctx->last_ws, (cudaStream_t)stream); break;
        case 64: rc = launch_chunk_n<64>(ctx, ar, ctx->last_ws, (cudaStream_t)stream); break;
        case 128: rc = launch_chunk_n<128>(ctx, ar, ctx->last_ws, (cudaStream_t)stream); break;
        default: rc = launch_chunk_n<256>(ctx, ar, ctx->last_ws, (cudaStream_t)stream); break;
    }
    ctx->split.pending = false;
    return rc;
}

// strided factor arrays -> pointer arrays (on the device), then the pointer-array path
static int tlr_batched_strided(hcore_b200_handle_t ctx, int ntiles, int nb, double alpha,
                        const void *d_AU, const void *d_AV, const int32_t *d_Ark, int lda, int64_t strideA,
                        const void *d_BU, const void *d_BV, const int32_t *d_Brk, int ldb, int64_t strideB,
                        double beta, void *d_CU, void *d_CV, int32_t *d_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, double acc, int32_t *d_status, void *stream, int f32,
                        int phase = 0, int32_t *d_newrank = nullptr, int scan_from = 1) {
    if (!ctx || ntiles < 0) return HCORE_B200_ERR_ARG;
    if (ntiles == 0) return HCORE_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(ctx, ctx->ptrs.reserve(6ull * ntiles * sizeof(double *)));
    const double **pp = (const double **)ctx->ptrs.p;
    const void *bases[6] = {d_AU, d_AV, d_BU, d_BV, d_CU, d_CV};
    const int64_t strides[6] = {strideA, strideA, strideB, strideB, strideC, strideC};
    const int es = f32 ? 4 : 8;
    const int thr = 256, blk = (ntiles + thr - 1) / thr;
    for (int i = 0; i < 6; ++i) {
        fill_ptrs_kernel<<<blk, thr, 0, st>>>(pp + (size_t)i * ntiles, (const char *)bases[i], strides[i] * es, ntiles);
        ctx->launches++;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return tlr_batched(ctx, ntiles, nb, alpha, pp, pp + ntiles, d_Ark, lda, pp + 2 * (size_t)ntiles,
                       pp + 3 * (size_t)ntiles, d_Brk, ldb, beta, (double *const *)(pp + 4 * (size_t)ntiles),
                       (double *const *)(pp + 5 * (size_t)ntiles), d_Crk, ldc, rk, maxrk, acc, d_status, stream, f32,
                       phase, d_newrank, nullptr, scan_from);
}

extern "C" {

int HCORE_dgemm_batched(hcore_b200_handle_t ctx, int ntiles, int nb, double alpha,
                        const double *const *d_AU, const double *const *d_AV, const int32_t *d_Ark, int lda,
                        const double *const *d_BU, const double *const *d_BV, const int32_t *d_Brk, int ldb,
                        double beta, double *const *d_CU, double *const *d_CV, int32_t *d_Crk, int ldc,
                        int rk, int maxrk, double acc, int32_t *d_status, void *stream) {
    return tlr_batched(ctx, ntiles, nb, alpha, d_AU, d_AV, d_Ark, lda, d_BU, d_BV, d_Brk, ldb, beta, d_CU, d_CV, d_Crk, ldc,
                       rk, maxrk, acc, d_status, stream, 0);
}

// ---- split-phase update (reference HCORE_dgemm_qr_svd / HCORE_dgemm_ormqr, src/compute/hcore_dgemm.c:1312-1699) ----
int HCORE_dgemm_batched_qr_svd(hcore_b200_handle_t ctx, int ntiles, int nb, double alpha,
                        const double *const *d_AU, const double *const *d_AV, const int32_t *d_Ark, int lda,
                        const double *const *d_BU, const double *const *d_BV, const int32_t *d_Brk, int ldb,
                        double beta, const double *const *d_CU, const double *const *d_CV, const int32_t *d_Crk, int ldc,
                        int rk, int maxrk, double acc, int32_t *d_newrank, int32_t *d_status, void *stream) {
    return tlr_batched(ctx, ntiles, nb, alpha, d_AU, d_AV, d_Ark, lda, d_BU, d_BV, d_Brk, ldb, beta,
                       (double *const *)d_CU, (double *const *)d_CV, const_cast<int32_t *>(d_Crk), ldc, rk, maxrk, acc,
                       d_status, stream, 0, 1, d_newrank);
}

int HCORE_dgemm_batched_ormqr(hcore_b200_handle_t ctx, int ntiles, int nb, double *const *d_CU, double *const *d_CV,
                        int32_t *d_Crk, int ldc, void *stream) {
    return tlr_batched_apply(ctx, ntiles, nb, d_CU, d_CV, d_Crk, ldc, stream);
}

int HCORE_dgemm_batched_qr_svd_strided(hcore_b200_handle_t ctx, int ntiles, int nb, double alpha,
                        const double *d_AU, const double *d_AV, const int32_t *d_Ark, int lda, int64_t strideA,
                        const double *d_BU, const double *d_BV, const int32_t *d_Brk, int ldb, int64_t strideB,
                        double beta, const double *d_CU, const double *d_CV, const int32_t *d_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, double acc, int32_t *d_newrank, int32_t *d_status, void *stream) {
    return tlr_batched_strided(ctx, ntiles, nb, alpha, d_AU, d_AV, d_Ark, lda, strideA, d_BU, d_BV, d_Brk, ldb, strideB, beta,
                               const_cast<double *>(d_CU), const_cast<double *>(d_CV), const_cast<int32_t *>(d_Crk), ldc,
                               strideC, rk, maxrk, acc, d_status, stream, 0, 1, d_newrank);
}

int HCORE_dgemm_batched_ormqr_strided(hcore_b200_handle_t ctx, int ntiles, int nb, double *d_CU, double *d_CV,
                        int32_t *d_Crk, int ldc, int64_t strideC, void *stream) {
    if (!ctx || ntiles <= 0) return ntiles == 0 ? HCORE_B200_OK : HCORE_B200_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    // the pointer table of phase 1 is still in ctx->ptrs (6 x ntiles): slots 4 and 5 are re-filled
    CUDA_TRY(ctx, ctx->ptrs.reserve(6ull * ntiles * sizeof(double *)));
    const double **pp = (const double **)ctx->ptrs.p;
    const int thr = 256, blk = (ntiles + thr - 1) / thr;
    fill_ptrs_kernel<<<blk, thr, 0, st>>>(pp + 4 * (size_t)ntiles, (const char *)d_CU, strideC * 8, ntiles);
    fill_ptrs_kernel<<<blk, thr, 0, st>>>(pp + 5 * (size_t)ntiles, (const char *)d_CV, strideC * 8, ntiles);
    ctx->launches += 2;
    return tlr_batched_apply(ctx, ntiles, nb, (double *const *)(pp + 4 * (size_t)ntiles),
                             (double *const *)(pp + 5 * (size_t)ntiles), d_Crk, ldc, stream);
}

// ---- B-dense variant (reference HCORE_dgemm_qr_svd_b_dense, src/compute/hcore_dgemm.c:1022-1303) ---------
int HCORE_dgemm_batched_b_dense(hcore_b200_handle_t ctx, int ntiles, int nb, double alpha,
                        const double *const *d_AU, const double *const *d_AV, const int32_t *d_Ark, int lda,
                        const double *const *d_B, int ldb, double beta,
                        double *const *d_CU, double *const *d_CV, int32_t *d_Crk, int ldc,
                        int rk, int maxrk, double acc, int32_t *d_status, void *stream) {
    if (!d_B) return HCORE_B200_ERR_ARG;
    return tlr_batched(ctx, ntiles, nb, alpha, d_AU, d_AV, d_Ark, lda, nullptr, nullptr, d_Ark, ldb, beta, d_CU, d_CV, d_Crk,
                       ldc, rk, maxrk, acc, d_status, stream, 0, 0, nullptr, d_B);
}

int HCORE_dgemm_batched_qr_svd_b_dense(hcore_b200_handle_t ctx, int ntiles, int nb, double alpha,
                        const double *const *d_AU, const double *const *d_AV, const int32_t *d_Ark, int lda,
                        const double *const *d_B, int ldb, double beta,
                        const double *const *d_CU, const double *const *d_CV, const int32_t *d_Crk, int ldc,
                        int rk, int maxrk, double acc, int32_t *d_newrank, int32_t *d_status, void *stream) {
    if (!d_B) return HCORE_B200_ERR_ARG;
    return tlr_batched(ctx, ntiles, nb, alpha, d_AU, d_AV, d_Ark, lda, nullptr, nullptr, d_Ark, ldb, beta,
                       (double *const *)d_CU, (double *const *)d_CV, const_cast<int32_t *>(d_Crk), ldc, rk, maxrk, acc,
                       d_status, stream, 0, 1, d_newrank, d_B);
}

int HCORE_dgemm_batched_b_dense_strided(hcore_b200_handle_t ctx, int ntiles, int nb, double alpha,
                        const double *d_AU, const double *d_AV, const int32_t *d_Ark, int lda, int64_t strideA,
                        const double *d_B, int ldb, int64_t strideB, double beta,
                        double *d_CU, double *d_CV, int32_t *d_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, double acc, int32_t *d_status, void *stream) {
    if (!ctx || ntiles < 0 || !d_B) return HCORE_B200_ERR_ARG;
    if (ntiles == 0) return HCORE_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(ctx, ctx->ptrs.reserve(6ull * ntiles * sizeof(double *)));
    const double **pp = (const double **)ctx->ptrs.p;
    const void *bases[5] = {d_AU, d_AV, d_B, d_CU, d_CV};
    const int64_t strides[5] = {strideA, strideA, strideB, strideC, strideC};
    const int thr = 256, blk = (ntiles + thr - 1) / thr;
    for (int i = 0; i < 5; ++i) {
        fill_ptrs_kernel<<<blk, thr, 0, st>>>(pp + (size_t)i * ntiles, (const char *)bases[i], strides[i] * 8, ntiles);
        ctx->launches++;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return HCORE_dgemm_batched_b_dense(ctx, ntiles, nb, alpha, pp, pp + ntiles, d_Ark, lda, pp + 2 * (size_t)ntiles, ldb, beta,
                                       (double *const *)(pp + 3 * (size_t)ntiles), (double *const *)(pp + 4 * (size_t)ntiles),
                                       d_Crk, ldc, rk, maxrk, acc, d_status, stream);
}

int HCORE_sgemm_batched(hcore_b200_handle_t ctx, int ntiles, int nb, float alpha,
                        const float *const *d_AU, const float *const *d_AV, const int32_t *d_Ark, int lda,
                        const float *const *d_BU, const float *const *d_BV, const int32_t *d_Brk, int ldb,
                        float beta, float *const *d_CU, float *const *d_CV, int32_t *d_Crk, int ldc,
                        int rk, int maxrk, float acc, int32_t *d_status, void *stream) {
    return tlr_batched(ctx, ntiles, nb, alpha, (const double *const *)d_AU, (const double *const *)d_AV, d_Ark, lda,
                       (const double *const *)d_BU, (const double *const *)d_BV, d_Brk, ldb, beta, (double *const *)d_CU,
                       (double *const *)d_CV, d_Crk, ldc, rk, maxrk, acc, d_status, stream, 1);
}

int HCORE_dgemm_batched_strided(hcore_b200_handle_t ctx, int ntiles, int nb, double alpha,
                        const double *d_AU, const double *d_AV, const int32_t *d_Ark, int lda, int64_t strideA,
                        const double *d_BU, const double *d_BV, const int32_t *d_Brk, int ldb, int64_t strideB,
                        double beta, double *d_CU, double *d_CV, int32_t *d_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, double acc, int32_t *d_status, void *stream) {
    return tlr_batched_strided(ctx, ntiles, nb, alpha, d_AU, d_AV, d_Ark, lda, strideA, d_BU, d_BV, d_Brk, ldb, strideB, beta,
                               d_CU, d_CV, d_Crk, ldc, strideC, rk, maxrk, acc, d_status, stream, 0);
}

int HCORE_sgemm_batched_strided(hcore_b200_handle_t ctx, int ntiles, int nb, float alpha,
                        const float *d_AU, const float *d_AV, const int32_t *d_Ark, int lda, int64_t strideA,
                        const float *d_BU, const float *d_BV, const int32_t *d_Brk, int ldb, int64_t strideB,
                        float beta, float *d_CU, float *d_CV, int32_t *d_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, float acc, int32_t *d_status, void *stream) {
    return tlr_batched_strided(ctx, ntiles, nb, alpha, d_AU, d_AV, d_Ark, lda, strideA, d_BU, d_BV, d_Brk, ldb, strideB, beta,
                               d_CU, d_CV, d_Crk, ldc, strideC, rk, maxrk, acc, d_status, stream, 1);
}

}  // extern "C"

// =====================================================================================================
// Host-buffer forms: every array is a HOST array (pinned memory gives full overlap).  Slices of the batch travel
// through kHostBufs device staging buffers on THREE streams -- host-to-device copies, kernels, device-to-host copies --
// chained by events only: the calling thread never blocks inside the loop, so the copy engines of both directions
// and the SMs stay busy at the same time (PCIe is full duplex).  The number of C columns that come back is fixed per
// slice from what the host knows (the stack width Ark + Crk bounds the new rank), so no rank has to reach the host
// before the copy is enqueued; the columns between a tile's new rank and that bound are zero-filled on the device.
// =====================================================================================================
namespace {

// zero the byte range [lo[t], hi) of two strided arrays per tile (columns past the new rank)
__global__ void zero_tail_kernel(char *CU, char *CV, size_t stride_bytes, size_t col_bytes, const int *rank_new,
                                 const int *status, const int *rank_old, int ncopy, int ntiles) {
    const int t = blockIdx.y;
    if (t >= ntiles) return;
    const int r = (status && status[t] != HCORE_B200_OK) ? rank_old[t] : rank_new[t];
    if (r >= ncopy) return;
    const size_t lo = (size_t)r * col_bytes, n16 = ((size_t)(ncopy - r) * col_bytes) >> 4;
    for (int a = 0; a < 2; ++a) {
        uint4 *dst = reinterpret_cast<uint4 *>((a ? CV : CU) + (size_t)t * stride_bytes + lo);
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
            dst[i] = make_uint4(0u, 0u, 0u, 0u);
    }
}

template <class FIn, class FComp, class FOut>
int host_pipeline(hcore_b200_context *ctx, HostPipe &hp, int ntiles, int hc, size_t stage_bytes, FIn &&in, FComp &&comp,
                  FOut &&out) {
    if (!hp.ready) {
        for (cudaStream_t *st : {&hp.s_in, &hp.s_comp, &hp.s_out}) CUDA_TRY(ctx, cudaStreamCreateWithFlags(st, cudaStreamNonBlocking));
        for (int b = 0; b < kHostBufs; ++b) {
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&hp.in_done[b], cudaEventDisableTiming));
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&hp.comp_done[b], cudaEventDisableTiming));
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&hp.out_done[b], cudaEventDisableTiming));
        }
        hp.ready = true;
    }
    const int nb_used = std::min(kHostBufs, (ntiles + hc - 1) / hc);
    for (int b = 0; b < nb_used; ++b) {
        CUDA_TRY(ctx, ctx->hstage[b].reserve(stage_bytes * hc));
        CUDA_TRY(ctx, ctx->hstage_ints[b].reserve((size_t)hc * 5 * sizeof(int)));
    }
    int slice = 0;
    for (int t0 = 0; t0 < ntiles; t0 += hc, ++slice) {
        const int b = slice % kHostBufs, cnt = std::min(hc, ntiles - t0);
        if (slice >= kHostBufs) CUDA_TRY(ctx, cudaStreamWaitEvent(hp.s_in, hp.out_done[b], 0));   // staging buffer b is free again
        CUDA_TRY(ctx, in(t0, cnt, b, hp.s_in));
        CUDA_TRY(ctx, cudaEventRecord(hp.in_done[b], hp.s_in));
        CUDA_TRY(ctx, cudaStreamWaitEvent(hp.s_comp, hp.in_done[b], 0));
        const int rc = comp(t0, cnt, b, hp.s_comp);
        if (rc != HCORE_B200_OK) { cudaDeviceSynchronize(); return rc; }
        CUDA_TRY(ctx, cudaEventRecord(hp.comp_done[b], hp.s_comp));
        CUDA_TRY(ctx, cudaStreamWaitEvent(hp.s_out, hp.comp_done[b], 0));
        CUDA_TRY(ctx, out(t0, cnt, b, hp.s_out));
        CUDA_TRY(ctx, cudaEventRecord(hp.out_done[b], hp.s_out));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(hp.s_out));
    CUDA_TRY(ctx, cudaStreamSynchronize(hp.s_comp));
    CUDA_TRY(ctx, cudaStreamSynchronize(hp.s_in));
    CUDA_TRY(ctx, cudaGetLastError());
    return HCORE_B200_OK;
}

inline cudaError_t first_error(std::initializer_list<cudaError_t> es) {
    for (cudaError_t e : es) if (e != cudaSuccess) return e;
    return cudaSuccess;
}

// kind: 0 = double, 1 = float (HCORE_sgemm), 2 = complex double (HCORE_zgemm); es = element size in bytes
int tlr_host(hcore_b200_handle_t ctx, int kind, int ntiles, int nb, const void *alpha, const void *h_AU, const void *h_AV,
             const int32_t *h_Ark, int lda, int64_t strideA, const void *h_BU, const void *h_BV, const int32_t *h_Brk, int ldb,
             int64_t strideB, const void *beta, void *h_CU, void *h_CV, int32_t *h_Crk, int ldc, int64_t strideC, int rk, int maxrk,
             double acc, int32_t *h_status) {
    if (!ctx || ntiles < 0 || nb <= 0 || maxrk <= 0) return HCORE_B200_ERR_ARG;
    if (ntiles == 0) return HCORE_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t es = kind == 0 ? 8 : (kind == 1 ? 4 : 16);
    // The path is PCIe-bound (3.1 MB in, 2.1 MB out per tile against ~17 us of GPU time at nb = 1024, rank 64): what a
    // slice costs is its share of the pipeline fill (the first H2D) and drain (the last kernels + D2H).  One wave of the
    // SVD kernel (148 tiles) per slice measured best: 64-tile slices lose 4 % (14.1 k vs 14.7 k tiles/s, 2048 tiles).
    const int hc = std::min(ntiles, ctx->host_slice > 0 ? ctx->host_slice : std::max(ctx->num_sms, 64));
    const size_t perA = (size_t)strideA * es, perB = (size_t)strideB * es, perC = (size_t)strideC * es;
    const size_t per_tile = 2 * perA + 2 * perB + 2 * perC;
    auto carve = [&](int b, char *&dAU, char *&dAV, char *&dBU, char *&dBV, char *&dCU, char *&dCV, int *&di) {
        char *base = (char *)ctx->hstage[b].p;
        dAU = base; dAV = base + perA * hc; dBU = base + 2 * perA * hc; dBV = dBU + perB * hc;
        dCU = base + 2 * perA * hc + 2 * perB * hc; dCV = dCU + perC * hc;
        di = (int *)ctx->hstage_ints[b].p;
    };
    std::vector<int> ncopy_of;       // C columns that travel back, per slice
    auto in = [&](int t0, int cnt, int b, cudaStream_t st) {
        char *dAU, *dAV, *dBU, *dBV, *dCU, *dCV; int *di;
        carve(b, dAU, dAV, dBU, dBV, dCU, dCV, di);
        // only the columns that carry data cross PCIe: the first max(Crk) columns of CU / CV go in
        int cmax = 1;
        for (int t = 0; t < cnt; ++t) cmax = std::max(cmax, (int)h_Crk[t0 + t]);
        const size_t win = std::min(perC, (size_t)ldc * cmax * es);
        return first_error({
            cudaMemcpyAsync(dAU, (const char *)h_AU + (size_t)t0 * perA, perA * cnt, cudaMemcpyHostToDevice, st),
            cudaMemcpyAsync(dAV, (const char *)h_AV + (size_t)t0 * perA, perA * cnt, cudaMemcpyHostToDevice, st),
            cudaMemcpyAsync(dBU, (const char *)h_BU + (size_t)t0 * perB, perB * cnt, cudaMemcpyHostToDevice, st),
            cudaMemcpyAsync(dBV, (const char *)h_BV + (size_t)t0 * perB, perB * cnt, cudaMemcpyHostToDevice, st),
            cudaMemcpy2DAsync(dCU, perC, (const char *)h_CU + (size_t)t0 * perC, perC, win, cnt, cudaMemcpyHostToDevice, st),
            cudaMemcpy2DAsync(dCV, perC, (const char *)h_CV + (size_t)t0 * perC, perC, win, cnt, cudaMemcpyHostToDevice, st),
            cudaMemcpyAsync(di, h_Ark + t0, sizeof(int) * cnt, cudaMemcpyHostToDevice, st),
            cudaMemcpyAsync(di + hc, h_Brk + t0, sizeof(int) * cnt, cudaMemcpyHostToDevice, st),
            cudaMemcpyAsync(di + 2 * hc, h_Crk + t0, sizeof(int) * cnt, cudaMemcpyHostToDevice, st)});
    };
    // the ranks that go in are read again when the results come back: keep a private copy (h_Crk is overwritten)
    std::vector<int32_t> crk_in(h_Crk, h_Crk + ntiles);
    auto comp = [&](int t0, int cnt, int b, cudaStream_t st) -> int {
        char *dAU, *dAV, *dBU, *dBV, *dCU, *dCV; int *di;
        carve(b, dAU, dAV, dBU, dBV, dCU, dCV, di);
        int *dArk = di, *dBrk = di + hc, *dCrk = di + 2 * hc, *dOld = di + 3 * hc, *dSt = di + 4 * hc;
        int nmax = 1;
        for (int t = 0; t < cnt; ++t) nmax = std::max(nmax, (int)h_Ark[t0 + t] + (int)crk_in[t0 + t]);
        const int ncopy = (int)std::min<int64_t>({(int64_t)nmax, (int64_t)maxrk, strideC / ldc});
        ncopy_of.push_back(ncopy);
        // a failed tile keeps its old rank: a copy of the ranks is taken before the update overwrites them
        int rc;
        CUDA_TRY(ctx, cudaMemcpyAsync(dOld, dCrk, sizeof(int) * cnt, cudaMemcpyDeviceToDevice, st));
        if (kind == 2) {
            const hcore_b200_dcomplex *al = (const hcore_b200_dcomplex *)alpha, *be = (const hcore_b200_dcomplex *)beta;
            rc = HCORE_zgemm_batched_strided(ctx, cnt, nb, *al, dAU, dAV, dArk, lda, strideA, dBU, dBV, dBrk, ldb, strideB, *be, dCU,
                                             dCV, dCrk, ldc, strideC, rk, maxrk, acc, dSt, st);
        } else {
            rc = tlr_batched_strided(ctx, cnt, nb, *(const double *)alpha, dAU, dAV, dArk, lda, strideA, dBU, dBV, dBrk, ldb, strideB,
                                     *(const double *)beta, dCU, dCV, dCrk, ldc, strideC, rk, maxrk, acc, dSt, st, kind == 1);
        }
        if (rc != HCORE_B200_OK) return rc;
        zero_tail_kernel<<<dim3(8, cnt), 256, 0, st>>>(dCU, dCV, perC, (size_t)ldc * es, dCrk, dSt, dOld, ncopy, cnt);
        ctx->launches++;
        return HCORE_B200_OK;
    };
    size_t slice_idx = 0;
    auto out = [&](int t0, int cnt, int b, cudaStream_t st) {
        char *dAU, *dAV, *dBU, *dBV, *dCU, *dCV; int *di;
        carve(b, dAU, dAV, dBU, dBV, dCU, dCV, di);
        const size_t wout = std::min(perC, (size_t)ldc * ncopy_of[slice_idx++] * es);
        return first_error({
            cudaMemcpy2DAsync((char *)h_CU + (size_t)t0 * perC, perC, dCU, perC, wout, cnt, cudaMemcpyDeviceToHost, st),
            cudaMemcpy2DAsync((char *)h_CV + (size_t)t0 * perC, perC, dCV, perC, wout, cnt, cudaMemcpyDeviceToHost, st),
            cudaMemcpyAsync(h_Crk + t0, di + 2 * hc, sizeof(int) * cnt, cudaMemcpyDeviceToHost, st),
            h_status ? cudaMemcpyAsync(h_status + t0, di + 4 * hc, sizeof(int) * cnt, cudaMemcpyDeviceToHost, st) : cudaSuccess});
    };
    return host_pipeline(ctx, ctx->hpipe, ntiles, hc, per_tile, in, comp, out);
}

}  // namespace

extern "C" {

int HCORE_dgemm_batched_host(hcore_b200_handle_t ctx, int ntiles, int nb, double alpha,
                        const double *h_AU, const double *h_AV, const int32_t *h_Ark, int lda, int64_t strideA,
                        const double *h_BU, const double *h_BV, const int32_t *h_Brk, int ldb, int64_t strideB,
                        double beta, double *h_CU, double *h_CV, int32_t *h_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, double acc, int32_t *h_status) {
    return tlr_host(ctx, 0, ntiles, nb, &alpha, h_AU, h_AV, h_Ark, lda, strideA, h_BU, h_BV, h_Brk, ldb, strideB, &beta, h_CU, h_CV,
                    h_Crk, ldc, strideC, rk, maxrk, acc, h_status);
}

int HCORE_sgemm_batched_host(hcore_b200_handle_t ctx, int ntiles, int nb, float alpha,
                        const float *h_AU, const float *h_AV, const int32_t *h_Ark, int lda, int64_t strideA,
                        const float *h_BU, const float *h_BV, const int32_t *h_Brk, int ldb, int64_t strideB,
                        float beta, float *h_CU, float *h_CV, int32_t *h_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, float acc, int32_t *h_status) {
    const double al = alpha, be = beta;
    return tlr_host(ctx, 1, ntiles, nb, &al, h_AU, h_AV, h_Ark, lda, strideA, h_BU, h_BV, h_Brk, ldb, strideB, &be, h_CU, h_CV,
                    h_Crk, ldc, strideC, rk, maxrk, acc, h_status);
}

int HCORE_zgemm_batched_host(hcore_b200_handle_t ctx, int ntiles, int nb, hcore_b200_dcomplex alpha,
                        const void *h_AU, const void *h_AV, const int32_t *h_Ark, int lda, int64_t strideA,
                        const void *h_BU, const void *h_BV, const int32_t *h_Brk, int ldb, int64_t strideB,
                        hcore_b200_dcomplex beta, void *h_CU, void *h_CV, int32_t *h_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, double acc, int32_t *h_status) {
    return tlr_host(ctx, 2, ntiles, nb, &alpha, h_AU, h_AV, h_Ark, lda, strideA, h_BU, h_BV, h_Brk, ldb, strideB, &beta, h_CU, h_CV,
                    h_Crk, ldc, strideC, rk, maxrk, acc, h_status);
}

// ---- complex double (HCORE_zgemm) ------------------------------------------------------------
}  // extern "C"
static int zgemm_launch(hcore_b200_handle_t ctx, int ntiles, int nb, hcore_b200_dcomplex alpha,
                        const void *d_AU, const void *d_AV, const int32_t *d_Ark, int lda, int64_t strideA,
                        const void *d_BU, const void *d_BV, const int32_t *d_Brk, int ldb, int64_t strideB,
                        hcore_b200_dcomplex beta, void *d_CU, void *d_CV, int32_t *d_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, double acc, int32_t *d_status, void *stream, const void *const *d_ptrs = nullptr) {
    if (!ctx || ntiles < 0 || nb <= 0 || maxrk <= 0 || lda < nb || ldb < nb || ldc < nb) return HCORE_B200_ERR_ARG;
    if (ntiles == 0) return HCORE_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int N = pick_width(maxrk, HCORE_B200_MAX_STACK_Z);
    const size_t per = (2ull * nb * N + 2ull * N * (1 + kZB) + 3ull * N * N) * sizeof(zc) + N * 8 + 16;
    int chunk = ctx->chunk_limit > 0 ? ctx->chunk_limit : (int)std::min<size_t>(ctx->num_sms * 4, (4ull << 30) / per);
    chunk = std::max(1, std::min(std::min(chunk, ntiles), 32767));     // tile index in gridDim.z (apply) / 2 x tiles in gridDim.x (QR)
    CUDA_TRY(ctx, ctx->zpanel.reserve(2ull * chunk * nb * N * sizeof(zc)));
    CUDA_TRY(ctx, ctx->ztau.reserve(2ull * chunk * N * (1 + kZB) * sizeof(zc)));     // tau + the 8 x 8 compact-WY factors
    CUDA_TRY(ctx, ctx->zT.reserve((size_t)chunk * N * N * sizeof(zc)));
    CUDA_TRY(ctx, ctx->zY.reserve(2ull * chunk * N * N * sizeof(zc)));
    CUDA_TRY(ctx, ctx->zsigma.reserve((size_t)chunk * N * 8));
    CUDA_TRY(ctx, ctx->zints.reserve((size_t)chunk * 4 * sizeof(int)));
    ZWorkspace ws{};
    ws.panel = (zc *)ctx->zpanel.p; ws.tau = (zc *)ctx->ztau.p; ws.tfac = ws.tau + 2ull * chunk * N; ws.T = (zc *)ctx->zT.p; ws.Y = (zc *)ctx->zY.p;
    ws.sigma = (double *)ctx->zsigma.p;
    ws.ncols = (int *)ctx->zints.p; ws.newrank = ws.ncols + chunk; ws.st = ws.newrank + chunk; ws.sweeps = ws.st + chunk;
    ws.chunk = chunk; ws.N = N;
    // apply kernel: W column buffers + 2 reflector buffers of nb complex each; W = 8, 4 or 2 warps
    int apw = 8;
    while (apw > 2 && (size_t)(apw + 4) * nb * sizeof(zc) > 227 * 1024) apw >>= 1;      // W columns + two pairs of reflectors
    const size_t sm_qr = (2ull * nb + 40) * sizeof(double), sm_ap = (size_t)(apw + 4) * nb * sizeof(zc);
    if (sm_ap > 227 * 1024) { ctx->err = "HCORE_zgemm: nb too large for the apply kernel's shared memory"; return HCORE_B200_ERR_ARG; }
    const size_t sm_qrb = (2ull * kZB * (nb + 1) + 32 + 2 * kZB + 6 * kZB * kZB + 2 * 16 * 64) * sizeof(double);
    CUDA_TRY(ctx, cudaFuncSetAttribute(ztlr_qr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_qr));
    if (nb <= 32 * kZRows)
        CUDA_TRY(ctx, cudaFuncSetAttribute(ztlr_qr_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_qrb));
    CUDA_TRY(ctx, cudaFuncSetAttribute(ztlr_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_ap));
    static const bool z_apply_scalar = getenv("HCORE_B200_Z_APPLY_SCALAR") != nullptr;   // measurement switch: one-column-per-warp application
    for (int t0 = 0; t0 < ntiles; t0 += chunk) {
        ZArgs ar{};
        ar.AU = (const zc *)d_AU; ar.AV = (const zc *)d_AV; ar.BU = (const zc *)d_BU; ar.BV = (const zc *)d_BV;
        ar.CU = (zc *)d_CU; ar.CV = (zc *)d_CV;
        ar.strideA = strideA; ar.strideB = strideB; ar.strideC = strideC;
        ar.Ark = d_Ark; ar.Brk = d_Brk; ar.Crk = d_Crk; ar.status = d_status;
        ar.nb = nb; ar.lda = lda; ar.ldb = ldb; ar.ldc = ldc;
        ar.alpha = make_double2(alpha.re, alpha.im); ar.beta = make_double2(beta.re, beta.im);
        ar.rk = rk; ar.maxrk = maxrk; ar.acc = acc; ar.tile0 = t0; ar.ntiles = std::min(chunk, ntiles - t0);
        ar.ptrs = (const zc *const *)d_ptrs; ar.ptr_stride = ntiles;
        ztlr_stack_kernel<<<ar.ntiles, 256, 0, st>>>(ar, ws);
        if (nb <= 32 * kZRows) ztlr_qr_blocked_kernel<<<2 * ar.ntiles, 512, sm_qrb, st>>>(ar, ws);
        else ztlr_qr_kernel<<<2 * ar.ntiles, 512, sm_qr, st>>>(ar, ws);
        if (N <= 128) ztlr_svd_kernel<8><<<ar.ntiles, 512, 0, st>>>(ar, ws);
        else ztlr_svd_kernel<16><<<ar.ntiles, 256, 0, st>>>(ar, ws);
        if (nb <= 32 * kZRows && !z_apply_scalar) ztlr_apply_wy_kernel<<<dim3((N + 7) / 8, 2, ar.ntiles), 512, 0, st>>>(ar, ws);
        else ztlr_apply_kernel<<<dim3((N + apw - 1) / apw, 2, ar.ntiles), 32 * apw, sm_ap, st>>>(ar, ws);
        ctx->launches += 4;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return HCORE_B200_OK;
}

extern "C" {
int HCORE_zgemm_batched_strided(hcore_b200_handle_t ctx, int ntiles, int nb, hcore_b200_dcomplex alpha,
                        const void *d_AU, const void *d_AV, const int32_t *d_Ark, int lda, int64_t strideA,
                        const void *d_BU, const void *d_BV, const int32_t *d_Brk, int ldb, int64_t strideB,
                        hcore_b200_dcomplex beta, void *d_CU, void *d_CV, int32_t *d_Crk, int ldc, int64_t strideC,
                        int rk, int maxrk, double acc, int32_t *d_status, void *stream) {
    return zgemm_launch(ctx, ntiles, nb, alpha, d_AU, d_AV, d_Ark, lda, strideA, d_BU, d_BV, d_Brk, ldb, strideB, beta, d_CU, d_CV,
                        d_Crk, ldc, strideC, rk, maxrk, acc, d_status, stream);
}

// pointer-array form: six device arrays of per-tile pointers (complex arrays, interleaved re / im)
int HCORE_zgemm_batched(hcore_b200_handle_t ctx, int ntiles, int nb, hcore_b200_dcomplex alpha,
                        const void *const *d_AU, const void *const *d_AV, const int32_t *d_Ark, int lda,
                        const void *const *d_BU, const void *const *d_BV, const int32_t *d_Brk, int ldb,
                        hcore_b200_dcomplex beta, void *const *d_CU, void *const *d_CV, int32_t *d_Crk, int ldc,
                        int rk, int maxrk, double acc, int32_t *d_status, void *stream) {
    if (!ctx || ntiles < 0) return HCORE_B200_ERR_ARG;
    if (ntiles == 0) return HCORE_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    // the six tables are gathered into one [6][ntiles] table the kernels index (device-to-device, 48 bytes per tile)
    CUDA_TRY(ctx, ctx->ptrs.reserve(6ull * ntiles * sizeof(void *)));
    const void *const *src[6] = {d_AU, d_AV, d_BU, d_BV, (const void *const *)d_CU, (const void *const *)d_CV};
    for (int q = 0; q < 6; ++q)
        CUDA_TRY(ctx, cudaMemcpyAsync((void **)ctx->ptrs.p + (size_t)q * ntiles, src[q], (size_t)ntiles * sizeof(void *),
                                      cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return zgemm_launch(ctx, ntiles, nb, alpha, nullptr, nullptr, d_Ark, lda, 0, nullptr, nullptr, d_Brk, ldb, 0, beta, nullptr, nullptr,
                        d_Crk, ldc, 0, rk, maxrk, acc, d_status, stream, (const void *const *)ctx->ptrs.p);
}
}  // extern "C"

// ---- dense-C updates ----------------------------------------------------------------------
// kcap = upper bound of the B-side rank (columns of the X workspace).  mc = rows contracted by AV^T BV.
typedef CUresult (*tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn resolve_tmap_encode() {      // cuTensorMapEncodeTiled through the runtime: no link-time dependency on libcuda
    static tmap_encode_fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<tmap_encode_fn>(p);
    }();
    return fn;
}
// The persistent streaming kernel (dense_stream_kernel) takes strided batches of whole 128 x 32 blocks whose operands
// the TMA bulk copies can address: everything else goes through dense_update_kernel.  HCORE_B200_DENSE_STREAM=0
// disables it (A/B measurements).
static bool dense_stream_ok(const DenseArgs &ar, int kcap) {
    static const bool enabled = [] { const char *e = getenv("HCORE_B200_DENSE_STREAM"); return !(e && e[0] == '0'); }();
    if (!enabled || ar.pCD || ar.pBU) return false;
    if (ar.nb % kDsRows || ar.ncols % kDsCols || kcap > kDsKmax) return false;
    if ((reinterpret_cast<uintptr_t>(ar.CD) | reinterpret_cast<uintptr_t>(ar.Yop)) & 15) return false;
    if ((ar.ldc | ar.ldy) & 1) return false;
    if ((ar.strideC | ar.strideY) & 1) return false;
    if (!resolve_tmap_encode()) return false;
    return true;
}
static int dense_stream_launch(hcore_b200_context *ctx, const DenseArgs &ar, int cnt, cudaStream_t st) {
    // two 256-thread CTAs per SM, three slots each, contraction lengths up to 32 (longer ones are bound by the FP64 pipe,
    // not by HBM, and the one-CTA-per-SM shape that would hold their X fragments is slower than dense_update_kernel:
    // 166 k vs 174 k calls/s for HCORE_duncompress at k = 64)
    auto *k32 = dense_stream_kernel<256, kDsKmax, 3, 2, 1>;
    constexpr size_t sm32 = DenseStreamSmem<256, kDsKmax, 3, 1>::kBytes;
    if (!ctx->attr_done[6]) {
        ctx->attr_done[6] = true;
        CUDA_TRY(ctx, cudaFuncSetAttribute(k32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm32));
    }
    // tensor map of the batch of dense tiles: rows x columns x tiles, FP64, 128 x 32 x 1 boxes
    alignas(64) CUtensorMap tmap;
    const cuuint64_t dims[3] = {(cuuint64_t)ar.nb, (cuuint64_t)ar.ncols, (cuuint64_t)cnt};
    const cuuint64_t strides[2] = {(cuuint64_t)ar.ldc * 8, (cuuint64_t)(cnt > 1 ? ar.strideC : (long)ar.ldc * ar.ncols) * 8};
    const cuuint32_t box[3] = {kDsRows, kDsCols, 1}, estr[3] = {1, 1, 1};
    const CUresult r = resolve_tmap_encode()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, ar.CD, dims, strides, box, estr,
                                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ctx->err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return HCORE_B200_ERR_CUDA; }
    const long nstrips = (long)cnt * (ar.nb / kDsRows);
    // an ODD number of CTAs: the strips of a CTA then cycle through all row blocks, which balances HCORE_dsyrk (the
    // strip of row block i has 4 (i + 1) blocks to update)
    const int grid = (int)std::min<long>(nstrips, 2l * ctx->num_sms - 1 + ((2 * ctx->num_sms) & 1));
    k32<<<grid, 256, sm32, st>>>(ar, cnt, tmap);
    return HCORE_B200_OK;
}

static int dense_launch(hcore_b200_handle_t ctx, int ntiles, int nb, double scale, int syrk,
                        const double *d_AU, const double *d_AV, const int32_t *d_Ark, int64_t strideA, int lda,
                        const double *d_BU, const double *d_BV, const int32_t *d_Brk, int64_t strideB, int ldb, int mc,
                        double *d_CD, int64_t strideC, int ldc, int kcap, void *stream,
                        const double *const *p_AU = nullptr, const double *const *p_AV = nullptr, const double *const *p_BU = nullptr,
                        const double *const *p_BV = nullptr, double *const *p_CD = nullptr) {
    if (!ctx || ntiles < 0 || nb <= 0 || lda < nb || ldb < nb || ldc < nb || kcap <= 0) return HCORE_B200_ERR_ARG;
    if (ntiles == 0) return HCORE_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int kc8 = round_up(kcap, 8);
    const int m_pad = round_up(nb, kPgRows);
    const size_t sm_prep = PgSmem::kBytes;
    const size_t sm_upd = DenseUpdSmem::kBytes;
    if (!ctx->attr_done[4]) {
        ctx->attr_done[4] = true;
        CUDA_TRY(ctx, cudaFuncSetAttribute(dense_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_prep));
        CUDA_TRY(ctx, cudaFuncSetAttribute(dense_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_upd));
    }
    const size_t per = (size_t)m_pad * kc8 * 8;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>({(size_t)ntiles, (size_t)4096, (4ull << 30) / per}));
    CUDA_TRY(ctx, ctx->X.reserve((size_t)chunk * per));
    for (int t0 = 0; t0 < ntiles; t0 += chunk) {
        const int cnt = std::min(chunk, ntiles - t0);
        DenseArgs ar{};
        ar.AU = d_AU + (size_t)t0 * strideA; ar.AV = d_AV + (size_t)t0 * strideA;
        ar.BV = d_BV + (size_t)t0 * strideB;
        ar.Ark = d_Ark + t0; ar.Brk = d_Brk + t0;
        ar.strideA = strideA; ar.strideB = strideB; ar.lda = lda; ar.ldb = ldb; ar.mc = mc;
        ar.X = (double *)ctx->X.p; ar.scale = scale; ar.m_pad = m_pad; ar.kcap = kc8;
        ar.Xop = ar.X; ar.strideX = (long)m_pad * kc8; ar.ldx = m_pad;
        ar.Yop = d_BU + (size_t)t0 * strideB; ar.strideY = strideB; ar.ldy = ldb;
        ar.Krk = d_Brk + t0;
        ar.CD = d_CD + (size_t)t0 * strideC; ar.strideC = strideC; ar.ldc = ldc;
        ar.nb = nb; ar.ncols = nb; ar.su = 1.0; ar.beta = 1.0; ar.syrk = syrk;
        if (p_CD) { ar.pAU = p_AU + t0; ar.pAV = p_AV + t0; ar.pBU = p_BU + t0; ar.pBV = p_BV + t0; ar.pCD = p_CD + t0; }
        dense_prep_kernel<<<cnt, 256, sm_prep, st>>>(ar);
        if (dense_stream_ok(ar, kc8)) { const int rc = dense_stream_launch(ctx, ar, cnt, st); if (rc != HCORE_B200_OK) return rc; }
        else dense_update_kernel<<<dim3((nb + kDuRows - 1) / kDuRows, (nb + kDuCols - 1) / kDuCols, cnt), 256, sm_upd, st>>>(ar);
        ctx->launches += 2;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return HCORE_B200_OK;
}

// CD (m x n, ld ldc) = beta CD + su * X (m x k, ld ldx) Y^T (n x k, ld ldy): the update kernel alone
static int dense_xyt(hcore_b200_handle_t ctx, int ntiles, int m, int n, double su, const double *d_X, int ldx, int64_t strideX,
                     const double *d_Y, int ldy, int64_t strideY, const int32_t *d_K, double beta, double *d_CD, int ldc,
                     int64_t strideC, cudaStream_t st, int kcap = 1 << 30) {      // kcap: upper bound of the contraction lengths
    if (!ctx->attr_done[4]) {
        ctx->attr_done[4] = true;
        CUDA_TRY(ctx, cudaFuncSetAttribute(dense_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PgSmem::kBytes));
        CUDA_TRY(ctx, cudaFuncSetAttribute(dense_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DenseUpdSmem::kBytes));
    }
    for (int t0 = 0; t0 < ntiles; t0 += 32768) {
        const int cnt = std::min(32768, ntiles - t0);
        DenseArgs ar{};
        ar.Xop = d_X + (size_t)t0 * strideX; ar.strideX = strideX; ar.ldx = ldx;
        ar.Yop = d_Y + (size_t)t0 * strideY; ar.strideY = strideY; ar.ldy = ldy;
        ar.Krk = d_K + t0;
        ar.CD = d_CD + (size_t)t0 * strideC; ar.strideC = strideC; ar.ldc = ldc;
        ar.nb = m; ar.ncols = n; ar.su = su; ar.beta = beta; ar.syrk = 0;
        if (dense_stream_ok(ar, kcap)) { const int rc = dense_stream_launch(ctx, ar, cnt, st); if (rc != HCORE_B200_OK) return rc; }
        else dense_update_kernel<<<dim3((m + kDuRows - 1) / kDuRows, (n + kDuCols - 1) / kDuCols, cnt), 256, DenseUpdSmem::kBytes, st>>>(ar);
        ctx->launches++;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return HCORE_B200_OK;
}

static int rank_cap(int64_t stride, int ld, int fallback) {      // columns a strided factor array can hold
    return stride > 0 ? (int)std::max<int64_t>(1, stride / ld) : fallback;
}

extern "C" {

int HCORE_dgemmcd_batched_strided(hcore_b200_handle_t ctx, int ntiles, int nb,
                        const double *d_AU, const double *d_AV, const int32_t *d_Ark, int64_t strideA,
                        const double *d_BU, const double *d_BV, const int32_t *d_Brk, int64_t strideB,
                        double *d_CD, int64_t strideC, void *stream) {
    return dense_launch(ctx, ntiles, nb, -1.0, 0, d_AU, d_AV, d_Ark, strideA, nb, d_BU, d_BV, d_Brk, strideB, nb, nb, d_CD,
                        strideC, nb, rank_cap(strideB, nb, HCORE_B200_MAX_STACK), stream);
}

int HCORE_dsyrk_batched_strided(hcore_b200_handle_t ctx, int ntiles, int nb, double alpha,
                        const double *d_AU, const double *d_AV, const int32_t *d_Ark, int64_t strideA,
                        double *d_CD, int64_t strideC, void *stream) {
    return dense_launch(ctx, ntiles, nb, alpha, 1, d_AU, d_AV, d_Ark, strideA, nb, d_AU, d_AV, d_Ark, strideA, nb, nb, d_CD,
                        strideC, nb, rank_cap(strideA, nb, HCORE_B200_MAX_STACK), stream);
}

// pointer-array forms (device arrays of per-tile device pointers); kmax = largest rank in the batch (workspace sizing)
int HCORE_dgemmcd_batched(hcore_b200_handle_t ctx, int ntiles, int nb, int kmax,
                        const double *const *d_AU, const double *const *d_AV, const int32_t *d_Ark,
                        const double *const *d_BU, const double *const *d_BV, const int32_t *d_Brk,
                        double *const *d_CD, void *stream) {
    if (!d_AU || !d_AV || !d_BU || !d_BV || !d_CD || kmax <= 0) return HCORE_B200_ERR_ARG;
    return dense_launch(ctx, ntiles, nb, -1.0, 0, nullptr, nullptr, d_Ark, 0, nb, nullptr, nullptr, d_Brk, 0, nb, nb, nullptr, 0, nb, kmax,
                        stream, d_AU, d_AV, d_BU, d_BV, d_CD);
}

int HCORE_dsyrk_batched(hcore_b200_handle_t ctx, int ntiles, int nb, int kmax, double alpha,
                        const double *const *d_AU, const double *const *d_AV, const int32_t *d_Ark,
                        double *const *d_CD, void *stream) {
    if (!d_AU || !d_AV || !d_CD || kmax <= 0) return HCORE_B200_ERR_ARG;
    return dense_launch(ctx, ntiles, nb, alpha, 1, nullptr, nullptr, d_Ark, 0, nb, nullptr, nullptr, d_Ark, 0, nb, nb, nullptr, 0, nb, kmax,
                        stream, d_AU, d_AV, d_AU, d_AV, d_CD);
}

// ---- HCORE_dgemmbdcd / HCORE_duncompress / HCORE_zuncompress (device pointers) -------------------------------------
// reference src/compute/hcore_dgemmbdcd.c:35-122: CD (M x N) <- alpha AU (AV^T BD) + beta CD, BD dense (rows x N).
static int gemmbdcd_launch(hcore_b200_handle_t ctx, int ntiles, int M, int mc, int N, double alpha,
                        const double *d_AU, const double *d_AV, const int32_t *d_Ark, int lda, int64_t strideA,
                        const double *d_BD, int ldb, int64_t strideB, double beta,
                        double *d_CD, int ldc, int64_t strideC, void *stream, int kcap_in = 0) {
    if (!ctx || ntiles < 0 || M <= 0 || N <= 0 || mc <= 0 || lda < M || lda < mc || ldb < mc || ldc < M) return HCORE_B200_ERR_ARG;
    if (ntiles == 0) return HCORE_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int kcap = round_up(kcap_in > 0 ? kcap_in : rank_cap(strideA, lda, HCORE_B200_MAX_STACK_D), 8), n_pad = round_up(N, 64);
    const size_t per = (size_t)n_pad * kcap * 8;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>({(size_t)ntiles, (size_t)8192, (4ull << 30) / per}));
    CUDA_TRY(ctx, ctx->Ywork.reserve((size_t)chunk * per));
    for (int t0 = 0; t0 < ntiles; t0 += chunk) {
        const int cnt = std::min(chunk, ntiles - t0);
        BdtAvArgs pa{};
        pa.AV = d_AV + (size_t)t0 * strideA; pa.BD = d_BD + (size_t)t0 * strideB; pa.Ark = d_Ark + t0;
        pa.strideA = strideA; pa.strideB = strideB; pa.lda = lda; pa.ldb = ldb; pa.mc = mc; pa.ncols = N;
        pa.Y = (double *)ctx->Ywork.p; pa.n_pad = n_pad; pa.ldy = n_pad; pa.strideY = (long)n_pad * kcap;
        dense_bdtav_kernel<<<dim3(n_pad / 64, cnt), 256, 0, st>>>(pa);
        ctx->launches++;
        int rc = dense_xyt(ctx, cnt, M, N, alpha, d_AU + (size_t)t0 * strideA, lda, strideA, (const double *)ctx->Ywork.p, n_pad,
                           (int64_t)n_pad * kcap, d_Ark + t0, beta, d_CD + (size_t)t0 * strideC, ldc, strideC, st, kcap);
        if (rc != HCORE_B200_OK) return rc;
    }
    return HCORE_B200_OK;
}

int HCORE_dgemmbdcd_batched_strided(hcore_b200_handle_t ctx, int ntiles, int M, int N, double alpha,
                        const double *d_AU, const double *d_AV, const int32_t *d_Ark, int lda, int64_t strideA,
                        const double *d_BD, int ldb, int64_t strideB, double beta,
                        double *d_CD, int ldc, int64_t strideC, void *stream) {
    return gemmbdcd_launch(ctx, ntiles, M, M, N, alpha, d_AU, d_AV, d_Ark, lda, strideA, d_BD, ldb, strideB, beta, d_CD, ldc,
                           strideC, stream);
}

// reference src/compute/hcore_duncompress.c:33-85: CD (M x N) <- AU BV^T with Ark == Brk columns (its alpha / beta
// arguments are ignored: cblas_dgemm is called with one / zero).  A tile whose two ranks differ gets
// HCORE_B200_ERR_RANK_MISMATCH (the reference asserts) and is left untouched.
__global__ void uncompress_check_kernel(const int *Ark, const int *Brk, int *K, int *status, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool ok = Ark[i] == Brk[i];
    K[i] = ok ? Ark[i] : -1;
    if (status) status[i] = ok ? HCORE_B200_OK : HCORE_B200_ERR_RANK_MISMATCH;
}

int HCORE_duncompress_batched_strided(hcore_b200_handle_t ctx, int ntiles, int M, int N,
                        const double *d_AU, const int32_t *d_Ark, int lda, int64_t strideA,
                        const double *d_BV, const int32_t *d_Brk, int ldb, int64_t strideB,
                        double *d_CD, int ldc, int64_t strideC, int32_t *d_status, void *stream) {
    if (!ctx || ntiles < 0 || M <= 0 || N <= 0 || lda < M || ldb < N || ldc < M) return HCORE_B200_ERR_ARG;
    if (ntiles == 0) return HCORE_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(ctx, ctx->info.reserve((size_t)ntiles * sizeof(int)));
    int *K = (int *)ctx->info.p;
    uncompress_check_kernel<<<(ntiles + 255) / 256, 256, 0, st>>>(d_Ark, d_Brk, K, d_status, ntiles);
    ctx->launches++;
    // a mismatched tile carries K = -1: the update kernel skips it (see dense_update_kernel)
    return dense_xyt(ctx, ntiles, M, N, 1.0, d_AU, lda, strideA, d_BV, ldb, strideB, K, 0.0, d_CD, ldc, strideC, st,
                     std::min(rank_cap(strideA, lda, 1 << 30), rank_cap(strideB, ldb, 1 << 30)));
}

int HCORE_zuncompress_batched_strided(hcore_b200_handle_t ctx, int ntiles, int M, int N,
                        const void *d_AU, const int32_t *d_Ark, int lda, int64_t strideA,
                        const void *d_BV, const int32_t *d_Brk, int ldb, int64_t strideB,
                        void *d_CD, int ldc, int64_t strideC, int32_t *d_status, void *stream) {
    if (!ctx || ntiles < 0 || M <= 0 || N <= 0 || lda < M || ldb < N || ldc < M) return HCORE_B200_ERR_ARG;
    if (ntiles == 0) return HCORE_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(ctx, ctx->info.reserve((size_t)ntiles * sizeof(int)));
    int *K = (int *)ctx->info.p;
    uncompress_check_kernel<<<(ntiles + 255) / 256, 256, 0, st>>>(d_Ark, d_Brk, K, d_status, ntiles);
    for (int t0 = 0; t0 < ntiles; t0 += 32768) {
        const int cnt = std::min(32768, ntiles - t0);
        ZUncArgs ar{};
        ar.AU = (const double2 *)d_AU + (size_t)t0 * strideA; ar.BV = (const double2 *)d_BV + (size_t)t0 * strideB;
        ar.Ark = K + t0; ar.strideA = strideA; ar.strideB = strideB; ar.strideC = strideC;
        ar.lda = lda; ar.ldb = ldb; ar.ldc = ldc; ar.m = M; ar.n = N;
        ar.CD = (double2 *)d_CD + (size_t)t0 * strideC;
        zdense_update_kernel<<<dim3((M + 63) / 64, (N + 63) / 64, cnt), 256, 0, st>>>(ar);
        ctx->launches++;
    }
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return HCORE_B200_OK;
}

// ---- tile compression: dense nb x nb tile -> U V^T ------------------------------------------
// The reference declares a generate-and-compress kernel (HCORE_zgytlr, src/include/hcore_z.h:63-75) but never defines it;
// its tests compress with LAPACKE_dgesvd and keep rank = #{sigma >= acc}, capped at the storage rank
// (testing/testing_dgemm.c:119-236).  Here, built from the kernels of the update path:
//   1. Y = A Omega (Omega: nb x p deterministic test matrix, p = maxrk + 16), Q = orth(Y) -- the orthonormalisation IS one
//      pass of the update pipeline on the tile (CU = Y, CV = [I; 0]) with a zero update and fixed rank p: its new CU are
//      the left singular vectors of Y;
//   2. optional power iteration (default 1): Q <- orth(A orth(A^T Q)), which makes the captured subspace accurate to
//      far below acc even for slowly decaying spectra;
//   3. V0 = A^T Q, so Q V0^T = Q Q^T A; one more pass of the update pipeline with the caller's accuracy truncates it by
//      the update path's own rule: U orthonormal, V carries the singular values, new rank in d_rank (capped at maxrk).
__global__ void fill_int_kernel(int *p, int v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

int HCORE_dcompress_batched_strided(hcore_b200_handle_t ctx, int ntiles, int nb, const double *d_A, int lda, int64_t strideA,
                        double acc, int maxrk, int power_iters, double *d_U, double *d_V, int ldu, int64_t strideUV,
                        int32_t *d_rank, int32_t *d_status, void *stream) {
    if (!ctx || ntiles < 0 || nb < 16 || lda < nb || ldu < nb || maxrk < 1 || power_iters < 0 || !d_rank) return HCORE_B200_ERR_ARG;
    if (strideUV > 0 && strideUV < (int64_t)ldu * maxrk) return HCORE_B200_ERR_ARG;
    if (ntiles == 0) return HCORE_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int p = std::min(round_up(maxrk + 16, 8), 248);
    p = std::min(p, (nb - 1) & ~7);                      // the stack [Q | 0] has p + 1 columns and must not be wider than the tile
    if (p < 8) return HCORE_B200_ERR_ARG;
    const int pc = p + 1;
    const size_t per = 2ull * nb * pc * 8;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>({(size_t)ntiles, (size_t)592, (3ull << 30) / per}));
    CUDA_TRY(ctx, ctx->cmpOm.reserve((size_t)nb * p * 8));
    CUDA_TRY(ctx, ctx->cmpCU.reserve((size_t)chunk * nb * pc * 8));
    CUDA_TRY(ctx, ctx->cmpCV.reserve((size_t)chunk * nb * pc * 8));
    CUDA_TRY(ctx, ctx->cmpZero.reserve((size_t)nb * 8));
    CUDA_TRY(ctx, ctx->cmpInts.reserve((size_t)chunk * 4 * sizeof(int)));
    double *Om = (double *)ctx->cmpOm.p, *bufA = (double *)ctx->cmpCU.p, *bufB = (double *)ctx->cmpCV.p, *zero = (double *)ctx->cmpZero.p;
    int *ones = (int *)ctx->cmpInts.p, *crk = ones + chunk, *stw = crk + chunk;
    const long strideI = (long)nb * pc;
    CUDA_TRY(ctx, cudaMemsetAsync(zero, 0, (size_t)nb * 8, st));
    fill_omega_kernel<<<(unsigned)(((size_t)nb * p + 255) / 256), 256, 0, st>>>(Om, nb, p, nb, 0x243F6A8885A308D3ull);
    fill_int_kernel<<<(chunk + 255) / 256, 256, 0, st>>>(ones, 1, chunk);
    ctx->launches += 2;
    const int n_pad = round_up(nb, 64);
    // one pass of the update pipeline on (CU, CV) with a zero update: rk > 0 = fixed rank (orthonormalisation), rk = 0 = accuracy
    auto pass = [&](int cnt, double *CU, double *CV, int rk, double a) -> int {
        fill_int_kernel<<<(cnt + 255) / 256, 256, 0, st>>>(crk, p, cnt);
        ctx->launches++;
        return tlr_batched_strided(ctx, cnt, nb, 0.0, zero, zero, ones, nb, 0, zero, zero, ones, nb, 0, 1.0, CU, CV, crk, nb, strideI, rk,
                                   pc, a, stw, st, 0);
    };
    auto identity = [&](int cnt, double *E) {
        fill_identity_kernel<<<dim3((unsigned)(((size_t)nb * p + 255) / 256), cnt), 256, 0, st>>>(E, strideI, nb, nb, p);
        ctx->launches++;
    };
    auto ax = [&](int cnt, const double *A, const double *X, long strideX, double *Y) {       // Y = A X
        AxArgs a{};
        a.A = A; a.X = X; a.Prk = nullptr; a.strideA = strideA; a.strideX = strideX; a.strideY = strideI;
        a.lda = lda; a.ldx = nb; a.ldy = nb; a.m = nb; a.n = nb; a.p = p; a.Y = Y;
        dense_ax_kernel<<<dim3(n_pad / 64, cnt), 256, 0, st>>>(a);
        ctx->launches++;
    };
    auto atq = [&](int cnt, const double *A, const double *Q, double *Y) {                      // Y = A^T Q
        BdtAvArgs a{};
        a.AV = Q; a.BD = A; a.Ark = crk; a.strideA = strideI; a.strideB = strideA; a.lda = nb; a.ldb = lda; a.mc = nb; a.ncols = nb;
        a.Y = Y; a.strideY = strideI; a.ldy = nb; a.n_pad = n_pad;
        dense_bdtav_kernel<<<dim3(n_pad / 64, cnt), 256, 0, st>>>(a);
        ctx->launches++;
    };
    for (int t0 = 0; t0 < ntiles; t0 += chunk) {
        const int cnt = std::min(chunk, ntiles - t0);
        const double *A = d_A + (size_t)t0 * strideA;
        int rc;
        ax(cnt, A, Om, 0, bufA);                                   // Y = A Omega
        identity(cnt, bufB);
        if ((rc = pass(cnt, bufA, bufB, p, acc)) != HCORE_B200_OK) return rc;      // bufA <- Q
        for (int it = 0; it < power_iters; ++it) {
            atq(cnt, A, bufA, bufB);                                // Z = A^T Q
            identity(cnt, bufA);
            if ((rc = pass(cnt, bufB, bufA, p, acc)) != HCORE_B200_OK) return rc;  // bufB <- orth(Z)
            ax(cnt, A, bufB, strideI, bufA);                       // Y = A orth(Z)
            identity(cnt, bufB);
            if ((rc = pass(cnt, bufA, bufB, p, acc)) != HCORE_B200_OK) return rc;  // bufA <- Q
        }
        fill_int_kernel<<<(cnt + 255) / 256, 256, 0, st>>>(crk, p, cnt);          // atq reads the column count from crk
        atq(cnt, A, bufA, bufB);                                    // V0 = A^T Q
        if ((rc = pass(cnt, bufA, bufB, 0, acc)) != HCORE_B200_OK) return rc;      // truncation to acc
        compress_copy_out_kernel<<<dim3(16, cnt), 256, 0, st>>>(bufA, bufB, strideI, nb, d_U + (size_t)t0 * strideUV,
                                                                 d_V + (size_t)t0 * strideUV, strideUV, ldu, nb, crk, stw, maxrk,
                                                                 d_rank + t0, d_status ? d_status + t0 : nullptr);
        ctx->launches += 2;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return HCORE_B200_OK;
}

// ---- HCORE_dtrsm / HCORE_dpotrf (device pointers) ---------------------------------------------------------------
// reference src/compute/hcore_dtrsm.c:24-48 (cblas_dtrsm): B (M x Brank) <- alpha op(A)^-1 B (side = Left, A is M x M)
// or alpha B op(A)^-1 (side = Right, A is Brank x Brank).  strideA = 0: one triangular tile for the whole batch.
int HCORE_dtrsm_batched_strided(hcore_b200_handle_t ctx, int ntiles, int side, int uplo, int transA, int diag, int M, int Brank,
                        double alpha, const double *d_A, int lda, int64_t strideA,
                        double *d_BV, const int32_t *d_Brk, int ldbv, int64_t strideB, void *stream) {
    if (!ctx || ntiles < 0 || M <= 0 || Brank <= 0 || ldbv < M) return HCORE_B200_ERR_ARG;
    if ((side != 141 && side != 142) || (uplo != 121 && uplo != 122) || (transA < 111 || transA > 113) || (diag != 131 && diag != 132))
        return HCORE_B200_ERR_ARG;
    const bool right = side == 142, tr = transA != 111;
    if (lda < (right ? Brank : M)) return HCORE_B200_ERR_ARG;
    if (right && d_Brk && strideA == 0 && ntiles > 1) { ctx->err = "HCORE_dtrsm_batched: side = Right with per-tile ranks needs one A per tile"; return HCORE_B200_ERR_ARG; }
    if (ntiles == 0) return HCORE_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int msz_max = right ? Brank : M;
    const int nblk = (msz_max + kTb - 1) / kTb, msz_pad = nblk * kTb;
    const int tilesA = strideA ? ntiles : 1;
    CUDA_TRY(ctx, ctx->Winv.reserve((size_t)tilesA * nblk * kTb * kTb * 8));
    TrsmArgs ar{};
    // op(A)(i, j): strides (1, lda) or transposed (lda, 1); op(A) is lower iff (uplo == Lower) xor trans
    long si = tr ? lda : 1, sj = tr ? 1 : lda;
    bool lower = (uplo == 122) != tr;
    long bi = 1, bj = ldbv;
    if (right) {                       // X op(A) = alpha B  <=>  op(A)^T X^T = alpha B^T
        std::swap(si, sj);
        lower = !lower;
        std::swap(bi, bj);
    }
    if (!lower) { si = -si; sj = -sj; bi = -bi; }     // backward substitution = forward on the reversed index
    ar.A = d_A; ar.strideA = strideA; ar.a_si = si; ar.a_sj = sj; ar.a_off = 0;
    ar.B = d_BV; ar.strideB = strideB; ar.b_si = bi; ar.b_sj = bj; ar.b_off = 0;
    ar.Brk = d_Brk; ar.rank = Brank; ar.M = M; ar.side_right = right; ar.unit_diag = (diag == 132); ar.alpha = alpha;
    ar.Winv = (double *)ctx->Winv.p; ar.nblk_max = nblk;
    if (!ctx->attr_done[5]) {
        ctx->attr_done[5] = true;
        CUDA_TRY(ctx, cudaFuncSetAttribute(trsm_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCholSmemBytes));
        CUDA_TRY(ctx, cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCholSmemBytes));
        CUDA_TRY(ctx, cudaFuncSetAttribute(potrf_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCholSmemBytes));
        CUDA_TRY(ctx, cudaFuncSetAttribute(potrf_trail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCholSmemBytes));
    }
    trsm_inv_kernel<<<dim3(nblk, tilesA), 256, kCholSmemBytes, st>>>(ar);
    const int nrhs_max = right ? M : Brank;
    size_t sm16 = (size_t)trsm_ldx(msz_pad) * 16 * 8, sm8 = (size_t)trsm_ldx(msz_pad) * 8 * 8;
    if (sm16 <= 200 * 1024) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(trsm_solve_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm16));
        trsm_solve_kernel<16><<<dim3((nrhs_max + 15) / 16, ntiles), 256, sm16, st>>>(ar, msz_pad);
    } else if (sm8 <= 227 * 1024) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(trsm_solve_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm8));
        trsm_solve_kernel<8><<<dim3((nrhs_max + 7) / 8, ntiles), 256, sm8, st>>>(ar, msz_pad);
    } else {
        ctx->err = "HCORE_dtrsm_batched: triangular dimension above 3584 is not supported";
        return HCORE_B200_ERR_ARG;
    }
    ctx->launches += 2;
    CUDA_TRY(ctx, cudaGetLastError());
    return HCORE_B200_OK;
}

// reference src/compute/hcore_dpotrf.c:23-35 (LAPACKE_dpotrf_work): A = L L^T (uplo = Lower) or U^T U (Upper) in place,
// d_info[t] = 0 or the order of the first leading minor that is not positive definite.
int HCORE_dpotrf_batched_strided(hcore_b200_handle_t ctx, int ntiles, int uplo, int N, double *d_A, int lda, int64_t strideA,
                        int32_t *d_info, void *stream) {
    if (!ctx || ntiles < 0 || N <= 0 || lda < N || !d_info || (uplo != 121 && uplo != 122)) return HCORE_B200_ERR_ARG;
    if (ntiles == 0) return HCORE_B200_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(ctx, ctx->Winv.reserve((size_t)ntiles * kTb * kTb * 8));
    CUDA_TRY(ctx, cudaMemsetAsync(d_info, 0, (size_t)ntiles * sizeof(int), st));
    if (!ctx->attr_done[5]) {
        ctx->attr_done[5] = true;
        CUDA_TRY(ctx, cudaFuncSetAttribute(trsm_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCholSmemBytes));
        CUDA_TRY(ctx, cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCholSmemBytes));
        CUDA_TRY(ctx, cudaFuncSetAttribute(potrf_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCholSmemBytes));
        CUDA_TRY(ctx, cudaFuncSetAttribute(potrf_trail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCholSmemBytes));
    }
    PotrfArgs ar{};
    ar.A = d_A; ar.strideA = strideA; ar.n = N; ar.info = d_info; ar.Winv = (double *)ctx->Winv.p;
    ar.si = (uplo == 122) ? 1 : lda; ar.sj = (uplo == 122) ? lda : 1;
    const int nblk = (N + kTb - 1) / kTb;
    for (int t0 = 0; t0 < ntiles; t0 += 32768) {
        const int cnt = std::min(32768, ntiles - t0);
        PotrfArgs a2 = ar;
        a2.A = d_A + (size_t)t0 * strideA; a2.info = d_info + t0;
        for (int jb = 0; jb < nblk; ++jb) {
            const int j0 = jb * kTb, rem = nblk - jb - 1;
            potrf_diag_kernel<<<cnt, 256, kCholSmemBytes, st>>>(a2, j0);
            ctx->launches++;
            if (rem > 0) {
                potrf_panel_kernel<<<dim3(rem, cnt), 256, kCholSmemBytes, st>>>(a2, j0);
                potrf_trail_kernel<<<dim3(rem * (rem + 1) / 2, cnt), 256, kCholSmemBytes, st>>>(a2, j0);
                ctx->launches += 2;
            }
        }
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return HCORE_B200_OK;
}

// ---- debug / profiling -------------------------------------------------------------------------
#ifdef HCORE_B200_PROF
// phase counters of the instrumented build (tools/phase_prof.py)
int hcore_b200_prof_read(unsigned long long *out, int n, int reset) {
    unsigned long long h[64];
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(h, hcore_b200::g_prof, sizeof(h)) != cudaSuccess) return HCORE_B200_ERR_CUDA;
    for (int i = 0; i < n && i < 64; ++i) out[i] = h[i];
    if (reset) {
        memset(h, 0, sizeof(h));
        cudaMemcpyToSymbol(hcore_b200::g_prof, h, sizeof(h));
    }
    return HCORE_B200_OK;
}
#endif

int hcore_b200_debug_set_stages(hcore_b200_handle_t ctx, int nstages) {
    if (!ctx) return HCORE_B200_ERR_ARG;
    ctx->debug_stages = nstages;
    return HCORE_B200_OK;
}
int hcore_b200_set_stage_timing(hcore_b200_handle_t ctx, int enable) {
    if (!ctx) return HCORE_B200_ERR_ARG;
    ctx->stage_timing = enable != 0;
    return HCORE_B200_OK;
}
double hcore_b200_stage_ms(hcore_b200_handle_t ctx, int stage) {
    return (ctx && stage >= 0 && stage < 4) ? ctx->stage_ms[stage] : -1.0;
}

int hcore_b200_debug_fetch(hcore_b200_handle_t ctx, int what, int tile, double *h_out, size_t nelem) {
    if (!ctx || !h_out || ctx->last_N == 0) return HCORE_B200_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    const TlrWorkspace &ws = ctx->last_ws;
    const int N = ctx->last_N;
    const size_t nn = (size_t)N * N, rr = (size_t)ws.nblk * (N / kSubW) * kRecDoubles, gg = (size_t)ws.m_pad * N;
    const double *src = nullptr;
    size_t avail = 0;
    switch (what) {
        case 0: src = ws.rec + ((size_t)0 * ws.chunk + tile) * rr; avail = rr; break;   // raw records, U panel
        case 1: src = ws.rec + ((size_t)1 * ws.chunk + tile) * rr; avail = rr; break;   // raw records, V panel
        case 2: src = ws.R + ((size_t)0 * ws.chunk + tile) * nn; avail = nn; break;
        case 3: src = ws.R + ((size_t)1 * ws.chunk + tile) * nn; avail = nn; break;
        case 4: src = ws.Y + ((size_t)0 * ws.chunk + tile) * nn; avail = nn; break;
        case 5: src = ws.Y + ((size_t)1 * ws.chunk + tile) * nn; avail = nn; break;
        case 6: src = ws.sigma + (size_t)tile * N; avail = N; break;
        case 7: src = ws.G + (size_t)tile * gg; avail = gg; break;
        case 9: {   // ints: ncols, newrank, status, sweeps (as doubles)
            int v[4];
            CUDA_TRY(ctx, cudaMemcpy(&v[0], ws.ncols + tile, 4, cudaMemcpyDeviceToHost));
            CUDA_TRY(ctx, cudaMemcpy(&v[1], ws.newrank + tile, 4, cudaMemcpyDeviceToHost));
            CUDA_TRY(ctx, cudaMemcpy(&v[2], ws.st + tile, 4, cudaMemcpyDeviceToHost));
            CUDA_TRY(ctx, cudaMemcpy(&v[3], ws.sweeps + tile, 4, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < std::min<size_t>(4, nelem); ++i) h_out[i] = v[i];
            return HCORE_B200_OK;
        }
        default: return HCORE_B200_ERR_ARG;
    }
    CUDA_TRY(ctx, cudaMemcpy(h_out, src, std::min(avail, nelem) * 8, cudaMemcpyDeviceToHost));
    return HCORE_B200_OK;
}

// ==================================================================================================
// Legacy single-tile drop-in symbols (host pointers).  Re-entrant like the reference with use_scratch = 0
// (reference src/compute/hcore_dgemm.c:761-923): every calling thread owns its context.
// ==================================================================================================
// One context (workspace, staging buffers, stream) PER CALLING THREAD: a task runtime with one worker thread per core
// (HiCMA over StarPU / PaRSEC) gets as many independent streams as it has workers and their single-tile updates
// overlap on the GPU; nothing is shared between threads, so the split-phase pair of one thread cannot be disturbed
// by another thread.  Contexts are released when their thread exits.
struct LegacySlot {
    hcore_b200_handle_t ctx = nullptr;
    ~LegacySlot() { if (ctx) hcore_b200_destroy(ctx); }
};
static thread_local LegacySlot tl_slot;

static hcore_b200_handle_t legacy_ctx(const char *who) {
    if (!tl_slot.ctx) {
        if (hcore_b200_create(&tl_slot.ctx, -1) != HCORE_B200_OK) {
            fprintf(stderr, "%s: cannot create a CUDA context (no GPU?). hcore_b200 has no CPU path.\n", who);
            exit(-1);
        }
    }
    return tl_slot.ctx;
}

static void legacy_fail(const char *who, const char *what, int code) {
    // same convention as the reference: message on stderr, then exit(-1)
    fprintf(stderr, "%s: %s (code %d)\n", who, what, code);
    exit(-1);
}
static inline void lchk(const char *who, cudaError_t e) {          // every staging copy is checked
    if (e != cudaSuccess) legacy_fail(who, cudaGetErrorString(e), HCORE_B200_ERR_CUDA);
}

// Shared body of the legacy HCORE_dgemm / HCORE_sgemm: host arrays of element size `es`.
static int legacy_gemm(const char *who, int es, int M, double alpha, const void *AU, const void *AV, int a, int LDA,
                       const void *BU, const void *BV, int b, int LDB, double beta, void *CU, void *CV, int c, int LDC,
                       int rk, int maxrk, double acc, int scan_from = 1) {
    hcore_b200_handle_t ctx = legacy_ctx(who);
    if (a == 0 || b == 0 || c == 0) legacy_fail(who, "rank values should not be zero", HCORE_B200_ERR_ZERO_RANK);
    if (maxrk < a + c) legacy_fail(who, "increase maxrank (maxrk < Ark + Crk)", HCORE_B200_ERR_MAXRK);
    cudaSetDevice(ctx->device);
    // Only the columns the reference reads (c) and writes (new rank) are moved, so callers
    // whose CU/CV buffers are smaller than LDC x maxrk keep working (reference test:
    // testing/testing_dgemm.c:287-290 allocates M x storage_maxrank).
    const size_t fa = (size_t)LDA * a * es, fb = (size_t)LDB * b * es, fcin = (size_t)LDC * c * es;
    const size_t fc = ((size_t)LDC * (a + c) * es + 15) / 16 * 16;
    const size_t fa16 = (fa + 15) / 16 * 16, fb16 = (fb + 15) / 16 * 16;
    if (ctx->stage[0].reserve(2 * fa16 + 2 * fb16 + 2 * fc) != cudaSuccess ||
        ctx->stage_ints[0].reserve(8 * sizeof(int)) != cudaSuccess)
        legacy_fail(who, "device allocation failed", HCORE_B200_ERR_CUDA);
    char *base = (char *)ctx->stage[0].p;
    char *dAU = base, *dAV = base + fa16, *dBU = base + 2 * fa16, *dBV = base + 2 * fa16 + fb16,
         *dCU = base + 2 * fa16 + 2 * fb16, *dCV = base + 2 * fa16 + 2 * fb16 + fc;
    int *di = (int *)ctx->stage_ints[0].p;
    // The caller's arrays are pageable: they are gathered into a PINNED mirror of the device staging block (same
    // layout) and travel in ONE asynchronous copy on this thread's stream.  A pageable cudaMemcpyAsync goes through a
    // driver-wide staging buffer and serialises the worker threads of a task runtime against each other.
    const size_t total = 2 * fa16 + 2 * fb16 + 2 * fc;
    if (ctx->pin_reserve(total + 64) != cudaSuccess) legacy_fail(who, "pinned host allocation failed", HCORE_B200_ERR_CUDA);
    char *hb = (char *)ctx->pin;
    int *hi = (int *)(hb + total);
    hi[0] = a; hi[1] = b; hi[2] = c; hi[3] = 0;
    memcpy(hb, AU, fa); memcpy(hb + fa16, AV, fa);
    memcpy(hb + 2 * fa16, BU, fb); memcpy(hb + 2 * fa16 + fb16, BV, fb);
    memcpy(hb + 2 * fa16 + 2 * fb16, CU, fcin); memcpy(hb + 2 * fa16 + 2 * fb16 + fc, CV, fcin);
    cudaStream_t st = ctx->streams[0];
    lchk(who, cudaMemcpyAsync(base, hb, 2 * fa16 + 2 * fb16 + fcin, cudaMemcpyHostToDevice, st));          // A, B factors and CU
    lchk(who, cudaMemcpyAsync(dCV, hb + 2 * fa16 + 2 * fb16 + fc, fcin, cudaMemcpyHostToDevice, st));
    lchk(who, cudaMemcpyAsync(di, hi, 4 * sizeof(int), cudaMemcpyHostToDevice, st));
    int rc = tlr_batched_strided(ctx, 1, M, alpha, dAU, dAV, di, LDA, 0, dBU, dBV, di + 1, LDB, 0, beta, dCU, dCV, di + 2,
                                 LDC, 0, rk, maxrk, acc, di + 3, st, es == 4, 0, nullptr, scan_from);
    if (rc != HCORE_B200_OK) legacy_fail(who, hcore_b200_last_error(ctx), rc);
    lchk(who, cudaMemcpyAsync(hi, di, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (cudaStreamSynchronize(st) != cudaSuccess) legacy_fail(who, "CUDA failure", HCORE_B200_ERR_CUDA);
    if (hi[3] != HCORE_B200_OK) legacy_fail(who, "tile update failed", hi[3]);
    const int rnew = hi[2];
    const size_t fout = (size_t)LDC * rnew * es;
    lchk(who, cudaMemcpyAsync(hb, dCU, fout, cudaMemcpyDeviceToHost, st));
    lchk(who, cudaMemcpyAsync(hb + fc, dCV, fout, cudaMemcpyDeviceToHost, st));
    if (cudaStreamSynchronize(st) != cudaSuccess) legacy_fail(who, "CUDA failure", HCORE_B200_ERR_CUDA);
    memcpy(CU, hb, fout); memcpy(CV, hb + fc, fout);
    return rnew;
}

void HCORE_dgemm(HCORE_enum transA, int transB, int M, int N, double alpha, double *AU, double *AV, double *Ark,
                 int LDA, double *BU, double *BV, double *Brk, int LDB, double beta, double *CU, double *CV,
                 double *Crk, int LDC, int rk, int maxrk, double acc, double *work, flop_counter *flops) {
    (void)transA; (void)transB; (void)N; (void)work;   // never read by the reference either (hcore_dgemm.c:703-755)
    const int a = (int)Ark[0], b = (int)Brk[0], c = (int)Crk[0];
    const int rnew = legacy_gemm("HCORE_dgemm", 8, M, alpha, AU, AV, a, LDA, BU, BV, b, LDB, beta, CU, CV, c, LDC, rk, maxrk, acc);
    if (flops) flops->update += hcore_b200_gemm_flops(M, a, b, c, rnew);
    Crk[0] = (double)rnew;
}

// ---- split-phase legacy symbols (reference src/compute/hcore_dgemm.c:1312-1597 and :1602-1699) -----------
// HCORE_dgemm_qr_svd reveals the new rank, the caller may re-size its C factors, HCORE_dgemm_ormqr then
// forms them.  The reference hands its intermediate host buffers from one call to the other through 19
// pass-through parameters; here the intermediates (block reflectors, U_r, W) stay on the device in the
// process-wide context, so the pass-through pointers are returned as NULL / the caller's own CU, CV, and
// the two calls must be paired (qr_svd, then ormqr for the same tile, no other legacy call in between).
struct SplitState { bool pending = false; int M = 0, n = 0, rnew = 0, LDC = 0; char *dCU = nullptr, *dCV = nullptr;
                    const double *hCU = nullptr, *hCV = nullptr; };
static thread_local SplitState g_split;     // per thread, and checked against the caller's CU / CV pointers

static uint64_t ormqr_flops(uint64_t m, uint64_t n, uint64_t r) {       // the two dormqr terms of the model
    return (4 * r * m * n - 2 * r * n * n + 3 * r * n) + (4 * m * r * n - 2 * r * n * n + 2 * r * n + m * n - n * n / 2 + n / 2);
}

// Shared body of HCORE_dgemm_qr_svd (B = BU BV^T) and HCORE_dgemm_qr_svd_b_dense (B dense, Bd != null).
static int legacy_qr_svd(const char *who, int M, double alpha, const double *AU, const double *AV, int a, int LDA,
                         const double *BU, const double *BV, int b, const double *Bd, int LDB, double beta,
                         const double *CU, const double *CV, int c, int LDC, int rk, int maxrk, double acc,
                         flop_counter *flops) {
    hcore_b200_handle_t ctx = legacy_ctx(who);
    if (a == 0 || c == 0 || (!Bd && b == 0)) legacy_fail(who, "rank values should not be zero", HCORE_B200_ERR_ZERO_RANK);
    // the split-phase variant accepts Ark + Crk up to 2 * maxrk (reference :1186, :1259)
    if (2 * maxrk < a + c) legacy_fail(who, "increase maxrank (2 * maxrk < Ark + Crk)", HCORE_B200_ERR_MAXRK);
    const int cap = std::max(maxrk, a + c);
    cudaSetDevice(ctx->device);
    const size_t fa = (size_t)LDA * a * 8, fcin = (size_t)LDC * c * 8;
    const size_t fb = Bd ? (size_t)LDB * M * 8 : (size_t)LDB * b * 8;       // dense B: one M x M array
    const size_t fc = ((size_t)LDC * (a + c) * 8 + 15) / 16 * 16;
    const size_t fa16 = (fa + 15) / 16 * 16, fb16 = (fb + 15) / 16 * 16;
    if (ctx->stage[0].reserve(2 * fa16 + 2 * fb16 + 2 * fc) != cudaSuccess ||
        ctx->stage_ints[0].reserve(8 * sizeof(int)) != cudaSuccess)
        legacy_fail(who, "device allocation failed", HCORE_B200_ERR_CUDA);
    char *base = (char *)ctx->stage[0].p;
    char *dAU = base, *dAV = base + fa16, *dBU = base + 2 * fa16, *dBV = base + 2 * fa16 + fb16,
         *dCU = base + 2 * fa16 + 2 * fb16, *dCV = base + 2 * fa16 + 2 * fb16 + fc;
    int *di = (int *)ctx->stage_ints[0].p;
    int hi[5] = {a, Bd ? a : b, c, 0, 0};
    cudaStream_t st = ctx->streams[0];
    lchk(who, cudaMemcpyAsync(dAU, AU, fa, cudaMemcpyHostToDevice, st));
    lchk(who, cudaMemcpyAsync(dAV, AV, fa, cudaMemcpyHostToDevice, st));
    if (Bd) {
        lchk(who, cudaMemcpyAsync(dBU, Bd, fb, cudaMemcpyHostToDevice, st));
    } else {
        lchk(who, cudaMemcpyAsync(dBU, BU, fb, cudaMemcpyHostToDevice, st));
        lchk(who, cudaMemcpyAsync(dBV, BV, fb, cudaMemcpyHostToDevice, st));
    }
    lchk(who, cudaMemcpyAsync(dCU, CU, fcin, cudaMemcpyHostToDevice, st));
    lchk(who, cudaMemcpyAsync(dCV, CV, fcin, cudaMemcpyHostToDevice, st));
    lchk(who, cudaMemcpyAsync(di, hi, sizeof(hi), cudaMemcpyHostToDevice, st));
    int rc;
    if (Bd) {
        // one tile: the pointer "arrays" are single device pointers kept in the context's pointer table
        if (ctx->ptrs.reserve(6 * sizeof(double *)) != cudaSuccess) legacy_fail(who, "device allocation failed", HCORE_B200_ERR_CUDA);
        const double *hp[6] = {(double *)dAU, (double *)dAV, (double *)dBU, (double *)dCU, (double *)dCV, nullptr};
        lchk(who, cudaMemcpyAsync(ctx->ptrs.p, hp, sizeof(hp), cudaMemcpyHostToDevice, st));
        const double **pp = (const double **)ctx->ptrs.p;
        rc = HCORE_dgemm_batched_qr_svd_b_dense(ctx, 1, M, alpha, pp, pp + 1, di, LDA, pp + 2, LDB, beta, pp + 3, pp + 4, di + 2,
                                                LDC, rk, cap, acc, di + 4, di + 3, st);
    } else {
        rc = tlr_batched_strided(ctx, 1, M, alpha, dAU, dAV, di, LDA, 0, dBU, dBV, di + 1, LDB, 0, beta, dCU, dCV, di + 2,
                                 LDC, 0, rk, cap, acc, di + 3, st, 0, 1, di + 4);
    }
    if (rc != HCORE_B200_OK) legacy_fail(who, hcore_b200_last_error(ctx), rc);
    lchk(who, cudaMemcpyAsync(hi, di, sizeof(hi), cudaMemcpyDeviceToHost, st));
    if (cudaStreamSynchronize(st) != cudaSuccess) legacy_fail(who, "CUDA failure", HCORE_B200_ERR_CUDA);
    if (hi[3] != HCORE_B200_OK) legacy_fail(who, "tile update failed", hi[3]);
    const int rnew = hi[4];
    g_split.pending = true; g_split.M = M; g_split.n = a + c; g_split.rnew = rnew; g_split.LDC = LDC;
    g_split.dCU = dCU; g_split.dCV = dCV; g_split.hCU = CU; g_split.hCV = CV;
    if (flops) {
        if (Bd) {
            // reference model of the b_dense path: QR of both stacks + the M x Ark x M product (:272) + SVD parts
            const uint64_t m = M, n = (uint64_t)a + c, r = rnew;
            auto geqrf = [](uint64_t mm, uint64_t nn) {
                return 2 * mm * nn * nn - (uint64_t)(2 * nn * nn * nn / 3.0f) + 2 * mm * nn + (uint64_t)(17 * nn / 3.0f);
            };
            flops->update += 2 * geqrf(m, n) + 2 * m * (uint64_t)a * m + 22 * n * n * n + r * n;
        } else {
            flops->update += hcore_b200_gemm_flops(M, a, b, c, rnew) - ormqr_flops(M, a + c, rnew);
        }
    }
    return rnew;
}

#define HCORE_SPLIT_OUTS                                                                                          \
    *_p_work_new = work; *_p__CU = CU; *_p__CV = CV; *_p_CU_ncols = a + c; *_p_new_UVrk = rnew;                   \
    *_p_newU = nullptr; *_p_ld_newU = M; *_p_qrtauA = nullptr; *_p_CV_ncols = a + c; *_p_newV = nullptr;           \
    *_p_ld_newV = maxrk; *_p_qrtauB = nullptr; *_p_use_CUV_clone = 0; *_p_CUclone = nullptr; *_p_ld_CUclone = M;   \
    *_p__CU_save = CU; *_p_CVclone = nullptr; *_p_ld_CVclone = M; *_p__CV_save = CV;

void HCORE_dgemm_qr_svd(int transA, int transB, int M, int N, double alpha, double *AU, double *AV, int *Ark, int LDA,
                        double *BU, double *BV, int *Brk, int LDB, double beta, double *CU, double *CV, int *Crk, int LDC,
                        int rk, int storage_maxrank, int maxrk, double acc, double *work, flop_counter *flops,
                        double **_p_work_new, double **_p__CU, double **_p__CV, int *_p_CU_ncols, int *_p_new_UVrk,
                        double **_p_newU, int *_p_ld_newU, double **_p_qrtauA, int *_p_CV_ncols, double **_p_newV,
                        int *_p_ld_newV, double **_p_qrtauB, int *_p_use_CUV_clone, double **_p_CUclone,
                        int *_p_ld_CUclone, double **_p__CU_save, double **_p_CVclone, int *_p_ld_CVclone,
                        double **_p__CV_save) {
    (void)transA; (void)transB; (void)N; (void)storage_maxrank;
    const int a = Ark[0], b = Brk[0], c = Crk[0];
    const int rnew = legacy_qr_svd("HCORE_dgemm_qr_svd", M, alpha, AU, AV, a, LDA, BU, BV, b, nullptr, LDB, beta, CU, CV, c, LDC,
                                   rk, maxrk, acc, flops);
    HCORE_SPLIT_OUTS
}

// reference src/compute/hcore_dgemm.c:1022-1303: B is a dense M x M tile (leading dimension LDB)
void HCORE_dgemm_qr_svd_b_dense(int transA, int transB, int M, int N, double alpha, double *AU, double *AV, int *Ark, int LDA,
                        double *B, int LDB, double beta, double *CU, double *CV, int *Crk, int LDC,
                        int rk, int storage_maxrank, int maxrk, double acc, double *work, flop_counter *flops,
                        double **_p_work_new, double **_p__CU, double **_p__CV, int *_p_CU_ncols, int *_p_new_UVrk,
                        double **_p_newU, int *_p_ld_newU, double **_p_qrtauA, int *_p_CV_ncols, double **_p_newV,
                        int *_p_ld_newV, double **_p_qrtauB, int *_p_use_CUV_clone, double **_p_CUclone,
                        int *_p_ld_CUclone, double **_p__CU_save, double **_p_CVclone, int *_p_ld_CVclone,
                        double **_p__CV_save) {
    (void)transA; (void)transB; (void)N; (void)storage_maxrank;
    const int a = Ark[0], c = Crk[0];
    const int rnew = legacy_qr_svd("HCORE_dgemm_qr_svd_b_dense", M, alpha, AU, AV, a, LDA, nullptr, nullptr, 0, B, LDB, beta, CU,
                                   CV, c, LDC, rk, maxrk, acc, flops);
    HCORE_SPLIT_OUTS
}

void HCORE_dgemm_ormqr(int transA, int transB, int M, int N, double alpha, double *AU, double *AV, int *Ark, int LDA,
                       double beta, double *CU, double *CV, int *Crk, int LDC, int rk, int storage_maxrank, int maxrk,
                       double acc, double *work, flop_counter *flops, double *_CU, double *_CV, int CU_ncols,
                       int new_UVrk, double *newU, int ld_newU, double *qrtauA, int CV_ncols, double *newV, int ld_newV,
                       double *qrtauB, int use_CUV_clone, double *CUclone, int ld_CUclone, double *_CU_save,
                       double *CVclone, int ld_CVclone, double *_CV_save) {
    (void)transA; (void)transB; (void)N; (void)alpha; (void)AU; (void)AV; (void)Ark; (void)LDA; (void)beta; (void)rk;
    (void)maxrk; (void)acc; (void)work; (void)CU_ncols; (void)newU; (void)ld_newU; (void)qrtauA; (void)CV_ncols; (void)newV;
    (void)ld_newV; (void)qrtauB; (void)use_CUV_clone; (void)CUclone; (void)ld_CUclone; (void)CVclone; (void)ld_CVclone;
    const char *who = "HCORE_dgemm_ormqr";
    hcore_b200_handle_t ctx = legacy_ctx(who);
    // the pair is matched on the calling thread, the tile size, the revealed rank and the identity of the C factors
    // the first half handed back through *_p__CU / *_p__CV (the reference passes its state the same way)
    if (!g_split.pending || g_split.M != M || g_split.rnew != new_UVrk || (_CU && _CU != g_split.hCU) || (_CV && _CV != g_split.hCV))
        legacy_fail(who, "no matching HCORE_dgemm_qr_svd call precedes this one on this thread", HCORE_B200_ERR_ARG);
    if (new_UVrk > storage_maxrank) {                     // reference :1649-1652
        fprintf(stderr, "%s: Increase storage maxrank %d. Observed rank for U and V is %d\n", who, storage_maxrank, new_UVrk);
        exit(-1);
    }
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->streams[0];
    int *di = (int *)ctx->stage_ints[0].p;
    // the device copies keep the leading dimension of the first call; the host arrays may have a new one
    int rc = HCORE_dgemm_batched_ormqr_strided(ctx, 1, M, (double *)g_split.dCU, (double *)g_split.dCV, di + 2, g_split.LDC, 0, st);
    if (rc != HCORE_B200_OK) legacy_fail(who, hcore_b200_last_error(ctx), rc);
    double *outU = _CU_save ? _CU_save : CU, *outV = _CV_save ? _CV_save : CV;
    lchk(who, cudaMemcpy2DAsync(outU, (size_t)LDC * 8, g_split.dCU, (size_t)g_split.LDC * 8, (size_t)M * 8, new_UVrk, cudaMemcpyDeviceToHost, st));
    lchk(who, cudaMemcpy2DAsync(outV, (size_t)LDC * 8, g_split.dCV, (size_t)g_split.LDC * 8, (size_t)M * 8, new_UVrk, cudaMemcpyDeviceToHost, st));
    if (cudaStreamSynchronize(st) != cudaSuccess) legacy_fail(who, "CUDA failure", HCORE_B200_ERR_CUDA);
    if (flops) flops->update += ormqr_flops(M, g_split.n, new_UVrk);
    Crk[0] = new_UVrk;
    g_split.pending = false;
}

// reference src/compute/hcore_sgemm.c:703 (float clone of HCORE_dgemm; ranks travel as float)
void HCORE_sgemm(HCORE_enum transA, HCORE_enum transB, int M, int N, float alpha, float *AU, float *AV, float *Ark,
                 int LDA, float *BU, float *BV, float *Brk, int LDB, float beta, float *CU, float *CV, float *Crk,
                 int LDC, int rk, int maxrk, float acc, float *work, flop_counter *flops) {
    (void)transA; (void)transB; (void)N; (void)work;
    const int a = (int)Ark[0], b = (int)Brk[0], c = (int)Crk[0];
    const int rnew = legacy_gemm("HCORE_sgemm", 4, M, alpha, AU, AV, a, LDA, BU, BV, b, LDB, beta, CU, CV, c, LDC, rk, maxrk, acc);
    if (flops) flops->update += hcore_b200_gemm_flops(M, a, b, c, rnew);
    Crk[0] = (float)rnew;
}

// reference src/compute/hcore_zgemm.c:663-975.  `double _Complex` by value and a struct of two
// doubles are passed identically (two SSE registers) under the x86-64 SysV ABI.
void HCORE_zgemm(HCORE_enum transA, int transB, int M, int N, hcore_b200_dcomplex alpha, void *AU, void *AV,
                 double *Ark, int LDA, void *BU, void *BV, double *Brk, int LDB, hcore_b200_dcomplex beta, void *CU,
                 void *CV, double *Crk, int LDC, int rk, int maxrk, double acc, void *work, flop_counter *flops) {
    (void)transA; (void)transB; (void)N; (void)work;
    const char *who = "HCORE_zgemm";
    const int a = (int)Ark[0], b = (int)Brk[0], c = (int)Crk[0];
    hcore_b200_handle_t ctx = legacy_ctx(who);
    if (a == 0 || b == 0 || c == 0) legacy_fail(who, "rank values should not be zero", HCORE_B200_ERR_ZERO_RANK);
    if (maxrk < a + c) legacy_fail(who, "increase maxrank (maxrk < Ark + Crk)", HCORE_B200_ERR_MAXRK);
    cudaSetDevice(ctx->device);
    const size_t es = 16;
    const size_t fa = (size_t)LDA * a * es, fb = (size_t)LDB * b * es, fcin = (size_t)LDC * c * es, fc = (size_t)LDC * (a + c) * es;
    if (ctx->stage[0].reserve(2 * fa + 2 * fb + 2 * fc) != cudaSuccess || ctx->stage_ints[0].reserve(8 * sizeof(int)) != cudaSuccess)
        legacy_fail(who, "device allocation failed", HCORE_B200_ERR_CUDA);
    char *base = (char *)ctx->stage[0].p;
    char *dAU = base, *dAV = base + fa, *dBU = base + 2 * fa, *dBV = base + 2 * fa + fb, *dCU = base + 2 * fa + 2 * fb,
         *dCV = base + 2 * fa + 2 * fb + fc;
    int *di = (int *)ctx->stage_ints[0].p;
    int hi[4] = {a, b, c, 0};
    cudaStream_t st = ctx->streams[0];
    lchk(who, cudaMemcpyAsync(dAU, AU, fa, cudaMemcpyHostToDevice, st));
    lchk(who, cudaMemcpyAsync(dAV, AV, fa, cudaMemcpyHostToDevice, st));
    lchk(who, cudaMemcpyAsync(dBU, BU, fb, cudaMemcpyHostToDevice, st));
    lchk(who, cudaMemcpyAsync(dBV, BV, fb, cudaMemcpyHostToDevice, st));
    lchk(who, cudaMemcpyAsync(dCU, CU, fcin, cudaMemcpyHostToDevice, st));
    lchk(who, cudaMemcpyAsync(dCV, CV, fcin, cudaMemcpyHostToDevice, st));
    lchk(who, cudaMemcpyAsync(di, hi, sizeof(hi), cudaMemcpyHostToDevice, st));
    int rc = HCORE_zgemm_batched_strided(ctx, 1, M, alpha, dAU, dAV, di, LDA, 0, dBU, dBV, di + 1, LDB, 0, beta, dCU, dCV,
                                         di + 2, LDC, 0, rk, maxrk, acc, di + 3, st);
    if (rc != HCORE_B200_OK) legacy_fail(who, hcore_b200_last_error(ctx), rc);
    lchk(who, cudaMemcpyAsync(hi, di, sizeof(hi), cudaMemcpyDeviceToHost, st));
    if (cudaStreamSynchronize(st) != cudaSuccess) legacy_fail(who, "CUDA failure", HCORE_B200_ERR_CUDA);
    if (hi[3] != HCORE_B200_OK) legacy_fail(who, "tile update failed", hi[3]);
    const int rnew = hi[2];
    lchk(who, cudaMemcpyAsync(CU, dCU, (size_t)LDC * rnew * es, cudaMemcpyDeviceToHost, st));
    lchk(who, cudaMemcpyAsync(CV, dCV, (size_t)LDC * rnew * es, cudaMemcpyDeviceToHost, st));
    if (cudaStreamSynchronize(st) != cudaSuccess) legacy_fail(who, "CUDA failure", HCORE_B200_ERR_CUDA);
    if (flops) flops->update += hcore_b200_gemm_flops(M, a, b, c, rnew);
    Crk[0] = (double)rnew;
}

// Shared body of the legacy HCORE_dgemmcd / HCORE_dsyrk.  lda / ldb / ldc are the leading dimensions the REFERENCE
// uses for the arrays (dgemmcd: always M, hcore_dgemmcd.c:66-99; dsyrk: LDAU for AU and AV, LDCD for CD,
// hcore_dsyrk.c:57-58,:103,:144,:165-181), rows_av = rows of AV the reference contracts over.
static void legacy_dense(const char *who, int syrk, int M, double scale, const double *Au, const double *Av, int a, int lda,
                         int rows_av, const double *Bu, const double *Bv, int b, int ldb, double *CD, int ldc) {
    hcore_b200_handle_t ctx = legacy_ctx(who);
    cudaSetDevice(ctx->device);
    if (lda < M || ldb < M || ldc < M || rows_av > lda) legacy_fail(who, "leading dimension smaller than the tile", HCORE_B200_ERR_ARG);
    const size_t fa = (size_t)lda * a * 8, fb = (size_t)ldb * b * 8, fc = (size_t)ldc * M * 8;
    const size_t fa16 = (fa + 15) / 16 * 16, fb16 = (fb + 15) / 16 * 16;
    if (ctx->stage[0].reserve(2 * fa16 + 2 * fb16 + fc) != cudaSuccess || ctx->stage_ints[0].reserve(16) != cudaSuccess)
        legacy_fail(who, "device allocation failed", HCORE_B200_ERR_CUDA);
    char *base = (char *)ctx->stage[0].p;
    double *dAu = (double *)base, *dAv = (double *)(base + fa16), *dBu = (double *)(base + 2 * fa16),
           *dBv = (double *)(base + 2 * fa16 + fb16), *dC = (double *)(base + 2 * fa16 + 2 * fb16);
    int *di = (int *)ctx->stage_ints[0].p;
    int hi[2] = {a, b};
    cudaStream_t st = ctx->streams[0];
    // the arrays are copied with their padding rows: ld * (cols - 1) + rows elements each
    lchk(who, cudaMemcpyAsync(dAu, Au, ((size_t)lda * (a - 1) + M) * 8, cudaMemcpyHostToDevice, st));
    lchk(who, cudaMemcpyAsync(dAv, Av, ((size_t)lda * (a - 1) + rows_av) * 8, cudaMemcpyHostToDevice, st));
    if (!syrk) {
        lchk(who, cudaMemcpyAsync(dBu, Bu, ((size_t)ldb * (b - 1) + M) * 8, cudaMemcpyHostToDevice, st));
        lchk(who, cudaMemcpyAsync(dBv, Bv, ((size_t)ldb * (b - 1) + M) * 8, cudaMemcpyHostToDevice, st));
    }
    lchk(who, cudaMemcpy2DAsync(dC, (size_t)ldc * 8, CD, (size_t)ldc * 8, (size_t)M * 8, M, cudaMemcpyHostToDevice, st));
    lchk(who, cudaMemcpyAsync(di, hi, sizeof(hi), cudaMemcpyHostToDevice, st));
    int rc = syrk ? dense_launch(ctx, 1, M, scale, 1, dAu, dAv, di, 0, lda, dAu, dAv, di, 0, lda, rows_av, dC, 0, ldc, a, st)
                  : dense_launch(ctx, 1, M, scale, 0, dAu, dAv, di, 0, lda, dBu, dBv, di + 1, 0, ldb, M, dC, 0, ldc, b, st);
    if (rc != HCORE_B200_OK) legacy_fail(who, hcore_b200_last_error(ctx), rc);
    lchk(who, cudaMemcpy2DAsync(CD, (size_t)ldc * 8, dC, (size_t)ldc * 8, (size_t)M * 8, M, cudaMemcpyDeviceToHost, st));
    if (cudaStreamSynchronize(st) != cudaSuccess) legacy_fail(who, "CUDA failure", HCORE_B200_ERR_CUDA);
}

void HCORE_dgemmcd(HCORE_enum transA, HCORE_enum transB, int M, double alpha, double *Au, double *Av, int Arank,
                   int LDA, double *Bu, double *Bv, int Brank, int LDB, double beta, double *CD, int LDC, int maxrank,
                   double *work, flop_counter *flops) {
    // alpha, beta and the leading dimensions are ignored by the reference too: -1/+1 and ld = M for every array
    // (src/compute/hcore_dgemmcd.c:66-99); flops is never touched there.
    (void)transA; (void)transB; (void)alpha; (void)LDA; (void)LDB; (void)beta; (void)LDC; (void)maxrank;
    (void)work; (void)flops;
    if (Arank <= 0 || Brank <= 0) return;              // k = 0 products: CD unchanged (cblas_dgemm with beta = 1)
    legacy_dense("HCORE_dgemmcd", 0, M, -1.0, Au, Av, Arank, M, M, Bu, Bv, Brank, M, CD, M);
}

void HCORE_dsyrk(HCORE_enum uplo, HCORE_enum trans, int N, int K, double alpha, const double *AU, int LDAU,
                 const double *AV, int LDAV, double beta, double *CD, int LDCD, double *work, flop_counter *flops) {
    // uplo, trans, beta and LDAV are ignored by the reference (src/compute/hcore_dsyrk.c:47-189).  LDAU is the leading
    // dimension of BOTH factor arrays and the contraction length of AV^T AV (:57-58,:103), LDCD the one of CD
    // (:165-181): honoured here, so padded tiles (LDAU, LDCD > N) give the reference's result.
    (void)uplo; (void)trans; (void)LDAV; (void)beta; (void)work;
    legacy_dense("HCORE_dsyrk", 1, N, alpha, AU, AV, K, LDAU, LDAU, AU, AV, K, LDAU, CD, LDCD);
    if (flops) {   // reference hcore_dsyrk.c:104,145,172,188 with its contraction length N = LDAU
        const unsigned long n = (unsigned long)LDAU, k = (unsigned long)K, m = (unsigned long)N;
        flops->syrk += 2 * k * k * n + 2 * m * k * k + 2 * m * m * k;
    }
}

// ---- HCORE_dgemm_fast (reference src/compute/hcore_dgemm_fast.c:59-426) -----------------------------------------
// The same update as HCORE_dgemm with a leaner data flow on the CPU (triangular core product, no clones when the
// stack fits C's storage).  Observable differences, reproduced here: it asserts use_trmm == 1 and use_scratch == 1
// (:82-83), has no flop counter, no "new rank > maxrk" check, and its truncation scan starts at index 2 (:314).
void HCORE_dgemm_fast(HCORE_enum transA, int transB, int M, int N, double alpha, double *AU, double *AV, double *Ark,
                      int LDA, double *BU, double *BV, double *Brk, int LDB, double beta, double *CU, double *CV,
                      double *Crk, int LDC, int rk, int maxrk, double acc, double *work) {
    (void)transA; (void)transB; (void)N; (void)work;
    if (use_trmm != 1 || (&use_scratch != nullptr && use_scratch != 1)) {
        fprintf(stderr, "HCORE_dgemm_fast: Assertion `use_trmm == 1 && use_scratch == 1' failed.\n");   // reference :82-83
        abort();
    }
    const int a = (int)Ark[0], b = (int)Brk[0], c = (int)Crk[0];
    const int rnew = legacy_gemm("HCORE_dgemm_fast", 8, M, alpha, AU, AV, a, LDA, BU, BV, b, LDB, beta, CU, CV, c, LDC, rk,
                                 std::max(maxrk, a + c), acc, 2);
    Crk[0] = (double)rnew;
}

// ---- HCORE_dtrsm / HCORE_dpotrf / HCORE_dgemmbdcd / HCORE_duncompress / HCORE_zuncompress (host pointers) ------------
// reference src/compute/hcore_dtrsm.c:24-48
void HCORE_dtrsm(HCORE_enum side, HCORE_enum uplo, HCORE_enum transA, HCORE_enum diag, int M, int Brank, double alpha,
                 double *A, int LDA, double *BV, int LDBV, flop_counter *flops) {
    const char *who = "HCORE_dtrsm";
    if (side != 141 && side != 142) {                   // reference :47: assert(0 == "side is not CblasLeft or CblasRight")
        fprintf(stderr, "%s: side is not CblasLeft or CblasRight\n", who);
        abort();
    }
    if (M > 0 && Brank > 0) {
        hcore_b200_handle_t ctx = legacy_ctx(who);
        cudaSetDevice(ctx->device);
        const int na = (side == 141) ? M : Brank;
        const size_t fa = ((size_t)LDA * na * 8 + 15) / 16 * 16, fb = (size_t)LDBV * Brank * 8;
        if (ctx->stage[0].reserve(fa + fb) != cudaSuccess) legacy_fail(who, "device allocation failed", HCORE_B200_ERR_CUDA);
        double *dA = (double *)ctx->stage[0].p, *dB = (double *)((char *)ctx->stage[0].p + fa);
        cudaStream_t st = ctx->streams[0];
        lchk(who, cudaMemcpy2DAsync(dA, (size_t)LDA * 8, A, (size_t)LDA * 8, (size_t)na * 8, na, cudaMemcpyHostToDevice, st));
        lchk(who, cudaMemcpy2DAsync(dB, (size_t)LDBV * 8, BV, (size_t)LDBV * 8, (size_t)M * 8, Brank, cudaMemcpyHostToDevice, st));
        int rc = HCORE_dtrsm_batched_strided(ctx, 1, side, uplo, transA, diag, M, Brank, alpha, dA, LDA, 0, dB, nullptr, LDBV, 0, st);
        if (rc != HCORE_B200_OK) legacy_fail(who, hcore_b200_last_error(ctx), rc);
        lchk(who, cudaMemcpy2DAsync(BV, (size_t)LDBV * 8, dB, (size_t)LDBV * 8, (size_t)M * 8, Brank, cudaMemcpyDeviceToHost, st));
        if (cudaStreamSynchronize(st) != cudaSuccess) legacy_fail(who, "CUDA failure", HCORE_B200_ERR_CUDA);
    }
    if (flops) {                                        // reference :42-45 with flop_counts('t', ...)
        const unsigned long m = (unsigned long)M, n = (unsigned long)Brank;
        flops->trsm += (side == 141) ? n * m * m : m * n * n;
    }
}

// reference src/compute/hcore_dpotrf.c:23-35
void HCORE_dpotrf(HCORE_enum uplo, int N, double *A, int LDA, int *info, flop_counter *flops) {
    const char *who = "HCORE_dpotrf";
    *info = 0;
    // The reference forwards `uplo` unchanged to LAPACKE's `char uplo` (:29-32), so its working callers pass the
    // CHARACTER codes 'L' / 'U'; the HCORE_enum values are accepted here as well.  Anything else is rejected the way
    // LAPACKE_dpotrf_work rejects it: message, info = -2, matrix untouched.
    if (uplo == 'L' || uplo == 'l') uplo = HCORE_Lower;
    else if (uplo == 'U' || uplo == 'u') uplo = HCORE_Upper;
    if (uplo != HCORE_Lower && uplo != HCORE_Upper) {
        fprintf(stderr, " ** On entry to DPOTRF parameter number  1 had an illegal value\n");
        *info = -2;
    } else if (N > 0) {
        hcore_b200_handle_t ctx = legacy_ctx(who);
        cudaSetDevice(ctx->device);
        if (ctx->stage[0].reserve((size_t)LDA * N * 8) != cudaSuccess || ctx->stage_ints[0].reserve(16) != cudaSuccess)
            legacy_fail(who, "device allocation failed", HCORE_B200_ERR_CUDA);
        double *dA = (double *)ctx->stage[0].p;
        int *di = (int *)ctx->stage_ints[0].p;
        cudaStream_t st = ctx->streams[0];
        lchk(who, cudaMemcpy2DAsync(dA, (size_t)LDA * 8, A, (size_t)LDA * 8, (size_t)N * 8, N, cudaMemcpyHostToDevice, st));
        int rc = HCORE_dpotrf_batched_strided(ctx, 1, uplo, N, dA, LDA, 0, di, st);
        if (rc != HCORE_B200_OK) legacy_fail(who, hcore_b200_last_error(ctx), rc);
        lchk(who, cudaMemcpy2DAsync(A, (size_t)LDA * 8, dA, (size_t)LDA * 8, (size_t)N * 8, N, cudaMemcpyDeviceToHost, st));
        lchk(who, cudaMemcpyAsync(info, di, sizeof(int), cudaMemcpyDeviceToHost, st));
        if (cudaStreamSynchronize(st) != cudaSuccess) legacy_fail(who, "CUDA failure", HCORE_B200_ERR_CUDA);
    }
    if (flops) {                                        // flop_counts('c', N): n^3/3 - n^2/2.0 + n/6 in the reference's integer/double mix
        const unsigned long n = (unsigned long)N;
        flops->potrf += (unsigned long)(n * n * n / 3 - n * n / 2.0 + n / 6);
    }
}

// reference src/compute/hcore_dgemmbdcd.c:35-122
void HCORE_dgemmbdcd(HCORE_enum transA, HCORE_enum transB, int M, int N, double alpha, double *AU, double *AV, double *Ark,
                     int LDA, double *BD, int LDB, double beta, double *CD, int LDC, double *work, flop_counter *flops) {
    (void)transB; (void)work;
    const char *who = "HCORE_dgemmbdcd";
    const int a = (int)Ark[0];
    // the two branches of the reference differ only in which size they take for the contraction and for the rows of
    // CD: (M, LDA) when transA == Trans (:74-93), (LDB, M) otherwise (:95-114)
    const int mc = (transA == HCORE_Trans) ? M : LDB, rows = (transA == HCORE_Trans) ? LDA : M;
    if (a > 0 && rows > 0 && N > 0) {
        hcore_b200_handle_t ctx = legacy_ctx(who);
        cudaSetDevice(ctx->device);
        if (LDA < rows || LDA < mc || LDB < mc || LDC < rows) legacy_fail(who, "leading dimension smaller than the tile", HCORE_B200_ERR_ARG);
        const size_t fa = ((size_t)LDA * a * 8 + 15) / 16 * 16, fb = ((size_t)LDB * N * 8 + 15) / 16 * 16, fc = (size_t)LDC * N * 8;
        if (ctx->stage[0].reserve(2 * fa + fb + fc) != cudaSuccess || ctx->stage_ints[0].reserve(16) != cudaSuccess)
            legacy_fail(who, "device allocation failed", HCORE_B200_ERR_CUDA);
        char *base = (char *)ctx->stage[0].p;
        double *dAU = (double *)base, *dAV = (double *)(base + fa), *dBD = (double *)(base + 2 * fa), *dC = (double *)(base + 2 * fa + fb);
        int *di = (int *)ctx->stage_ints[0].p;
        cudaStream_t st = ctx->streams[0];
        lchk(who, cudaMemcpyAsync(dAU, AU, ((size_t)LDA * (a - 1) + rows) * 8, cudaMemcpyHostToDevice, st));
        lchk(who, cudaMemcpyAsync(dAV, AV, ((size_t)LDA * (a - 1) + mc) * 8, cudaMemcpyHostToDevice, st));
        lchk(who, cudaMemcpyAsync(dBD, BD, ((size_t)LDB * (N - 1) + mc) * 8, cudaMemcpyHostToDevice, st));
        if (beta != 0.0)
            lchk(who, cudaMemcpy2DAsync(dC, (size_t)LDC * 8, CD, (size_t)LDC * 8, (size_t)rows * 8, N, cudaMemcpyHostToDevice, st));
        lchk(who, cudaMemcpyAsync(di, &a, sizeof(int), cudaMemcpyHostToDevice, st));
        int rc = gemmbdcd_launch(ctx, 1, rows, mc, N, alpha, dAU, dAV, di, LDA, 0, dBD, LDB, 0, beta, dC, LDC, 0, st, a);
        if (rc != HCORE_B200_OK) legacy_fail(who, hcore_b200_last_error(ctx), rc);
        lchk(who, cudaMemcpy2DAsync(CD, (size_t)LDC * 8, dC, (size_t)LDC * 8, (size_t)rows * 8, N, cudaMemcpyDeviceToHost, st));
        if (cudaStreamSynchronize(st) != cudaSuccess) legacy_fail(who, "CUDA failure", HCORE_B200_ERR_CUDA);
    }
    if (flops) {                                        // reference :91-92 / :112-113
        const unsigned long k = (unsigned long)a, n = (unsigned long)N;
        flops->update += k * n * (unsigned long)mc * 2 + (unsigned long)rows * n * k * 2;
    }
}

// reference src/compute/hcore_duncompress.c:33-85
void HCORE_duncompress(HCORE_enum transA, HCORE_enum transB, int M, int N, double alpha, double *AU, double *Ark, int LDA,
                       double *BV, double *Brk, int LDB, double beta, double *CD, int LDC) {
    (void)transA; (void)transB; (void)alpha; (void)beta;      // never read by the reference: the product is formed with 1 / 0
    const char *who = "HCORE_duncompress";
    const int a = (int)Ark[0], b = (int)Brk[0];
    if (a != b) {                                       // reference :69: assert(_Ark == _Brk)
        fprintf(stderr, "%s: Assertion `_Ark == _Brk' failed.\n", who);
        abort();
    }
    if (M <= 0 || N <= 0) return;
    hcore_b200_handle_t ctx = legacy_ctx(who);
    cudaSetDevice(ctx->device);
    const size_t fa = ((size_t)LDA * std::max(a, 1) * 8 + 15) / 16 * 16, fb = ((size_t)LDB * std::max(a, 1) * 8 + 15) / 16 * 16, fc = (size_t)LDC * N * 8;
    if (ctx->stage[0].reserve(fa + fb + fc) != cudaSuccess || ctx->stage_ints[0].reserve(16) != cudaSuccess)
        legacy_fail(who, "device allocation failed", HCORE_B200_ERR_CUDA);
    char *base = (char *)ctx->stage[0].p;
    double *dAU = (double *)base, *dBV = (double *)(base + fa), *dC = (double *)(base + fa + fb);
    int *di = (int *)ctx->stage_ints[0].p;
    int hi[2] = {a, b};
    cudaStream_t st = ctx->streams[0];
    if (a > 0) {
        lchk(who, cudaMemcpyAsync(dAU, AU, ((size_t)LDA * (a - 1) + M) * 8, cudaMemcpyHostToDevice, st));
        lchk(who, cudaMemcpyAsync(dBV, BV, ((size_t)LDB * (a - 1) + N) * 8, cudaMemcpyHostToDevice, st));
    }
    lchk(who, cudaMemcpyAsync(di, hi, sizeof(hi), cudaMemcpyHostToDevice, st));
    int rc = HCORE_duncompress_batched_strided(ctx, 1, M, N, dAU, di, LDA, 0, dBV, di + 1, LDB, 0, dC, LDC, 0, nullptr, st);
    if (rc != HCORE_B200_OK) legacy_fail(who, hcore_b200_last_error(ctx), rc);
    lchk(who, cudaMemcpy2DAsync(CD, (size_t)LDC * 8, dC, (size_t)LDC * 8, (size_t)M * 8, N, cudaMemcpyDeviceToHost, st));
    if (cudaStreamSynchronize(st) != cudaSuccess) legacy_fail(who, "CUDA failure", HCORE_B200_ERR_CUDA);
}

// reference src/compute/hcore_zuncompress.c:35-95
void HCORE_zuncompress(HCORE_enum transA, HCORE_enum transB, int M, int N, hcore_b200_dcomplex alpha, void *AU, double *Ark,
                       int LDA, void *BV, double *Brk, int LDB, hcore_b200_dcomplex beta, void *CD, int LDC) {
    (void)transA; (void)transB; (void)alpha; (void)beta;
    const char *who = "HCORE_zuncompress";
    const int a = (int)Ark[0], b = (int)Brk[0];
    if (a != b) {                                       // reference :77: assert(_Ark == _Brk)
        fprintf(stderr, "%s: Assertion `_Ark == _Brk' failed.\n", who);
        abort();
    }
    if (M <= 0 || N <= 0) return;
    hcore_b200_handle_t ctx = legacy_ctx(who);
    cudaSetDevice(ctx->device);
    const size_t fa = (size_t)LDA * std::max(a, 1) * 16, fb = (size_t)LDB * std::max(a, 1) * 16, fc = (size_t)LDC * N * 16;
    if (ctx->stage[0].reserve(fa + fb + fc) != cudaSuccess || ctx->stage_ints[0].reserve(16) != cudaSuccess)
        legacy_fail(who, "device allocation failed", HCORE_B200_ERR_CUDA);
    char *base = (char *)ctx->stage[0].p;
    void *dAU = base, *dBV = base + fa, *dC = base + fa + fb;
    int *di = (int *)ctx->stage_ints[0].p;
    int hi[2] = {a, b};
    cudaStream_t st = ctx->streams[0];
    if (a > 0) {
        lchk(who, cudaMemcpyAsync(dAU, AU, ((size_t)LDA * (a - 1) + M) * 16, cudaMemcpyHostToDevice, st));
        lchk(who, cudaMemcpyAsync(dBV, BV, ((size_t)LDB * (a - 1) + N) * 16, cudaMemcpyHostToDevice, st));
    }
    lchk(who, cudaMemcpyAsync(di, hi, sizeof(hi), cudaMemcpyHostToDevice, st));
    int rc = HCORE_zuncompress_batched_strided(ctx, 1, M, N, dAU, di, LDA, 0, dBV, di + 1, LDB, 0, dC, LDC, 0, nullptr, st);
    if (rc != HCORE_B200_OK) legacy_fail(who, hcore_b200_last_error(ctx), rc);
    lchk(who, cudaMemcpy2DAsync(CD, (size_t)LDC * 16, dC, (size_t)LDC * 16, (size_t)M * 16, N, cudaMemcpyDeviceToHost, st));
    if (cudaStreamSynchronize(st) != cudaSuccess) legacy_fail(who, "CUDA failure", HCORE_B200_ERR_CUDA);
}

}  // extern "C"
