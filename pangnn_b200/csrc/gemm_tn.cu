// Tall-skinny weight-gradient GEMM  C[M,K] = A^T B,  A: [N,M], B: [N,K], N ~ 1e6, M,K in {64,128}.
// This is dW = dH^T X of every GCN layer and of the hoisted scorer layer (autograd's mm backward
// behind pangnn.py:207).  cuBLAS runs these at 26-74 CTAs (split-K) and 0.6-1.2 ms; the reduction
// over N is what parallelises, so: persistent CTAs stream 32-row chunks of A and B through shared
// memory (cp.async, double-buffered), every CTA accumulates the FULL M x K tile in registers
// (TM x TK per thread, 16 x 16 threads), per-CTA partials are summed by a fixed-order second stage
// (deterministic, no atomics).
// Roofline: fp32 FMA pipe (2 N M K FLOP vs 4 N (M+K) bytes: 32-64 FLOP/B), ~2x under HBM time.
#include <cuda_pipeline.h>

#include "common.cuh"

namespace pangnn {

int reduce_partials(const float *partial, int64_t nblocks, int32_t width, int32_t stride, float *out,
                    cudaStream_t st);

constexpr int kRows = 32;            // rows per chunk
constexpr int kGemmThreads = 256;

template <int TM, int TK>
__global__ void __launch_bounds__(kGemmThreads)
gemm_tn_kernel(const float *__restrict__ A, int64_t lda, const float *__restrict__ B, int64_t ldb,
               int64_t N, float *__restrict__ partial) {
    constexpr int M = 16 * TM, K = 16 * TK;
    extern __shared__ __align__(16) float sm[];
    float *sA[2] = {sm, sm + kRows * M};
    float *sB[2] = {sm + 2 * kRows * M, sm + 2 * kRows * M + kRows * K};
    const int tid = threadIdx.x;
    const int tm = tid >> 4, tk = tid & 15;
    float acc[TM][TK];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TK; ++j) acc[i][j] = 0.f;

    const int64_t nchunks = (N + kRows - 1) / kRows;
    auto issue = [&](int64_t chunk, int buf) {
        const int64_t r0 = chunk * kRows;
        for (int i = tid; i < kRows * (M / 4); i += kGemmThreads) {
            const int r = i / (M / 4), c = i % (M / 4);
            float *dst = sA[buf] + r * M + c * 4;
            if (r0 + r < N) __pipeline_memcpy_async(dst, A + (r0 + r) * lda + c * 4, 16);
            else *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int i = tid; i < kRows * (K / 4); i += kGemmThreads) {
            const int r = i / (K / 4), c = i % (K / 4);
            float *dst = sB[buf] + r * K + c * 4;
            if (r0 + r < N) __pipeline_memcpy_async(dst, B + (r0 + r) * ldb + c * 4, 16);
            else *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __pipeline_commit();
    };

    int buf = 0;
    int64_t chunk = blockIdx.x;
    if (chunk < nchunks) issue(chunk, 0);
    for (; chunk < nchunks; chunk += gridDim.x) {
        const int64_t next = chunk + gridDim.x;
        if (next < nchunks) {
            issue(next, buf ^ 1);
            __pipeline_wait_prior(1);
        } else {
            __pipeline_wait_prior(0);
        }
        __syncthreads();
        const float *a = sA[buf] + tm * TM, *b = sB[buf] + tk * TK;
#pragma unroll 8
        for (int r = 0; r < kRows; ++r) {
            float av[TM], bv[TK];
#pragma unroll
            for (int i = 0; i < TM; i += 4)
                *reinterpret_cast<float4 *>(av + i) = *reinterpret_cast<const float4 *>(a + r * M + i);
#pragma unroll
            for (int j = 0; j < TK; j += 4)
                *reinterpret_cast<float4 *>(bv + j) = *reinterpret_cast<const float4 *>(b + r * K + j);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TK; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
        buf ^= 1;
    }
    float *out = partial + (int64_t)blockIdx.x * (M * K);
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TK; j += 4)
            *reinterpret_cast<float4 *>(out + (tm * TM + i) * K + tk * TK + j) =
                make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
}

static int gemm_tn_grid(int64_t N) {
    const int64_t nchunks = (N + kRows - 1) / kRows;
    const int64_t cap = 2 * kNumSMs;
    return (int)(nchunks < cap ? (nchunks > 0 ? nchunks : 1) : cap);
}

template <int TM, int TK>
static int launch_gemm_tn(const float *A, int64_t lda, const float *B, int64_t ldb, int64_t N, float *C,
                          float *partial, cudaStream_t st) {
    constexpr int M = 16 * TM, K = 16 * TK;
    const size_t smem = (size_t)2 * kRows * (M + K) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        int rc = check_cuda(cudaFuncSetAttribute(gemm_tn_kernel<TM, TK>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(gemm_tn)");
        if (rc) return rc;
        attr = true;
    }
    const int grid = gemm_tn_grid(N);
    gemm_tn_kernel<TM, TK><<<grid, kGemmThreads, smem, st>>>(A, lda, B, ldb, N, partial);
    PANGNN_CHECK_LAUNCH("gemm_tn");
    return reduce_partials(partial, grid, M * K, M * K, C, st);
}

}  // namespace pangnn

using namespace pangnn;

extern "C" {

size_t pangnn_gemm_tn_workspace_bytes(int64_t N, int32_t M, int32_t K) {
    return (size_t)gemm_tn_grid(N) * M * K * sizeof(float) + 256;
}

int pangnn_gemm_tn(const float *A, int64_t lda, const float *B, int64_t ldb, int64_t N, int32_t M,
                   int32_t K, float *C, void *ws, size_t ws_bytes, void *stream) {
    PANGNN_REQUIRE(C && ws, "null pointer");
    PANGNN_REQUIRE((M == 64 || M == 128) && (K == 64 || K == 128), "M and K must be 64 or 128");
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 0) return check_cuda(cudaMemsetAsync(C, 0, (size_t)M * K * sizeof(float), st), "memset");
    PANGNN_REQUIRE(A && B, "null pointer");
    PANGNN_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && (uintptr_t)A % 16 == 0 && (uintptr_t)B % 16 == 0,
                   "rows must be 16-byte aligned");
    if (ws_bytes < pangnn_gemm_tn_workspace_bytes(N, M, K)) {
        set_error("gemm_tn: workspace too small");
        return PANGNN_EWORKSPACE;
    }
    float *partial = static_cast<float *>(ws);
    if (M == 64 && K == 64) return launch_gemm_tn<4, 4>(A, lda, B, ldb, N, C, partial, st);
    if (M == 64 && K == 128) return launch_gemm_tn<4, 8>(A, lda, B, ldb, N, C, partial, st);
    if (M == 128 && K == 64) return launch_gemm_tn<8, 4>(A, lda, B, ldb, N, C, partial, st);
    return launch_gemm_tn<8, 8>(A, lda, B, ldb, N, C, partial, st);
}

}  // extern "C"
