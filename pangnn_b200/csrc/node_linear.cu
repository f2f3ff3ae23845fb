// Node linear transform  Y = act(X W^T + b)  on the 5th-generation tensor cores (tcgen05.mma,
// kind::tf32, accumulators in TMEM) with the 3xTF32 split, i.e. fp32-grade results (the 1e-5 parity
// bar rules out plain TF32).  Replaces the X W^T inside every torch_geometric GCNConv call
// (src/gnn.py:129-165), linear_out (src/gnn.py:104,148) and the first scorer layer hoisted to the
// nodes (src/gnn.py:110,177), plus their dX = dY W backward products (pangnn.py:207).
//
// Shapes: X [M, K] with M ~ 1e6 rows, W [N, K] (or [K, N] for the backward product), N, K in
// {64, 128}: a real dense contraction, but with K this small the kernel is HBM-bound
// (4(K+N) bytes per row against 6 K N tensor-core FLOP): the design goal is to stream X once.
//
// One persistent CTA per SM slot: W is split into TF32 hi / lo once and stays in shared memory in
// the chunk-interleaved no-swizzle layout (umma.cuh); X row tiles of 128 x 32 are read with
// coalesced 128-bit loads, split in registers and written to a shared-memory stage; one thread
// issues the 3 x 4 tcgen05.mma of the chunk (M = 128, N, K = 8 each) and commits to an mbarrier; the
// next chunk's global loads are in flight meanwhile.  Epilogue: tcgen05.ld (thread = row), bias,
// ELU, 128-bit stores.
#include "common.cuh"
#include "umma.cuh"

namespace pangnn {

namespace {

constexpr int kTileM = 128;
constexpr int kKC = 32;                              // K columns per stage
constexpr int kThreadsNL = 256;
constexpr uint32_t kChunkA = kTileM * 16 + 16;       // bytes between 4-k chunks of the X stage (+16: bank spread)

__device__ __forceinline__ float elu1f(float x) { return x > 0.f ? x : expm1f(x); }

template <int N, int K>
struct NLSmem {
    static constexpr uint32_t chunkB = N * 16 + 16;
    static constexpr uint32_t bytesB = (K / 4) * chunkB;          // one of hi / lo
    static constexpr uint32_t bytesA = (kKC / 4) * kChunkA;       // one of hi / lo
    static constexpr uint32_t offBhi = 0, offBlo = bytesB, offAhi = 2 * bytesB, offAlo = 2 * bytesB + bytesA;
    static constexpr uint32_t total = 2 * bytesB + 2 * bytesA + 64;
    static constexpr int tmemCols = N;                            // 64 or 128 (powers of two >= 32)
};

template <int N, int K>
__global__ void __launch_bounds__(kThreadsNL)
node_linear_kernel(const float *__restrict__ x, int64_t ldx, int64_t M, const float *__restrict__ w,
                   int64_t ldw, int w_is_kn, const float *__restrict__ bias, int act,
                   float *__restrict__ y, int64_t ldy) {
    using S = NLSmem<N, K>;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = umma::smem_u32(smem);
    constexpr int NCH = K / kKC;

    // ---- one-time setup: TMEM, barrier, W -> smem (hi / lo, chunk-interleaved K-major [n][k])
    if (warp == 0) umma::tmem_alloc(&tmem_base_s, S::tmemCols);
    if (tid == 32) {
        umma::mbar_init(&mma_bar, 1);
        umma::fence_mbar_init();
    }
    for (int idx = tid; idx < N * K; idx += kThreadsNL) {
        int n, k;
        float v;
        if (!w_is_kn) { n = idx / K; k = idx % K; v = w[(int64_t)n * ldw + k]; }
        else          { k = idx / N; n = idx % N; v = w[(int64_t)k * ldw + n]; }
        const float hi = umma::tf32_hi(v);
        const uint32_t off = (uint32_t)(k >> 2) * S::chunkB + (uint32_t)n * 16 + (uint32_t)(k & 3) * 4;
        *reinterpret_cast<float *>(smem + S::offBhi + off) = hi;
        *reinterpret_cast<float *>(smem + S::offBlo + off) = umma::tf32_lo(v, hi);
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_d = tmem_base_s;

    const int64_t num_tiles = (M + kTileM - 1) / kTileM;
    const int f = tid & 7, rb = tid >> 3;                        // float4 slot in the 32-k chunk, base row
    constexpr uint32_t idesc = umma::idesc_tf32(kTileM, N, false, false);

    float4 pre[4];
    auto prefetch = [&](int64_t tile, int c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t row = tile * kTileM + rb + 32 * i;
            pre[i] = row < M ? ld_stream_f4(x + row * ldx + c * kKC + f * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };

    uint32_t commits = 0;
    int64_t tile = blockIdx.x;
    if (tile < num_tiles) prefetch(tile, 0);
    for (; tile < num_tiles; tile += gridDim.x) {
        for (int c = 0; c < NCH; ++c) {
            if (commits > 0) umma::mbar_wait(&mma_bar, (commits - 1) & 1);   // stage free (previous MMAs done)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 hi, lo;
                umma::split4(pre[i], hi, lo);
                const uint32_t off = (uint32_t)f * kChunkA + (uint32_t)(rb + 32 * i) * 16;
                *reinterpret_cast<float4 *>(smem + S::offAhi + off) = hi;
                *reinterpret_cast<float4 *>(smem + S::offAlo + off) = lo;
            }
            umma::fence_async_smem();
            umma::fence_before_sync();      // orders this thread's earlier tcgen05.ld before the barrier
            __syncthreads();
            // next item's loads fly while the tensor core works
            if (c + 1 < NCH) prefetch(tile, c + 1);
            else if (tile + gridDim.x < num_tiles) prefetch(tile + gridDim.x, 0);
            if (tid == 0) {
                umma::fence_after_sync();
                umma::mma_3xtf32(tmem_d, sbase + S::offAhi, sbase + S::offAlo,
                                 sbase + S::offBhi + (uint32_t)c * (kKC / 4) * S::chunkB,
                                 sbase + S::offBlo + (uint32_t)c * (kKC / 4) * S::chunkB,
                                 kChunkA, 128, 2 * kChunkA, S::chunkB, 128, 2 * S::chunkB, kKC / 8, idesc, c > 0);
                umma::mma_commit(&mma_bar);
            }
            ++commits;
        }
        // ---- epilogue: accumulator complete -> registers -> bias / ELU -> global
        umma::mbar_wait(&mma_bar, (commits - 1) & 1);
        umma::fence_after_sync();
        {
            const int q = warp & 3, h = warp >> 2;
            const int64_t row = tile * kTileM + q * 32 + lane;
            constexpr int COLS = N / 2;                           // columns per warp half
#pragma unroll
            for (int part = 0; part < COLS / 32; ++part) {
                const int c0 = h * COLS + part * 32;
                float v[32];
                umma::tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                if (row < M) {
                    float *dst = y + row * ldy + c0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        if (bias) {
                            const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + c0 + j));
                            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                        }
                        if (act == PANGNN_ACT_ELU) { o.x = elu1f(o.x); o.y = elu1f(o.y); o.z = elu1f(o.z); o.w = elu1f(o.w); }
                        *reinterpret_cast<float4 *>(dst + j) = o;
                    }
                }
            }
        }
    }
    // ---- teardown
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_d, S::tmemCols);
}

template <int N, int K>
int launch_node_linear(const float *x, int64_t ldx, int64_t M, const float *w, int64_t ldw, int w_is_kn,
                       const float *bias, int act, float *y, int64_t ldy, cudaStream_t st) {
    using S = NLSmem<N, K>;
    static bool attr_set = false;
    if (!attr_set) {
        int rc = check_cuda(cudaFuncSetAttribute(node_linear_kernel<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)S::total), "cudaFuncSetAttribute(node_linear)");
        if (rc) return rc;
        attr_set = true;
    }
    const int per_sm = (int)((227u * 1024u) / (S::total + 1024u));
    const int64_t tiles = (M + kTileM - 1) / kTileM;
    const int64_t cap = (int64_t)kNumSMs * (per_sm < 1 ? 1 : (per_sm > 3 ? 3 : per_sm));
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    node_linear_kernel<N, K><<<grid, kThreadsNL, S::total, st>>>(x, ldx, M, w, ldw, w_is_kn, bias, act, y, ldy);
    PANGNN_CHECK_LAUNCH("node_linear");
    return PANGNN_OK;
}

}  // namespace
}  // namespace pangnn

using namespace pangnn;

extern "C" int pangnn_node_linear(const float *x, int64_t ldx, int64_t num_rows, int32_t k, const float *w,
                                  int64_t ldw, int w_is_kn, int32_t n, const float *bias, int act, float *y,
                                  int64_t ldy, void *stream) {
    PANGNN_REQUIRE(num_rows >= 0, "negative row count");
    if (num_rows == 0) return PANGNN_OK;
    PANGNN_REQUIRE(x && w && y, "null pointer");
    PANGNN_REQUIRE((n == 64 || n == 128) && (k == 64 || k == 128), "n and k must be 64 or 128");
    PANGNN_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0 && ldx >= k && ldy >= n, "bad row stride");
    PANGNN_REQUIRE(ldw >= (w_is_kn ? n : k), "bad weight stride");
    PANGNN_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0 && (!bias || (uintptr_t)bias % 16 == 0),
                   "pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 64 && k == 64) return launch_node_linear<64, 64>(x, ldx, num_rows, w, ldw, w_is_kn, bias, act, y, ldy, st);
    if (n == 128 && k == 64) return launch_node_linear<128, 64>(x, ldx, num_rows, w, ldw, w_is_kn, bias, act, y, ldy, st);
    if (n == 64 && k == 128) return launch_node_linear<64, 128>(x, ldx, num_rows, w, ldw, w_is_kn, bias, act, y, ldy, st);
    return launch_node_linear<128, 128>(x, ldx, num_rows, w, ldw, w_is_kn, bias, act, y, ldy, st);
}
