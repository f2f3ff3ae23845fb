// Tall-skinny weight-gradient GEMM  C[M,K] = A^T B  (A: [R,M], B: [R,K], R ~ 1e6 rows, M,K in
// {64,128}) on the tcgen05 tensor cores, 3xTF32 (fp32-grade).  This is dW = dH^T X of every GCN layer
// and of the hoisted scorer layer (autograd's mm backward behind pangnn.py:207).
//
// The contraction runs over the ROWS, so both operands are needed "transposed": a K-major tensor-core
// operand wants, for every output row m, the contraction index contiguous.  (For tf32 the tensor
// core accepts MN-major operands only in the 128B_BASE32B swizzle, see umma.cuh; the transposition
// is therefore done on the way into shared memory.)  A warp loads 4 rows x 128 B with coalesced
// float4 loads and scatters the 4-byte elements into the chunk-interleaved layout
//         offset(m, r) = (r / 4) * CHUNK + m * 16 + (r % 4) * 4
// rotating the element order per lane so that the 32 lanes of every store hit 32 distinct banks.
// Each CTA accumulates the whole M x K tile in TMEM over its 32-row tiles and drains it to a global
// partial every kFlush tiles: the tensor core adds into its fp32 accumulator with truncation, an
// error that grows with the chain length (measured ~2e-8 relative per tcgen05.mma), so chains are
// kept short.  All partials are summed in fixed order (fp64) by reduce_partials — deterministic, no
// atomics.
//
// Roofline: HBM (4 (M + K) bytes per row; 6 M K tensor-core FLOP per row is ~10x under the tf32 peak).
#include "common.cuh"
#include "umma.cuh"

namespace pangnn {

int reduce_partials(const float *partial, int64_t nblocks, int32_t width, int32_t stride, float *out,
                    cudaStream_t st);

namespace {

constexpr int kTR = 32;                  // rows (contraction length) per tile
constexpr int kThreadsTN = 256;
constexpr int kFlush = 16;               // tiles per accumulation chain (16 x 4 k-steps x 3 products)

// CA = columns of the A-side source matrix (128, or 64 when STACK), NB = columns of the B side.
// STACK: the A operand is [A_hi ; A_lo] stacked to 128 rows (M = K = 64): 2 products per k-step and
// C = D[0:64] + D[64:128]; otherwise A_hi / A_lo are separate 128-row operands and 3 products.
template <int CA, int NB, bool STACK>
struct TNSmem {
    static constexpr uint32_t chA = 128 * 16 + 16;                       // A operand always has 128 rows
    static constexpr uint32_t chB = NB * 16 + 16;
    static constexpr uint32_t bytesA = (kTR / 4) * chA;                   // one 128-row operand
    static constexpr uint32_t bytesB = (kTR / 4) * chB;
    static constexpr uint32_t oAh = 0, oAl = STACK ? 0 : bytesA;          // STACK: lo = rows 64..127 of the same operand
    static constexpr uint32_t oBh = STACK ? bytesA : 2 * bytesA, oBl = oBh + bytesB;
    static constexpr uint32_t total = oBl + bytesB + 64;
};

// coalesced load of this thread's float4s of a [kTR x C] tile: item = (row group of 4) x (128-byte segment)
template <int C>
__device__ __forceinline__ void tile_load(const float *__restrict__ src, int64_t ld, int64_t r0, int64_t R,
                                          int warp, int lane, float4 (&v)[C / 32]) {
    constexpr int SEG = C / 32;                                           // segments per row
    const int fl = lane & 7, rr = lane >> 3;
#pragma unroll
    for (int i = 0; i < SEG; ++i) {
        const int item = warp + 8 * i;                                    // 0 .. 8 * SEG - 1
        const int g = item / SEG, seg = item % SEG;
        const int64_t row = r0 + 4 * g + rr;
        v[i] = row < R ? ld_stream_f4(src + row * ld + seg * 32 + fl * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// transposed, bank-conflict-free store of those float4s: hi at off_hi (+ m * 16), lo at off_lo
template <int C>
__device__ __forceinline__ void tile_store_t(uint8_t *smem, uint32_t off_hi, uint32_t off_lo, uint32_t chunk,
                                             int warp, int lane, const float4 (&v)[C / 32]) {
    constexpr int SEG = C / 32;
    const int fl = lane & 7, rr = lane >> 3;
    const int s = (fl >> 1) & 3;                                          // per-lane rotation of the 4 elements
#pragma unroll
    for (int i = 0; i < SEG; ++i) {
        const int item = warp + 8 * i;
        const int g = item / SEG, seg = item % SEG;
        float e[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
        if (s & 1) { const float t = e[0]; e[0] = e[1]; e[1] = e[2]; e[2] = e[3]; e[3] = t; }
        if (s & 2) { float t = e[0]; e[0] = e[2]; e[2] = t; t = e[1]; e[1] = e[3]; e[3] = t; }
        // now e[j] = original element (j + s) & 3
        const uint32_t base = (uint32_t)g * chunk + (uint32_t)rr * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = seg * 32 + fl * 4 + ((j + s) & 3);
            const float hi = umma::tf32_hi(e[j]);
            *reinterpret_cast<float *>(smem + off_hi + base + (uint32_t)m * 16) = hi;
            *reinterpret_cast<float *>(smem + off_lo + base + (uint32_t)m * 16) = umma::tf32_lo(e[j], hi);
        }
    }
}

template <int CA, int NB, bool STACK>
__global__ void __launch_bounds__(kThreadsTN)
gemm_tn_tc_kernel(const float *__restrict__ A, int64_t lda, const float *__restrict__ B, int64_t ldb, int64_t R,
                  int transpose_out, int slots_per_cta, float *__restrict__ partial) {
    using S = TNSmem<CA, NB, STACK>;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sb = umma::smem_u32(smem);
    constexpr uint32_t kTmemCols = NB;                                    // 64 or 128
    constexpr uint32_t idesc = umma::idesc_tf32(128, NB, false, false);
    // STACK: lo rows sit 64 rows (64 * 16 bytes) below the hi rows inside each chunk
    constexpr uint32_t offAl = STACK ? 64 * 16 : S::oAl;

    if (warp == 0) umma::tmem_alloc(&tmem_base_s, kTmemCols);
    if (tid == 32) {
        umma::mbar_init(&bar, 1);
        umma::fence_mbar_init();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tD = tmem_base_s;

    const int64_t num_tiles = (R + kTR - 1) / kTR;
    const int Mc = STACK ? 64 : (transpose_out ? NB : 128);               // rows / cols of C
    const int Kc = STACK ? 64 : (transpose_out ? 128 : NB);
    const int q = warp & 3, h = warp >> 2;
    const int ma = q * 32 + lane;                                         // D row held by this thread
    constexpr int COLS = NB / 2;
    float *red = reinterpret_cast<float *>(smem);                         // STACK: [128][64] floats

    // D (or zeros when `have` is false) -> partial slot; every thread of the CTA calls this
    auto drain = [&](int slot, bool have) {
        float *out = partial + ((int64_t)blockIdx.x * slots_per_cta + slot) * Mc * Kc;
#pragma unroll
        for (int part = 0; part < COLS / 32; ++part) {
            const int c0 = h * COLS + part * 32;
            float v[32];
            if (have) {
                umma::tmem_ld32(tD + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = 0.f;
            }
            if (STACK) {
#pragma unroll
                for (int c = 0; c < 32; ++c) red[ma * 64 + c0 + c] = v[c];
            } else if (!transpose_out) {
#pragma unroll
                for (int c = 0; c < 32; c += 4)
                    *reinterpret_cast<float4 *>(out + (int64_t)ma * Kc + c0 + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) out[(int64_t)(c0 + c) * Kc + ma] = v[c];   // C[m = nb][k = ma]
            }
        }
        if (STACK) {
            __syncthreads();
            for (int i = tid; i < 64 * 64; i += kThreadsTN) out[i] = red[i] + red[64 * 64 + i];
        }
        umma::fence_before_sync();
        __syncthreads();                                                  // D / red may be rewritten after this
    };

    float4 va[CA / 32], vb[NB / 32];
    uint32_t commits = 0;
    int in_chain = 0, slot = 0;                                           // tiles in the open chain, next partial slot
    int64_t tile = blockIdx.x;
    if (tile < num_tiles) {
        tile_load<CA>(A, lda, tile * kTR, R, warp, lane, va);
        tile_load<NB>(B, ldb, tile * kTR, R, warp, lane, vb);
    }
    for (; tile < num_tiles; tile += gridDim.x) {
        if (commits > 0) umma::mbar_wait(&bar, (commits - 1) & 1);        // previous MMAs finished reading smem
        if (in_chain == kFlush) {                                         // close the chain (uniform branch)
            umma::fence_after_sync();
            drain(slot++, true);
            in_chain = 0;
        }
        tile_store_t<CA>(smem, S::oAh, offAl, S::chA, warp, lane, va);
        tile_store_t<NB>(smem, S::oBh, S::oBl, S::chB, warp, lane, vb);
        umma::fence_async_smem();
        __syncthreads();
        const int64_t next = tile + gridDim.x;
        if (next < num_tiles) {                                           // in flight while the tensor core works
            tile_load<CA>(A, lda, next * kTR, R, warp, lane, va);
            tile_load<NB>(B, ldb, next * kTR, R, warp, lane, vb);
        }
        if (tid == 0) {
            umma::fence_after_sync();
            const uint64_t bh0 = umma::smem_desc(sb + S::oBh, S::chB, 128), bl0 = umma::smem_desc(sb + S::oBl, S::chB, 128);
            const uint64_t ah0 = umma::smem_desc(sb + S::oAh, S::chA, 128);
            const uint64_t al0 = umma::smem_desc(sb + S::oAl, S::chA, 128);          // unused when STACK
#pragma unroll
            for (int s = 0; s < kTR / 8; ++s) {
                const uint64_t bh = umma::desc_advance(bh0, s * 2 * S::chB), bl = umma::desc_advance(bl0, s * 2 * S::chB);
                const uint64_t ah = umma::desc_advance(ah0, s * 2 * S::chA);
                const uint32_t acc0 = (in_chain > 0 || s > 0) ? 1u : 0u;
                if (STACK) {
                    umma::mma_tf32(tD, ah, bl, idesc, acc0);
                    umma::mma_tf32(tD, ah, bh, idesc, 1u);
                } else {
                    const uint64_t al = umma::desc_advance(al0, s * 2 * S::chA);
                    umma::mma_tf32(tD, al, bh, idesc, acc0);
                    umma::mma_tf32(tD, ah, bl, idesc, 1u);
                    umma::mma_tf32(tD, ah, bh, idesc, 1u);
                }
            }
            umma::mma_commit(&bar);
        }
        ++commits;
        ++in_chain;
    }
    // ---- last chain, then zero the slots this CTA did not need
    if (commits > 0) {
        umma::mbar_wait(&bar, (commits - 1) & 1);
        umma::fence_after_sync();
    }
    if (in_chain > 0) drain(slot++, true);
    for (; slot < slots_per_cta; ++slot) drain(slot, false);
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tD, kTmemCols);
}

template <int CA, int NB, bool STACK>
struct TNLaunch {
    static int ctas_per_sm() {
        const int by_smem = (int)((227u * 1024u) / (TNSmem<CA, NB, STACK>::total + 1024u));
        return by_smem > 3 ? 3 : (by_smem < 1 ? 1 : by_smem);
    }
    static int grid(int64_t R) {
        const int64_t tiles = (R + kTR - 1) / kTR;
        const int64_t cap = (int64_t)kNumSMs * ctas_per_sm();
        return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
    }
    static int slots(int64_t R) {                                         // partial slots per CTA
        const int64_t tiles = (R + kTR - 1) / kTR;
        const int64_t per_cta = (tiles + grid(R) - 1) / grid(R);
        return (int)((per_cta + kFlush - 1) / kFlush);
    }
    static size_t partial_floats(int64_t R, int Mc, int Kc) { return (size_t)grid(R) * slots(R) * Mc * Kc; }
    static int run(const float *A, int64_t lda, const float *B, int64_t ldb, int64_t R, int transpose_out, int Mc,
                   int Kc, float *C, float *partial, cudaStream_t st) {
        using S = TNSmem<CA, NB, STACK>;
        static bool attr = false;
        if (!attr) {
            int rc = check_cuda(cudaFuncSetAttribute(gemm_tn_tc_kernel<CA, NB, STACK>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::total),
                                "cudaFuncSetAttribute(gemm_tn_tc)");
            if (rc) return rc;
            attr = true;
        }
        const int g = grid(R), sl = slots(R);
        gemm_tn_tc_kernel<CA, NB, STACK><<<g, kThreadsTN, S::total, st>>>(A, lda, B, ldb, R, transpose_out, sl, partial);
        PANGNN_CHECK_LAUNCH("gemm_tn_tc");
        return reduce_partials(partial, (int64_t)g * sl, Mc * Kc, Mc * Kc, C, st);
    }
};

}  // namespace
}  // namespace pangnn

using namespace pangnn;

extern "C" {

size_t pangnn_gemm_tn_workspace_bytes(int64_t N, int32_t M, int32_t K) {
    if (N <= 0) return 256;
    size_t f;
    if (M == 128 && K == 128) f = TNLaunch<128, 128, false>::partial_floats(N, M, K);
    else if (M == 64 && K == 64) f = TNLaunch<64, 64, true>::partial_floats(N, M, K);
    else f = TNLaunch<128, 64, false>::partial_floats(N, M, K);
    return f * sizeof(float) + 256;
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
    if (M == 128 && K == 128) return TNLaunch<128, 128, false>::run(A, lda, B, ldb, N, 0, M, K, C, partial, st);
    if (M == 128 && K == 64) return TNLaunch<128, 64, false>::run(A, lda, B, ldb, N, 0, M, K, C, partial, st);
    if (M == 64 && K == 128) return TNLaunch<128, 64, false>::run(B, ldb, A, lda, N, 1, M, K, C, partial, st);
    return TNLaunch<64, 64, true>::run(A, lda, B, ldb, N, 0, M, K, C, partial, st);
}

}  // extern "C"
