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
// One persistent CTA per SM: W is split into TF32 hi / lo once and stays in shared memory in the
// chunk-interleaved no-swizzle layout (umma.cuh); X row tiles of 128 x 32 stream through a cp.async
// ring (up to 5 chunks = 80 KB in flight per SM — the first version, one chunk of register prefetch,
// was latency-bound at 1.6-2.8 TB/s), are split in registers and written to the operand stage; one
// thread issues the 3 x 4 tcgen05.mma of the chunk (M = 128, N, K = 8 each) and commits to an
// mbarrier.  Two TMEM accumulators ping-pong so that the epilogue (tcgen05.ld, thread = row, bias,
// ELU, 128-bit stores) of a tile overlaps the next tile's MMAs.  Warp-specialised: warps 0-7 move and
// convert data and run the epilogues, warp 8 only issues MMAs; they meet on mbarriers (operand stage
// ready / stage free / accumulator full / accumulator drained), never on a CTA-wide barrier.
// Measured and rejected (round 2): four DEDICATED epilogue warps (one per TMEM lane quadrant) beside the eight
// producers — the epilogue (two accumulators out of tensor memory, staging transpose, stores) is the larger half of
// a tile's instruction work, and four warps are slower at it than the eight producer warps in turn
// (N = 128: 0.39-0.45 ms vs 0.27-0.41 ms per call at 1e6 rows; only N = 64, K = 128 gained, 0.24 vs 0.26 ms).
#include "common.cuh"
#include "umma.cuh"

namespace pangnn {

namespace {

constexpr int kTileM = 128;
constexpr int kKC = 32;                              // K columns per stage
constexpr int kThreadsNL = 256;                      // producer / epilogue threads (warps 0-7)
constexpr int kThreadsNLAll = kThreadsNL + 32;      // + warp 8: the MMA issue warp
constexpr uint32_t kChunkA = kTileM * 16 + 16;       // bytes between 4-k chunks of the X stage (+16: bank spread)

// ELU in the GEMM epilogue: expm1f costs ~25 instructions and was 77 % of this kernel's issue slots
// (ncu); ex2.approx-based exp(x) - 1 has ~1e-7 ABSOLUTE error, far inside the 1e-5 (of the tensor's
// scale) bar on activations.
__device__ __forceinline__ float elu1f(float x) { return x > 0.f ? x : __expf(x) - 1.0f; }

// Fused GEMM -> halo all-gather of the genome-partitioned path: rows that a neighbouring rank needs
// (slot[row] >= 0) are ALSO stored, from the same coalesced epilogue, into that rank's extended
// activation buffer through its NVLink peer mapping.  Up to two peers (left / right neighbour of a
// contiguous genome block); the row stride of the peer buffers equals ldy.
struct NLPush {
    const int32_t *slot0, *slot1;      // [M] destination row in the peer's buffer, or -1
    float *peer0, *peer1;
    int64_t lo0, hi0, lo1, hi1;        // rows outside [lo, hi) are known not to be pushed (no map lookup)
};

template <int N, int K>
struct NLSmem {
    static constexpr uint32_t chunkB = N * 16 + 16;
    static constexpr uint32_t bytesB = (K / 4) * chunkB;          // one of hi / lo
    static constexpr uint32_t bytesA = (kKC / 4) * kChunkA;       // one of hi / lo
    static constexpr uint32_t ringSlot = kTileM * kKC * 4;        // 16 KB raw fp32 chunk
    static constexpr uint32_t budget = 227u * 1024u - 2048u;
    // Two operand stages where shared memory allows (conversion of item i+1 under the MMAs of item i;
    // only pays off with the dedicated issue warp — with the issuing thread inside a converting warp
    // everybody waited for it at the barrier anyway).  The epilogue goes through a per-warp
    // shared-memory transpose (32 rows x PARTC columns, rows padded by 16 B) and leaves as 64/128-byte
    // row segments: row-scattered stores cost 32 LSU wavefronts per instruction.
    static constexpr int stages = ((int64_t)budget - (int64_t)(2 * bytesB + 4 * bytesA + 256) - 8 * 32 * (32 * 4 + 16)) / (int64_t)ringSlot >= 3 ? 2 : 1;
    static constexpr int partC = (budget - (2 * bytesB + 2 * bytesA + 256) - 8 * 32 * (32 * 4 + 16)) / ringSlot >= 3 ? 32 : 16;
    static constexpr uint32_t stageRow = partC * 4 + 16;          // bytes per staged row
    static constexpr uint32_t bytesStage = 8 * 32 * stageRow;     // 8 warps x 32 rows
    static constexpr uint32_t fixed = 2 * bytesB + 2 * stages * bytesA + bytesStage + 256;
    static constexpr int ringSlots = ((budget - fixed) / ringSlot) > 6 ? 6 : (int)((budget - fixed) / ringSlot);
    static constexpr uint32_t offBhi = 0, offBlo = bytesB, offA = 2 * bytesB;     // stage s: hi at offA + s*2*bytesA, lo + bytesA
    static constexpr uint32_t offStage = (offA + 2 * stages * bytesA + 127) / 128 * 128;
    static constexpr uint32_t offRing = (offStage + bytesStage + 127) / 128 * 128;
    static constexpr uint32_t total = offRing + ringSlots * ringSlot + 64;
    static constexpr int tmemCols = 4 * N;                        // (main + small-terms) x ping-pong: 256 or 512
    static_assert(ringSlots >= 2, "ring too small");
};

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(NPENDING) : "memory"); }

// One persistent CTA per SM.  Work is the flattened sequence of (row tile, 32-column chunk) items.
//   * every thread streams ITS OWN four float4 of each item through a cp.async ring in shared
//     memory (thread-private slots: no barrier between landing and use), ringSlots - 1 items ahead;
//   * per item: read the landed chunk, split hi / lo, wait until the previous item's MMAs have
//     released the operand stage, store, fence, barrier, one thread issues 12 tcgen05.mma + commit;
//   * accumulators ping-pong between two TMEM buffers; the epilogue of tile t (tcgen05.ld, bias,
//     ELU, stores) runs after the first MMAs of tile t + 1 have been issued, so the tensor core and
//     the loads keep going underneath it.
template <int N, int K>
__global__ void __launch_bounds__(kThreadsNLAll, 1)
node_linear_kernel(const float *__restrict__ x, int64_t ldx, int64_t M, const float *__restrict__ w,
                   int64_t ldw, int w_is_kn, const float *__restrict__ bias, int act,
                   float *__restrict__ y, int64_t ldy, const NLPush push) {
    using S = NLSmem<N, K>;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_stage[2], bar_ready[2], bar_full[2], bar_dfree[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = umma::smem_u32(smem);
    constexpr int NCH = K / kKC;
    constexpr int RS = S::ringSlots;

    // ---- one-time setup: TMEM, barriers, W -> smem (hi / lo, chunk-interleaved K-major [n][k])
    if (warp == 0) umma::tmem_alloc(&tmem_base_s, S::tmemCols);
    if (tid == 32) {
        umma::mbar_init(&bar_stage[0], 1);
        umma::mbar_init(&bar_stage[1], 1);
        umma::mbar_init(&bar_full[0], 1);
        umma::mbar_init(&bar_full[1], 1);
        umma::mbar_init(&bar_ready[0], kThreadsNL);              // every producer thread arrives
        umma::mbar_init(&bar_ready[1], kThreadsNL);
        umma::mbar_init(&bar_dfree[0], kThreadsNL);
        umma::mbar_init(&bar_dfree[1], kThreadsNL);
        umma::fence_mbar_init();
    }
    for (int idx = tid; idx < N * K; idx += kThreadsNLAll) {
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
    const int64_t my_tiles = blockIdx.x < num_tiles ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t n_items = my_tiles * NCH;
    const int f = tid & 7, rb = tid >> 3;                        // float4 slot in the 32-k chunk, base row
    constexpr uint32_t idesc = umma::idesc_tf32(kTileM, N, false, false);

    auto issue_item = [&](int64_t it) {                          // cp.async of this thread's part of item `it`
        if (it < n_items) {
            const int64_t tile = blockIdx.x + (it / NCH) * gridDim.x;
            const int c = (int)(it % NCH);
            const uint32_t slot = sbase + S::offRing + (uint32_t)(it % RS) * S::ringSlot;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = rb + 32 * i;
                const int64_t row = tile * kTileM + r;
                if (row < M) cp_async16(slot + (uint32_t)r * 128 + (uint32_t)f * 16, x + row * ldx + c * kKC + f * 4);
            }
        }
        cp_async_commit();                                       // one group per item, empty or not
    };

    auto epilogue = [&](int64_t t_local) {                       // tile index local to this CTA
        const int64_t tile = blockIdx.x + t_local * gridDim.x;
        const int buf = (int)(t_local & 1);
        umma::mbar_wait(&bar_full[buf], (uint32_t)((t_local >> 1) & 1));
        umma::fence_after_sync();
        const int q = warp & 3, h = warp >> 2;
        constexpr int COLS = N / 2;                              // columns per warp half
        constexpr int PC = S::partC;
        uint8_t *stg = smem + S::offStage + (uint32_t)warp * 32 * S::stageRow;    // this warp's 32 x PC staging tile
        const int64_t row0 = tile * kTileM + q * 32;
#pragma unroll
        for (int part = 0; part < COLS / PC; ++part) {
            const int c0 = h * COLS + part * PC;
            float v[PC], vs[PC];
            const uint32_t ta = tmem_d + (uint32_t)(buf * 2 * N) + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            umma::tmem_ld<PC>(ta, v);
            umma::tmem_ld<PC>(ta + N, vs);
#pragma unroll
            for (int j = 0; j < PC; ++j) v[j] += vs[j];
            // thread = row: bias / activation, then its PC values into its (padded) staging row
#pragma unroll
            for (int j = 0; j < PC; j += 4) {
                float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                if (bias) {
                    const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + c0 + j));
                    o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                }
                if (act == PANGNN_ACT_ELU) { o.x = elu1f(o.x); o.y = elu1f(o.y); o.z = elu1f(o.z); o.w = elu1f(o.w); }
                *reinterpret_cast<float4 *>(stg + (uint32_t)lane * S::stageRow + j * 4) = o;
            }
            __syncwarp();
            // coalesced: LPR lanes cover one row segment of PC floats, 32 / LPR rows per store
            constexpr int LPR = PC / 4;
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += 32 / LPR) {
                const int r = r0 + lane / LPR, fl = lane % LPR;
                const float4 o = *reinterpret_cast<const float4 *>(stg + (uint32_t)r * S::stageRow + fl * 16);
                if (row0 + r < M) {
                    *reinterpret_cast<float4 *>(y + (row0 + r) * ldy + c0 + fl * 4) = o;
                    const int64_t gr = row0 + r;
                    if (push.slot0 && gr >= push.lo0 && gr < push.hi0) {
                        const int32_t ps = __ldg(push.slot0 + gr);
                        if (ps >= 0) *reinterpret_cast<float4 *>(push.peer0 + (int64_t)ps * ldy + c0 + fl * 4) = o;
                    }
                    if (push.slot1 && gr >= push.lo1 && gr < push.hi1) {
                        const int32_t ps = __ldg(push.slot1 + gr);
                        if (ps >= 0) *reinterpret_cast<float4 *>(push.peer1 + (int64_t)ps * ldy + c0 + fl * 4) = o;
                    }
                }
            }
            __syncwarp();
        }
        umma::fence_before_sync();                               // orders the tcgen05.ld before the next barrier
    };

    constexpr int ST = S::stages;
    if (warp == kThreadsNL / 32) {
        // ===== MMA issue warp: waits for an operand stage, issues its 12 tcgen05.mma, commits =====
#pragma unroll 1
        for (int64_t it = 0; it < n_items; ++it) {
            const int64_t t_local = it / NCH;
            const int c = (int)(it % NCH);
            const int st = (int)(it % ST);
            umma::mbar_wait(&bar_ready[st], (uint32_t)((it / ST) & 1));          // converted operands are in smem
            // accumulator buffer t_local & 1 was last used by tile t_local - 2: its epilogue must have drained it
            if (c == 0 && t_local >= 2) umma::mbar_wait(&bar_dfree[t_local & 1], (uint32_t)(((t_local >> 1) - 1) & 1));
            if (lane == 0) {
                umma::fence_after_sync();
                const uint32_t offAhi = S::offA + (uint32_t)st * 2 * S::bytesA, offAlo = offAhi + S::bytesA;
                const uint32_t dm = tmem_d + (uint32_t)((t_local & 1) * 2 * N);      // main | small-terms accumulator
                umma::mma_3xtf32<kKC / 8>(dm, dm + N, sbase + offAhi, sbase + offAlo,
                                 sbase + S::offBhi + (uint32_t)c * (kKC / 4) * S::chunkB,
                                 sbase + S::offBlo + (uint32_t)c * (kKC / 4) * S::chunkB,
                                 kChunkA, 128, 2 * kChunkA, S::chunkB, 128, 2 * S::chunkB, idesc, c > 0);
                umma::mma_commit(&bar_stage[st]);
                if (c == NCH - 1) umma::mma_commit(&bar_full[t_local & 1]);
            }
            __syncwarp();
        }
    } else {
        // ===== producer / epilogue warps =====
#pragma unroll 1
        for (int64_t it = 0; it < RS - 1; ++it) issue_item(it);      // prologue: RS - 1 items in flight
#pragma unroll 1
        for (int64_t it = 0; it < n_items; ++it) {
            const int64_t t_local = it / NCH;
            const int c = (int)(it % NCH);
            issue_item(it + RS - 1);                                 // slot (it - 1) % RS: consumed by this thread last time
            cp_async_wait<RS - 1>();                                 // item `it` has landed (for this thread's own copies)
            float4 raw[4];
            {
                const int64_t tile = blockIdx.x + t_local * gridDim.x;
                const uint8_t *slot = smem + S::offRing + (uint32_t)(it % RS) * S::ringSlot;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = rb + 32 * i;
                    raw[i] = (tile * kTileM + r < M) ? *reinterpret_cast<const float4 *>(slot + r * 128 + f * 16)
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            const int st = (int)(it % ST);                           // operand stage of this item
            // stage free: the MMAs of item it - ST (the n-th commit on this stage's barrier, n = (it - ST) / ST)
            if (it >= ST) umma::mbar_wait(&bar_stage[st], (uint32_t)(((it - ST) / ST) & 1));
            const uint32_t offAhi = S::offA + (uint32_t)st * 2 * S::bytesA, offAlo = offAhi + S::bytesA;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 hi, lo;
                umma::split4(raw[i], hi, lo);
                const uint32_t off = (uint32_t)f * kChunkA + (uint32_t)(rb + 32 * i) * 16;
                *reinterpret_cast<float4 *>(smem + offAhi + off) = hi;
                *reinterpret_cast<float4 *>(smem + offAlo + off) = lo;
            }
            umma::fence_async_smem();                                // my stores -> visible to the tensor core
            umma::mbar_arrive(&bar_ready[st]);
            // the previous tile's epilogue runs underneath this tile's MMAs
            if (c == 0 && t_local > 0) {
                epilogue(t_local - 1);
                umma::mbar_arrive(&bar_dfree[(t_local - 1) & 1]);    // (the epilogue ends with tcgen05.fence::before_thread_sync)
            }
        }
        if (my_tiles > 0) epilogue(my_tiles - 1);
        cp_async_wait<0>();
    }
    // ---- teardown
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_d, S::tmemCols);
}

template <int N, int K>
int launch_node_linear(const float *x, int64_t ldx, int64_t M, const float *w, int64_t ldw, int w_is_kn,
                       const float *bias, int act, float *y, int64_t ldy, const NLPush &push, cudaStream_t st) {
    using S = NLSmem<N, K>;
    static bool attr_set = false;
    if (!attr_set) {
        int rc = check_cuda(cudaFuncSetAttribute(node_linear_kernel<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)S::total), "cudaFuncSetAttribute(node_linear)");
        if (rc) return rc;
        attr_set = true;
    }
    const int64_t tiles = (M + kTileM - 1) / kTileM;
    const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
    node_linear_kernel<N, K><<<grid, kThreadsNLAll, S::total, st>>>(x, ldx, M, w, ldw, w_is_kn, bias, act, y, ldy, push);
    PANGNN_CHECK_LAUNCH("node_linear");
    return PANGNN_OK;
}

}  // namespace
}  // namespace pangnn

using namespace pangnn;

static int node_linear_dispatch(const float *x, int64_t ldx, int64_t num_rows, int32_t k, const float *w, int64_t ldw,
                                int w_is_kn, int32_t n, const float *bias, int act, float *y, int64_t ldy,
                                const NLPush &push, void *stream) {
    PANGNN_REQUIRE(num_rows >= 0, "negative row count");
    if (num_rows == 0) return PANGNN_OK;
    PANGNN_REQUIRE(x && w && y, "null pointer");
    PANGNN_REQUIRE((n == 64 || n == 128) && (k == 64 || k == 128), "n and k must be 64 or 128");
    PANGNN_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0 && ldx >= k && ldy >= n, "bad row stride");
    PANGNN_REQUIRE(ldw >= (w_is_kn ? n : k), "bad weight stride");
    PANGNN_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0 && (!bias || (uintptr_t)bias % 16 == 0),
                   "pointers must be 16-byte aligned");
    PANGNN_REQUIRE((!push.slot0 || (push.peer0 && (uintptr_t)push.peer0 % 16 == 0)) &&
                       (!push.slot1 || (push.peer1 && (uintptr_t)push.peer1 % 16 == 0)), "bad push target");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 64 && k == 64) return launch_node_linear<64, 64>(x, ldx, num_rows, w, ldw, w_is_kn, bias, act, y, ldy, push, st);
    if (n == 128 && k == 64) return launch_node_linear<128, 64>(x, ldx, num_rows, w, ldw, w_is_kn, bias, act, y, ldy, push, st);
    if (n == 64 && k == 128) return launch_node_linear<64, 128>(x, ldx, num_rows, w, ldw, w_is_kn, bias, act, y, ldy, push, st);
    return launch_node_linear<128, 128>(x, ldx, num_rows, w, ldw, w_is_kn, bias, act, y, ldy, push, st);
}

extern "C" int pangnn_node_linear(const float *x, int64_t ldx, int64_t num_rows, int32_t k, const float *w,
                                  int64_t ldw, int w_is_kn, int32_t n, const float *bias, int act, float *y,
                                  int64_t ldy, void *stream) {
    return node_linear_dispatch(x, ldx, num_rows, k, w, ldw, w_is_kn, n, bias, act, y, ldy, NLPush{nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0},
                                stream);
}

extern "C" int pangnn_node_linear_push(const float *x, int64_t ldx, int64_t num_rows, int32_t k, const float *w,
                                       int64_t ldw, int w_is_kn, int32_t n, const float *bias, int act, float *y,
                                       int64_t ldy, const int32_t *push_slot0, float *peer_y0, int64_t lo0, int64_t hi0,
                                       const int32_t *push_slot1, float *peer_y1, int64_t lo1, int64_t hi1,
                                       void *stream) {
    return node_linear_dispatch(x, ldx, num_rows, k, w, ldw, w_is_kn, n, bias, act, y, ldy,
                                NLPush{push_slot0, push_slot1, peer_y0, peer_y1, lo0, hi0, lo1, hi1}, stream);
}
