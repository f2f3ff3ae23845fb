// Fused edge scorer on the tcgen05 tensor cores (3xTF32, fp32-grade): endpoint gather + 3-layer MLP
// + BCE-with-logits loss + the whole backward in ONE pass over the scored edges.  Replaces
// src/gnn.py:171-177, pangnn.py:98,203 and their autograd backward (pangnn.py:207).
//
// Per 128-edge tile (layer 1 is hoisted to the nodes, pq[n] = [h W1a^T | h W1b^T]):
//   gather   r1 = relu(pq[src,0:64] + pq[dst,64:128] + w1c*skip + b1)  -> smem X (TF32 hi / lo)
//   G1       a2 = r1 W2^T                               tcgen05, D1 in TMEM   [128 x 64]
//            (while it runs: r1 is re-read per edge slot and stored TRANSPOSED into smem Y)
//   epi-1    r2 = relu(a2 + b2); z = r2 . w3 + b3; loss; dz; da2 = dz w3 [a2>0] -> TENSOR MEMORY (hi / lo; the A operand
//            of G2 in the accumulator's own thread <-> edge layout, tcgen05.st) and, row-major, -> smem X (for G3)
//   G2       dr1 = da2 W2                               tcgen05 (A from TMEM), D2   [128 x 64]
//   epi-2    da1 = dr1 [r1>0] -> HBM [E,64];  db1, dw1c column sums
//   G3       [dW2_hi ; dW2_lo] += [da2_hi ; da2_lo]^T (r1_hi + r1_lo)   tcgen05, D3 [128 x 64]:
//            the contraction runs over the EDGE index, so both operands are read MN-major — the row-major
//            tiles themselves, r1 [e][k] (written beside the K-major copy by the gather) and da2 [e][j]
//            (written over X by epilogue 1, once G1 is done; G2 and G3 are issued back to back), in the one layout the tensor core accepts for MN-major tf32
//            (128B swizzle with 32-byte atoms, umma.cuh) with 16-byte stores; D3 accumulates in TMEM across
//            the tiles of the CTA; its completion is only awaited when the next tile is about to overwrite
//            X / R.  (Round 1 wrote both operands TRANSPOSED with 4-byte scatters: 96 STS.32 + 8 LDS.128
//            per thread and tile, ~25 % of the kernel's instructions.)
// G1 / G2 operands are K-major in the chunk-interleaved no-swizzle layout of umma.cuh.
// W2 is kept both as [j][k] (G1) and as [k][j] (G2).
// Column sums (db2, dw3, db1, dw1c) are kept per edge slot in registers across tiles and reduced
// once per CTA in fixed order; per-CTA partials are summed by reduce_partials (no atomics).
//
// Roofline: HBM.  Bytes / edge: 8 idx + 512 gathered rows + 4 logit (+4 skip, +4 y), training
// adds 256 for da1 -> 528 / 784 B per edge; 24.6 kFLOP / edge now run on the tensor pipe.
#include "edge_scorer.cuh"
#include "umma.cuh"

namespace pangnn {

#define PROF_T(i) do { } while (0)

namespace {

constexpr int D = kScD;
constexpr int BM = 128;
constexpr uint32_t CH = BM * 16 + 16;            // chunk stride of a [128 rows] K-major operand (X: r1, then da2)
constexpr uint32_t CHW = D * 16 + 16;            // chunk stride of a [64 rows] operand (W2, W2^T)
constexpr uint32_t kOpBytes = (D / 4) * CH;      // 33024: one [128 x 64] operand (hi or lo)
constexpr uint32_t kWBytes = (D / 4) * CHW;      // 16640: one [64 x 64] operand

// shared-memory map (bytes).  X = r1 (hi | lo) K-major, later da2 (hi | lo) K-major, later da2 row-major (G3)
constexpr uint32_t oXh = 0, oXl = oXh + kOpBytes;
constexpr uint32_t oWh = oXl + kOpBytes, oWl = oWh + kWBytes;
constexpr uint32_t oVec = oWl + kWBytes;                     // b1, w1c, b2, w3: 4 * 64 floats
constexpr uint32_t oZp = oVec + 4 * D * 4;                   // [<= 4][128] partial logits
constexpr uint32_t oSkip = oZp + 4 * BM * 4;                 // [2][128]: this tile's and the next tile's staging
constexpr uint32_t oSrc = oSkip + 2 * BM * 4, oDst = oSrc + 2 * BM * 4;
constexpr uint32_t oM1 = oDst + 2 * BM * 4;                  // [128][16] relu-mask nibbles of r1 (TRAIN)
constexpr uint32_t oFwdEnd = oM1 + BM * 16;
constexpr uint32_t oWTh = (oFwdEnd + 127) / 128 * 128, oWTl = oWTh + kWBytes;   // W2^T (TRAIN)
// MN-major operands of G3 (TRAIN): R = r1 row-major [e][k], hi and lo, 2 panels of 32 k each; the row-major
// [da2_hi | da2_lo] (4 panels of 32 j) is written over X after G2.  Offsets are relative to a 1024-byte aligned base.
constexpr uint32_t kPanel = BM * 128;                                           // 16 KB: 128 edge rows x 32 floats
constexpr uint32_t oRh = (oWTl + kWBytes + 1023) / 1024 * 1024, oRl = oRh + 2 * kPanel;
constexpr uint32_t oTrainEnd = oRl + 2 * kPanel;
static_assert(4 * kPanel <= 2 * kOpBytes, "X must also hold the row-major [da2_hi | da2_lo] operand");
static_assert(oTrainEnd + 1024 + 2048 <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ float4 shfl_xor4(float4 v, int m) {
    v.x = __shfl_xor_sync(0xffffffffu, v.x, m); v.y = __shfl_xor_sync(0xffffffffu, v.y, m);
    v.z = __shfl_xor_sync(0xffffffffu, v.z, m); v.w = __shfl_xor_sync(0xffffffffu, v.w, m);
    return v;
}
// 4 x 4 transpose of float4 elements inside every quad of lanes: in, lane r of the quad holds S[c] = M[r][c];
// out, lane i holds S[j] = M[j][i].  Two exchange rounds (lane ^ 2, lane ^ 1), half of the data each.
__device__ __forceinline__ void quad_transpose(float4 (&S)[4], int lane) {
    const bool up = lane & 2, odd = lane & 1;
    float4 s0 = up ? S[0] : S[2], s1 = up ? S[1] : S[3];
    s0 = shfl_xor4(s0, 2); s1 = shfl_xor4(s1, 2);
    if (up) { S[0] = s0; S[1] = s1; } else { S[2] = s0; S[3] = s1; }
    s0 = odd ? S[0] : S[1]; s1 = odd ? S[2] : S[3];
    s0 = shfl_xor4(s0, 1); s1 = shfl_xor4(s1, 1);
    if (odd) { S[0] = s0; S[2] = s1; } else { S[1] = s0; S[3] = s1; }
}

// NT compute threads = NT/32 warps: warp w serves TMEM lane group w % 4 (edge slots 32 (w%4) .. +31) and
// the column slice w / 4 of width CPT = 64 / (NT / 128).
// TRAIN: one more warp GROUP (warps NT/32 .. NT/32+3; only the first one works) does nothing but issue the
// tcgen05.mma groups and hands registers to the compute warps (setmaxnreg: 24 vs 112 per thread — the CTA's pool is
// what it was launched with, 640 x 96 registers, so 512 x 120 + 128 x 24 does not fit: asking for 120 deadlocked
// on hardware; a single extra warp does not work either, 17 warps put 5 on one scheduler = 96 registers each).
// With the issuing thread
// inside a compute warp, that warp — and through the CTA barriers everybody — waited while the tensor pipe's
// queue drained (the 32 MMAs of G3 held thread 0 for ~2000 cycles per tile: per-phase cycle counters of the
// profiling build, tools/scorer_phases.py).  Hand-offs: every compute warp arrives on `bar_ops` when its part
// of an operand set is in shared memory, the issuer waits for all NT/32 arrivals, issues, and commits to `bar`.
template <bool TRAIN, int NT>
__global__ void __launch_bounds__(NT + (TRAIN ? 128 : 0), TRAIN ? 1 : 2)
edge_score_tc_kernel(const ScorerArgs p) {
    constexpr int kThreads = NT;
    constexpr int CPT = D / (NT / 128);                      // columns per thread in the epilogues
    static_assert(CPT == 16 || CPT == 32, "256 or 512 threads");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // the swizzled MN-major operands want a 1024-byte aligned base (dynamic shared memory starts after the statics)
    uint8_t *smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bar, bar_ops, bar3;      // G1 / G2 done; operands ready; G3 done
    __shared__ uint32_t tmem_base_s;
    __shared__ double lred[BM];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, h = warp >> 2;                   // TMEM lane group, column slice
    const int row = q * 32 + lane;                           // this thread's edge slot in the epilogues
    const uint32_t sb = umma::smem_u32(smem);
    float *sVec = reinterpret_cast<float *>(smem + oVec);
    float *sZp = reinterpret_cast<float *>(smem + oZp);
    float *sSkip = reinterpret_cast<float *>(smem + oSkip);
    int32_t *sSrc = reinterpret_cast<int32_t *>(smem + oSrc);
    int32_t *sDst = reinterpret_cast<int32_t *>(smem + oDst);
    uint8_t *sM1 = smem + oM1;
    constexpr uint32_t kTmemCols = TRAIN ? 512 : 128;            // D1 | D1s | D2 | D2s | D3 | da2_hi | da2_lo (64 columns each)

    // ---- one-time setup
    if (warp == 0) umma::tmem_alloc(&tmem_base_s, kTmemCols);
    if (tid == 32) {
        umma::mbar_init(&bar, 1);
        umma::mbar_init(&bar_ops, NT / 32);
        umma::mbar_init(&bar3, 1);
        umma::fence_mbar_init();
    }
    const bool issuer = TRAIN && warp >= NT / 32;               // the dedicated MMA-issue warp group
    // barrier of the NT compute threads (the issuer warp never joins it)
    auto sync_compute = [&]() {
        if constexpr (TRAIN) asm volatile("bar.sync 1, %0;" :: "n"(NT) : "memory");
        else __syncthreads();
    };
    // this warp's part of an operand set is in shared memory (visible to the async proxy): tell the issuer
    auto ops_ready = [&]() {
        umma::fence_async_smem();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(&bar_ops);
    };
    for (int i = tid; i < D * D && !issuer; i += kThreads) {            // W2[j][k]: rows j over k, and rows k over j
        const int j = i / D, k = i % D;
        const float v = p.w2[i];
        const float hi = umma::tf32_hi(v), lo = umma::tf32_lo(v, hi);
        const uint32_t off = (uint32_t)(k >> 2) * CHW + (uint32_t)j * 16 + (uint32_t)(k & 3) * 4;
        *reinterpret_cast<float *>(smem + oWh + off) = hi;
        *reinterpret_cast<float *>(smem + oWl + off) = lo;
        if (TRAIN) {
            const uint32_t offT = (uint32_t)(j >> 2) * CHW + (uint32_t)k * 16 + (uint32_t)(j & 3) * 4;
            *reinterpret_cast<float *>(smem + oWTh + offT) = hi;
            *reinterpret_cast<float *>(smem + oWTl + offT) = lo;
        }
    }
    if (tid < D) {      // (compute threads)
        sVec[tid] = p.b1[tid];
        sVec[D + tid] = (p.skip && p.w1c) ? p.w1c[tid] : 0.f;
        sVec[2 * D + tid] = p.b2[tid];
        sVec[3 * D + tid] = p.w3[tid];
    }
    const float b3 = p.b3[0];
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tD1 = tmem_base_s, tD1s = tmem_base_s + 64, tD2 = tmem_base_s + 128,
                   tD3 = tmem_base_s + 256, tAh = tmem_base_s + 320, tAl = tmem_base_s + 384;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    constexpr uint32_t idesc = umma::idesc_tf32(BM, D, false, false);     // M = 128, N = 64, K-major x K-major

    // per-edge-slot column sums, live across tiles (columns h*32 .. h*32+31)
    float gw3[CPT], gb2[CPT];
    // db1 / dw1c: after the quad transpose of epilogue 2 a thread sums 4 columns (h*16 + 4 (lane % 4) ..) over its
    // quad's 4 edge slots
    float4 gb1q = make_float4(0.f, 0.f, 0.f, 0.f), gw1cq = make_float4(0.f, 0.f, 0.f, 0.f);
    float gb3 = 0.f, loss_acc = 0.f;
    // The tensor core adds into its fp32 accumulator with truncation, an error that grows with the
    // length of the accumulation chain (measured ~2e-8 relative per tcgen05.mma).  D3 is therefore
    // drained into these registers (round-to-nearest adds) every kG3Flush tiles.
    constexpr int kG3Flush = 4;
    float g3acc[CPT];
    int g3_tiles = 0;               // tiles accumulated in D3 since the last drain
    if (TRAIN) {
#pragma unroll
        for (int c = 0; c < CPT; ++c) gw3[c] = gb2[c] = g3acc[c] = 0.f;
    }

    uint32_t commits = 0;           // tcgen05.commit count on `bar` (uniform); commit n completes barrier phase (n-1)&1
    uint32_t g3_commits = 0;        // ... on `bar3` (one per tile)
    bool g3_pending = false;        // the last G3 has not been waited for yet
    const int64_t num_tiles = (p.E + BM - 1) / BM;
    // indices of the NEXT tile travel in registers (threads 0..127), its endpoint rows are pulled
    // into L2 while the current tile computes
    int32_t nsrc = 0, ndst = 0;
    float nskip = 0.f;
    auto load_indices = [&](int64_t t) {
        nsrc = ndst = 0;
        nskip = 0.f;
        if (tid < BM && t < num_tiles) {
            const int64_t e = t * BM + tid;
            if (e < p.E) {
                nsrc = p.src[e];
                ndst = p.dst[e];
                if (p.skip) nskip = p.skip[e];
            }
        }
    };
    if (issuer) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        // ---- the MMA-issue warp: three operand sets per tile (G1, G2, G3), one elected lane issues
        if (warp == NT / 32 && lane == 0) {
            // small code on purpose (this warp group runs on 24 registers per thread): rolled loops,
            // descriptors advanced incrementally
            uint32_t ready = 0;         // completed phases of bar_ops
            int64_t it = 0;             // local tile counter: D3 restarts (accumulate = 0) after every drain
            const uint64_t dXh = umma::smem_desc(sb + oXh, CH, 128), dXl = umma::smem_desc(sb + oXl, CH, 128);
            const uint64_t dWh = umma::smem_desc(sb + oWh, CHW, 128), dWl = umma::smem_desc(sb + oWl, CHW, 128);
            const uint64_t dWTh = umma::smem_desc(sb + oWTh, CHW, 128), dWTl = umma::smem_desc(sb + oWTl, CHW, 128);
            const uint64_t dRh = umma::smem_desc_mn(sb + oRh, kPanel), dRl = umma::smem_desc_mn(sb + oRl, kPanel);
            const uint64_t dAmn = umma::smem_desc_mn(sb + oXh, kPanel);                     // 4 panels: hi j 0-63, lo j 0-63
            constexpr uint32_t idesc_mn = umma::idesc_tf32(BM, D, true, true);
            constexpr uint64_t stepX = (2 * CH) >> 4, stepW = (2 * CHW) >> 4;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                // G1: D1 = r1 W2^T (A = X K-major in shared memory)
                umma::mbar_wait(&bar_ops, ready++ & 1);
                umma::fence_after_sync();
                {
                    uint64_t ah = dXh, al = dXl, bh = dWh, bl = dWl;
#pragma unroll 1
                    for (int s = 0; s < D / 8; ++s, ah += stepX, al += stepX, bh += stepW, bl += stepW) {
                        umma::mma_tf32(tD1, al, bh, idesc, s > 0 ? 1u : 0u);      // small terms first
                        umma::mma_tf32(tD1, ah, bl, idesc, 1u);
                        umma::mma_tf32(tD1, ah, bh, idesc, 1u);
                    }
                }
                umma::mma_commit(&bar);
                // G2: D2 = da2 W2 (A = da2 hi / lo in tensor memory, 8 columns per k-step), then G3 on the same hand-off
                umma::mbar_wait(&bar_ops, ready++ & 1);
                umma::fence_after_sync();
                {
                    uint64_t bh = dWTh, bl = dWTl;
                    uint32_t ah = tAh, al = tAl;
#pragma unroll 1
                    for (int s = 0; s < D / 8; ++s, ah += 8, al += 8, bh += stepW, bl += stepW) {
                        umma::mma_tf32_ts(tD2, al, bh, idesc, s > 0 ? 1u : 0u);
                        umma::mma_tf32_ts(tD2, ah, bl, idesc, 1u);
                        umma::mma_tf32_ts(tD2, ah, bh, idesc, 1u);
                    }
                }
                umma::mma_commit(&bar);
                // G3: D3 += [da2_hi | da2_lo]^T (r1_hi + r1_lo), both operands MN-major
                uint32_t acc = (it % kG3Flush) == 0 ? 0u : 1u;
                uint64_t a = dAmn, bh = dRh, bl = dRl;
#pragma unroll 1
                for (int s = 0; s < BM / 8; ++s, a += 64, bh += 64, bl += 64) {    // 8 edge rows = 1024 B per k-step
                    umma::mma_tf32(tD3, a, bl, idesc_mn, acc);
                    umma::mma_tf32(tD3, a, bh, idesc_mn, 1u);
                    acc = 1u;
                }
                umma::mma_commit(&bar3);
            }
        }
        __syncwarp();
    } else {
    if constexpr (TRAIN) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // gather mapping: 16 lanes x float4 per endpoint row, NT/16 edges per pass, groups of 4 passes
    const int fl = tid & 15, sub = tid >> 4;                 // float4 slot of the 64-wide row, edge within a pass
    constexpr int EPP = NT / 16;                             // edges per pass
    constexpr int NG = BM / EPP / 4;                         // groups of 4 passes per tile (1 at 512 threads)
    // Register prefetch of the next tile's endpoint rows (requested after epilogue 1, carried across epilogue 2):
    // measured SLOWER on B200 (3.81 vs 3.53 ms) — the L2 prefetch below already hides the gather latency and the 32
    // extra live registers spill; what the "gather" phase costs is its 2048 shared-memory store wavefronts.
    constexpr bool PRE = TRAIN && false;
    static_assert(!PRE || NG == 1, "the prefetching form keeps one whole tile of endpoint rows in registers");
    float4 pv[4], qv[4];                                     // endpoint row pieces of the tile about to be processed
    auto issue_gather = [&](int buf, int g) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = (g * 4 + u) * EPP + sub;
            pv[u] = __ldg(reinterpret_cast<const float4 *>(p.pq + (int64_t)sSrc[buf * BM + e] * (2 * D)) + fl);
            qv[u] = __ldg(reinterpret_cast<const float4 *>(p.pq + (int64_t)sDst[buf * BM + e] * (2 * D) + D) + fl);
        }
    };
    load_indices(blockIdx.x);
    if constexpr (PRE) {
        // PRE: the endpoint rows of tile t+1 are requested in the middle of tile t (after epilogue 1) and travel in
        // registers across epilogue 2 — the gather was the longest exposed latency of the tile (3600 of 14600 cycles).
        // Index staging is double-buffered: buffer b = this tile, b ^ 1 = the next one.
        if (tid < BM) {
            sSrc[tid] = nsrc;
            sDst[tid] = ndst;
            sSkip[tid] = nskip;
        }
        sync_compute();
        if (blockIdx.x < num_tiles) issue_gather(0, 0);
        load_indices((int64_t)blockIdx.x + gridDim.x);
    }
    int buf = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, buf ^= PRE ? 1 : 0) {
        const int64_t e0 = tile * BM;
        if (tid < BM) {                                      // PRE: the NEXT tile's indices; else this tile's
            const int o = PRE ? (buf ^ 1) * BM : 0;
            sSrc[o + tid] = nsrc;
            sDst[o + tid] = ndst;
            sSkip[o + tid] = nskip;
        }
        sync_compute();
        PROF_T(0);
        load_indices(tile + (PRE ? 2 : 1) * (int64_t)gridDim.x);
        // ---- gather + layer-1 epilogue
        {
            const float4 b1v = *reinterpret_cast<const float4 *>(sVec + fl * 4);
            const float4 w1cv = *reinterpret_cast<const float4 *>(sVec + D + fl * 4);
#pragma unroll
            for (int g = 0; g < NG; ++g) {
            if constexpr (!PRE) issue_gather(0, g);
            if (TRAIN && g3_pending) {                       // X / R are still being read by the previous tile's G3
                umma::mbar_wait(&bar3, (g3_commits - 1) & 1);
                g3_pending = false;
                if (g3_tiles == kG3Flush) {                  // drain D3 (uniform branch)
                    umma::fence_after_sync();
                    float t[CPT];
                    umma::tmem_ld<CPT>(tD3 + lane_off + (uint32_t)(h * CPT), t);
#pragma unroll
                    for (int c = 0; c < CPT; ++c) g3acc[c] += t[c];
                    g3_tiles = 0;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = (g * 4 + u) * EPP + sub;
                const float sk = sSkip[buf * BM + e];
                float4 a;
                a.x = fmaxf(pv[u].x + qv[u].x + fmaf(w1cv.x, sk, b1v.x), 0.f);
                a.y = fmaxf(pv[u].y + qv[u].y + fmaf(w1cv.y, sk, b1v.y), 0.f);
                a.z = fmaxf(pv[u].z + qv[u].z + fmaf(w1cv.z, sk, b1v.z), 0.f);
                a.w = fmaxf(pv[u].w + qv[u].w + fmaf(w1cv.w, sk, b1v.w), 0.f);
                float4 ahi, alo;
                umma::split4(a, ahi, alo);
                const uint32_t offK = (uint32_t)fl * CH + (uint32_t)e * 16;
                *reinterpret_cast<float4 *>(smem + oXh + offK) = ahi;
                *reinterpret_cast<float4 *>(smem + oXl + offK) = alo;
                if (TRAIN) {                                 // the same row once more, row-major, for G3
                    const uint32_t offM = umma::mn_off((uint32_t)e, (uint32_t)fl * 4, kPanel);
                    *reinterpret_cast<float4 *>(smem + oRh + offM) = ahi;
                    *reinterpret_cast<float4 *>(smem + oRl + offM) = alo;
                    // relu mask of these 4 columns for epilogue 2 (r1 >= 0: positive <=> non-zero)
                    sM1[e * 16 + fl] = (uint8_t)((a.x > 0.f ? 1u : 0u) | (a.y > 0.f ? 2u : 0u) |
                                                 (a.z > 0.f ? 4u : 0u) | (a.w > 0.f ? 8u : 0u));
                }
            }
            }
        }
        PROF_T(1);
        // ---- G1: D1 = r1 W2^T
        if constexpr (TRAIN) {
            ops_ready();
            sync_compute();             // X is complete for the r1 re-read below
        } else {
            umma::fence_async_smem();
            umma::fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                umma::fence_after_sync();
                umma::mma_3xtf32<D / 8>(tD1, tD1s, sb + oXh, sb + oXl, sb + oWh, sb + oWl,
                                        CH, 128, 2 * CH, CHW, 128, 2 * CHW, idesc, false);
                umma::mma_commit(&bar);
            }
        }
        PROF_T(2);
        ++commits;
        // ---- while G1 runs: relu mask of this edge slot's r1 columns h*16 .. +15 (4 nibbles written by the gather)
        uint32_t m1 = 0;                                     // bit c: r1[row][h*CPT + c] > 0
        if (TRAIN) {
            const uint32_t nb = *reinterpret_cast<const uint32_t *>(sM1 + row * 16 + h * 4);
            m1 = (nb & 0xfu) | ((nb >> 4) & 0xf0u) | ((nb >> 8) & 0xf00u) | ((nb >> 12) & 0xf000u);
        }
        if (!PRE && tid < BM && tile + gridDim.x < num_tiles) {
            const char *ps = reinterpret_cast<const char *>(p.pq + (int64_t)nsrc * (2 * D));
            const char *pd = reinterpret_cast<const char *>(p.pq + (int64_t)ndst * (2 * D) + D);
            asm volatile("prefetch.global.L2 [%0];" :: "l"(ps));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(ps + 128));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(pd));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(pd + 128));
        }
        PROF_T(3);
        umma::mbar_wait(&bar, (commits - 1) & 1);
        umma::fence_after_sync();
        PROF_T(4);
        // ---- epilogue 1: thread = edge slot `row`, columns h*32 .. h*32+31
        float v[CPT];
        {
            if constexpr (TRAIN) {
                // training form: ONE accumulator per contraction (24 tcgen05.mma in the chain, truncation ~5e-7 relative,
                // measured) — reading a second 32 KB accumulator costs 512 cycles of tensor-memory bandwidth per tile
                umma::tmem_ld<CPT>(tD1 + lane_off + (uint32_t)(h * CPT), v);
            } else {
                float vs[CPT];
                umma::tmem_ld2<CPT>(tD1 + lane_off + (uint32_t)(h * CPT), v, tD1s + lane_off + (uint32_t)(h * CPT), vs);
#pragma unroll
                for (int c = 0; c < CPT; ++c) v[c] += vs[c];
            }
        }
        uint32_t m2 = 0;                                     // bit c: a2[row][h*32 + c] > 0
        {
            float zp = 0.f;
#pragma unroll
            for (int c = 0; c < CPT; c += 4) {
                const float4 b2v = *reinterpret_cast<const float4 *>(sVec + 2 * D + h * CPT + c);
                const float4 w3v = *reinterpret_cast<const float4 *>(sVec + 3 * D + h * CPT + c);
                v[c + 0] = fmaxf(v[c + 0] + b2v.x, 0.f);                          // r2
                v[c + 1] = fmaxf(v[c + 1] + b2v.y, 0.f);
                v[c + 2] = fmaxf(v[c + 2] + b2v.z, 0.f);
                v[c + 3] = fmaxf(v[c + 3] + b2v.w, 0.f);
                m2 |= (v[c + 0] > 0.f ? 1u : 0u) << (c + 0);
                m2 |= (v[c + 1] > 0.f ? 1u : 0u) << (c + 1);
                m2 |= (v[c + 2] > 0.f ? 1u : 0u) << (c + 2);
                m2 |= (v[c + 3] > 0.f ? 1u : 0u) << (c + 3);
                zp = fmaf(v[c + 0], w3v.x, zp); zp = fmaf(v[c + 1], w3v.y, zp);
                zp = fmaf(v[c + 2], w3v.z, zp); zp = fmaf(v[c + 3], w3v.w, zp);
            }
            sZp[h * BM + row] = zp;
        }
        sync_compute();
        const int64_t e = e0 + row;
        const bool ok = e < p.E;
        float zz = sZp[row];
#pragma unroll
        for (int hh = 1; hh < NT / 128; ++hh) zz += sZp[hh * BM + row];
        zz += b3;
        const float yy = (ok && p.y) ? p.y[e] : 0.f;
        if (h == 0 && ok) {
            if (p.logits) p.logits[e] = zz;
            if (p.prob || p.pred) {
                const float pr = 1.f / (1.f + expf(-zz));
                if (p.prob) p.prob[e] = pr;
                if (p.pred) p.pred[e] = pr >= p.threshold ? 1 : 0;
            }
            if (p.y && p.loss_partial) {
                // torch BCEWithLogits(pos_weight): (1-y) z + (1+(pw-1)y) (log1p(exp(-|z|)) + max(-z,0))
                const float lw = fmaf(p.pos_weight - 1.f, yy, 1.f);
                loss_acc += (1.f - yy) * zz + lw * (log1pf(expf(-fabsf(zz))) + fmaxf(-zz, 0.f));
            }
        }
        if (TRAIN) {
            // read before this warp's next arrival on bar_ops: threads 0..127 overwrite the index / skip staging
            // at the top of the next tile, ordered after every warp's G2 arrival through the issuer's commit
            const float4 sk4 = *reinterpret_cast<const float4 *>(sSkip + buf * BM + (row & ~3));      // this quad's 4 edge slots
            float dz = 0.f;
            if (ok) {
                if (p.dlogits) {
                    dz = p.dlogits[e] * p.scale;
                } else {
                    // torch's backward: ((pw*y + 1 - y) * sigmoid(z) - pw*y) * grad
                    const float sg = 1.f / (1.f + expf(-zz));
                    const float t = p.pos_weight * yy;
                    dz = ((t + 1.f - yy) * sg - t) * p.scale;
                }
            }
            if (h == 0) gb3 += dz;
            // da2 = dz * w3 * [a2 > 0]: hi / lo -> tensor memory (A operand of G2: this thread's lane, columns j) and,
            // row-major, -> X (MN-major A operand of G3: panels hi j 0-31, 32-63, lo j 0-31, 32-63).  X's K-major r1 is
            // dead: G1 has completed and every warp's mask read precedes the barrier above.
#pragma unroll
            for (int c0 = 0; c0 < CPT; c0 += 8) {
                float dh8[8], dl8[8];
#pragma unroll
                for (int c = c0; c < c0 + 8; c += 4) {
                    float4 d;
                    const int j = h * CPT + c;
                    const float4 w3v = *reinterpret_cast<const float4 *>(sVec + 3 * D + j);
                    d.x = (m2 >> (c + 0)) & 1u ? dz * w3v.x : 0.f;
                    d.y = (m2 >> (c + 1)) & 1u ? dz * w3v.y : 0.f;
                    d.z = (m2 >> (c + 2)) & 1u ? dz * w3v.z : 0.f;
                    d.w = (m2 >> (c + 3)) & 1u ? dz * w3v.w : 0.f;
                    gw3[c + 0] = fmaf(dz, v[c + 0], gw3[c + 0]); gw3[c + 1] = fmaf(dz, v[c + 1], gw3[c + 1]);
                    gw3[c + 2] = fmaf(dz, v[c + 2], gw3[c + 2]); gw3[c + 3] = fmaf(dz, v[c + 3], gw3[c + 3]);
                    gb2[c + 0] += d.x; gb2[c + 1] += d.y; gb2[c + 2] += d.z; gb2[c + 3] += d.w;
                    float4 dh, dl;
                    umma::split4(d, dh, dl);
                    const uint32_t off = umma::mn_off((uint32_t)row, (uint32_t)j, kPanel);
                    *reinterpret_cast<float4 *>(smem + oXh + off) = dh;
                    *reinterpret_cast<float4 *>(smem + oXh + 2 * kPanel + off) = dl;
                    dh8[c - c0 + 0] = dh.x; dh8[c - c0 + 1] = dh.y; dh8[c - c0 + 2] = dh.z; dh8[c - c0 + 3] = dh.w;
                    dl8[c - c0 + 0] = dl.x; dl8[c - c0 + 1] = dl.y; dl8[c - c0 + 2] = dl.z; dl8[c - c0 + 3] = dl.w;
                }
                umma::tmem_st8(tAh + lane_off + (uint32_t)(h * CPT + c0), dh8);
                umma::tmem_st8(tAl + lane_off + (uint32_t)(h * CPT + c0), dl8);
            }
            umma::tmem_st_wait();
            umma::fence_before_sync();
            PROF_T(5);
            // ---- G2: D2 = da2 W2 and G3: D3 += da2^T r1, issued back to back by the issuer warp
            ops_ready();
            PROF_T(6);
            if (PRE && tile + gridDim.x < num_tiles) issue_gather(buf ^ 1, 0);     // lands while G2 / epilogue 2 run
            ++commits;
            ++g3_commits;
            g3_pending = true;
            umma::mbar_wait(&bar, (commits - 1) & 1);
            umma::fence_after_sync();
            PROF_T(7);
            // ---- epilogue 2: da1 = dr1 * [r1 > 0] -> HBM; db1, dw1c
            {
                umma::tmem_ld<CPT>(tD2 + lane_off + (uint32_t)(h * CPT), v);
            }
            // The 4 lanes of a quad exchange their float4 pieces so that every store instruction writes 64 contiguous
            // bytes per edge row (8 rows per warp instruction instead of 32 rows x 16 bytes: the row-scattered stores
            // were 2048 of the tile's ~6000 L1 wavefronts, the busiest pipe of this kernel).
            static_assert(!TRAIN || CPT == 16, "the quad transpose handles 4 float4 per thread");
            {
                float4 T[4];
#pragma unroll
                for (int c = 0; c < 16; c += 4) {
                    T[c / 4].x = (m1 >> (c + 0)) & 1u ? v[c + 0] : 0.f;
                    T[c / 4].y = (m1 >> (c + 1)) & 1u ? v[c + 1] : 0.f;
                    T[c / 4].z = (m1 >> (c + 2)) & 1u ? v[c + 2] : 0.f;
                    T[c / 4].w = (m1 >> (c + 3)) & 1u ? v[c + 3] : 0.f;
                }
                quad_transpose(T, lane);        // T[j] = da1[edge slot (row & ~3) + j][h*16 + 4 (lane & 3) .. + 3]
                const int64_t eq = e0 + (row & ~3);
                float *dst = p.da1 + eq * D + h * 16 + 4 * (lane & 3);
                const float skj[4] = {sk4.x, sk4.y, sk4.z, sk4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    gb1q.x += T[j].x; gb1q.y += T[j].y; gb1q.z += T[j].z; gb1q.w += T[j].w;
                    gw1cq.x = fmaf(T[j].x, skj[j], gw1cq.x); gw1cq.y = fmaf(T[j].y, skj[j], gw1cq.y);
                    gw1cq.z = fmaf(T[j].z, skj[j], gw1cq.z); gw1cq.w = fmaf(T[j].w, skj[j], gw1cq.w);
                    if (eq + j < p.E) *reinterpret_cast<float4 *>(dst + j * D) = T[j];
                }
            }
            PROF_T(9);
            ++g3_tiles;
        } else {
            umma::fence_before_sync();
            __syncthreads();      // D1 and the index buffers are rewritten by the next tile
        }
    }
    if (TRAIN && g3_pending) {
        umma::mbar_wait(&bar3, (g3_commits - 1) & 1);
        umma::fence_after_sync();
    }

    // ---- CTA epilogue
    if (p.loss_partial) {
        if (tid < BM) lred[tid] = (double)loss_acc;          // tid < 128 <=> h == 0, row == tid
        sync_compute();
        if (tid == 0) {
            double s = 0.0;
            for (int i = 0; i < BM; ++i) s += lred[i];
            p.loss_partial[blockIdx.x] = s;
        }
    }
    if (TRAIN) {
        float *out = p.partial + (int64_t)blockIdx.x * kScNGP;
        float *red = reinterpret_cast<float *>(smem);        // [128][64] floats = 32 KB (reuses X)
        // dW2[j][k] = D3[j][k] + D3[64 + j][k]
        float v[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) v[c] = 0.f;
        if (g3_tiles > 0) umma::tmem_ld<CPT>(tD3 + lane_off + (uint32_t)(h * CPT), v);
#pragma unroll
        for (int c = 0; c < CPT; ++c) v[c] += g3acc[c];
        sync_compute();
#pragma unroll
        for (int c = 0; c < CPT; ++c) red[row * D + h * CPT + c] = v[c];
        sync_compute();
        for (int i = tid; i < D * D; i += kThreads) out[kG_W2 + i] = red[i] + red[D * D + i];
        // column sums over the 128 edge slots, fixed order
        auto reduce_cols = [&](const float (&acc)[CPT], int off) {
            sync_compute();
#pragma unroll
            for (int c = 0; c < CPT; ++c) red[row * D + h * CPT + c] = acc[c];
            sync_compute();
            if (tid < D) {
                float s = 0.f;
                for (int r = 0; r < BM; ++r) s += red[r * D + tid];
                out[off + tid] = s;
            }
        };
        reduce_cols(gb2, kG_B2);
        reduce_cols(gw3, kG_W3);
        // db1 / dw1c: 32 partial rows (4 lane groups x 8 quads), 4 columns per thread
        auto reduce_cols_q = [&](const float4 acc, int off) {
            sync_compute();
            *reinterpret_cast<float4 *>(red + (q * 8 + (lane >> 2)) * D + h * 16 + 4 * (lane & 3)) = acc;
            sync_compute();
            if (tid < D) {
                float s = 0.f;
                for (int r = 0; r < 32; ++r) s += red[r * D + tid];
                out[off + tid] = s;
            }
        };
        reduce_cols_q(gb1q, kG_B1);
        reduce_cols_q(gw1cq, kG_W1C);
        sync_compute();
        if (h == 0) red[row] = gb3;
        sync_compute();
        if (tid == 0) {
            float s = 0.f;
            for (int r = 0; r < BM; ++r) s += red[r];
            out[kG_B3] = s;
        }
    }
    }   // compute warps
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_base_s, kTmemCols);
}

}  // namespace

int edge_score_tc_max_grid() { return kNumSMs * 2; }

int launch_edge_score_tc(const ScorerArgs &a, bool train, int *grid_out, cudaStream_t st) {
    if (train) return launch_edge_score_train(a, grid_out, st);          // edge_scorer_train.cu
    const size_t smem_fwd = oFwdEnd + 1024 + 128;            // + alignment slack
    static bool attr_set = false;
    if (!attr_set) {
        int rc = check_cuda(cudaFuncSetAttribute(edge_score_tc_kernel<false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem_fwd), "cudaFuncSetAttribute(edge_score fwd)");
        if (rc) return rc;
        attr_set = true;
    }
    const int64_t tiles = (a.E + BM - 1) / BM;
    const int64_t cap = (int64_t)kNumSMs * (train ? 1 : 2);
    const int grid = (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
    *grid_out = grid;
    edge_score_tc_kernel<false, 256><<<grid, 256, smem_fwd, st>>>(a);
    PANGNN_CHECK_LAUNCH("edge_score_tc");
    return PANGNN_OK;
}

}  // namespace pangnn

