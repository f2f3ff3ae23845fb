// Inference form of the fused edge scorer on the tcgen05 tensor cores (3xTF32, fp32-grade): endpoint gather +
// 3-layer MLP (+ optional sigmoid / threshold / BCE-with-logits loss) in ONE pass over the scored edges.  Replaces
// src/gnn.py:171-177 and pangnn.py:220-221 / src/predict.py:54-55.  The training form (loss + the whole backward)
// lives in edge_scorer_train.cu.
//
// Per 128-edge tile (layer 1 is hoisted to the nodes, pq[n] = [h W1a^T | h W1b^T]):
//   gather   r1 = relu(pq[src,0:64] + pq[dst,64:128] + w1c*skip + b1)  -> smem X (TF32 hi / lo, K-major)
//   G1       a2 = r1 W2^T                               tcgen05, D1 (+ small-terms accumulator D1s) in TMEM [128 x 64]
//   epilogue r2 = relu(a2 + b2); z = r2 . w3 + b3 (partial sums of the column slices through shared memory);
//            logit / probability / prediction / loss
// Operands are K-major in the chunk-interleaved no-swizzle layout of umma.cuh.  256 threads, two CTAs per SM: the
// gather of one CTA overlaps the contraction and epilogue of the other.  Next tile's indices ride in registers and
// its endpoint rows are prefetched into L2.
//
// Roofline: HBM.  Bytes / edge: 8 idx + 512 gathered rows + 4 logit (+4 skip, +4 y) = 528; measured 0.90 ms for
// 1e7 edges = 5.9 TB/s = 0.90 of the measured HBM copy peak.
#include "edge_scorer.cuh"
#include "umma.cuh"

namespace pangnn {

namespace {

constexpr int D = kScD;
constexpr int BM = 128;
constexpr int NT = 256;                          // warp w serves TMEM lane group w % 4 (edge slots 32 (w%4) .. +31)
constexpr int CPT = D / (NT / 128);              // and the column slice w / 4 of width 32
constexpr uint32_t CH = BM * 16 + 16;            // chunk stride of a [128 rows] K-major operand (X = r1)
constexpr uint32_t CHW = D * 16 + 16;            // chunk stride of a [64 rows] operand (W2)
constexpr uint32_t kOpBytes = (D / 4) * CH;      // 33024: one [128 x 64] operand (hi or lo)
constexpr uint32_t kWBytes = (D / 4) * CHW;      // 16640: one [64 x 64] operand

// shared-memory map (bytes)
constexpr uint32_t oXh = 0, oXl = oXh + kOpBytes;
constexpr uint32_t oWh = oXl + kOpBytes, oWl = oWh + kWBytes;
constexpr uint32_t oVec = oWl + kWBytes;                     // b1, w1c, b2, w3: 4 * 64 floats
constexpr uint32_t oZp = oVec + 4 * D * 4;                   // [2][128] partial logits
constexpr uint32_t oSkip = oZp + (NT / 128) * BM * 4;        // [128]
constexpr uint32_t oSrc = oSkip + BM * 4, oDst = oSrc + BM * 4;
constexpr uint32_t oEnd = oDst + BM * 4;
static_assert(2 * (oEnd + 1024 + 2048) <= 227 * 1024, "two CTAs per SM");

__global__ void __launch_bounds__(NT, 2)
edge_score_fwd_kernel(const ScorerArgs p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ double lred[BM];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, h = warp >> 2;                   // TMEM lane group, column slice
    const int row = q * 32 + lane;                           // this thread's edge slot in the epilogue
    const uint32_t sb = umma::smem_u32(smem);
    float *sVec = reinterpret_cast<float *>(smem + oVec);
    float *sZp = reinterpret_cast<float *>(smem + oZp);
    float *sSkip = reinterpret_cast<float *>(smem + oSkip);
    int32_t *sSrc = reinterpret_cast<int32_t *>(smem + oSrc);
    int32_t *sDst = reinterpret_cast<int32_t *>(smem + oDst);
    constexpr uint32_t kTmemCols = 128;                      // D1 | D1s (64 columns each)

    // ---- one-time setup
    if (warp == 0) umma::tmem_alloc(&tmem_base_s, kTmemCols);
    if (tid == 32) {
        umma::mbar_init(&bar, 1);
        umma::fence_mbar_init();
    }
    for (int i = tid; i < D * D; i += NT) {                  // W2[j][k]: rows j over k
        const int j = i / D, k = i % D;
        const float v = p.w2[i];
        const float hi = umma::tf32_hi(v), lo = umma::tf32_lo(v, hi);
        const uint32_t off = (uint32_t)(k >> 2) * CHW + (uint32_t)j * 16 + (uint32_t)(k & 3) * 4;
        *reinterpret_cast<float *>(smem + oWh + off) = hi;
        *reinterpret_cast<float *>(smem + oWl + off) = lo;
    }
    if (tid < D) {
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
    // The tensor core adds into its fp32 accumulator with truncation (~2e-8 relative per tcgen05.mma in the chain):
    // the hi*hi products and the 2^-11-times-smaller correction products go to separate accumulators
    const uint32_t tD1 = tmem_base_s, tD1s = tmem_base_s + 64;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    constexpr uint32_t idesc = umma::idesc_tf32(BM, D, false, false);     // M = 128, N = 64, K-major x K-major

    float loss_acc = 0.f;
    uint32_t commits = 0;           // tcgen05.commit count (uniform); commit n completes barrier phase (n-1)&1
    const int64_t num_tiles = (p.E + BM - 1) / BM;
    // indices of the NEXT tile travel in registers (threads 0..127)
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
    // gather mapping: 16 lanes x float4 per endpoint row, 16 edges per pass, 2 groups of 4 passes
    const int fl = tid & 15, sub = tid >> 4;
    constexpr int EPP = NT / 16;
    load_indices(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t e0 = tile * BM;
        if (tid < BM) {
            sSrc[tid] = nsrc;
            sDst[tid] = ndst;
            sSkip[tid] = nskip;
        }
        __syncthreads();
        load_indices(tile + gridDim.x);
        // ---- gather + layer-1 epilogue
        {
            const float4 b1v = *reinterpret_cast<const float4 *>(sVec + fl * 4);
            const float4 w1cv = *reinterpret_cast<const float4 *>(sVec + D + fl * 4);
#pragma unroll
            for (int g = 0; g < BM / EPP / 4; ++g) {
                float4 pv[4], qv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = (g * 4 + u) * EPP + sub;
                    pv[u] = __ldg(reinterpret_cast<const float4 *>(p.pq + (int64_t)sSrc[e] * (2 * D)) + fl);
                    qv[u] = __ldg(reinterpret_cast<const float4 *>(p.pq + (int64_t)sDst[e] * (2 * D) + D) + fl);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = (g * 4 + u) * EPP + sub;
                    const float sk = sSkip[e];
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
                }
            }
        }
        // ---- G1: D1 = r1 W2^T
        umma::fence_async_smem();
        umma::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after_sync();
            umma::mma_3xtf32<D / 8>(tD1, tD1s, sb + oXh, sb + oXl, sb + oWh, sb + oWl,
                                    CH, 128, 2 * CH, CHW, 128, 2 * CHW, idesc, false);
            umma::mma_commit(&bar);
        }
        ++commits;
        // while G1 runs: pull the next tile's endpoint rows into L2
        if (tid < BM && tile + gridDim.x < num_tiles) {
            const char *ps = reinterpret_cast<const char *>(p.pq + (int64_t)nsrc * (2 * D));
            const char *pd = reinterpret_cast<const char *>(p.pq + (int64_t)ndst * (2 * D) + D);
            asm volatile("prefetch.global.L2 [%0];" :: "l"(ps));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(ps + 128));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(pd));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(pd + 128));
        }
        umma::mbar_wait(&bar, (commits - 1) & 1);
        umma::fence_after_sync();
        // ---- epilogue: thread = edge slot `row`, columns h*32 .. h*32+31
        float v[CPT];
        {
            float vs[CPT];
            umma::tmem_ld2<CPT>(tD1 + lane_off + (uint32_t)(h * CPT), v, tD1s + lane_off + (uint32_t)(h * CPT), vs);
#pragma unroll
            for (int c = 0; c < CPT; ++c) v[c] += vs[c];
        }
        {
            float zp = 0.f;
#pragma unroll
            for (int c = 0; c < CPT; c += 4) {
                const float4 b2v = *reinterpret_cast<const float4 *>(sVec + 2 * D + h * CPT + c);
                const float4 w3v = *reinterpret_cast<const float4 *>(sVec + 3 * D + h * CPT + c);
                zp = fmaf(fmaxf(v[c + 0] + b2v.x, 0.f), w3v.x, zp);
                zp = fmaf(fmaxf(v[c + 1] + b2v.y, 0.f), w3v.y, zp);
                zp = fmaf(fmaxf(v[c + 2] + b2v.z, 0.f), w3v.z, zp);
                zp = fmaf(fmaxf(v[c + 3] + b2v.w, 0.f), w3v.w, zp);
            }
            sZp[h * BM + row] = zp;
        }
        __syncthreads();
        const int64_t e = e0 + row;
        if (h == 0 && e < p.E) {
            float zz = sZp[row];
#pragma unroll
            for (int hh = 1; hh < NT / 128; ++hh) zz += sZp[hh * BM + row];
            zz += b3;
            if (p.logits) p.logits[e] = zz;
            if (p.prob || p.pred) {
                const float pr = 1.f / (1.f + expf(-zz));
                if (p.prob) p.prob[e] = pr;
                if (p.pred) p.pred[e] = pr >= p.threshold ? 1 : 0;
            }
            if (p.y && p.loss_partial) {
                // torch BCEWithLogits(pos_weight): (1-y) z + (1+(pw-1)y) (log1p(exp(-|z|)) + max(-z,0))
                const float yy = p.y[e];
                const float lw = fmaf(p.pos_weight - 1.f, yy, 1.f);
                loss_acc += (1.f - yy) * zz + lw * (log1pf(expf(-fabsf(zz))) + fmaxf(-zz, 0.f));
            }
        }
        umma::fence_before_sync();
        __syncthreads();      // D1, the partial logits and the index buffers are rewritten by the next tile
    }

    // ---- CTA epilogue
    if (p.loss_partial) {
        if (tid < BM) lred[tid] = (double)loss_acc;          // tid < 128 <=> h == 0, row == tid
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int i = 0; i < BM; ++i) s += lred[i];
            p.loss_partial[blockIdx.x] = s;
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_base_s, kTmemCols);
}

}  // namespace

int edge_score_tc_max_grid() { return kNumSMs * 2; }

int launch_edge_score_tc(const ScorerArgs &a, bool train, int *grid_out, cudaStream_t st) {
    if (train) return launch_edge_score_train(a, grid_out, st);          // edge_scorer_train.cu
    const size_t smem_fwd = oEnd + 128;
    static bool attr_set = false;
    if (!attr_set) {
        int rc = check_cuda(cudaFuncSetAttribute(edge_score_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem_fwd), "cudaFuncSetAttribute(edge_score fwd)");
        if (rc) return rc;
        attr_set = true;
    }
    const int64_t tiles = (a.E + BM - 1) / BM;
    const int64_t cap = (int64_t)kNumSMs * 2;
    const int grid = (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
    *grid_out = grid;
    edge_score_fwd_kernel<<<grid, NT, smem_fwd, st>>>(a);
    PANGNN_CHECK_LAUNCH("edge_score_fwd");
    return PANGNN_OK;
}

}  // namespace pangnn
