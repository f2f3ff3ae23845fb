// Fused edge scorer: endpoint gather + 3-layer MLP + (optional) BCE-with-logits loss and the full
// backward, one pass over the scored edges.  Replaces src/gnn.py:171-177 (two index_selects, a
// [E, 2D(+1)] concat, three addmm + two relu), pangnn.py:98,203 (BCEWithLogitsLoss(pos_weight)) and
// their autograd backward (index_put / index_add, pangnn.py:207).
//
// Layer 1 is hoisted to the nodes by the caller:  pq[n] = [h[n] W1a^T | h[n] W1b^T]  (one node GEMM),
// so per edge:   a1 = pq[src, 0:64] + pq[dst, 64:128] (+ w1c * skip) + b1          (2 x 256 B gather)
//                z  = w3 . relu(W2 relu(a1) + b2) + b3                             (64x64 contraction)
// Backward (recomputes the forward for the tile, nothing is stored between passes):
//                da2 = dz * w3 * [a2 > 0];  dr1 = W2^T da2;  da1 = dr1 * [a1 > 0]  -> HBM, [E,64]
//                dW2 += da2 r1^T (register accumulators across the block's tiles), db2, dw3, db3,
//                db1, dw1c likewise; per-block partials are summed by a fixed-order second stage.
// The per-node gradients are sorted-segment reductions of da1 (pangnn_gcn_aggregate with val = NULL
// over the by-source / by-destination CSR of the scored edges) — no atomics anywhere.
//
// Roofline: mixed.  Bytes/edge fwd = 8 idx + 512 rows + 4 logit (+4 skip +4 y) = 528;  the 64x64
// layer is 8.2 kFLOP/edge fwd and 24.6 kFLOP/edge fwd+bwd on the fp32 FMA pipe, which bounds this
// kernel (ridge ~11 FLOP/B): 148 SMs x 128 FMA/clk.  fp32 throughout (1e-5 parity bar), so the
// tensor-core path would need 3xTF32; see DESIGN.md.
#include "common.cuh"

namespace pangnn {

int reduce_partials(const float *partial, int64_t nblocks, int32_t width, int32_t stride, float *out,
                    cudaStream_t st);

constexpr int D = PANGNN_SCORER_D;      // 64
constexpr int BM = 128;                 // edges per tile
constexpr int LDA = D + 4;              // padded smem row stride (floats): rows 16 B aligned, +4 banks
constexpr int kThreads = 256;
constexpr int NG = PANGNN_SCORER_NGRADS;
constexpr int NGP = (NG + 3) / 4 * 4;   // per-block partial stride: keeps the float4 stores 16 B aligned
// layout of the gradient vector
constexpr int G_W2 = 0, G_B2 = D * D, G_W3 = G_B2 + D, G_B3 = G_W3 + D, G_B1 = G_B3 + 1,
              G_W1C = G_B1 + D;

struct ScorerArgs {
    const float *pq;
    const int32_t *src, *dst;
    const float *skip, *w1c, *b1, *w2, *b2, *w3, *b3;
    int64_t E;
    const float *y, *dlogits;
    float pos_weight, scale;
    float *logits;      // fwd
    float *da1;         // bwd
    float *partial;     // [grid][NG] (bwd) ; loss partials [grid] doubles (fwd) via loss_partial
    double *loss_partial;
};

// acc[i][c] += sum_k A[(te + 16 i)][k] * B[k][c0 + c],  A: [BM][LDA], B: [D][D]
__device__ __forceinline__ void tile_gemm(const float *__restrict__ A, const float *__restrict__ B,
                                          int te, int c0, float (&acc)[8][4]) {
#pragma unroll 4
    for (int k = 0; k < D; k += 4) {
        float4 b[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) b[kk] = *reinterpret_cast<const float4 *>(B + (k + kk) * D + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 a = *reinterpret_cast<const float4 *>(A + (te + 16 * i) * LDA + k);
            acc[i][0] = fmaf(a.x, b[0].x, acc[i][0]); acc[i][1] = fmaf(a.x, b[0].y, acc[i][1]);
            acc[i][2] = fmaf(a.x, b[0].z, acc[i][2]); acc[i][3] = fmaf(a.x, b[0].w, acc[i][3]);
            acc[i][0] = fmaf(a.y, b[1].x, acc[i][0]); acc[i][1] = fmaf(a.y, b[1].y, acc[i][1]);
            acc[i][2] = fmaf(a.y, b[1].z, acc[i][2]); acc[i][3] = fmaf(a.y, b[1].w, acc[i][3]);
            acc[i][0] = fmaf(a.z, b[2].x, acc[i][0]); acc[i][1] = fmaf(a.z, b[2].y, acc[i][1]);
            acc[i][2] = fmaf(a.z, b[2].z, acc[i][2]); acc[i][3] = fmaf(a.z, b[2].w, acc[i][3]);
            acc[i][0] = fmaf(a.w, b[3].x, acc[i][0]); acc[i][1] = fmaf(a.w, b[3].y, acc[i][1]);
            acc[i][2] = fmaf(a.w, b[3].z, acc[i][2]); acc[i][3] = fmaf(a.w, b[3].w, acc[i][3]);
        }
    }
}

template <bool TRAIN>
__global__ void __launch_bounds__(kThreads, 2)
edge_score_kernel(const ScorerArgs p) {
    extern __shared__ __align__(16) float smem[];
    float *sR1 = smem;                                   // [BM][LDA]  relu(a1)
    float *sW2T = sR1 + BM * LDA;                        // [D][D]     W2T[k][j] = W2[j][k]
    float *sVec = sW2T + D * D;                          // b1, w1c, b2, w3 : 4*D
    float *sZ = sVec + 4 * D;                            // [BM] logits of the tile
    float *sSkip = sZ + BM;                              // [BM]
    int32_t *sSrc = reinterpret_cast<int32_t *>(sSkip + BM);   // [BM]
    int32_t *sDst = sSrc + BM;                           // [BM]
    float *sDZ = reinterpret_cast<float *>(sDst + BM);   // [BM]  (TRAIN)
    float *sW2 = sDZ + BM;                               // [D][D] W2[j][k]       (TRAIN)
    float *sDA2 = sW2 + D * D;                           // [BM][LDA]             (TRAIN)

    const int tid = threadIdx.x;
    const int tj = tid & 15, te = tid >> 4;              // GEMM-1/2 mapping: 4 cols x 8 edges
    const int c0 = tj * 4;

    // ---- one-time: weights to shared memory
    for (int i = tid; i < D * D; i += kThreads) {
        const float v = p.w2[i];                         // W2[j][k], i = j*D + k
        sW2T[(i % D) * D + (i / D)] = v;
        if (TRAIN) sW2[i] = v;
    }
    if (tid < D) {
        sVec[tid] = p.b1[tid];
        sVec[D + tid] = (p.skip && p.w1c) ? p.w1c[tid] : 0.f;
        sVec[2 * D + tid] = p.b2[tid];
        sVec[3 * D + tid] = p.w3[tid];
    }
    const float b3 = p.b3[0];
    __syncthreads();
    const float4 b2v = *reinterpret_cast<const float4 *>(sVec + 2 * D + c0);
    const float4 w3v = *reinterpret_cast<const float4 *>(sVec + 3 * D + c0);

    // per-thread gradient accumulators (live across tiles)
    float gW2[4][4];                                     // dW2[j0+a][k0+b], j0 = te*4, k0 = tj*4
    float gb2[4], gw3[4], gb1[4], gw1c[4], gb3 = 0.f;
    float loss_acc = 0.f;
    if (TRAIN) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            gb2[a] = gw3[a] = gb1[a] = gw1c[a] = 0.f;
#pragma unroll
            for (int b = 0; b < 4; ++b) gW2[a][b] = 0.f;
        }
    }

    const int64_t num_tiles = (p.E + BM - 1) / BM;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t e0 = tile * BM;
        // ---- stage indices (coalesced)
        if (tid < BM) {
            const int64_t e = e0 + tid;
            const bool ok = e < p.E;
            sSrc[tid] = ok ? p.src[e] : 0;
            sDst[tid] = ok ? p.dst[e] : 0;
            sSkip[tid] = (ok && p.skip) ? p.skip[e] : 0.f;
        }
        __syncthreads();
        // ---- gather: 16 lanes x float4 per endpoint row, 16 edges per pass, 8 passes
        {
            const int fl = tid & 15, sub = tid >> 4;
            const float4 b1v = *reinterpret_cast<const float4 *>(sVec + fl * 4);
            const float4 w1cv = *reinterpret_cast<const float4 *>(sVec + D + fl * 4);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float4 pv[4], qv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = (half * 4 + u) * 16 + sub;
                    pv[u] = __ldg(reinterpret_cast<const float4 *>(p.pq + (int64_t)sSrc[e] * (2 * D)) + fl);
                    qv[u] = __ldg(reinterpret_cast<const float4 *>(p.pq + (int64_t)sDst[e] * (2 * D) + D) + fl);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = (half * 4 + u) * 16 + sub;
                    const float sk = sSkip[e];
                    float4 a;
                    a.x = fmaxf(pv[u].x + qv[u].x + fmaf(w1cv.x, sk, b1v.x), 0.f);
                    a.y = fmaxf(pv[u].y + qv[u].y + fmaf(w1cv.y, sk, b1v.y), 0.f);
                    a.z = fmaxf(pv[u].z + qv[u].z + fmaf(w1cv.z, sk, b1v.z), 0.f);
                    a.w = fmaxf(pv[u].w + qv[u].w + fmaf(w1cv.w, sk, b1v.w), 0.f);
                    *reinterpret_cast<float4 *>(sR1 + e * LDA + fl * 4) = a;
                }
            }
        }
        __syncthreads();
        // ---- GEMM-1: a2[e][j] = sum_k r1[e][k] W2[j][k]
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        tile_gemm(sR1, sW2T, te, c0, acc);
        // ---- layer 3 + logits
        float z[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[i][0] = fmaxf(acc[i][0] + b2v.x, 0.f); acc[i][1] = fmaxf(acc[i][1] + b2v.y, 0.f);
            acc[i][2] = fmaxf(acc[i][2] + b2v.z, 0.f); acc[i][3] = fmaxf(acc[i][3] + b2v.w, 0.f);
            float s = acc[i][0] * w3v.x;
            s = fmaf(acc[i][1], w3v.y, s); s = fmaf(acc[i][2], w3v.z, s); s = fmaf(acc[i][3], w3v.w, s);
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            z[i] = s + b3;
        }
        if (tj == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) sZ[te + 16 * i] = z[i];
        }
        __syncthreads();
        // ---- per-edge epilogue: logits out, loss, dz
        if (tid < BM) {
            const int64_t e = e0 + tid;
            const bool ok = e < p.E;
            const float zz = sZ[tid];
            if (ok && p.logits) p.logits[e] = zz;
            const float yy = (ok && p.y) ? p.y[e] : 0.f;
            if (ok && p.y && p.loss_partial) {
                // torch BCEWithLogits(pos_weight): (1-y) z + (1+(pw-1)y) (log1p(exp(-|z|)) + max(-z,0))
                const float lw = fmaf(p.pos_weight - 1.f, yy, 1.f);
                loss_acc += (1.f - yy) * zz + lw * (log1pf(expf(-fabsf(zz))) + fmaxf(-zz, 0.f));
            }
            if (TRAIN) {
                float dzv = 0.f;
                if (ok) {
                    if (p.dlogits) {
                        dzv = p.dlogits[e] * p.scale;
                    } else {
                        // torch's backward: ((pw*y + 1 - y) * sigmoid(z) - pw*y) * grad
                        const float sg = 1.f / (1.f + expf(-zz));
                        const float t = p.pos_weight * yy;
                        dzv = ((t + 1.f - yy) * sg - t) * p.scale;
                    }
                }
                sDZ[tid] = dzv;
            }
        }
        if (TRAIN) {
            __syncthreads();
            // ---- da2 = dz * w3 * [a2 > 0]   (acc holds r2 = relu(a2))
            float dz[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                dz[i] = sDZ[te + 16 * i];
                gw3[0] = fmaf(dz[i], acc[i][0], gw3[0]); gw3[1] = fmaf(dz[i], acc[i][1], gw3[1]);
                gw3[2] = fmaf(dz[i], acc[i][2], gw3[2]); gw3[3] = fmaf(dz[i], acc[i][3], gw3[3]);
                float4 d;
                d.x = acc[i][0] > 0.f ? dz[i] * w3v.x : 0.f;
                d.y = acc[i][1] > 0.f ? dz[i] * w3v.y : 0.f;
                d.z = acc[i][2] > 0.f ? dz[i] * w3v.z : 0.f;
                d.w = acc[i][3] > 0.f ? dz[i] * w3v.w : 0.f;
                gb2[0] += d.x; gb2[1] += d.y; gb2[2] += d.z; gb2[3] += d.w;
                if (tj == 0) gb3 += dz[i];
                *reinterpret_cast<float4 *>(sDA2 + (te + 16 * i) * LDA + c0) = d;
                acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            }
            __syncthreads();
            // ---- GEMM-2: dr1[e][k] = sum_j da2[e][j] W2[j][k];  da1 = dr1 * [r1 > 0]
            tile_gemm(sDA2, sW2, te, c0, acc);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int er = te + 16 * i;
                const float4 r = *reinterpret_cast<const float4 *>(sR1 + er * LDA + c0);
                float4 d;
                d.x = r.x > 0.f ? acc[i][0] : 0.f; d.y = r.y > 0.f ? acc[i][1] : 0.f;
                d.z = r.z > 0.f ? acc[i][2] : 0.f; d.w = r.w > 0.f ? acc[i][3] : 0.f;
                const float sk = sSkip[er];
                gb1[0] += d.x; gb1[1] += d.y; gb1[2] += d.z; gb1[3] += d.w;
                gw1c[0] = fmaf(d.x, sk, gw1c[0]); gw1c[1] = fmaf(d.y, sk, gw1c[1]);
                gw1c[2] = fmaf(d.z, sk, gw1c[2]); gw1c[3] = fmaf(d.w, sk, gw1c[3]);
                const int64_t e = e0 + er;
                if (e < p.E) *reinterpret_cast<float4 *>(p.da1 + e * D + c0) = d;
            }
            // ---- GEMM-3: dW2[j][k] += sum_e da2[e][j] r1[e][k]   (j0 = te*4, k0 = tj*4)
            {
                const int j0 = te * 4;
#pragma unroll 4
                for (int e = 0; e < BM; ++e) {
                    const float4 a = *reinterpret_cast<const float4 *>(sDA2 + e * LDA + j0);
                    const float4 r = *reinterpret_cast<const float4 *>(sR1 + e * LDA + c0);
                    gW2[0][0] = fmaf(a.x, r.x, gW2[0][0]); gW2[0][1] = fmaf(a.x, r.y, gW2[0][1]);
                    gW2[0][2] = fmaf(a.x, r.z, gW2[0][2]); gW2[0][3] = fmaf(a.x, r.w, gW2[0][3]);
                    gW2[1][0] = fmaf(a.y, r.x, gW2[1][0]); gW2[1][1] = fmaf(a.y, r.y, gW2[1][1]);
                    gW2[1][2] = fmaf(a.y, r.z, gW2[1][2]); gW2[1][3] = fmaf(a.y, r.w, gW2[1][3]);
                    gW2[2][0] = fmaf(a.z, r.x, gW2[2][0]); gW2[2][1] = fmaf(a.z, r.y, gW2[2][1]);
                    gW2[2][2] = fmaf(a.z, r.z, gW2[2][2]); gW2[2][3] = fmaf(a.z, r.w, gW2[2][3]);
                    gW2[3][0] = fmaf(a.w, r.x, gW2[3][0]); gW2[3][1] = fmaf(a.w, r.y, gW2[3][1]);
                    gW2[3][2] = fmaf(a.w, r.z, gW2[3][2]); gW2[3][3] = fmaf(a.w, r.w, gW2[3][3]);
                }
            }
        }
        __syncthreads();      // sR1 / sDA2 / index buffers are rewritten by the next tile
    }

    // ---- block epilogue
    if (p.loss_partial) {
        // fixed-order block reduction of the 128 per-edge-slot partial losses
        __shared__ double lred[BM];
        if (tid < BM) lred[tid] = (double)loss_acc;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int i = 0; i < BM; ++i) s += lred[i];
            p.loss_partial[blockIdx.x] = s;
        }
    }
    if (!TRAIN) return;
    float *out = p.partial + (int64_t)blockIdx.x * NGP;
    {
        const int j0 = te * 4;
#pragma unroll
        for (int a = 0; a < 4; ++a)
            *reinterpret_cast<float4 *>(out + G_W2 + (j0 + a) * D + c0) =
                make_float4(gW2[a][0], gW2[a][1], gW2[a][2], gW2[a][3]);
    }
    // column partials held by the 16 `te` groups -> fixed-order sum through shared memory
    float *red = smem;                                   // reuse sR1: [4 kinds][16 te][D]
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        red[(0 * 16 + te) * D + c0 + a] = gb2[a];
        red[(1 * 16 + te) * D + c0 + a] = gw3[a];
        red[(2 * 16 + te) * D + c0 + a] = gb1[a];
        red[(3 * 16 + te) * D + c0 + a] = gw1c[a];
    }
    if (tj == 0) red[4 * 16 * D + te] = gb3;
    __syncthreads();
    {
        const int kind = tid >> 6, c = tid & 63;         // 4 kinds x 64 columns = 256 threads
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < 16; ++t) s += red[(kind * 16 + t) * D + c];
        const int off = kind == 0 ? G_B2 : kind == 1 ? G_W3 : kind == 2 ? G_B1 : G_W1C;
        out[off + c] = s;
        if (tid == 0) {
            float b = 0.f;
            for (int t = 0; t < 16; ++t) b += red[4 * 16 * D + t];
            out[G_B3] = b;
        }
    }
}

__global__ void reduce_loss_kernel(const double *__restrict__ partial, int n, double *__restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += partial[i];
        *out += s;
    }
}

// cosine similarity (F.cosine_similarity, eps = 1e-8 on each norm) / row-wise dot, warp per edge
__global__ void __launch_bounds__(256)
edge_pair_score_kernel(const float *__restrict__ h, int64_t ldh, int32_t feat,
                       const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int64_t E,
                       int mode, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (e >= E) return;
    const float *a = h + (int64_t)src[e] * ldh, *b = h + (int64_t)dst[e] * ldh;
    float ab = 0.f, aa = 0.f, bb = 0.f;
    for (int f = lane; f < feat; f += 32) {
        const float x = a[f], yv = b[f];
        ab = fmaf(x, yv, ab); aa = fmaf(x, x, aa); bb = fmaf(yv, yv, bb);
    }
    ab = warp_sum(ab); aa = warp_sum(aa); bb = warp_sum(bb);
    if (lane == 0) {
        if (mode == 0) {
            const float eps = 1e-8f;
            out[e] = ab / (fmaxf(sqrtf(aa), eps) * fmaxf(sqrtf(bb), eps));
        } else {
            out[e] = ab;
        }
    }
}

static size_t fwd_smem() { return (size_t)(BM * LDA + D * D + 4 * D + 2 * BM + 2 * BM) * 4; }
static size_t bwd_smem() { return fwd_smem() + (size_t)(BM + D * D + BM * LDA) * 4; }

constexpr size_t kLossOff = (size_t)kNumSMs * 2 * NGP * sizeof(float);   // doubles live after the float partials

static int scorer_grid(int64_t E, int blocks_per_sm) {
    const int64_t tiles = (E + BM - 1) / BM;
    const int64_t cap = (int64_t)kNumSMs * blocks_per_sm;
    return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}

}  // namespace pangnn

using namespace pangnn;

extern "C" {

size_t pangnn_edge_score_workspace_bytes(int64_t E) {
    (void)E;
    return kLossOff + (size_t)kNumSMs * 4 * sizeof(double) + 1024;
}

int pangnn_edge_score_fwd(const float *pq, const int32_t *src, const int32_t *dst, const float *skip,
                          const float *w1c, const float *b1, const float *w2, const float *b2,
                          const float *w3, const float *b3, int64_t E, const float *y,
                          float pos_weight, float *logits, double *loss_sum, void *ws,
                          size_t ws_bytes, void *stream) {
    PANGNN_REQUIRE(E >= 0, "negative edge count");
    if (E == 0) return PANGNN_OK;
    PANGNN_REQUIRE(pq && src && dst && b1 && w2 && b2 && w3 && b3, "null pointer");
    PANGNN_REQUIRE(!skip || w1c, "skip feature needs w1c");
    PANGNN_REQUIRE(!loss_sum || (y && ws), "loss needs labels and a workspace");
    PANGNN_REQUIRE((uintptr_t)pq % 16 == 0, "pq must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = scorer_grid(E, 4);
    if (loss_sum && ws_bytes < pangnn_edge_score_workspace_bytes(E)) {
        set_error("edge_score_fwd: workspace too small");
        return PANGNN_EWORKSPACE;
    }
    ScorerArgs a{};
    a.pq = pq; a.src = src; a.dst = dst; a.skip = skip; a.w1c = w1c; a.b1 = b1; a.w2 = w2; a.b2 = b2;
    a.w3 = w3; a.b3 = b3; a.E = E; a.y = loss_sum ? y : nullptr; a.pos_weight = pos_weight;
    a.logits = logits;
    a.loss_partial = loss_sum ? reinterpret_cast<double *>(static_cast<char *>(ws) + kLossOff) : nullptr;
    static bool attr_set = false;
    if (!attr_set) {
        int rc = check_cuda(cudaFuncSetAttribute(edge_score_kernel<false>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)fwd_smem()), "cudaFuncSetAttribute(fwd)");
        if (rc) return rc;
        attr_set = true;
    }
    edge_score_kernel<false><<<grid, kThreads, fwd_smem(), st>>>(a);
    PANGNN_CHECK_LAUNCH("edge_score_fwd");
    if (loss_sum) {
        reduce_loss_kernel<<<1, 32, 0, st>>>(a.loss_partial, grid, loss_sum);
        PANGNN_CHECK_LAUNCH("reduce_loss");
    }
    return PANGNN_OK;
}

int pangnn_edge_score_bwd(const float *pq, const int32_t *src, const int32_t *dst, const float *skip,
                          const float *w1c, const float *b1, const float *w2, const float *b2,
                          const float *w3, const float *b3, int64_t E, const float *dlogits,
                          const float *y, float pos_weight, float scale, float *da1, float *grads,
                          float *logits, double *loss_sum, void *ws, size_t ws_bytes, void *stream) {
    PANGNN_REQUIRE(E >= 0 && grads, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (E == 0) return check_cuda(cudaMemsetAsync(grads, 0, NG * sizeof(float), st), "memset");
    PANGNN_REQUIRE(pq && src && dst && b1 && w2 && b2 && w3 && b3 && da1 && ws, "null pointer");
    PANGNN_REQUIRE(dlogits || y, "need dlogits or labels");
    PANGNN_REQUIRE(!skip || w1c, "skip feature needs w1c");
    PANGNN_REQUIRE((uintptr_t)pq % 16 == 0 && (uintptr_t)da1 % 16 == 0, "pointers must be 16-byte aligned");
    if (ws_bytes < pangnn_edge_score_workspace_bytes(E)) {
        set_error("edge_score_bwd: workspace too small");
        return PANGNN_EWORKSPACE;
    }
    const int grid = scorer_grid(E, 2);
    ScorerArgs a{};
    a.pq = pq; a.src = src; a.dst = dst; a.skip = skip; a.w1c = w1c; a.b1 = b1; a.w2 = w2; a.b2 = b2;
    a.w3 = w3; a.b3 = b3; a.E = E; a.y = y; a.dlogits = dlogits; a.pos_weight = pos_weight;
    a.scale = scale; a.da1 = da1; a.partial = static_cast<float *>(ws);
    a.logits = logits;
    PANGNN_REQUIRE(!loss_sum || y, "loss needs labels");
    a.loss_partial = loss_sum ? reinterpret_cast<double *>(static_cast<char *>(ws) + kLossOff) : nullptr;
    static bool attr_set = false;
    if (!attr_set) {
        int rc = check_cuda(cudaFuncSetAttribute(edge_score_kernel<true>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)bwd_smem()), "cudaFuncSetAttribute(bwd)");
        if (rc) return rc;
        attr_set = true;
    }
    edge_score_kernel<true><<<grid, kThreads, bwd_smem(), st>>>(a);
    PANGNN_CHECK_LAUNCH("edge_score_bwd");
    if (loss_sum) {
        reduce_loss_kernel<<<1, 32, 0, st>>>(a.loss_partial, grid, loss_sum);
        PANGNN_CHECK_LAUNCH("reduce_loss");
    }
    return reduce_partials(a.partial, grid, NG, NGP, grads, st);
}

int pangnn_edge_pair_score(const float *h, int64_t ldh, int32_t feat, const int32_t *src,
                           const int32_t *dst, int64_t E, int mode, float *out, void *stream) {
    PANGNN_REQUIRE(E >= 0, "negative edge count");
    if (E == 0) return PANGNN_OK;
    PANGNN_REQUIRE(h && src && dst && out && feat > 0, "null pointer");
    const unsigned blocks = (unsigned)((E * 32 + 255) / 256);
    edge_pair_score_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(h, ldh, feat, src, dst, E, mode, out);
    PANGNN_CHECK_LAUNCH("edge_pair_score");
    return PANGNN_OK;
}

}  // extern "C"
