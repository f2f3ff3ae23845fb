// C ABI of the fused edge scorer (argument checks, workspace carving, second-stage reductions) and
// the cosine / dot decoders.  The tile kernel itself — endpoint gather, 3-layer MLP, BCE loss and the
// whole backward on the tcgen05 tensor cores — is edge_scorer_tc.cu; see its header for the design.
// Replaces src/gnn.py:171-180,202-207, pangnn.py:98,203 and their autograd backward (pangnn.py:207).
// The per-node gradients are sorted-segment reductions of da1 (pangnn_gcn_aggregate with val = NULL
// over the by-source / by-destination CSR of the scored edges) — no atomics anywhere.
#include "edge_scorer.cuh"

namespace pangnn {

int reduce_partials(const float *partial, int64_t nblocks, int32_t width, int32_t stride, float *out,
                    cudaStream_t st);

constexpr int NG = kScNG;
constexpr int NGP = kScNGP;

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

// doubles (per-CTA loss partials) live after the float partials
constexpr size_t kLossOff = (size_t)kNumSMs * 2 * NGP * sizeof(float);

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
    if (loss_sum && ws_bytes < pangnn_edge_score_workspace_bytes(E)) {
        set_error("edge_score_fwd: workspace too small");
        return PANGNN_EWORKSPACE;
    }
    ScorerArgs a{};
    a.pq = pq; a.src = src; a.dst = dst; a.skip = skip; a.w1c = w1c; a.b1 = b1; a.w2 = w2; a.b2 = b2;
    a.w3 = w3; a.b3 = b3; a.E = E; a.y = loss_sum ? y : nullptr; a.pos_weight = pos_weight;
    a.logits = logits;
    a.loss_partial = loss_sum ? reinterpret_cast<double *>(static_cast<char *>(ws) + kLossOff) : nullptr;
    int grid = 0;
    int rc = launch_edge_score_tc(a, false, &grid, st);
    if (rc) return rc;
    if (loss_sum) {
        reduce_loss_kernel<<<1, 32, 0, st>>>(a.loss_partial, grid, loss_sum);
        PANGNN_CHECK_LAUNCH("reduce_loss");
    }
    return PANGNN_OK;
}

int pangnn_edge_score_predict(const float *pq, const int32_t *src, const int32_t *dst, const float *skip,
                              const float *w1c, const float *b1, const float *w2, const float *b2,
                              const float *w3, const float *b3, int64_t E, float threshold, float *logits,
                              float *prob, int32_t *pred, void *stream) {
    PANGNN_REQUIRE(E >= 0, "negative edge count");
    if (E == 0) return PANGNN_OK;
    PANGNN_REQUIRE(pq && src && dst && b1 && w2 && b2 && w3 && b3, "null pointer");
    PANGNN_REQUIRE(!skip || w1c, "skip feature needs w1c");
    PANGNN_REQUIRE((uintptr_t)pq % 16 == 0, "pq must be 16-byte aligned");
    ScorerArgs a{};
    a.pq = pq; a.src = src; a.dst = dst; a.skip = skip; a.w1c = w1c; a.b1 = b1; a.w2 = w2; a.b2 = b2;
    a.w3 = w3; a.b3 = b3; a.E = E; a.logits = logits; a.prob = prob; a.pred = pred; a.threshold = threshold;
    int grid = 0;
    return launch_edge_score_tc(a, false, &grid, (cudaStream_t)stream);
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
    ScorerArgs a{};
    a.pq = pq; a.src = src; a.dst = dst; a.skip = skip; a.w1c = w1c; a.b1 = b1; a.w2 = w2; a.b2 = b2;
    a.w3 = w3; a.b3 = b3; a.E = E; a.y = y; a.dlogits = dlogits; a.pos_weight = pos_weight;
    a.scale = scale; a.da1 = da1; a.partial = static_cast<float *>(ws);
    a.logits = logits;
    PANGNN_REQUIRE(!loss_sum || y, "loss needs labels");
    a.loss_partial = loss_sum ? reinterpret_cast<double *>(static_cast<char *>(ws) + kLossOff) : nullptr;
    int grid = 0;
    int rc = launch_edge_score_tc(a, true, &grid, st);
    if (rc) return rc;
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
