// Shared declarations of the fused edge scorer (edge_scorer.cu: C ABI + small kernels;
// edge_scorer_tc.cu: the tcgen05 3xTF32 tile kernel).
#pragma once
#include "common.cuh"

namespace pangnn {

constexpr int kScD = PANGNN_SCORER_D;            // 64
constexpr int kScNG = PANGNN_SCORER_NGRADS;
constexpr int kScNGP = (kScNG + 3) / 4 * 4;      // per-block partial stride (16 B aligned rows)
// layout of the small-gradient vector
constexpr int kG_W2 = 0, kG_B2 = kScD * kScD, kG_W3 = kG_B2 + kScD, kG_B3 = kG_W3 + kScD,
              kG_B1 = kG_B3 + 1, kG_W1C = kG_B1 + kScD;

struct ScorerArgs {
    const float *pq;
    const int32_t *src, *dst;
    const float *skip, *w1c, *b1, *w2, *b2, *w3, *b3;
    int64_t E;
    const float *y, *dlogits;
    float pos_weight, scale;
    float *logits;      // optional
    float *prob;        // optional: sigmoid(logit)                      (pangnn.py:220, src/predict.py:54)
    int32_t *pred;      // optional: sigmoid(logit) >= threshold         (pangnn.py:221, src/predict.py:55)
    float threshold;
    float *da1;         // TRAIN: [E, 64]
    float *partial;     // TRAIN: [grid][kScNGP]
    double *loss_partial;   // optional: [grid]
};

// launches the tile kernel; *grid_out = number of CTAs (= rows of `partial` / `loss_partial`)
int launch_edge_score_tc(const ScorerArgs &a, bool train, int *grid_out, cudaStream_t st);
int edge_score_tc_max_grid();
// training form (edge_scorer_train.cu): one CTA per SM
int launch_edge_score_train(const ScorerArgs &a, int *grid_out, cudaStream_t st);

}  // namespace pangnn
