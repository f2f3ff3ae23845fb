// Candidate normalisation on the device: hit table -> unique sorted hits -> per (query, target
// genome) segment temperature softmax + Q-score -> compacted (src, dst, w, y) edge list.
//
// Replaces (reference, all pure-Python dict loops):
//   src/preprocessing.py:413-416  groupby / dict(zip)   (duplicate (query,target): last row wins)
//   src/preprocessing.py:370-385  remove_trivial_cases
//   src/preprocessing.py:430-443  softmax_with_temperature (scipy logsumexp, fp64)
//   src/preprocessing.py:454-548  normalize_sim_scores
//   src/preprocessing.py:73-118   build_edge_index, :264-325 map_edge_weights,
//   src/preprocessing.py:122-156  map_labels_to_edge_index
//
// Roofline: HBM; algorithmic bytes of the softmax kernel = n*(4 q + 4 t + 8 bits + 4 w + 4 y + 1 keep)
// + 8 per segment.  One warp per segment; lanes stride the segment, fp64 shuffle reductions.
#include "common.cuh"

namespace pangnn {

int exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total, void *ws,
                       size_t ws_bytes, cudaStream_t st);
int sort_pairs_u64(const uint64_t *keys_in, const uint32_t *vals_in, uint64_t *keys_out,
                   uint32_t *vals_out, int64_t n, int key_bits, void *ws, size_t ws_bytes,
                   cudaStream_t st);

static int bits_for_nodes(int64_t n) {
    int b = 1;
    while (((int64_t)1 << b) < n) ++b;
    return b;
}

__global__ void hits_pack_kernel(const int32_t *__restrict__ q, const int32_t *__restrict__ t,
                                 int64_t n, int nbits, uint64_t *__restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ((uint64_t)(uint32_t)q[i] << nbits) | (uint64_t)(uint32_t)t[i];
}

// flag = 1 on the LAST element of every run of equal keys (stable sort => last original row)
__global__ void hits_last_flag_kernel(const uint64_t *__restrict__ keys, int64_t n,
                                      uint32_t *__restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == n - 1 || keys[i] != keys[i + 1]) ? 1u : 0u;
}

__global__ void hits_unpack_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                   const uint32_t *__restrict__ pos, const double *__restrict__ bits,
                                   int64_t n, int nbits, int32_t *__restrict__ q,
                                   int32_t *__restrict__ t, double *__restrict__ bits_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i != n - 1 && keys[i] == keys[i + 1]) return;
    const uint32_t p = pos[i];
    const uint64_t k = keys[i];
    q[p] = (int32_t)(k >> nbits);
    t[p] = (int32_t)(k & ((1ull << nbits) - 1ull));
    bits_out[p] = bits[vals[i]];
}

// head[i] = 1 where a (query, genome_of[target]) segment starts
__global__ void seg_head_kernel(const int32_t *__restrict__ q, const int32_t *__restrict__ t,
                                const int32_t *__restrict__ genome_of, int64_t n,
                                uint32_t *__restrict__ head) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = 1u;
    if (i > 0) h = (q[i] != q[i - 1] || genome_of[t[i]] != genome_of[t[i - 1]]) ? 1u : 0u;
    head[i] = h;
}

// seg_start[seg_id] = i for heads; seg_start[num_seg] = n
__global__ void seg_start_kernel(const uint32_t *__restrict__ head_pos, const int32_t *__restrict__ q,
                                 const int32_t *__restrict__ t, const int32_t *__restrict__ genome_of,
                                 int64_t n, const uint32_t *__restrict__ num_seg,
                                 int64_t *__restrict__ seg_start) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool h = (i == 0) || q[i] != q[i - 1] || genome_of[t[i]] != genome_of[t[i - 1]];
    if (h) seg_start[head_pos[i]] = i;
    if (i == n - 1) seg_start[*num_seg] = n;
}

__global__ void __launch_bounds__(256)
segment_softmax_q_kernel(const int32_t *__restrict__ q, const int32_t *__restrict__ t,
                         const double *__restrict__ bits, const int64_t *__restrict__ seg_start,
                         const uint32_t *__restrict__ num_seg, const int32_t *__restrict__ group_of,
                         double temp, double eps, double pseudo, int drop_trivial,
                         float *__restrict__ w, float *__restrict__ y, uint32_t *__restrict__ keep) {
    const int lane = threadIdx.x & 31;
    const int64_t seg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (seg >= (int64_t)*num_seg) return;
    const int64_t s0 = seg_start[seg], s1 = seg_start[seg + 1];
    const bool trivial = drop_trivial && (s1 - s0) == 1;        // the self hit counts (a1)

    // pass 1: max and member count over non-self members
    double mx = -INFINITY;
    int members = 0;
    for (int64_t i = s0 + lane; i < s1; i += 32) {
        if (q[i] != t[i]) {
            mx = fmax(mx, bits[i] / temp);
            ++members;
        }
    }
    mx = warp_max(mx);
    members = __reduce_add_sync(0xffffffffu, members);
    // pass 2: sum of exponentials (fp64)
    double sum = 0.0;
    if (members > 1)
        for (int64_t i = s0 + lane; i < s1; i += 32)
            if (q[i] != t[i]) sum += exp(bits[i] / temp - mx);
    sum = warp_sum(sum);
    // pass 3: emit
    for (int64_t i = s0 + lane; i < s1; i += 32) {
        const int32_t qi = q[i], ti = t[i];
        const bool self = qi == ti;
        float wi = 0.f;
        if (!self) {
            double om = 0.0;                                       // 1 - p
            if (members > 1) {
                const double e = exp(bits[i] / temp - mx);
                om = (sum - e) / sum;                              // = 1 - softmax, no 1-p cancellation
            }
            om = fmin(fmax(om, eps), 1.0 - eps);                   // np.clip(1-p, eps, 1-eps)
            wi = (float)(-10.0 * log10(om) + pseudo);
        }
        w[i] = wi;
        float yi = 0.f;
        if (group_of && !self) {
            const int32_t gq = group_of[qi];
            yi = (gq >= 0 && gq == group_of[ti]) ? 1.f : 0.f;
        }
        y[i] = yi;
        keep[i] = (!self && !trivial) ? 1u : 0u;
    }
}

__global__ void compact_edges_kernel(const int32_t *__restrict__ q, const int32_t *__restrict__ t,
                                     const float *__restrict__ w, const float *__restrict__ y,
                                     const uint32_t *__restrict__ keep, const uint32_t *__restrict__ pos,
                                     int64_t n, int32_t *__restrict__ src, int32_t *__restrict__ dst,
                                     float *__restrict__ w_out, float *__restrict__ y_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const uint32_t p = pos[i];
    src[p] = q[i];
    dst[p] = t[i];
    w_out[p] = w[i];
    y_out[p] = y[i];
}

}  // namespace pangnn

using namespace pangnn;

extern "C" {

size_t pangnn_scan_workspace_bytes(int64_t n);
size_t pangnn_sort_pairs_workspace_bytes(int64_t n);

size_t pangnn_hits_sort_unique_workspace_bytes(int64_t n) {
    return 2 * align_up((size_t)n * 8, 256) + 3 * align_up((size_t)n * 4, 256) +
           pangnn_sort_pairs_workspace_bytes(n) + pangnn_scan_workspace_bytes(n) + 2048;
}

int pangnn_hits_sort_unique(const int32_t *q, const int32_t *t, const double *bits, int64_t n,
                            int32_t num_nodes, int32_t *q_out, int32_t *t_out, double *bits_out,
                            uint32_t *count, void *ws, size_t ws_bytes, void *stream) {
    PANGNN_REQUIRE(count, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return check_cuda(cudaMemsetAsync(count, 0, 4, st), "memset");
    PANGNN_REQUIRE(q && t && bits && q_out && t_out && bits_out && ws, "null pointer");
    if (ws_bytes < pangnn_hits_sort_unique_workspace_bytes(n)) {
        set_error("hits_sort_unique: workspace too small");
        return PANGNN_EWORKSPACE;
    }
    Workspace w(ws, ws_bytes);
    uint64_t *ka = w.take<uint64_t>(n);
    uint64_t *kb = w.take<uint64_t>(n);
    uint32_t *vals = w.take<uint32_t>(n);
    uint32_t *flag = w.take<uint32_t>(n);
    const size_t scan_bytes = pangnn_scan_workspace_bytes(n);
    void *scan_ws = w.take<char>(scan_bytes);
    const size_t sort_bytes = pangnn_sort_pairs_workspace_bytes(n);
    void *sort_ws = w.take<char>(sort_bytes);
    const int nbits = bits_for_nodes(num_nodes > 1 ? num_nodes : 2);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    hits_pack_kernel<<<blocks, 256, 0, st>>>(q, t, n, nbits, ka);
    PANGNN_CHECK_LAUNCH("hits_pack");
    int rc = sort_pairs_u64(ka, nullptr, kb, vals, n, 2 * nbits, sort_ws, sort_bytes, st);
    if (rc) return rc;
    hits_last_flag_kernel<<<blocks, 256, 0, st>>>(kb, n, flag);
    PANGNN_CHECK_LAUNCH("hits_last_flag");
    rc = exclusive_scan_u32(flag, flag, n, count, scan_ws, scan_bytes, st);
    if (rc) return rc;
    hits_unpack_kernel<<<blocks, 256, 0, st>>>(kb, vals, flag, bits, n, nbits, q_out, t_out, bits_out);
    PANGNN_CHECK_LAUNCH("hits_unpack");
    return PANGNN_OK;
}

size_t pangnn_hits_normalize_workspace_bytes(int64_t n) {
    return align_up((size_t)(n + 1) * 8, 256) + 4 * align_up((size_t)n * 4, 256) +
           pangnn_scan_workspace_bytes(n) + 2048;
}

int pangnn_hits_normalize(const int32_t *q, const int32_t *t, const double *bits, int64_t n,
                          const int32_t *genome_of, const int32_t *group_of, double temp, double eps,
                          double pseudo, int drop_trivial, int32_t *src, int32_t *dst, float *w_out,
                          float *y_out, uint32_t *count, void *ws, size_t ws_bytes, void *stream) {
    PANGNN_REQUIRE(count, "null pointer");
    PANGNN_REQUIRE(temp > 0.0, "temperature must be > 0");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return check_cuda(cudaMemsetAsync(count, 0, 4, st), "memset");
    PANGNN_REQUIRE(q && t && bits && genome_of && src && dst && w_out && y_out && ws, "null pointer");
    if (ws_bytes < pangnn_hits_normalize_workspace_bytes(n)) {
        set_error("hits_normalize: workspace too small");
        return PANGNN_EWORKSPACE;
    }
    Workspace wk(ws, ws_bytes);
    int64_t *seg_start = wk.take<int64_t>(n + 1);
    uint32_t *flag = wk.take<uint32_t>(n);           // heads, then keep flags / positions
    float *w = wk.take<float>(n);
    float *y = wk.take<float>(n);
    uint32_t *num_seg = wk.take<uint32_t>(64);
    const size_t scan_bytes = pangnn_scan_workspace_bytes(n);
    void *scan_ws = wk.take<char>(scan_bytes);
    uint32_t *keep = wk.take<uint32_t>(n);
    const unsigned blocks = (unsigned)((n + 255) / 256);

    seg_head_kernel<<<blocks, 256, 0, st>>>(q, t, genome_of, n, flag);
    PANGNN_CHECK_LAUNCH("seg_head");
    int rc = exclusive_scan_u32(flag, flag, n, num_seg, scan_ws, scan_bytes, st);
    if (rc) return rc;
    seg_start_kernel<<<blocks, 256, 0, st>>>(flag, q, t, genome_of, n, num_seg, seg_start);
    PANGNN_CHECK_LAUNCH("seg_start");
    // one warp per potential segment (num_seg <= n lives on the device; surplus warps exit)
    const unsigned wblocks = (unsigned)((n * 32 + 255) / 256);
    segment_softmax_q_kernel<<<wblocks, 256, 0, st>>>(q, t, bits, seg_start, num_seg, group_of,
                                                      temp, eps, pseudo, drop_trivial, w, y, keep);
    PANGNN_CHECK_LAUNCH("segment_softmax_q");
    rc = exclusive_scan_u32(keep, flag, n, count, scan_ws, scan_bytes, st);
    if (rc) return rc;
    compact_edges_kernel<<<blocks, 256, 0, st>>>(q, t, w, y, keep, flag, n, src, dst, w_out, y_out);
    PANGNN_CHECK_LAUNCH("compact_edges");
    return PANGNN_OK;
}

}  // extern "C"
