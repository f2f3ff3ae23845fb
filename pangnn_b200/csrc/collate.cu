// Batch collation on the device (replaces torch_geometric's Batch.from_data_list as the
// reference uses it through DataLoader, pangnn.py:121,152-153; semantics in SURVEY A.3): the graphs of a
// split live packed back to back in one device arena per attribute; a batch is a list of graph ids.
//   * attributes whose name contains "index" ([2, e] int64): concatenated on the last dim, every value
//     shifted by the cumulative node count of the graphs before it in the batch;
//   * every other tensor attribute ([k, ...]): concatenated on dim 0;
//   * `batch` (graph slot of every node) and `ptr` (cumulative node counts) are produced alongside.
// Two launches per batch whatever the number of attributes: (1) one warp per attribute scans the segment
// sizes of the batch's graphs into output offsets, (2) one CTA per (graph slot, attribute) copies its
// segment.  The reference's regime is 32 sub-graphs of ~12 nodes per step: the host-side
// collation it replaces costs ~1 ms of Python per batch, half of the whole step.
#include "common.cuh"

namespace pangnn {

struct CollateArgs {
    pangnn_collate_attr a[PANGNN_COLLATE_MAX_ATTRS];
    int32_t n;
    int32_t node_attr;        // index of the attribute whose segments are the graphs' node ranges
};

__global__ void __launch_bounds__(32)
collate_offsets_kernel(const CollateArgs args, const int32_t *__restrict__ ids, int32_t B,
                       int64_t *__restrict__ off /* [n, B+1] */) {
    const pangnn_collate_attr &at = args.a[blockIdx.x];
    int64_t *o = off + (int64_t)blockIdx.x * (B + 1);
    const int lane = threadIdx.x;
    int64_t running = 0;
    for (int32_t base = 0; base < B; base += 32) {
        const int32_t i = base + lane;
        int64_t sz = 0;
        if (i < B) {
            const int32_t g = ids[i];
            sz = at.seg_ptr[g + 1] - at.seg_ptr[g];
        }
        int64_t inc = sz;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (i < B) o[i] = running + inc - sz;
        running += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) o[B] = running;
}

template <typename T>
__device__ __forceinline__ void copy_segment(const pangnn_collate_attr &at, int64_t s0, int64_t len, int64_t d0,
                                             int64_t add) {
    for (int r = 0; r < at.rows; ++r) {
        const T *src = reinterpret_cast<const T *>(at.src) + (int64_t)r * at.src_row_stride + s0;
        T *dst = reinterpret_cast<T *>(at.dst) + (int64_t)r * at.dst_row_stride + d0;
        for (int64_t i = threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i] + (T)add;
    }
}

__global__ void __launch_bounds__(128)
collate_copy_kernel(const CollateArgs args, const int32_t *__restrict__ ids, int32_t B,
                    const int64_t *__restrict__ off /* [n, B+1] */) {
    const int32_t b = blockIdx.x;
    const pangnn_collate_attr &at = args.a[blockIdx.y];
    const int32_t g = ids[b];
    const int64_t w = at.width;
    const int64_t s0 = at.seg_ptr[g] * w, len = (at.seg_ptr[g + 1] - at.seg_ptr[g]) * w;
    const int64_t d0 = off[(int64_t)blockIdx.y * (B + 1) + b] * w;
    if (at.kind == PANGNN_COLLATE_FILL_SLOT) {               // `batch`: graph slot of every node
        int64_t *dst = reinterpret_cast<int64_t *>(at.dst) + d0;
        for (int64_t i = threadIdx.x; i < len; i += blockDim.x) dst[i] = b;
        return;
    }
    const int64_t add = at.kind == PANGNN_COLLATE_INDEX ? off[(int64_t)args.node_attr * (B + 1) + b] : 0;
    if (at.elem_bytes == 8) copy_segment<int64_t>(at, s0, len, d0, add);
    else copy_segment<int32_t>(at, s0, len, d0, 0);           // 4-byte payloads are moved as bit patterns
}

}  // namespace pangnn

using namespace pangnn;

extern "C" {

int pangnn_collate(const pangnn_collate_attr *attrs, int32_t num_attrs, int32_t node_attr, const int32_t *graph_ids,
                   int32_t batch_size, int64_t *offsets, void *stream) {
    PANGNN_REQUIRE(attrs && num_attrs > 0 && num_attrs <= PANGNN_COLLATE_MAX_ATTRS, "1..PANGNN_COLLATE_MAX_ATTRS attributes");
    PANGNN_REQUIRE(node_attr >= 0 && node_attr < num_attrs, "node_attr out of range");
    PANGNN_REQUIRE(batch_size >= 0 && batch_size <= 65535 * 32, "bad batch size");
    if (batch_size == 0) return PANGNN_OK;
    PANGNN_REQUIRE(graph_ids && offsets, "null pointer");
    CollateArgs args;
    args.n = num_attrs;
    args.node_attr = node_attr;
    for (int i = 0; i < num_attrs; ++i) {
        const pangnn_collate_attr &a = attrs[i];
        PANGNN_REQUIRE(a.seg_ptr && (a.elem_bytes == 4 || a.elem_bytes == 8) && (a.rows == 1 || a.rows == 2) && a.width >= 1,
                       "bad attribute descriptor");
        PANGNN_REQUIRE(a.kind == PANGNN_COLLATE_COPY || a.elem_bytes == 8, "index / batch attributes are int64");
        args.a[i] = a;
    }
    cudaStream_t st = (cudaStream_t)stream;
    collate_offsets_kernel<<<num_attrs, 32, 0, st>>>(args, graph_ids, batch_size, offsets);
    PANGNN_CHECK_LAUNCH("collate_offsets");
    dim3 grid((unsigned)batch_size, (unsigned)num_attrs);
    PANGNN_REQUIRE(batch_size <= 0x7fffffff, "batch too large");
    collate_copy_kernel<<<grid, 128, 0, st>>>(args, graph_ids, batch_size, offsets);
    PANGNN_CHECK_LAUNCH("collate_copy");
    return PANGNN_OK;
}

}  // extern "C"
