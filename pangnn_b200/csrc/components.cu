// Output side of the path: ortholog groups = connected components of the edges the model
// predicts positive.  The reference's writer (src/postprocessing.py:5-36, unused) walks the edges once and
// appends / grows Python sets; as written it never merges two existing sets and re-appends every pair, so the
// behaviour DEFINED here is the intended one: true connected components, label = smallest gene id of the
// component (deterministic whatever the order of the atomics).
//   init      parent[i] = i
//   hook      for every selected edge (u, v): ru = root(u), rv = root(v); atomicMin(parent[max], min)
//   compress  parent[i] = root(i)            (pointer jumping)
// repeated until a hook pass changes nothing (a handful of rounds: the trees are flattened every round).
#include "common.cuh"

namespace pangnn {

__device__ __forceinline__ int32_t cc_root(const int32_t *parent, int32_t i) {
    int32_t p = parent[i];
    while (p != i) {
        i = p;
        p = parent[i];
    }
    return i;
}

__global__ void cc_init_kernel(int32_t *parent, int32_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) parent[i] = (int32_t)i;
}

__global__ void __launch_bounds__(256)
cc_hook_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst, const int32_t *__restrict__ select,
               int64_t E, int32_t *parent, int32_t *changed) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E || (select && !select[e])) return;
    int32_t ru = cc_root(parent, src[e]), rv = cc_root(parent, dst[e]);
    while (ru != rv) {                                        // hook the larger root under the smaller one
        const int32_t hi = ru > rv ? ru : rv, lo = ru > rv ? rv : ru;
        const int32_t old = atomicMin(&parent[hi], lo);
        *changed = 1;
        if (old == hi) break;                                 // hi was a root: hooked
        ru = cc_root(parent, old);                            // somebody re-parented hi meanwhile: merge with that tree
        rv = lo;
    }
}

__global__ void cc_compress_kernel(int32_t *parent, int32_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) parent[i] = cc_root(parent, (int32_t)i);
}

}  // namespace pangnn

using namespace pangnn;

extern "C" {

int pangnn_components_init(int32_t *labels, int32_t num_nodes, void *stream) {
    if (num_nodes <= 0) return PANGNN_OK;
    PANGNN_REQUIRE(labels, "null pointer");
    cc_init_kernel<<<(unsigned)((num_nodes + 255) / 256), 256, 0, (cudaStream_t)stream>>>(labels, num_nodes);
    PANGNN_CHECK_LAUNCH("cc_init");
    return PANGNN_OK;
}

int pangnn_components_round(const int32_t *src, const int32_t *dst, const int32_t *select, int64_t num_edges,
                            int32_t *labels, int32_t num_nodes, int32_t *changed, void *stream) {
    PANGNN_REQUIRE(labels && changed && num_nodes >= 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_cuda(cudaMemsetAsync(changed, 0, sizeof(int32_t), st), "memset");
    if (rc) return rc;
    if (num_edges > 0) {
        PANGNN_REQUIRE(src && dst, "null pointer");
        cc_hook_kernel<<<(unsigned)((num_edges + 255) / 256), 256, 0, st>>>(src, dst, select, num_edges, labels, changed);
        PANGNN_CHECK_LAUNCH("cc_hook");
    }
    if (num_nodes > 0) {
        cc_compress_kernel<<<(unsigned)((num_nodes + 255) / 256), 256, 0, st>>>(labels, num_nodes);
        PANGNN_CHECK_LAUNCH("cc_compress");
    }
    return PANGNN_OK;
}

}  // extern "C"
