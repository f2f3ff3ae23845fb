// Halo exchange of the genome-partitioned path over NVLink peer memory (new — the
// reference has no multi-GPU path).  The destination / source pointers of these kernels are PEER
// device pointers (symmetric-memory allocations mapped into this process), so the rows travel as
// plain coalesced 128-bit stores / loads through NVSwitch, without staging buffers or a separate
// communication kernel:
//   push:      peer_ext[slot0 + k][:] = own[send_idx[k]][:]          (forward: layer inputs of remote sources)
//   pull-add:  own[send_idx[k]][:]  += peer_dext[slot0 + k][:]       (backward: gradients of those rows;
//              indices are unique per peer and peers are processed in rank order -> deterministic)
// node_linear.cu can do the push from its epilogue (fused GEMM -> halo all-gather); these standalone
// kernels serve the layers whose producer is not ours and the backward direction.
#include "common.cuh"

namespace pangnn {

// one float4 per thread; a row of F floats = F/4 consecutive threads
__global__ void __launch_bounds__(256)
rows_gather_copy_kernel(const float *__restrict__ src, int64_t ld_src, const int32_t *__restrict__ idx, int64_t n,
                        int32_t fq, float *__restrict__ dst, int64_t ld_dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * fq) return;
    const int64_t k = i / fq;
    const int f = (int)(i % fq);
    const int64_t r = idx ? idx[k] : k;
    const float4 v = __ldg(reinterpret_cast<const float4 *>(src + r * ld_src) + f);
    reinterpret_cast<float4 *>(dst + k * ld_dst)[f] = v;
}

__global__ void __launch_bounds__(256)
rows_scatter_add_kernel(const float *__restrict__ src, int64_t ld_src, const int32_t *__restrict__ idx, int64_t n,
                        int32_t fq, float *__restrict__ dst, int64_t ld_dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * fq) return;
    const int64_t k = i / fq;
    const int f = (int)(i % fq);
    const int64_t r = idx ? idx[k] : k;
    const float4 v = *(reinterpret_cast<const float4 *>(src + k * ld_src) + f);      // peer memory: plain load
    float4 *d = reinterpret_cast<float4 *>(dst + r * ld_dst) + f;
    float4 o = *d;
    o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
    *d = o;
}

}  // namespace pangnn

using namespace pangnn;

extern "C" {

int pangnn_rows_gather_copy(const float *src, int64_t ld_src, const int32_t *idx, int64_t n, int32_t feat, float *dst,
                            int64_t ld_dst, void *stream) {
    if (n <= 0) return PANGNN_OK;
    PANGNN_REQUIRE(src && dst && feat > 0 && feat % 4 == 0, "bad arguments");
    PANGNN_REQUIRE(ld_src % 4 == 0 && ld_dst % 4 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0,
                   "rows must be 16-byte aligned");
    const int64_t total = n * (feat / 4);
    rows_gather_copy_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, ld_src, idx, n, feat / 4,
                                                                                             dst, ld_dst);
    PANGNN_CHECK_LAUNCH("rows_gather_copy");
    return PANGNN_OK;
}

int pangnn_rows_scatter_add(const float *src, int64_t ld_src, const int32_t *idx, int64_t n, int32_t feat, float *dst,
                            int64_t ld_dst, void *stream) {
    if (n <= 0) return PANGNN_OK;
    PANGNN_REQUIRE(src && dst && feat > 0 && feat % 4 == 0, "bad arguments");
    PANGNN_REQUIRE(ld_src % 4 == 0 && ld_dst % 4 == 0 && (uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0,
                   "rows must be 16-byte aligned");
    const int64_t total = n * (feat / 4);
    rows_scatter_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, ld_src, idx, n, feat / 4,
                                                                                             dst, ld_dst);
    PANGNN_CHECK_LAUNCH("rows_scatter_add");
    return PANGNN_OK;
}

}  // extern "C"
