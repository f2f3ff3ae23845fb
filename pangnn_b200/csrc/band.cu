// Whole-graph neighbour band (src/dataset.py:351-366): edges i -> j for
// j in [i-n, i+n] ∩ [0, N), self loop included, over the concatenated genome order, emitted in the
// reference's loop order (i ascending, j ascending).  Closed form, no scan and no host sync: row i
// holds deg(i) = 2n+1 - max(n-i, 0) - max(i+n-(N-1), 0) edges and starts at
//   prefix(i) = (2n+1) i - L(i) - H(i),  L(i) = sum_{r<i} max(n-r, 0),  H(i) = sum_{r<i} max(r+n-(N-1), 0).
// One thread per (i, offset) slot of the dense [N, 2n+1] band; writes are contiguous apart from the
// 2 x n(n+1)/2 clipped slots.  The outputs are two separate pointers so that the band can be written
// straight into the tail of a union edge list (a11, src/dataset.py:373-381).
#include "common.cuh"

namespace pangnn {

__host__ __device__ inline int64_t band_prefix(int64_t i, int64_t N, int64_t n) {
    const int64_t a = i < n ? i : n;                               // rows r < i clipped below: r < n
    const int64_t low = a * n - a * (a - 1) / 2;                   // sum_{r<a} (n - r)
    const int64_t f = N - n;                                       // rows r >= f are clipped above by r - f + 1
    const int64_t f0 = f > 0 ? f : 0;
    const int64_t m = i > f0 ? i - f0 : 0;                         // clipped rows below i: r = f0 .. i-1
    const int64_t high = m * (f0 - f + 1) + m * (m - 1) / 2;
    return (2 * n + 1) * i - low - high;
}

__global__ void __launch_bounds__(256)
neighbour_band_kernel(int64_t N, int32_t n, int64_t *__restrict__ out_src, int64_t *__restrict__ out_dst) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t width = 2 * (int64_t)n + 1;
    if (slot >= N * width) return;
    const int64_t i = slot / width;
    const int64_t j = i - n + (slot % width);
    if (j < 0 || j >= N) return;
    const int64_t lo = i - n > 0 ? i - n : 0;
    const int64_t pos = band_prefix(i, N, n) + (j - lo);
    out_src[pos] = i;
    out_dst[pos] = j;
}

// CSR of the union graph [sim ; band] (a11) from the CSR of the sim edges, without sorting the union list:
// row r of the union is the merge of the sim row (columns ascending, ties in edge order) with the band
// columns max(r-n,0) .. min(r+n,N-1); on equal columns the sim entry comes first (its edge id is smaller),
// which is exactly what the stable sort of the concatenated list produces.  One thread per row.
__global__ void __launch_bounds__(256)
csr_merge_band_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                      const uint32_t *__restrict__ perm, int64_t E, int64_t N, int32_t n, int by_dst,
                      int64_t *__restrict__ rowptr_u, int32_t *__restrict__ col_u, uint32_t *__restrict__ perm_u) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    int64_t a = rowptr[r];
    const int64_t b = rowptr[r + 1];
    const int64_t start = band_prefix(r, N, n);
    int64_t out = a + start;
    rowptr_u[r] = out;
    if (r == N - 1) rowptr_u[N] = E + band_prefix(N, N, n);
    const int64_t lo = r - n > 0 ? r - n : 0, hi = r + n < N - 1 ? r + n : N - 1;
    int64_t c = lo;
    int32_t sc = a < b ? col[a] : 0;
    while (a < b || c <= hi) {
        if (a < b && (c > hi || (int64_t)sc <= c)) {
            col_u[out] = sc;
            perm_u[out] = perm[a];
            ++a;
            if (a < b) sc = col[a];
        } else {
            // band edge (src, dst) = by_dst ? (c, r) : (r, c); its position in the band list
            const int64_t src = by_dst ? c : r, dst = by_dst ? r : c;
            const int64_t slo = src - n > 0 ? src - n : 0;
            col_u[out] = (int32_t)c;
            perm_u[out] = (uint32_t)(E + (by_dst ? band_prefix(src, N, n) : start) + (dst - slo));
            ++c;
        }
        ++out;
    }
}

}  // namespace pangnn

using namespace pangnn;

extern "C" {

int64_t pangnn_neighbour_band_edges(int64_t num_nodes, int32_t n) {
    if (num_nodes <= 0 || n < 0) return 0;
    return band_prefix(num_nodes, num_nodes, n);
}

int pangnn_neighbour_band(int64_t num_nodes, int32_t n, int64_t *out_src, int64_t *out_dst, void *stream) {
    if (num_nodes <= 0) return PANGNN_OK;
    PANGNN_REQUIRE(n >= 0 && out_src && out_dst, "bad arguments");
    const int64_t slots = num_nodes * (2 * (int64_t)n + 1);
    PANGNN_REQUIRE((slots + 255) / 256 < ((int64_t)1 << 31), "band too large for one launch");
    neighbour_band_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, (cudaStream_t)stream>>>(num_nodes, n, out_src, out_dst);
    PANGNN_CHECK_LAUNCH("neighbour_band");
    return PANGNN_OK;
}

int pangnn_csr_merge_band(const int64_t *rowptr, const int32_t *col, const uint32_t *perm, int64_t num_edges,
                          int64_t num_nodes, int32_t n, int by_dst, int64_t *rowptr_u, int32_t *col_u,
                          uint32_t *perm_u, void *stream) {
    PANGNN_REQUIRE(num_nodes > 0 && n >= 0 && num_edges >= 0, "bad arguments");
    PANGNN_REQUIRE(rowptr && rowptr_u && col_u && perm_u && (num_edges == 0 || (col && perm)), "null pointer");
    PANGNN_REQUIRE(num_edges + band_prefix(num_nodes, num_nodes, n) < ((int64_t)1 << 32), "union too large for uint32 perm");
    csr_merge_band_kernel<<<(unsigned)((num_nodes + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        rowptr, col, perm, num_edges, num_nodes, n, by_dst, rowptr_u, col_u, perm_u);
    PANGNN_CHECK_LAUNCH("csr_merge_band");
    return PANGNN_OK;
}

}  // extern "C"
