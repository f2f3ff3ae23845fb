// Device primitives for graph assembly: exclusive scan, stable LSD radix sort of (u64 key, u32 value)
// pairs, and the CSR build on top of them.  Replaces the Python dict/set/list loops of
// src/preprocessing.py:73-118 (build_edge_index) and src/helper.py:420-433 (dedupe) in the reference.
//
// All of this is HBM-bound integer work: per radix pass every key/value is read twice (histogram,
// scatter) and written once; the CSR build needs ceil(2*ceil(log2 N)/8) passes over 12 B/edge.
#include <stdarg.h>

#include "common.cuh"

namespace pangnn {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// exclusive scan (u32), three-phase, recursive over block sums
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;   // 2048

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *smem /*[9]*/,
                                                         uint32_t &block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t s = lane < (kScanThreads / 32) ? smem[lane] : 0u;
        uint32_t si = s;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += n;
        }
        if (lane < 8) smem[lane] = si - s;          // exclusive warp offsets
        if (lane == 7) smem[8] = si;                // block total
    }
    __syncthreads();
    block_total = smem[8];
    return smem[warp] + inc - v;
}

__global__ void __launch_bounds__(kScanThreads)
scan_block_sums_kernel(const uint32_t *__restrict__ in, int64_t n, uint32_t *__restrict__ sums) {
    __shared__ uint32_t red[8];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        int64_t idx = base + (int64_t)i * kScanThreads + threadIdx.x;
        if (idx < n) s += in[idx];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += red[w];
        sums[blockIdx.x] = t;
    }
}

// Each thread owns kScanItems CONSECUTIVE items.  offsets == nullptr -> single-block mode.
__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const uint32_t *in, uint32_t *out, int64_t n, const uint32_t *__restrict__ offsets,
                  uint32_t *__restrict__ total) {
    __shared__ uint32_t smem[9];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t tsum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0u;
        tsum += v[i];
    }
    uint32_t block_total;
    uint32_t run = block_exclusive_scan(tsum, smem, block_total);
    run += offsets ? offsets[blockIdx.x] : 0u;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
    if (total && !offsets && threadIdx.x == 0) *total = block_total;
}

static size_t scan_ws_elems(int64_t n) {
    size_t elems = 0;
    int64_t m = n;
    while (m > kScanTile) {
        m = (m + kScanTile - 1) / kScanTile;
        elems += align_up((size_t)m, 64);
    }
    return elems + 64;
}

static int scan_rec(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total, uint32_t *ws,
                    cudaStream_t st) {
    const int64_t nb = (n + kScanTile - 1) / kScanTile;
    if (nb <= 1) {
        scan_apply_kernel<<<1, kScanThreads, 0, st>>>(in, out, n, nullptr, total);
        PANGNN_CHECK_LAUNCH("scan_apply(single)");
        return PANGNN_OK;
    }
    uint32_t *sums = ws;
    scan_block_sums_kernel<<<(unsigned)nb, kScanThreads, 0, st>>>(in, n, sums);
    PANGNN_CHECK_LAUNCH("scan_block_sums");
    int rc = scan_rec(sums, sums, nb, total, ws + align_up((size_t)nb, 64), st);
    if (rc) return rc;
    scan_apply_kernel<<<(unsigned)nb, kScanThreads, 0, st>>>(in, out, n, sums, nullptr);
    PANGNN_CHECK_LAUNCH("scan_apply");
    return PANGNN_OK;
}

int exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total, void *ws,
                       size_t ws_bytes, cudaStream_t st) {
    if (ws_bytes < scan_ws_elems(n) * sizeof(uint32_t)) {
        set_error("exclusive_scan: workspace too small");
        return PANGNN_EWORKSPACE;
    }
    if (n <= 0) {
        if (total) return check_cuda(cudaMemsetAsync(total, 0, sizeof(uint32_t), st), "memset");
        return PANGNN_OK;
    }
    return scan_rec(in, out, n, total, static_cast<uint32_t *>(ws), st);
}

// ------------------------------------------------------------------------------------------------
// LSD radix sort, 8-bit digits, (u64 key, u32 value) pairs, stable
// ------------------------------------------------------------------------------------------------
constexpr int kSortWarps = 8;
constexpr int kSortThreads = kSortWarps * 32;
constexpr int kSortItems = 16;                                  // keys per lane
constexpr int kSortWarpChunk = 32 * kSortItems;                 // 512 consecutive keys per warp
constexpr int kSortTile = kSortWarps * kSortWarpChunk;          // 4096 keys per block
constexpr int kRadix = 256;

__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int shift, uint32_t mask,
                  uint32_t *__restrict__ hist /*[256][nb]*/, uint32_t nb) {
    __shared__ uint32_t h[kRadix];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll 4
    for (int i = 0; i < kSortItems; ++i) {
        int64_t idx = base + (int64_t)i * kSortThreads + threadIdx.x;
        if (idx < n) atomicAdd(&h[(uint32_t)(keys[idx] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}

// One pass: every block ranks its 4096 keys (warp-private 512-key chunks, __match_any ballots), places
// them at their BLOCK-LOCAL sorted position in shared memory, and then writes the tile out digit run by
// digit run: consecutive threads write consecutive addresses (a run holds 16 keys = 128 B on average),
// where a direct scatter from registers wrote 1-2 key fragments (measured 1.6 TB/s per pass).
constexpr size_t kScatterSmem = (size_t)kSortTile * (sizeof(uint64_t) + sizeof(uint32_t)) +
                                (size_t)(kSortWarps + 2) * kRadix * sizeof(uint32_t);

__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                     uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n,
                     int shift, uint32_t mask, const uint32_t *__restrict__ offs /*[256][nb]*/,
                     uint32_t nb) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    uint64_t *skey = reinterpret_cast<uint64_t *>(sm_raw);                       // [4096] tile in sorted order
    uint32_t *sval = reinterpret_cast<uint32_t *>(skey + kSortTile);             // [4096]
    uint32_t (*wh)[kRadix] = reinterpret_cast<uint32_t (*)[kRadix]>(sval + kSortTile);   // [warps][256]
    uint32_t *lstart = &wh[kSortWarps][0];                                       // [256] local start of each digit run
    uint32_t *gbase = lstart + kRadix;                                           // [256] global start of each digit run
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&wh[0][0])[i] = 0;
    __syncthreads();

    // phase A: load this warp's 512 consecutive keys (32 per round, coalesced) + per-warp histogram
    const int64_t wbase = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * kSortWarpChunk;
    uint64_t k[kSortItems];
    uint32_t v[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int64_t idx = wbase + r * 32 + lane;
        const bool ok = idx < n;
        k[r] = ok ? keys_in[idx] : 0ull;
        v[r] = ok ? (vals_in ? vals_in[idx] : (uint32_t)idx) : 0u;
        if (ok) atomicAdd(&wh[warp][(uint32_t)(k[r] >> shift) & mask], 1u);
    }
    __syncthreads();

    // phase B: digit d (one thread each): block count -> block-local exclusive prefix over digits (block scan),
    // then per-warp local start = digit start + exclusive prefix over warps
    {
        const int d = threadIdx.x;
        uint32_t cnt = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) cnt += wh[w][d];
        // exclusive scan of cnt over the 256 digits
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        __shared__ uint32_t wsum[kSortWarps];
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) woff += (w < warp) ? wsum[w] : 0u;
        uint32_t run = woff + inc - cnt;                     // local start of digit d
        lstart[d] = run;
        gbase[d] = offs[(size_t)d * nb + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t c = wh[w][d];
            wh[w][d] = run;
            run += c;
        }
    }
    __syncthreads();

    // phase C: stable rank inside the warp, round by round -> block-local sorted position in shared memory
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int64_t idx = wbase + r * 32 + lane;
        const bool ok = idx < n;
        const uint32_t d = ok ? ((uint32_t)(k[r] >> shift) & mask) : (uint32_t)kRadix;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t rank = __popc(peers & lt);
        uint32_t base = 0;
        if (ok) base = wh[warp][d];
        __syncwarp();
        if (ok && rank == 0) wh[warp][d] = base + __popc(peers);
        __syncwarp();
        if (ok) {
            skey[base + rank] = k[r];
            sval[base + rank] = v[r];
        }
    }
    __syncthreads();

    // phase D: coalesced write-out, run by run
    const int64_t tile_n = min((int64_t)kSortTile, n - (int64_t)blockIdx.x * kSortTile);
    for (int i = threadIdx.x; i < tile_n; i += kSortThreads) {
        const uint64_t key = skey[i];
        const uint32_t d = (uint32_t)(key >> shift) & mask;
        const uint32_t g = gbase[d] + ((uint32_t)i - lstart[d]);
        keys_out[g] = key;
        vals_out[g] = sval[i];
    }
}

static size_t sort_ws_bytes(int64_t n) {
    const size_t nb = (size_t)((n + kSortTile - 1) / kSortTile);
    size_t b = 0;
    b += align_up((size_t)n * sizeof(uint64_t), 256);            // tmp keys
    b += align_up((size_t)n * sizeof(uint32_t), 256);            // tmp vals
    b += align_up(kRadix * nb * sizeof(uint32_t), 256);          // histogram / offsets
    b += align_up(scan_ws_elems((int64_t)(kRadix * nb)) * sizeof(uint32_t), 256);
    return b + 1024;
}

// Sorts by the key bits [first_bit, key_bits) only (stable): first_bit > 0 is the CSR transpose, whose input
// is already ordered by the low bits.
int sort_pairs_u64(const uint64_t *keys_in, const uint32_t *vals_in, uint64_t *keys_out,
                   uint32_t *vals_out, int64_t n, int key_bits, void *ws, size_t ws_bytes,
                   cudaStream_t st, int first_bit = 0) {
    if (n >= ((int64_t)1 << 32)) {
        set_error("sort_pairs: n must be < 2^32");
        return PANGNN_EINVAL;
    }
    if (ws_bytes < sort_ws_bytes(n)) {
        set_error("sort_pairs: workspace too small (%zu < %zu)", ws_bytes, sort_ws_bytes(n));
        return PANGNN_EWORKSPACE;
    }
    if (n == 0) return PANGNN_OK;
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 64) key_bits = 64;
    const uint32_t nb = (uint32_t)((n + kSortTile - 1) / kSortTile);
    Workspace w(ws, ws_bytes);
    uint64_t *tk = w.take<uint64_t>(n);
    uint32_t *tv = w.take<uint32_t>(n);
    uint32_t *hist = w.take<uint32_t>((size_t)kRadix * nb);
    const size_t scan_elems = scan_ws_elems((int64_t)kRadix * nb);
    uint32_t *scan_ws = w.take<uint32_t>(scan_elems);

    static bool attr_set = false;
    if (!attr_set) {
        int rc = check_cuda(cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)kScatterSmem), "cudaFuncSetAttribute(radix_scatter)");
        if (rc) return rc;
        attr_set = true;
    }
    if (first_bit < 0 || first_bit >= key_bits) first_bit = 0;
    const int passes = (key_bits - first_bit + 7) / 8;
    const uint64_t *ksrc = keys_in;
    const uint32_t *vsrc = vals_in;
    for (int p = 0; p < passes; ++p) {
        // the LAST pass must land in (keys_out, vals_out)
        const bool to_out = ((passes - 1 - p) % 2) == 0;
        uint64_t *kdst = to_out ? keys_out : tk;
        uint32_t *vdst = to_out ? vals_out : tv;
        const int shift = first_bit + 8 * p;
        const int width = (key_bits - shift) < 8 ? (key_bits - shift) : 8;
        const uint32_t mask = (1u << width) - 1u;
        radix_hist_kernel<<<nb, kSortThreads, 0, st>>>(ksrc, n, shift, mask, hist, nb);
        PANGNN_CHECK_LAUNCH("radix_hist");
        int rc = exclusive_scan_u32(hist, hist, (int64_t)kRadix * nb, nullptr, scan_ws,
                                    scan_elems * sizeof(uint32_t), st);
        if (rc) return rc;
        radix_scatter_kernel<<<nb, kSortThreads, kScatterSmem, st>>>(ksrc, vsrc, kdst, vdst, n, shift, mask,
                                                                     hist, nb);
        PANGNN_CHECK_LAUNCH("radix_scatter");
        ksrc = kdst;
        vsrc = vdst;
    }
    return PANGNN_OK;
}

// ------------------------------------------------------------------------------------------------
// CSR build
// ------------------------------------------------------------------------------------------------
__global__ void csr_pack_keys_kernel(const int64_t *__restrict__ edge_index, int64_t E, int nbits,
                                     int by_dst, uint64_t *__restrict__ keys) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const uint64_t s = (uint64_t)edge_index[e], d = (uint64_t)edge_index[E + e];
    keys[e] = by_dst ? ((d << nbits) | s) : ((s << nbits) | d);
}

__global__ void csr_unpack_kernel(const uint64_t *__restrict__ keys, int64_t E, int32_t N, int nbits,
                                  int64_t *__restrict__ rowptr, int32_t *__restrict__ col) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    const uint64_t key = keys[i];
    const int64_t row = (int64_t)(key >> nbits);
    col[i] = (int32_t)(key & ((1ull << nbits) - 1ull));
    const int64_t prev = (i == 0) ? -1 : (int64_t)(keys[i - 1] >> nbits);
    for (int64_t r = prev + 1; r <= row; ++r) rowptr[r] = i;       // rows (prev, row] start at i
    if (i == E - 1)
        for (int64_t r = row + 1; r <= N; ++r) rowptr[r] = E;
}

// CSR transpose: entry p of row r (binary search over rowptr) becomes key (col << nbits | r) carrying its
// perm; the entries are already ordered by (r, col, perm), so a stable sort on the HIGH nbits alone yields
// the (col, r, perm) order of the other orientation: ceil(nbits / 8) radix passes instead of ceil(2 nbits / 8).
__global__ void csr_transpose_pack_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                          int64_t E, int32_t N, int nbits, uint64_t *__restrict__ keys) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= E) return;
    int64_t a = 0, b = N;                                     // last row r with rowptr[r] <= p
    while (b - a > 1) {
        const int64_t m = (a + b) >> 1;
        if (__ldg(rowptr + m) <= p) a = m; else b = m;
    }
    keys[p] = ((uint64_t)(uint32_t)col[p] << nbits) | (uint64_t)a;
}

// Edge lists that arrive already in canonical (src, dst) order — every table this package's preprocessing emits —
// need no sort for the by-source orientation: one pass checks the order, one pass writes rowptr / col / perm
// (identity); the by-destination orientation then comes from the 3-pass transpose.
// flags: bit 0 = not in (src, dst) order, bit 1 = an endpoint outside [0, N) (torch index ops would raise; the
// packed (row << b | col) keys and the rowptr writes would silently go out of bounds)
__global__ void edges_sorted_kernel(const int64_t *__restrict__ edge_index, int64_t E, int64_t N,
                                    int32_t *__restrict__ flags) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int64_t s0 = edge_index[e], d0 = edge_index[E + e];
    int f = (s0 < 0 || s0 >= N || d0 < 0 || d0 >= N) ? 2 : 0;
    if (e + 1 < E) {
        const int64_t s1 = edge_index[e + 1];
        if (s0 > s1 || (s0 == s1 && d0 > edge_index[E + e + 1])) f |= 1;
    }
    if (f) atomicOr(flags, f);
}

// One thread per edge (col, perm) and per row (rowptr[r] = first position whose source is >= r, a binary search
// over the sorted source row: no serial gap filling, whatever the distribution of empty rows).
__global__ void csr_from_sorted_kernel(const int64_t *__restrict__ edge_index, int64_t E, int32_t N,
                                       int64_t *__restrict__ rowptr, int32_t *__restrict__ col,
                                       uint32_t *__restrict__ perm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < E) {
        col[i] = (int32_t)edge_index[E + i];
        perm[i] = (uint32_t)i;
    }
    if (i <= N) {
        int64_t a = 0, b = E;
        while (a < b) {
            const int64_t m = (a + b) >> 1;
            if (__ldg(edge_index + m) < i) a = m + 1; else b = m;
        }
        rowptr[i] = a;
    }
}

// Small graphs (the reference's actual training regime, SURVEY F7: batches of 32 sub-graphs, a few hundred
// edges): the whole CSR build in ONE single-CTA launch instead of ~35 (pack, 5 radix passes x 3 kernels,
// unpack) — those steps are launch-latency-bound.  Bitonic sort of (key, original position) pairs in shared
// memory (the position breaks ties, so the result equals the stable radix sort), then col / perm / rowptr.
constexpr int kSmallCsrMaxE = 4096;

__global__ void __launch_bounds__(1024)
csr_build_small_kernel(const int64_t *__restrict__ edge_index, int32_t E, int32_t N, int nbits, int by_dst,
                       int64_t *__restrict__ rowptr, int32_t *__restrict__ col, uint32_t *__restrict__ perm) {
    __shared__ uint64_t skey[kSmallCsrMaxE];
    __shared__ uint32_t sidx[kSmallCsrMaxE];
    int P = 1;
    while (P < E) P <<= 1;                                    // padded length (power of two)
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        uint64_t k = ~0ull;                                   // padding sorts last
        if (i < E) {
            const uint64_t sv = (uint64_t)edge_index[i], dv = (uint64_t)edge_index[E + i];
            k = by_dst ? ((dv << nbits) | sv) : ((sv << nbits) | dv);
        }
        skey[i] = k;
        sidx[i] = (uint32_t)i;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));    // index of the lower element of the pair
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;             // ascending half
                const uint64_t ka = skey[lo], kb = skey[hi];
                const uint32_t ia = sidx[lo], ib = sidx[hi];
                const bool gt = ka > kb || (ka == kb && ia > ib);
                if (gt == up) {
                    skey[lo] = kb; skey[hi] = ka;
                    sidx[lo] = ib; sidx[hi] = ia;
                }
            }
            __syncthreads();
        }
    }
    const uint64_t mask = (1ull << nbits) - 1ull;
    for (int i = threadIdx.x; i < E; i += blockDim.x) {
        col[i] = (int32_t)(skey[i] & mask);
        perm[i] = sidx[i];
    }
    // rowptr[r] = first sorted position whose row is >= r
    for (int r = threadIdx.x; r <= N; r += blockDim.x) {
        int a = 0, b = E;
        while (a < b) {
            const int m = (a + b) >> 1;
            if ((int64_t)(skey[m] >> nbits) < (int64_t)r) a = m + 1; else b = m;
        }
        rowptr[r] = a;
    }
}

__global__ void fill_i64_kernel(int64_t *p, int64_t n, int64_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

static int bits_for(int64_t n) {
    int b = 1;
    while (((int64_t)1 << b) < n) ++b;
    return b;
}

}  // namespace pangnn

using namespace pangnn;

extern "C" {

int pangnn_abi_version(void) { return PANGNN_ABI_VERSION; }
const char *pangnn_last_error(void) { return g_err; }
#ifndef PANGNN_SRC_DIGEST
#define PANGNN_SRC_DIGEST "unknown"
#endif
const char *pangnn_source_digest(void) { return PANGNN_SRC_DIGEST; }

size_t pangnn_scan_workspace_bytes(int64_t n) { return scan_ws_elems(n) * sizeof(uint32_t); }

int pangnn_exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total, void *ws,
                              size_t ws_bytes, void *stream) {
    PANGNN_REQUIRE(n == 0 || (in && out), "null pointer");
    return exclusive_scan_u32(in, out, n, total, ws, ws_bytes, (cudaStream_t)stream);
}

size_t pangnn_sort_pairs_workspace_bytes(int64_t n) { return sort_ws_bytes(n); }

int pangnn_sort_pairs_u64(const uint64_t *keys_in, const uint32_t *vals_in, uint64_t *keys_out,
                          uint32_t *vals_out, int64_t n, int key_bits, void *ws, size_t ws_bytes,
                          void *stream) {
    PANGNN_REQUIRE(n == 0 || (keys_in && keys_out && vals_out && ws), "null pointer");
    PANGNN_REQUIRE(keys_in != keys_out, "in-place sort is not supported");
    return sort_pairs_u64(keys_in, vals_in, keys_out, vals_out, n, key_bits, ws, ws_bytes,
                          (cudaStream_t)stream);
}

size_t pangnn_csr_build_workspace_bytes(int64_t E) {
    return 2 * align_up((size_t)E * sizeof(uint64_t), 256) + sort_ws_bytes(E) + 1024;
}

int pangnn_csr_build(const int64_t *edge_index, int64_t E, int32_t N, int by_dst, int64_t *rowptr,
                     int32_t *col, uint32_t *perm, void *ws, size_t ws_bytes, void *stream) {
    PANGNN_REQUIRE(rowptr && N >= 0 && E >= 0, "bad arguments");
    PANGNN_REQUIRE(E == 0 || (edge_index && col && perm && ws), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (E == 0) {
        fill_i64_kernel<<<(unsigned)((N + 1 + 255) / 256), 256, 0, st>>>(rowptr, (int64_t)N + 1, 0);
        PANGNN_CHECK_LAUNCH("fill_rowptr");
        return PANGNN_OK;
    }
    if (E <= kSmallCsrMaxE) {                                 // one launch for small graphs
        csr_build_small_kernel<<<1, 1024, 0, st>>>(edge_index, (int32_t)E, N, bits_for(N > 1 ? N : 2), by_dst, rowptr, col, perm);
        PANGNN_CHECK_LAUNCH("csr_build_small");
        return PANGNN_OK;
    }
    if (ws_bytes < pangnn_csr_build_workspace_bytes(E)) {
        set_error("csr_build: workspace too small");
        return PANGNN_EWORKSPACE;
    }
    Workspace w(ws, ws_bytes);
    uint64_t *ka = w.take<uint64_t>(E);
    uint64_t *kb = w.take<uint64_t>(E);
    w.off = align_up(w.off, 256);
    void *sort_ws = w.base + w.off;
    const size_t sort_bytes = ws_bytes - w.off;
    const int nbits = bits_for(N > 1 ? N : 2);
    const unsigned blocks = (unsigned)((E + 255) / 256);
    csr_pack_keys_kernel<<<blocks, 256, 0, st>>>(edge_index, E, nbits, by_dst, ka);
    PANGNN_CHECK_LAUNCH("csr_pack_keys");
    int rc = sort_pairs_u64(ka, nullptr, kb, perm, E, 2 * nbits, sort_ws, sort_bytes, st);
    if (rc) return rc;
    csr_unpack_kernel<<<blocks, 256, 0, st>>>(kb, E, N, nbits, rowptr, col);
    PANGNN_CHECK_LAUNCH("csr_unpack");
    return PANGNN_OK;
}

/* *unsorted (device int32, must be zeroed by the caller... it is set here) = 1 unless the edge list is in
 * non-decreasing (src, dst) order. */
int pangnn_edges_sorted(const int64_t *edge_index, int64_t E, int32_t N, int32_t *flags, void *stream) {
    PANGNN_REQUIRE(flags && E >= 0 && N >= 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_cuda(cudaMemsetAsync(flags, 0, sizeof(int32_t), st), "memset");
    if (rc || E < 1) return rc;
    PANGNN_REQUIRE(edge_index, "null pointer");
    edges_sorted_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(edge_index, E, (int64_t)N, flags);
    PANGNN_CHECK_LAUNCH("edges_sorted");
    return PANGNN_OK;
}

/* By-source CSR of an edge list that IS in (src, dst) order (pangnn_edges_sorted): no sort, perm = identity;
 * identical to pangnn_csr_build(..., by_dst = 0). */
int pangnn_csr_from_sorted(const int64_t *edge_index, int64_t E, int32_t N, int64_t *rowptr, int32_t *col,
                           uint32_t *perm, void *stream) {
    PANGNN_REQUIRE(rowptr && N >= 0 && E >= 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (E == 0) {
        fill_i64_kernel<<<(unsigned)((N + 1 + 255) / 256), 256, 0, st>>>(rowptr, (int64_t)N + 1, 0);
        PANGNN_CHECK_LAUNCH("fill_rowptr");
        return PANGNN_OK;
    }
    PANGNN_REQUIRE(edge_index && col && perm, "null pointer");
    const int64_t work = E > (int64_t)N + 1 ? E : (int64_t)N + 1;
    csr_from_sorted_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(edge_index, E, N, rowptr, col, perm);
    PANGNN_CHECK_LAUNCH("csr_from_sorted");
    return PANGNN_OK;
}

/* CSR of the transposed graph from an existing CSR (rows <-> columns), same canonical order and the same
 * perm (csr position -> original edge position) as pangnn_csr_build of the other orientation. */
int pangnn_csr_transpose(const int64_t *rowptr, const int32_t *col, const uint32_t *perm, int64_t E, int32_t N,
                         int64_t *rowptr_t, int32_t *col_t, uint32_t *perm_t, void *ws, size_t ws_bytes,
                         void *stream) {
    PANGNN_REQUIRE(rowptr_t && N >= 0 && E >= 0, "bad arguments");
    PANGNN_REQUIRE(E == 0 || (rowptr && col && perm && col_t && perm_t && ws), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (E == 0) {
        fill_i64_kernel<<<(unsigned)((N + 1 + 255) / 256), 256, 0, st>>>(rowptr_t, (int64_t)N + 1, 0);
        PANGNN_CHECK_LAUNCH("fill_rowptr");
        return PANGNN_OK;
    }
    if (ws_bytes < pangnn_csr_build_workspace_bytes(E)) {
        set_error("csr_transpose: workspace too small");
        return PANGNN_EWORKSPACE;
    }
    Workspace w(ws, ws_bytes);
    uint64_t *ka = w.take<uint64_t>(E);
    uint64_t *kb = w.take<uint64_t>(E);
    w.off = align_up(w.off, 256);
    void *sort_ws = w.base + w.off;
    const size_t sort_bytes = ws_bytes - w.off;
    const int nbits = bits_for(N > 1 ? N : 2);
    const unsigned blocks = (unsigned)((E + 255) / 256);
    csr_transpose_pack_kernel<<<blocks, 256, 0, st>>>(rowptr, col, E, N, nbits, ka);
    PANGNN_CHECK_LAUNCH("csr_transpose_pack");
    int rc = sort_pairs_u64(ka, perm, kb, perm_t, E, 2 * nbits, sort_ws, sort_bytes, st, nbits);
    if (rc) return rc;
    csr_unpack_kernel<<<blocks, 256, 0, st>>>(kb, E, N, nbits, rowptr_t, col_t);
    PANGNN_CHECK_LAUNCH("csr_unpack");
    return PANGNN_OK;
}

}  // extern "C"
