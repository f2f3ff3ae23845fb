// MMseqs2 hit table (16-column TSV; replaces pandas read_csv + the per-row dict lookups of
// src/preprocessing.py:388-426) parsed on the device: the file's bytes go to HBM once, then
//   1. line index: newlines counted per 4096-byte tile, scanned, line starts written;
//   2. one thread per line: fields split at tabs, query / target ids hashed (FNV-1a 64) and looked up in the
//      sorted hash table of the known gene ids (binary search -> node id, -1 = unknown), last column parsed as a
//      decimal number (sign, digits, fraction, exponent; exact for <= 19 significant digits and |exp10| <= 22,
//      which covers MMseqs2 bit scores).
// Lines that are empty or start with '#' yield q = -2 (dropped by the caller).  Byte work, HBM/L2 bound.
#include "common.cuh"

namespace pangnn {

int exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total, void *ws, size_t ws_bytes,
                       cudaStream_t st);

constexpr int kTsvTile = 4096;

__global__ void __launch_bounds__(256)
tsv_count_newlines_kernel(const uint8_t *__restrict__ text, int64_t n, uint32_t *__restrict__ counts) {
    __shared__ uint32_t red[8];
    const int64_t base = (int64_t)blockIdx.x * kTsvTile;
    uint32_t c = 0;
    for (int i = threadIdx.x; i < kTsvTile; i += 256) {
        const int64_t p = base + i;
        if (p < n && text[p] == '\n') ++c;
    }
    c = (uint32_t)warp_sum((float)c);                        // <= 128 per warp: exact in fp32
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int w = 0; w < 8; ++w) s += red[w];
        counts[blockIdx.x] = s;
    }
}

// line_start[0] = 0; line_start[k] = position after the k-th newline.  One warp per 32-byte-groups walk would be
// faster; a tile's newlines are few (~40), so one thread per tile writes them in order.
__global__ void __launch_bounds__(256)
tsv_line_starts_kernel(const uint8_t *__restrict__ text, int64_t n, const uint32_t *__restrict__ offs,
                       int64_t *__restrict__ line_start, int64_t num_tiles) {
    const int64_t tile = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= num_tiles) return;
    if (tile == 0) line_start[0] = 0;
    int64_t k = (int64_t)offs[tile] + 1;
    const int64_t base = tile * kTsvTile;
    const int64_t end = min(base + (int64_t)kTsvTile, n);
    for (int64_t p = base; p < end; ++p)
        if (text[p] == '\n') line_start[k++] = p + 1;
}

__device__ __forceinline__ int32_t tsv_lookup(const uint64_t *__restrict__ table, const int32_t *__restrict__ pos,
                                              int32_t n, uint64_t h) {
    int32_t a = 0, b = n;
    while (a < b) {
        const int32_t m = (a + b) >> 1;
        if (table[m] < h) a = m + 1; else b = m;
    }
    return (a < n && table[a] == h) ? pos[a] : -1;
}

__constant__ double kPow10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                  1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

__device__ double tsv_parse_number(const uint8_t *s, int64_t a, int64_t b) {
    bool neg = false;
    if (a < b && (s[a] == '-' || s[a] == '+')) neg = s[a++] == '-';
    uint64_t mant = 0;
    int digits = 0, exp10 = 0;
    bool frac = false, any = false;
    for (; a < b; ++a) {
        const uint8_t ch = s[a];
        if (ch >= '0' && ch <= '9') {
            any = true;
            if (digits < 19) { mant = mant * 10 + (ch - '0'); digits += (mant != 0); if (frac) --exp10; }
            else if (!frac) ++exp10;                          // digits beyond 19: dropped (scaled)
        } else if (ch == '.' && !frac) {
            frac = true;
        } else {
            break;
        }
    }
    if (a < b && (s[a] == 'e' || s[a] == 'E')) {
        ++a;
        bool eneg = false;
        if (a < b && (s[a] == '-' || s[a] == '+')) eneg = s[a++] == '-';
        int e = 0;
        for (; a < b && s[a] >= '0' && s[a] <= '9'; ++a) e = e < 10000 ? e * 10 + (s[a] - '0') : e;
        exp10 += eneg ? -e : e;
    }
    if (!any) return __longlong_as_double(0x7ff8000000000000LL);   // NaN
    double v = (double)mant;
    if (exp10 > 0) v = exp10 <= 22 ? v * kPow10[exp10] : v * pow(10.0, (double)exp10);
    else if (exp10 < 0) v = -exp10 <= 22 ? v / kPow10[-exp10] : v * pow(10.0, (double)exp10);
    return neg ? -v : v;
}

__global__ void __launch_bounds__(256)
tsv_parse_lines_kernel(const uint8_t *__restrict__ text, int64_t n, const int64_t *__restrict__ line_start,
                       int64_t num_lines, int64_t num_newlines, int32_t score_col, const uint64_t *__restrict__ table,
                       const int32_t *__restrict__ pos, int32_t table_n, int32_t *__restrict__ q,
                       int32_t *__restrict__ t, double *__restrict__ bits) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= num_lines) return;
    int64_t a = line_start[l];
    int64_t b = l < num_newlines ? line_start[l + 1] - 1 : n;          // exclusive end: the line's newline, or EOF
    if (b > a && text[b - 1] == '\r') --b;
    if (b <= a || text[a] == '#') {
        q[l] = -2; t[l] = -2; bits[l] = 0.0;
        return;
    }
    int field = 0;
    int64_t fs = a;
    int32_t qi = -1, ti = -1;
    double sc = __longlong_as_double(0x7ff8000000000000LL);
    uint64_t h = 0xcbf29ce484222325ull;                       // FNV-1a offset basis
    for (int64_t p = a; p <= b; ++p) {
        const bool sep = (p == b) || text[p] == '\t';
        if (!sep) {
            if (field < 2) h = (h ^ (uint64_t)text[p]) * 0x100000001b3ull;
            continue;
        }
        if (field == 0) qi = tsv_lookup(table, pos, table_n, h);
        else if (field == 1) ti = tsv_lookup(table, pos, table_n, h);
        if (field == score_col) sc = tsv_parse_number(text, fs, p);
        ++field;
        fs = p + 1;
        h = 0xcbf29ce484222325ull;
        if (field > score_col) break;
    }
    q[l] = qi; t[l] = ti; bits[l] = sc;
}

}  // namespace pangnn

using namespace pangnn;

extern "C" {

size_t pangnn_parse_hits_tsv_workspace_bytes(int64_t num_bytes) {
    const int64_t tiles = (num_bytes + kTsvTile - 1) / kTsvTile;
    return 2 * align_up((size_t)(tiles + 1) * sizeof(uint32_t), 256) + pangnn_scan_workspace_bytes(tiles) + 1024;
}

/* Pass 1: line index.  num_lines (device) = number of newlines, + 1 if the text does not end with one. */
int pangnn_tsv_line_index(const uint8_t *text, int64_t num_bytes, int64_t *line_start, int64_t max_lines,
                          uint32_t *num_newlines /* device */, void *ws, size_t ws_bytes, void *stream) {
    PANGNN_REQUIRE(num_bytes >= 0 && num_newlines && ws, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (num_bytes == 0) return check_cuda(cudaMemsetAsync(num_newlines, 0, sizeof(uint32_t), st), "memset");
    PANGNN_REQUIRE(text, "null pointer");
    if (ws_bytes < pangnn_parse_hits_tsv_workspace_bytes(num_bytes)) {
        set_error("tsv_line_index: workspace too small");
        return PANGNN_EWORKSPACE;
    }
    const int64_t tiles = (num_bytes + kTsvTile - 1) / kTsvTile;
    Workspace w(ws, ws_bytes);
    uint32_t *counts = w.take<uint32_t>(tiles + 1);
    w.off = align_up(w.off, 256);
    uint32_t *offs = w.take<uint32_t>(tiles + 1);
    w.off = align_up(w.off, 256);
    void *scan_ws = w.base + w.off;
    tsv_count_newlines_kernel<<<(unsigned)tiles, 256, 0, st>>>(text, num_bytes, counts);
    PANGNN_CHECK_LAUNCH("tsv_count_newlines");
    int rc = exclusive_scan_u32(counts, offs, tiles, num_newlines, scan_ws, ws_bytes - w.off, st);
    if (rc) return rc;
    if (line_start) {                                          // second call, once the caller knows the count
        PANGNN_REQUIRE(max_lines >= 1, "max_lines");
        tsv_line_starts_kernel<<<(unsigned)((tiles + 255) / 256), 256, 0, st>>>(text, num_bytes, offs, line_start, tiles);
        PANGNN_CHECK_LAUNCH("tsv_line_starts");
    }
    return PANGNN_OK;
}

/* Pass 2: q / t = node ids of columns 0 / 1 through the sorted FNV-1a-64 table (-1 unknown, -2 blank / comment
 * line), bits = column `score_col` as a double. */
int pangnn_tsv_parse_hits(const uint8_t *text, int64_t num_bytes, const int64_t *line_start, int64_t num_lines,
                          int64_t num_newlines, int32_t score_col, const uint64_t *id_hash_sorted, const int32_t *id_pos, int32_t num_ids,
                          int32_t *q, int32_t *t, double *bits, void *stream) {
    if (num_lines <= 0) return PANGNN_OK;
    PANGNN_REQUIRE(text && line_start && q && t && bits && score_col >= 2, "bad arguments");
    PANGNN_REQUIRE(num_lines == num_newlines || num_lines == num_newlines + 1, "num_lines must be num_newlines (+ 1)");
    PANGNN_REQUIRE(num_ids == 0 || (id_hash_sorted && id_pos), "null id table");
    tsv_parse_lines_kernel<<<(unsigned)((num_lines + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        text, num_bytes, line_start, num_lines, num_newlines, score_col, id_hash_sorted, id_pos, num_ids, q, t, bits);
    PANGNN_CHECK_LAUNCH("tsv_parse_lines");
    return PANGNN_OK;
}

}  // extern "C"
