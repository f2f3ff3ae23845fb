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

// ---- GFF3 annotation and RIBAP group table (src/preprocessing.py:329-367, 159-193) ----------
// Both are tab-separated text read by pandas with comment = '#': everything from the first '#' of a line on is
// ignored, lines that are empty after that are no records.  A field is "missing" when it is empty or one of
// pandas' default NA spellings.
__device__ bool tsv_is_na(const uint8_t *s, int64_t a, int64_t b) {
    const int n = (int)(b - a);
    if (n == 0) return true;
    if (n > 8) return false;
    char f[9];
    for (int i = 0; i < n; ++i) f[i] = (char)s[a + i];
    f[n] = 0;
    const char *na[] = {"NA", "N/A", "NULL", "NaN", "nan", "n/a", "null", "#N/A", "#NA", "-NaN", "-nan", "1.#IND",
                        "1.#QNAN", "<NA>", "#N/A N/A", "None"};
    for (const char *w : na) {
        int i = 0;
        while (w[i] && w[i] == f[i]) ++i;
        if (w[i] == 0 && f[i] == 0) return true;
    }
    return false;
}

// [a, b) of line l with the comment tail and a trailing '\r' removed
__device__ void tsv_line_range(const uint8_t *text, int64_t n, const int64_t *line_start, int64_t l, int64_t num_newlines,
                               int64_t &a, int64_t &b) {
    a = line_start[l];
    b = l < num_newlines ? line_start[l + 1] - 1 : n;
    for (int64_t p = a; p < b; ++p)
        if (text[p] == '#') { b = p; break; }
    if (b > a && text[b - 1] == '\r') --b;
}

constexpr int kGffRecord = 1, kGffComplete = 2, kGffStart = 4, kGffGeneId = 8, kGffComplex = 16;

// One thread per line of a GFF3 file.  flags: record (non-empty after comment removal) | complete (9 fields, none
// missing: survives dropna) | the attribute column mentions `start_gene` | the id looks like [A-Z]+_[0-9]+ |
// complex ("ID=" occurs elsewhere than at the start of the attribute: the caller resolves that line on the host).
// id = attribute up to the first ';' without the leading "ID=": byte range + FNV-1a 64 hash.
__global__ void __launch_bounds__(256)
gff_parse_lines_kernel(const uint8_t *__restrict__ text, int64_t n, const int64_t *__restrict__ line_start,
                       int64_t num_lines, int64_t num_newlines, const uint8_t *__restrict__ start_gene, int32_t sg_len,
                       int32_t *__restrict__ flags, int64_t *__restrict__ id_off, int32_t *__restrict__ id_len,
                       uint64_t *__restrict__ id_hash) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= num_lines) return;
    int64_t a, b;
    tsv_line_range(text, n, line_start, l, num_newlines, a, b);
    int f = 0;
    int64_t io = 0;
    int32_t il = 0;
    uint64_t h = 0xcbf29ce484222325ull;
    if (b > a) {
        f = kGffRecord;
        int field = 0, missing = 0;
        int64_t fs = a, as = -1, ae = -1;
        for (int64_t p = a; p <= b; ++p) {
            if (p < b && text[p] != '\t') continue;
            if (field < 9 && tsv_is_na(text, fs, p)) ++missing;
            if (field == 8) { as = fs; ae = b; break; }       // the attribute column runs to the end of the line
            ++field;
            fs = p + 1;
        }
        if (as >= 0 && missing == 0) f |= kGffComplete;
        if (as >= 0) {
            // a 10th tab-separated column would end the attribute early
            for (int64_t p = as; p < ae; ++p)
                if (text[p] == '\t') { ae = p; break; }
            for (int64_t p = as; p + sg_len <= ae && sg_len > 0; ++p) {
                int i = 0;
                while (i < sg_len && text[p + i] == start_gene[i]) ++i;
                if (i == sg_len) { f |= kGffStart; break; }
            }
            int64_t e = as;
            while (e < ae && text[e] != ';') ++e;
            int64_t s0 = as;
            if (e - s0 >= 3 && text[s0] == 'I' && text[s0 + 1] == 'D' && text[s0 + 2] == '=') s0 += 3;
            for (int64_t p = s0; p + 3 <= e; ++p)
                if (text[p] == 'I' && text[p + 1] == 'D' && text[p + 2] == '=') f |= kGffComplex;
            io = s0;
            il = (int32_t)(e - s0);
            for (int64_t p = s0; p < e; ++p) {
                h = (h ^ (uint64_t)text[p]) * 0x100000001b3ull;
                if (p + 2 < e && text[p] >= 'A' && text[p] <= 'Z' && text[p + 1] == '_' && text[p + 2] >= '0' && text[p + 2] <= '9')
                    f |= kGffGeneId;
            }
        }
    }
    flags[l] = f;
    id_off[l] = io;
    id_len[l] = il;
    id_hash[l] = h;
}

// One thread per line of a tab-separated table: node id of the gene named in every KEPT column (col_slot[c] = output
// slot of column c or -1), through the sorted hash table: out[l * K + slot] = node id, -1 = unknown id, -2 = missing
// cell / absent column.  line_flag[l] = 1 when the line is a record.
__global__ void __launch_bounds__(256)
tsv_lookup_columns_kernel(const uint8_t *__restrict__ text, int64_t n, const int64_t *__restrict__ line_start,
                          int64_t num_lines, int64_t num_newlines, const int32_t *__restrict__ col_slot, int32_t ncols,
                          int32_t K, const uint64_t *__restrict__ table, const int32_t *__restrict__ pos, int32_t table_n,
                          int32_t *__restrict__ out, int32_t *__restrict__ line_flag) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= num_lines) return;
    int64_t a, b;
    tsv_line_range(text, n, line_start, l, num_newlines, a, b);
    for (int k = 0; k < K; ++k) out[l * K + k] = -2;
    line_flag[l] = b > a ? 1 : 0;
    if (b <= a) return;
    int field = 0;
    int64_t fs = a;
    uint64_t h = 0xcbf29ce484222325ull;
    for (int64_t p = a; p <= b; ++p) {
        if (p < b && text[p] != '\t') {
            h = (h ^ (uint64_t)text[p]) * 0x100000001b3ull;
            continue;
        }
        if (field < ncols) {
            const int32_t slot = col_slot[field];
            if (slot >= 0 && !tsv_is_na(text, fs, p)) out[l * K + slot] = tsv_lookup(table, pos, table_n, h);
        }
        ++field;
        fs = p + 1;
        h = 0xcbf29ce484222325ull;
    }
}

}  // namespace pangnn

using namespace pangnn;

extern "C" {

int pangnn_gff_parse_lines(const uint8_t *text, int64_t num_bytes, const int64_t *line_start, int64_t num_lines,
                           int64_t num_newlines, const uint8_t *start_gene, int32_t start_gene_len, int32_t *flags,
                           int64_t *id_off, int32_t *id_len, uint64_t *id_hash, void *stream) {
    if (num_lines <= 0) return PANGNN_OK;
    PANGNN_REQUIRE(text && line_start && flags && id_off && id_len && id_hash, "null pointer");
    PANGNN_REQUIRE(num_lines == num_newlines || num_lines == num_newlines + 1, "num_lines must be num_newlines (+ 1)");
    PANGNN_REQUIRE(start_gene_len == 0 || start_gene, "null start gene");
    gff_parse_lines_kernel<<<(unsigned)((num_lines + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        text, num_bytes, line_start, num_lines, num_newlines, start_gene, start_gene_len, flags, id_off, id_len, id_hash);
    PANGNN_CHECK_LAUNCH("gff_parse_lines");
    return PANGNN_OK;
}

int pangnn_tsv_lookup_columns(const uint8_t *text, int64_t num_bytes, const int64_t *line_start, int64_t num_lines,
                              int64_t num_newlines, const int32_t *col_slot, int32_t num_cols, int32_t num_slots,
                              const uint64_t *id_hash_sorted, const int32_t *id_pos, int32_t num_ids, int32_t *out,
                              int32_t *line_flag, void *stream) {
    if (num_lines <= 0) return PANGNN_OK;
    PANGNN_REQUIRE(text && line_start && col_slot && out && line_flag && num_cols > 0 && num_slots > 0, "bad arguments");
    PANGNN_REQUIRE(num_lines == num_newlines || num_lines == num_newlines + 1, "num_lines must be num_newlines (+ 1)");
    PANGNN_REQUIRE(num_ids == 0 || (id_hash_sorted && id_pos), "null id table");
    tsv_lookup_columns_kernel<<<(unsigned)((num_lines + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        text, num_bytes, line_start, num_lines, num_newlines, col_slot, num_cols, num_slots, id_hash_sorted, id_pos, num_ids,
        out, line_flag);
    PANGNN_CHECK_LAUNCH("tsv_lookup_columns");
    return PANGNN_OK;
}

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
