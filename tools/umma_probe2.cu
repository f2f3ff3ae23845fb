// Hardware probe (development aid, not part of the library): which shared-memory layouts does
// tcgen05.mma kind::tf32 accept for MN-major operands, and does the 128B_BASE32B swizzle also work
// for K-major descriptors (so that ONE copy of a row-major [e][k] tile can feed a K-major MMA and an
// MN-major MMA)?  Host builds raw byte images under a hypothesis, the kernel runs the MMAs and dumps D.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I pangnn_b200/csrc tools/umma_probe2.cu -o tools/bin/umma_probe2
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <functional>
#include <vector>
#include "umma.cuh"
using namespace pangnn;

__global__ void probe(const uint8_t *imgA, uint32_t bytesA, const uint8_t *imgB, uint32_t bytesB,
                      const uint64_t *descs, int n_mma, uint32_t idesc, int ncols, float *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tb;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sb = umma::smem_u32(smem);
    const uint32_t pad = (1024 - (sb & 1023)) & 1023;          // 1024-byte aligned operand images
    uint8_t *sA = smem + pad, *sB = sA + ((bytesA + 1023) / 1024) * 1024;
    for (uint32_t i = tid; i < bytesA / 4; i += blockDim.x) ((uint32_t *)sA)[i] = ((const uint32_t *)imgA)[i];
    for (uint32_t i = tid; i < bytesB / 4; i += blockDim.x) ((uint32_t *)sB)[i] = ((const uint32_t *)imgB)[i];
    if (warp == 0) umma::tmem_alloc(&tb, 256);
    if (tid == 32) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
    umma::fence_async_smem(); umma::fence_before_sync(); __syncthreads(); umma::fence_after_sync();
    const uint32_t a0 = umma::smem_u32(sA), b0 = umma::smem_u32(sB);
    if (tid == 0) {
        for (int i = 0; i < n_mma; ++i) {
            // descs hold offsets relative to the image base in the address field; add the base (16-byte units)
            const uint64_t da = descs[2 * i] + (uint64_t)((a0 >> 4) & 0x3fff), db = descs[2 * i + 1] + (uint64_t)((b0 >> 4) & 0x3fff);
            umma::mma_tf32(tb, da, db, idesc, i > 0);
        }
        umma::mma_commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    if (warp < 4) {
        for (int part = 0; part < ncols / 32; ++part) {
            float v[32];
            umma::tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + part * 32, v);
            for (int c = 0; c < 32; ++c) out[(warp * 32 + lane) * ncols + part * 32 + c] = v[c];
        }
    }
    umma::fence_before_sync(); __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tb, 256);
}

static float tf32r(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float r; memcpy(&r, &u, 4); return r; }

// descriptor with a RELATIVE start address (bytes), layout type in bits 61..63
static uint64_t mkdesc(uint32_t rel, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
    return (uint64_t)((rel >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout_type << 61);
}
static uint32_t mkidesc(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte-address swizzles
static uint32_t swz_none(uint32_t a) { return a; }
static uint32_t swz_128(uint32_t a) { return a ^ (((a >> 7) & 7u) << 4); }        // Swizzle<3,4,3>
static uint32_t swz_128_32(uint32_t a) { return a ^ (((a >> 7) & 3u) << 5); }     // Swizzle<2,5,2>
static uint32_t swz_128_32b(uint32_t a) { return a ^ (((a >> 8) & 3u) << 5); }    // variant: rows of 256 B?
static uint32_t swz_128_32c(uint32_t a) { return a ^ (((a >> 7) & 7u) >> 1 << 5); }  // variant: (row % 8) / 2

struct Mat { int R, C; std::vector<float> v; float at(int r, int c) const { return v[r * C + c]; } };

int main(int argc, char **argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;   // run ONE experiment per process: a fault is sticky
    int test_id = 0;
    srand(3);
    Mat X{128, 64}, W{64, 64}, Z{128, 64};      // X: [e][k] row-major, W: [j][k], Z: [e][j]
    X.v.resize(128 * 64); W.v.resize(64 * 64); Z.v.resize(128 * 64);
    for (auto &v : X.v) v = tf32r((rand() % 2001 - 1000) / 1000.f);
    for (auto &v : W.v) v = tf32r((rand() % 2001 - 1000) / 1000.f);
    for (auto &v : Z.v) v = tf32r((rand() % 2001 - 1000) / 1000.f);
    uint8_t *dA, *dB; uint64_t *dD; float *dO;
    const uint32_t IMG = 96 * 1024;
    cudaMalloc(&dA, IMG); cudaMalloc(&dB, IMG); cudaMalloc(&dD, 64 * 16); cudaMalloc(&dO, 128 * 256 * 4);
    const size_t smem = 2 * IMG + 2048;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<float> out(128 * 256);

    // run one experiment: images, descriptor list, idesc, expected D [M][N]
    auto run = [&](const char *name, const std::vector<uint8_t> &ia, const std::vector<uint8_t> &ib,
                   const std::vector<uint64_t> &descs, uint32_t idesc, int M, int N,
                   const std::function<double(int, int)> &ref, bool m64 = false) {
        if (only >= 0 && test_id++ != only) return;
        cudaMemcpy(dA, ia.data(), ia.size(), cudaMemcpyHostToDevice);
        cudaMemcpy(dB, ib.data(), ib.size(), cudaMemcpyHostToDevice);
        cudaMemcpy(dD, descs.data(), descs.size() * 8, cudaMemcpyHostToDevice);
        cudaMemset(dO, 0, 128 * 256 * 4);
        probe<<<1, 128, smem>>>(dA, (uint32_t)ia.size(), dB, (uint32_t)ib.size(), dD, (int)descs.size() / 2, idesc, N, dO);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-58s CUDA ERROR: %s\n", name, cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(out.data(), dO, 128 * N * 4, cudaMemcpyDeviceToHost);
        double err = 0, mx = 0, sa = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                const double r = ref(m, n);
                const int lane = m64 ? 32 * (m / 16) + m % 16 : m;      // M = 64: row m sits in TMEM lane 32 (m/16) + m%16
                err = fmax(err, fabs(r - out[lane * N + n])); mx = fmax(mx, fabs(r)); sa += fabs(out[lane * N + n]);
            }
        printf("%-58s err %.3e (max ref %.2f) sum|out| %.3e  %s %s\n", name, err, mx, sa, err < 1e-3 * mx ? "OK  " : "FAIL",
               e == cudaSuccess ? "" : cudaGetErrorString(e));
        if (e != cudaSuccess) exit(1);
    };
    auto put = [](std::vector<uint8_t> &img, uint32_t addr, float v) { memcpy(&img[addr], &v, 4); };

    // ---------------- T0: K-major x K-major, no swizzle (the layout the library uses) -------------------
    {
        const uint32_t CH = 128 * 16, CHW = 64 * 16;
        std::vector<uint8_t> ia(16 * CH, 0), ib(16 * CHW, 0);
        for (int e = 0; e < 128; ++e) for (int k = 0; k < 64; ++k) put(ia, (k >> 2) * CH + e * 16 + (k & 3) * 4, X.at(e, k));
        for (int j = 0; j < 64; ++j) for (int k = 0; k < 64; ++k) put(ib, (k >> 2) * CHW + j * 16 + (k & 3) * 4, W.at(j, k));
        std::vector<uint64_t> d;
        for (int s = 0; s < 8; ++s) { d.push_back(mkdesc(s * 2 * CH, CH, 128, 0)); d.push_back(mkdesc(s * 2 * CHW, CHW, 128, 0)); }
        run("T0  K x K, no swizzle", ia, ib, d, mkidesc(128, 64, false, false), 128, 64,
            [&](int m, int n) { double r = 0; for (int k = 0; k < 64; ++k) r += (double)X.at(m, k) * W.at(n, k); return r; });
    }
    // row-major [rows][64] image in two 32-column panels of `rows` x 128 B, with a byte swizzle
    auto panel_img = [&](const Mat &A, uint32_t (*swz)(uint32_t), uint32_t &panel_stride) {
        panel_stride = A.R * 128;
        std::vector<uint8_t> img(2 * panel_stride, 0);
        for (int r = 0; r < A.R; ++r) for (int c = 0; c < 64; ++c)
            put(img, swz((c >> 5) * panel_stride + r * 128 + (c & 31) * 4), A.at(r, c));
        return img;
    };
    // ---------------- T1: K-major x K-major, standard 128B swizzle (layout type 2) -----------------------
    for (int variant = 0; variant < 2; ++variant) {
        uint32_t pa, pb;
        auto ia = panel_img(X, variant == 0 ? swz_128 : swz_128_32, pa), ib = panel_img(W, variant == 0 ? swz_128 : swz_128_32, pb);
        const uint32_t lt = variant == 0 ? 2 : 1;
        std::vector<uint64_t> d;
        for (int s = 0; s < 8; ++s) {       // k-step s: panel s / 4, 32 B per step inside the 128 B row
            d.push_back(mkdesc((s >> 2) * pa + (s & 3) * 32, 16, 1024, lt));
            d.push_back(mkdesc((s >> 2) * pb + (s & 3) * 32, 16, 1024, lt));
        }
        run(variant == 0 ? "T1a K x K, SW128 (type 2), Swizzle<3,4,3>" : "T1b K x K, type 1 (BASE32B), Swizzle<2,5,2>", ia, ib, d,
            mkidesc(128, 64, false, false), 128, 64,
            [&](int m, int n) { double r = 0; for (int k = 0; k < 64; ++k) r += (double)X.at(m, k) * W.at(n, k); return r; });
    }
    // ---------------- T2: A K-major (no swizzle), B MN-major type 1 over row-major W'[j][k] --------------
    //   D[e][k] = sum_j Z[e][j] W[j][k]:  B = W read as N = k (contiguous), K = j (rows of 128 B)
    {
        const uint32_t CH = 128 * 16;
        std::vector<uint8_t> ia(16 * CH, 0);
        for (int e = 0; e < 128; ++e) for (int j = 0; j < 64; ++j) put(ia, (j >> 2) * CH + e * 16 + (j & 3) * 4, Z.at(e, j));
        struct V { const char *name; uint32_t (*swz)(uint32_t); uint32_t lt; bool swap; };
        const V vs[] = {{"T2a B MN type1 <2,5,2> lbo=panel sbo=512", swz_128_32, 1, false},
                        {"T2b B MN type1 <2,5,2> lbo=512 sbo=panel", swz_128_32, 1, true},
                        {"T2c B MN type1 swz(row/2%4) lbo=panel sbo=512", swz_128_32c, 1, false},
                        {"T2d B MN type1 swz(a>>8) lbo=panel sbo=512", swz_128_32b, 1, false},
                        {"T2e B MN type2 <3,4,3> lbo=panel sbo=1024", swz_128, 2, false},
                        {"T2f B MN type0 none lbo=panel sbo=512", swz_none, 0, false}};
        for (const V &v : vs) {
            uint32_t pb;
            auto ib = panel_img(W, v.swz, pb);      // rows j (K), 32 k (N) per 128 B row, panel = k / 32
            std::vector<uint64_t> d;
            for (int s = 0; s < 8; ++s) {           // k-step s covers j = 8 s .. 8 s + 7 = 8 rows = 1024 B
                const uint32_t katom = v.lt == 2 ? 1024 : 512;
                d.push_back(mkdesc(s * 2 * CH, CH, 128, 0));
                d.push_back(v.swap ? mkdesc(s * 1024, katom, pb, v.lt) : mkdesc(s * 1024, pb, katom, v.lt));
            }
            run(v.name, ia, ib, d, mkidesc(128, 64, false, true), 128, 64,
                [&](int m, int n) { double r = 0; for (int j = 0; j < 64; ++j) r += (double)Z.at(m, j) * W.at(j, n); return r; });
        }
    }
    // ---------------- T3: both MN-major type 1: D[j][k] = sum_e Z[e][j] X[e][k]  (the dW2 contraction) ----
    {
        struct V { const char *name; uint32_t (*swz)(uint32_t); bool swap; };
        const V vs[] = {{"T3a A,B MN type1 <2,5,2> lbo=panel sbo=512", swz_128_32, false},
                        {"T3b A,B MN type1 <2,5,2> lbo=512 sbo=panel", swz_128_32, true}};
        for (const V &v : vs) {
            uint32_t pa, pb;
            auto ia = panel_img(Z, v.swz, pa), ib = panel_img(X, v.swz, pb);
            std::vector<uint64_t> d;
            for (int s = 0; s < 16; ++s) {          // 8 edges per step = 8 rows = 1024 B
                d.push_back(v.swap ? mkdesc(s * 1024, 512, pa, 1) : mkdesc(s * 1024, pa, 512, 1));
                d.push_back(v.swap ? mkdesc(s * 1024, 512, pb, 1) : mkdesc(s * 1024, pb, 512, 1));
            }
            run(v.name, ia, ib, d, mkidesc(64, 64, true, true), 64, 64,
                [&](int m, int n) { double r = 0; for (int e = 0; e < 128; ++e) r += (double)Z.at(e, m) * X.at(e, n); return r; }, true);
        }
    }
    // ---------------- T4: the SAME type-1 image of X as K-major A (G1) : D[e][j] = sum_k X[e][k] W[j][k] ----
    {
        uint32_t pa, pb;
        auto ia = panel_img(X, swz_128_32, pa), ib = panel_img(W, swz_128_32, pb);
        for (int sbo_variant = 0; sbo_variant < 2; ++sbo_variant) {
            std::vector<uint64_t> d;
            for (int s = 0; s < 8; ++s) {
                d.push_back(mkdesc((s >> 2) * pa + (s & 3) * 32, 16, sbo_variant ? 512 : 1024, 1));
                d.push_back(mkdesc((s >> 2) * pb + (s & 3) * 32, 16, sbo_variant ? 512 : 1024, 1));
            }
            run(sbo_variant ? "T4b K x K on type-1 images, sbo=512" : "T4a K x K on type-1 images, sbo=1024", ia, ib, d,
                mkidesc(128, 64, false, false), 128, 64,
                [&](int m, int n) { double r = 0; for (int k = 0; k < 64; ++k) r += (double)X.at(m, k) * W.at(n, k); return r; });
        }
    }
    // ---------------- T5: M = 128 MN-major A over two stacked images ([Z ; 2Z]: 4 panels of 32 j) ----------
    {
        uint32_t pa, pb;
        auto iz = panel_img(Z, swz_128_32, pa);
        std::vector<uint8_t> ia(4 * pa, 0);
        memcpy(&ia[0], iz.data(), 2 * pa);
        for (uint32_t i = 0; i < 2 * pa; i += 4) { float v; memcpy(&v, &iz[i], 4); v *= 2.f; memcpy(&ia[2 * pa + i], &v, 4); }
        auto ib = panel_img(X, swz_128_32, pb);
        std::vector<uint64_t> d;
        for (int s = 0; s < 16; ++s) { d.push_back(mkdesc(s * 1024, pa, 512, 1)); d.push_back(mkdesc(s * 1024, pb, 512, 1)); }
        run("T5  A MN M=128 (4 panels), B MN, type 1", ia, ib, d, mkidesc(128, 64, true, true), 128, 64,
            [&](int m, int n) { double r = 0; for (int e = 0; e < 128; ++e) r += (double)(m < 64 ? 1.0 : 2.0) * Z.at(e, m % 64) * X.at(e, n); return r; });
    }
    return 0;
}
