// Hardware probe for the no-swizzle MN-major descriptor fields (development aid; not built into the
// library).  nvcc -gencode arch=compute_100a,code=sm_100a -I pangnn_b200/csrc tools/umma_probe.cu -o /tmp/umma_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "umma.cuh"
using namespace pangnn;

constexpr int R = 128, C = 64;
constexpr uint32_t CH = R * 16 + 16;

// X [R=128 rows][C=64 cols] stored chunk-interleaved over cols; W [64][64] likewise (CHW)
constexpr uint32_t CHW = 64 * 16 + 16;

__global__ void probe(const float *X, const float *W, float *out, int mode, uint32_t lbo, uint32_t sbo, uint32_t step,
                      uint32_t albo, uint32_t asbo, uint32_t astep) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tb;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *sX = smem, *sW = smem + 16 * CH * 2;
    for (int i = tid; i < 2 * 16 * CH / 4; i += blockDim.x) ((float *)sX)[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < R * C; i += blockDim.x) {
        int r = i / C, c = i % C;
        *(float *)(sX + (c >> 2) * CH + r * 16 + (c & 3) * 4) = X[i];
        *(float *)(sX + 16 * CH + (c >> 2) * CH + r * 16 + (c & 3) * 4) = 2.f * X[i];   // "lo" buffer = 2X for G3 probe
    }
    for (int i = tid; i < 64 * 64; i += blockDim.x) {
        int r = i / 64, c = i % 64;
        *(float *)(sW + (c >> 2) * CHW + r * 16 + (c & 3) * 4) = W[i];
    }
    if (warp == 0) umma::tmem_alloc(&tb, 128);
    if (tid == 32) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
    umma::fence_async_smem(); umma::fence_before_sync(); __syncthreads(); umma::fence_after_sync();
    const uint32_t sb = umma::smem_u32(smem), sx = sb, sw = sb + 16 * CH * 2;
    if (tid == 0) {
        if (mode == 0) {          // D[e][j] = sum_k X[e][k] W[j][k]   (K-major, K-major)
            for (int s = 0; s < 8; ++s)
                umma::mma_tf32(tb, umma::smem_desc(sx + s * 2 * CH, CH, 128), umma::smem_desc(sw + s * 2 * CHW, CHW, 128),
                               umma::idesc_tf32(128, 64, false, false), s > 0);
        } else if (mode == 1) {   // D[e][k] = sum_j X[e][j] W[j][k]   (A K-major, B MN-major)
            for (int s = 0; s < 8; ++s)
                umma::mma_tf32(tb, umma::smem_desc(sx + s * 2 * CH, CH, 128), umma::smem_desc(sw + s * step, lbo, sbo),
                               umma::idesc_tf32(128, 64, false, true), s > 0);
        } else {                  // D[j'][k] = sum_e [X ; 2X][e][j'] X[e][k]   (A MN-major M=128, B MN-major)
            for (int s = 0; s < 16; ++s)
                umma::mma_tf32(tb, umma::smem_desc(sx + s * astep, albo, asbo), umma::smem_desc(sx + s * step, lbo, sbo),
                               umma::idesc_tf32(128, 64, true, true), s > 0);
        }
        umma::mma_commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    if (warp < 4) {
        for (int part = 0; part < 2; ++part) {
            float v[32];
            umma::tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + part * 32, v);
            for (int c = 0; c < 32; ++c) out[(warp * 32 + lane) * 64 + part * 32 + c] = v[c];
        }
    }
    umma::fence_before_sync(); __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tb, 128);
}

static float tf32r(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float r; memcpy(&r, &u, 4); return r; }

int main() {
    std::vector<float> X(R * C), W(64 * 64), out(128 * 64);
    srand(1);
    for (auto &v : X) v = tf32r((rand() % 2001 - 1000) / 1000.f);
    for (auto &v : W) v = tf32r((rand() % 2001 - 1000) / 1000.f);
    float *dX, *dW, *dO;
    cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dO, out.size() * 4);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = 16 * CH * 2 + 16 * CHW + 256;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    auto run = [&](int mode, uint32_t lbo, uint32_t sbo, uint32_t step, uint32_t albo, uint32_t asbo, uint32_t astep, const char *name) {
        cudaMemset(dO, 0, out.size() * 4);
        probe<<<1, 256, smem>>>(dX, dW, dO, mode, lbo, sbo, step, albo, asbo, astep);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
        double err = 0, mx = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
                double ref = 0;
                if (mode == 0) for (int k = 0; k < 64; ++k) ref += (double)X[m * 64 + k] * W[n * 64 + k];
                else if (mode == 1) for (int j = 0; j < 64; ++j) ref += (double)X[m * 64 + j] * W[j * 64 + n];
                else for (int e2 = 0; e2 < 128; ++e2) ref += (double)(m < 64 ? 1.0 : 2.0) * X[e2 * 64 + (m % 64)] * X[e2 * 64 + n];
                err = fmax(err, fabs(ref - out[m * 64 + n])); mx = fmax(mx, fabs(ref));
            }
        double sa = 0; for (float v : out) sa += fabs(v);
        printf("%-40s mode %d  err %.3e (max ref %.3f)  sum|out| %.3e  out[0..3] %g %g %g %g  cuda=%s\n", name, mode, err, mx, sa,
               out[0], out[1], out[2], out[3], cudaGetErrorString(e));
    };
    run(0, 0, 0, 0, 0, 0, 0, "K-major x K-major");
    run(1, 128, CHW, 128, 0, 0, 0, "B MN: lbo=128 sbo=CHW step=128");
    run(1, CHW, 128, 128, 0, 0, 0, "B MN: lbo=CHW sbo=128 step=128");
    run(2, 128, CH, 128, 128, CH, 128, "A,B MN: lbo=128 sbo=CH step=128");
    run(2, CH, 128, 128, CH, 128, 128, "A,B MN: lbo=CH sbo=128 step=128");
    return 0;
}
