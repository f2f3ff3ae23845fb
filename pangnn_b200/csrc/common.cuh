// Shared helpers for the pangnn_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pangnn_b200.h"

namespace pangnn {

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

void set_error(const char *fmt, ...);

inline int check_cuda(cudaError_t e, const char *what) {
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return PANGNN_ECUDA;
    }
    return PANGNN_OK;
}

#define PANGNN_CHECK_LAUNCH(name)                                           \
    do {                                                                    \
        int _rc = ::pangnn::check_cuda(cudaGetLastError(), name);           \
        if (_rc != PANGNN_OK) return _rc;                                   \
    } while (0)

#define PANGNN_REQUIRE(cond, msg)                                           \
    do {                                                                    \
        if (!(cond)) {                                                      \
            ::pangnn::set_error("%s: %s", __func__, msg);                   \
            return PANGNN_EINVAL;                                           \
        }                                                                   \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Carves aligned sub-buffers out of a caller-provided workspace.
struct Workspace {
    char *base;
    size_t size, off;
    Workspace(void *p, size_t n) : base(static_cast<char *>(p)), size(n), off(0) {}
    template <typename T>
    T *take(size_t count) {
        off = align_up(off, 256);
        T *r = reinterpret_cast<T *>(base + off);
        off += count * sizeof(T);
        return r;
    }
    bool ok() const { return off <= size; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Streaming (read-once) 128-bit load that does not allocate in L1.
__device__ __forceinline__ float4 ld_stream_f4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

}  // namespace pangnn
