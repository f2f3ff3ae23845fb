// Thin inline-PTX layer over the sm_100a tensor-core path (tcgen05.mma with TMEM accumulators),
// used by the 3xTF32 kernels (node_linear.cu, edge_scorer_tc.cu).
//
// Operand layout used throughout: "chunk-interleaved", no swizzle.  A [R x K] fp32 operand lives in
// shared memory as K/4 chunks; chunk c holds, for every row r, the 16 bytes (k = 4c .. 4c+3):
//         byte offset(r, k) = (k / 4) * CHUNK + r * 16 + (k % 4) * 4 ,   CHUNK >= R * 16.
// Read as a K-major operand (rows = M or N) this is the canonical no-swizzle layout
// ((8,n),2):((1,SBO),LBO) in 16-byte units with SBO = 128 B (next 8 rows) and LBO = CHUNK (next 4 k);
// read as an MN-major operand whose MN index is k and whose K index is r it is the canonical
// ((1,n),(8,k)):((X,SBO),(1,LBO)) with SBO = CHUNK (next 4 mn) and LBO = 128 B (next 8 k) — the same
// bytes serve as  X  (K-major) and as  X^T  (MN-major), which the edge scorer uses for dW2.
// Descriptor bit fields follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor / InstrDescriptor).
//
// 3xTF32: an fp32 value x is split into hi = rn_tf32(x) and lo = x - hi (exact in fp32; the tensor
// core reads its top 19 bits); a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi accumulated in fp32
// (dropped terms ~2^-21 relative, signs random).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pangnn {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- 3xTF32 split -----------------------------------------------------------------------------
// hi = x rounded to TF32 (nearest, ties away from zero): two integer ops on the bit pattern — the
// cvt.rna.tf32.f32 instruction is emulated in SASS with Inf/NaN guards (~5 instructions) and showed
// up as 20 % of the scorer's issue slots.  Inputs are finite activations / weights.
__device__ __forceinline__ float tf32_hi(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
// lo = x - hi is exact in fp32 (|lo| <= 2^-11 |x|, random sign).  It is handed to the tensor core
// as is: the hardware ignores its 13 low mantissa bits, an error of <= 2^-21 |x| with random sign.
__device__ __forceinline__ float tf32_lo(float x, float hi) { return x - hi; }
__device__ __forceinline__ void split4(const float4 v, float4 &hi, float4 &lo) {
    hi.x = tf32_hi(v.x); hi.y = tf32_hi(v.y); hi.z = tf32_hi(v.z); hi.w = tf32_hi(v.w);
    lo.x = tf32_lo(v.x, hi.x); lo.y = tf32_lo(v.y, hi.y); lo.z = tf32_lo(v.z, hi.z); lo.w = tf32_lo(v.w, hi.w);
}

// ---- descriptors --------------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}
// MN-major tf32 operands exist in ONE layout only, "128B swizzle with 32-byte atomicity" (layout type 1;
// CUTLASS UMMA::LayoutType::SWIZZLE_128B_BASE32B, Swizzle<2,5,2>) — probed on B200 (tools/umma_probe2.cu: type 0
// and type 2 MN-major descriptors yield zeros, type 1 on a K-major descriptor traps "misaligned address").
// The operand is the ROW-MAJOR matrix itself, [K rows][MN contiguous], in panels of 32 MN elements:
//     byte(k, mn) = (mn / 32) * PANEL + k * 128 + (((mn % 32) / 8) ^ (k & 3)) * 32 + (mn % 8) * 4
// with LBO = PANEL (next 32 MN), SBO = 512 B (next 4 K rows); a k-step of 8 rows advances the start by 1024 B.
// Panels and the operand base must be 512-byte aligned (the swizzle uses address bits 7..8).
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t saddr, uint32_t panel_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((panel_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((512u >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// byte offset of the 16-byte piece holding mn .. mn+3 (mn % 4 == 0) of K-row k inside such an operand
__device__ __forceinline__ uint32_t mn_off(uint32_t k, uint32_t mn, uint32_t panel_bytes) {
    return (mn >> 5) * panel_bytes + k * 128u + ((((mn & 31u) >> 3) ^ (k & 3u)) << 5) + ((mn & 7u) << 2);
}

// instruction descriptor: kind::tf32, fp32 accumulate, dense
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) |
           ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM ------------------------------------------------------------------------------------------
// one full warp; writes the allocated base address to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 lanes x 32 consecutive columns: thread i of the warp receives lane (taddr.lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
template <int C>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[C]) {
    static_assert(C == 16 || C == 32, "16 or 32 columns");
    if constexpr (C == 16) tmem_ld16(taddr, v); else tmem_ld32(taddr, v);
}

// two loads in flight, one wait (the epilogues read a main and a small-terms accumulator)
template <int C>
__device__ __forceinline__ void tmem_ld2(uint32_t ta, float (&a)[C], uint32_t tb, float (&b)[C]) {
    if constexpr (C != 16) {
        tmem_ld<C>(ta, a);
        tmem_ld<C>(tb, b);
        return;
    } else {
    uint32_t r[16], q[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(ta) : "memory");
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]),
          "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
        : "r"(tb) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(q[i]); }
    }
}

// 32 lanes x 8 consecutive columns, registers -> TMEM (thread i of the warp writes lane taddr.lane + i).  The caller
// issues tmem_st_wait() before handing the data to another thread / the tensor core.
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
        :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
           "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
           "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- MMA + completion -------------------------------------------------------------------------------
// A operand read from TENSOR MEMORY (lanes = M rows, one 32-bit column per tf32 K element; K-major by construction),
// B from shared memory: D[tmem] (+)= A[tmem] * B[smem].  Used where the A tile is produced by the epilogue threads in
// the accumulator's own thread <-> row layout (the scorer's da2), which then never touches shared memory.
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the mbarrier when every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}

// ---- mbarrier -----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "W_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\tbra W_%=;\n"
        "D_%=:\n\t}\n"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Advancing a descriptor's start address (low 14 bits, 16-byte units) by `bytes`: one 64-bit add.
// Building descriptors from scratch inside the issue loop cost ~60 dependent cycles each on the
// single issuing thread (ncu: 14 % of the scorer's stall samples sat on the issue branch).
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }

// 3xTF32 product of two chunk-interleaved operands over KSTEPS MMA k-steps (8 k each), fully
// unrolled.  a_step / b_step: byte advance of the start address per k-step (K-major: 2 * CHUNK).
// The tensor core adds into its fp32 accumulator with truncation (~2e-8 relative per tcgen05.mma,
// same sign every time), so the large hi*hi terms and the 2^-11-times-smaller correction terms go to
// SEPARATE accumulators: the main chain is a third as long and the truncation of the small chain is
// negligible; the epilogue adds the two in fp32 (round-to-nearest).  Pass d_small == d_main to use
// one accumulator.
template <int KSTEPS>
__device__ __forceinline__ void mma_3xtf32(uint32_t d_main, uint32_t d_small, uint32_t a_hi, uint32_t a_lo,
                                           uint32_t b_hi, uint32_t b_lo, uint32_t a_lbo, uint32_t a_sbo,
                                           uint32_t a_step, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_step,
                                           uint32_t idesc, bool accumulate_first) {
    const uint64_t ah0 = smem_desc(a_hi, a_lbo, a_sbo), al0 = smem_desc(a_lo, a_lbo, a_sbo);
    const uint64_t bh0 = smem_desc(b_hi, b_lbo, b_sbo), bl0 = smem_desc(b_lo, b_lbo, b_sbo);
    const bool split = d_small != d_main;
#pragma unroll
    for (int s = 0; s < KSTEPS; ++s) {
        const uint64_t ah = desc_advance(ah0, s * a_step), al = desc_advance(al0, s * a_step);
        const uint64_t bh = desc_advance(bh0, s * b_step), bl = desc_advance(bl0, s * b_step);
        const uint32_t acc = (accumulate_first || s > 0) ? 1u : 0u;
        mma_tf32(d_small, al, bh, idesc, acc);                                      // small terms
        mma_tf32(d_small, ah, bl, idesc, 1u);
        mma_tf32(d_main, ah, bh, idesc, split ? acc : 1u);
    }
}

}  // namespace umma
}  // namespace pangnn
