// nww_tc.cuh — thin PTX wrappers for the 5th-generation tensor core path (tcgen05 + TMEM),
// shared by the dense-layer GEMM (nww_gemm_tc.cuh) and the fused CNN stage (nww_cnn2.cuh).
//
// The NWW_CPUSIM branch is a functional model of the same operations (shared-memory matrix
// descriptors decoded in software, TMEM as a plain array) used only by the developer tool under
// tools/cpusim/ to check index arithmetic without a GPU; it is never compiled into the library.
#pragma once

#include "nww_common.cuh"

namespace nww {

// ---- shared-memory matrix descriptors (cute::UMMA::SmemDescriptor bit layout) ------------------
// bits [0,14)  start address >> 4        bits [16,30) leading-dimension byte offset >> 4
// bits [32,46) stride byte offset >> 4   bits [46,48) version = 1 (sm_100)
// bits [61,64) layout: 0 = no swizzle ("interleave"), 2 = SWIZZLE_128B
//
// K-major, no swizzle: the operand is made of 8-row x 16-byte core matrices stored as 128
// contiguous bytes; SBO = byte distance between core matrices adjacent in M/N (next 8 rows),
// LBO = byte distance between the two core matrices adjacent in K (the two 16-byte halves of a
// K = 32-byte MMA slice).
__host__ __device__ constexpr uint64_t umma_desc_fields(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ uint64_t umma_desc_noswz(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | umma_desc_fields(lbo_bytes, sbo_bytes, 0);
}
// K-major SWIZZLE_128B tile (rows of 128 bytes, 8-row swizzle atoms): SBO = 1024, LBO unused (1).
__device__ __forceinline__ uint64_t umma_desc_sw128_addr(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// instruction descriptors: D = F32 (bit 4), A/B format at bits 7/10 (F16 = 0, BF16 = 1, TF32 = 2),
// both operands K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n) {       // kind::f16 with IEEE half operands
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

#ifndef NWW_CPUSIM
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 32 bit, 16 consecutive columns: thread `lane` of the warp gets TMEM[lane_base + lane][col .. col + 15]
__device__ __forceinline__ void tmem_ld_32x32b_x16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// named barrier among `count` threads (count a multiple of 32); id 0 is __syncthreads' barrier
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ float bf16_bits_to_float(uint32_t b) { return __uint_as_float(b << 16); }
// round to nearest even.  The packed form compiles to F2FP.BF16.F32.PACK_AB on the ALU pipe; the scalar cvt.rn.bf16.f32
// becomes F2F on the quarter-rate XU pipe, and the integer emulation used before cost three ALU operations.
__device__ __forceinline__ uint32_t float_to_bf16_bits(float x) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(0.0f), "f"(x));
    return d & 0xFFFFu;
}
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x));
    return __uint_as_float(t);
}
#else
// ------------------------------------------------------------------------- functional model
namespace sim {
inline float g_tmem[128][512];
inline float bf16f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }
inline float halff(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu;
    float f;
    if (e == 0) {
        f = (float)m * (1.0f / 16777216.0f);
        return sign ? -f : f;
    }
    const uint32_t u = sign | ((e + 112u) << 23) | (m << 13);
    memcpy(&f, &u, 4);
    return f;
}
}
__device__ __forceinline__ void tc_fence_before() {}
__device__ __forceinline__ void tc_fence_after() {}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t) { *dst_smem = 0; }
__device__ __forceinline__ void tmem_dealloc(uint32_t, uint32_t) {}
inline void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    const int M = (int)((idesc >> 24) & 0x1F) << 4, N = (int)((idesc >> 17) & 0x3F) << 3;
    auto field = [](uint64_t d, int sh) { return (uint32_t)((d >> sh) & 0x3FFF) << 4; };
    const unsigned char* base = cudasim::g_dyn_smem;
    const uint32_t a0 = field(da, 0), alb = field(da, 16), asb = field(da, 32);
    const uint32_t b0 = field(db, 0), blb = field(db, 16), bsb = field(db, 32);
    const int col0 = (int)(tmem_d & 0xFFFF), lane0 = (int)(tmem_d >> 16);
    const bool half_fmt = ((idesc >> 7) & 7u) == 0u;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0.f;
            for (int k = 0; k < 16; ++k) {
                uint16_t av, bv;
                memcpy(&av, base + a0 + (m / 8) * asb + (m % 8) * 16 + (k / 8) * alb + (k % 8) * 2, 2);
                memcpy(&bv, base + b0 + (n / 8) * bsb + (n % 8) * 16 + (k / 8) * blb + (k % 8) * 2, 2);
                s += half_fmt ? sim::halff(av) * sim::halff(bv) : sim::bf16f(av) * sim::bf16f(bv);
            }
            float& d = sim::g_tmem[lane0 + m][col0 + n];
            d = accumulate ? d + s : s;
        }
}
inline void umma_commit(uint64_t*) {}        // MMAs are synchronous here; mbar_wait is a CTA barrier in the model
__device__ __forceinline__ void tmem_ld_wait() {}
inline void tmem_ld_32x32b_x16_nowait(uint32_t taddr, uint32_t* r) {
    const int lane = (int)(taddr >> 16) + (cudasim::linear_tid() & 31), col = (int)(taddr & 0xFFFF);
    for (int i = 0; i < 16; ++i) memcpy(&r[i], &sim::g_tmem[lane][col + i], 4);
}
inline void tmem_ld_32x32b_x32(uint32_t taddr, float* v) {
    const int lane = (int)(taddr >> 16) + (cudasim::linear_tid() & 31), col = (int)(taddr & 0xFFFF);
    for (int i = 0; i < 32; ++i) v[i] = sim::g_tmem[lane][col + i];
}
inline void named_bar_sync(int id, int count) { cudasim::named_barrier(id, count); }
inline float bf16_bits_to_float(uint32_t b) { return sim::bf16f((uint16_t)b); }
inline uint32_t float_to_bf16_bits(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return u >> 16;
}
inline float round_tf32(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}
#endif

}  // namespace nww
