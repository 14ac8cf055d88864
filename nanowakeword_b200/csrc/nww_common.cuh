// nww_common.cuh — shared device helpers for the B200 wake-word engine.
//
// Everything here is sm_100a CUDA.  The NWW_CPUSIM branch exists only so that the
// developer tool under tools/cpusim/ can run the same device code on host threads to
// triage indexing/numerics without a GPU; it is never compiled into libnwwb200.so and
// nothing in the product, tests, bench or smoke() uses it.
#pragma once

#include <stdint.h>

#ifdef NWW_CPUSIM
#include "cuda_sim.h"
#else
#include <cuda_runtime.h>
#endif

namespace nww {

constexpr int kMaxTailLayers = 8;

enum Activation : int { ACT_RELU = 0, ACT_GELU = 1, ACT_SILU = 2 };

template <typename T> struct cplx { T x, y; };

template <typename T> __device__ __forceinline__ cplx<T> cadd(cplx<T> a, cplx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T> __device__ __forceinline__ cplx<T> csub(cplx<T> a, cplx<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T> __device__ __forceinline__ cplx<T> cmul(cplx<T> a, cplx<T> b) {
    return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
// multiply by -i (forward-DFT quarter turn): (x + iy)(-i) = y - ix
template <typename T> __device__ __forceinline__ cplx<T> mul_mi(cplx<T> a) { return {a.y, -a.x}; }
// multiply by +i
template <typename T> __device__ __forceinline__ cplx<T> mul_pi(cplx<T> a) { return {-a.y, a.x}; }

__device__ __forceinline__ float apply_act(float x, int act) {
    if (act == ACT_RELU) return fmaxf(x, 0.0f);
    if (act == ACT_GELU) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));   // exact-erf GELU
    return x / (1.0f + expf(-x));                                                        // SiLU
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------------------------------
// Asynchronous staging of one window's PCM into shared memory: a 1-D TMA bulk copy
// (cp.async.bulk, SASS UBLKCP) completing on an mbarrier.  One elected thread issues it;
// everybody waits on the barrier phase.
// ---------------------------------------------------------------------------------------
#ifndef NWW_CPUSIM
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// a wait executed by ONE thread of the CTA (the MMA / copy issuer): the same instruction on the device; in the host
// model MMAs and bulk copies are synchronous, so it is a no-op there (mbar_wait itself models a CTA-wide barrier)
__device__ __forceinline__ void mbar_wait_one(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)((const unsigned char*)p - cudasim::g_dyn_smem); }
__device__ __forceinline__ void mbar_init(uint64_t*, uint32_t) {}
__device__ __forceinline__ void fence_mbar_init() {}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t*, uint32_t) {}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t*) { memcpy(dst, src, bytes); }
// every thread of the block calls wait(), so a block barrier orders the (synchronous) copy
__device__ __forceinline__ void mbar_wait(uint64_t*, uint32_t) { __syncthreads(); }
__device__ __forceinline__ void mbar_wait_one(uint64_t*, uint32_t) {}
#endif

}  // namespace nww
