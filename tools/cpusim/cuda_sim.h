// cuda_sim.h — DEVELOPER TOOL ONLY.  A minimal host-thread emulation of the CUDA execution
// model (one std::thread per CUDA thread, one block at a time) so the device code in
// nanowakeword_b200/csrc/*.cuh can be exercised for indexing/numerics triage in a container
// without a GPU.  It is not part of the product, the tests, the bench or smoke().
#pragma once
#ifndef __shared__
#define __shared__ static      // one block runs at a time in the host model
#endif
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <barrier>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __constant__

struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct int4 { int x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return {a, b, c, d}; }
inline float __uint_as_float(unsigned v) { float f; memcpy(&f, &v, 4); return f; }
inline unsigned __float_as_uint(float v) { unsigned f; memcpy(&f, &v, 4); return f; }
inline float2 make_float2(float a, float b) { return {a, b}; }

namespace cudasim {
inline thread_local dim3 t_threadIdx;
inline dim3 g_blockIdx, g_blockDim, g_gridDim;
inline std::barrier<>* g_block_barrier = nullptr;
inline std::vector<std::unique_ptr<std::barrier<>>> g_warp_barriers;
inline std::vector<uint64_t> g_shfl;     // one slot per thread
inline unsigned char* g_dyn_smem = nullptr;
inline std::vector<std::unique_ptr<std::barrier<>>> g_named_barriers;   // id -> barrier (fixed participant count)
inline void named_barrier(int id, int count) {
    (void)count;
    g_named_barriers[id]->arrive_and_wait();
}
inline int linear_tid() { return (int)(t_threadIdx.x + g_blockDim.x * (t_threadIdx.y + g_blockDim.y * t_threadIdx.z)); }
}  // namespace cudasim

#define threadIdx cudasim::t_threadIdx
#define blockIdx cudasim::g_blockIdx
#define blockDim cudasim::g_blockDim
#define gridDim cudasim::g_gridDim
#define warpSize 32

inline void __syncthreads() { cudasim::g_block_barrier->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { cudasim::g_warp_barriers[cudasim::linear_tid() / 32]->arrive_and_wait(); }
inline void __threadfence() {}
inline void __threadfence_block() {}

template <typename T> inline T __shfl_generic(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload too large");
    const int tid = cudasim::linear_tid();
    const int warp = tid / 32;
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    cudasim::g_shfl[tid] = raw;
    cudasim::g_warp_barriers[warp]->arrive_and_wait();
    uint64_t got = cudasim::g_shfl[warp * 32 + (src_lane & 31)];
    cudasim::g_warp_barriers[warp]->arrive_and_wait();
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}
template <typename T> inline T __shfl_sync(unsigned, T v, int lane) { return __shfl_generic(v, lane); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int mask) { return __shfl_generic(v, (cudasim::linear_tid() & 31) ^ mask); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, int d) {
    int lane = cudasim::linear_tid() & 31;
    return __shfl_generic(v, lane + d < 32 ? lane + d : lane);
}
template <typename T> inline T __ldg(const T* p) { return *p; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return {fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }   // FFMA2: two IEEE FMAs
inline float __fdividef(float a, float b) { return a / b; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
inline int __float_as_int(float v) { int f; memcpy(&f, &v, 4); return f; }
using std::max;
using std::min;

namespace cudasim {
// Run `kernel(args...)` for every block (sequentially) with block_threads host threads each.
template <typename F> void launch(dim3 grid, dim3 block, size_t dyn_smem, F&& body) {
    g_gridDim = grid;
    g_blockDim = block;
    const int nthreads = (int)(block.x * block.y * block.z);
    std::vector<unsigned char> smem(dyn_smem + 1024);
    g_dyn_smem = (unsigned char*)(((uintptr_t)smem.data() + 1023) & ~(uintptr_t)1023);
    g_shfl.assign(nthreads, 0);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_blockIdx = dim3(bx, by, bz);
                std::barrier<> bar(nthreads);
                g_block_barrier = &bar;
                g_warp_barriers.clear();
                g_named_barriers.clear();
                for (int i = 0; i < 16; ++i) g_named_barriers.emplace_back(new std::barrier<>(128));
                for (int w = 0; w < (nthreads + 31) / 32; ++w)
                    g_warp_barriers.emplace_back(new std::barrier<>(std::min(32, nthreads - 32 * w)));
                std::vector<std::thread> ts;
                ts.reserve(nthreads);
                for (int t = 0; t < nthreads; ++t)
                    ts.emplace_back([&, t] {
                        t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                        body();
                    });
                for (auto& th : ts) th.join();
            }
}
}  // namespace cudasim

// `extern __shared__ T name[]` becomes a reference to the per-launch dynamic buffer.
#define NWW_DYN_SMEM(name) unsigned char* name = cudasim::g_dyn_smem
