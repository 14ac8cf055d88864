// nww_gemm_tc.cuh — the dense layer that follows the per-window stage, on 5th-gen tensor cores.
//
//   out[m][n] = post( sum_k X[m][k] * W[n][k] + b[n] ),   X = feature rows of a chunk of windows
//
// (fc1 of CNNModel, architectures.py:64,77; layer1 + LayerNorm of Net, :107-119; fc1 + folded
//  BatchNorm1d of E2E_MelSpectrogram_CNN, :861-862.)
//
// tcgen05.mma kind::tf32 with FP32 accumulation in TMEM, fed by TMA (SWIZZLE_128B, K-major tiles
// of 128 rows x 32 floats) through a 3-stage mbarrier pipeline; one elected thread issues the MMAs.
// Numerics: TF32 keeps 10 mantissa bits, which is NOT enough for the 1e-3 score budget after the
// classifier gain, so both operands are split  x = x_hi + x_lo  (x_hi = RN_tf32(x), x_lo = x - x_hi)
// and three products are accumulated:  x_hi*w_hi + x_lo*w_hi + x_hi*w_lo  (the dropped x_lo*w_lo is
// ~2^-22 relative) — "3xTF32", FP32-class accuracy at tensor-core rate.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> bias / LayerNorm / activation -> global).
// One CTA per 128-row tile of the chunk; accumulator 128 lanes x N columns.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include "nww_common.cuh"
#include "nww_tail.cuh"
#include "nww_tc.cuh"

namespace nww {

constexpr int kTcBM = 128;          // rows (windows) per CTA
constexpr int kTcBK = 32;           // floats per K tile = one 128-byte swizzle row
constexpr int kTcStages = 3;
constexpr int kTcThreads = 192;
constexpr int kTcMaxN = 128;
constexpr int kTcMaxSplits = 16;    // split-K slabs of the partial-sum buffer
constexpr int kTcKbPerSplit = 20;   // K blocks (of 32 floats) per split

__host__ __device__ constexpr size_t tc_stage_bytes(int n) { return (size_t)(2 * kTcBM + 2 * n) * kTcBK * sizeof(float); }
__host__ __device__ constexpr size_t tc_smem_bytes(int n) { return kTcStages * tc_stage_bytes(n) + 1024 /*align*/ + 256 /*barriers*/; }

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
struct GemmTcArgs {
    const float* bias;      // [N]
    const float* ln_g;      // [N] or null
    const float* ln_b;
    float* out;             // [M][N]
    int M, N, K;
    int post, act;
    // split-K: gridDim.y CTAs share one row tile; CTA y reduces K blocks [y * kb_per_split, ...) and
    // writes its raw FP32 partial sums to partial[y][M_pad][N]; bias / post-op then happen in the
    // consumer (tail_kernel's pre-stage), which adds the partials in a fixed order (deterministic).
    float* partial;         // null: single-pass GEMM with the fused epilogue
    int kb_per_split;
    int m_pad;              // rows per split slab of `partial`
};

// ReLU inline, GELU / SiLU out of line: the epilogue below is unrolled over up to 128 columns, and 256 inlined three-way
// activation switches made this the largest kernel of the library (20 k instructions)
__device__ __noinline__ float tc_act_slow(float x, int act) { return apply_act(x, act); }
__device__ __forceinline__ float tc_act(float x, int act) { return act == ACT_RELU ? fmaxf(x, 0.0f) : tc_act_slow(x, act); }

__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo,
                   const __grid_constant__ CUtensorMap tm_whi, const __grid_constant__ CUtensorMap tm_wlo, GemmTcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const size_t stage_bytes = tc_stage_bytes(a.N);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kTcStages * stage_bytes);
    uint64_t* empty = full + kTcStages;
    uint64_t* acc_full = empty + kTcStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kTcBM;
    const int nkb_total = a.K / kTcBK;
    const int kb0 = a.partial ? (int)blockIdx.y * a.kb_per_split : 0;
    const int nkb = a.partial ? min(a.kb_per_split, nkb_total - kb0) : nkb_total;   // K blocks of this CTA
    const uint32_t tmem_cols = a.N <= 32 ? 32 : a.N <= 64 ? 64 : 128;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_xhi);
        tma_prefetch_desc(&tm_xlo);
        tma_prefetch_desc(&tm_whi);
        tma_prefetch_desc(&tm_wlo);
        for (int s = 0; s < kTcStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kTcStages;
                const uint32_t ph = (kb / kTcStages) & 1;
                mbar_wait(&empty[s], ph ^ 1);                  // slot free (first round passes immediately)
                unsigned char* st = smem + (size_t)s * stage_bytes;
                float* xhi = reinterpret_cast<float*>(st);
                float* xlo = xhi + kTcBM * kTcBK;
                float* whi = xlo + kTcBM * kTcBK;
                float* wlo = whi + a.N * kTcBK;
                mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
                tma_load_2d(xhi, &tm_xhi, (kb0 + kb) * kTcBK, m0, &full[s]);
                tma_load_2d(xlo, &tm_xlo, (kb0 + kb) * kTcBK, m0, &full[s]);
                tma_load_2d(whi, &tm_whi, (kb0 + kb) * kTcBK, 0, &full[s]);
                tma_load_2d(wlo, &tm_wlo, (kb0 + kb) * kTcBK, 0, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc = umma_idesc_tf32(kTcBM, a.N);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kTcStages;
                const uint32_t ph = (kb / kTcStages) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                unsigned char* st = smem + (size_t)s * stage_bytes;
                const uint64_t d_xhi = umma_desc_sw128_addr(smem_u32(st));
                const uint64_t d_xlo = umma_desc_sw128_addr(smem_u32(st + (size_t)kTcBM * kTcBK * 4));
                const uint64_t d_whi = umma_desc_sw128_addr(smem_u32(st + (size_t)2 * kTcBM * kTcBK * 4));
                const uint64_t d_wlo = umma_desc_sw128_addr(smem_u32(st + (size_t)(2 * kTcBM + a.N) * kTcBK * 4));
#pragma unroll
                for (int k = 0; k < kTcBK / 8; ++k) {          // one MMA = K 8 floats = 32 bytes = 2 x 16 B
                    const uint64_t off = (uint64_t)(k * 2);
                    umma_tf32(tmem_base, d_xhi + off, d_whi + off, idesc, (kb | k) != 0);
                    umma_tf32(tmem_base, d_xlo + off, d_whi + off, idesc, 1);
                    umma_tf32(tmem_base, d_xhi + off, d_wlo + off, idesc, 1);
                }
                umma_commit(&empty[s]);                        // frees the smem slot when these MMAs retire
            }
            umma_commit(acc_full);                             // accumulator complete
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        float v[kTcMaxN];
#pragma unroll
        for (int c = 0; c < kTcMaxN / 32; ++c) {
            if (c * 32 < a.N) tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v + c * 32);
        }
        const int m = m0 + row;
        if (a.partial != nullptr) {
            if (m < a.M) {
                float4* dst = reinterpret_cast<float4*>(a.partial + ((size_t)blockIdx.y * a.m_pad + m) * a.N);
#pragma unroll
                for (int n4 = 0; n4 < kTcMaxN / 4; ++n4)
                    if (n4 * 4 < a.N) dst[n4] = make_float4(v[4 * n4], v[4 * n4 + 1], v[4 * n4 + 2], v[4 * n4 + 3]);
            }
        } else if (m < a.M) {
            const int N = a.N;
#pragma unroll
            for (int n = 0; n < kTcMaxN; ++n)
                if (n < N) v[n] += __ldg(a.bias + n);
            if (a.post == POST_LN_ACT) {
                float s = 0.f;
#pragma unroll
                for (int n = 0; n < kTcMaxN; ++n)
                    if (n < N) s += v[n];
                const float mu = s / (float)N;
                float var = 0.f;
#pragma unroll
                for (int n = 0; n < kTcMaxN; ++n)
                    if (n < N) { const float d = v[n] - mu; var = fmaf(d, d, var); }
                const float rstd = 1.0f / sqrtf(var / (float)N + 1e-5f);
#pragma unroll
                for (int n = 0; n < kTcMaxN; ++n)
                    if (n < N) v[n] = tc_act((v[n] - mu) * rstd * __ldg(a.ln_g + n) + __ldg(a.ln_b + n), a.act);
            } else if (a.post == POST_ACT) {
#pragma unroll
                for (int n = 0; n < kTcMaxN; ++n)
                    if (n < N) v[n] = tc_act(v[n], a.act);
            }
            float4* dst = reinterpret_cast<float4*>(a.out + (size_t)m * N);
#pragma unroll
            for (int n4 = 0; n4 < kTcMaxN / 4; ++n4)
                if (n4 * 4 < N) dst[n4] = make_float4(v[4 * n4], v[4 * n4 + 1], v[4 * n4 + 2], v[4 * n4 + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// x [rows][K] -> (hi, lo) [rows][Kp] with hi = RN_tf32(x) (low 13 mantissa bits zero), lo = RN_tf32(x - hi).
// Kp >= K is the K extent padded to a multiple of the 32-float K tile; the pad columns are zeroed once at
// allocation time and never written.  K must be a multiple of 4.
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo,
                                                         long long rows, int K, int Kp) {
    const int k4 = K / 4;
    const long long n4 = rows * k4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / k4;
        const int c = (int)(i - r * k4);
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        float4 h, l;
        h.x = round_tf32(v.x); h.y = round_tf32(v.y); h.z = round_tf32(v.z); h.w = round_tf32(v.w);
        l.x = round_tf32(v.x - h.x); l.y = round_tf32(v.y - h.y); l.z = round_tf32(v.z - h.z); l.w = round_tf32(v.w - h.w);
        reinterpret_cast<float4*>(hi + r * Kp)[c] = h;
        reinterpret_cast<float4*>(lo + r * Kp)[c] = l;
    }
}

// ---- host side: tensor maps through the driver entry point (no libcuda link dependency) -------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled get_tmap_encoder() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }
    return fn;
}

// 2-D row-major float matrix [rows][cols], box = box_rows x 32 floats, SWIZZLE_128B
inline bool make_tmap_2d(CUtensorMap* tm, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    PFN_tmapEncodeTiled enc = get_tmap_encoder();
    if (!enc) return false;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kTcBK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline bool tc_layer_eligible(int N, int K) { return N % 16 == 0 && N >= 16 && N <= kTcMaxN && K % 4 == 0 && K >= 512; }
inline int tc_padded_k(int K) { return (K + kTcBK - 1) / kTcBK * kTcBK; }

}  // namespace nww
