// nww_rowgemm.cuh — dense layers over many rows on the FP32 pipes with the register-tiled, shared-memory
// staged row GEMM of nww_tcn.cuh:  out[r][n] = post(bias[n] + sum_k A[r][k] W[k][n]).
// Used for the CRNN head's GRU input projections (x W_ih^T + b_ih over all 12 steps at once,
// reference CRNNModel architectures.py:242-282 / torch.nn.GRU) and, inside gru2_kernel, for the recurrent
// product h W_hh^T of every step.
#pragma once

#ifndef NWW_CPUSIM
#include <cooperative_groups.h>
#endif

#include <string.h>
#include <vector>

#include "nww_tc.cuh"
#include "nww_stage.cuh"
#include "nww_tcn.cuh"

namespace nww {

constexpr int kRgRows = 112;          // rows per CTA tile
inline size_t rowgemm_smem_bytes(int K) { return sizeof(float) * ((size_t)2 * kTcnWBuf + (size_t)kRgRows * K); }

// A [rows * a_row_mul + a_row_off][K] (row r of the GEMM reads row r * a_row_mul + a_row_off of A), W [K][N], out [rows][N]
__global__ void __launch_bounds__(kTcnNT, 1)
rowgemm_kernel(const float* __restrict__ A, long long a_row_mul, long long a_row_off, const float* __restrict__ W,
               const float* __restrict__ bias, float* __restrict__ out, long long rows, int K, int N) {
    NWW_DYN_SMEM(smem);
    float* wbuf = reinterpret_cast<float*>(smem);
    float* a_s = wbuf + 2 * kTcnWBuf;
    const int tid = threadIdx.x;
    const int k4 = K / 4;
    for (long long r0 = (long long)blockIdx.x * kRgRows; r0 < rows; r0 += (long long)gridDim.x * kRgRows) {
        const int nr = (int)((rows - r0 < kRgRows) ? (rows - r0) : kRgRows);
        __syncthreads();
        for (int i = tid; i < nr * k4; i += kTcnNT) {
            const int r = i / k4, c = i - r * k4;
            reinterpret_cast<float4*>(a_s)[i] = __ldg(reinterpret_cast<const float4*>(A + ((r0 + r) * a_row_mul + a_row_off) * K) + c);
        }
        __syncthreads();
        for (int n0 = 0; n0 < N; n0 += 128) {
            const int nc = (N - n0 < 128) ? (N - n0) : 128;
            tcn_layer_any(TcnLayerArgs{a_s, 0, 1, 0, K, K, W + n0, bias + n0, out + r0 * N + n0, 0, nc, nr, 1, 0, nullptr, 0, 0, 0, N, N, 0, 0},
                          wbuf, tid);
        }
    }
}

// ---------------------------------------------------------------------------------------
// GRU recurrence, 32 windows per CTA (gate order r, z, n; h0 = 0):
//   gi_f [B*S][3H] = x W_ih^T + b_ih (rowgemm_kernel),  whh [H][3H], bhh [3H]
//   r = sig(gi_r + gh_r)  z = sig(gi_z + gh_z)  n = tanh(gi_n + r * gh_n)  h = (1 - z) n + z h
// The reverse direction contributes its first step only (x_{S-1}, h0 = 0) to out[:, -1, :]; gi_b [B][3H].
// feat [B][2H] = [h_fwd(S-1) | h_bwd(first step)].  gh = h W_hh^T + b_hh is the row GEMM above per step.
// ---------------------------------------------------------------------------------------
constexpr int kGru2TM = 32;
inline size_t gru2_smem_bytes(int Hd) { return sizeof(float) * ((size_t)2 * kTcnWBuf + (size_t)kGru2TM * Hd * 7); }

__global__ void __launch_bounds__(kTcnNT, 1)
gru2_kernel(const float* __restrict__ gi_f, const float* __restrict__ gi_b, const float* __restrict__ whh,
            const float* __restrict__ bhh, const float* __restrict__ bhh_b, float* __restrict__ feat, long long B, int S, int Hd) {
    NWW_DYN_SMEM(smem);
    float* wbuf = reinterpret_cast<float*>(smem);
    float* h = wbuf + 2 * kTcnWBuf;                       // [TM][Hd]
    float* gh = h + (size_t)kGru2TM * Hd;                 // [TM][3Hd]
    float* gis = gh + (size_t)kGru2TM * 3 * Hd;           // [TM][3Hd] this step's input projections, fetched behind the GEMM
    const int tid = threadIdx.x;
    const int G = 3 * Hd;
    for (long long w0 = (long long)blockIdx.x * kGru2TM; w0 < B; w0 += (long long)gridDim.x * kGru2TM) {
        const int mt = (B - w0 < kGru2TM) ? (int)(B - w0) : kGru2TM;
        __syncthreads();
        for (int i = tid; i < kGru2TM * Hd; i += kTcnNT) h[i] = 0.0f;
        __syncthreads();
        for (int s = 0; s < S; ++s) {
            // gi of this step -> shared memory (cp.async; the row GEMM's own wait_group calls drain it)
            for (int i = tid; i < mt * (G / 4); i += kTcnNT) {
                const int m = i / (G / 4), c4 = i - m * (G / 4);
                tcn_cp_async16(gis + (size_t)m * G + 4 * c4, gi_f + ((w0 + m) * S + s) * (long long)G + 4 * c4);
            }
            tcn_cp_commit();
            for (int n0 = 0; n0 < G; n0 += 128) {
                const int nc = (G - n0 < 128) ? (G - n0) : 128;
                tcn_layer_any(TcnLayerArgs{h, 0, 1, 0, Hd, Hd, whh + n0, bhh + n0, gh + n0, 0, nc, mt, 1, 0, nullptr, 0, 0, 0, G, G, 0, 0},
                              wbuf, tid);
            }
            __syncthreads();
            for (int i = tid; i < mt * Hd; i += kTcnNT) {
                const int m = i / Hd, j = i - m * Hd;
                const float* gi = gis + (size_t)m * G;
                const float* g = gh + m * G;
                const float r = sigmoidf_acc(gi[j] + g[j]);
                const float z = sigmoidf_acc(gi[Hd + j] + g[Hd + j]);
                const float nn = tanhf(gi[2 * Hd + j] + r * g[2 * Hd + j]);
                h[m * Hd + j] = (1.0f - z) * nn + z * h[m * Hd + j];
            }
            __syncthreads();
        }
        for (int i = tid; i < mt * Hd; i += kTcnNT) {
            const int m = i / Hd, j = i - m * Hd;
            feat[(w0 + m) * (long long)(2 * Hd) + j] = h[m * Hd + j];
            const float* gi = gi_b + (w0 + m) * (long long)G;
            const float r = sigmoidf_acc(gi[j] + __ldg(bhh_b + j));
            const float z = sigmoidf_acc(gi[Hd + j] + __ldg(bhh_b + Hd + j));
            const float nn = tanhf(gi[2 * Hd + j] + r * __ldg(bhh_b + 2 * Hd + j));
            feat[(w0 + m) * (long long)(2 * Hd) + Hd + j] = (1.0f - z) * nn;
        }
    }
}

// ---------------------------------------------------------------------------------------
// GRU recurrence with the recurrent product on tcgen05 (default for H = 128).
// Same contract as gru2_kernel.  One CTA = 32 windows = rows 0..31 of a 128-row MMA tile (rows 32..127 of the A
// descriptor fall on whatever shared memory follows and produce accumulator rows nobody reads).  Per step:
//   h (FP32, shared) -> bf16 hi / lo un-swizzled K-major operand [K group 16][32 rows][8]   (LBO 512 B, SBO 128 B)
//   for each block of 128 gate columns: pre-split W_hh^T block (engine: [block][hi|lo][K group][128][8] bf16, 64 KB)
//     -> shared; 8 K steps x 3 split products of tcgen05.mma 128 x 128 x 16 -> TMEM columns [128 nb, +128)
//   TMEM lanes 0..31 -> gh in shared memory (warps 0 and 4), then the gate update by all 256 threads with this
//   step's input projections (prefetched by cp.async while the MMAs ran).
// ---------------------------------------------------------------------------------------
constexpr int kGruTcH = 128, kGruTcTM = 32, kGruTcNT = 256;
constexpr int kGruTcGhPitch = 3 * kGruTcH + 4;           // floats; +4 spreads the 32 row starts over the banks
constexpr size_t kGruTcABytes = (size_t)2 * 16 * 512;    // hi, lo x 16 K groups x (32 rows x 16 B)
constexpr size_t kGruTcBBytes = (size_t)2 * 16 * 128 * 16;   // hi, lo x 16 K groups x 128 columns x 16 B = 64 KB
inline size_t gru_tc_smem_bytes() {
    return kGruTcABytes + 4096 /* rows 32..127 of the last K group read past the operand */ + kGruTcBBytes +
           sizeof(float) * ((size_t)kGruTcTM * kGruTcH + (size_t)kGruTcTM * kGruTcGhPitch + (size_t)kGruTcTM * 3 * kGruTcH) + 128;
}

// host: w_hh (H, 3H) = [k][n] FP32 -> [block 3][hi|lo][K group 16][n 128][8 k] bf16
inline void gru_tc_pack_whh(const float* whh, std::vector<uint16_t>* out) {
    auto bf16_rn = [](float x) {
        uint32_t u;
        memcpy(&u, &x, 4);
        u += 0x7FFFu + ((u >> 16) & 1u);
        return (uint16_t)(u >> 16);
    };
    auto bf16_f = [](uint16_t b) {
        uint32_t u = (uint32_t)b << 16;
        float f;
        memcpy(&f, &u, 4);
        return f;
    };
    const int H = kGruTcH, G = 3 * H;
    const size_t op = (size_t)16 * 128 * 8;                 // elements per (block, hi|lo)
    out->assign((size_t)3 * 2 * op, 0);
    for (int k = 0; k < H; ++k)
        for (int n = 0; n < G; ++n) {
            const float v = whh[(size_t)k * G + n];
            const uint16_t hi = bf16_rn(v), lo = bf16_rn(v - bf16_f(hi));
            const int nb = n / 128, nn = n % 128;
            const size_t base = (size_t)nb * 2 * op + (size_t)(k >> 3) * 128 * 8 + (size_t)nn * 8 + (k & 7);
            (*out)[base] = hi;
            (*out)[base + op] = lo;
        }
}

__global__ void __launch_bounds__(kGruTcNT, 1)
gru_tc_kernel(const float* __restrict__ gi_f, const float* __restrict__ gi_b, const uint4* __restrict__ whh_q,
              const float* __restrict__ bhh, const float* __restrict__ bhh_b, float* __restrict__ feat, long long B, int S) {
    constexpr int H = kGruTcH, G = 3 * H, TM = kGruTcTM;
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* a_s = smem;                                            // [hi|lo][kg 16][32 rows][16 B]
    unsigned char* b_s = smem + kGruTcABytes + 4096;
    float* h = reinterpret_cast<float*>(b_s + kGruTcBBytes);              // [TM][H]
    float* gh = h + TM * H;                                               // [TM][kGruTcGhPitch]
    float* gis = gh + TM * kGruTcGhPitch;                                 // [TM][3H]
    uint64_t* bar = reinterpret_cast<uint64_t*>(gis + TM * G);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint64_t da_hi = umma_desc_noswz(smem_u32(a_s), 512, 128);
    const uint64_t da_lo = da_hi + (uint64_t)((16 * 512) / 16);
    const uint64_t db_hi = umma_desc_noswz(smem_u32(b_s), 128 * 16, 128);
    const uint64_t db_lo = db_hi + (uint64_t)((16 * 128 * 16) / 16);
    const uint32_t idesc = umma_idesc_bf16(128, 128);
    uint32_t phase = 0;

    for (long long w0 = (long long)blockIdx.x * TM; w0 < B; w0 += (long long)gridDim.x * TM) {
        const int mt = (B - w0 < TM) ? (int)(B - w0) : TM;
        __syncthreads();
        for (int i = tid; i < TM * H; i += kGruTcNT) h[i] = 0.0f;
        __syncthreads();
        for (int s = 0; s < S; ++s) {
            // this step's input projections -> shared memory, in flight during the MMAs
            for (int i = tid; i < mt * (G / 4); i += kGruTcNT) {
                const int m = i / (G / 4), c4 = i - m * (G / 4);
                tcn_cp_async16(gis + (size_t)m * G + 4 * c4, gi_f + ((w0 + m) * S + s) * (long long)G + 4 * c4);
            }
            tcn_cp_commit();
            // h -> bf16 hi / lo operand: thread = (row m, K group g)
            for (int i = tid; i < TM * 16; i += kGruTcNT) {
                const int g = i / TM, m = i - g * TM;
                const float4 v0 = *reinterpret_cast<const float4*>(h + m * H + 8 * g);
                const float4 v1 = *reinterpret_cast<const float4*>(h + m * H + 8 * g + 4);
                const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                uint32_t hb[8], lb[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    hb[k] = float_to_bf16_bits(v[k]);
                    lb[k] = float_to_bf16_bits(v[k] - bf16_bits_to_float(hb[k]));
                }
                unsigned char* dst = a_s + (size_t)g * 512 + (size_t)m * 16;
                *reinterpret_cast<uint4*>(dst) = make_uint4(hb[0] | (hb[1] << 16), hb[2] | (hb[3] << 16), hb[4] | (hb[5] << 16), hb[6] | (hb[7] << 16));
                *reinterpret_cast<uint4*>(dst + 16 * 512) =
                    make_uint4(lb[0] | (lb[1] << 16), lb[2] | (lb[3] << 16), lb[4] | (lb[5] << 16), lb[6] | (lb[7] << 16));
            }
            for (int nb = 0; nb < 3; ++nb) {
                // W_hh^T block nb (64 KB, L2-resident) -> shared
                const uint4* src = whh_q + (size_t)nb * (kGruTcBBytes / 16);
                for (int i = tid; i < (int)(kGruTcBBytes / 16); i += kGruTcNT) reinterpret_cast<uint4*>(b_s)[i] = __ldg(src + i);
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(nb * 128);
#pragma unroll
                    for (int ks = 0; ks < H / 16; ++ks) {                 // 16 K = 2 K groups per MMA
                        const uint64_t ao = (uint64_t)((2 * ks * 512) / 16), bo = (uint64_t)((2 * ks * 128 * 16) / 16);
                        umma_bf16(d_tmem, da_hi + ao, db_hi + bo, idesc, ks != 0);
                        umma_bf16(d_tmem, da_lo + ao, db_hi + bo, idesc, 1);
                        umma_bf16(d_tmem, da_hi + ao, db_lo + bo, idesc, 1);
                    }
                    umma_commit(bar);
                }
                mbar_wait(bar, phase);                                    // the block's MMAs are done: b_s can be refilled
                phase ^= 1;
                tc_fence_after();
            }
            // TMEM lanes 0..31 (the 32 windows) -> gh; warps 0 and 4 own lane quarter 0, six 32-column loads each
            if ((warp & 3) == 0) {
                const int half = warp >> 2;
                for (int c = half * 6; c < half * 6 + 6; ++c) {
                    float v[32];
                    tmem_ld_32x32b_x32(tmem_base + (uint32_t)(c * 32), v);
                    float4* dst = reinterpret_cast<float4*>(gh + (size_t)lane * kGruTcGhPitch + c * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
            tcn_cp_wait<0>();
            tc_fence_before();
            __syncthreads();
            for (int i = tid; i < mt * H; i += kGruTcNT) {
                const int m = i / H, j = i - m * H;
                const float* gi = gis + (size_t)m * G;
                const float* g = gh + (size_t)m * kGruTcGhPitch;
                const float r = sigmoidf_acc(gi[j] + g[j] + __ldg(bhh + j));
                const float z = sigmoidf_acc(gi[H + j] + g[H + j] + __ldg(bhh + H + j));
                const float nn = tanhf(gi[2 * H + j] + r * (g[2 * H + j] + __ldg(bhh + 2 * H + j)));
                h[m * H + j] = (1.0f - z) * nn + z * h[m * H + j];
            }
            __syncthreads();
        }
        for (int i = tid; i < mt * H; i += kGruTcNT) {
            const int m = i / H, j = i - m * H;
            feat[(w0 + m) * (long long)(2 * H) + j] = h[m * H + j];
            const float* gi = gi_b + (w0 + m) * (long long)G;
            const float r = sigmoidf_acc(gi[j] + __ldg(bhh_b + j));
            const float z = sigmoidf_acc(gi[H + j] + __ldg(bhh_b + H + j));
            const float nn = tanhf(gi[2 * H + j] + r * __ldg(bhh_b + 2 * H + j));
            feat[(w0 + m) * (long long)(2 * H) + H + j] = (1.0f - z) * nn;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------
// Row GEMM on tcgen05: out[r][n] = bias[n] + sum_k A[r * a_row_mul + a_row_off][k] W[k][n]   (K % 16 == 0, N % 64 == 0)
// Tiles of 128 rows: FP32 rows -> bf16 hi / lo un-swizzled K-major operand; weights pre-split by the engine in
// 64-column chunks ([chunk][hi|lo][K group][64][8] bf16); (K / 16) x 3 tcgen05.mma 128 x 64 x 16 per chunk.
// Used for the GRU input projections (K = 160, N = 384).
// ---------------------------------------------------------------------------------------
constexpr int kRuRows = 128, kRuNC = 64, kRuNT = 256;
inline size_t rowgemm_umma_smem_bytes(int K) { return (size_t)2 * kRuRows * K * 2 + (size_t)2 * kRuNC * K * 2 + 128; }

inline void rowgemm_umma_pack(const float* w /* [K][N] */, int K, int N, std::vector<uint16_t>* out) {
    auto bf16_rn = [](float x) {
        uint32_t u;
        memcpy(&u, &x, 4);
        u += 0x7FFFu + ((u >> 16) & 1u);
        return (uint16_t)(u >> 16);
    };
    auto bf16_f = [](uint16_t b) {
        uint32_t u = (uint32_t)b << 16;
        float f;
        memcpy(&f, &u, 4);
        return f;
    };
    const size_t op = (size_t)(K / 8) * kRuNC * 8;
    out->assign((size_t)(N / kRuNC) * 2 * op, 0);
    for (int k = 0; k < K; ++k)
        for (int n = 0; n < N; ++n) {
            const float v = w[(size_t)k * N + n];
            const uint16_t hi = bf16_rn(v), lo = bf16_rn(v - bf16_f(hi));
            const size_t base = (size_t)(n / kRuNC) * 2 * op + (size_t)(k >> 3) * kRuNC * 8 + (size_t)(n % kRuNC) * 8 + (k & 7);
            (*out)[base] = hi;
            (*out)[base + op] = lo;
        }
}

__global__ void __launch_bounds__(kRuNT, 1)
rowgemm_umma_kernel(const float* __restrict__ A, long long a_row_mul, long long a_row_off, const uint4* __restrict__ wq,
                    const float* __restrict__ bias, float* __restrict__ out, long long rows, int K, int N) {
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kg_n = K / 8;
    const int a_op = kRuRows * K * 2, b_op = kRuNC * K * 2;          // bytes of one (hi | lo) operand
    unsigned char* a_s = smem;
    unsigned char* b_s = smem + 2 * a_op;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * a_op + 2 * b_op);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = umma_idesc_bf16(128, kRuNC);
    const uint64_t da_h0 = umma_desc_noswz(smem_u32(a_s), kRuRows * 16, 128), da_l0 = da_h0 + (uint64_t)(a_op >> 4);
    const uint64_t db_h0 = umma_desc_noswz(smem_u32(b_s), kRuNC * 16, 128), db_l0 = db_h0 + (uint64_t)(b_op >> 4);
    uint32_t phase = 0;
    for (long long r0 = (long long)blockIdx.x * kRuRows; r0 < rows; r0 += (long long)gridDim.x * kRuRows) {
        for (int i = tid; i < kRuRows * kg_n; i += kRuNT) {
            const int g = i / kRuRows, r = i - g * kRuRows;
            uint4 hv = make_uint4(0, 0, 0, 0), lv = hv;
            if (r0 + r < rows) {
                const float4* p = reinterpret_cast<const float4*>(A + ((r0 + r) * a_row_mul + a_row_off) * K + 8 * g);
                const float4 v0 = __ldg(p), v1 = __ldg(p + 1);
                const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                uint32_t h[8], l[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    h[e] = float_to_bf16_bits(v[e]);
                    l[e] = float_to_bf16_bits(v[e] - bf16_bits_to_float(h[e]));
                }
                hv = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
                lv = make_uint4(l[0] | (l[1] << 16), l[2] | (l[3] << 16), l[4] | (l[5] << 16), l[6] | (l[7] << 16));
            }
            const int off = (g * kRuRows + r) * 16;
            *reinterpret_cast<uint4*>(a_s + off) = hv;
            *reinterpret_cast<uint4*>(a_s + a_op + off) = lv;
        }
        for (int nc = 0; nc < N / kRuNC; ++nc) {
            const uint4* src = wq + (size_t)nc * (2 * b_op / 16);
            for (int i = tid; i < 2 * b_op / 16; i += kRuNT) reinterpret_cast<uint4*>(b_s)[i] = __ldg(src + i);
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                for (int ks = 0; ks < K / 16; ++ks) {
                    const uint64_t ao = (uint64_t)((2 * ks * kRuRows * 16) >> 4), bo = (uint64_t)((2 * ks * kRuNC * 16) >> 4);
                    umma_bf16(tmem_base, da_h0 + ao, db_h0 + bo, idesc, ks != 0);
                    umma_bf16(tmem_base, da_l0 + ao, db_h0 + bo, idesc, 1);
                    umma_bf16(tmem_base, da_h0 + ao, db_l0 + bo, idesc, 1);
                }
                umma_commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
            tc_fence_after();
            {
                const int q = warp & 3, hcol = warp >> 2;
                float v[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hcol * 32), v);
                tc_fence_before();
                const long long r = r0 + q * 32 + lane;
                if (r < rows) {
                    const int n0 = nc * kRuNC + hcol * 32;
                    float4* dst = reinterpret_cast<float4*>(out + r * N + n0);
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        dst[j4] = make_float4(v[4 * j4] + __ldg(bias + n0 + 4 * j4), v[4 * j4 + 1] + __ldg(bias + n0 + 4 * j4 + 1),
                                              v[4 * j4 + 2] + __ldg(bias + n0 + 4 * j4 + 2), v[4 * j4 + 3] + __ldg(bias + n0 + 4 * j4 + 3));
                }
            }
            __syncthreads();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 64);
    }
}

// ---------------------------------------------------------------------------------------
// Row GEMM on tcgen05 for long K and wide N (QuartzNet's pointwise + residual 1x1 convolutions, K up to 512,
// N up to 512):  out[r][n] = act(bias[n] + sum_k A[r][k] W[k][n] + res[r][n])   (K % 64 == 0, N % 64 == 0, N <= 512)
// Tiles of 128 rows; K is walked in chunks of 64: the FP32 rows of a chunk become bf16 hi / lo un-swizzled K-major
// operands in one of two shared-memory buffers (the conversion of chunk c + 1 overlaps the MMAs of chunk c), the
// weights stream from L2 in 16 KB sub-blocks (chunk, 64 columns, hi | lo) through a ring of four slots filled by
// cp.async.bulk, 4 K steps x 3 products of tcgen05.mma 128 x 64 x 16 per sub-block accumulate into TMEM columns
// [64 nc, +64) across all chunks; one epilogue per tile.
// Round 2: the operand producer reads 8 rows x 128 contiguous bytes per warp instruction with the K-group offsets of a
// row tabulated once per launch; a ninth warp issues the weight copies and the MMAs; layers of up to 256 columns run two
// CTAs per SM, wider ones as 256-column slices (n_off / n_pitch) of the same launch shape.
// ---------------------------------------------------------------------------------------
constexpr int kKcRows = 128, kKcNT = 256, kKcKC = 64, kKcNC = 64, kKcRing = 4;
// kKcNT threads convert operands and run the epilogue; ONE MORE WARP issues the weight copies and the MMAs (its lane 0): the
// issuer waits for weights to land and for MMAs to free ring slots, and as warp 0 it held up its own share of the next chunk's
// conversion — and with it every CTA barrier.
constexpr int kKcBlock = kKcNT + 32, kKcIssuer = kKcNT;
constexpr int kKcABuf = 2 * (kKcKC / 8) * kKcRows * 16;           // hi | lo of one chunk: 32 KB
constexpr int kKcSub = 2 * (kKcKC / 8) * kKcNC * 16;              // one weight sub-block: 16 KB
constexpr int kKcKoff = 256;                                       // K groups (of 8 columns) whose offsets are tabulated: K <= 2048
inline size_t rowgemm_kc_smem_bytes(int ring = kKcRing) {
    return (size_t)2 * kKcABuf + (size_t)ring * kKcSub + 256 + (kKcRows + kKcKoff) * sizeof(long long);
}
// Outputs up to 256 columns leave a tile little MMA work between its operand conversion and its epilogue: a two-slot
// weight ring makes the CTA small enough (101 KB, <= 256 TMEM columns, 128 registers) for two CTAs per SM to overlap
// those phases (N <= 128: +39 % on the raw-audio models; N = 256: +4 % on QuartzNet).
struct KcLaunch { int ring, grid; size_t smem; };
inline KcLaunch rowgemm_kc_launch(long long rows, int N, int sm_count) {
    const int ring = N <= 256 ? 2 : kKcRing;                          // (N <= 256: two CTAs share the SM's 512 TMEM columns)
    const long long tiles = (rows + kKcRows - 1) / kKcRows;
    const long long cap = (long long)sm_count * (N <= 256 ? 2 : 1);
    return KcLaunch{ring, (int)(tiles < cap ? tiles : cap), rowgemm_kc_smem_bytes(ring)};
}

// host: w [K][N] FP32 -> the stream the kernel consumes: for every K chunk, for every 64-column group:
// [hi | lo][K group 8][64 columns][8 k] bf16
inline void rowgemm_kc_pack(const float* w, int K, int N, std::vector<uint16_t>* out) {
    auto bf16_rn = [](float x) {
        uint32_t u;
        memcpy(&u, &x, 4);
        u += 0x7FFFu + ((u >> 16) & 1u);
        return (uint16_t)(u >> 16);
    };
    auto bf16_f = [](uint16_t b) {
        uint32_t u = (uint32_t)b << 16;
        float f;
        memcpy(&f, &u, 4);
        return f;
    };
    const size_t plane = (size_t)(kKcKC / 8) * kKcNC * 8, sub = 2 * plane;
    const int n_nc = N / kKcNC;
    out->assign((size_t)(K / kKcKC) * n_nc * sub, 0);
    for (int k = 0; k < K; ++k)
        for (int n = 0; n < N; ++n) {
            const float v = w[(size_t)k * N + n];
            const uint16_t hi = bf16_rn(v), lo = bf16_rn(v - bf16_f(hi));
            const int kk = k % kKcKC;
            const size_t base = ((size_t)(k / kKcKC) * n_nc + (size_t)(n / kKcNC)) * sub + (size_t)(kk >> 3) * kKcNC * 8 +
                                (size_t)(n % kKcNC) * 8 + (kk & 7);
            (*out)[base] = hi;
            (*out)[base + plane] = lo;
        }
}

// Where GEMM row r lives: rows are grouped per window (rpw rows each) and, inside a window, in lines of `inner` rows
// (an image row of a strided Conv2d; inner = rpw for sequences).  Row r is at
//   base + (r / rpw) * win_stride + (t / inner) * outer_stride + (t % inner + row_off) * row_stride,  t = r % rpw   (floats)
// A plain [rows][pitch] matrix is kc_plain(rows, pitch).  For strided convolutions on channel-last buffers the A rows
// OVERLAP (row_stride = conv stride * C_in < K) and the output view skips the next layer's zero padding.
struct KcView {
    long long rpw, win_stride, row_stride, row_off, inner, outer_stride;
    __device__ __forceinline__ long long at(long long r) const {
        const long long w = r / rpw, t = r - w * rpw, o = t / inner;
        return w * win_stride + o * outer_stride + (t - o * inner + row_off) * row_stride;
    }
};
inline KcView kc_plain(long long rows, long long pitch) { return KcView{rows, 0, pitch, 0, rows, 0}; }
inline KcView kc_seq(long long rpw, long long win_stride, long long row_stride, long long row_off) {
    return KcView{rpw, win_stride, row_stride, row_off, rpw, 0};
}
// The K axis of an A row may be cut into segments of seg_len floats (a multiple of 8) that lie seg_stride apart: the three
// kernel rows of a 3x3 convolution on an NHWC image (seg_len = 3 C_in, seg_stride = one padded image row).  Columns
// k >= k_valid (the padding of K to a multiple of 64) read as zero.
struct KcSegs { int seg_len; long long seg_stride; int k_valid; };
inline KcSegs kc_one_seg(int k_valid) { return KcSegs{1 << 30, 0, k_valid}; }

// n_valid (a multiple of 4, <= N): columns actually stored (the weight matrix is padded to 64 columns)
// act: 0 none, 1 ReLU, 2 GELU, 3 SiLU (apply_act codes + 1)
// VIEWS = false is the fast path for plain matrices (A = [rows][K], out = [rows][N], all columns stored, act 0 / 1):
// no per-row address table, no segment arithmetic, no column masks.
// Read path of the operand / residual loads: the read-only (non-coherent) path for a launch per layer; plain cached loads
// inside the fused multi-layer kernel, where a layer reads what OTHER SMs wrote earlier in the same launch.  That is safe
// because every activation buffer of the cone is written exactly once per launch, before its first read (the layer
// program gives each level its own residual buffer), so no SM can hold a stale L1 line; L2-only loads (ld.global.cg)
// were measured 40 % slower — the rows of a layer overlap (three taps) and want L1.
template <bool COHERENT> __device__ __forceinline__ float4 kc_ld(const float4* p) {
#ifndef NWW_CPUSIM
    if (COHERENT) return *p;
#endif
    return __ldg(p);
}

// GELU / SiLU of four values, out of line: one copy of the erf / exp code per kernel instead of one per unrolled call site
__device__ __noinline__ float4 kc_act4(float4 o, int act) {
    return make_float4(apply_act(o.x, act), apply_act(o.y, act), apply_act(o.z, act), apply_act(o.w, act));
}

// mbarrier / ring bookkeeping that lives across tiles (and, in the fused kernel, across layers)
struct KcState { uint32_t c_slot = 0, c_par = 0, p_slot = 0, done_phase = 0, g = 0; };

// All tiles of one row GEMM for this CTA (see rowgemm_kc_umma_kernel for the arguments); shared-memory carve-up, barriers
// and TMEM were set up by the caller.
template <bool VIEWS, bool COHERENT>
__device__ __forceinline__ void rowgemm_kc_run(const float* __restrict__ A, const KcView& av, const KcSegs& sg, int K,
                                               const uint4* __restrict__ wq, const float* __restrict__ bias,
                                               const float* __restrict__ res, float* __restrict__ out, const KcView& ov, long long rows,
                                               int N, int n_valid, int act, int ring, const KcView& rv, int pre_relu, KcState& S,
                                               int n_off, int n_pitch /* this launch computes columns [n_off, n_off + N) of n_pitch */,
                                               unsigned char* a_s, unsigned char* b_s, uint64_t* bar_full, uint64_t* bar_empty,
                                               uint64_t* bar_afree, uint64_t* bar_done, long long* row_at, uint32_t tmem_base) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t idesc = umma_idesc_bf16(128, kKcNC);
    const uint64_t da_buf0 = umma_desc_noswz(smem_u32(a_s), kKcRows * 16, 128);
    const uint64_t db_ring = umma_desc_noswz(smem_u32(b_s), kKcNC * 16, 128);
    const int n_kc = K / kKcKC, n_nc = N / kKcNC, total = n_kc * n_nc;
    const int nc_total = n_pitch / kKcNC, nc_off = n_off / kKcNC;      // the weight stream holds every 64-column group of every K chunk
    uint32_t& c_slot = S.c_slot; uint32_t& c_par = S.c_par; uint32_t& p_slot = S.p_slot; uint32_t& done_phase = S.done_phase;
    uint32_t& g = S.g;                                                 // chunks converted so far: buffer g & 1
    // where K group k8 / 8 of an A row starts relative to the row (segments resolved once per launch instead of a
    // division per cell: the operand producer spent a quarter of its instructions on this address arithmetic); -1 = all padding
    long long* koff = row_at + kKcRows;
    const bool tab = VIEWS && K / 8 <= kKcKoff;
    if (tab) {
        __syncthreads();                                               // (fused kernel) the previous layer's last conversion has read the table
        for (int g8 = tid; g8 < K / 8; g8 += kKcBlock) {
            const int k8 = 8 * g8, seg = k8 / sg.seg_len;
            koff[g8] = k8 < sg.k_valid ? seg * sg.seg_stride + (k8 - seg * sg.seg_len) : -1;
        }
    }
    for (long long r0 = (long long)blockIdx.x * kKcRows; r0 < rows; r0 += (long long)gridDim.x * kKcRows) {   // persistent over the launch's tiles
        int p_pos = 0, prev_slot = -1;
        uint32_t prev_par = 0;
        auto produce = [&]() {                                         // thread 0: next sub-block of the tile's stream
            mbar_expect_tx(bar_full + p_slot, kKcSub);
            bulk_g2s(b_s + (size_t)p_slot * kKcSub,
                     reinterpret_cast<const unsigned char*>(wq) + (size_t)((p_pos / n_nc) * nc_total + nc_off + p_pos % n_nc) * kKcSub, kKcSub,
                     bar_full + p_slot);
            p_slot = (int)p_slot + 1 == ring ? 0 : p_slot + 1;
            ++p_pos;
        };
        if (tid == kKcIssuer)
            for (int i = 0; i < ring && p_pos < total; ++i) produce();
        if (VIEWS) {
            __syncthreads();                                           // the previous tile's last conversion has read row_at
            if (tid < kKcRows) row_at[tid] = r0 + tid < rows ? av.at(r0 + tid) : -1;
            __syncthreads();
        }
        for (int kc = 0; kc < n_kc; ++kc, ++g) {
            const uint32_t buf = g & 1u;
            if (g >= 2) mbar_wait(bar_afree + buf, ((g >> 1) - 1u) & 1u);   // chunk g - 2's MMAs have read this buffer
            unsigned char* ab = a_s + (size_t)buf * kKcABuf;
            // every thread owns kPer (row, K group) cells of the chunk: ALL their global loads are issued before the first
            // conversion, so a thread has 2 * kPer 128-bit loads in flight instead of 2 (the operand producer was
            // latency-bound on L2: profiles/r02_rowgemm_tcn_before.txt, long-scoreboard 57 % of the samples)
            constexpr int kPer = kKcRows * (kKcKC / 8) / kKcNT;
            if (tid < kKcNT) {
            float4 v0[kPer], v1[kPer];
            int k8s[kPer];
#pragma unroll
            for (int j = 0; j < kPer; ++j) {
                // a warp instruction covers 8 rows x 4 K groups (128 contiguous bytes per row: 8 cache lines per request; one
                // row per lane was 32 lines per request and the L1 tag stage showed it) at the price of 4-way conflicts on the
                // operand stores, which the shared-memory pipe (8 % busy) does not notice
                const int cw = warp + (kKcNT / 32) * j;
                const int gq = (cw & 1) * 4 + (lane & 3), r = (cw >> 1) * 8 + (lane >> 2);
                const int k8 = kc * kKcKC + 8 * gq;
                const long long ra = VIEWS ? row_at[r] : (r0 + r < rows ? (r0 + r) * (long long)K : -1);
                k8s[j] = -1;
                v0[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                v1[j] = v0[j];
                long long ko = k8;
                if (VIEWS) {
                    if (tab) {
                        ko = koff[k8 >> 3];
                    } else {
                        const int seg = k8 / sg.seg_len;
                        ko = k8 < sg.k_valid ? seg * sg.seg_stride + (k8 - seg * sg.seg_len) : -1;
                    }
                }
                if (ra >= 0 && ko >= 0) {
                    const float4* p = reinterpret_cast<const float4*>(A + ra + ko);
                    v0[j] = kc_ld<COHERENT>(p);
                    v1[j] = kc_ld<COHERENT>(p + 1);
                    k8s[j] = k8;
                }
            }
#pragma unroll
            for (int j = 0; j < kPer; ++j) {
                const int cw = warp + (kKcNT / 32) * j;
                const int i = ((cw & 1) * 4 + (lane & 3)) * kKcRows + (cw >> 1) * 8 + (lane >> 2);     // [K group][row]
                uint4 hv = make_uint4(0, 0, 0, 0), lv = hv;
                if (k8s[j] >= 0) {
                    float v[8] = {v0[j].x, v0[j].y, v0[j].z, v0[j].w, v1[j].x, v1[j].y, v1[j].z, v1[j].w};
                    if (VIEWS && k8s[j] + 8 > sg.k_valid) {            // the group straddles the end of the valid columns
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            if (k8s[j] + e >= sg.k_valid) v[e] = 0.0f;
                    }
                    uint32_t h[8], l[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        h[e] = float_to_bf16_bits(v[e]);
                        l[e] = float_to_bf16_bits(v[e] - bf16_bits_to_float(h[e]));
                    }
                    hv = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
                    lv = make_uint4(l[0] | (l[1] << 16), l[2] | (l[3] << 16), l[4] | (l[5] << 16), l[6] | (l[7] << 16));
                }
                *reinterpret_cast<uint4*>(ab + i * 16) = hv;
                *reinterpret_cast<uint4*>(ab + kKcABuf / 2 + i * 16) = lv;
            }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == kKcIssuer) {
                tc_fence_after();
                const uint64_t da_h = da_buf0 + (uint64_t)((buf * (uint32_t)kKcABuf) >> 4), da_l = da_h + (uint64_t)((kKcABuf / 2) >> 4);
                for (int nc = 0; nc < n_nc; ++nc) {
                    mbar_wait_one(bar_full + c_slot, c_par);
                    tc_fence_after();
                    const uint64_t db_h = db_ring + (uint64_t)((c_slot * (uint32_t)kKcSub) >> 4), db_l = db_h + (uint64_t)((kKcSub / 2) >> 4);
                    const uint32_t d = tmem_base + (uint32_t)(nc * kKcNC);
#pragma unroll
                    for (int ks = 0; ks < kKcKC / 16; ++ks) {
                        const uint64_t ao = (uint64_t)((2 * ks * kKcRows * 16) >> 4), bo = (uint64_t)((2 * ks * kKcNC * 16) >> 4);
                        umma_bf16(d, da_h + ao, db_h + bo, idesc, (kc | ks) != 0);
                        umma_bf16(d, da_l + ao, db_h + bo, idesc, 1);
                        umma_bf16(d, da_h + ao, db_l + bo, idesc, 1);
                    }
                    umma_commit(bar_empty + c_slot);
                    if (prev_slot >= 0) {                              // refill the previous sub-block's slot
                        mbar_wait_one(bar_empty + prev_slot, prev_par);
                        if (p_pos < total) produce();
                    }
                    prev_slot = (int)c_slot;
                    prev_par = c_par;
                    if ((int)++c_slot == ring) { c_slot = 0; c_par ^= 1u; }
                }
                umma_commit(bar_afree + buf);
            }
        }
        if (tid == kKcIssuer) {
            umma_commit(bar_done);
            mbar_wait_one(bar_empty + prev_slot, prev_par);
        }
        mbar_wait(bar_done, done_phase);
        done_phase ^= 1u;
        tc_fence_after();
        if (tid < kKcNT) {
            const int q = warp & 3, hcol = warp >> 2;
            const long long r = r0 + q * 32 + lane;
            const long long o_at = r < rows ? (VIEWS ? ov.at(r) : r * (long long)n_pitch) + n_off : 0;
            for (int c0 = hcol * (N / 2); c0 < (hcol + 1) * (N / 2) && c0 < n_valid; c0 += 32) {
                float v[32];
                float4 bq[8];                                          // the bias of these 32 columns: eight 128-bit loads in flight
#pragma unroll                                                         // behind the TMEM load (32 scalar loads stalled the epilogue)
                for (int j4 = 0; j4 < 8; ++j4)
                    bq[j4] = c0 + 4 * j4 < n_valid ? __ldg(reinterpret_cast<const float4*>(bias + n_off + c0) + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                if (r < rows) {
                    float4* dst = reinterpret_cast<float4*>(out + o_at + c0);
                    const float4* rs = res ? reinterpret_cast<const float4*>(res + (VIEWS && rv.rpw ? rv.at(r) : r * (long long)n_pitch) + n_off + c0) : nullptr;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        if (VIEWS && c0 + 4 * j4 >= n_valid) break;
                        float4 o = make_float4(v[4 * j4] + bq[j4].x, v[4 * j4 + 1] + bq[j4].y, v[4 * j4 + 2] + bq[j4].z, v[4 * j4 + 3] + bq[j4].w);
                        if (VIEWS && pre_relu) { o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f); }
                        if (rs) {
                            const float4 t = kc_ld<COHERENT>(rs + j4);
                            o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
                        }
                        if (!VIEWS) {
                            if (act) { o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f); }
                        } else if (act == 1) {                         // ReLU: kept apart so that the unrolled loop does not carry
                            o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f);   // eight copies of erf / exp
                        } else if (act) {
                            o = kc_act4(o, act - 1);
                        }
                        dst[j4] = o;
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
    }
}

template <bool VIEWS>
__global__ void __launch_bounds__(kKcBlock, 2)
rowgemm_kc_umma_kernel(const float* __restrict__ A, KcView av, KcSegs sg, int K, const uint4* __restrict__ wq,
                       const float* __restrict__ bias, const float* __restrict__ res, float* __restrict__ out, KcView ov, long long rows,
                       int N, int n_valid, int act, int ring,
                       KcView rv = KcView{0, 0, 0, 0, 0, 0} /* where the residual row of GEMM row r lives; rpw == 0: res + r * N */,
                       int pre_relu = 0 /* ReLU before the residual is added: TemporalBlock's relu(relu(conv2) + res) */,
                       int n_off = 0, int n_pitch = 0 /* columns [n_off, n_off + N) of an n_pitch-wide layer (0: N); plain matrices only */) {
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* a_s = smem;                                         // two chunk buffers
    unsigned char* b_s = smem + 2 * kKcABuf;                           // weight ring
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(b_s + ring * kKcSub);
    uint64_t* bar_empty = bar_full + kKcRing;
    uint64_t* bar_afree = bar_full + 2 * kKcRing;                      // [2]: the MMAs that read chunk buffer b are done
    uint64_t* bar_done = bar_full + 2 * kKcRing + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 2 * kKcRing + 3);
    long long* row_at = reinterpret_cast<long long*>(b_s + ring * kKcSub + 256);   // A offsets of the tile's rows
    const uint32_t tmem_cols = N <= 64 ? 64u : N <= 128 ? 128u : N <= 256 ? 256u : 512u;
    if (tid == 0) {
        for (int i = 0; i < 2 * kKcRing + 3; ++i) mbar_init(bar_full + i, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    KcState S;
    rowgemm_kc_run<VIEWS, false>(A, av, sg, K, wq, bias, res, out, ov, rows, N, n_valid, act, ring, rv, pre_relu, S, n_off, n_pitch ? n_pitch : N, a_s, b_s, bar_full,
                                 bar_empty, bar_afree, bar_done, row_at, tmem_base);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

#ifndef NWW_CPUSIM
// ---------------------------------------------------------------------------------------
// The TCN dependency cone (nww_tcn.cuh) as ONE cooperative launch: every layer is the row GEMM above over all windows
// of the launch group, with a grid-wide barrier between layers (layer l + 1 reads what other CTAs wrote in layer l, so
// its operand / residual loads go to L2: kc_ld<true>).  In stream mode the launch starts with the gather of the cone's
// frames out of the mel ring (stream_mel_tail_kernel's job: one warp per stream, transposed through shared memory).
// Replaces 9 launches per push (tail gather + 8 layers) by one; same tiles, same order: bit-identical.
// ---------------------------------------------------------------------------------------
struct TcnFusedLayer {
    const float* A; KcView av; int k_valid, K;
    const uint4* wq; const float* bias;
    const float* res; KcView rv;
    float* out; KcView ov;
    long long rows; int N, act, pre_relu;
};
constexpr int kTcnFusedMaxLayers = 9;
struct TcnFusedParams {
    int n_layers;
    TcnFusedLayer L[kTcnFusedMaxLayers];
    MelRingRef ring;            // ring.ring != nullptr: gather first
    long long n;                // windows (streams) of the launch group
    float* mel;                 // time-major (n, T, F) log-mel buffer the first layer reads
    int t0, n_tail;
};

__global__ void __maxnreg__(96)          // two CTAs of 256 threads per SM with room to spare (registers come in groups of four warps)
tcn_rows_fused_kernel(const __grid_constant__ TcnFusedParams P) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int ring = 2;
    unsigned char* a_s = smem;
    unsigned char* b_s = smem + 2 * kKcABuf;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(b_s + ring * kKcSub);
    uint64_t* bar_empty = bar_full + kKcRing;
    uint64_t* bar_afree = bar_full + 2 * kKcRing;
    uint64_t* bar_done = bar_full + 2 * kKcRing + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 2 * kKcRing + 3);
    long long* row_at = reinterpret_cast<long long*>(b_s + ring * kKcSub + 256);
    constexpr uint32_t tmem_cols = 128;                                // every layer of the cone has N <= 128
    if (tid == 0) {
        for (int i = 0; i < 2 * kKcRing + 3; ++i) mbar_init(bar_full + i, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (P.ring.ring != nullptr) {
        // the cone's n_tail frames of every stream, time-major, where the batch front end would have written them
        float* tile = reinterpret_cast<float*>(a_s) + (size_t)warp * (kMelTailMax * (SMel::F + 1));
        const int nw = kKcNT / 32;
        for (long long w = (long long)blockIdx.x * nw + warp; warp < nw && w < P.n; w += (long long)gridDim.x * nw) {
            const long long s = P.ring.stream(w);
            const int head = smel_slot(P.ring.count[s] / SMel::HOP - 3 + 1);
            const float* src = P.ring.ring + s * SMel::STREAM_FLOATS + head + P.t0;
            if (lane < P.n_tail)
                for (int m = 0; m < SMel::F; ++m) tile[lane * (SMel::F + 1) + m] = src[m * SMel::ROW + lane];
            __syncwarp();
            float* dst = P.mel + w * (long long)(SMel::F * SMel::T) + (long long)P.t0 * SMel::F;
            for (int i = lane; i < P.n_tail * SMel::F; i += 32) dst[i] = tile[(i / SMel::F) * (SMel::F + 1) + i % SMel::F];
            __syncwarp();
        }
        __threadfence();
        grid.sync();
    }
    KcState S;
    for (int l = 0; l < P.n_layers; ++l) {
        const TcnFusedLayer& L = P.L[l];
        rowgemm_kc_run<true, true>(L.A, L.av, KcSegs{1 << 30, 0, L.k_valid}, L.K, L.wq, L.bias, L.res, L.out, L.ov, L.rows, L.N, L.N, L.act, ring,
                                   L.rv, L.pre_relu, S, 0, L.N, a_s, b_s, bar_full, bar_empty, bar_afree, bar_done, row_at, tmem_base);
        if (l + 1 < P.n_layers) {
            __threadfence();
            grid.sync();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}
#endif

// QuartzNet depthwise Conv1d (padding 'same', no bias: it is folded into the pointwise bias) on channel-last rows.
// x [n][T][in_pitch] (first C channels used), dw [k][Cp];  a [n * T][K]: columns [0, Cp) = depthwise output
// (zero beyond C), and, when K == 2 Cp, columns [Cp, 2 Cp) = x itself (the residual 1x1 convolution's input).
__global__ void __launch_bounds__(256)
qn_dw_kernel(const float* __restrict__ x, int in_pitch, const float* __restrict__ dw, float* __restrict__ a, long long n, int T,
             int C, int Cp, int k, int K) {
    const int left = (k - 1) / 2, q4 = Cp / 4;
    const long long total = n * T * q4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % q4) * 4;
        const long long row = i / q4;
        const int t = (int)(row % T);
        const float* xw = x + (row - t) * in_pitch;                    // the window's first row
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), ctr = acc;
        if (c < C) {                                                   // C % 4 == 0
            const int j_lo = left - t > 0 ? left - t : 0, j_hi = (T - 1 - t + left < k - 1) ? T - 1 - t + left : k - 1;
            for (int j = j_lo; j <= j_hi; ++j) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(xw + (long long)(t + j - left) * in_pitch + c));
                const float4 w = __ldg(reinterpret_cast<const float4*>(dw + (long long)j * Cp + c));
                acc.x = fmaf(v.x, w.x, acc.x); acc.y = fmaf(v.y, w.y, acc.y);
                acc.z = fmaf(v.z, w.z, acc.z); acc.w = fmaf(v.w, w.w, acc.w);
            }
            ctr = __ldg(reinterpret_cast<const float4*>(xw + (long long)t * in_pitch + c));
        }
        *reinterpret_cast<float4*>(a + row * K + c) = acc;
        if (K == 2 * Cp) *reinterpret_cast<float4*>(a + row * K + Cp + c) = ctr;
    }
}

// Raw-audio front end, step 0: int16 window -> float samples / 32768 (nanointerpreter.py:750) in a zero-padded row
// p [n][len]: p[w][i] = pcm[w][i - pad] for pad <= i < pad + clip, else 0
__global__ void __launch_bounds__(256)
raw_pcm_kernel(WindowSource src, long long n, float* __restrict__ p, int len, int pad) {
    const int q4 = len / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * q4; i += (long long)gridDim.x * blockDim.x) {
        const long long w = i / q4;
        const int i0 = (int)(i - w * q4) * 4;
        float v[4];
        if (src.fbase != nullptr) {                       // float feed: the samples are already x / 32768
            const float* xf = src.atf(w);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = i0 + e - pad;
                v[e] = (j >= 0 && j < src.clip) ? xf[j] : 0.0f;
            }
        } else {
            const int16_t* x = src.at(w);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = i0 + e - pad;
                v[e] = (j >= 0 && j < src.clip) ? (float)x[j] * (1.0f / 32768.0f) : 0.0f;
            }
        }
        *reinterpret_cast<float4*>(p + w * len + i0) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// zero the padding rows of a per-window buffer: floats [0, head) and [tail_off, tail_off + tail_len) of every window
__global__ void __launch_bounds__(256)
zero_pads_kernel(float* __restrict__ buf, long long n, long long win_stride, int head, int tail_off, int tail_len) {
    const int per = head + tail_len;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * per; i += (long long)gridDim.x * blockDim.x) {
        const long long w = i / per;
        const int j = (int)(i - w * per);
        buf[w * win_stride + (j < head ? j : tail_off + j - head)] = 0.0f;
    }
}

// RawAudioBackbone conv1 (architectures.py:741-745): 3x3, 1 -> CO channels, stride (1, 2), padding 1, BatchNorm folded,
// on the raw front end's output read as a one-channel image img[h][w] = bf[w][h] (h = front-end channel, w = time step).
// Writes the zero-padded NHWC image o1[(oh + 1)][(ow + 1)][co] the next layer's GEMM view reads.
__global__ void __launch_bounds__(256)
rawcnn_conv1_kernel(const float* __restrict__ bf, const float* __restrict__ w1 /* [9][CO] */, const float* __restrict__ b1,
                    float* __restrict__ o1, long long n, int H, int W, int Wo, int CO, int act) {
    const int q4 = CO / 4;
    const long long per = (long long)H * Wo * q4, total = n * per;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long w = i / per;
        int r = (int)(i - w * per);
        const int c = (r % q4) * 4;
        r /= q4;
        const int ow = r % Wo, oh = r / Wo;
        const float* img = bf + w * (long long)W * H;
        float4 acc = __ldg(reinterpret_cast<const float4*>(b1 + c));
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int ih = oh + kh - 1, iw = 2 * ow + kw - 1;
                if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
                const float v = __ldg(img + (long long)iw * H + ih);
                const float4 k4 = __ldg(reinterpret_cast<const float4*>(w1 + (kh * 3 + kw) * CO + c));
                acc.x = fmaf(v, k4.x, acc.x); acc.y = fmaf(v, k4.y, acc.y); acc.z = fmaf(v, k4.z, acc.z); acc.w = fmaf(v, k4.w, acc.w);
            }
        acc.x = apply_act(acc.x, act); acc.y = apply_act(acc.y, act); acc.z = apply_act(acc.z, act); acc.w = apply_act(acc.w, act);
        *reinterpret_cast<float4*>(o1 + ((w * (H + 2) + oh + 1) * (long long)(Wo + 2) + ow + 1) * CO + c) = acc;
    }
}

// the same with one thread per output pixel and all CO = 4 Q4 channels in registers (9 input loads per pixel instead of 9 per
// channel quad; the default CO = 24 takes this one)
template <int Q4>
__global__ void __launch_bounds__(256)
rawcnn_conv1_px_kernel(const float* __restrict__ bf, const float* __restrict__ w1 /* [9][4 Q4] */, const float* __restrict__ b1,
                       float* __restrict__ o1, long long n, int H, int W, int Wo, int act) {
    constexpr int CO = 4 * Q4;
    const long long per = (long long)H * Wo, total = n * per;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long w = i / per;
        const int r = (int)(i - w * per);
        const int oh = r % H, ow = r / H;           // consecutive threads walk h: img[h][w] = bf[w][h] is contiguous in h
        const float* img = bf + w * (long long)W * H;
        float4 acc[Q4];
#pragma unroll
        for (int q = 0; q < Q4; ++q) acc[q] = __ldg(reinterpret_cast<const float4*>(b1) + q);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int ih = oh + kh - 1, iw = 2 * ow + kw - 1;
                const float v = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(img + (long long)iw * H + ih) : 0.0f;
#pragma unroll
                for (int q = 0; q < Q4; ++q) {
                    const float4 k4 = __ldg(reinterpret_cast<const float4*>(w1 + (kh * 3 + kw) * CO) + q);
                    acc[q].x = fmaf(v, k4.x, acc[q].x); acc[q].y = fmaf(v, k4.y, acc[q].y);
                    acc[q].z = fmaf(v, k4.z, acc[q].z); acc[q].w = fmaf(v, k4.w, acc[q].w);
                }
            }
        float4* dst = reinterpret_cast<float4*>(o1 + ((w * (H + 2) + oh + 1) * (long long)(Wo + 2) + ow + 1) * CO);
        if (act == ACT_RELU) {
#pragma unroll
            for (int q = 0; q < Q4; ++q)
                dst[q] = make_float4(fmaxf(acc[q].x, 0.0f), fmaxf(acc[q].y, 0.0f), fmaxf(acc[q].z, 0.0f), fmaxf(acc[q].w, 0.0f));
        } else {
#pragma unroll
            for (int q = 0; q < Q4; ++q) dst[q] = kc_act4(acc[q], act);
        }
    }
}

// zero the one-pixel border (and any spare rows / columns) of padded NHWC images buf[n][Hp][Wp][C]; interior = H x W at (1, 1).
// Only border cells are visited: the full rows 0 and H + 1 .. Hp - 1, then columns 0 and W + 1 .. Wp - 1 of rows 1 .. H.
__global__ void __launch_bounds__(256)
zero_border_kernel(float* __restrict__ buf, long long n, int Hp, int Wp, int C, int H, int W) {
    const int q4 = C / 4, row_cells = Wp * (Hp - H), side = Wp - W, cells = row_cells + H * side;
    const long long total = n * cells * q4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % q4);
        const long long t = i / q4;
        const long long w = t / cells;
        const int c = (int)(t - w * cells);
        int y, x;
        if (c < row_cells) {
            const int j = c / Wp;
            y = j == 0 ? 0 : H + j;
            x = c - j * Wp;
        } else {
            const int cc = c - row_cells, j = cc / side, xx = cc - j * side;
            y = 1 + j;
            x = xx == 0 ? 0 : W + xx;
        }
        reinterpret_cast<float4*>(buf)[((w * Hp + y) * Wp + x) * q4 + c4] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// The same depthwise FIR with the window's rows staged in shared memory and the taps in registers: one CTA iteration =
// (window, 32-channel slab); thread = (channel, segment of 13 output steps), a sliding window of inputs feeds the 13
// accumulators (KW + 12 shared-memory loads for 13 KW FMAs).  T = 98 (NS40x98).
constexpr int kQnSeg = 13, kQnSegs = 8;
template <int KW>
__global__ void __launch_bounds__(256)
qn_dw_tile_kernel(const float* __restrict__ x, int in_pitch, const float* __restrict__ dw, float* __restrict__ a, long long n, int T,
                  int C, int Cp, int K) {
    constexpr int LEFT = (KW - 1) / 2, ROWS = kQnSeg * kQnSegs + KW - 1;
    __shared__ float xs[ROWS * 32];
    const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
    const int slabs = Cp / 32;
    for (long long it = blockIdx.x; it < n * slabs; it += gridDim.x) {
        const long long w = it / slabs;
        const int c = (int)(it - w * slabs) * 32 + lane;
        const float* xw = x + w * T * in_pitch;
        __syncthreads();
        for (int i = threadIdx.x; i < ROWS * 32; i += 256) {
            const int t = (i >> 5) - LEFT;                             // source step of shared row i / 32
            xs[i] = (t >= 0 && t < T && c < C) ? __ldg(xw + (long long)t * in_pitch + c) : 0.0f;
        }
        float wt[KW];
#pragma unroll
        for (int j = 0; j < KW; ++j) wt[j] = c < C ? __ldg(dw + (long long)j * Cp + c) : 0.0f;
        __syncthreads();
        float acc[kQnSeg];
#pragma unroll
        for (int r = 0; r < kQnSeg; ++r) acc[r] = 0.0f;
        const int t0 = seg * kQnSeg;
#pragma unroll
        for (int j = 0; j < kQnSeg + KW - 1; ++j) {
            const float v = xs[(t0 + j) * 32 + lane];
#pragma unroll
            for (int r = 0; r < kQnSeg; ++r)
                if (j - r >= 0 && j - r < KW) acc[r] = fmaf(wt[j - r], v, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < kQnSeg; ++r) {
            const int t = t0 + r;
            if (t < T) {
                float* row = a + (w * T + t) * (long long)K;
                row[c] = acc[r];
                if (K == 2 * Cp) row[Cp + c] = xs[(t + LEFT) * 32 + lane];
            }
        }
    }
}

}  // namespace nww
