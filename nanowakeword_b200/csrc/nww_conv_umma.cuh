// nww_conv_umma.cuh — 3x3 convolution (pad 1, stride 1) + bias + activation (+ 2x2 max pool) on channel-last
// activations as a tcgen05 implicit GEMM with NO im2col, for small images that fit one CTA:
//   E2E mel-CNN body: conv2 16 -> 32 on 32 x 50 (+pool), conv3 32 -> 64 on 16 x 25   (architectures.py:840-856)
//   CRNN: conv3 32 -> 32 on 10 x 24 (+pool)                                               (architectures.py:222-230)
// BatchNorm is folded into weights / bias by the packer.
//
// Same scheme as conv2 of nww_cnn2.cuh, generalised.  One window per CTA iteration:
//   * the window's input [H][W][Cin] (FP32, global) is converted to bf16 hi / lo and stored as zero-bordered
//     position lists, 16 bytes (8 channels) per position and K group:
//       no pool: one plane,   s = (y + 1) * P + (x + 1),              P = W + 2,     GEMM row m = y * P + x
//       pool:    four parity planes (y & 1, x & 1), s = ((y >> 1) + 1) * P + (x >> 1) + 1, P = W / 2 + 2,
//                GEMM row m = ph * P + pw, one accumulator per pooling quad;
//   * the A operand of a tap is that plane behind a descriptor whose start address is shifted by whole positions
//     (K-major, un-swizzled: SBO = 128 B, LBO = plane K-group stride);
//   * weights arrive pre-split from the engine: [tap][hi|lo][K group][Cout][8 ic] bf16;
//   * per 128-row tile: (quads) x 9 taps x (Cin / 16) x 3 split products of tcgen05.mma 128 x Cout x 16, FP32
//     accumulation in TMEM; epilogue = bias, activation, (max over the four quads), channel-last FP32 store.
//
// Schedule (round 2).  Where two sets of position lists and two sets of accumulators fit (conv3 of both models) the
// windows are software-pipelined: iteration i issues MMAs(i), then runs epilogue(i - 1) and the loads / conversions of
// window i + 1 under them; one CTA barrier per window.  The E2E conv2 (one set: 147 KB of position lists, all 512
// TMEM columns) instead takes its input from the log-mel and computes the FIRST layer — Conv2d(1, 16) + act + pool —
// inside the loader (FRONT1), so that layer's activations never exist in HBM.  The activation is a template parameter
// (the run-time form put an erf / exp / branch triple at ~100 unrolled call sites and pushed the kernel out of the
// instruction cache); MMAs are issued by one lane of each of the last warps from tap offsets tabulated once.
#pragma once

#include <string.h>
#include <algorithm>
#include <vector>

#include "nww_tc.cuh"

#ifndef NWW_CPUSIM
#ifndef NWW_DYN_SMEM
#define NWW_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#endif
#endif

namespace nww {

constexpr int kCuNT = 512;            // 16 warps: the load / convert phase is latency-bound, the epilogue has (tile, chunk) tasks for all

struct ConvUmmaPlan {
    int H, W, Cin, Cout, pool;
    int P;            // position pitch
    int rows;         // GEMM rows that hold real outputs: (pool ? H/2 : H) * P
    int tiles;        // 128-row tiles
    int npos;         // positions allocated per plane and K group
    int planes;       // 1 or 4
    int kg;           // Cin / 8
    int tmem_cols;    // power of two >= nbuf * tiles * quads * Cout
    int acc_cols;     // tiles * quads * Cout: one set of accumulators
    int nbuf;         // 2: two sets of position lists AND two sets of accumulators: window i + 1's MMAs run under window i's
                      // epilogue and window i + 2's loads / conversions (the tensor pipe is the only thing that is not overlapped)
    size_t a_bytes, b_bytes, smem_bytes;
    // fused first layer (conv_umma_plan_front1): the loader computes Conv2d(1, Cin, 3, pad 1) + act + MaxPool2d(2) from the
    // log-mel (Hm x Wm per window) straight into the position lists, so that layer's activations never exist in HBM
    int f1_hm = 0, f1_wm = 0, f1_pitch = 0;
    size_t f1_tile = 0, f1_stage = 0, f1_w = 0;      // shared-memory offsets: zero-bordered mel tile | raw mel (TMA target) | weights + bias
};

inline bool conv_umma_plan(int H, int W, int Cin, int Cout, int pool, ConvUmmaPlan* p) {
    p->H = H; p->W = W; p->Cin = Cin; p->Cout = Cout; p->pool = pool;
    if (Cin % 16 || Cout % 16 || Cout > 64 || (pool && ((H | W) & 1))) return false;
    const int Ho = pool ? H / 2 : H;
    p->P = (pool ? W / 2 : W) + 2;
    p->rows = Ho * p->P;
    p->tiles = (p->rows + 127) / 128;
    p->planes = pool ? 4 : 1;
    p->kg = Cin / 8;
    p->npos = p->tiles * 128 + 2 * p->P + 8;
    const int cols = p->tiles * (pool ? 4 : 1) * Cout;
    if (cols > 512) return false;
    p->acc_cols = cols;
    p->a_bytes = (size_t)p->planes * 2 * p->kg * p->npos * 16;
    p->b_bytes = (size_t)9 * 2 * p->kg * Cout * 16;
    p->nbuf = (2 * cols <= 512 && 2 * p->a_bytes + p->b_bytes + Cout * sizeof(float) + 256 <= 226 * 1024) ? 2 : 1;
    const int tc = p->nbuf * cols;
    p->tmem_cols = tc <= 32 ? 32 : tc <= 64 ? 64 : tc <= 128 ? 128 : tc <= 256 ? 256 : 512;
    p->smem_bytes = p->nbuf * p->a_bytes + p->b_bytes + Cout * sizeof(float) + 256;
    return p->smem_bytes <= 226 * 1024;
}

// Adds the fused first layer to a plan: mel (Hm, Wm) -> conv 3x3 (1 -> Cin = 16) -> pool -> (H, W) = (Hm / 2, Wm / 2).
inline bool conv_umma_plan_front1(ConvUmmaPlan* plan, int Hm, int Wm) {
    ConvUmmaPlan q = *plan;                                               // the plan changes only if everything fits
    ConvUmmaPlan* p = &q;
    if (!p->pool || p->Cin != 16 || p->H != Hm / 2 || p->W != Wm / 2 || (Hm * Wm) % 4) return false;
    if (p->nbuf == 2) {                                                    // the fused first layer works on one set of position lists
        p->nbuf = 1;
        p->smem_bytes -= p->a_bytes;
        p->tmem_cols = p->acc_cols <= 32 ? 32 : p->acc_cols <= 64 ? 64 : p->acc_cols <= 128 ? 128 : p->acc_cols <= 256 ? 256 : 512;
    }
    p->f1_hm = Hm; p->f1_wm = Wm;
    p->f1_pitch = (std::max(Wm + 2, 2 * p->W + 4) + 3) / 4 * 4;           // even: the 4 x 4 patches are read as float2
    size_t off = (p->smem_bytes + 127) / 128 * 128;
    p->f1_tile = off;  off += (size_t)(Hm + 2) * p->f1_pitch * sizeof(float);
    off = (off + 127) / 128 * 128;
    p->f1_stage = off; off += (size_t)Hm * Wm * sizeof(float);
    p->f1_w = off;     off += (size_t)(9 * 16 + 16) * sizeof(float) + 16;  // + the mel mbarrier
    p->smem_bytes = off;
    if (p->smem_bytes > 227 * 1024) return false;
    *plan = q;
    return true;
}

// host: folded (Cin, 9, Cout) FP32 weights -> [tap][hi|lo][K group][Cout][8] bf16
inline void conv_umma_pack_weights(const float* w, int Cin, int Cout, std::vector<uint16_t>* out) {
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
    const size_t op = (size_t)(Cin / 8) * Cout * 8;              // elements per (tap, hi|lo) operand
    out->assign((size_t)9 * 2 * op, 0);
    for (int ic = 0; ic < Cin; ++ic)
        for (int tap = 0; tap < 9; ++tap)
            for (int oc = 0; oc < Cout; ++oc) {
                const float v = w[((size_t)ic * 9 + tap) * Cout + oc];
                const uint16_t hi = bf16_rn(v), lo = bf16_rn(v - bf16_f(hi));
                const size_t base = (size_t)tap * 2 * op + (size_t)(ic >> 3) * Cout * 8 + (size_t)oc * 8 + (ic & 7);
                (*out)[base] = hi;
                (*out)[base + op] = lo;
            }
}

// in [n][H*W][Cin] -> out [n][Ho*Wo][Cout]  (Ho, Wo = H/2, W/2 with pool)
// FRONT1: `in` is the log-mel [n][Hm][Wm]; w1 [9][16] (tap-major) and b1 [16] are the first layer's folded weights (the
// arithmetic of bc_init_conv_kernel, FMA for FMA, so both routes give the same activations).
// ACT: the activation is a template parameter — with a run-time code every one of the ~100 unrolled call sites carried
// the erf / exp paths too, and the kernel no longer fitted the instruction cache (13 % of the stall samples were
// instruction fetches).
template <bool FRONT1, int ACT>
__global__ void __launch_bounds__(kCuNT, 1)
conv3x3_umma_kernel(const float* __restrict__ in, const uint4* __restrict__ wq, const float* __restrict__ bias,
                    float* __restrict__ out, long long n_windows, ConvUmmaPlan P,
                    const float* __restrict__ w1 = nullptr, const float* __restrict__ b1 = nullptr) {
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* a_s = smem;                                     // P.nbuf sets of position lists
    const size_t a_all = (size_t)P.nbuf * P.a_bytes;
    unsigned char* b_s = smem + a_all;
    float* bias_s = reinterpret_cast<float*>(smem + a_all + P.b_bytes);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + a_all + P.b_bytes + P.Cout * sizeof(float));
    bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bar) + 7) & ~(uintptr_t)7);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);    // bar[2]: one per accumulator set
    const int n_acc = P.tiles * (P.pool ? 4 : 1);                  // accumulators = (tile, quad) pairs
    const int n_issuers = n_acc < 4 ? n_acc : 4;                   // one lane of the last n_issuers warps each issues the MMAs of its accumulators
    if (tid == 0) {
        mbar_init(bar, (uint32_t)n_issuers);
        mbar_init(bar + 1, (uint32_t)n_issuers);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
    for (int i = tid; i < (int)(a_all / 16); i += kCuNT) reinterpret_cast<uint4*>(a_s)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < (int)(P.b_bytes / 16); i += kCuNT) reinterpret_cast<uint4*>(b_s)[i] = __ldg(wq + i);
    for (int i = tid; i < P.Cout; i += kCuNT) bias_s[i] = bias[i];
    float* tile1 = reinterpret_cast<float*>(smem + P.f1_tile);
    float* stage1 = reinterpret_cast<float*>(smem + P.f1_stage);
    float* w1s = reinterpret_cast<float*>(smem + P.f1_w);                  // [9][16] | bias [16]
    uint64_t* mel_bar = reinterpret_cast<uint64_t*>(smem + P.f1_w + (9 * 16 + 16) * sizeof(float));
    const uint32_t mel_bytes = (uint32_t)(P.f1_hm * P.f1_wm) * (uint32_t)sizeof(float);
    if (FRONT1) {
        for (int i = tid; i < (P.f1_hm + 2) * P.f1_pitch; i += kCuNT) tile1[i] = 0.0f;     // the border stays zero
        for (int i = tid; i < 9 * 16; i += kCuNT) w1s[i] = w1[i];
        for (int i = tid; i < 16; i += kCuNT) w1s[9 * 16 + i] = b1[i];
        if (tid == 0) {
            mbar_init(mel_bar, 1);
            fence_mbar_init();
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (FRONT1 && tid == 0 && (long long)blockIdx.x < n_windows) {
        mbar_expect_tx(mel_bar, mel_bytes);
        bulk_g2s(stage1, in + (long long)blockIdx.x * (P.f1_hm * P.f1_wm), mel_bytes, mel_bar);
    }
    uint32_t mel_phase = 0;
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lbo_a = (uint32_t)P.npos * 16, lbo_b = (uint32_t)P.Cout * 16;
    const uint32_t plane_bytes = (uint32_t)P.kg * lbo_a;            // one (plane, hi|lo)
    const uint32_t bop_bytes = (uint32_t)P.kg * lbo_b;              // one (tap, hi|lo)
    const uint64_t da_base = umma_desc_noswz(smem_u32(a_s), lbo_a, 128);
    const uint64_t db_base = umma_desc_noswz(smem_u32(b_s), lbo_b, 128);
    const uint32_t idesc = umma_idesc_bf16(128, P.Cout);
    const int quads = P.pool ? 4 : 1;
    const int Ho = P.pool ? P.H / 2 : P.H, Wo = P.pool ? P.W / 2 : P.W;

    // input window w (channel-last FP32) -> bf16 hi / lo position lists at ab
    auto load_window = [&](long long w, unsigned char* ab) {
        // ---- input window -> bf16 hi / lo position lists ------------------------------------------------------------
        const float* src = in + w * (long long)P.H * P.W * P.Cin;
        // (four cells per thread and trip, all eight 128-bit loads issued before the first conversion)
        const int n_cells = P.H * P.W * P.kg;
        for (int i0 = tid; i0 < n_cells; i0 += 4 * kCuNT) {
            float4 v0[4], v1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + j * kCuNT;
                if (i < n_cells) {
                    const int g = i % P.kg, pix = i / P.kg;
                    const float4* p4 = reinterpret_cast<const float4*>(src + (size_t)pix * P.Cin + 8 * g);
                    v0[j] = __ldg(p4);
                    v1[j] = __ldg(p4 + 1);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + j * kCuNT;
                if (i >= n_cells) break;
                const int g = i % P.kg, pix = i / P.kg;
                const int y = pix / P.W, x = pix - y * P.W;
                const float v[8] = {v0[j].x, v0[j].y, v0[j].z, v0[j].w, v1[j].x, v1[j].y, v1[j].z, v1[j].w};
                uint32_t h[8], l[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    h[k] = float_to_bf16_bits(v[k]);
                    l[k] = float_to_bf16_bits(v[k] - bf16_bits_to_float(h[k]));
                }
                int plane, s;
                if (P.pool) {
                    plane = ((y & 1) << 1) | (x & 1);
                    s = ((y >> 1) + 1) * P.P + (x >> 1) + 1;
                } else {
                    plane = 0;
                    s = (y + 1) * P.P + x + 1;
                }
                unsigned char* dst = ab + (size_t)plane * 2 * plane_bytes + (size_t)g * lbo_a + (size_t)s * 16;
                *reinterpret_cast<uint4*>(dst) = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
                *reinterpret_cast<uint4*>(dst + plane_bytes) =
                    make_uint4(l[0] | (l[1] << 16), l[2] | (l[3] << 16), l[4] | (l[5] << 16), l[6] | (l[7] << 16));
            }
        }
    };

    // MMA issue.  One lane of each of the LAST n_issuers warps (the ones with the least epilogue work when a window has few
    // (tile, chunk) tasks) issues the MMAs of "its" accumulators: with four pooling quads issuer i always owns quad i, without
    // pooling quad 0, so the nine taps' operand offsets are per-thread constants worked out once — the issue loop is a
    // single-thread dependent chain (an ELECT + four R2UR + the MMA per product as it is) and was the longest pole of a window
    // when the descriptors were re-derived per tap (ncu: 57 % of the stall samples at the CTA barrier behind it).
    const int issuer = warp - (kCuNT / 32 - n_issuers);                // 0 .. n_issuers - 1 for the issuing warps, negative otherwise
    // A start offset of tap (r, c) for quad q, 16-byte units: a small shared table (registers are what the loader and the
    // epilogue are short of)
    uint32_t* a_tab = reinterpret_cast<uint32_t*>(tmem_slot + 2);        // [4 quads][9 taps], behind the barriers / TMEM slot
    if (tid < 36) {
        const int quad = tid / 9, tap = tid - quad * 9, dy = quad >> 1, dx = quad & 1;
        {
            const int r = tap / 3, c = tap % 3;
            int plane, s0;
            if (P.pool) {
                const int ry = dy + r - 1, cx = dx + c - 1;
                plane = ((ry & 1) << 1) | (cx & 1);
                s0 = (1 + (ry >> 1)) * P.P + 1 + (cx >> 1);
            } else {
                plane = 0;
                s0 = r * P.P + c;
            }
            a_tab[tid] = (uint32_t)(((size_t)plane * 2 * plane_bytes) / 16 + s0);
        }
    }
    __syncthreads();
    const uint32_t a_lo_off = plane_bytes / 16, b_lo_off = bop_bytes / 16, b_tap_step = 2 * bop_bytes / 16;
    const uint32_t a_kh_step = 2 * lbo_a / 16, b_kh_step = 2 * lbo_b / 16;
    const int n_kh = P.Cin / 16;                                         // 16 channels = 2 K groups per MMA
    auto issue_mmas = [&](uint64_t da_win, uint32_t tm_off, uint64_t* bar_set) {
        if (lane == 0 && issuer >= 0) {
            tc_fence_after();
            for (int acc = issuer; acc < n_acc; acc += n_issuers) {
                const int t = acc / quads;
                const uint32_t d_tmem = tmem_base + tm_off + (uint32_t)(acc * P.Cout);
                const uint64_t a_t = da_win + (uint64_t)(t * 128);
                const uint32_t* a_tap = a_tab + (quads == 4 ? (acc & 3) : 0) * 9;
                uint32_t accumulate = 0u;
                uint32_t a_off[9];
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) a_off[tap] = a_tap[tap];          // nine independent shared loads up front
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    uint64_t a_hi = a_t + a_off[tap], b_hi = db_base + (uint64_t)(tap * b_tap_step);
                    for (int kh = 0; kh < n_kh; ++kh, a_hi += a_kh_step, b_hi += b_kh_step) {
                        umma_bf16(d_tmem, a_hi, b_hi, idesc, accumulate);
                        umma_bf16(d_tmem, a_hi + a_lo_off, b_hi, idesc, 1);
                        umma_bf16(d_tmem, a_hi, b_hi + b_lo_off, idesc, 1);
                        accumulate = 1u;
                    }
                }
            }
            umma_commit(bar_set);
        }
    };
    // bias, activation (+ max over the pooling quads), channel-last FP32 store of window w from accumulator set tm_off
    auto epilogue = [&](long long w, uint32_t tm_off) {
        // ---- epilogue: warp -> TMEM lane quarter; tasks (tile, 16-column chunk) dealt to the two warps of a quarter ----
        {
            const int q = warp & 3, sub = warp >> 2;                 // 16 warps: four per quarter
            const int chunks = P.Cout / 16;
            float* dst_w = out + w * (long long)Ho * Wo * P.Cout;
            for (int task = sub; task < P.tiles * chunks; task += kCuNT / 128) {
                const int t = task / chunks, ch = task - t * chunks;
                uint32_t r[4][16];
                const uint32_t taddr = tmem_base + tm_off + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * quads * P.Cout + ch * 16);
                tmem_ld_32x32b_x16_nowait(taddr, r[0]);
                if (P.pool) {
                    tmem_ld_32x32b_x16_nowait(taddr + P.Cout, r[1]);
                    tmem_ld_32x32b_x16_nowait(taddr + 2 * P.Cout, r[2]);
                    tmem_ld_32x32b_x16_nowait(taddr + 3 * P.Cout, r[3]);
                }
                tmem_ld_wait();
                const int m = t * 128 + q * 32 + lane;
                const int py = m / P.P, px = m - py * P.P;
                if (py < Ho && px < Wo) {
                    float v[16];
#pragma unroll
                    for (int o = 0; o < 16; ++o) {
                        const float b = bias_s[ch * 16 + o];
                        float x0 = apply_act(__uint_as_float(r[0][o]) + b, ACT);
                        if (P.pool) {
                            x0 = fmaxf(x0, apply_act(__uint_as_float(r[1][o]) + b, ACT));
                            x0 = fmaxf(x0, apply_act(__uint_as_float(r[2][o]) + b, ACT));
                            x0 = fmaxf(x0, apply_act(__uint_as_float(r[3][o]) + b, ACT));
                        }
                        v[o] = x0;
                    }
                    float4* d4 = reinterpret_cast<float4*>(dst_w + ((size_t)py * Wo + px) * P.Cout + ch * 16);
#pragma unroll
                    for (int j = 0; j < 4; ++j) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
            tc_fence_before();
        }
    };

    // the first layer's loader (FRONT1): window w's log-mel -> conv1 + act + pool -> the position lists at a_s
    auto load_front1 = [&](long long w) {
        // ---- log-mel (TMA, fetched behind the previous window) -> zero-bordered tile; next window's fetch starts ----
        mbar_wait(mel_bar, mel_phase);
        mel_phase ^= 1;
        for (int i = tid; i < P.f1_hm * P.f1_wm; i += kCuNT) {
            const int r = i / P.f1_wm, c = i - r * P.f1_wm;
            tile1[(r + 1) * P.f1_pitch + c + 1] = stage1[i];
        }
        __syncthreads();
        if (tid == 0 && w + gridDim.x < n_windows) {
            fence_proxy_async();
            mbar_expect_tx(mel_bar, mel_bytes);
            bulk_g2s(stage1, in + (w + gridDim.x) * (long long)(P.f1_hm * P.f1_wm), mel_bytes, mel_bar);
        }
        // ---- first layer: task = (pooled pixel, 8 channels) -> bf16 hi / lo rows of the parity planes -------------
        for (int T = tid; T < P.H * P.W * 2; T += kCuNT) {
            const int pix = T >> 1, cg = T & 1;
            const int y = pix / P.W, x = pix - y * P.W;
            float pin[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float2 lo = *reinterpret_cast<const float2*>(tile1 + (2 * y + r) * P.f1_pitch + 2 * x);
                const float2 hi = *reinterpret_cast<const float2*>(tile1 + (2 * y + r) * P.f1_pitch + 2 * x + 2);
                pin[r][0] = lo.x; pin[r][1] = lo.y; pin[r][2] = hi.x; pin[r][3] = hi.y;
            }
            // packed FP32 FMAs (FFMA2: two IEEE FMAs per instruction, channels o and o + 1 of one conv output; bit-identical to the
            // scalar form): with the activation templated this loader is issue-bound, and 144 slots instead of 288 per task count
            float2 acc[4][4];                                          // [channel pair][pooling quad]
#pragma unroll
            for (int o2 = 0; o2 < 4; ++o2) {
                const float2 bv = *reinterpret_cast<const float2*>(w1s + 9 * 16 + cg * 8 + 2 * o2);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) acc[o2][q4] = bv;
            }
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float4 wa = *reinterpret_cast<const float4*>(w1s + (r * 3 + c) * 16 + cg * 8);
                    const float4 wb = *reinterpret_cast<const float4*>(w1s + (r * 3 + c) * 16 + cg * 8 + 4);
                    const float2 wv[4] = {make_float2(wa.x, wa.y), make_float2(wa.z, wa.w), make_float2(wb.x, wb.y), make_float2(wb.z, wb.w)};
                    const float2 p00 = make_float2(pin[r][c], pin[r][c]), p01 = make_float2(pin[r][c + 1], pin[r][c + 1]);
                    const float2 p10 = make_float2(pin[r + 1][c], pin[r + 1][c]), p11 = make_float2(pin[r + 1][c + 1], pin[r + 1][c + 1]);
#pragma unroll
                    for (int o2 = 0; o2 < 4; ++o2) {
                        acc[o2][0] = __ffma2_rn(p00, wv[o2], acc[o2][0]);
                        acc[o2][1] = __ffma2_rn(p01, wv[o2], acc[o2][1]);
                        acc[o2][2] = __ffma2_rn(p10, wv[o2], acc[o2][2]);
                        acc[o2][3] = __ffma2_rn(p11, wv[o2], acc[o2][3]);
                    }
                }
            uint32_t h[8], l[8];
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const float a0 = (o & 1) ? acc[o >> 1][0].y : acc[o >> 1][0].x, a1 = (o & 1) ? acc[o >> 1][1].y : acc[o >> 1][1].x;
                const float a2 = (o & 1) ? acc[o >> 1][2].y : acc[o >> 1][2].x, a3 = (o & 1) ? acc[o >> 1][3].y : acc[o >> 1][3].x;
                const float best = fmaxf(fmaxf(apply_act(a0, ACT), apply_act(a1, ACT)), fmaxf(apply_act(a2, ACT), apply_act(a3, ACT)));
                h[o] = float_to_bf16_bits(best);
                l[o] = float_to_bf16_bits(best - bf16_bits_to_float(h[o]));
            }
            const int plane = ((y & 1) << 1) | (x & 1);
            const int s = ((y >> 1) + 1) * P.P + (x >> 1) + 1;
            unsigned char* dst = a_s + (size_t)plane * 2 * plane_bytes + (size_t)cg * lbo_a + (size_t)s * 16;
            *reinterpret_cast<uint4*>(dst) = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
            *reinterpret_cast<uint4*>(dst + plane_bytes) =
                make_uint4(l[0] | (l[1] << 16), l[2] | (l[3] << 16), l[4] | (l[5] << 16), l[6] | (l[7] << 16));
        }
    };

    // One schedule for every configuration, one call site per phase (the kernel is instruction-cache sized, so the phases
    // must not be duplicated).  Iteration i (i = -1 .. n_it):
    //   issue MMAs(i)  ->  [wait] epilogue(e)  ->  load window i + 1  ->  CTA barrier
    // with e = i - 1 when there are two sets of position lists / accumulators (MMAs(i) run under epilogue(i - 1) and the loads of
    // window i + 1; the buffer those loads overwrite was last read by MMAs(i - 1), which the wait has seen complete) and
    // e = i with one set (the loads of window i + 1 follow the wait for MMAs(i)).
    const int n_it = (long long)blockIdx.x < n_windows ? (int)((n_windows - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    const int lag = P.nbuf == 2 ? 1 : 0;
    uint32_t ph[2] = {0, 0};
    for (int it = -1; it < n_it + lag; ++it) {
        if (it >= 0 && it < n_it) {
            const int set = it & (P.nbuf - 1);
            issue_mmas(da_base + (uint64_t)(((size_t)set * P.a_bytes) >> 4), (uint32_t)(set * P.acc_cols), bar + set);
        }
        const int e = it - lag;
        if (e >= 0 && e < n_it) {
            const int set = e & (P.nbuf - 1);
            mbar_wait(bar + set, ph[set]);
            ph[set] ^= 1u;
            tc_fence_after();
            epilogue((long long)blockIdx.x + (long long)e * gridDim.x, (uint32_t)(set * P.acc_cols));
        }
        if (it + 1 < n_it) {
            const long long wn = (long long)blockIdx.x + (long long)(it + 1) * gridDim.x;
            if (FRONT1) load_front1(wn);
            else load_window(wn, a_s + (size_t)((it + 1) & (P.nbuf - 1)) * P.a_bytes);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
    }
}

// host: pick the instantiation for a run-time activation code and launch it
template <bool FRONT1>
inline cudaError_t conv3x3_umma_launch(int act, int grid, cudaStream_t st, const float* in, const uint4* wq, const float* bias, float* out,
                                       long long n, const ConvUmmaPlan& P, const float* w1 = nullptr, const float* b1 = nullptr) {
    auto k = act == ACT_RELU ? conv3x3_umma_kernel<FRONT1, ACT_RELU> : act == ACT_GELU ? conv3x3_umma_kernel<FRONT1, ACT_GELU>
                                                                                          : conv3x3_umma_kernel<FRONT1, ACT_SILU>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem_bytes);
    if (e != cudaSuccess) return e;
    k<<<grid, kCuNT, P.smem_bytes, st>>>(in, wq, bias, out, n, P, w1, b1);
    return cudaGetLastError();
}

// AdaptiveAvgPool2d((1, OW)) in its deployed AvgPool2d form (_export/onnx.py:139-147) on channel-last input:
// in [B][H*W][C] -> out [B][C*OW]  (feature index c * OW + j, the reference's flatten of (C, 1, OW))
// One thread = one (window, channel quad): it streams its four channels' H x W activations once as 128-bit loads (a warp
// reads two pixels' 256 contiguous bytes per instruction, a whole image row of W loads in flight) and adds each pixel to
// the bins that contain its column — rows outer, columns inner, the order a per-bin loop uses.  W and OW are
// compile-time so that the bin tests fold away and the row loop is fully unrolled.
template <int W, int OW>
__global__ void __launch_bounds__(256)
avgpool_row_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, long long B, int C, int H) {
    constexpr int sw = W / OW, kw = W - (OW - 1) * sw;
    const int c4n = C / 4;
    const long long total = B * c4n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(t % c4n);
        const long long b = t / c4n;
        const float4* src = reinterpret_cast<const float4*>(in + b * (long long)H * W * C) + c4;
        float4 s[OW];
#pragma unroll
        for (int j = 0; j < OW; ++j) s[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int y = 0; y < H; ++y) {
            const float4* row = src + (long long)y * W * c4n;
            float4 v[W];
#pragma unroll
            for (int x = 0; x < W; ++x) v[x] = __ldg(row + (long long)x * c4n);
#pragma unroll
            for (int x = 0; x < W; ++x)
#pragma unroll
                for (int j = 0; j < OW; ++j)
                    if (x >= j * sw && x < j * sw + kw) { s[j].x += v[x].x; s[j].y += v[x].y; s[j].z += v[x].z; s[j].w += v[x].w; }
        }
        float* o = out + b * (long long)C * OW + (long long)(4 * c4) * OW;
#pragma unroll
        for (int j = 0; j < OW; ++j) {
            o[j] = s[j].x / (float)(H * kw); o[OW + j] = s[j].y / (float)(H * kw);
            o[2 * OW + j] = s[j].z / (float)(H * kw); o[3 * OW + j] = s[j].w / (float)(H * kw);
        }
    }
}

}  // namespace nww
