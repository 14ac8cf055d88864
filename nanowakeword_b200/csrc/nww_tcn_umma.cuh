// nww_tcn_umma.cuh — the TCN dependency-cone kernel (nww_tcn.cuh) with every layer's GEMM on tcgen05.
//
// Same cone, same activation buffers and the same reference contract as tcn_cone_kernel
// (TCNModel / TemporalBlock, nanowakeword/modules/architectures.py:290-362); only the inner products move:
//   * a tile is WT = 4 windows, so a layer has at most 4 x 27 = 108 rows (window, position): ONE 128-row MMA tile;
//   * per layer the A operand is built explicitly (the tile is small, im2col costs a few thousand elements):
//     row r, K group g = 8 consecutive input channels of one tap, FP32 from the shared activation buffer ->
//     bf16 hi / lo, un-swizzled K-major [hi|lo][K group][rows8][16 B] (LBO = rows8 * 16, SBO = 128); rows past
//     rows8 read whatever follows in shared memory and land in accumulator rows nobody reads;
//   * the folded weights come pre-split from the engine in the chunks the kernel consumes
//     ([chunk][hi|lo][K group][columns][8] bf16, at most 48 KB per chunk: wide layers are cut into 64-column
//     and 24-K-group pieces), each chunk = (K groups / 2) x 3 tcgen05.mma 128 x columns x 16 accumulated in TMEM;
//   * epilogue from TMEM: bias, ReLU, residual (identity from the block input or the 1x1 downsample result),
//     ReLU, FP32 store into the next activation buffer.
#pragma once

#include <string.h>
#include <vector>

#include "nww_tc.cuh"
#include "nww_tcn.cuh"

namespace nww {

constexpr int kTuNT = 256, kTuWT = 4, kTuMaxLayers = 12, kTuMaxChunks = 24;
constexpr int kTuChunkElems = 1536;        // K groups x columns per weight chunk (x 32 B = 48 KB)

struct TuLayer {
    int in_off, in_mul, in_add, Cin, taps, kgs, Cout, n_pos, out_off, relu;
    int res_off, res_mul, res_add, res_relu;          // res_off < 0: no residual
    int rows8, chunk0, nchunks;
    const float* bias;
};
struct TuChunk { int kg0, kgs, n0, ncols; unsigned w_off16; };

struct TcnUmmaParams {
    int n_layers, per_window, n_in, c_in, T, off_in, c_last, last_off;
    unsigned a_bytes, b_bytes;
    TuLayer layers[kTuMaxLayers];
    TuChunk chunks[kTuMaxChunks];
};

inline size_t tcn_umma_smem_bytes(const TcnUmmaParams& P) {
    return sizeof(float) * (size_t)P.per_window * kTuWT + P.a_bytes + 2048 + P.b_bytes + 128;
}

// Host: lay out the layer program and pack the weights.  `w` / `bias` are host / device pointers per layer in the
// order conv1, [down], conv2 of every level (the order the kernel runs them).
struct TuHostLayer { const float* w_host; const float* bias_dev; int Cin, taps, Cout; };

inline bool tcn_umma_build(const TcnConeParams& C, const std::vector<TuHostLayer>& hl, TcnUmmaParams* P, std::vector<uint16_t>* wq) {
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
    *P = TcnUmmaParams{};
    P->per_window = C.per_window; P->n_in = C.n_in; P->c_in = C.c_in; P->T = C.T; P->off_in = C.off_in;
    P->c_last = C.ch[C.levels - 1]; P->last_off = C.off_out[C.levels - 1];
    wq->clear();
    int li = 0, ci = 0, hi_idx = 0;
    int x_off = C.off_in, cin = C.c_in;
    unsigned a_max = 0, b_max = 0;
    auto add_layer = [&](const TuHostLayer& H, int in_off, int in_mul, int in_add, int n_pos, int out_off, int relu, int res_off,
                         int res_mul, int res_add, int res_relu) -> bool {
        if (li >= kTuMaxLayers || H.Cin % 8 || H.Cout % 64 || H.Cout > 128) return false;
        TuLayer& L = P->layers[li++];
        const int K = H.taps * H.Cin;
        L.in_off = in_off; L.in_mul = in_mul; L.in_add = in_add; L.Cin = H.Cin; L.taps = H.taps;
        L.kgs = ((K + 15) / 16) * 2;                       // K padded to a multiple of 16 (one MMA = 2 K groups)
        L.Cout = H.Cout; L.n_pos = n_pos; L.out_off = out_off; L.relu = relu;
        L.res_off = res_off; L.res_mul = res_mul; L.res_add = res_add; L.res_relu = res_relu;
        L.rows8 = (kTuWT * n_pos + 7) & ~7;
        if (kTuWT * n_pos > 128) return false;
        L.bias = H.bias_dev;
        L.chunk0 = ci;
        a_max = std::max<unsigned>(a_max, 2u * L.kgs * L.rows8 * 16u);
        const int ncols = (L.kgs * H.Cout <= kTuChunkElems) ? H.Cout : 64;
        int kgc = std::min(L.kgs, kTuChunkElems / ncols) & ~1;
        for (int n0 = 0; n0 < H.Cout; n0 += ncols)
            for (int kg0 = 0; kg0 < L.kgs; kg0 += kgc) {
                if (ci >= kTuMaxChunks) return false;
                TuChunk& Ck = P->chunks[ci++];
                Ck.kg0 = kg0; Ck.kgs = std::min(kgc, L.kgs - kg0); Ck.n0 = n0; Ck.ncols = ncols;
                Ck.w_off16 = (unsigned)(wq->size() / 8);
                const size_t op = (size_t)Ck.kgs * ncols * 8;
                const size_t base0 = wq->size();
                wq->resize(base0 + 2 * op, 0);
                for (int kg = 0; kg < Ck.kgs; ++kg)
                    for (int n = 0; n < ncols; ++n)
                        for (int e = 0; e < 8; ++e) {
                            const int k = (kg0 + kg) * 8 + e;
                            const float v = k < K ? H.w_host[(size_t)k * H.Cout + n0 + n] : 0.0f;
                            const uint16_t hi = bf16_rn(v), lo = bf16_rn(v - bf16_f(hi));
                            (*wq)[base0 + ((size_t)kg * ncols + n) * 8 + e] = hi;
                            (*wq)[base0 + op + ((size_t)kg * ncols + n) * 8 + e] = lo;
                        }
                b_max = std::max<unsigned>(b_max, (unsigned)(2 * op * 2));
            }
        L.nchunks = ci - L.chunk0;
        return true;
    };
    for (int l = 0; l < C.levels; ++l) {
        const int Cc = C.ch[l];
        if (hi_idx >= (int)hl.size()) return false;
        if (!add_layer(hl[hi_idx++], x_off, 1, 0, C.n_mid[l], C.off_mid[l], 1, -1, 0, 0, 0)) return false;   // conv1
        int res_off = x_off, res_mul = 2, res_add = 4;
        if (cin != Cc) {                                                                                       // downsample
            if (!add_layer(hl[hi_idx++], x_off, 2, 4, C.n_out[l], C.off_res, 0, -1, 0, 0, 0)) return false;
            res_off = C.off_res; res_mul = 1; res_add = 0;
        }
        if (!add_layer(hl[hi_idx++], C.off_mid[l], 2, 0, C.n_out[l], C.off_out[l], 1, res_off, res_mul, res_add, 1)) return false;   // conv2
        x_off = C.off_out[l];
        cin = Cc;
    }
    P->n_layers = li;
    P->a_bytes = a_max;
    P->b_bytes = b_max;
    return tcn_umma_smem_bytes(*P) <= (size_t)224 * 1024;
}

__global__ void __launch_bounds__(kTuNT, 1)
tcn_cone_umma_kernel(const float* __restrict__ mel_tm, long long mel_win_stride, MelRingRef ring, long long n_windows,
                     TcnUmmaParams P, const uint4* __restrict__ wq, float* __restrict__ feat) {
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* act = reinterpret_cast<float*>(smem);
    unsigned char* a_s = smem + sizeof(float) * (size_t)P.per_window * kTuWT;
    unsigned char* b_s = a_s + P.a_bytes + 2048;
    uint64_t* bar = reinterpret_cast<uint64_t*>(b_s + P.b_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int pw = P.per_window;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_addr = smem_u32(a_s), b_addr = smem_u32(b_s);
    uint32_t phase = 0;

    for (long long w0 = (long long)blockIdx.x * kTuWT; w0 < n_windows; w0 += (long long)gridDim.x * kTuWT) {
        const int nw = (int)((n_windows - w0 < kTuWT) ? (n_windows - w0) : kTuWT);
        // ---- the last n_in frames of each window, [pos][mel] rows -----------------------------------------------
        const int n_in_f = P.n_in * P.c_in;
        if (ring.ring != nullptr) {
            for (int i = tid; i < nw * n_in_f; i += kTuNT) {
                const int w = i / n_in_f, r = i - w * n_in_f;
                const int pos = r % P.n_in, m = r / P.n_in;
                const long long s = ring.stream(w0 + w);
                const int head = smel_slot(ring.count[s] / SMel::HOP - 3 + 1);
                act[(size_t)w * pw + P.off_in + pos * P.c_in + m] =
                    ring.ring[s * SMel::STREAM_FLOATS + m * SMel::ROW + head + (P.T - P.n_in) + pos];
            }
        } else {
            for (int i = tid; i < nw * n_in_f; i += kTuNT) {
                const int w = i / n_in_f, r = i - w * n_in_f;
                act[(size_t)w * pw + P.off_in + r] = mel_tm[(w0 + w) * mel_win_stride + (size_t)(P.T - P.n_in) * P.c_in + r];
            }
        }
        __syncthreads();
        for (int li = 0; li < P.n_layers; ++li) {
            const TuLayer L = P.layers[li];
            const int rows = nw * L.n_pos;
            const uint32_t lbo_a = (uint32_t)L.rows8 * 16, hl_a = (uint32_t)L.kgs * lbo_a;
            // ---- A operand: (row, K group) -> 8 channels of one tap, bf16 hi / lo -----------------------------------
            for (int i = tid; i < L.kgs * rows; i += kTuNT) {
                const int kg = i / rows, r = i - kg * rows;
                const int w = r / L.n_pos, p = r - w * L.n_pos;
                const int k0 = kg * 8, j = k0 / L.Cin, ic0 = k0 - j * L.Cin;
                uint4 hv = make_uint4(0, 0, 0, 0), lv = hv;
                if (j < L.taps) {
                    const float* src = act + (size_t)w * pw + L.in_off + (size_t)(L.in_mul * p + L.in_add + j) * L.Cin + ic0;
                    const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
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
                unsigned char* dst = a_s + (size_t)kg * lbo_a + (size_t)r * 16;
                *reinterpret_cast<uint4*>(dst) = hv;
                *reinterpret_cast<uint4*>(dst + hl_a) = lv;
            }
            // ---- weight chunks -> MMAs ---------------------------------------------------------------------------------
            for (int c = 0; c < L.nchunks; ++c) {
                const TuChunk Ck = P.chunks[L.chunk0 + c];
                const int n16 = 2 * Ck.kgs * Ck.ncols;                       // uint4 in this chunk (hi + lo)
                const uint4* src = wq + Ck.w_off16;
                for (int i = tid; i < n16; i += kTuNT) reinterpret_cast<uint4*>(b_s)[i] = __ldg(src + i);
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after();
                    const uint32_t lbo_b = (uint32_t)Ck.ncols * 16, hl_b = (uint32_t)Ck.kgs * lbo_b;
                    const uint32_t idesc = umma_idesc_bf16(128, Ck.ncols);
                    const uint32_t d_tmem = tmem_base + (uint32_t)Ck.n0;
                    // descriptors = one base per operand + a multiple of 16 bytes in the start-address field
                    uint64_t da_h = umma_desc_noswz(a_addr + (uint32_t)Ck.kg0 * lbo_a, lbo_a, 128);
                    uint64_t db_h = umma_desc_noswz(b_addr, lbo_b, 128);
                    const uint64_t a_lo = (uint64_t)(hl_a >> 4), b_lo = (uint64_t)(hl_b >> 4);
                    const uint64_t a_step = (uint64_t)((2 * lbo_a) >> 4), b_step = (uint64_t)((2 * lbo_b) >> 4);
                    for (int ks = 0; ks < Ck.kgs / 2; ++ks, da_h += a_step, db_h += b_step) {
                        umma_bf16(d_tmem, da_h, db_h, idesc, (Ck.kg0 | ks) != 0);
                        umma_bf16(d_tmem, da_h + a_lo, db_h, idesc, 1);
                        umma_bf16(d_tmem, da_h, db_h + b_lo, idesc, 1);
                    }
                    umma_commit(bar);
                }
                mbar_wait(bar, phase);
                phase ^= 1;
                tc_fence_after();
            }
            // ---- epilogue: TMEM lane = row; warp -> (lane quarter, 32-column chunks) ----------------------------------
            {
                const int q = warp & 3;
                const int r = q * 32 + lane;
                const int w = r / L.n_pos, p = r - w * L.n_pos;
                for (int ch = warp >> 2; ch < L.Cout / 32; ch += 2) {
                    float v[32];
                    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
                    if (r < rows) {
                        const float* rr = L.res_off >= 0
                                              ? act + (size_t)w * pw + L.res_off + (size_t)(L.res_mul * p + L.res_add) * L.Cout + ch * 32
                                              : nullptr;
                        float* o = act + (size_t)w * pw + L.out_off + (size_t)p * L.Cout + ch * 32;
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            float x[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float t = v[4 * j4 + e] + __ldg(L.bias + ch * 32 + 4 * j4 + e);
                                if (L.relu) t = fmaxf(t, 0.0f);
                                x[e] = t;
                            }
                            if (rr != nullptr) {
                                const float4 rv = *reinterpret_cast<const float4*>(rr + 4 * j4);
                                x[0] += rv.x; x[1] += rv.y; x[2] += rv.z; x[3] += rv.w;
                                if (L.res_relu) {
#pragma unroll
                                    for (int e = 0; e < 4; ++e) x[e] = fmaxf(x[e], 0.0f);
                                }
                            }
                            *reinterpret_cast<float4*>(o + 4 * j4) = make_float4(x[0], x[1], x[2], x[3]);
                        }
                    }
                }
                tc_fence_before();
            }
            __syncthreads();
        }
        for (int i = tid; i < nw * P.c_last; i += kTuNT) {
            const int w = i / P.c_last, c = i - w * P.c_last;
            feat[(w0 + w) * (long long)P.c_last + c] = act[(size_t)w * pw + P.last_off + c];
        }
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 128);
    }
}

}  // namespace nww
