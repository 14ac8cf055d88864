// nww_tail.cuh — stage B, the "dense tail": a chain of Linear(+LayerNorm)(+activation) layers
// applied to a tile of windows, ending in the classifier and the sigmoid
// (reference: Net/FCNBlock architectures.py:102-126, fc layers of the other heads,
//  Model.classifier modules/model.py:291-296, sigmoid + view(-1,1,1) _export/onnx.py:169-172).
//
// FP32 CUDA-core version: one CTA owns TM windows; thread (n, half) accumulates TM/2 windows
// for output feature n, streaming its weight row from global/L2 in 128-byte pieces while the
// activations sit in shared memory (broadcast reads).
#pragma once

#include "nww_common.cuh"

#ifndef NWW_CPUSIM
#ifndef NWW_DYN_SMEM
#define NWW_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#endif
#endif

namespace nww {

enum TailPost : int { POST_NONE = 0, POST_ACT = 1, POST_LN_ACT = 2 };

struct TailLayer {
    const float* W;      // [N][K] row-major (torch Linear layout), BatchNorm already folded
    const float* b;      // [N]
    const float* ln_g;   // [N] or null
    const float* ln_b;   // [N] or null
    int K, N, post;
};

struct TailParams {
    TailLayer layers[kMaxTailLayers];
    int n_layers;
    int act;
    int max_width;       // max over layers of N and of K for layers >= 1 (shared-memory row pitch)
    int raw_out;         // 1: write the last layer's [rows][N] output instead of sigmoid scores
    int x_row_mul;       // input row r of the first layer lives at feat[(r * x_row_mul + x_row_off) * K]
    int x_row_off;       //   (0 is read as 1: plain row-major input)
    // Optional pre-stage: the first activations are not read from `feat` but rebuilt from the
    // split-K partial sums of the tensor-core GEMM (nww_gemm_tc.cuh):
    //   x[m][n] = post( sum_{s < pre_splits} pre_part[s][m][n] + pre_b[n] ),  s in increasing order.
    const float* pre_part;   // [pre_splits][pre_mpad][pre_N] or null
    const float* pre_b;
    const float* pre_g;      // LayerNorm terms when pre_post == POST_LN_ACT
    const float* pre_beta;
    int pre_splits, pre_mpad, pre_N, pre_post;
};

// LayerNorm over N (biased variance, eps 1e-5) then activation on a [TM][pitch] tile: one warp per row.
__device__ __forceinline__ void tail_layernorm_act(float* tile, int pitch, int rows, int N, const float* g,
                                                   const float* b, int act, int tid, int nthreads) {
    const int warp = tid >> 5, lane = tid & 31;
    for (int m = warp; m < rows; m += nthreads / 32) {
        float* row = tile + (size_t)m * pitch;
        float s = 0.0f;
        for (int n = lane; n < N; n += 32) s += row[n];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mu = s / (float)N;
        float v = 0.0f;
        for (int n = lane; n < N; n += 32) {
            const float d = row[n] - mu;
            v = fmaf(d, d, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const float rstd = 1.0f / sqrtf(v / (float)N + 1e-5f);
        for (int n = lane; n < N; n += 32) row[n] = apply_act((row[n] - mu) * rstd * g[n] + b[n], act);
    }
}

constexpr int kTailTM = 32;      // windows per CTA (wide first layers: weight rows are re-used by 16 windows per thread)
constexpr int kTailTMSmall = 8;  // windows per CTA for chains of small layers (more CTAs, less latency)
constexpr int kTailKC = 32;      // K chunk staged in shared memory
constexpr int kTailNT = 256;

__host__ __device__ inline size_t tail_smem_bytes(int max_width, int tm = kTailTM) {
    // two ping-pong activation tiles [TM][max_width] + one K chunk [TM][KC]
    return sizeof(float) * ((size_t)2 * tm * max_width + (size_t)tm * kTailKC);
}

template <int TM>
__global__ void __launch_bounds__(kTailNT, 2)
tail_kernel_t(const float* __restrict__ feat, long long n_windows, TailParams P, float* __restrict__ scores,
            float* __restrict__ logits /* nullable */, float* __restrict__ emb_dump /* nullable: input of classifier */) {
    NWW_DYN_SMEM(smem);
    float* actA = reinterpret_cast<float*>(smem);
    float* actB = actA + (size_t)TM * P.max_width;
    float* xch = actB + (size_t)TM * P.max_width;
    const int tid = threadIdx.x;
    const int lane_n = tid & 127;
    const int half = tid >> 7;                       // 0/1 -> windows [0,16) / [16,32)
    constexpr int MH = TM / 2;
    const long long xmul = P.x_row_mul > 0 ? P.x_row_mul : 1;

    for (long long w0 = (long long)blockIdx.x * TM; w0 < n_windows; w0 += (long long)gridDim.x * TM) {
        const int mt = (n_windows - w0 < TM) ? (int)(n_windows - w0) : TM;
        float* cur = actA;
        float* nxt = actB;
        const bool pre = P.pre_part != nullptr;
        if (pre) {
            for (int i = tid; i < TM * P.pre_N; i += kTailNT) {
                const int m = i / P.pre_N, n = i - m * P.pre_N;
                float v = 0.0f;
                if (m < mt) {
                    const float* src = P.pre_part + (size_t)(w0 + m) * P.pre_N + n;
                    for (int sidx = 0; sidx < P.pre_splits; ++sidx) v += src[(size_t)sidx * P.pre_mpad * P.pre_N];
                    v += P.pre_b[n];
                    if (P.pre_post == POST_ACT) v = apply_act(v, P.act);
                }
                cur[(size_t)m * P.max_width + n] = v;
            }
            __syncthreads();
            if (P.pre_post == POST_LN_ACT) {
                tail_layernorm_act(cur, P.max_width, TM, P.pre_N, P.pre_g, P.pre_beta, P.act, tid, kTailNT);
                __syncthreads();
            }
            if (emb_dump != nullptr && P.n_layers == 2) {           // the pre-stage layer is the backbone's last layer
                for (int i = tid; i < mt * P.pre_N; i += kTailNT)
                    emb_dump[(w0 + i / P.pre_N) * (long long)P.pre_N + i % P.pre_N] = cur[(size_t)(i / P.pre_N) * P.max_width + i % P.pre_N];
            }
        }
        for (int li = 0; li < P.n_layers; ++li) {
            const TailLayer L = P.layers[li];
            const bool from_feat = (li == 0) && !pre;
            const int pitch_in = from_feat ? 0 : P.max_width;
            for (int nb = 0; nb < L.N; nb += 128) {
                const int n = nb + lane_n;
                const bool nvalid = n < L.N;
                float acc[MH];
#pragma unroll
                for (int m = 0; m < MH; ++m) acc[m] = 0.0f;
                const float* wrow = L.W + (size_t)(nvalid ? n : 0) * L.K;
                for (int k0 = 0; k0 < L.K; k0 += kTailKC) {
                    const int kc = (L.K - k0 < kTailKC) ? (L.K - k0) : kTailKC;
                    const float* xs;
                    int xpitch;
                    if (from_feat) {
                        __syncthreads();             // previous chunk fully consumed
                        for (int i = tid; i < TM * kTailKC; i += kTailNT) {
                            const int m = i / kTailKC, k = i - m * kTailKC;
                            xch[i] = (m < mt && k < kc) ? feat[((w0 + m) * xmul + P.x_row_off) * (long long)L.K + k0 + k] : 0.0f;
                        }
                        __syncthreads();
                        xs = xch;
                        xpitch = kTailKC;
                    } else {
                        xs = cur + k0;
                        xpitch = pitch_in;
                    }
                    if (nvalid) {
                        if (kc == kTailKC && (((size_t)(wrow + k0)) & 15) == 0) {
                            float wv[kTailKC];
#pragma unroll
                            for (int q = 0; q < kTailKC / 4; ++q) {
                                const float4 v = __ldg(reinterpret_cast<const float4*>(wrow + k0) + q);
                                wv[4 * q] = v.x; wv[4 * q + 1] = v.y; wv[4 * q + 2] = v.z; wv[4 * q + 3] = v.w;
                            }
#pragma unroll
                            for (int m = 0; m < MH; ++m) {
                                const float* xr = xs + (size_t)(half * MH + m) * xpitch;
                                float s = acc[m];
#pragma unroll
                                for (int k = 0; k < kTailKC; ++k) s = fmaf(xr[k], wv[k], s);
                                acc[m] = s;
                            }
                        } else {
                            for (int k = 0; k < kc; ++k) {
                                const float wk = __ldg(wrow + k0 + k);
#pragma unroll
                                for (int m = 0; m < MH; ++m)
                                    acc[m] = fmaf(xs[(size_t)(half * MH + m) * xpitch + k], wk, acc[m]);
                            }
                        }
                    }
                }
                if (nvalid) {
                    const float bias = L.b ? L.b[n] : 0.0f;
#pragma unroll
                    for (int m = 0; m < MH; ++m) {
                        float v = acc[m] + bias;
                        if (L.post == POST_ACT) v = apply_act(v, P.act);
                        nxt[(size_t)(half * MH + m) * P.max_width + n] = v;
                    }
                }
            }
            __syncthreads();
            if (L.post == POST_LN_ACT) {
                tail_layernorm_act(nxt, P.max_width, TM, L.N, L.ln_g, L.ln_b, P.act, tid, kTailNT);
                __syncthreads();
            }
            if (emb_dump != nullptr && li == P.n_layers - 3) {      // output of the backbone's last layer
                for (int i = tid; i < mt * L.N; i += kTailNT)
                    emb_dump[(w0 + i / L.N) * (long long)L.N + i % L.N] = nxt[(size_t)(i / L.N) * P.max_width + i % L.N];
            }
            float* t = cur; cur = nxt; nxt = t;
        }
        if (P.raw_out) {
            const int NL = P.layers[P.n_layers - 1].N;
            for (int i = tid; i < mt * NL; i += kTailNT)
                scores[(w0 + i / NL) * (long long)NL + i % NL] = cur[(size_t)(i / NL) * P.max_width + i % NL];
            __syncthreads();
            continue;
        }
        // last layer has N == 1: logit in cur[m * max_width]
        for (int m = tid; m < mt; m += kTailNT) {
            const float z = cur[(size_t)m * P.max_width];
            if (logits) logits[w0 + m] = z;
            scores[w0 + m] = sigmoidf_acc(z);
        }
        __syncthreads();
    }
}

// the historical name: 32 windows per CTA
#define tail_kernel tail_kernel_t<kTailTM>

}  // namespace nww
