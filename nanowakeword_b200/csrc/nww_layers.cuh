// nww_layers.cuh — layer kernels for the heads that are not (yet) fused per window:
// BcResNet, CRNN-GRU, TCN and the E2E mel-CNN body.  FP32 CUDA-core kernels over
// channel-major activations kept in an L2-sized scratch arena; each thread owns a small
// register tile (8 output channels x one output position or one 2x2 pooling quad), weights
// are laid out [in-channel][tap][out-channel] so a warp reads them as uniform 128-bit loads.
//
// Reference modules (nanowakeword/modules/architectures.py):
//   conv3x3 + BN + act + MaxPool2d(2)        CRNNModel :222-230, E2E_MelSpectrogram_CNN :840-856,
//                                            BcResNetModel.init_conv :627-632
//   depthwise 3x3 / pointwise 1x1 / shortcut BcResNetBlock :620-648  (activation BEFORE the add)
//   causal dilated conv1d + chomp + residual TemporalBlock :295-328
//   bidirectional GRU, last step             CRNNModel :242-282
//   AdaptiveAvgPool2d((1,4)) as AvgPool2d    _export/onnx.py:139-147
#pragma once

#include "nww_common.cuh"

namespace nww {

constexpr int kOCT = 8;     // output channels per thread in the conv kernels

// ---------------------------------------------------------------------------------------
// 3x3, stride 1, pad 1, + bias + activation (+ 2x2 max pool, floor).  BN is pre-folded.
// in [B][Cin][H][W]   w [Cin][9][Cout]   out [B][Cout][Ho][Wo]
// ---------------------------------------------------------------------------------------
template <bool POOL>
__global__ void __launch_bounds__(256)
conv3x3_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
               float* __restrict__ out, long long B, int Cin, int Cout, int H, int W, int act) {
    const int Ho = POOL ? H / 2 : H, Wo = POOL ? W / 2 : W;
    const int groups = Cout / kOCT;
    const long long total = B * groups * Ho * Wo;
    constexpr int Q = POOL ? 4 : 1;         // conv outputs per thread and channel
    constexpr int PS = POOL ? 4 : 3;        // input patch side
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(t % Wo);
        const int y = (int)((t / Wo) % Ho);
        const int g = (int)((t / ((long long)Wo * Ho)) % groups);
        const long long b = t / ((long long)Wo * Ho * groups);
        const int y0 = (POOL ? 2 * y : y) - 1, x0 = (POOL ? 2 * x : x) - 1;
        float acc[kOCT][Q];
#pragma unroll
        for (int o = 0; o < kOCT; ++o)
#pragma unroll
            for (int q = 0; q < Q; ++q) acc[o][q] = 0.0f;
        const float* inb = in + b * (long long)Cin * H * W;
        for (int ic = 0; ic < Cin; ++ic) {
            float p[PS][PS];
            const float* src = inb + (long long)ic * H * W;
#pragma unroll
            for (int r = 0; r < PS; ++r)
#pragma unroll
                for (int c = 0; c < PS; ++c) {
                    const int yy = y0 + r, xx = x0 + c;
                    p[r][c] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(src + yy * W + xx) : 0.0f;
                }
            const float* wk = w + ((long long)ic * 9) * Cout + g * kOCT;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float4 wa = __ldg(reinterpret_cast<const float4*>(wk + (r * 3 + c) * Cout));
                    const float4 wb = __ldg(reinterpret_cast<const float4*>(wk + (r * 3 + c) * Cout) + 1);
                    const float wv[kOCT] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                    for (int o = 0; o < kOCT; ++o) {
                        acc[o][0] = fmaf(p[r][c], wv[o], acc[o][0]);
                        if (POOL) {
                            acc[o][1] = fmaf(p[r][c + 1], wv[o], acc[o][1]);
                            acc[o][2] = fmaf(p[r + 1][c], wv[o], acc[o][2]);
                            acc[o][3] = fmaf(p[r + 1][c + 1], wv[o], acc[o][3]);
                        }
                    }
                }
        }
#pragma unroll
        for (int o = 0; o < kOCT; ++o) {
            const int oc = g * kOCT + o;
            const float bv = bias ? __ldg(bias + oc) : 0.0f;
            float v = apply_act(acc[o][0] + bv, act);
            if (POOL) {
                v = fmaxf(v, apply_act(acc[o][1] + bv, act));
                v = fmaxf(v, apply_act(acc[o][2] + bv, act));
                v = fmaxf(v, apply_act(acc[o][3] + bv, act));
            }
            out[((b * Cout + oc) * Ho + y) * (long long)Wo + x] = v;
        }
    }
}

// depthwise 3x3, stride (sh, sw), pad 1, no bias / activation.  w [C][9]
__global__ void __launch_bounds__(256)
dw3x3_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out, long long B, int C, int H,
             int W, int sh, int sw) {
    const int Ho = (H - 1) / sh + 1, Wo = (W - 1) / sw + 1;
    const long long total = B * C * Ho * Wo;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(t % Wo);
        const int y = (int)((t / Wo) % Ho);
        const int c = (int)((t / ((long long)Wo * Ho)) % C);
        const long long b = t / ((long long)Wo * Ho * C);
        const float* src = in + (b * C + c) * (long long)H * W;
        const float* k = w + c * 9;
        float s = 0.0f;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int yy = y * sh - 1 + r, xx = x * sw - 1 + q;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) s = fmaf(__ldg(src + yy * W + xx), __ldg(k + r * 3 + q), s);
            }
        out[t] = s;
    }
}

// BcResNet block tail:  out = act(pointwise(dw) + pb) + (shortcut(in, stride) + sb)
// dw [B][Cin][Ho][Wo]  in [B][Cin][H][W]  pw, sc [Cin][Cout]  out [B][Cout][Ho][Wo]
__global__ void __launch_bounds__(256)
bc_pw_res_kernel(const float* __restrict__ dw, const float* __restrict__ in, const float* __restrict__ pw,
                 const float* __restrict__ pb, const float* __restrict__ sc, const float* __restrict__ sb,
                 float* __restrict__ out, long long B, int Cin, int Cout, int H, int W, int sh, int sw, int act) {
    const int Ho = (H - 1) / sh + 1, Wo = (W - 1) / sw + 1;
    const int groups = Cout / kOCT;
    const long long total = B * groups * Ho * Wo;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(t % Wo);
        const int y = (int)((t / Wo) % Ho);
        const int g = (int)((t / ((long long)Wo * Ho)) % groups);
        const long long b = t / ((long long)Wo * Ho * groups);
        float a[kOCT], r[kOCT];
#pragma unroll
        for (int o = 0; o < kOCT; ++o) a[o] = r[o] = 0.0f;
        const float* dsrc = dw + (b * Cin) * (long long)Ho * Wo + y * Wo + x;
        const float* isrc = in + (b * Cin) * (long long)H * W + (y * sh) * W + x * sw;
        for (int ic = 0; ic < Cin; ++ic) {
            const float dv = __ldg(dsrc + (long long)ic * Ho * Wo);
            const float iv = __ldg(isrc + (long long)ic * H * W);
            const float4 pa = __ldg(reinterpret_cast<const float4*>(pw + (long long)ic * Cout + g * kOCT));
            const float4 pbb = __ldg(reinterpret_cast<const float4*>(pw + (long long)ic * Cout + g * kOCT) + 1);
            const float4 sa = __ldg(reinterpret_cast<const float4*>(sc + (long long)ic * Cout + g * kOCT));
            const float4 sbb = __ldg(reinterpret_cast<const float4*>(sc + (long long)ic * Cout + g * kOCT) + 1);
            const float pv[kOCT] = {pa.x, pa.y, pa.z, pa.w, pbb.x, pbb.y, pbb.z, pbb.w};
            const float sv[kOCT] = {sa.x, sa.y, sa.z, sa.w, sbb.x, sbb.y, sbb.z, sbb.w};
#pragma unroll
            for (int o = 0; o < kOCT; ++o) {
                a[o] = fmaf(dv, pv[o], a[o]);
                r[o] = fmaf(iv, sv[o], r[o]);
            }
        }
#pragma unroll
        for (int o = 0; o < kOCT; ++o) {
            const int oc = g * kOCT + o;
            out[((b * Cout + oc) * Ho + y) * (long long)Wo + x] =
                apply_act(a[o] + __ldg(pb + oc), act) + (r[o] + __ldg(sb + oc));
        }
    }
}

// mean over the spatial plane: in [B*C][HW] -> out [B*C]; one warp per plane
__global__ void __launch_bounds__(256) gap_kernel(const float* __restrict__ in, float* __restrict__ out, long long planes, int hw) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp; p < planes; p += nwarps) {
        float s = 0.0f;
        for (int i = lane; i < hw; i += 32) s += in[p * hw + i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) out[p] = s / (float)hw;
    }
}

// AdaptiveAvgPool2d((1, OW)) in its deployed AvgPool2d form: kernel (H, W-(OW-1)*(W/OW)), stride (H, W/OW)
// in [B][C][H][W] -> out [B][C*OW]
__global__ void __launch_bounds__(256)
avgpool_row_kernel(const float* __restrict__ in, float* __restrict__ out, long long B, int C, int H, int W, int OW) {
    const int sw = W / OW, kw = W - (OW - 1) * sw;
    const long long total = B * C * OW;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % OW);
        const long long plane = t / OW;
        const float* src = in + plane * (long long)H * W + j * sw;
        float s = 0.0f;
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < kw; ++x) s += src[y * W + x];
        out[t] = s / (float)(H * kw);
    }
}

// ---------------------------------------------------------------------------------------
// TCN: causal dilated conv1d (k taps) on [B][C][T], computed only for t >= t_lo — the last
// time step is all the head reads (architectures.py:358), so every layer only needs the
// suffix of its dependency cone; results are bit-identical to computing all T positions.
//   mode 0: out = relu(conv(in) + b)
//   mode 1: out = relu(relu(conv(in) + b) + res),  res = down(res_in) + db  or  res_in
// w [k][Cin][Cout]   down [Cres][Cout]
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tcn_conv_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                const float* __restrict__ res_in, const float* __restrict__ down, const float* __restrict__ db,
                float* __restrict__ out, long long B, int Cin, int Cout, int Cres, int T, int k, int dil, int t_lo, int mode) {
    const int npos = T - t_lo;
    const long long total = B * npos * Cout;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int oc = (int)(i % Cout);
        const int t = t_lo + (int)((i / Cout) % npos);
        const long long b = i / ((long long)Cout * npos);
        const float* xb = in + b * (long long)Cin * T;
        float s = __ldg(bias + oc);
        for (int j = 0; j < k; ++j) {
            const int tt = t - (k - 1 - j) * dil;          // tap j of the chomped, left-padded conv
            if (tt < 0) continue;
            const float* wj = w + ((long long)j * Cin) * Cout + oc;
            for (int ic = 0; ic < Cin; ++ic) s = fmaf(__ldg(xb + (long long)ic * T + tt), __ldg(wj + (long long)ic * Cout), s);
        }
        s = fmaxf(s, 0.0f);
        if (mode == 1) {
            float r;
            if (down != nullptr) {
                r = __ldg(db + oc);
                const float* rb = res_in + b * (long long)Cres * T + t;
                for (int ic = 0; ic < Cres; ++ic) r = fmaf(__ldg(rb + (long long)ic * T), __ldg(down + (long long)ic * Cout + oc), r);
            } else {
                r = res_in[(b * Cout + oc) * (long long)T + t];
            }
            s = fmaxf(s + r, 0.0f);
        }
        out[(b * Cout + oc) * (long long)T + t] = s;
    }
}

// feat[b][c] = x[b][c][T-1]
__global__ void __launch_bounds__(256) last_step_kernel(const float* __restrict__ x, float* __restrict__ feat, long long B, int C, int T) {
    const long long total = B * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        feat[i] = x[i * T + (T - 1)];
}

// CRNN: conv output [B][C][H][W] -> GRU input sequence [B][W][C*H]  (view + permute, :272-276)
__global__ void __launch_bounds__(256) seq_pack_kernel(const float* __restrict__ a, float* __restrict__ seq, long long B, int C, int H, int W) {
    const long long total = B * C * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i % (C * H));
        const int wi = (int)((i / (C * H)) % W);
        const long long b = i / ((long long)C * H * W);
        seq[i] = a[(b * C * H + f) * (long long)W + wi];
    }
}

// ---------------------------------------------------------------------------------------
// GRU recurrence for a tile of 32 windows per CTA (gate order r, z, n; h0 = 0):
//   gi_f [B*S][3H] = x W_ih^T + b_ih  (precomputed by the dense kernel),  whh [H][3H], bhh [3H]
//   r = sig(gi_r + gh_r)  z = sig(gi_z + gh_z)  n = tanh(gi_n + r*gh_n)  h = (1-z) n + z h
// The reverse direction contributes its first step only (x_{S-1}, h0 = 0) to out[:, -1, :]:
//   gi_b [B][3H];  gh = b_hh_b.
// feat [B][2H] = [h_fwd(S-1) | h_bwd(first step)]
// ---------------------------------------------------------------------------------------
constexpr int kGruTM = 32;
constexpr int kGruNT = 256;
__host__ __device__ inline size_t gru_smem_bytes(int Hd) { return sizeof(float) * (size_t)kGruTM * (Hd + 3 * Hd); }

__global__ void __launch_bounds__(kGruNT)
gru_kernel(const float* __restrict__ gi_f, const float* __restrict__ gi_b, const float* __restrict__ whh,
           const float* __restrict__ bhh, const float* __restrict__ bhh_b, float* __restrict__ feat, long long B, int S, int Hd) {
    NWW_DYN_SMEM(smem);
    float* h = reinterpret_cast<float*>(smem);            // [TM][Hd]
    float* gh = h + (size_t)kGruTM * Hd;                   // [TM][3Hd]
    const int tid = threadIdx.x;
    const int lane_n = tid & 127, half = tid >> 7;
    constexpr int MH = kGruTM / 2;
    const int G = 3 * Hd;
    for (long long w0 = (long long)blockIdx.x * kGruTM; w0 < B; w0 += (long long)gridDim.x * kGruTM) {
        const int mt = (B - w0 < kGruTM) ? (int)(B - w0) : kGruTM;
        for (int i = tid; i < kGruTM * Hd; i += kGruNT) h[i] = 0.0f;
        __syncthreads();
        for (int s = 0; s < S; ++s) {
            for (int nb = 0; nb < G; nb += 128) {
                const int n = nb + lane_n;
                if (n < G) {
                    float acc[MH];
                    const float bv = __ldg(bhh + n);
#pragma unroll
                    for (int m = 0; m < MH; ++m) acc[m] = bv;
                    for (int k = 0; k < Hd; ++k) {
                        const float wv = __ldg(whh + (long long)k * G + n);
#pragma unroll
                        for (int m = 0; m < MH; ++m) acc[m] = fmaf(h[(half * MH + m) * Hd + k], wv, acc[m]);
                    }
#pragma unroll
                    for (int m = 0; m < MH; ++m) gh[(half * MH + m) * G + n] = acc[m];
                }
            }
            __syncthreads();
            for (int i = tid; i < mt * Hd; i += kGruNT) {
                const int m = i / Hd, j = i - m * Hd;
                const float* gi = gi_f + ((w0 + m) * S + s) * (long long)G;
                const float* g = gh + m * G;
                const float r = sigmoidf_acc(gi[j] + g[j]);
                const float z = sigmoidf_acc(gi[Hd + j] + g[Hd + j]);
                const float nn = tanhf(gi[2 * Hd + j] + r * g[2 * Hd + j]);
                h[m * Hd + j] = (1.0f - z) * nn + z * h[m * Hd + j];
            }
            __syncthreads();
        }
        for (int i = tid; i < mt * Hd; i += kGruNT) {
            const int m = i / Hd, j = i - m * Hd;
            feat[(w0 + m) * (long long)(2 * Hd) + j] = h[m * Hd + j];
            const float* gi = gi_b + (w0 + m) * (long long)G;
            const float r = sigmoidf_acc(gi[j] + __ldg(bhh_b + j));
            const float z = sigmoidf_acc(gi[Hd + j] + __ldg(bhh_b + Hd + j));
            const float nn = tanhf(gi[2 * Hd + j] + r * __ldg(bhh_b + 2 * Hd + j));
            feat[(w0 + m) * (long long)(2 * Hd) + Hd + j] = (1.0f - z) * nn;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// CRNN third conv on the channel-last output of cnn2_stage_kernel: conv3x3(C -> Cout, pad 1) + folded BN + act +
// MaxPool2d(2), written straight as the GRU input sequence (reference CRNNModel architectures.py:222-230, 272-276:
// view (B, C*H, W) -> permute (B, W, C*H), i.e. feature index c * Ho + h at step w).
// in [B][H*W][C]   w [C][9][Cout]   seq [B][Wo][Cout*Ho];  one thread = one pooled pixel x 8 output channels.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
crnn_conv3_seq_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                      float* __restrict__ seq, long long B, int C, int Cout, int H, int W, int act) {
    const int Ho = H / 2, Wo = W / 2, groups = Cout / kOCT;
    const long long total = B * Ho * Wo * groups;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(t % groups);
        const long long pix = t / groups;
        const int x = (int)(pix % Wo), y = (int)((pix / Wo) % Ho);
        const long long b = pix / ((long long)Wo * Ho);
        const float* src = in + b * (long long)H * W * C;
        float acc[kOCT][4];
#pragma unroll
        for (int o = 0; o < kOCT; ++o) {
            const float bv = __ldg(bias + g * kOCT + o);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[o][q] = bv;
        }
        for (int ic4 = 0; ic4 < C; ic4 += 4) {
            float4 p[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int yy = 2 * y - 1 + r, xx = 2 * x - 1 + c;
                    p[r][c] = (yy >= 0 && yy < H && xx >= 0 && xx < W)
                                  ? __ldg(reinterpret_cast<const float4*>(src + ((long long)yy * W + xx) * C + ic4))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float* wk = w + ((long long)(ic4 + k) * 9) * Cout + g * kOCT;
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float4 wa = __ldg(reinterpret_cast<const float4*>(wk + (r * 3 + c) * Cout));
                        const float4 wb = __ldg(reinterpret_cast<const float4*>(wk + (r * 3 + c) * Cout) + 1);
                        const float wv[kOCT] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                        auto comp = [&](const float4& v) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; };
                        const float i00 = comp(p[r][c]), i01 = comp(p[r][c + 1]), i10 = comp(p[r + 1][c]), i11 = comp(p[r + 1][c + 1]);
#pragma unroll
                        for (int o = 0; o < kOCT; ++o) {
                            acc[o][0] = fmaf(i00, wv[o], acc[o][0]);
                            acc[o][1] = fmaf(i01, wv[o], acc[o][1]);
                            acc[o][2] = fmaf(i10, wv[o], acc[o][2]);
                            acc[o][3] = fmaf(i11, wv[o], acc[o][3]);
                        }
                    }
            }
        }
#pragma unroll
        for (int o = 0; o < kOCT; ++o) {
            const float v = fmaxf(fmaxf(apply_act(acc[o][0], act), apply_act(acc[o][1], act)),
                                  fmaxf(apply_act(acc[o][2], act), apply_act(acc[o][3], act)));
            seq[(b * Wo + x) * (long long)(Cout * Ho) + (g * kOCT + o) * Ho + y] = v;
        }
    }
}

// channel-last pooled conv output [B][Ho*Wo][C] -> GRU sequence [B][Wo][C*Ho] (feature index c * Ho + h at step w)
__global__ void __launch_bounds__(256)
seq_pack_nhwc_kernel(const float* __restrict__ a, float* __restrict__ seq, long long B, int C, int Ho, int Wo) {
    const long long total = B * C * Ho * Wo;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int x = (int)((i / C) % Wo);
        const int y = (int)((i / ((long long)C * Wo)) % Ho);
        const long long b = i / ((long long)C * Wo * Ho);
        seq[(b * Wo + x) * (long long)(C * Ho) + c * Ho + y] = a[i];
    }
}

}  // namespace nww
