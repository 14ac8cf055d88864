// nww_bc.cuh — BcResNet head on channel-last activations, with the pointwise + shortcut 1x1
// convolutions (72 % of the head's time as scalar layer kernels) as shared-memory row GEMMs.
//
// Reference: BcResNetModel / BcResNetBlock, nanowakeword/modules/architectures.py:620-687
//   init: conv3x3(1 -> 32, no bias) + BN + act + MaxPool2d(2)                (40,98) -> (32,20,49)
//   block(Cin -> Cout, stride s): o = act(BN(pointwise(depthwise3x3_s(x))))  (activation BEFORE the add, :646-647)
//                                 out = o + BN(shortcut1x1_s(x))
//   blocks (32->64, s(2,2)), (64->128, s(2,2)), (128->256, s(2,1)); global average pool; fc.
// BatchNorm is folded by the packer (weights.py).
//
// Layout: activations are [window][pixel][channel] ("NHWC"), so
//   * a pixel's channels are one contiguous GEMM row: pointwise and shortcut are plain
//     [pixels x Cin] x [Cin x Cout] products, done by the register-tiled row GEMM of nww_tcn.cuh with the
//     weights streamed through shared memory (cp.async) — FP32, exact;
//   * the depthwise kernel reads / writes 128-bit channel quads and also emits the centre tap of its
//     window, which IS the strided 1x1 shortcut's input pixel (stride s, pad 1: centre = (s y, s x)).
#pragma once

#include "nww_tcn.cuh"

namespace nww {

// init conv (Cin = 1) + folded BN + act + 2x2 max pool.  mel (F, T) per window -> out [n][(F/2)*(T/2)][C0].
// w [9][C0] (tap-major), one thread = one pooled pixel x 8 output channels.
__global__ void __launch_bounds__(256)
bc_init_conv_kernel(const float* __restrict__ mel, const float* __restrict__ w, const float* __restrict__ bias,
                    float* __restrict__ out, long long n, int F, int T, int C0, int act) {
    const int H1 = F / 2, W1 = T / 2, groups = C0 / 8;
    const long long total = n * H1 * W1 * groups;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(t % groups);
        const long long pix = t / groups;
        const int x = (int)(pix % W1), y = (int)((pix / W1) % H1);
        const long long b = pix / ((long long)W1 * H1);
        const float* m = mel + b * (long long)F * T;
        float in[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int yy = 2 * y - 1 + r, xx = 2 * x - 1 + c;
                in[r][c] = (yy >= 0 && yy < F && xx >= 0 && xx < T) ? __ldg(m + yy * T + xx) : 0.0f;
            }
        float acc[8][4];
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const float bv = __ldg(bias + g * 8 + o);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[o][q] = bv;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 wa = __ldg(reinterpret_cast<const float4*>(w + (r * 3 + c) * C0 + g * 8));
                const float4 wb = __ldg(reinterpret_cast<const float4*>(w + (r * 3 + c) * C0 + g * 8) + 1);
                const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    acc[o][0] = fmaf(in[r][c], wv[o], acc[o][0]);
                    acc[o][1] = fmaf(in[r][c + 1], wv[o], acc[o][1]);
                    acc[o][2] = fmaf(in[r + 1][c], wv[o], acc[o][2]);
                    acc[o][3] = fmaf(in[r + 1][c + 1], wv[o], acc[o][3]);
                }
            }
        float v[8];
#pragma unroll
        for (int o = 0; o < 8; ++o)
            v[o] = fmaxf(fmaxf(apply_act(acc[o][0], act), apply_act(acc[o][1], act)),
                         fmaxf(apply_act(acc[o][2], act), apply_act(acc[o][3], act)));
        float4* dst = reinterpret_cast<float4*>(out + pix * C0 + g * 8);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
}

// depthwise 3x3, stride (sh, sw), pad 1, no bias / activation, channel-last.  w [C][9].
// in [n][H*W][C] -> dwo [n][Ho*Wo][C] and ctr [n][Ho*Wo][C] = in at (sh y, sw x) (the shortcut's input).
__global__ void __launch_bounds__(256)
bc_dw_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ dwo, float* __restrict__ ctr,
             long long n, int C, int H, int W, int sh, int sw) {
    const int Ho = (H - 1) / sh + 1, Wo = (W - 1) / sw + 1, c4n = C / 4;
    const long long total = n * Ho * Wo * c4n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(t % c4n);
        const long long pix = t / c4n;
        const int x = (int)(pix % Wo), y = (int)((pix / Wo) % Ho);
        const long long b = pix / ((long long)Wo * Ho);
        const float* src = in + b * (long long)H * W * C + 4 * c4;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f), centre = s;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int yy = y * sh - 1 + r, xx = x * sw - 1 + q;
                if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                const float4 v = __ldg(reinterpret_cast<const float4*>(src + ((long long)yy * W + xx) * C));
                const int tap = r * 3 + q;
                s.x = fmaf(v.x, __ldg(w + (4 * c4 + 0) * 9 + tap), s.x);
                s.y = fmaf(v.y, __ldg(w + (4 * c4 + 1) * 9 + tap), s.y);
                s.z = fmaf(v.z, __ldg(w + (4 * c4 + 2) * 9 + tap), s.z);
                s.w = fmaf(v.w, __ldg(w + (4 * c4 + 3) * 9 + tap), s.w);
                if (r == 1 && q == 1) centre = v;
            }
        reinterpret_cast<float4*>(dwo + pix * C)[c4] = s;
        reinterpret_cast<float4*>(ctr + pix * C)[c4] = centre;
    }
}

// Block tail as two row GEMMs per tile of rows:  out = act(dwo Wpw + bpw) + (ctr Wsc + bsc).
constexpr int kBcRows = 112;          // rows (pixels) per CTA tile
inline size_t bc_block_smem_bytes(int Cin) {
    return sizeof(float) * ((size_t)2 * kTcnWBuf + (size_t)2 * kBcRows * Cin + (size_t)kBcRows * 128);
}

__global__ void __launch_bounds__(kTcnNT, 1)
bc_block_gemm_kernel(const float* __restrict__ dwo, const float* __restrict__ ctr, const float* __restrict__ Wpw,
                     const float* __restrict__ bpw, const float* __restrict__ Wsc, const float* __restrict__ bsc,
                     float* __restrict__ out, long long rows, int Cin, int Cout, int act) {
    NWW_DYN_SMEM(smem);
    float* wbuf = reinterpret_cast<float*>(smem);
    float* a_dw = wbuf + 2 * kTcnWBuf;
    float* a_ct = a_dw + (size_t)kBcRows * Cin;
    float* tmp = a_ct + (size_t)kBcRows * Cin;           // [rows][<= 128] shortcut result of the current column half
    const int tid = threadIdx.x;
    for (long long r0 = (long long)blockIdx.x * kBcRows; r0 < rows; r0 += (long long)gridDim.x * kBcRows) {
        const int nr = (int)((rows - r0 < kBcRows) ? (rows - r0) : kBcRows);
        __syncthreads();
        const int n4 = nr * Cin / 4;
        const float4* gd = reinterpret_cast<const float4*>(dwo + r0 * Cin);
        const float4* gc = reinterpret_cast<const float4*>(ctr + r0 * Cin);
        for (int i = tid; i < n4; i += kTcnNT) {
            reinterpret_cast<float4*>(a_dw)[i] = __ldg(gd + i);
            reinterpret_cast<float4*>(a_ct)[i] = __ldg(gc + i);
        }
        __syncthreads();
        for (int n0 = 0; n0 < Cout; n0 += 128) {
            const int nc = (Cout - n0 < 128) ? (Cout - n0) : 128;
            // shortcut: tmp = ctr Wsc + bsc
            tcn_layer_any(TcnLayerArgs{a_ct, 0, 1, 0, Cin, Cin, Wsc + n0, bsc + n0, tmp, 0, nc, nr, 1, 0, nullptr, 0, 0, 0, Cout, nc, 0, 0},
                          wbuf, tid);
            // pointwise: out = act(dwo Wpw + bpw) + tmp
            tcn_layer_any(TcnLayerArgs{a_dw, 0, 1, 0, Cin, Cin, Wpw + n0, bpw + n0, out + r0 * Cout + n0, 0, nc, nr, 1, 2 + act, tmp, 0,
                                       1, 0, Cout, Cout, nc, 0},
                          wbuf, tid);
        }
    }
}

// global average pool, channel-last: in [n][P][C] -> out [n][C]
__global__ void __launch_bounds__(256) bc_gap_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, int P, int C) {
    const long long total = n * C;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long b = t / C;
        const int c = (int)(t - b * C);
        const float* src = in + b * (long long)P * C + c;
        float s = 0.0f;
        for (int p = 0; p < P; ++p) s += __ldg(src + (long long)p * C);
        out[t] = s / (float)P;
    }
}

}  // namespace nww
