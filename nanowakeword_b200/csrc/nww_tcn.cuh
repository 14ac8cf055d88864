// nww_tcn.cuh — the TCN head as one fused kernel over the DEPENDENCY CONE of the last time step.
//
// Reference: TCNModel / TemporalBlock, nanowakeword/modules/architectures.py:290-362.  Each block is
//   conv1(k=3, dilation d, left pad (k-1)d, chomp) -> ReLU -> conv2(same) -> ReLU -> ReLU(out + res)
// with res = x or a 1x1 "downsample" conv when the channel count changes (:306, :327), d = 2^level,
// and the model reads ONLY the last time step (:358).  As written that is 26.8 MFLOP per window; the
// last step depends on 1 + 4 (2^L - 1) input frames (29 of 98 for L = 3) and, per layer, only on an
// arithmetic progression of positions.  Walking the blocks backwards from {T-1}:
//   block output positions: c values, step 2d   (c = 1 for the last block)
//   conv1 output positions: 2c + 1, step d      (tap j of conv2 reads index 2p + j)
//   block input  positions: 2c + 3, step d      (tap j of conv1 reads index p + j; the residual of output
//                                                p reads input index 2p + 4)
// and the block input list is the previous block's output list.  For channels [64, 64, 128] that is
// 29 mel frames -> 27 -> 13 -> 11 -> 5 -> 3 -> 1 positions and 0.73 M MACs: every MAC that feeds the
// score is done exactly once and exactly as the reference orders the taps, nothing else.
//
// One CTA owns WT windows end to end; activations stay in shared memory as [window][position][channel]
// rows; each layer is a small GEMM (rows = window x position, K = 3 Cin, N = Cout) on the FP32 pipes:
// a thread owns RM x RN outputs, reads activations as 128-bit shared loads (4 input channels) and
// weights as 128-bit read-only loads shared by the whole CTA through L1.
#pragma once

#include "nww_common.cuh"

#ifndef NWW_CPUSIM
#ifndef NWW_DYN_SMEM
#define NWW_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#endif
#endif

namespace nww {

constexpr int kTcnMaxLevels = 4;
constexpr int kTcnWT = 8;            // windows per CTA tile
constexpr int kTcnNT = 256;
constexpr int kTcnRM = 4;            // rows per thread tile

struct TcnConeParams {
    int levels, c_in, T;                         // T = frames per window in the mel input
    int ch[kTcnMaxLevels];
    const float* w1[kTcnMaxLevels];              // [3][Cin][C]
    const float* b1[kTcnMaxLevels];
    const float* w2[kTcnMaxLevels];              // [3][C][C]
    const float* b2[kTcnMaxLevels];
    const float* wd[kTcnMaxLevels];              // [Cin][C] or null (identity residual)
    const float* bd[kTcnMaxLevels];
    // shared-memory plan (floats per window): input, then per level mid / out
    int n_in;                                    // input positions (2 c0 + 3 of level 0)
    int off_in, off_mid[kTcnMaxLevels], off_out[kTcnMaxLevels], per_window;
    int n_mid[kTcnMaxLevels], n_out[kTcnMaxLevels];
};

// positions of the cone, host side.  Returns false when the cone does not fit in T frames.
inline bool tcn_plan(TcnConeParams* P) {
    int c = 1;
    int n_out[kTcnMaxLevels], n_mid[kTcnMaxLevels], n_inp[kTcnMaxLevels];
    for (int i = P->levels - 1; i >= 0; --i) {
        n_out[i] = c;
        n_mid[i] = 2 * c + 1;
        n_inp[i] = 2 * c + 3;
        c = n_inp[i];
    }
    if (1 + 4 * ((1 << P->levels) - 1) > P->T) return false;
    P->n_in = n_inp[0];
    int off = 0;
    P->off_in = off;
    off += P->n_in * P->c_in;
    for (int i = 0; i < P->levels; ++i) {
        P->n_mid[i] = n_mid[i];
        P->n_out[i] = n_out[i];
        P->off_mid[i] = off;
        off += n_mid[i] * P->ch[i];
        P->off_out[i] = off;
        off += n_out[i] * P->ch[i];
    }
    P->per_window = (off + 3) & ~3;
    return true;
}

// out[w][p][oc] = post( b[oc] + sum_{j < taps, ic} W[j][ic][oc] * in[w][in_mul * p + j][ic] )
//   RES = 0: ReLU.   RES = 1: ReLU( ReLU(.) + in_res[w][2 p + 4][oc] )   (identity residual)
//   RES = 2: ReLU( ReLU(.) + bd[oc] + sum_ic Wd[ic][oc] * in_res[w][2 p + 4][ic] )   (1x1 downsample)
template <int RN>
__device__ __forceinline__ void tcn_layer(const float* __restrict__ in, int in_pitch /*floats per window*/, int Cin, int in_mul,
                                          const float* __restrict__ W, const float* __restrict__ bias, int taps,
                                          float* __restrict__ out, int out_pitch, int Cout, int n_pos, int n_win, int res_mode,
                                          const float* __restrict__ in_res, int res_pitch, int Cres,
                                          const float* __restrict__ Wd, const float* __restrict__ bd, int tid) {
    const int tn = tid & 15, tm = tid >> 4;
    const int oc0 = tn * RN;
    if (oc0 >= Cout) return;
    const int rows = n_win * n_pos;
    for (int r0 = tm * kTcnRM; r0 < rows; r0 += (kTcnNT / 16) * kTcnRM) {
        float acc[kTcnRM][RN];
        const float* arow[kTcnRM];
        int rw[kTcnRM], rp[kTcnRM];
#pragma unroll
        for (int i = 0; i < kTcnRM; ++i) {
            const int r = (r0 + i < rows) ? r0 + i : rows - 1;          // clamp: duplicates are computed, not stored
            rw[i] = r / n_pos;
            rp[i] = r - rw[i] * n_pos;
            arow[i] = in + (size_t)rw[i] * in_pitch + (size_t)(in_mul * rp[i]) * Cin;
#pragma unroll
            for (int c = 0; c < RN; ++c) acc[i][c] = __ldg(bias + oc0 + c);
        }
        for (int j = 0; j < taps; ++j) {
            const float* wj = W + (size_t)j * Cin * Cout + oc0;
            for (int ic = 0; ic < Cin; ic += 4) {
                float4 a[kTcnRM];
#pragma unroll
                for (int i = 0; i < kTcnRM; ++i) a[i] = *reinterpret_cast<const float4*>(arow[i] + (size_t)j * Cin + ic);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float wv[RN];
#pragma unroll
                    for (int c4 = 0; c4 < RN / 4; ++c4) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(wj + (size_t)(ic + q) * Cout) + c4);
                        wv[4 * c4] = t.x; wv[4 * c4 + 1] = t.y; wv[4 * c4 + 2] = t.z; wv[4 * c4 + 3] = t.w;
                    }
#pragma unroll
                    for (int i = 0; i < kTcnRM; ++i) {
                        const float av = (q == 0) ? a[i].x : (q == 1) ? a[i].y : (q == 2) ? a[i].z : a[i].w;
#pragma unroll
                        for (int c = 0; c < RN; ++c) acc[i][c] = fmaf(av, wv[c], acc[i][c]);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < kTcnRM; ++i)
#pragma unroll
            for (int c = 0; c < RN; ++c) acc[i][c] = fmaxf(acc[i][c], 0.0f);
        if (res_mode != 0) {
            float rs[kTcnRM][RN];
            if (res_mode == 1) {
#pragma unroll
                for (int i = 0; i < kTcnRM; ++i)
#pragma unroll
                    for (int c = 0; c < RN; ++c)
                        rs[i][c] = in_res[(size_t)rw[i] * res_pitch + (size_t)(2 * rp[i] + 4) * Cres + oc0 + c];
            } else {
#pragma unroll
                for (int i = 0; i < kTcnRM; ++i)
#pragma unroll
                    for (int c = 0; c < RN; ++c) rs[i][c] = __ldg(bd + oc0 + c);
                for (int ic = 0; ic < Cres; ++ic) {
                    float wv[RN];
#pragma unroll
                    for (int c4 = 0; c4 < RN / 4; ++c4) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(Wd + (size_t)ic * Cout + oc0) + c4);
                        wv[4 * c4] = t.x; wv[4 * c4 + 1] = t.y; wv[4 * c4 + 2] = t.z; wv[4 * c4 + 3] = t.w;
                    }
#pragma unroll
                    for (int i = 0; i < kTcnRM; ++i) {
                        const float av = in_res[(size_t)rw[i] * res_pitch + (size_t)(2 * rp[i] + 4) * Cres + ic];
#pragma unroll
                        for (int c = 0; c < RN; ++c) rs[i][c] = fmaf(av, wv[c], rs[i][c]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < kTcnRM; ++i)
#pragma unroll
                for (int c = 0; c < RN; ++c) acc[i][c] = fmaxf(acc[i][c] + rs[i][c], 0.0f);
        }
#pragma unroll
        for (int i = 0; i < kTcnRM; ++i) {
            if (r0 + i >= rows) continue;
            float* o = out + (size_t)rw[i] * out_pitch + (size_t)rp[i] * Cout + oc0;
#pragma unroll
            for (int c4 = 0; c4 < RN / 4; ++c4)
                reinterpret_cast<float4*>(o)[c4] = make_float4(acc[i][4 * c4], acc[i][4 * c4 + 1], acc[i][4 * c4 + 2], acc[i][4 * c4 + 3]);
        }
    }
}

__device__ __forceinline__ void tcn_layer_any(int Cout, const float* in, int in_pitch, int Cin, int in_mul, const float* W,
                                              const float* bias, int taps, float* out, int out_pitch, int n_pos, int n_win,
                                              int res_mode, const float* in_res, int res_pitch, int Cres, const float* Wd,
                                              const float* bd, int tid) {
    if (Cout <= 64)
        tcn_layer<4>(in, in_pitch, Cin, in_mul, W, bias, taps, out, out_pitch, Cout, n_pos, n_win, res_mode, in_res, res_pitch,
                     Cres, Wd, bd, tid);
    else
        tcn_layer<8>(in, in_pitch, Cin, in_mul, W, bias, taps, out, out_pitch, Cout, n_pos, n_win, res_mode, in_res, res_pitch,
                     Cres, Wd, bd, tid);
}

// mel_tm: time-major log-mel, window w at mel_tm + w * mel_win_stride, frame t at + t * c_in (only the last n_in
// frames are read).  feat: [n][C_last].
__global__ void __launch_bounds__(kTcnNT, 1)
tcn_cone_kernel(const float* __restrict__ mel_tm, long long mel_win_stride, long long n_windows, TcnConeParams P,
                float* __restrict__ feat) {
    NWW_DYN_SMEM(smem);
    float* act = reinterpret_cast<float*>(smem);
    const int tid = threadIdx.x;
    const int pw = P.per_window;
    const int c_last = P.ch[P.levels - 1];
    for (long long w0 = (long long)blockIdx.x * kTcnWT; w0 < n_windows; w0 += (long long)gridDim.x * kTcnWT) {
        const int nw = (int)((n_windows - w0 < kTcnWT) ? (n_windows - w0) : kTcnWT);
        __syncthreads();
        // the last n_in frames of each window, [pos][mel] rows
        const int n_in_f = P.n_in * P.c_in;
        for (int i = tid; i < nw * n_in_f; i += kTcnNT) {
            const int w = i / n_in_f, r = i - w * n_in_f;
            act[(size_t)w * pw + P.off_in + r] = mel_tm[(w0 + w) * mel_win_stride + (size_t)(P.T - P.n_in) * P.c_in + r];
        }
        __syncthreads();
        const float* x = act + P.off_in;
        int cin = P.c_in;
        for (int l = 0; l < P.levels; ++l) {
            const int C = P.ch[l];
            float* mid = act + P.off_mid[l];
            float* out = act + P.off_out[l];
            tcn_layer_any(C, x, pw, cin, 1, P.w1[l], P.b1[l], 3, mid, pw, P.n_mid[l], nw, 0, nullptr, 0, 0, nullptr, nullptr, tid);
            __syncthreads();
            tcn_layer_any(C, mid, pw, C, 2, P.w2[l], P.b2[l], 3, out, pw, P.n_out[l], nw, P.wd[l] ? 2 : 1, x, pw, cin, P.wd[l],
                          P.bd[l], tid);
            __syncthreads();
            x = out;
            cin = C;
        }
        for (int i = tid; i < nw * c_last; i += kTcnNT) {
            const int w = i / c_last, c = i - w * c_last;
            feat[(w0 + w) * (long long)c_last + c] = x[(size_t)w * pw + c];
        }
    }
}

}  // namespace nww
