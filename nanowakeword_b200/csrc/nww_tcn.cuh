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
#include "nww_stream_mel.cuh"

#ifndef NWW_CPUSIM
#ifndef NWW_DYN_SMEM
#define NWW_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#endif
#endif

namespace nww {

constexpr int kTcnKC = 32;            // weight rows per staged chunk of the row GEMM
constexpr int kTcnMaxLevels = 4;
constexpr int kTcnWTMax = 8;         // windows per CTA tile (upper bound; the plan picks what fits in shared memory)
constexpr int kTcnNT = 256;

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
    int off_in, off_mid[kTcnMaxLevels], off_out[kTcnMaxLevels], off_res, per_window;
    int n_mid[kTcnMaxLevels], n_out[kTcnMaxLevels];
    int wt;                                      // windows per CTA tile
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
    P->off_res = off;                            // downsampled residual of the block being computed
    int max_res = 0;
    for (int i = 0; i < P->levels; ++i) max_res = n_out[i] * P->ch[i] > max_res ? n_out[i] * P->ch[i] : max_res;
    off += max_res;
    P->per_window = (off + 3) & ~3;
    P->wt = 0;
    for (int wt = kTcnWTMax; wt >= 1; --wt)
        if (sizeof(float) * ((size_t)2 * kTcnKC * 128 + (size_t)P->per_window * wt) <= (size_t)220 * 1024) {
            P->wt = wt;
            break;
        }
    return P->wt > 0;
}

// ---- one layer as a small GEMM with the weights streamed through shared memory -------------------------
//   out[w][p][oc] = post( b[oc] + sum_{k < K} W[k][oc] * in[w][in_mul * p + in_off][k] )        K = taps * Cin
// A row of the A operand is CONTIGUOUS: activations are [position][channel] rows, so the taps j = 0..2 of
// position q are the 3 Cin floats that start at in[w][q].  W arrives in chunks of kTcnKC rows by cp.async,
// double-buffered; thread (tn, tm) keeps RN = Cout / 16 columns of up to kTcnRows rows (r = tm + 16 i) in
// registers across the whole K loop.
//   post: relu ? max(., 0) : identity;  then, with res != nullptr,  max(. + res[w][p][oc], 0)
//   (res rows have pitch res_pitch per window and Cout per position, positions res_mul * p + res_off).
constexpr int kTcnRows = 14;          // rows per thread at 4 columns (16 x 14 = 224 rows per pass); half of that at 8 columns
constexpr int kTcnWBuf = kTcnKC * 128;   // floats per weight buffer (Cout <= 128)

#ifndef NWW_CPUSIM
__device__ __forceinline__ void tcn_cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void tcn_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tcn_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#else
inline void tcn_cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
inline void tcn_cp_commit() {}
template <int N> inline void tcn_cp_wait() {}
#endif

struct TcnLayerArgs {
    const float* in; int in_pitch, in_mul, in_off;      // A rows: in + w * in_pitch + (in_mul * p + in_off) * Cin
    int Cin, K;                                         // K = taps * Cin (multiple of 4)
    const float* W; const float* bias;                  // W [K][ldw], columns [0, Cout) of it are used
    float* out; int out_pitch;                          // out + w * out_pitch + p * ldo
    int Cout, n_pos, n_win, relu;                       // relu: 0 none, 1 ReLU, 2 + a = activation a (nww_common.cuh)
    const float* res; int res_pitch, res_mul, res_off;  // nullable; rows of ldr floats
    int ldw, ldo, ldr;                                  // leading dimensions; 0 = Cout
    int res_relu;                                       // 1: max(. + res, 0) (TCN)   0: . + res (BcResNet)
};

template <int RN, int ROWS>
__device__ __forceinline__ void tcn_layer(const TcnLayerArgs& L, float* __restrict__ wbuf /* [2][kTcnWBuf] shared */, int tid) {
    const int tn = tid & 15, tm = tid >> 4;
    // a thread's columns: group c4 (of 4 columns) starts at c4 * 64 + tn * 4 — the 16 lanes of a half-warp read 256
    // contiguous bytes of a weight row per group, so the 128-bit shared loads are conflict-free for 64 and 128 columns
    const int oc0 = tn * 4;
    const bool col_ok = oc0 < L.Cout;
    const int rows = L.n_win * L.n_pos;
    const int n_chunks = (L.K + kTcnKC - 1) / kTcnKC;
    const int row_f4 = L.Cout / 4;                       // float4 per weight row
    const int ldw = L.ldw ? L.ldw : L.Cout, ldo = L.ldo ? L.ldo : L.Cout, ldr = L.ldr ? L.ldr : L.Cout;
    auto stage = [&](int chunk, int buf) {
        const int k0 = chunk * kTcnKC;
        const int kc = (L.K - k0 < kTcnKC) ? (L.K - k0) : kTcnKC;
        for (int i = tid; i < kc * row_f4; i += kTcnNT) {
            const int kr = i / row_f4, c4 = i - kr * row_f4;
            tcn_cp_async16(wbuf + buf * kTcnWBuf + kr * L.Cout + 4 * c4, L.W + (size_t)(k0 + kr) * ldw + 4 * c4);
        }
        tcn_cp_commit();
    };
    for (int rbase = 0; rbase < rows; rbase += 16 * ROWS) {
        float acc[ROWS][RN];
        int aoff[ROWS];                                  // float offset of the row's A vector from L.in
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            int r = rbase + tm + 16 * i;
            if (r >= rows) r = rows - 1;                 // clamp: computed, never stored
            const int w = r / L.n_pos, p = r - w * L.n_pos;
            aoff[i] = w * L.in_pitch + (L.in_mul * p + L.in_off) * L.Cin;
#pragma unroll
            for (int c = 0; c < RN; ++c) {
                const int col = (c >> 2) * 64 + oc0 + (c & 3);
                acc[i][c] = col < L.Cout ? __ldg(L.bias + col) : 0.0f;
            }
        }
        // rows this warp really has (its two tm values differ by one row at most): the FMA blocks of the
        // other register rows are skipped, which is most of them in the small late layers
        const int left = rows - rbase - (tm & ~1);
        const int nv = left <= 0 ? 0 : ((left + 15) >> 4) < ROWS ? ((left + 15) >> 4) : ROWS;
        __syncthreads();                                 // previous users of wbuf are done
        stage(0, 0);
        for (int ch = 0; ch < n_chunks; ++ch) {
            if (ch + 1 < n_chunks) {
                stage(ch + 1, (ch + 1) & 1);
                tcn_cp_wait<1>();
            } else {
                tcn_cp_wait<0>();
            }
            __syncthreads();                             // chunk ch is visible to everybody
            const float* wb = wbuf + (ch & 1) * kTcnWBuf + oc0;
            const int k0 = ch * kTcnKC;
            const int kc = (L.K - k0 < kTcnKC) ? (L.K - k0) : kTcnKC;
            if (col_ok && nv > 0) {
                for (int kk = 0; kk < kc; kk += 4) {
                    float wv[4][RN];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int c4 = 0; c4 < RN / 4; ++c4) {
                            const float4 t = *reinterpret_cast<const float4*>(wb + (kk + q) * L.Cout + 64 * c4);
                            wv[q][4 * c4] = t.x; wv[q][4 * c4 + 1] = t.y; wv[q][4 * c4 + 2] = t.z; wv[q][4 * c4 + 3] = t.w;
                        }
#pragma unroll
                    for (int i = 0; i < ROWS; ++i) {
                        if (i >= nv) break;
                        const float4 a = *reinterpret_cast<const float4*>(L.in + aoff[i] + k0 + kk);
#pragma unroll
                        for (int c = 0; c < RN; ++c) {
                            float s = acc[i][c];
                            s = fmaf(a.x, wv[0][c], s);
                            s = fmaf(a.y, wv[1][c], s);
                            s = fmaf(a.z, wv[2][c], s);
                            s = fmaf(a.w, wv[3][c], s);
                            acc[i][c] = s;
                        }
                    }
                }
            }
            __syncthreads();                             // everybody is done with buffer ch & 1 before it is refilled
        }
        if (col_ok) {
#pragma unroll
            for (int i = 0; i < ROWS; ++i) {
                const int r = rbase + tm + 16 * i;
                if (r >= rows) continue;
                const int w = r / L.n_pos, p = r - w * L.n_pos;
                float v[RN];
#pragma unroll
                for (int c = 0; c < RN; ++c)
                    v[c] = L.relu == 0 ? acc[i][c] : L.relu == 1 ? fmaxf(acc[i][c], 0.0f) : apply_act(acc[i][c], L.relu - 2);
                const float* rr = L.res ? L.res + (size_t)w * L.res_pitch + (size_t)(L.res_mul * p + L.res_off) * ldr + oc0 : nullptr;
                float* o = L.out + (size_t)w * L.out_pitch + (size_t)p * ldo + oc0;
#pragma unroll
                for (int c4 = 0; c4 < RN / 4; ++c4) {
                    if (64 * c4 + oc0 >= L.Cout) continue;
                    if (rr != nullptr) {
                        const float4 rv = *reinterpret_cast<const float4*>(rr + 64 * c4);
                        const float r4[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            v[4 * c4 + c] = L.res_relu ? fmaxf(v[4 * c4 + c] + r4[c], 0.0f) : v[4 * c4 + c] + r4[c];
                    }
                    *reinterpret_cast<float4*>(o + 64 * c4) = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
                }
            }
        }
    }
}

__device__ __forceinline__ void tcn_layer_any(const TcnLayerArgs& L, float* wbuf, int tid) {
    if (L.Cout <= 64) tcn_layer<4, kTcnRows>(L, wbuf, tid);
    else tcn_layer<8, kTcnRows / 2>(L, wbuf, tid);
}

inline size_t tcn_cone_smem_bytes(const TcnConeParams& P) {
    return sizeof(float) * ((size_t)2 * kTcnWBuf + (size_t)P.per_window * P.wt);
}

// mel_tm: time-major log-mel, window w at mel_tm + w * mel_win_stride, frame t at + t * c_in (only the last n_in
// frames are read).  feat: [n][C_last].
__global__ void __launch_bounds__(kTcnNT, 1)
tcn_cone_kernel(const float* __restrict__ mel_tm, long long mel_win_stride, MelRingRef ring, long long n_windows,
                TcnConeParams P, float* __restrict__ feat) {
    NWW_DYN_SMEM(smem);
    float* wbuf = reinterpret_cast<float*>(smem);                       // [2][kTcnWBuf]
    float* act = wbuf + 2 * kTcnWBuf;
    const int tid = threadIdx.x;
    const int pw = P.per_window;
    const int c_last = P.ch[P.levels - 1];
    for (long long w0 = (long long)blockIdx.x * P.wt; w0 < n_windows; w0 += (long long)gridDim.x * P.wt) {
        const int nw = (int)((n_windows - w0 < P.wt) ? (n_windows - w0) : P.wt);
        __syncthreads();
        // the last n_in frames of each window, [pos][mel] rows
        const int n_in_f = P.n_in * P.c_in;
        if (ring.ring != nullptr) {
            // stream mode: straight out of the mirrored mel ring ([mel][2 T] rows, window = T slots from `head`)
            for (int i = tid; i < nw * n_in_f; i += kTcnNT) {
                const int w = i / n_in_f, r = i - w * n_in_f;
                const int pos = r % P.n_in, m = r / P.n_in;            // consecutive threads walk the time axis
                const long long s = ring.stream(w0 + w);
                const int head = smel_slot(ring.count[s] / SMel::HOP - 3 + 1);
                act[(size_t)w * pw + P.off_in + pos * P.c_in + m] =
                    ring.ring[s * SMel::STREAM_FLOATS + m * SMel::ROW + head + (P.T - P.n_in) + pos];
            }
        } else {
            for (int i = tid; i < nw * n_in_f; i += kTcnNT) {
                const int w = i / n_in_f, r = i - w * n_in_f;
                act[(size_t)w * pw + P.off_in + r] = mel_tm[(w0 + w) * mel_win_stride + (size_t)(P.T - P.n_in) * P.c_in + r];
            }
        }
        __syncthreads();
        const float* x = act + P.off_in;
        int cin = P.c_in;
        for (int l = 0; l < P.levels; ++l) {
            const int C = P.ch[l];
            float* mid = act + P.off_mid[l];
            float* out = act + P.off_out[l];
            float* tmp = act + P.off_res;
            // conv1: taps of position p start at input position p
            tcn_layer_any(TcnLayerArgs{x, pw, 1, 0, cin, 3 * cin, P.w1[l], P.b1[l], mid, pw, C, P.n_mid[l], nw, 1, nullptr, 0, 0, 0, 0, 0, 0, 1},
                          wbuf, tid);
            const float* res = x;
            int res_mul = 2, res_off = 4;
            if (P.wd[l] != nullptr) {                    // 1x1 downsample of the block input at the output positions
                tcn_layer_any(TcnLayerArgs{x, pw, 2, 4, cin, cin, P.wd[l], P.bd[l], tmp, pw, C, P.n_out[l], nw, 0, nullptr, 0, 0, 0, 0, 0, 0, 1},
                              wbuf, tid);
                res = tmp;
                res_mul = 1;
                res_off = 0;
            }
            // conv2 on the conv1 output (positions 2 p + j), then ReLU(ReLU(.) + res)
            tcn_layer_any(TcnLayerArgs{mid, pw, 2, 0, C, 3 * C, P.w2[l], P.b2[l], out, pw, C, P.n_out[l], nw, 1, res, pw, res_mul, res_off, 0, 0, 0, 1},
                          wbuf, tid);
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
