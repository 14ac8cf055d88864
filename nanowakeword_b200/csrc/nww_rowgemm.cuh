// nww_rowgemm.cuh — dense layers over many rows on the FP32 pipes with the register-tiled, shared-memory
// staged row GEMM of nww_tcn.cuh:  out[r][n] = post(bias[n] + sum_k A[r][k] W[k][n]).
// Used for the CRNN head's GRU input projections (x W_ih^T + b_ih over all 12 steps at once,
// reference CRNNModel architectures.py:242-282 / torch.nn.GRU) and, inside gru2_kernel, for the recurrent
// product h W_hh^T of every step.
#pragma once

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

}  // namespace nww
