// nww_cnn.cuh — stage A of the CNN head: PCM -> log-mel -> conv1+act+pool -> conv2+act+pool,
// one window per CTA iteration, every intermediate in shared memory; the 7680-wide pooled
// activation vector goes to global memory for the dense tail (fc1 is a GEMM across windows).
//
// Reference: CNNModel, nanowakeword/modules/architectures.py:51-80
//   conv1 = Conv2d(1, 16, 3, pad 1) -> act -> MaxPool2d(2)      (40,98) -> (16,20,49)
//   conv2 = Conv2d(16, 32, 3, pad 1) -> act -> MaxPool2d(2)     -> (32,10,24)   [49 -> 24: floor]
//   flatten (c, h, w) row-major -> 7680
#pragma once

#include "nww_stage.cuh"

namespace nww {

struct CnnWeights {
    const float* w1;   // [16][9]
    const float* b1;   // [16]
    const float* w2;   // [16 ic][9 tap][32 oc]   (re-laid out by the host packer)
    const float* b2;   // [32]
};

template <typename G> struct CnnDims {
    static constexpr int F = G::N_MELS, TT = G::N_FRAMES;       // input (F, T) = (H, W)
    static constexpr int C1 = 16, H1 = F / 2, W1 = TT / 2;      // after conv1+pool
    static constexpr int C2 = 32, H2 = H1 / 2, W2 = W1 / 2;     // after conv2+pool
    static constexpr int MEL_P = TT + 2;                        // padded mel pitch
    static constexpr int MEL_ROWS = F + 2;
    static constexpr int A1_P = W1 + 2, A1_ROWS = H1 + 2;       // padded conv1 output
    static constexpr int FEAT = C2 * H2 * W2;
};

template <typename T, typename G, int NFB> struct CnnSmem {
    using D = CnnDims<G>;
    static constexpr size_t kWorkFft = sizeof(cplx<T>) * G::N_FFT * NFB;
    static constexpr size_t kA1 = sizeof(float) * D::C1 * D::A1_ROWS * D::A1_P;
    static constexpr size_t kWork = align_up(kWorkFft > kA1 ? kWorkFft : kA1, 128);   // FFT buffer, later conv1 output
    static constexpr size_t kMel = align_up(sizeof(float) * D::MEL_ROWS * D::MEL_P, 128);
    static constexpr size_t kW = align_up(sizeof(float) * (16 * 9 + 16 + 16 * 9 * 32 + 32), 128);
    static constexpr size_t kTotal = kWork + kMel + kW + PcmStager<G::CLIP>::kBytes;
};

template <typename T, typename G, int NFB, int NT>
__global__ void __launch_bounds__(NT, 1)
cnn_stage_kernel(WindowSource src, long long n_windows, FrontendTables<T> tab, CnnWeights wt, int act,
                 float* __restrict__ feat_out, float* __restrict__ mel_dump /* nullable, (F,T) */) {
    using D = CnnDims<G>;
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x;
    cplx<T>* work = reinterpret_cast<cplx<T>*>(smem);
    float* a1 = reinterpret_cast<float*>(smem);                                   // overlays the FFT buffer
    float* melp = reinterpret_cast<float*>(smem + CnnSmem<T, G, NFB>::kWork);
    float* w1s = reinterpret_cast<float*>(smem + CnnSmem<T, G, NFB>::kWork + CnnSmem<T, G, NFB>::kMel);
    float* b1s = w1s + 16 * 9;
    float* w2s = b1s + 16;
    float* b2s = w2s + 16 * 9 * 32;
    PcmStager<G::CLIP> stager;
    stager.carve(smem + CnnSmem<T, G, NFB>::kWork + CnnSmem<T, G, NFB>::kMel + CnnSmem<T, G, NFB>::kW);
    stager.init(tid);

    for (int i = tid; i < 16 * 9; i += NT) w1s[i] = wt.w1[i];
    for (int i = tid; i < 16; i += NT) b1s[i] = wt.b1[i];
    for (int i = tid; i < 16 * 9 * 32; i += NT) w2s[i] = wt.w2[i];
    for (int i = tid; i < 32; i += NT) b2s[i] = wt.b2[i];
    for (int i = tid; i < D::MEL_ROWS * D::MEL_P; i += NT) melp[i] = 0.0f;        // zero border, written once
    __syncthreads();

    long long w = blockIdx.x;
    if (w < n_windows) stager.issue(0, src.at(w), tid);
    for (int it = 0; w < n_windows; w += gridDim.x, ++it) {
        const long long wn = w + gridDim.x;
        if (wn < n_windows) stager.issue((it + 1) & 1, src.at(wn), tid);
        const int16_t* x = stager.wait(it & 1, (it >> 1) & 1, src.at(w));

        // ---- K1: log-mel into the padded (F+2, T+2) plane --------------------------------
        logmel_window<T, G, NFB, int16_t>(x, work, tab, melp + D::MEL_P + 1, D::MEL_P, 1, tid, NT);
        if (mel_dump != nullptr) {
            float* md = mel_dump + w * (long long)(D::F * D::TT);
            for (int i = tid; i < D::F * D::TT; i += NT) md[i] = melp[(i / D::TT + 1) * D::MEL_P + (i % D::TT) + 1];
        }

        // ---- conv1 (1 -> 16) + act + 2x2 max pool; zero the padded border of a1 -----------
        for (int i = tid; i < D::C1 * D::A1_ROWS * D::A1_P; i += NT) {
            const int xx = i % D::A1_P, yy = (i / D::A1_P) % D::A1_ROWS;
            if (xx == 0 || xx == D::A1_P - 1 || yy == 0 || yy == D::A1_ROWS - 1) a1[i] = 0.0f;
        }
        for (int p = tid; p < D::H1 * D::W1; p += NT) {
            const int ph = p / D::W1, pw = p - ph * D::W1;
            float in[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) in[r][c] = melp[(2 * ph + r) * D::MEL_P + 2 * pw + c];
#pragma unroll 4
            for (int oc = 0; oc < D::C1; ++oc) {
                const float* k = w1s + oc * 9;
                const float bias = b1s[oc];
                float best = -3.4e38f;
#pragma unroll
                for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 2; ++dx) {
                        float s = bias;
#pragma unroll
                        for (int r = 0; r < 3; ++r)
#pragma unroll
                            for (int c = 0; c < 3; ++c) s = fmaf(in[dy + r][dx + c], k[r * 3 + c], s);
                        best = fmaxf(best, apply_act(s, act));
                    }
                a1[(oc * D::A1_ROWS + ph + 1) * D::A1_P + pw + 1] = best;
            }
        }
        __syncthreads();

        // ---- conv2 (16 -> 32) + act + 2x2 max pool -> global feature vector ----------------
        float* fo = feat_out + w * (long long)D::FEAT;
        constexpr int OCG = 8;                                   // output channels per thread
        constexpr int NTASK2 = (D::C2 / OCG) * D::H2 * D::W2;
        for (int t = tid; t < NTASK2; t += NT) {
            const int g = t / (D::H2 * D::W2);
            const int p = t - g * (D::H2 * D::W2);
            const int ph = p / D::W2, pw = p - ph * D::W2;
            float acc[OCG][4];
#pragma unroll
            for (int o = 0; o < OCG; ++o)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[o][q] = 0.0f;
            for (int ic = 0; ic < D::C1; ++ic) {
                float in[4][4];
                const float* src = a1 + (ic * D::A1_ROWS + 2 * ph) * D::A1_P + 2 * pw;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) in[r][c] = src[r * D::A1_P + c];
                const float* wk = w2s + (ic * 9) * 32 + g * OCG;
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float4 wa = *reinterpret_cast<const float4*>(wk + (r * 3 + c) * 32);
                        const float4 wb = *reinterpret_cast<const float4*>(wk + (r * 3 + c) * 32 + 4);
                        const float wv[OCG] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                        for (int o = 0; o < OCG; ++o) {
                            acc[o][0] = fmaf(in[r][c], wv[o], acc[o][0]);
                            acc[o][1] = fmaf(in[r][c + 1], wv[o], acc[o][1]);
                            acc[o][2] = fmaf(in[r + 1][c], wv[o], acc[o][2]);
                            acc[o][3] = fmaf(in[r + 1][c + 1], wv[o], acc[o][3]);
                        }
                    }
            }
#pragma unroll
            for (int o = 0; o < OCG; ++o) {
                const int oc = g * OCG + o;
                const float bias = b2s[oc];
                float best = apply_act(acc[o][0] + bias, act);
                best = fmaxf(best, apply_act(acc[o][1] + bias, act));
                best = fmaxf(best, apply_act(acc[o][2] + bias, act));
                best = fmaxf(best, apply_act(acc[o][3] + bias, act));
                fo[(oc * D::H2 + ph) * D::W2 + pw] = best;
            }
        }
        __syncthreads();   // a1 (= FFT work buffer) is free again
    }
}

}  // namespace nww
