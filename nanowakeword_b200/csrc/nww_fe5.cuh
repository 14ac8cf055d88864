// nww_fe5.cuh — front end for the REF64x101 geometry (the reference's own e2e_dnn mel: n_fft = win = 400, hop 160,
// centred frames with reflect padding, 201 bins -> 64 HTK mel filters -> dB; architectures.py:830-837, 869-878,
// deployed as the conv1d DFT of _export/onnx.py:27-83; the /32768 of nanointerpreter.py:750 is in the window table).
//
// One WARP per packed FFT (two real frames in one complex FFT-400), in the style of nww_fe3.cuh, with the FFT
// factored 20 x 20 so that there is ONE shared-memory exchange per transform:
//   pass 1  lane j (0..19): 20 windowed samples n = j + 20 m of both frames -> radix-20 DFT in registers (a 4 x 5
//           prime-factor DFT: no internal twiddles) -> x W400^(jq) -> row q of the warp's private 20 x 21 buffer;
//   pass 2  lane q: row q -> radix-20 DFT -> register p holds Z[q + 20 p];
//   power   Z[400 - k] for k = q + 20 p sits in lane 20 - q, register 19 - p: ten register shuffles give every lane
//           the partners of its bins p = 0..9 (lane 0 pairs inside itself and owns bin 200), so the two frames'
//           power spectra come straight out of registers;
//   mel     FP32 sparse filters from the power rows (they overlay the work buffer), a lane owns a filter for both
//           frames, widest filters first; 10 log10 with the reference's 1e-10 floor.
// The generic batch-interleaved radix 5 x 5 x 4 x 4 code of nww_frontend.cuh (four CTA-wide barriers per batch,
// every pass through shared memory) took 1.15 ms per 4096 windows; see DESIGN.md §4.2 for this one.
//
// The warps of a CTA are fully decoupled: packed FFT number G = 51 * (window iteration) + f goes to warp G mod 16,
// across window boundaries, so the 51 FFTs of a window (not a multiple of 16) leave no warp idle.  PCM windows are
// staged by TMA bulk copies into two slots; a slot is refilled by whichever warp is the last to leave its window.
#pragma once

#include "nww_stage.cuh"
#include "nww_fe2.cuh"

namespace nww {

struct Fe5 {
    using G = GeoREF64x101;
    static constexpr int NT = 384, NWARP = 12;                  // 168 registers per thread: the radix-20 butterflies do not spill
    static constexpr int N_PACKED = (G::N_FRAMES + 1) / 2;      // 51 packed FFTs per window
    static constexpr int P = 21;                                // pitch (complex) of a pass-1 output row: odd, so that
                                                                // lanes reading different rows hit different 16-byte banks
    static constexpr int WB = 20 * P;                           // 420 complex per warp
    static constexpr int PW_PITCH = 208;                        // 201 power bins + the over-read of the padded mel rows
    static constexpr int MEL_ROW = 20;                          // padded filter row: (first bin & 3) + count, rounded up to 4
    static constexpr int SLOT = G::CLIP + 8;                    // int16 per PCM slot (16-byte skew, see PcmStager)
    static constexpr size_t kWorkBytes = (size_t)NWARP * WB * sizeof(cplx<double>);                  // 107520
    static constexpr size_t kWinBytes = (size_t)G::WIN * sizeof(double);                             // 3200
    static constexpr size_t kTwBytes = (size_t)400 * sizeof(cplx<double>);                           // 6400
    static constexpr size_t kMelBytes = (size_t)G::N_MELS * MEL_ROW * sizeof(float) + (size_t)G::N_MELS * sizeof(int2);   // 5632
    static constexpr size_t kPcmBytes = align_up((size_t)2 * SLOT * sizeof(int16_t), 128);          // 64128
    static constexpr size_t kOffWin = kWorkBytes;
    static constexpr size_t kOffTw = kOffWin + kWinBytes;
    static constexpr size_t kOffMel = kOffTw + kTwBytes;
    static constexpr size_t kOffPcm = align_up(kOffMel + kMelBytes, 128);
    static constexpr size_t kOffBars = kOffPcm + kPcmBytes;
    static constexpr size_t kTotal = kOffBars + 128;
    static_assert(2 * PW_PITCH * sizeof(float) <= WB * sizeof(cplx<double>), "the power rows overlay the work buffer");
};

// Host-side check (engine create): every filter fits the padded row.
static inline bool fe5_tables_fit(const int* mel_start, const int* mel_count, int n_mels) {
    if (n_mels != GeoREF64x101::N_MELS) return false;
    for (int m = 0; m < n_mels; ++m) {
        if (mel_count[m] && (mel_start[m] & 3) + mel_count[m] > Fe5::MEL_ROW) return false;
        if (mel_start[m] + mel_count[m] > GeoREF64x101::N_FREQS) return false;
    }
    return true;
}

// 20-point DFT in registers, natural order in and out: the 4 x 5 prime-factor algorithm (input map
// n = (5 n1 + 4 n2) mod 20, output map k = (5 k1 + 16 k2) mod 20 — tools/probe/fft400_model.py checks the maps).
// load(n) hands out input n when the first stage needs it and store(k, X[k]) takes the outputs as the second stage
// makes them, so that a caller which reads from / writes to memory never holds all twenty values besides t[][].
template <typename LoadFn, typename StoreFn> __device__ __forceinline__ void fe5_dft20(LoadFn load, StoreFn store) {
    cplx<double> t[4][5];
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) {
        cplx<double> x[4];
#pragma unroll
        for (int n1 = 0; n1 < 4; ++n1) x[n1] = load((5 * n1 + 4 * n2) % 20);
        SmallDft<double, 4>::run(x);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) t[k1][n2] = x[k1];
    }
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        SmallDft<double, 5>::run(t[k1]);
#pragma unroll
        for (int k2 = 0; k2 < 5; ++k2) store((5 * k1 + 16 * k2) % 20, t[k1][k2]);
    }
}

__device__ __forceinline__ void fe5_build_tables(unsigned char* smem, const FrontendTables<double>& tab, int tid, int nthreads) {
    using G = GeoREF64x101;
    double* win_s = reinterpret_cast<double*>(smem + Fe5::kOffWin);
    cplx<double>* tw_s = reinterpret_cast<cplx<double>*>(smem + Fe5::kOffTw);
    float* wpad = reinterpret_cast<float*>(smem + Fe5::kOffMel);
    int2* meta = reinterpret_cast<int2*>(wpad + G::N_MELS * Fe5::MEL_ROW);
    // Hann * 2^-15 * 1/2: the half makes Z[k] +- conj(Z[N-k]) the spectra of the two packed frames without a scale
    for (int i = tid; i < G::WIN; i += nthreads) win_s[i] = 0.5 * tab.window[i];
    for (int i = tid; i < 400; i += nthreads) tw_s[i] = tab.twiddle[((i / 20) * (i % 20)) % 400];      // [q][j] = W400^(jq)
    for (int i = tid; i < G::N_MELS * Fe5::MEL_ROW; i += nthreads) {
        const int m = i / Fe5::MEL_ROW, j = i - m * Fe5::MEL_ROW;
        const int ks = tab.mel_start[m], cnt = tab.mel_count[m];
        const int bin = (ks & ~3) + j;
        wpad[i] = (bin >= ks && bin < ks + cnt) ? tab.mel_w[tab.mel_woff[m] + bin - ks] : 0.0f;
    }
    for (int m = tid; m < G::N_MELS; m += nthreads) {
        const int ks = tab.mel_start[m], cnt = tab.mel_count[m];
        meta[m] = make_int2(ks & ~3, cnt ? ((ks & 3) + cnt + 3) >> 2 : 0);
    }
}

// sum_k w_m[k] P[k] for both frames of the packed FFT: four partial sums over the bins k mod 4 (as fe2_mel_dot2)
__device__ __forceinline__ void fe5_mel_dot2(const float* __restrict__ prow_a, const float* __restrict__ prow_b, int m,
                                             const float* __restrict__ wpad, float* __restrict__ out_a, float* __restrict__ out_b) {
    const int2 mt = reinterpret_cast<const int2*>(wpad + GeoREF64x101::N_MELS * Fe5::MEL_ROW)[m];
    const int k0 = mt.x, n4 = mt.y;
    const float4* __restrict__ pa4 = reinterpret_cast<const float4*>(prow_a + k0);
    const float4* __restrict__ pb4 = reinterpret_cast<const float4*>(prow_b + k0);
    const float4* __restrict__ w4 = reinterpret_cast<const float4*>(wpad + m * Fe5::MEL_ROW);
    constexpr int MAXG = Fe5::MEL_ROW / 4;                      // five groups of four bins
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f, b0 = 0.0f, b1 = 0.0f, b2 = 0.0f, b3 = 0.0f;
    float4 pa[MAXG], pb[MAXG], w[MAXG];
#pragma unroll
    for (int j = 0; j < MAXG; ++j)
        if (j < n4) {
            w[j] = w4[j];
            pa[j] = pa4[j];
            pb[j] = pb4[j];
        }
#pragma unroll
    for (int j = 0; j < MAXG; ++j)
        if (j < n4) {
            a0 = fmaf(w[j].x, pa[j].x, a0); a1 = fmaf(w[j].y, pa[j].y, a1);
            a2 = fmaf(w[j].z, pa[j].z, a2); a3 = fmaf(w[j].w, pa[j].w, a3);
            b0 = fmaf(w[j].x, pb[j].x, b0); b1 = fmaf(w[j].y, pb[j].y, b1);
            b2 = fmaf(w[j].z, pb[j].z, b2); b3 = fmaf(w[j].w, pb[j].w, b3);
        }
    *out_a = (a0 + a1) + (a2 + a3);
    *out_b = (b0 + b1) + (b2 + b3);
}

// Clip index of padded-signal sample i (reflect padding of 200 on both sides, torch.stft center=True)
__device__ __forceinline__ int fe5_reflect(int i) {
    i -= GeoREF64x101::PAD;
    if (i < 0) i = -i;
    if (i >= GeoREF64x101::CLIP) i = 2 * (GeoREF64x101::CLIP - 1) - i;
    return i;
}

// One packed FFT f (frames 2f, 2f + 1) of the window at x (16000 int16, shared) by one warp.
// store(frame, mel bin, dB) is called for the frames that exist.
template <typename StoreFn>
__device__ __forceinline__ void fe5_warp_fft(const int16_t* __restrict__ x, int f, cplx<double>* __restrict__ wb,
                                             const unsigned char* __restrict__ smem, const FrontendTables<double>& tab,
                                             StoreFn store, int lane) {
    using G = GeoREF64x101;
    const double* win_s = reinterpret_cast<const double*>(smem + Fe5::kOffWin);
    const cplx<double>* tw_s = reinterpret_cast<const cplx<double>*>(smem + Fe5::kOffTw);
    const float* wpad = reinterpret_cast<const float*>(smem + Fe5::kOffMel);
    const bool active = lane < 20;
    const int j = active ? lane : lane - 20;                    // lanes 20..31 shadow lanes 0..11 (no stores)
    const int fa = 2 * f;
    const bool has_b = fa + 1 < G::N_FRAMES;
    // ---- pass 1 ------------------------------------------------------------------------------------------------
    auto put = [&](int q, cplx<double> y) {
        if (active) wb[q * Fe5::P + j] = (q == 0) ? y : cmul(y, tw_s[q * 20 + j]);
    };
    if (f >= 1 && fa + 1 <= 98) {                               // both frames inside the clip: no reflection
        const int16_t* xa = x + fa * G::HOP - G::PAD + j;
        fe5_dft20(
            [&](int m) {
                const double w = win_s[j + 20 * m];
                return cplx<double>{w * fe2_i16_to_f64(xa[20 * m]), w * fe2_i16_to_f64(xa[20 * m + G::HOP])};
            },
            put);
    } else {
        fe5_dft20(
            [&](int m) {
                const int n = j + 20 * m;
                const double w = win_s[n];
                const double sa = fe2_i16_to_f64(x[fe5_reflect(fa * G::HOP + n)]);
                const double sb = has_b ? fe2_i16_to_f64(x[fe5_reflect((fa + 1) * G::HOP + n)]) : 0.0;
                return cplx<double>{w * sa, w * sb};
            },
            put);
    }
    __syncwarp();
    // ---- pass 2: lane q, register p = Z[q + 20 p] -----------------------------------------------------------------
    cplx<double> v[20];
    {
        const cplx<double>* row = wb + j * Fe5::P;
        fe5_dft20([&](int m) { return row[m]; }, [&](int p, cplx<double> z) { v[p] = z; });
    }
    // ---- power of both frames: bins k = q + 20 p, p = 0..9; the partner Z[400 - k] is register 19 - p of lane 20 - q ---
    const int q = j;
    const int partner = (20 - q) % 20;
    float pa_v[10], pb_v[10];
#pragma unroll
    for (int p = 0; p < 10; ++p) {
        cplx<double> W;
        W.x = __shfl_sync(0xffffffffu, v[19 - p].x, partner);
        W.y = __shfl_sync(0xffffffffu, v[19 - p].y, partner);
        if (q == 0) W = v[(20 - p) % 20];                       // k = 20 p pairs with 20 (20 - p): the same lane
        const cplx<double> U = v[p];
        const double ar = U.x + W.x, ai = U.y - W.y;
        const double br = U.y + W.y, bi = U.x - W.x;
        pa_v[p] = (float)(ar * ar + ai * ai);
        pb_v[p] = (float)(br * br + bi * bi);
    }
    const double nr = 2.0 * v[10].x, ni = 2.0 * v[10].y;        // bin 200 (lane 0 only)
    __syncwarp();                                               // every lane has read its pass-2 row
    float* pwa = reinterpret_cast<float*>(wb);                  // the power rows overlay the work buffer
    float* pwb = pwa + Fe5::PW_PITCH;
    if (active) {
#pragma unroll
        for (int p = 0; p < 10; ++p) {
            pwa[q + 20 * p] = pa_v[p];
            pwb[q + 20 * p] = pb_v[p];
        }
        if (lane == 0) {
            pwa[200] = (float)(nr * nr);
            pwb[200] = (float)(ni * ni);
        } else if (lane < 8) {                                  // the padded mel rows read a few bins past Nyquist
            pwa[200 + lane] = 0.0f;
            pwb[200 + lane] = 0.0f;
        }
    }
    __syncwarp();
    // ---- mel + dB: a lane owns one filter for both frames, the 32 widest filters first --------------------------------
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int m = G::N_MELS - 1 - (lane + 32 * r);
        float pm_a, pm_b;
        fe5_mel_dot2(pwa, pwb, m, wpad, &pm_a, &pm_b);
        store(fa, m, (pm_a <= tab.amin) ? tab.floor_db : 10.0f * log10f(pm_a));
        if (has_b) store(fa + 1, m, (pm_b <= tab.amin) ? tab.floor_db : 10.0f * log10f(pm_b));
    }
    __syncwarp();                                               // the buffer is free for the warp's next FFT
}

// ----------------------------------------------------------------------------------------
// Front end only (REF64x101): log-mel to global memory, (F, T) or (T, F) per window.  Persistent, one CTA per SM.
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(Fe5::NT, 1)
frontend5_kernel(WindowSource src, long long n_windows, FrontendTables<double> tab, float* __restrict__ mel_out, int time_major) {
    using G = GeoREF64x101;
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int16_t* pcm_s = reinterpret_cast<int16_t*>(smem + Fe5::kOffPcm);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Fe5::kOffBars);          // [2] a slot's window has landed
    int* left = reinterpret_cast<int*>(smem + Fe5::kOffBars + 16);               // [2] warps that have left the slot's window
    cplx<double>* wb = reinterpret_cast<cplx<double>*>(smem) + (size_t)warp * Fe5::WB;

    const long long w0 = blockIdx.x;
    const int n_it = (w0 < n_windows) ? (int)((n_windows - w0 + gridDim.x - 1) / gridDim.x) : 0;
    auto issue = [&](int it) {                                                   // one thread: window `it` of this CTA -> slot it & 1
        const int16_t* s = src.at(w0 + (long long)it * gridDim.x);
        const int skew = PcmStager<G::CLIP>::skew_of(s);
        const uint32_t bytes = (uint32_t)G::CLIP * (uint32_t)sizeof(int16_t) + (skew ? 16u : 0u);
        fence_proxy_async();
        mbar_expect_tx(&full[it & 1], bytes);
        bulk_g2s(pcm_s + (size_t)(it & 1) * Fe5::SLOT, s - skew, bytes, &full[it & 1]);
    };
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
        left[0] = left[1] = 0;
    }
    fe5_build_tables(smem, tab, tid, Fe5::NT);
    __syncthreads();
    if (tid == 0) {
        if (n_it > 0) issue(0);
        if (n_it > 1) issue(1);
    }

    const int stride_m = time_major ? 1 : G::N_FRAMES;
    const int stride_t = time_major ? G::N_MELS : 1;
    const long long total = (long long)n_it * Fe5::N_PACKED;
    int cur = -1;
    const int16_t* x = nullptr;
    float* mw = nullptr;
#pragma unroll 1
    for (long long g = warp; g < total; g += Fe5::NWARP) {
        const int it = (int)(g / Fe5::N_PACKED);
        const int f = (int)(g - (long long)it * Fe5::N_PACKED);
        if (it != cur) {
            if (cur >= 0) {
                // this warp is done with window `cur`: the last warp to say so refills the slot with window cur + 2
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    if (atomicAdd(&left[cur & 1], 1) == Fe5::NWARP - 1) {
                        atomicExch(&left[cur & 1], 0);
                        if (cur + 2 < n_it) issue(cur + 2);
                    }
                }
            }
            mbar_wait(&full[it & 1], (uint32_t)((it >> 1) & 1));
            cur = it;
            const long long w = w0 + (long long)it * gridDim.x;
            x = pcm_s + (size_t)(it & 1) * Fe5::SLOT + PcmStager<G::CLIP>::skew_of(src.at(w));
            mw = mel_out + w * (long long)(G::N_MELS * G::N_FRAMES);
        }
        fe5_warp_fft(x, f, wb, smem, tab, [&](int fr, int m, float db) { mw[m * stride_m + fr * stride_t] = db; }, lane);
    }
}

}  // namespace nww
