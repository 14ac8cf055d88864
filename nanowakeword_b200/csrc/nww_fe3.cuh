// nww_fe3.cuh — front end v3 for the NS40x98 geometry: one WARP per packed FFT.
//
// Same arithmetic as nww_fe2.cuh (FP64 radix 8 x 8 x 8 DIF, two real frames per complex FFT, the
// (k, N-k) pairs of pass 3 inside one thread, FP32 sparse mel + 10 log10) and the same reference
// contract (architectures.py:830-837, 869-878; _export/onnx.py:27-83; nanointerpreter.py:750), but
// with no barrier between the passes other than __syncwarp: every warp owns a private 520-element
// FP64 work buffer and carries whole FFTs from PCM to dB on its own.
//
// Why: the v2 profile (profiles/r01_v21_*) showed 33 % of all warp time parked at group / CTA barriers
// and the FP64 pipe 16 % busy — four warps per scheduler are too few to hide barrier skew.  Here the
// sixteen warps of the CTA are fully decoupled during the FFT phase: warp w transforms FFTs
// w, w + 16, w + 32 (frames 0..95); the one FFT left over (frames 96, 97) is done by warps 0..3
// together with the group code of nww_fe2.cuh.
//
// Per FFT a lane runs 2 radix-8 butterflies in pass 1 (j = lane, lane + 32), 2 in pass 2 and the
// pair butterfly of pass 3; the power of both frames (2 x 257 floats) then overlays the warp's own
// work buffer and the 80 (filter, frame) sums are dealt to the lanes longest-filter-first.
#pragma once

#include "nww_fe2.cuh"

namespace nww {

struct Fe3 {
    static constexpr int NT = 512, NWARP = 16;
    static constexpr int NPAD = Fe2::NPAD;
    static constexpr size_t kWorkBytes = (size_t)NWARP * NPAD * sizeof(cplx<double>);      // 133120
    static constexpr size_t kWinBytes = (GeoNS40x98::WIN * sizeof(double) + 127) / 128 * 128;   // 3200
    static constexpr size_t kTwBytes = (Fe2::kTwBytes + 127) / 128 * 128;                  // 8064
    static constexpr int N_PRIVATE = 48;                    // FFTs done warp-privately (3 per warp)
    static_assert(kWorkBytes >= Fe2::kScratchBytes, "the shared last batch reuses the private buffers");
};

// Hann * 2^-15 * 1/2 in shared memory (see nww_fe2.cuh for the two factors)
__device__ __forceinline__ void fe3_build_window(double* win_s, const double* __restrict__ window, int tid, int nthreads) {
    for (int i = tid; i < GeoNS40x98::WIN; i += nthreads) win_s[i] = 0.5 * window[i];
}

// One packed FFT (frames at x and x + 160) by one warp; store(frame 0|1, mel bin, dB).
template <typename StoreFn>
__device__ __forceinline__ void fe3_warp_fft(const int16_t* __restrict__ x, cplx<double>* __restrict__ wb,
                                             const double* __restrict__ win_s, const cplx<double>* __restrict__ tw_smem,
                                             const FrontendTables<double>& tab, StoreFn store, int lane) {
    const cplx<double>* tw1 = tw_smem;
    const cplx<double>* tw2 = tw_smem + 7 * 64;
    // ---- pass 1: butterflies j = lane, lane + 32 (unrolled: the two independent butterflies interleave) -------
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int jj = lane + 32 * h;
        const int16_t* xa = x + jj;
        cplx<double> v[8];
#pragma unroll
        for (int m = 0; m < 7; ++m) {
            const int n = jj + 64 * m;
            const double w = (n < GeoNS40x98::WIN) ? win_s[n] : 0.0;
            const double sa = fe2_i16_to_f64(xa[64 * m]);
            const double sb = fe2_i16_to_f64(xa[64 * m + GeoNS40x98::HOP]);
            v[m] = {w * sa, w * sb};
        }
        v[7] = {0.0, 0.0};
        SmallDft<double, 8>::run(v);
        wb[jj] = v[0];
#pragma unroll
        for (int q = 1; q < 8; ++q) wb[jj + 65 * q] = cmul(v[q], tw1[(q - 1) * 64 + jj]);
    }
    __syncwarp();
    // ---- pass 2 ----------------------------------------------------------------------------------------------
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int jj = lane + 32 * h;
        const int b2 = jj >> 3, j2 = jj & 7;
        cplx<double>* blk = wb + 65 * b2 + j2;
        cplx<double> v[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) v[m] = blk[8 * m];
        SmallDft<double, 8>::run(v);
        blk[0] = v[0];
#pragma unroll
        for (int q = 1; q < 8; ++q) blk[8 * q] = cmul(v[q], tw2[(q - 1) * 8 + j2]);
    }
    __syncwarp();
    // ---- pass 3 + power of both frames (see nww_fe2.cuh for the pairing) ---------------------------------------
    float pa_v[9], pb_v[9];
    int bins[9];
    {
        const int c = lane;
        const int cb = (c == 0) ? 32 : 64 - c;
        const cplx<double>* pa = wb + 65 * (c & 7) + 8 * (c >> 3);
        const cplx<double>* pb = wb + 65 * (cb & 7) + 8 * (cb >> 3);
        cplx<double> A[8], B[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) A[m] = pa[m];
#pragma unroll
        for (int m = 0; m < 8; ++m) B[m] = pb[m];
        SmallDft<double, 8>::run(A);
        SmallDft<double, 8>::run(B);
        cplx<double> U[8], W[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            U[i] = A[i];
            W[i] = B[7 - i];
        }
        const bool special = (c == 0);
        if (special) {
            W[0] = A[0]; W[1] = A[7]; W[2] = A[6]; W[3] = A[5];
            U[4] = B[0]; U[5] = B[1]; U[6] = B[2]; U[7] = B[3];
            W[4] = B[7]; W[5] = B[6]; W[6] = B[5]; W[7] = B[4];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double ar = U[i].x + W[i].x, ai = U[i].y - W[i].y;
            const double br = U[i].y + W[i].y, bi = U[i].x - W[i].x;
            bins[i] = (i < 4) ? (64 * i + c) : (special ? (64 * (i - 4) + 32) : (64 * (7 - i) + 64 - c));
            pa_v[i] = (float)(ar * ar + ai * ai);
            pb_v[i] = (float)(br * br + bi * bi);
        }
        const double ar = 2.0 * A[4].x, br = 2.0 * A[4].y;      // bin 256 (lane 0 only)
        bins[8] = 256;
        pa_v[8] = (float)(ar * ar);
        pb_v[8] = (float)(br * br);
    }
    __syncwarp();                                                // every lane has read its pass-3 inputs
    float* pwa = reinterpret_cast<float*>(wb);                   // the power tables overlay the work buffer
    float* pwb = pwa + Fe2::PW_PITCH;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        pwa[bins[i]] = pa_v[i];
        pwb[bins[i]] = pb_v[i];
    }
    if (lane == 0) {
        pwa[256] = pa_v[8];
        pwb[256] = pb_v[8];
    } else if (lane < 8) {                                       // the padded mel rows read a few bins past Nyquist
        pwa[256 + lane] = 0.0f;
        pwb[256 + lane] = 0.0f;
    }
    __syncwarp();
    // ---- mel + dB: a lane owns one filter for BOTH frames (the weights are read once), the 32 widest filters first ---
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int slot = lane + 32 * r;
        if (slot < GeoNS40x98::N_MELS) {
            const int m = GeoNS40x98::N_MELS - 1 - slot;
            float pm_a, pm_b;
            fe2_mel_dot2(pwa, pwb, m, tw_smem, tab, &pm_a, &pm_b);
            store(0, m, (pm_a <= tab.amin) ? tab.floor_db : 10.0f * log10f(pm_a));
            store(1, m, (pm_b <= tab.amin) ? tab.floor_db : 10.0f * log10f(pm_b));
        }
    }
    __syncwarp();                                                // the buffer is free for the warp's next FFT
}

// One window: pcm (16000 int16, shared) -> mel[m * stride_m + t * stride_t].  All 512 threads; ends with
// __syncthreads().  scratch: Fe3::kWorkBytes.
__device__ __forceinline__ void fe3_logmel_window(const int16_t* __restrict__ pcm, unsigned char* __restrict__ scratch,
                                                  const double* __restrict__ win_s, const cplx<double>* __restrict__ tw_smem,
                                                  const FrontendTables<double>& tab, float* __restrict__ mel, int stride_m,
                                                  int stride_t, int tid, int f_lo = 0 /* first packed FFT = frame_lo / 2 */) {
    const int warp = tid >> 5, lane = tid & 31;
    cplx<double>* wb = reinterpret_cast<cplx<double>*>(scratch) + (size_t)warp * Fe3::NPAD;
#pragma unroll 1
    for (int f = f_lo + warp; f < Fe3::N_PRIVATE; f += Fe3::NWARP) {
        float* mf = mel + 2 * f * stride_t;
        fe3_warp_fft(pcm + 2 * f * GeoNS40x98::HOP, wb, win_s, tw_smem, tab,
                     [&](int fr, int m, float db) { mf[m * stride_m + fr * stride_t] = db; }, lane);
    }
    __syncthreads();
    // frames 96, 97: the one FFT left over, by group 0 (warps 0..3) of the v2 code
    fe2_run(
        1, [&](int) { return Fe2Batch{pcm + 2 * Fe3::N_PRIVATE * GeoNS40x98::HOP, 2}; },
        [&](int, int fr, int m, float db) { mel[m * stride_m + (2 * Fe3::N_PRIVATE + fr) * stride_t] = db; }, scratch, tw_smem,
        tab, tid);
}

// ----------------------------------------------------------------------------------------
// Front end only (NS40x98): log-mel to global memory, (F, T) or (T, F) per window.
// ----------------------------------------------------------------------------------------
struct Fe3KernelSmem {
    static constexpr size_t kTotal = Fe3::kWorkBytes + Fe3::kWinBytes + Fe3::kTwBytes + PcmStager<GeoNS40x98::CLIP>::kBytes;
};

__global__ void __launch_bounds__(Fe3::NT, 1)
frontend3_kernel(WindowSource src, long long n_windows, FrontendTables<double> tab, float* __restrict__ mel_out,
                 int time_major, int frame_lo /* even: frames below it are neither staged nor computed */) {
    using G = GeoNS40x98;
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x;
    double* win_s = reinterpret_cast<double*>(smem + Fe3::kWorkBytes);
    cplx<double>* tw = reinterpret_cast<cplx<double>*>(smem + Fe3::kWorkBytes + Fe3::kWinBytes);
    PcmStager<G::CLIP> stager;
    stager.carve(smem + Fe3::kWorkBytes + Fe3::kWinBytes + Fe3::kTwBytes);
    stager.init(tid);
    fe2_build_tables(tw, tab, tid, Fe3::NT);
    fe3_build_window(win_s, tab.window, tid, Fe3::NT);
    __syncthreads();

    const int stride_m = time_major ? 1 : G::N_FRAMES;
    const int stride_t = time_major ? G::N_MELS : 1;
    const int first_sample = frame_lo * G::HOP;
    long long w = blockIdx.x;
    if (w < n_windows) stager.issue(0, src.at(w), tid, first_sample);
    for (int it = 0; w < n_windows; w += gridDim.x, ++it) {
        const long long wn = w + gridDim.x;
        if (wn < n_windows) stager.issue((it + 1) & 1, src.at(wn), tid, first_sample);
        const int16_t* x = stager.wait(it & 1, (it >> 1) & 1, src.at(w));
        fe3_logmel_window(x, smem, win_s, tw, tab, mel_out + w * (long long)(G::N_MELS * G::N_FRAMES), stride_m, stride_t, tid,
                          frame_lo >> 1);
    }
}

}  // namespace nww
