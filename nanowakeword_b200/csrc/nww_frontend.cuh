// nww_frontend.cuh — K1: int16 PCM window -> log-mel (dB), computed by one CTA in shared memory.
//
// Replaces the reference's MelSpectrogram + AmplitudeToDB
// (reference nanowakeword/modules/architectures.py:830-837, 869-878; deployed as the
// conv1d-DFT of nanowakeword/_export/onnx.py:27-83) and the int16 -> float scaling of
// nanowakeword/interpreter/nanointerpreter.py:750.
//
// Algorithm (B200-first, not the reference's dense 400x402 DFT matmul):
//   * two real frames are packed into one complex N-point FFT (z = frameA + i*frameB);
//   * the CTA transforms NFB such FFTs at once with an in-place decimation-in-frequency
//     mixed-radix FFT held in shared memory in a batch-interleaved layout
//     work[point * NFB + fft]  — lanes walk the fft index, so every pass of every radix is
//     bank-conflict free; radices 8*8*8 (N=512) or 5*5*4*4 (N=400);
//   * the first pass reads PCM directly (window * sample, reflect padding when centred),
//     so framed audio is never materialised;
//   * |X_A|^2, |X_B|^2 are recovered from Z[k], Z[N-k] on the fly while accumulating the
//     sparse triangular mel filters; 10*log10(max(., amin)) closes the stage.
// T = float is the fast mode; T = double (B200 keeps half-rate FP64) makes the power
// spectrum exact enough that bins 90 dB below the frame peak still meet 1e-4 dB.
#pragma once

#include "nww_common.cuh"

namespace nww {

struct GeoNS40x98 {
    static constexpr int N_FFT = 512, WIN = 400, HOP = 160, N_MELS = 40, N_FRAMES = 98, N_FREQS = 257;
    static constexpr int CENTER = 0, PAD = 0, CLIP = 16000;
    static constexpr int N_PASS = 3;
    static constexpr int R0 = 8, R1 = 8, R2 = 8, R3 = 1;
};
struct GeoREF64x101 {
    static constexpr int N_FFT = 400, WIN = 400, HOP = 160, N_MELS = 64, N_FRAMES = 101, N_FREQS = 201;
    static constexpr int CENTER = 1, PAD = 200, CLIP = 16000;
    static constexpr int N_PASS = 4;
    static constexpr int R0 = 5, R1 = 5, R2 = 4, R3 = 4;
};

// Device-resident constant tables of one front end (built by the engine at create time).
template <typename T> struct FrontendTables {
    const T* window;            // [WIN]   Hann * 2^-15 (the /32768 of nanointerpreter.py:750, exact)
    const T* window_unscaled;   // [WIN]   for float PCM that is already in [-1, 1)
    const cplx<T>* twiddle;     // [N_FFT] exp(-2*pi*i*k/N)
    const uint16_t* binpos;     // [N_FFT] where bin k sits after the in-place DIF passes
    const int* mel_start;       // [N_MELS] first bin with a non-zero weight
    const int* mel_count;       // [N_MELS] number of consecutive bins
    const int* mel_woff;        // [N_MELS] offset of this filter's weights in mel_w
    const float* mel_w;         // packed filter weights (float32 values of the reference's fb)
    float amin;                 // 1e-10
    float floor_db;             // 10*log10(amin)
    int mel_vec_ok;             // every filter fits the padded row of the vectorised mel (nww_fe2.cuh: 32 bins; nww_fe5.cuh: 20)
};

// --------------------------------------------------------------------------- small DFTs
template <typename T, int R> struct SmallDft;

template <typename T> struct SmallDft<T, 2> {
    __device__ static __forceinline__ void run(cplx<T>* v) {
        cplx<T> a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};
template <typename T> struct SmallDft<T, 4> {
    __device__ static __forceinline__ void run(cplx<T>* v) {
        cplx<T> s0 = cadd(v[0], v[2]), s1 = csub(v[0], v[2]);
        cplx<T> s2 = cadd(v[1], v[3]), s3 = csub(v[1], v[3]);
        v[0] = cadd(s0, s2);
        v[2] = csub(s0, s2);
        v[1] = cadd(s1, mul_mi(s3));
        v[3] = cadd(s1, mul_pi(s3));
    }
};
template <typename T> struct SmallDft<T, 5> {
    __device__ static __forceinline__ void run(cplx<T>* v) {
        const T c1 = (T)0.30901699437494742410, c2 = (T)-0.80901699437494742410;
        const T s1 = (T)0.95105651629515357212, s2 = (T)0.58778525229247312917;
        cplx<T> t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]);
        cplx<T> t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
        cplx<T> a1 = {v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y};
        cplx<T> a2 = {v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y};
        cplx<T> b1 = {s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y};
        cplx<T> b2 = {s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y};
        v[0] = {v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y};
        v[1] = cadd(a1, mul_mi(b1));
        v[4] = cadd(a1, mul_pi(b1));
        v[2] = cadd(a2, mul_mi(b2));
        v[3] = cadd(a2, mul_pi(b2));
    }
};
template <typename T> struct SmallDft<T, 8> {
    __device__ static __forceinline__ void run(cplx<T>* v) {
        const T h = (T)0.70710678118654752440;
        cplx<T> a[4], b[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            a[m] = cadd(v[m], v[m + 4]);
            b[m] = csub(v[m], v[m + 4]);
        }
        // b[m] *= w8^m : w8 = (1 - i)/sqrt2, w8^2 = -i, w8^3 = (-1 - i)/sqrt2
        b[1] = {h * (b[1].x + b[1].y), h * (b[1].y - b[1].x)};
        b[2] = mul_mi(b[2]);
        b[3] = {h * (b[3].y - b[3].x), -h * (b[3].x + b[3].y)};
        SmallDft<T, 4>::run(a);
        SmallDft<T, 4>::run(b);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[2 * q] = a[q];
            v[2 * q + 1] = b[q];
        }
    }
};

// --------------------------------------------------------------------------- PCM access
template <typename G> __device__ __forceinline__ int pcm_index(int i) {
    // i indexes the (optionally reflect-padded) signal; returns the index into the clip.
    if (G::CENTER) {
        i -= G::PAD;
        if (i < 0) i = -i;
        if (i >= G::CLIP) i = 2 * (G::CLIP - 1) - i;
    }
    return i;
}

// One in-place DIF pass of radix R on sub-blocks of length L for all NFB interleaved FFTs.
template <typename T, typename G, int L, int R, bool FROM_PCM, int NFB, typename PcmT>
__device__ __forceinline__ void fft_pass(cplx<T>* __restrict__ work, int nf, int f0, const PcmT* __restrict__ pcm,
                                         const T* __restrict__ window, const cplx<T>* __restrict__ tw, int tid,
                                         int nthreads) {
    constexpr int N = G::N_FFT;
    constexpr int S = L / R;
    constexpr int NTASK = (N / R) * NFB;
    for (int t = tid; t < NTASK; t += nthreads) {
        const int f = t % NFB;
        if (f >= nf) continue;
        const int jj = t / NFB;
        const int b = jj / S;
        const int j = jj - b * S;
        const int base = b * L + j;
        cplx<T> v[R];
        if (FROM_PCM) {
            const int fa = 2 * (f0 + f);               // frame A; frame B = fa + 1
            const bool has_b = (fa + 1) < G::N_FRAMES;
#pragma unroll
            for (int m = 0; m < R; ++m) {
                const int n = base + m * S;
                T xa = (T)0, xb = (T)0;
                if (n < G::WIN) {
                    const T w = window[n];
                    xa = w * (T)pcm[pcm_index<G>(fa * G::HOP + n)];
                    if (has_b) xb = w * (T)pcm[pcm_index<G>((fa + 1) * G::HOP + n)];
                }
                v[m] = {xa, xb};
            }
        } else {
#pragma unroll
            for (int m = 0; m < R; ++m) v[m] = work[(base + m * S) * NFB + f];
        }
        SmallDft<T, R>::run(v);
        work[base * NFB + f] = v[0];
#pragma unroll
        for (int q = 1; q < R; ++q) {
            cplx<T> y = v[q];
            if (S > 1) y = cmul(y, tw[j * q * (N / L)]);
            work[(base + q * S) * NFB + f] = y;
        }
    }
}

// All passes for one batch of FFTs, with the CTA-wide barriers between them.
template <typename T, typename G, int NFB, typename PcmT>
__device__ __forceinline__ void fft_batch(cplx<T>* work, int nf, int f0, const PcmT* pcm, const T* window,
                                          const cplx<T>* tw, int tid, int nthreads) {
    constexpr int N = G::N_FFT;
    fft_pass<T, G, N, G::R0, true, NFB, PcmT>(work, nf, f0, pcm, window, tw, tid, nthreads);
    __syncthreads();
    fft_pass<T, G, N / G::R0, G::R1, false, NFB, PcmT>(work, nf, f0, pcm, window, tw, tid, nthreads);
    __syncthreads();
    fft_pass<T, G, N / (G::R0 * G::R1), G::R2, false, NFB, PcmT>(work, nf, f0, pcm, window, tw, tid, nthreads);
    __syncthreads();
    if constexpr (G::N_PASS > 3) {
        fft_pass<T, G, N / (G::R0 * G::R1 * G::R2), (G::R3 > 1 ? G::R3 : 2), false, NFB, PcmT>(work, nf, f0, pcm,
                                                                                              window, tw, tid, nthreads);
        __syncthreads();
    }
}

// Power spectrum of both packed frames -> sparse mel -> dB, written to mel[m*stride_m + t*stride_t].
template <typename T, typename G, int NFB>
__device__ __forceinline__ void mel_batch(const cplx<T>* __restrict__ work, int nf, int f0,
                                          const FrontendTables<T>& tab, float* __restrict__ mel, int stride_m,
                                          int stride_t, int tid, int nthreads) {
    constexpr int N = G::N_FFT;
    constexpr int NTASK = G::N_MELS * 2 * NFB;
    for (int t = tid; t < NTASK; t += nthreads) {
        const int f = t % NFB;
        const int r = t / NFB;
        const int which = r & 1;
        const int m = r >> 1;
        const int frame = 2 * (f0 + f) + which;
        if (f >= nf || frame >= G::N_FRAMES) continue;
        const int ks = tab.mel_start[m];
        const int cnt = tab.mel_count[m];
        const float* __restrict__ w = tab.mel_w + tab.mel_woff[m];
        const T sgn = which ? (T)-1 : (T)1;
        T acc = (T)0;
        for (int i = 0; i < cnt; ++i) {
            const int k = ks + i;
            const int kn = (k == 0) ? 0 : (N - k);
            const cplx<T> a = work[(int)tab.binpos[k] * NFB + f];
            const cplx<T> c = work[(int)tab.binpos[kn] * NFB + f];
            // X_A = (Z[k] + conj(Z[N-k]))/2 ; X_B = (Z[k] - conj(Z[N-k]))/(2i)
            const T re = a.x + sgn * c.x;
            const T im = a.y - sgn * c.y;
            acc += (T)w[i] * ((re * re + im * im) * (T)0.25);
        }
        const float p = (float)acc;
        const float db = (p <= tab.amin) ? tab.floor_db : 10.0f * log10f(p);
        mel[m * stride_m + frame * stride_t] = db;
    }
}

// Whole window: PCM (shared or global) -> log-mel (shared or global).  Ends with a barrier.
template <typename T, typename G, int NFB, typename PcmT>
__device__ __forceinline__ void logmel_window(const PcmT* __restrict__ pcm, cplx<T>* __restrict__ work,
                                              const FrontendTables<T>& tab, float* __restrict__ mel, int stride_m,
                                              int stride_t, int tid, int nthreads) {
    constexpr int NFFT_TOTAL = (G::N_FRAMES + 1) / 2;
    const T* window = (sizeof(PcmT) == 2) ? tab.window : tab.window_unscaled;
    for (int f0 = 0; f0 < NFFT_TOTAL; f0 += NFB) {
        const int nf = (NFFT_TOTAL - f0 < NFB) ? (NFFT_TOTAL - f0) : NFB;
        fft_batch<T, G, NFB, PcmT>(work, nf, f0, pcm, window, tab.twiddle, tid, nthreads);
        mel_batch<T, G, NFB>(work, nf, f0, tab, mel, stride_m, stride_t, tid, nthreads);
        __syncthreads();
    }
}

}  // namespace nww
