// nww_stage.cuh — persistent per-window "stage A" kernels: PCM staging + front end (+ the
// convolutional body of a head), one window per CTA iteration, grid = a multiple of the SM count.
#pragma once

#include "nww_frontend.cuh"

#ifndef NWW_CPUSIM
#define NWW_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#endif

namespace nww {

__host__ __device__ constexpr size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Double-buffered TMA bulk staging of int16 windows (global -> shared), one mbarrier per slot.
// A window may start at any int16 boundary (the stream rings of nww_stream.cuh hand out
// windows that begin wherever the stream's write position is): the bulk copy starts at the
// enclosing 16-byte boundary, moves 16 extra bytes, and wait() returns the skewed pointer.
template <int CLIP, int NSLOT = 2> struct PcmStager {
    int16_t* buf;      // [NSLOT][SLOT]
    uint64_t* bars;    // [NSLOT]
    static constexpr int SLOT = CLIP + 8;
    static constexpr size_t kBytes = align_up(NSLOT * SLOT * sizeof(int16_t), 128) + 128;
    __device__ __forceinline__ void carve(unsigned char* p) {
        buf = reinterpret_cast<int16_t*>(p);
        bars = reinterpret_cast<uint64_t*>(p + align_up(NSLOT * SLOT * sizeof(int16_t), 128));
    }
    __device__ __forceinline__ void init(int tid) {
        if (tid == 0) {
            for (int i = 0; i < NSLOT; ++i) mbar_init(&bars[i], 1);
            fence_mbar_init();
        }
        __syncthreads();
    }
    static __device__ __forceinline__ int skew_of(const int16_t* src) { return (int)((reinterpret_cast<uintptr_t>(src) & 15) >> 1); }
    // first_sample (a multiple of 8): stage only samples [first_sample, CLIP) of the window — heads that read
    // just the tail of the clip (the TCN's dependency cone) do not pay for the rest.  Indexing is unchanged.
    __device__ __forceinline__ void issue(int slot, const int16_t* src, int tid, int first_sample = 0) {
        if (tid == 0) {
            const int skew = skew_of(src);
            const uint32_t bytes = (uint32_t)(CLIP - first_sample) * (uint32_t)sizeof(int16_t) + (skew ? 16u : 0u);
            fence_proxy_async();
            mbar_expect_tx(&bars[slot], bytes);
            bulk_g2s(buf + (size_t)slot * SLOT + first_sample, src + first_sample - skew, bytes, &bars[slot]);
        }
    }
    __device__ __forceinline__ const int16_t* wait(int slot, uint32_t parity, const int16_t* src) {
        mbar_wait(&bars[slot], parity);
        return buf + (size_t)slot * SLOT + skew_of(src);
    }
};

// Where window w of a launch starts: densely packed windows, or (stream mode) an explicit
// element offset per window into the ring arena.
struct WindowSource {
    const int16_t* base;
    const long long* offsets;     // nullable
    int clip;
    // Float feeds (nww_run_windows_f32): densely packed float32 windows already scaled to [-1, 1) the way the
    // reference scales them (nanointerpreter.py:750); base is null then.  Only the launchers look at this field:
    // they pick the float front end (frontend_f32_kernel) / the float raw-audio loader.
    const float* fbase = nullptr;
    __device__ __forceinline__ const int16_t* at(long long w) const {
        return base + (offsets ? offsets[w] : w * (long long)clip);
    }
    __device__ __forceinline__ const float* atf(long long w) const { return fbase + w * (long long)clip; }
};

// ----------------------------------------------------------------------------------------
// Front end only: log-mel to global memory, either (F, T) or (T, F) per window.
// Used for the DNN head (whose "body" is the identity), for parity dumps and for streaming.
// ----------------------------------------------------------------------------------------
template <typename T, typename G, int NFB> struct FrontendSmem {
    static constexpr size_t kWork = align_up(sizeof(cplx<T>) * G::N_FFT * NFB, 128);
    static constexpr size_t kTotal = kWork + PcmStager<G::CLIP>::kBytes;
};

template <typename T, typename G, int NFB, int NT>
__global__ void __launch_bounds__(NT, 1)
frontend_kernel(WindowSource src, long long n_windows, FrontendTables<T> tab, float* __restrict__ mel_out,
                int time_major) {
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x;
    cplx<T>* work = reinterpret_cast<cplx<T>*>(smem);
    PcmStager<G::CLIP> stager;
    stager.carve(smem + FrontendSmem<T, G, NFB>::kWork);
    stager.init(tid);

    const int stride_m = time_major ? 1 : G::N_FRAMES;
    const int stride_t = time_major ? G::N_MELS : 1;
    long long w = blockIdx.x;
    if (w < n_windows) stager.issue(0, src.at(w), tid);
    for (int it = 0; w < n_windows; w += gridDim.x, ++it) {
        const long long wn = w + gridDim.x;
        if (wn < n_windows) stager.issue((it + 1) & 1, src.at(wn), tid);
        const int16_t* x = stager.wait(it & 1, (it >> 1) & 1, src.at(w));
        logmel_window<T, G, NFB, int16_t>(x, work, tab, mel_out + w * (long long)(G::N_MELS * G::N_FRAMES), stride_m,
                                          stride_t, tid, NT);
    }
}

// Same, for float32 PCM already scaled to [-1, 1) (what the reference feeds its session,
// nanointerpreter.py:750, 771-775): read straight from global, no staging.
template <typename T, typename G, int NFB, int NT>
__global__ void __launch_bounds__(NT, 1)
frontend_f32_kernel(const float* __restrict__ pcm, long long n_windows, FrontendTables<T> tab,
                    float* __restrict__ mel_out, int time_major) {
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x;
    cplx<T>* work = reinterpret_cast<cplx<T>*>(smem);
    const int stride_m = time_major ? 1 : G::N_FRAMES;
    const int stride_t = time_major ? G::N_MELS : 1;
    for (long long w = blockIdx.x; w < n_windows; w += gridDim.x)
        logmel_window<T, G, NFB, float>(pcm + w * G::CLIP, work, tab,
                                        mel_out + w * (long long)(G::N_MELS * G::N_FRAMES), stride_m, stride_t, tid, NT);
}

}  // namespace nww
