// nww_stream_mel.cuh — incremental log-mel for streams (NS40x98 geometry): the K9 "ring-buffer kernel".
//
// The reference re-runs its whole graph on the last clip_samples at every predict(chunk)
// (nanointerpreter.py:755-756, 783): with 1280-sample chunks, 90 of the 98 STFT frames of a window are
// the same samples, windowed the same way, as 90 frames of the previous window.  Without centring
// (NS40x98) frame t of a window is exactly frame t + 8 of the previous one, so the engine keeps, per
// stream, a ring of log-mel frames indexed by ABSOLUTE frame number a (frame a covers stream samples
// [160 a, 160 a + 400)) and computes only the frames a chunk completes — bit-identical to recomputing
// the window, 12x less front-end work at the reference's 80 ms hop.
//
// Layout: mel ring [n_streams][40][2 x 98] float, mirrored along time like the PCM ring, so the 98
// frames of the current window are contiguous per mel row starting at slot (a_new + 1) mod 98, where
// a_new = count / 160 - 3 is the newest complete frame.
// Requires every chunk since open/reset to be a multiple of the hop (160); the engine falls back to the
// full-window path otherwise.
#pragma once

#include "nww_fe3.cuh"
#include "nww_stream.cuh"

namespace nww {

struct SMel {
    static constexpr int T = GeoNS40x98::N_FRAMES, F = GeoNS40x98::N_MELS, HOP = GeoNS40x98::HOP;
    static constexpr int ROW = 2 * T;                              // mirrored row pitch
    static constexpr int STREAM_FLOATS = F * ROW;                  // 7840 floats = 31 360 B per stream
    static constexpr int SPB = 8;                                  // streams per CTA iteration
    static constexpr int MAX_NEW = 16;                             // frames per push handled incrementally
    static constexpr int PCM_SLOT = 160 * (MAX_NEW - 1) + 400 + 160 * 3 + 64;   // samples per stream in shared memory (+ slack for dropped frames)
    static constexpr size_t kScratch = (Fe2::kScratchBytes + 127) / 128 * 128;
    static constexpr size_t kTw = (Fe2::kTwBytes + 127) / 128 * 128;
    static constexpr size_t kPcm = ((size_t)SPB * PCM_SLOT * sizeof(int16_t) + 127) / 128 * 128;
    static constexpr size_t kTotal = kScratch + kTw + kPcm;
};

// Where a launch finds the log-mel of its windows in stream mode: window w of the launch is stream s0 + w.
// ring == nullptr means "not in stream mode".
struct MelRingRef {
    const float* ring;
    const long long* count;     // per-stream sample counters (position of the window in the ring)
    long long s0;
    // Selective scoring (nww_stream_push_select: the cascade's verifier scores only the streams its gate let through,
    // nanointerpreter.py:758-769): window s0 + w of the launch is stream ids[s0 + w] instead of stream s0 + w.
    const long long* ids = nullptr;
    __device__ __forceinline__ long long stream(long long w) const { return ids ? ids[s0 + w] : s0 + w; }
};

__device__ __forceinline__ int smel_slot(long long a) {
    int r = (int)(a % SMel::T);
    return r < 0 ? r + SMel::T : r;
}

// After stream_append_kernel: compute the n_new = chunk_len / 160 frames the chunk completed, for every stream.
__global__ void __launch_bounds__(Fe2::NT, 1)
stream_mel_update_kernel(StreamState st, float* __restrict__ mel_ring, FrontendTables<double> tab, int n_new) {
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x;
    cplx<double>* tw = reinterpret_cast<cplx<double>*>(smem + SMel::kScratch);
    int16_t* pcm_s = reinterpret_cast<int16_t*>(smem + SMel::kScratch + SMel::kTw);
    fe2_build_tables(tw, tab, tid, Fe2::NT);
    const int nbps = (n_new + 3) >> 2;                              // batches per stream
    const int n_samp = SMel::HOP * (n_new - 1) + GeoNS40x98::WIN;   // samples the new frames span
    const int w_off = st.R - SMel::HOP * (n_new + 2);               // their offset inside the stream's current window
    for (long long s0 = (long long)blockIdx.x * SMel::SPB; s0 < st.n_streams; s0 += (long long)gridDim.x * SMel::SPB) {
        const int ns = (int)((st.n_streams - s0 < SMel::SPB) ? (st.n_streams - s0) : SMel::SPB);
        __syncthreads();                                            // previous iteration done with pcm_s (and tw built)
        for (int i = tid; i < ns * SMel::PCM_SLOT; i += Fe2::NT) {
            const int sl = i / SMel::PCM_SLOT, j = i - sl * SMel::PCM_SLOT;
            pcm_s[i] = (j < n_samp) ? st.ring[st.win_off[s0 + sl] + w_off + j] : (int16_t)0;
        }
        __syncthreads();
        fe2_run(
            ns * nbps,
            [&](int b) {
                const int sl = b / nbps, q = b - sl * nbps;
                const int left = n_new - 4 * q;
                return Fe2Batch{pcm_s + sl * SMel::PCM_SLOT + 4 * q * SMel::HOP, left < 4 ? left : 4};
            },
            [&](int b, int fr, int m, float db) {
                const int sl = b / nbps, q = b - sl * nbps;
                const long long s = s0 + sl;
                const long long a = st.count[s] / SMel::HOP - 3 - (n_new - 1) + 4 * q + fr;   // absolute frame number
                float* row = mel_ring + s * SMel::STREAM_FLOATS + m * SMel::ROW + smel_slot(a);
                row[0] = db;
                row[SMel::T] = db;
            },
            smem, tw, tab, tid);
    }
}

// Current window of every stream out of the mel ring: (n, F, T) or, time_major, (n, T, F).
__global__ void __launch_bounds__(256)
stream_mel_gather_kernel(MelRingRef ring, long long n, float* __restrict__ out, int time_major) {
    const long long total = n * (long long)(SMel::F * SMel::T);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long w = i / (SMel::F * SMel::T);
        const int r = (int)(i - w * (SMel::F * SMel::T));
        const int m = r / SMel::T, t = r - m * SMel::T;             // reads are contiguous along t
        const long long s = ring.stream(w);
        const int head = smel_slot(ring.count[s] / SMel::HOP - 3 + 1);
        const float v = ring.ring[s * SMel::STREAM_FLOATS + m * SMel::ROW + head + t];
        out[w * (long long)(SMel::F * SMel::T) + (time_major ? t * SMel::F + m : r)] = v;
    }
}

// The last n_tail frames of every stream's current window out of the mel ring, TIME-major, at the place a time-major
// (n, T, F) log-mel buffer holds them: out[s][(t0 + pos) * F + m] — one contiguous run of n_tail * F floats per stream.
// One warp per stream: reads run along time inside the ring's mel rows, a padded shared tile transposes, writes are
// coalesced.  Feeds the TCN's row-GEMM layers in stream mode.
constexpr int kMelTailWarps = 8, kMelTailMax = 32;       // n_tail <= 32
__global__ void __launch_bounds__(kMelTailWarps * 32)
stream_mel_tail_kernel(MelRingRef ring, long long n, float* __restrict__ out, int t0, int n_tail) {
    __shared__ float tile[kMelTailWarps][kMelTailMax][SMel::F + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long w = (long long)blockIdx.x * kMelTailWarps + warp; w < n; w += (long long)gridDim.x * kMelTailWarps) {
        const long long s = ring.stream(w);
        const int head = smel_slot(ring.count[s] / SMel::HOP - 3 + 1);
        const float* src = ring.ring + s * SMel::STREAM_FLOATS + head + t0;
        if (lane < n_tail)
            for (int m = 0; m < SMel::F; ++m) tile[warp][lane][m] = src[m * SMel::ROW + lane];
        __syncwarp();
        float* dst = out + w * (long long)(SMel::F * SMel::T) + (long long)t0 * SMel::F;
        for (int i = lane; i < n_tail * SMel::F; i += 32) dst[i] = tile[warp][i / SMel::F][i % SMel::F];
        __syncwarp();
    }
}

// ----------------------------------------------------------------------------------------
// K9 ingest step, fused: append the chunk to the PCM ring AND compute the log-mel frames it completes, one launch.
// One WARP per stream (persistent CTAs of 12 warps): the warp stages [320 samples before the chunk | the chunk] in its
// shared slot — the 320 older samples are one contiguous run of the mirrored ring, the chunk comes straight from the
// caller's buffer and is written to both copies of the ring on the way —, then runs ceil(n_new / 2) warp-private packed
// FFTs (fe3_warp_fft: the arithmetic of the batch front end, so the ring stays bit-identical to recomputing windows) and
// writes the dB frames to both copies of the mel ring.  Replaces stream_append_kernel + stream_mel_update_kernel when
// every chunk is a multiple of the hop (<= 16 frames).
// ----------------------------------------------------------------------------------------
struct SPush {
    static constexpr int NW = 12, NT = NW * 32;          // 12 warps x 168 registers: at 14 x 128 the FFT loop spilled and its reloads (long scoreboard) were 22 % of the samples
    static constexpr int OLD = 2 * SMel::HOP;                                   // samples before the chunk that the new frames reach back to
    static constexpr int PCM_SLOT = OLD + SMel::HOP * SMel::MAX_NEW + SMel::HOP;   // + one hop: the dummy second frame of an odd count
    static constexpr size_t kWork = (size_t)NW * Fe3::NPAD * sizeof(cplx<double>);
    static constexpr size_t kPcm = ((size_t)NW * PCM_SLOT * sizeof(int16_t) + 127) / 128 * 128;
    static constexpr size_t kTotal = kWork + Fe3::kTwBytes + Fe3::kWinBytes + kPcm;
};

__global__ void __launch_bounds__(SPush::NT, 1)
stream_push_mel_kernel(StreamState st, const int16_t* __restrict__ chunks, int chunk_len, float* __restrict__ mel_ring,
                       FrontendTables<double> tab, long long s_begin, long long s_end /* streams [s_begin, s_end) of the bank */) {
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    cplx<double>* wb = reinterpret_cast<cplx<double>*>(smem) + (size_t)warp * Fe3::NPAD;
    cplx<double>* tw = reinterpret_cast<cplx<double>*>(smem + SPush::kWork);
    double* win_s = reinterpret_cast<double*>(smem + SPush::kWork + Fe3::kTwBytes);
    int16_t* slot = reinterpret_cast<int16_t*>(smem + SPush::kWork + Fe3::kTwBytes + Fe3::kWinBytes) + (size_t)warp * SPush::PCM_SLOT;
    fe2_build_tables(tw, tab, tid, SPush::NT);
    fe3_build_window(win_s, tab.window, tid, SPush::NT);
    for (int i = lane; i < SPush::PCM_SLOT; i += 32) slot[i] = 0;              // whatever the dummy frame reads is finite
    __syncthreads();
    const int R = st.R, n_new = chunk_len / SMel::HOP;
    for (long long s = s_begin + (long long)blockIdx.x * SPush::NW + warp; s < s_end; s += (long long)gridDim.x * SPush::NW) {
        int16_t* ring = st.ring + s * st.pitch();
        const int wp = st.wpos[s];                                              // a multiple of the hop: 16-byte aligned
        const long long cnt = st.count[s] + chunk_len;
        // the 320 samples before the write position: contiguous in the mirrored ring
        const int16_t* old = ring + (wp >= SPush::OLD ? wp - SPush::OLD : wp - SPush::OLD + R);
        for (int i = lane * 8; i < SPush::OLD; i += 32 * 8)
            *reinterpret_cast<uint4*>(slot + i) = *reinterpret_cast<const uint4*>(old + i);
        const int16_t* src = chunks + s * (long long)chunk_len;
        for (int i = lane * 8; i < chunk_len; i += 32 * 8) {
            const uint4 v = *reinterpret_cast<const uint4*>(src + i);
            *reinterpret_cast<uint4*>(slot + SPush::OLD + i) = v;
            int p = wp + i;
            if (p >= R) p -= R;
            *reinterpret_cast<uint4*>(ring + p) = v;
            *reinterpret_cast<uint4*>(ring + p + R) = v;
        }
        __syncwarp();
        const long long a0 = cnt / SMel::HOP - 3 - (n_new - 1);                 // absolute number of the first new frame
        float* mrow = mel_ring + s * SMel::STREAM_FLOATS;
#pragma unroll 1
        for (int f = 0; 2 * f < n_new; ++f) {
            const int sa = smel_slot(a0 + 2 * f), sb = smel_slot(a0 + 2 * f + 1);
            const bool has_b = 2 * f + 1 < n_new;
            fe3_warp_fft(slot + 2 * f * SMel::HOP, wb, win_s, tw, tab,
                         [&](int fr, int m, float db) {
                             if (fr && !has_b) return;
                             float* row = mrow + m * SMel::ROW + (fr ? sb : sa);
                             row[0] = db;
                             row[SMel::T] = db;
                         },
                         lane);
        }
        if (lane == 0) {
            int nw = wp + chunk_len;
            if (nw >= R) nw -= R;
            st.wpos[s] = nw;
            st.count[s] = cnt;
            st.win_off[s] = s * st.pitch() + nw;
        }
        __syncwarp();
    }
}

}  // namespace nww
