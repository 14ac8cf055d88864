// nww_stream.cuh — per-stream sliding audio rings in HBM (K9 data structure).
//
// Replaces the reference's per-model ``deque(maxlen=clip_samples)`` of Python floats and its
// cumulative sample counter (reference nanowakeword/interpreter/nanointerpreter.py:176-183 create,
// :750-753 append, :755-756 "score the last clip_samples once enough audio arrived", :719-733 reset)
// for MANY independent streams at once.
//
// Layout: stream s owns int16 ring[s][2 R + 8], R = clip_samples.  Every sample is written twice, at
// p and p + R ("mirrored ring"), so the most recent R samples are always one CONTIGUOUS run starting
// at the write position — which is what a TMA bulk copy wants; no wrap handling in the consumers.
// The run may start at any int16 boundary; PcmStager (nww_stage.cuh) copes with the 16-byte skew, and
// the 8 spare samples at the end keep its 16 extra bytes inside the allocation.
#pragma once

#include "nww_common.cuh"

namespace nww {

struct StreamState {
    int16_t* ring;            // [n_streams][2 R + 8]
    int* wpos;                // [n_streams] next write position in [0, R)
    long long* count;         // [n_streams] samples received since open / reset (never wraps in practice)
    long long* win_off;       // [n_streams] element offset of the stream's current window in `ring`
    long long n_streams;
    int R;                    // clip_samples
    __host__ __device__ long long pitch() const { return 2ll * R + 8; }
};

// One CTA per stream: append chunk_len samples (only the last R of them matter), advance the write
// position and counter, publish the window offset.
__global__ void __launch_bounds__(256)
stream_append_kernel(StreamState st, const int16_t* __restrict__ chunks, int chunk_len) {
    const long long s = blockIdx.x;
    if (s >= st.n_streams) return;
    const int R = st.R;
    int16_t* ring = st.ring + s * st.pitch();
    const int wp = st.wpos[s];
    const int16_t* src = chunks + s * (long long)chunk_len;
    const int skip = chunk_len > R ? chunk_len - R : 0;          // older samples would be overwritten anyway
    const int n = chunk_len - skip;
    src += skip;
    const int wp0 = (int)(((long long)wp + skip) % R);
    if (((wp0 | n) & 7) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        // common case (chunk of 1280 on an aligned stream): 16-byte moves
        for (int i = threadIdx.x * 8; i < n; i += blockDim.x * 8) {
            const uint4 v = *reinterpret_cast<const uint4*>(src + i);
            int p = wp0 + i;
            if (p >= R) p -= R;
            *reinterpret_cast<uint4*>(ring + p) = v;
            *reinterpret_cast<uint4*>(ring + p + R) = v;
        }
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int16_t v = src[i];
            int p = wp0 + i;
            if (p >= R) p -= R;
            ring[p] = v;
            ring[p + R] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = (int)(((long long)wp + chunk_len) % R);
        st.wpos[s] = nw;
        st.count[s] += chunk_len;
        st.win_off[s] = s * st.pitch() + nw;      // oldest of the last R samples sits at the write position
    }
}

// Reset streams (ids == nullptr: all): zero the ring, the write position and the counter.
__global__ void __launch_bounds__(256)
stream_reset_kernel(StreamState st, const long long* __restrict__ ids, long long n_ids) {
    const long long j = blockIdx.x;
    if (j >= n_ids) return;
    const long long s = ids ? ids[j] : j;
    if (s < 0 || s >= st.n_streams) return;
    int16_t* ring = st.ring + s * st.pitch();
    uint4* r4 = reinterpret_cast<uint4*>(ring);                 // pitch is a multiple of 8 samples = 16 bytes
    const int n16 = (int)(st.pitch() / 8);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) r4[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        st.wpos[s] = 0;
        st.count[s] = 0;
        st.win_off[s] = s * st.pitch();
    }
}

// Selective scoring: window offsets of the listed streams, and their scores back to the per-stream array.
__global__ void __launch_bounds__(256)
stream_select_offsets_kernel(StreamState st, const long long* __restrict__ ids, long long n_ids, long long* __restrict__ off) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_ids) off[i] = st.win_off[ids[i]];
}
__global__ void __launch_bounds__(256)
stream_select_scatter_kernel(StreamState st, const long long* __restrict__ ids, long long n_ids, const float* __restrict__ sel,
                             float* __restrict__ scores) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_ids) {
        const long long s = ids[i];
        scores[s] = st.count[s] < st.R ? 0.0f : sel[i];          // nanointerpreter.py:755: no score before clip_samples arrived
    }
}

// A stream that has not yet received clip_samples reports 0 (nanointerpreter.py:755, 785-786).
__global__ void __launch_bounds__(256)
stream_mask_kernel(StreamState st, float* __restrict__ scores) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < st.n_streams && st.count[s] < st.R) scores[s] = 0.0f;
}

}  // namespace nww
