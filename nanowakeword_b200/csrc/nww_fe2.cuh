// nww_fe2.cuh — front end v2 for the NS40x98 geometry (frame 400 -> FFT 512, hop 160, 40 mels):
// int16 PCM window in shared memory -> log-mel (dB), by one 512-thread CTA.
//
// Same arithmetic contract as nww_frontend.cuh (reference MelSpectrogram + AmplitudeToDB,
// nanowakeword/modules/architectures.py:830-837, 869-878; _export/onnx.py:27-83; int16 scaling of
// nanowakeword/interpreter/nanointerpreter.py:750) — FFT and power spectrum in FP64 — but organised
// for the SM instead of for generality:
//
//   * the CTA is four independent groups of 128 threads; a group transforms two packed complex
//     FFT-512 at a time (frames 4b .. 4b+3 of batch b) and synchronises only with itself through a
//     named barrier (bar.sync id, 128), so one group's barrier wait is another group's issue slot;
//   * radix 8 x 8 x 8 decimation in frequency; every thread owns one radix-8 butterfly in passes 1
//     and 2, so its Hann weights (pass 1) are registers and its twiddles come from two small
//     pre-arranged shared tables (conflict-free rows);
//   * the work buffer index is p + (p >> 6): with 16-byte complex elements every quarter-warp of
//     every pass touches 8 distinct 16-byte bank groups;
//   * pass 3 is one warp per FFT: lane c (1..31) owns the two final butterflies that hold bins
//     {64 i + c} and {64 i + 64 - c}, i.e. every (k, N-k) pair needed to split the two packed
//     real frames lives in one thread's registers; lane 0 owns the two self-paired butterflies
//     (bins 64 i and 64 i + 32).  The spectrum is never written back: the power of both frames
//     goes straight to a small FP32 table;
//   * the sparse triangular mel filters run in FP32 from that table, then 10*log10 — by the two
//     warps of the group that are NOT in pass 3, one batch behind (the power table is double-
//     buffered), so the filterbank costs no time on the critical path; the pass-3 pair alternates
//     between warps {0,1} and {2,3} every batch so that all four SM sub-partitions see the same
//     FP64 load.
#pragma once

#include "nww_stage.cuh"
#include "nww_tc.cuh"

namespace nww {

struct Fe2 {
    static constexpr int NT = 512, NGROUP = 4, GT = 128, NFB = 2;
    static constexpr int N = 512, NPAD = 520;                  // idx(p) = p + (p >> 6)
    static constexpr int PW_PITCH = 264;                       // 257 power bins per frame, padded
    static constexpr int N_FFT_TOTAL = 49, N_BATCH = 25;       // 98 frames = 49 packed FFTs = 24.5 batches of two
    static constexpr size_t kWorkBytes = (size_t)NGROUP * NFB * NPAD * sizeof(cplx<double>);   // 66560
    static constexpr size_t kPowBytes = (size_t)2 * NGROUP * NFB * 2 * PW_PITCH * sizeof(float);   // 33792 (double-buffered)
    static constexpr int N_TW = 7 * 64 + 7 * 8;                // twiddle entries
    static constexpr int MEL_ROW = 36;                         // padded filter row (32 weights + 4 to spread banks)
    static constexpr int MEL_MAX = 40;
    // "tables" block in shared memory: twiddles | padded mel weights [40][36] | per-filter (first bin, groups of 4)
    static constexpr size_t kTwBytes = (size_t)N_TW * sizeof(cplx<double>) + (size_t)MEL_MAX * MEL_ROW * sizeof(float) +
                                       (size_t)MEL_MAX * 2 * sizeof(int);                      // 14144
    static constexpr size_t kScratchBytes = kWorkBytes + kPowBytes;                            // reusable between windows
};

__device__ __forceinline__ int fe2_idx(int p) { return p + (p >> 6); }

// int16 -> double through the 2^52 trick: one integer op + one DADD instead of a 64-bit I2F.
__device__ __forceinline__ double fe2_i16_to_f64(int16_t v) {
#ifndef NWW_CPUSIM
    const uint32_t lo = (uint32_t)((int)v + 32768);
    return __hiloint2double(0x43300000, (int)lo) - 4503599627403264.0;     // 2^52 + 2^15
#else
    return (double)v;
#endif
}

// Tables in shared memory, built once per CTA.  Twiddles from the engine's exp(-2 pi i k / 512) table:
//   tw1[(q-1) * 64 + j]  = W512^(j q)      (pass 1, j = 0..63)
//   tw2[(q-1) * 8 + j2]  = W512^(8 j2 q)   (pass 2, j2 = 0..7)
// Mel filters for the vectorised dot product: filter m starts at bin ks and has cnt weights; its padded row holds
// the weights of bins [ks & ~3, ...) in groups of four (zeros outside the filter), so power and weights are both
// read as aligned 128-bit shared loads.
__device__ __forceinline__ void fe2_build_tables(cplx<double>* tw_smem, const FrontendTables<double>& tab, int tid,
                                                 int nthreads) {
    const cplx<double>* __restrict__ tw512 = tab.twiddle;
    for (int i = tid; i < 7 * 64; i += nthreads) tw_smem[i] = tw512[((i & 63) * ((i >> 6) + 1)) & 511];
    for (int i = tid; i < 7 * 8; i += nthreads) tw_smem[7 * 64 + i] = tw512[(8 * (i & 7) * ((i >> 3) + 1)) & 511];
    float* wpad = reinterpret_cast<float*>(tw_smem + Fe2::N_TW);
    int* meta = reinterpret_cast<int*>(wpad + Fe2::MEL_MAX * Fe2::MEL_ROW);
    if (tab.mel_vec_ok) {
        for (int i = tid; i < GeoNS40x98::N_MELS * Fe2::MEL_ROW; i += nthreads) {
            const int m = i / Fe2::MEL_ROW, j = i - m * Fe2::MEL_ROW;
            const int ks = tab.mel_start[m], cnt = tab.mel_count[m];
            const int bin = (ks & ~3) + j;
            wpad[i] = (bin >= ks && bin < ks + cnt) ? tab.mel_w[tab.mel_woff[m] + bin - ks] : 0.0f;
        }
        for (int m = tid; m < GeoNS40x98::N_MELS; m += nthreads) {
            const int ks = tab.mel_start[m], cnt = tab.mel_count[m];
            meta[2 * m] = ks & ~3;
            meta[2 * m + 1] = cnt ? ((ks & 3) + cnt + 3) >> 2 : 0;
        }
    }
}

// sum_k w_m[k] P[k] for filter m over one frame's power row (16-byte aligned, bins >= 257 must read as finite).
// The summation order is part of the contract between the two front ends (fe2 / fe3): the stream mel ring must
// be bit-identical to recomputing a window.
__device__ __forceinline__ float fe2_mel_dot(const float* __restrict__ prow, int m, const cplx<double>* __restrict__ tw_smem,
                                             const FrontendTables<double>& tab) {
    if (tab.mel_vec_ok) {
        const float* wpad = reinterpret_cast<const float*>(tw_smem + Fe2::N_TW);
        const int* meta = reinterpret_cast<const int*>(wpad + Fe2::MEL_MAX * Fe2::MEL_ROW);
        const int k0 = meta[2 * m], n4 = meta[2 * m + 1];
        const float4* __restrict__ p4 = reinterpret_cast<const float4*>(prow + k0);
        const float4* __restrict__ w4 = reinterpret_cast<const float4*>(wpad + m * Fe2::MEL_ROW);
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        for (int g = 0; g < n4; ++g) {
            const float4 p = p4[g], w = w4[g];
            a0 = fmaf(w.x, p.x, a0);
            a1 = fmaf(w.y, p.y, a1);
            a2 = fmaf(w.z, p.z, a2);
            a3 = fmaf(w.w, p.w, a3);
        }
        return (a0 + a1) + (a2 + a3);
    }
    const int ks = __ldg(tab.mel_start + m);
    const int cnt = __ldg(tab.mel_count + m);
    const float* __restrict__ w = tab.mel_w + __ldg(tab.mel_woff + m);
    const float* __restrict__ p = prow + ks;
    float acc0 = 0.0f, acc1 = 0.0f;
    int i = 0;
    for (; i + 1 < cnt; i += 2) {
        acc0 = fmaf(__ldg(w + i), p[i], acc0);
        acc1 = fmaf(__ldg(w + i + 1), p[i + 1], acc1);
    }
    if (i < cnt) acc0 = fmaf(__ldg(w + i), p[i], acc0);
    return acc0 + acc1;
}

// Both frames of a packed FFT for one filter: the weights are loaded once for the two dot products, and every load of
// a step is issued before the first multiply (the filter length is a run-time value, so the plain loop above exposes one
// shared-memory round trip per four bins — tools/probe/fft_probe.cu: 42 % of a lone warp's time per FFT).  Same
// operations in the same order per (filter, frame) as fe2_mel_dot: bit-identical results.
__device__ __forceinline__ void fe2_mel_dot2(const float* __restrict__ prow_a, const float* __restrict__ prow_b, int m,
                                             const cplx<double>* __restrict__ tw_smem, const FrontendTables<double>& tab,
                                             float* __restrict__ out_a, float* __restrict__ out_b) {
    if (!tab.mel_vec_ok) {
        *out_a = fe2_mel_dot(prow_a, m, tw_smem, tab);
        *out_b = fe2_mel_dot(prow_b, m, tw_smem, tab);
        return;
    }
    const float* wpad = reinterpret_cast<const float*>(tw_smem + Fe2::N_TW);
    const int2 mt = reinterpret_cast<const int2*>(wpad + Fe2::MEL_MAX * Fe2::MEL_ROW)[m];
    const int k0 = mt.x, n4 = mt.y;
    const float4* __restrict__ pa4 = reinterpret_cast<const float4*>(prow_a + k0);
    const float4* __restrict__ pb4 = reinterpret_cast<const float4*>(prow_b + k0);
    const float4* __restrict__ w4 = reinterpret_cast<const float4*>(wpad + m * Fe2::MEL_ROW);
    constexpr int MAXG = Fe2::MEL_ROW / 4;                     // 9 groups of four bins
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f, b0 = 0.0f, b1 = 0.0f, b2 = 0.0f, b3 = 0.0f;
#pragma unroll
    for (int g0 = 0; g0 < MAXG; g0 += 3) {                      // three groups (nine 128-bit loads) in flight at a time
        float4 pa[3], pb[3], w[3];
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (g0 + j < n4) {
                w[j] = w4[g0 + j];
                pa[j] = pa4[g0 + j];
                pb[j] = pb4[g0 + j];
            }
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (g0 + j < n4) {
                a0 = fmaf(w[j].x, pa[j].x, a0); a1 = fmaf(w[j].y, pa[j].y, a1);
                a2 = fmaf(w[j].z, pa[j].z, a2); a3 = fmaf(w[j].w, pa[j].w, a3);
                b0 = fmaf(w[j].x, pb[j].x, b0); b1 = fmaf(w[j].y, pb[j].y, b1);
                b2 = fmaf(w[j].z, pb[j].z, b2); b3 = fmaf(w[j].w, pb[j].w, b3);
            }
    }
    *out_a = (a0 + a1) + (a2 + a3);
    *out_b = (b0 + b1) + (b2 + b3);
}

// A batch = up to four consecutive frames (hop 160) = two packed FFTs.  `x` points at sample 0 of the
// batch's first frame in shared memory; frames beyond `nframes` are computed on whatever follows in
// shared memory (it must be readable up to x + 3 * 160 + 448) and dropped.
struct Fe2Batch {
    const int16_t* x;
    int nframes;
};

// The front-end core: `n_batches` batches dealt round-robin to the four groups.  batch_of(b) -> Fe2Batch,
// store(b, frame_in_batch, mel_bin, dB).  scratch: Fe2::kScratchBytes of shared memory; tw_smem: the
// tables above.  Must be called by all 512 threads; ends with __syncthreads().
template <typename BatchFn, typename StoreFn>
__device__ __forceinline__ void fe2_run(int n_batches, BatchFn batch_of, StoreFn store, unsigned char* __restrict__ scratch,
                                        const cplx<double>* __restrict__ tw_smem, const FrontendTables<double>& tab,
                                        int tid) {
    const int g = tid >> 7;                   // group
    const int t = tid & 127;                  // thread in group
    const int wg = t >> 5, lane = t & 31;
    cplx<double>* work = reinterpret_cast<cplx<double>*>(scratch) + (size_t)g * Fe2::NFB * Fe2::NPAD;
    // power tables: [buffer 2][group][fft 2][frame 2][PW_PITCH]
    float* pw_base = reinterpret_cast<float*>(scratch + Fe2::kWorkBytes) + (size_t)g * Fe2::NFB * 2 * Fe2::PW_PITCH;
    constexpr int kPwBuf = Fe2::NGROUP * Fe2::NFB * 2 * Fe2::PW_PITCH;       // floats per buffer
    const cplx<double>* tw1 = tw_smem;
    const cplx<double>* tw2 = tw_smem + 7 * 64;

    // passes 1 and 2: FFT slot f, butterfly jj
    const int f = wg >> 1;
    const int jj = ((wg & 1) << 5) | lane;
    cplx<double>* wf = work + f * Fe2::NPAD;
    // Hann * 2^-15 (int16 scaling) * 1/2 (so that |Z_k +- conj Z_{N-k}|^2 needs no final /4), zero beyond the frame
    double wreg[7];
#pragma unroll
    for (int m = 0; m < 7; ++m) {
        const int n = jj + 64 * m;
        wreg[m] = (n < GeoNS40x98::WIN) ? 0.5 * tab.window[n] : 0.0;
    }
    const int b2 = jj >> 3, j2 = jj & 7;

    // mel + dB of batch `mb` (4 frames x 40 filters = 160 tasks) from power table `p0`, by `nth` threads
    // of which this is number `t2`.  Filters are dealt in octets ordered long, short, ... so that the two
    // warps of a pair get about the same number of filter taps (the triangles widen with frequency).
    auto mel_batch2 = [&](int mb, int mb_frames, const float* __restrict__ p0, int t2, int nth) {
        for (int task = t2; task < 4 * GeoNS40x98::N_MELS; task += nth) {
            const int fr = task & 3;                       // FFT slot fr >> 1, packed frame fr & 1
            const int slot = task >> 2;                    // 0..39 -> filter octets in the order 4, 3, 0, 2, 1
            const int oct = slot >> 3;
            const int m = ((oct == 0) ? 32 : (oct == 1) ? 24 : (oct == 2) ? 0 : (oct == 3) ? 16 : 8) + (slot & 7);
            if (fr >= mb_frames) continue;
            const float pm = fe2_mel_dot(p0 + fr * Fe2::PW_PITCH, m, tw_smem, tab);
            store(mb, fr, m, (pm <= tab.amin) ? tab.floor_db : 10.0f * log10f(pm));
        }
    };

    int it = 0, b_prev = -1, nf_prev = 0;
    for (int b = g; b < n_batches; b += Fe2::NGROUP, ++it) {
        float* pw = pw_base + (it & 1) * kPwBuf;
        const Fe2Batch bt = batch_of(b);
        const bool fvalid = 2 * f < bt.nframes;
        // ---- pass 1: L = 512, inputs straight from PCM (frames 4b + 2f and 4b + 2f + 1) -------------
        if (fvalid) {
            const int16_t* xa = bt.x + 2 * f * GeoNS40x98::HOP + jj;
            cplx<double> v[8];
#pragma unroll
            for (int m = 0; m < 7; ++m) {
                const double sa = fe2_i16_to_f64(xa[64 * m]);
                const double sb = fe2_i16_to_f64(xa[64 * m + GeoNS40x98::HOP]);
                v[m] = {wreg[m] * sa, wreg[m] * sb};
            }
            v[7] = {0.0, 0.0};
            SmallDft<double, 8>::run(v);
            wf[jj] = v[0];
#pragma unroll
            for (int q = 1; q < 8; ++q) wf[jj + 65 * q] = cmul(v[q], tw1[(q - 1) * 64 + jj]);    // idx(jj + 64 q)
        }
        named_bar_sync(1 + g, Fe2::GT);
        // ---- pass 2: L = 64 inside block b2 -------------------------------------------------------------
        if (fvalid) {
            cplx<double>* blk = wf + 65 * b2 + j2;                 // idx(64 b2 + j2 + 8 m) = 65 b2 + j2 + 8 m
            cplx<double> v[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) v[m] = blk[8 * m];
            SmallDft<double, 8>::run(v);
            blk[0] = v[0];
#pragma unroll
            for (int q = 1; q < 8; ++q) blk[8 * q] = cmul(v[q], tw2[(q - 1) * 8 + j2]);
        }
        named_bar_sync(1 + g, Fe2::GT);
        // ---- pass 3 + power (one warp per FFT, lane c)  ||  mel + dB of the previous batch -----------------
        const int p3_first = (it & 1) << 1;                    // warps {0,1} on even batches, {2,3} on odd ones
        const int fs = wg - p3_first;                          // FFT slot for a pass-3 warp
        if (fs >= 0 && fs < Fe2::NFB) {
          if (2 * fs < bt.nframes) {
            const int c = lane;
            const int cb = (c == 0) ? 32 : 64 - c;
            // butterfly with residue r = 8 q2 + b holds positions 64 b + 8 q2 + m -> idx = 65 (r & 7) + 8 (r >> 3) + m
            const cplx<double>* pa = work + fs * Fe2::NPAD + 65 * (c & 7) + 8 * (c >> 3);
            const cplx<double>* pb = work + fs * Fe2::NPAD + 65 * (cb & 7) + 8 * (cb >> 3);
            cplx<double> A[8], B[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) A[m] = pa[m];
#pragma unroll
            for (int m = 0; m < 8; ++m) B[m] = pb[m];
            SmallDft<double, 8>::run(A);      // A[i] = Z[64 i + c]
            SmallDft<double, 8>::run(B);      // B[i] = Z[64 i + cb]
            cplx<double> U[8], W[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                U[i] = A[i];
                W[i] = B[7 - i];              // N - (64 i + c) = 64 (7 - i) + (64 - c)
            }
            const bool special = (c == 0);
            if (special) {
                // lane 0: bins 0, 64, 128, 192 pair inside A (i <-> 8 - i), bins 32 .. 224 inside B (i <-> 7 - i)
                W[0] = A[0]; W[1] = A[7]; W[2] = A[6]; W[3] = A[5];
                U[4] = B[0]; U[5] = B[1]; U[6] = B[2]; U[7] = B[3];
                W[4] = B[7]; W[5] = B[6]; W[6] = B[5]; W[7] = B[4];
            }
            float* pwa = pw + (fs * 2 + 0) * Fe2::PW_PITCH;
            float* pwb = pw + (fs * 2 + 1) * Fe2::PW_PITCH;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                // X_A = Z_k + conj(Z_{N-k}),  X_B = (Z_k - conj(Z_{N-k})) / i   (the 1/2 is in the window)
                const double ar = U[i].x + W[i].x, ai = U[i].y - W[i].y;
                const double br = U[i].y + W[i].y, bi = U[i].x - W[i].x;
                const int bin = (i < 4) ? (64 * i + c) : (special ? (64 * (i - 4) + 32) : (64 * (7 - i) + 64 - c));
                pwa[bin] = (float)(ar * ar + ai * ai);
                pwb[bin] = (float)(br * br + bi * bi);
            }
            if (special) {                     // bin 256 pairs with itself
                const double ar = 2.0 * A[4].x, br = 2.0 * A[4].y;
                pwa[256] = (float)(ar * ar);
                pwb[256] = (float)(br * br);
            } else if (c < 8) {                // the padded mel rows read a few bins past Nyquist: keep them finite
                pwa[256 + c] = 0.0f;
                pwb[256 + c] = 0.0f;
            }
          }
        } else if (b_prev >= 0) {
            const int wm = (wg - p3_first) & 3;                // 2 or 3 -> mel warp 0 / 1
            mel_batch2(b_prev, nf_prev, pw_base + ((it - 1) & 1) * kPwBuf, ((wm - 2) << 5) | lane, 64);
        }
        named_bar_sync(1 + g, Fe2::GT);
        b_prev = b;
        nf_prev = bt.nframes;
    }
    // the group's last batch: all four warps
    if (b_prev >= 0) mel_batch2(b_prev, nf_prev, pw_base + ((it - 1) & 1) * kPwBuf, t, Fe2::GT);
    __syncthreads();
}

// One window.  pcm: 16000 int16 in shared memory.  Writes mel[m * stride_m + t * stride_t] (shared or global).
__device__ __forceinline__ void fe2_logmel_window(const int16_t* __restrict__ pcm, unsigned char* __restrict__ scratch,
                                                  const cplx<double>* __restrict__ tw_smem,
                                                  const FrontendTables<double>& tab, float* __restrict__ mel,
                                                  int stride_m, int stride_t, int tid) {
    fe2_run(
        Fe2::N_BATCH,
        [&](int b) {
            const int left = GeoNS40x98::N_FRAMES - 4 * b;
            return Fe2Batch{pcm + 4 * b * GeoNS40x98::HOP, left < 4 ? left : 4};
        },
        [&](int b, int fr, int m, float db) { mel[m * stride_m + (4 * b + fr) * stride_t] = db; }, scratch, tw_smem, tab, tid);
}

}  // namespace nww
