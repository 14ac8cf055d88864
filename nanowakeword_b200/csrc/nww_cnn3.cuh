// nww_cnn3.cuh — stage A of the CNN head, v3: the same arithmetic as nww_cnn2.cuh (PCM -> FP64 FFT log-mel ->
// conv1 + act + pool -> conv2 on tcgen05 + act + pool -> feature row), software-pipelined ACROSS the phases.
//
// Why (profiles/r01_v8_ncu_full.txt, profiles/r02_v9base_*): in v2 the four phases of a window — FP64 FFT, FP32
// conv1, tcgen05 conv2, TMEM epilogue — use four different pipes and run one after another inside the CTA, so no
// pipe is busy more than 25 % of the time and 1.8 warps of 4 per scheduler sit at CTA barriers.  Here the CTA is
// warp-specialised and the phases of consecutive column groups / windows overlap:
//
//   warps 0..6   "front end": 7 warps x 7 packed FFT-512 = the 49 FFTs (98 frames) of a window, warp-private
//                (fe3_warp_fft, unchanged arithmetic).  Round r (FFTs 7r .. 7r+6) produces log-mel frames
//                14r .. 14r+13 = column group ("slot") r of the zero-bordered mel plane, then signals full_mel[r].
//   warp 7       control: one thread re-fills the PCM buffer by TMA, SEGMENT by segment: round r reads samples
//                [2240 r, 2240 r + 2480) only, so segment r of the NEXT window is fetched as soon as the seven
//                front-end warps have finished round r of this one (consumed[r] -> cp.async.bulk -> full_pcm[r]).
//                One 32 KB PCM buffer behaves like a 7-stage pipeline; the front end never waits for a window.
//   warps 8..15  "conv": conv1 tasks in column-group order, each warp waiting only for the slot its tasks read, so
//                conv1 of window w runs UNDER the FFTs of window w (FP32 FMA pipe vs FP64 pipe); then one thread
//                issues the 216 tcgen05.mma of conv2 (parity planes / shifted descriptors, see nww_cnn2.cuh) and
//                the eight warps run the TMEM epilogue while the front end is already 1-2 rounds into window w+1.
//                done[r] tells the front end that slot r of the mel plane may be overwritten.
//
// The parity planes no longer overlay the FFT scratch (their zero border is written once per kernel, not once per
// window), the log-mel plane is single-buffered behind the full_mel / done handshake, TMEM holds one window.
// Shared memory: 7 x 8320 FFT scratch + 78848 planes + 17408 tables + 17472 mel + 18432 conv2 weights + 32128 PCM
// = 223 KB.  Scores are bit-identical to v2 (same operations in the same order on every value).
//
// Stream mode / float feeds (msrc.ring != nullptr): the front-end warps copy the window's log-mel from the stream
// ring (or a plain (n, F, T) buffer) into the plane slot by slot instead of computing it; the control warp idles.
#pragma once

#include "nww_cnn2.cuh"

#ifndef NWW_EXP
#define NWW_EXP 0        // developer experiments: 1 = conv warps skip conv1 + MMAs, 2 = skip the MMAs only
#endif

namespace nww {

struct Cnn3 {
    using D = Cnn2;
    static constexpr int NT = 512;
    static constexpr int N_FE_WARPS = 7, CTRL_WARP = 7, CONV_WARP0 = 8, N_CONV_WARPS = 8, CONV_NT = 256;
    static constexpr int N_SLOTS = 7;                          // column groups of 14 frames
    static constexpr int SLOT_FRAMES = 14;
    static constexpr int SEG = 2240;                           // PCM samples per segment = 7 FFTs x 320
    static constexpr int PCM_SAMPLES = 16008;                  // + 8: a window may start at any int16 boundary (16-byte skew)
    static constexpr int MEL_P = 104, MEL_ROWS = 42;           // 2 * MEL_P = 16 (mod 32): pooled rows y, y + 1 hit different banks
    static constexpr int CONV_ROUNDS = (D::CONV1_TASKS + CONV_NT - 1) / CONV_NT;   // 8
    static constexpr int BAR_CONV = 1;                         // named barrier of the 256 conv threads
    static constexpr int BAR_FE = 2;                           // named barrier of the 224 front-end threads

    static constexpr size_t oWork = 0;
    static constexpr size_t kWork = (size_t)N_FE_WARPS * Fe3::NPAD * sizeof(cplx<double>);            // 58240
    static constexpr size_t oA1 = align_up(oWork + kWork, 1024);
    static constexpr size_t oTw = align_up(oA1 + D::A1_BYTES, 128);                                  // tables, then Hann
    static constexpr size_t oMel = oTw + Fe3::kTwBytes + Fe3::kWinBytes;
    static constexpr size_t kMel = align_up(sizeof(float) * MEL_ROWS * MEL_P, 128);
    static constexpr size_t oW2 = oMel + kMel;
    static constexpr size_t oSmall = oW2 + D::W2_BYTES;
    static constexpr size_t oBars = oSmall + D::kSmall;
    static constexpr size_t kBars = 256;                       // 30 mbarriers + TMEM base slot
    static constexpr size_t oPcm = oBars + kBars;
    static constexpr size_t kTotal = oPcm + align_up(PCM_SAMPLES * sizeof(int16_t), 128);
    static_assert(kTotal <= 232448, "shared memory budget of one CTA");
};

// first pooled column and number of pooled columns whose conv1 inputs are complete once mel slot r is: column x reads
// frames 2x - 1 .. 2x + 2, so slot r (frames <= 14 r + 13) completes x <= 7 r + 5 (x = 48 reads the zero border)
__host__ __device__ constexpr int cnn3_x0(int r) { return r == 0 ? 0 : 7 * r - 1; }
__host__ __device__ constexpr int cnn3_ncols(int r) { return r == 0 ? 6 : (r == 6 ? 8 : 7); }
__host__ __device__ constexpr int cnn3_task_end(int r) {     // tasks in slots <= r (20 rows x 2 channel groups per column)
    return 40 * (cnn3_x0(r) + cnn3_ncols(r));
}

#ifndef NWW_CPUSIM
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Wait that yields the issue slots: the consumer warps of the pipeline spend most of a window waiting for the front end,
// and a tight try_wait loop on 8 warps takes issue cycles from the 7 warps that are the critical path.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, unsigned ns) {
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        __nanosleep(ns);
    }
}
#endif

template <int ACT>
__global__ void __launch_bounds__(Cnn3::NT, 1)
cnn3_stage_kernel(WindowSource src, Cnn2MelSource msrc, long long n_windows, FrontendTables<double> tab, Cnn2Weights wt,
                  float* __restrict__ feat_hi, float* __restrict__ feat_lo /* as in cnn2_stage_kernel */,
                  float* __restrict__ mel_dump /* nullable, (F,T); PCM mode only */) {
    using D = Cnn2;
    using P = Cnn3;
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    unsigned char* a1b = smem + P::oA1;
    cplx<double>* tw = reinterpret_cast<cplx<double>*>(smem + P::oTw);
    double* win_s = reinterpret_cast<double*>(smem + P::oTw + Fe3::kTwBytes);
    float* melp = reinterpret_cast<float*>(smem + P::oMel);
    unsigned char* w2s = smem + P::oW2;
    float* w1s = reinterpret_cast<float*>(smem + P::oSmall);
    float* b1s = w1s + 16 * 9;
    float* b2s = b1s + 16;
    uint64_t* full_mel = reinterpret_cast<uint64_t*>(smem + P::oBars);      // [7] count 7: a front-end warp finished round r
    uint64_t* done = full_mel + P::N_SLOTS;                                 // [7] count 8: a conv warp finished the tasks of slot r
    uint64_t* full_pcm = done + P::N_SLOTS;                                 // [7] TMA transaction barriers
    uint64_t* consumed = full_pcm + P::N_SLOTS;                             // [7] count 7: a front-end warp is past round r
    uint64_t* mma_bar = consumed + P::N_SLOTS;                              // [2] one per M tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 2);
    int16_t* pcm_s = reinterpret_cast<int16_t*>(smem + P::oPcm);

    if (tid == 0) {
        for (int i = 0; i < P::N_SLOTS; ++i) {
            mbar_init(&full_mel[i], P::N_FE_WARPS);
            mbar_init(&done[i], P::N_CONV_WARPS);
            mbar_init(&full_pcm[i], 1);
            mbar_init(&consumed[i], P::N_FE_WARPS);
        }
        mbar_init(&mma_bar[0], 1);
        mbar_init(&mma_bar[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, D::TMEM_COLS);
    fe2_build_tables(tw, tab, tid, P::NT);
    fe3_build_window(win_s, tab.window, tid, P::NT);
    for (int i = tid; i < 16 * 9; i += P::NT) w1s[i] = wt.w1[i];
    for (int i = tid; i < 16; i += P::NT) b1s[i] = wt.b1[i];
    for (int i = tid; i < 32; i += P::NT) b2s[i] = wt.b2[i];
    for (int i = tid; i < D::W2_BYTES / 16; i += P::NT) reinterpret_cast<uint4*>(w2s)[i] = wt.w2_umma[i];
    for (int i = tid; i < P::MEL_ROWS * P::MEL_P; i += P::NT) melp[i] = 0.0f;    // zero border, written once
    // parity planes: conv1 writes every position that holds a real pixel, the others ARE the zero border of conv2
    for (int i = tid; i < D::A1_BYTES / 16; i += P::NT) reinterpret_cast<uint4*>(a1b)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();                                                         // w2s / planes are read by the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const bool from_mel = msrc.ring != nullptr;

    if (warp < P::N_FE_WARPS) {
        // =============================== front end: 7 warps x 7 FFTs per window ===============================
        cplx<double>* wb = reinterpret_cast<cplx<double>*>(smem + P::oWork) + (size_t)warp * Fe3::NPAD;
        int it = 0;
        for (long long w = blockIdx.x; w < n_windows; w += gridDim.x, ++it) {
            const uint32_t par = (uint32_t)(it & 1);
            if (from_mel) {
                const long long s = msrc.count ? msrc.stream(w) : msrc.s0 + w;
                const bool plain = msrc.count == nullptr;
                const int row = plain ? D::TT : SMel::ROW;
                const float* ring = plain ? msrc.ring + s * (long long)(D::F * D::TT)
                                          : msrc.ring + s * SMel::STREAM_FLOATS + smel_slot(msrc.count[s] / SMel::HOP - 3 + 1);
#pragma unroll 1
                for (int r = 0; r < P::N_SLOTS; ++r) {
                    if (it > 0) mbar_wait(&done[r < 6 ? r + 1 : 6], par ^ 1u);
                    for (int e = warp * 32 + lane; e < D::F * P::SLOT_FRAMES; e += P::N_FE_WARPS * 32) {
                        const int m = e / P::SLOT_FRAMES, t = P::SLOT_FRAMES * r + (e - m * P::SLOT_FRAMES);
                        melp[(m + 1) * P::MEL_P + t + 1] = ring[m * row + t];
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_mel[r]);
                }
                continue;
            }
            const int16_t* x = pcm_s + PcmStager<D::G::CLIP>::skew_of(src.at(w));
            float* md = mel_dump ? mel_dump + w * (long long)(D::F * D::TT) : nullptr;
#pragma unroll 1
            for (int r = 0; r < P::N_SLOTS; ++r) {
                const int f = P::N_FE_WARPS * r + warp;                         // packed FFT = frames 2f, 2f + 1
                if (r == 0) mbar_wait(&full_pcm[0], par);
                if (r < 6) mbar_wait(&full_pcm[r + 1], par);                    // the round's last FFTs reach 240 samples into it
                if (it > 0) mbar_wait(&done[r < 6 ? r + 1 : 6], par ^ 1u);      // conv1 of the previous window is past this slot
                // the seven warps start every round together: their instruction streams (64 KB of straight-line FFT code)
                // then stay within a few cache lines of each other instead of spreading over the whole instruction cache
                named_bar_sync(P::BAR_FE, P::N_FE_WARPS * 32);
                float* mf = melp + P::MEL_P + 1 + 2 * f;
                fe3_warp_fft(x + 2 * f * D::G::HOP, wb, win_s, tw, tab,
                             [&](int fr, int m, float db) {
                                 mf[m * P::MEL_P + fr] = db;
                                 if (md) md[m * D::TT + 2 * f + fr] = db;
                             },
                             lane);
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&full_mel[r]);
                    mbar_arrive(&consumed[r]);
                }
            }
        }
    } else if (warp == P::CTRL_WARP) {
        // =============================== control: PCM segments by TMA =====================================
        if (lane == 0 && !from_mel) {
            auto issue_segment = [&](long long w, int s) {
                const int16_t* p = src.at(w);
                const int skew = PcmStager<D::G::CLIP>::skew_of(p);
                const uint32_t bytes = s < 6 ? (uint32_t)(P::SEG * 2) : (uint32_t)((D::G::CLIP - 6 * P::SEG + (skew ? 8 : 0)) * 2);
                fence_proxy_async();
                mbar_expect_tx(&full_pcm[s], bytes);
                bulk_g2s(pcm_s + P::SEG * s, p - skew + P::SEG * s, bytes, &full_pcm[s]);
            };
            long long w = blockIdx.x;
            if (w < n_windows)
                for (int s = 0; s < P::N_SLOTS; ++s) issue_segment(w, s);
            int it = 0;
            for (; w < n_windows; w += gridDim.x, ++it) {
                const long long wn = w + gridDim.x;
                if (wn >= n_windows) break;
                for (int s = 0; s < P::N_SLOTS; ++s) {
                    mbar_wait_backoff(&consumed[s], (uint32_t)(it & 1), 500);
                    issue_segment(wn, s);
                }
            }
        }
    } else {
        // =============================== conv: conv1 -> conv2 (tcgen05) -> epilogue ===========================
        const int cw = warp - P::CONV_WARP0, ctid = tid - P::CONV_WARP0 * 32;
        const uint32_t a1_addr = smem_u32(a1b), w2_addr = smem_u32(w2s);
        constexpr uint32_t kIdesc = umma_idesc_bf16(128, 32);
        const uint64_t da_base = umma_desc_noswz(a1_addr, D::KG_BYTES, 128);
        const uint64_t db_base = umma_desc_noswz(w2_addr, 512, 128);
        // (the tile / quad loops stay rolled: the pipeline's three roles share one instruction cache, and 216 unrolled
        //  MMAs with their descriptor arithmetic are 17 KB of code that one thread walks once per window)
        auto issue_tile = [&](int tile) {                                       // see nww_cnn2.cuh
            const uint64_t da_t = da_base + (uint64_t)(tile * 128);
#pragma unroll 1
            for (int quad = 0; quad < 4; ++quad) {
                const int dy = quad >> 1, dx = quad & 1;
                const uint32_t d_tmem = tmem_base + (uint32_t)((tile * 4 + quad) * 32);
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int ry = dy + r - 1, cx = dx + c - 1;
                        const int plane = ((ry & 1) << 1) | (cx & 1);
                        const int s0 = 26 + 25 * (ry >> 1) + (cx >> 1);       // >> is arithmetic: ry, cx = -1 .. 2
                        const uint64_t da_hi = da_t + (uint64_t)(long long)((plane * 2 * D::PLANE_BYTES) / 16 + s0);
                        const uint64_t da_lo = da_hi + (uint64_t)(D::PLANE_BYTES / 16);
                        const uint64_t db_hi = db_base + (uint64_t)(((r * 3 + c) * 2 * D::W2_TAP_BYTES) / 16);
                        const uint64_t db_lo = db_hi + (uint64_t)(D::W2_TAP_BYTES / 16);
                        umma_bf16(d_tmem, da_hi, db_hi, kIdesc, (r | c) != 0);
                        umma_bf16(d_tmem, da_lo, db_hi, kIdesc, 1);
                        umma_bf16(d_tmem, da_hi, db_lo, kIdesc, 1);
                    }
            }
            umma_commit(&mma_bar[tile]);
        };

        // conv1 + act + 2x2 max pool of 8 channels of one pooled pixel -> bf16 hi / lo rows of the parity planes
        auto conv1_task = [&](int y, int x, int cg) {
            float in[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float2 lo = *reinterpret_cast<const float2*>(melp + (2 * y + r) * P::MEL_P + 2 * x);
                const float2 hi = *reinterpret_cast<const float2*>(melp + (2 * y + r) * P::MEL_P + 2 * x + 2);
                in[r][0] = lo.x; in[r][1] = lo.y; in[r][2] = hi.x; in[r][3] = hi.y;
            }
            float acc[8][4];
            {
                const float4 ba = *reinterpret_cast<const float4*>(b1s + cg * 8);
                const float4 bb = *reinterpret_cast<const float4*>(b1s + cg * 8 + 4);
                const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                for (int o = 0; o < 8; ++o)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[o][q] = bv[o];
            }
            const float* wk = w1s + cg * 72;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float4 wa = *reinterpret_cast<const float4*>(wk + (r * 3 + c) * 8);
                    const float4 wb4 = *reinterpret_cast<const float4*>(wk + (r * 3 + c) * 8 + 4);
                    const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb4.x, wb4.y, wb4.z, wb4.w};
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        acc[o][0] = fmaf(in[r][c], wv[o], acc[o][0]);
                        acc[o][1] = fmaf(in[r][c + 1], wv[o], acc[o][1]);
                        acc[o][2] = fmaf(in[r + 1][c], wv[o], acc[o][2]);
                        acc[o][3] = fmaf(in[r + 1][c + 1], wv[o], acc[o][3]);
                    }
                }
            uint32_t hb[8], lb[8];
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const float best = fmaxf(fmaxf(cnn2_act<ACT>(acc[o][0]), cnn2_act<ACT>(acc[o][1])),
                                         fmaxf(cnn2_act<ACT>(acc[o][2]), cnn2_act<ACT>(acc[o][3])));
                hb[o] = float_to_bf16_bits(best);
                lb[o] = float_to_bf16_bits(best - bf16_bits_to_float(hb[o]));
            }
            const int plane = ((y & 1) << 1) | (x & 1);
            const int s = ((y >> 1) + 1) * D::PITCH + (x >> 1) + 1;
            unsigned char* dst = a1b + plane * 2 * D::PLANE_BYTES + cg * D::KG_BYTES + s * 16;
            *reinterpret_cast<uint4*>(dst) = make_uint4(cnn2_pack_bf16(hb[0], hb[1]), cnn2_pack_bf16(hb[2], hb[3]),
                                                        cnn2_pack_bf16(hb[4], hb[5]), cnn2_pack_bf16(hb[6], hb[7]));
            *reinterpret_cast<uint4*>(dst + D::PLANE_BYTES) = make_uint4(cnn2_pack_bf16(lb[0], lb[1]), cnn2_pack_bf16(lb[2], lb[3]),
                                                                         cnn2_pack_bf16(lb[4], lb[5]), cnn2_pack_bf16(lb[6], lb[7]));
        };

        // conv1 tasks in slot order: task T = ctid + 256 i; inside slot r: (row y, column, channel group) with the column
        // fastest, so a warp reads runs of mel columns.  Decoded on the fly in a ROLLED loop (one copy of the 450-instruction
        // task body: see the instruction-cache note above).
        auto slot_of = [](int T) {
            int r = 0;
#pragma unroll
            for (int q = 0; q < 6; ++q) r += T >= cnn3_task_end(q) ? 1 : 0;
            return r;
        };

        int it = 0;
        for (long long w = blockIdx.x; w < n_windows; w += gridDim.x, ++it) {
            const uint32_t par = (uint32_t)(it & 1);
            int arrived = 0;                                                    // done[r] signalled for r < arrived
#pragma unroll 1
            for (int i = 0; i < P::CONV_ROUNDS; ++i) {
                const int T = ctid + P::CONV_NT * i;
                if (cw * 32 + P::CONV_NT * i < D::CONV1_TASKS) {
                    // the warp waits for the slot of its LAST task of the round (slots complete in order)
                    mbar_wait_backoff(&full_mel[slot_of(min(cw * 32 + 31 + P::CONV_NT * i, D::CONV1_TASKS - 1))], par, 200);
                    if (T < D::CONV1_TASKS) {
                        const int r = slot_of(T);
                        const int local = T - (r ? 40 * cnn3_x0(r) : 0);          // tasks before slot r = 40 per column
                        const int nc2 = r == 0 ? 12 : (r == 6 ? 16 : 14);
                        const int y = r == 0 ? local / 12 : (r == 6 ? local >> 4 : local / 14);
                        const int rem = local - y * nc2;
#if NWW_EXP != 1
                        conv1_task(y, cnn3_x0(r) + (rem >> 1), rem & 1);
#endif
                    }
                }
                // slots below the one of the warp's first task of the next round are finished by this warp
                const int t_nx = cw * 32 + P::CONV_NT * (i + 1);
                const int upto = (i + 1 < P::CONV_ROUNDS && t_nx < D::CONV1_TASKS) ? slot_of(t_nx) : 7;
                __syncwarp();
                if (lane == 0)
                    for (; arrived < upto; ++arrived) mbar_arrive(&done[arrived]);
                arrived = max(arrived, upto);
            }
            // ---- conv2: all planes written -> 2 x 108 MMAs by one thread --------------------------------------
            fence_proxy_async();
            named_bar_sync(P::BAR_CONV, P::CONV_NT);
#if NWW_EXP == 1 || NWW_EXP == 2
            if (ctid == 0) { umma_commit(&mma_bar[0]); umma_commit(&mma_bar[1]); }
#else
            if (ctid == 0) {
                tc_fence_after();
#pragma unroll 1
                for (int tile = 0; tile < 2; ++tile) issue_tile(tile);
            }
#endif
            // ---- epilogue: warp -> (TMEM lane quarter, M tile), both channel halves ------------------------------
            {
                const int q = cw & 3, tile = cw >> 2;
                mbar_wait_backoff(&mma_bar[tile], par, 100);
                tc_fence_after();
                const int m = tile * 128 + q * 32 + lane;
                const int ph = m / D::PITCH, pw = m - ph * D::PITCH;
                const bool valid = m < D::H2 * D::PITCH && pw < D::W2;
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    uint32_t r[4][16];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tile * 4 * 32 + half * 16);
#pragma unroll
                    for (int quad = 0; quad < 4; ++quad) tmem_ld_32x32b_x16_nowait(taddr + quad * 32, r[quad]);
                    tmem_ld_wait();
                    if (valid) {
                        float hi[16], lo[16];
#pragma unroll
                        for (int o = 0; o < 16; ++o) {
                            const float bias = b2s[half * 16 + o];
                            float v = cnn2_act<ACT>(__uint_as_float(r[0][o]) + bias);
                            v = fmaxf(v, cnn2_act<ACT>(__uint_as_float(r[1][o]) + bias));
                            v = fmaxf(v, cnn2_act<ACT>(__uint_as_float(r[2][o]) + bias));
                            v = fmaxf(v, cnn2_act<ACT>(__uint_as_float(r[3][o]) + bias));
                            hi[o] = feat_lo ? round_tf32(v) : v;
                            lo[o] = round_tf32(v - hi[o]);
                        }
                        const long long off = w * (long long)D::FEAT + (ph * D::W2 + pw) * 32 + half * 16;
                        float4* dh = reinterpret_cast<float4*>(feat_hi + off);
#pragma unroll
                        for (int j = 0; j < 4; ++j) dh[j] = make_float4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                        if (feat_lo != nullptr) {
                            float4* dl = reinterpret_cast<float4*>(feat_lo + off);
#pragma unroll
                            for (int j = 0; j < 4; ++j) dl[j] = make_float4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                        }
                    }
                }
                tc_fence_before();
            }
            named_bar_sync(P::BAR_CONV, P::CONV_NT);      // planes and TMEM are free again (both tiles' MMAs were waited for)
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, D::TMEM_COLS);
    }
}

}  // namespace nww
