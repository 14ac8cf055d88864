// nww_cnn4.cuh — stage A of the CNN head as TWO co-resident kernels (round 2, after the cnn3 experiment).
//
// What nww_cnn3.cuh showed: overlapping the FP64 front end with conv1 / conv2 / the epilogue pays only if the front end
// keeps (almost) all of its warps, and one CTA cannot hold ten FFT scratch buffers AND the parity planes (227 KB), nor
// can one kernel give the FFT warps 128 registers and still leave room for eight more warps.  Two kernels can:
//
//   fe4_mel_kernel    10 warp-private FFT warps + 1 control warp (352 threads x 128 registers), 126 KB of shared memory:
//                     the window's PCM comes through a THREE-slot ring of 3200-sample segments filled by TMA (round r of
//                     the FFTs reads segment r and the first 240 samples of segment r + 1; slot 0 is mirrored behind
//                     slot 2 so that the pair is always contiguous), 49 FFTs = 5 rounds of 10 (98 % balanced); the dB
//                     frames go to a (n, F, T) log-mel buffer in global memory (L2-resident: 15.7 KB per window).
//   conv4_mel_kernel  8 warps (256 threads, <= 80 registers), 98 KB: conv1 straight from that buffer (read-only cache, the
//                     zero border by predication), parity planes, 216 tcgen05.mma, TMEM epilogue — the conv role of
//                     nww_cnn3.cuh as its own kernel.
//
// Both fit on one SM together (126 + 98 KB, 256 TMEM columns; registers are allocated in groups of four warps, so the
// front end's 11 warps take 12 x 4096 and the conv kernel must stay at 64 registers to get the remaining 16 384), and the
// engine launches them on two streams — front end of sub-chunk k + 1 beside the convolution of sub-chunk k.  Same
// arithmetic as v2 / v3: bit-identical scores.
//
// MEASURED (tools/probe/split_probe.cu, B200): front end alone 38.8 k cycles per window, conv kernel alone 42 k, both
// co-resident 66.7 k per window — they slow each other (18 latency-bound warps share the issue slots and the
// shared-memory pipe) and the pair loses to v2's 62 k.  Opt-in (`cnn_stage="v4"`), kept as the measured end point of
// the "overlap the phases" line of work; see DESIGN.md §4.1.
#pragma once

#include "nww_cnn3.cuh"

namespace nww {

struct Fe4 {
    using G = GeoNS40x98;
    static constexpr int NW = 10, NT = (NW + 1) * 32;            // FFT warps + the control warp
    static constexpr int N_ROUNDS = 5;                           // 49 FFTs = 4 x 10 + 9
    static constexpr int SEG = NW * 2 * G::HOP;                  // 3200 samples per segment (one round's fresh samples)
    static constexpr int N_SEG = G::CLIP / SEG;                  // 5
    static constexpr int N_SLOT = 3;
    static_assert(N_SEG * SEG == G::CLIP, "segments tile the window");
    static constexpr size_t kWork = (size_t)NW * Fe3::NPAD * sizeof(cplx<double>);            // 83200
    static constexpr size_t oTw = kWork;
    static constexpr size_t oPcm = oTw + Fe3::kTwBytes + Fe3::kWinBytes;
    static constexpr size_t kPcm = (size_t)(N_SLOT + 1) * SEG * sizeof(int16_t);                  // + the mirror of slot 0
    static constexpr size_t oBars = oPcm + kPcm;
    static constexpr size_t kTotal = oBars + 128;
};

__global__ void __maxnreg__(128)      // 352 threads x 128 registers: leaves 20 480 registers for the conv kernel's CTA
fe4_mel_kernel(const int16_t* __restrict__ pcm /* [n][16000], 16-byte aligned */, long long n_windows, FrontendTables<double> tab,
               float* __restrict__ mel_out /* [n][F][T] */) {
    using P = Fe4;
    using G = GeoNS40x98;
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    cplx<double>* tw = reinterpret_cast<cplx<double>*>(smem + P::oTw);
    double* win_s = reinterpret_cast<double*>(smem + P::oTw + Fe3::kTwBytes);
    int16_t* ring = reinterpret_cast<int16_t*>(smem + P::oPcm);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + P::oBars);        // [3] TMA transaction barriers
    uint64_t* consumed = full + P::N_SLOT;                                // [3] count NW: a front-end warp is past the segment
    if (tid == 0) {
        for (int i = 0; i < P::N_SLOT; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&consumed[i], P::NW);
        }
        fence_mbar_init();
    }
    fe2_build_tables(tw, tab, tid, P::NT);
    fe3_build_window(win_s, tab.window, tid, P::NT);
    __syncthreads();

    long long n_mine = 0;                                                 // windows of this CTA
    if ((long long)blockIdx.x < n_windows) n_mine = (n_windows - 1 - blockIdx.x) / gridDim.x + 1;
    const long long n_segs = n_mine * P::N_SEG;                           // global segment g = 5 * it + s goes to slot g % 3

    if (warp < P::NW) {
        cplx<double>* wb = reinterpret_cast<cplx<double>*>(smem) + (size_t)warp * Fe3::NPAD;
        long long g = 0;
        for (long long it = 0; it < n_mine; ++it) {
            float* mw = mel_out + (blockIdx.x + it * gridDim.x) * (long long)(G::N_MELS * G::N_FRAMES);
#pragma unroll 1
            for (int r = 0; r < P::N_ROUNDS; ++r, ++g) {
                const int slot = (int)(g % P::N_SLOT);
                mbar_wait(&full[slot], (uint32_t)((g / P::N_SLOT) & 1));
                if (r + 1 < P::N_ROUNDS)                                  // the round's last FFTs reach 240 samples into the next segment
                    mbar_wait(&full[(slot + 1) % P::N_SLOT], (uint32_t)(((g + 1) / P::N_SLOT) & 1));
                const int f = P::NW * r + warp;                           // packed FFT = frames 2f, 2f + 1
                if (f < Fe2::N_FFT_TOTAL) {
                    float* mf = mw + 2 * f;
                    fe3_warp_fft(ring + slot * P::SEG + 2 * G::HOP * warp, wb, win_s, tw, tab,
                                 [&](int fr, int m, float db) { mf[m * G::N_FRAMES + fr] = db; }, lane);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&consumed[slot]);
            }
        }
    } else if (lane == 0) {
        // control: segment g of this CTA's window sequence -> slot g % 3 (and its mirror behind slot 2 when that is slot 0)
        for (long long g = 0; g < n_segs; ++g) {
            const int slot = (int)(g % P::N_SLOT);
            if (g >= P::N_SLOT) mbar_wait_backoff(&consumed[slot], (uint32_t)(((g / P::N_SLOT) - 1) & 1), 200);
            const long long it = g / P::N_SEG;
            const int s = (int)(g - it * P::N_SEG);
            const int16_t* src = pcm + (blockIdx.x + it * gridDim.x) * (long long)G::CLIP + (long long)s * P::SEG;
            const uint32_t bytes = (uint32_t)(P::SEG * sizeof(int16_t));
            fence_proxy_async();
            mbar_expect_tx(&full[slot], slot == 0 ? 2 * bytes : bytes);
            bulk_g2s(ring + slot * P::SEG, src, bytes, &full[slot]);
            if (slot == 0) bulk_g2s(ring + P::N_SLOT * P::SEG, src, bytes, &full[slot]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
struct Conv4 {
    using D = Cnn2;
    static constexpr int NT = 256, NWARP = 8;
    static constexpr int ROUNDS = (D::CONV1_TASKS + NT - 1) / NT;                                    // 8
    static constexpr size_t oA1 = 0;
    static constexpr size_t oW2 = align_up(D::A1_BYTES, 128);
    static constexpr size_t oSmall = oW2 + D::W2_BYTES;
    static constexpr size_t oBars = oSmall + D::kSmall;
    static constexpr size_t kTotal = oBars + 128;
};

template <int ACT>
__global__ void __maxnreg__(64)      // 256 threads; the kernel shares the SM (and its register file) with fe4_mel_kernel
conv4_mel_kernel(const float* __restrict__ mel /* [n][F][T] */, long long n_windows, Cnn2Weights wt, float* __restrict__ feat_hi,
                 float* __restrict__ feat_lo /* as in cnn2_stage_kernel */) {
    using D = Cnn2;
    using P = Conv4;
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* a1b = smem + P::oA1;
    unsigned char* w2s = smem + P::oW2;
    float* w1s = reinterpret_cast<float*>(smem + P::oSmall);
    float* b1s = w1s + 16 * 9;
    float* b2s = b1s + 16;
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(smem + P::oBars);            // [2], one per M tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 2);
    if (tid == 0) {
        mbar_init(&mma_bar[0], 1);
        mbar_init(&mma_bar[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, D::TMEM_COLS);
    for (int i = tid; i < 16 * 9; i += P::NT) w1s[i] = wt.w1[i];
    for (int i = tid; i < 16; i += P::NT) b1s[i] = wt.b1[i];
    for (int i = tid; i < 32; i += P::NT) b2s[i] = wt.b2[i];
    for (int i = tid; i < D::W2_BYTES / 16; i += P::NT) reinterpret_cast<uint4*>(w2s)[i] = wt.w2_umma[i];
    for (int i = tid; i < D::A1_BYTES / 16; i += P::NT) reinterpret_cast<uint4*>(a1b)[i] = make_uint4(0, 0, 0, 0);   // zero border, once
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a1_addr = smem_u32(a1b), w2_addr = smem_u32(w2s);
    constexpr uint32_t kIdesc = umma_idesc_bf16(128, 32);
    const uint64_t da_base = umma_desc_noswz(a1_addr, D::KG_BYTES, 128);
    const uint64_t db_base = umma_desc_noswz(w2_addr, 512, 128);

    auto issue_tile = [&](int tile) {                                           // see nww_cnn2.cuh / nww_cnn3.cuh
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
                    const int s0 = 26 + 25 * (ry >> 1) + (cx >> 1);
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

    int it = 0;
    for (long long w = blockIdx.x; w < n_windows; w += gridDim.x, ++it) {
        const float* mw = mel + w * (long long)(D::F * D::TT);
        // ---- conv1 + act + 2x2 max pool: task = (pooled pixel, 8 output channels); the 4 x 4 input patch comes straight
        // from the log-mel buffer (rows 2y - 1 .. 2y + 2, columns 2x - 1 .. 2x + 2; outside = the zero padding) ------------
#pragma unroll 1
        for (int rd = 0; rd < P::ROUNDS; ++rd) {
            const int T = rd * P::NT + tid;
            if (T >= D::CONV1_TASKS) break;
            const int pos = T >> 1, cg = T & 1;
            const int y = pos / D::W1, x = pos - y * D::W1;
            float in[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int my = 2 * y + r - 1;
                const bool row_ok = my >= 0 && my < D::F;
                const float* row = mw + my * D::TT;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int mx = 2 * x + c - 1;
                    in[r][c] = (row_ok && mx >= 0 && mx < D::TT) ? __ldg(row + mx) : 0.0f;
                }
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
        }
        // ---- conv2: 2 x 108 MMAs by one thread ---------------------------------------------------------------------------
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll 1
            for (int tile = 0; tile < 2; ++tile) issue_tile(tile);
        }
        // ---- epilogue: warp -> (TMEM lane quarter, M tile), both channel halves ------------------------------------------
        {
            const int q = warp & 3, tile = warp >> 2;
            mbar_wait_backoff(&mma_bar[tile], (uint32_t)(it & 1), 100);
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
                    const long long off = w * (long long)D::FEAT + (ph * D::W2 + pw) * 32 + half * 16;
                    float4* dh = reinterpret_cast<float4*>(feat_hi + off);
                    float4* dl = feat_lo ? reinterpret_cast<float4*>(feat_lo + off) : nullptr;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float hi[4], lo[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int o = 4 * j + e;
                            const float bias = b2s[half * 16 + o];
                            float v = cnn2_act<ACT>(__uint_as_float(r[0][o]) + bias);
                            v = fmaxf(v, cnn2_act<ACT>(__uint_as_float(r[1][o]) + bias));
                            v = fmaxf(v, cnn2_act<ACT>(__uint_as_float(r[2][o]) + bias));
                            v = fmaxf(v, cnn2_act<ACT>(__uint_as_float(r[3][o]) + bias));
                            hi[e] = feat_lo ? round_tf32(v) : v;
                            lo[e] = round_tf32(v - hi[e]);
                        }
                        dh[j] = make_float4(hi[0], hi[1], hi[2], hi[3]);
                        if (dl) dl[j] = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
            }
            tc_fence_before();
        }
        __syncthreads();          // planes and TMEM are free again
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, D::TMEM_COLS);
    }
}

}  // namespace nww
