// nww_cnn2.cuh — stage A of the CNN head, v2: PCM -> log-mel -> conv1+act+pool -> conv2+act+pool,
// one window per CTA iteration, persistent one CTA per SM; conv2 (81 % of the head's MACs) runs on
// the 5th-generation tensor cores as an implicit GEMM with no im2col.
//
// Reference: CNNModel, nanowakeword/modules/architectures.py:51-80
//   conv1 = Conv2d(1, 16, 3, pad 1) -> act -> MaxPool2d(2)      (40,98) -> (16,20,49)
//   conv2 = Conv2d(16, 32, 3, pad 1) -> act -> MaxPool2d(2)     -> (32,10,24)   [49 -> 24: floor]
//   flatten (c, h, w) row-major -> 7680 -> fc1 (nww_gemm_tc.cuh)
//
// conv2 as tcgen05 tiles.  Let a1[ic][y][x] be the pooled conv1 output (y < 20, x < 49) with a
// zero border.  The conv2 output needed by pooled position (ph, pw) and pooling quad (dy, dx) is
//   sum_{r,c,ic} w2[oc][ic][r][c] * a1[ic][2 ph + dy + r - 1][2 pw + dx + c - 1].
// a1 is stored as four PARITY PLANES (y & 1, x & 1), each a flat list of positions
//   s = ((y >> 1) + 1) * 25 + (x >> 1) + 1          (pitch 25: column 25 of a row is column 0 of the next)
// with the 16 input channels of a position as two 16-byte rows (8 bf16 each) in two K groups.  With
// m = ph * 25 + pw as the GEMM row, the operand of (quad, tap) is the plane
// ((dy + r - 1) & 1, (dx + c - 1) & 1) starting at position  m + 26 + 25 * ((dy + r - 1) >> 1) +
// ((dx + c - 1) >> 1): a K-major, un-swizzled UMMA operand whose core matrices are 8 consecutive
// positions (128 contiguous bytes, SBO = 128) and whose second K half is the other K group
// (LBO = plane stride) — every tap is the same buffer behind a different descriptor start address.
// The four quads accumulate into four TMEM column groups with identical row <-> (ph, pw) mapping, so
// the 2x2 max pool is an element-wise max of four tcgen05.ld results in registers.
// M = 250 rows = 2 tiles of 128; per tile 4 quads x 9 taps x 3 bf16 split products
// (a_hi w_hi + a_lo w_hi + a_hi w_lo, ~2^-16 relative) of 128 x 32 x 16.
//
// The epilogue writes the 7680 features directly as the TF32 hi / lo pair the fc1 tensor-core GEMM
// consumes, in the K order (ph, pw, oc) (fc1's weight columns are permuted to match at create time).
#pragma once

#include "nww_fe3.cuh"
#include "nww_stream_mel.cuh"

namespace nww {

struct Cnn2 {
    using G = GeoNS40x98;
    static constexpr int NT = 512;
    static constexpr int F = 40, TT = 98, C1 = 16, H1 = 20, W1 = 49, C2 = 32, H2 = 10, W2 = 24;
    static constexpr int FEAT = C2 * H2 * W2;                     // 7680
    static constexpr int MEL_P = 100, MEL_ROWS = 42;              // zero-bordered log-mel plane
    static constexpr int PITCH = 25;                              // GEMM row m = ph * 25 + pw
    static constexpr int NPOS = 308;                              // positions per parity plane (max read 307)
    static constexpr int KG_BYTES = NPOS * 16;                    // one K group of one plane (LBO)
    static constexpr int PLANE_BYTES = 2 * KG_BYTES;              // [kg][pos] x 16 B
    static constexpr int A1_BYTES = 4 * 2 * PLANE_BYTES;          // [plane][hi|lo] = 78848
    static constexpr int W2_TAP_BYTES = 1024;                     // [kg 2][oc 32] x 16 B per (tap, hi|lo)
    static constexpr int W2_BYTES = 9 * 2 * W2_TAP_BYTES;         // 18432
    static constexpr int CONV1_TASKS = H1 * W1 * 2;               // (position, channel group of 8) = 1960
    static constexpr int CONV1_SPLIT_ROUNDS = 3;                  // after 3 rounds of 512 tasks a1 rows 0..14 are complete
    static constexpr int TMEM_COLS = 256;                         // 2 tiles x 4 quads x 32 columns
    static constexpr int ZSLOTS = 62;                             // border positions per (plane, hi|lo, K group) copy

    static constexpr size_t kUnion = align_up(Fe3::kWorkBytes > (size_t)A1_BYTES ? Fe3::kWorkBytes : (size_t)A1_BYTES, 1024);
    static constexpr size_t kTw = Fe3::kTwBytes + Fe3::kWinBytes;     // twiddles, then the Hann table
    static constexpr size_t kMel = align_up(sizeof(float) * MEL_ROWS * MEL_P, 128);
    static constexpr size_t kW2 = W2_BYTES;
    static constexpr size_t kSmall = align_up(sizeof(float) * (16 * 9 + 16 + 32), 128);
    static constexpr size_t kBars = 128;                          // 2 mbarriers + TMEM base slot
    static constexpr size_t oTw = kUnion, oMel = oTw + kTw, oW2 = oMel + kMel, oSmall = oW2 + kW2, oBars = oSmall + kSmall,
                            oPcm = oBars + kBars;
    using Stager = PcmStager<G::CLIP, 1>;                          // one PCM slot: the next window is fetched as soon as
                                                                  // the FFT phase has consumed this one
    static constexpr size_t kTotal = oPcm + Stager::kBytes;
};

struct Cnn2Weights {
    const float* w1;          // [cg 2][tap 9][8 oc]: conv1 weights, 8 output channels of a tap contiguous
    const float* b1;          // [16]
    const uint4* w2_umma;     // Cnn2::W2_BYTES: [tap][hi|lo][kg][oc][8 ic] bf16, built by the engine
    const float* b2;          // [32]
};

// Stream mode: the log-mel of window w is read from the streams' mel ring (nww_stream_mel.cuh)
// instead of being computed from PCM.  ring == nullptr selects the PCM path.
using Cnn2MelSource = MelRingRef;

__device__ __forceinline__ uint32_t cnn2_pack_bf16(uint32_t lo16, uint32_t hi16) { return lo16 | (hi16 << 16); }

template <int ACT> __device__ __forceinline__ float cnn2_act(float x) {
    if (ACT == ACT_RELU) return fmaxf(x, 0.0f);
    return apply_act(x, ACT);
}

template <int ACT>
__global__ void __launch_bounds__(Cnn2::NT, 1)
cnn2_stage_kernel(WindowSource src, Cnn2MelSource msrc, long long n_windows, FrontendTables<double> tab, Cnn2Weights wt,
                  float* __restrict__ feat_hi, float* __restrict__ feat_lo /* [n][7680], K order (ph, pw, oc); feat_lo ==
                  nullptr: plain FP32 rows into feat_hi (channel-last input of the CRNN's third conv) */,
                  float* __restrict__ mel_dump /* nullable, (F,T) */) {
    using D = Cnn2;
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    unsigned char* a1b = smem;                                                   // overlays the FFT scratch
    cplx<double>* tw = reinterpret_cast<cplx<double>*>(smem + D::oTw);
    double* win_s = reinterpret_cast<double*>(smem + D::oTw + Fe3::kTwBytes);
    float* melp = reinterpret_cast<float*>(smem + D::oMel);
    unsigned char* w2s = smem + D::oW2;
    float* w1s = reinterpret_cast<float*>(smem + D::oSmall);
    float* b1s = w1s + 16 * 9;
    float* b2s = b1s + 16;
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(smem + D::oBars);            // [2], one per M tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + D::oBars + 64);
    typename D::Stager stager;
    stager.carve(smem + D::oPcm);
    stager.init(tid);

    if (tid == 0) {
        mbar_init(&mma_bar[0], 1);
        mbar_init(&mma_bar[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, D::TMEM_COLS);
    fe2_build_tables(tw, tab, tid, D::NT);
    fe3_build_window(win_s, tab.window, tid, D::NT);
    for (int i = tid; i < 16 * 9; i += D::NT) w1s[i] = wt.w1[i];
    for (int i = tid; i < 16; i += D::NT) b1s[i] = wt.b1[i];
    for (int i = tid; i < 32; i += D::NT) b2s[i] = wt.b2[i];
    for (int i = tid; i < D::W2_BYTES / 16; i += D::NT) reinterpret_cast<uint4*>(w2s)[i] = wt.w2_umma[i];
    for (int i = tid; i < D::MEL_ROWS * D::MEL_P; i += D::NT) melp[i] = 0.0f;    // zero border, written once
    fence_proxy_async();                                                         // w2s is read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a1_addr = smem_u32(a1b), w2_addr = smem_u32(w2s);
    constexpr uint32_t kIdesc = umma_idesc_bf16(128, 32);

    // One elected thread issues the 108 MMAs of M tile `tile` (4 quads x 9 taps x 3 split products).  Every
    // descriptor is a kernel-constant base plus a compile-time multiple of 16 bytes in the start-address field,
    // so the issue loop is one 64-bit add per operand.
    const uint64_t da_base = umma_desc_noswz(a1_addr, D::KG_BYTES, 128);
    const uint64_t db_base = umma_desc_noswz(w2_addr, 512, 128);
    auto issue_tile = [&](int tile) {
        tc_fence_after();
        const uint64_t da_t = da_base + (uint64_t)(tile * 128);               // 128 positions x 16 B, in 16-byte units
#pragma unroll
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
                    const uint64_t da_hi = da_t + (uint64_t)((plane * 2 * D::PLANE_BYTES) / 16 + s0);
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

    // conv1 + act + 2x2 max pool for task T = (position, channel group): 8 channels of one pooled pixel
    // -> bf16 hi / lo rows of the parity planes.  Weights come as two 128-bit shared loads per tap.
    auto conv1_task = [&](int T) {
        const int pos = T >> 1, cg = T & 1;
        const int y = pos / D::W1, x = pos - y * D::W1;
        float in[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float2 lo = *reinterpret_cast<const float2*>(melp + (2 * y + r) * D::MEL_P + 2 * x);
            const float2 hi = *reinterpret_cast<const float2*>(melp + (2 * y + r) * D::MEL_P + 2 * x + 2);
            in[r][0] = lo.x; in[r][1] = lo.y; in[r][2] = hi.x; in[r][3] = hi.y;
        }
        // packed FP32 FMAs (FFMA2: two IEEE FMAs per instruction — channels o, o + 1 of one conv output): the 288 FMAs of
        // a task take 144 issue slots, and this phase is issue-bound; bit-identical to the scalar form
        float2 acc[4][4];                                              // [channel pair][pooling quad]
#pragma unroll
        for (int o2 = 0; o2 < 4; ++o2) {
            const float2 bv = *reinterpret_cast<const float2*>(b1s + cg * 8 + 2 * o2);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[o2][q] = bv;
        }
        const float* wk = w1s + cg * 72;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 wa = *reinterpret_cast<const float4*>(wk + (r * 3 + c) * 8);
                const float4 wb = *reinterpret_cast<const float4*>(wk + (r * 3 + c) * 8 + 4);
                const float2 wv[4] = {make_float2(wa.x, wa.y), make_float2(wa.z, wa.w), make_float2(wb.x, wb.y), make_float2(wb.z, wb.w)};
                const float2 p00 = make_float2(in[r][c], in[r][c]), p01 = make_float2(in[r][c + 1], in[r][c + 1]);
                const float2 p10 = make_float2(in[r + 1][c], in[r + 1][c]), p11 = make_float2(in[r + 1][c + 1], in[r + 1][c + 1]);
#pragma unroll
                for (int o2 = 0; o2 < 4; ++o2) {
                    acc[o2][0] = __ffma2_rn(p00, wv[o2], acc[o2][0]);
                    acc[o2][1] = __ffma2_rn(p01, wv[o2], acc[o2][1]);
                    acc[o2][2] = __ffma2_rn(p10, wv[o2], acc[o2][2]);
                    acc[o2][3] = __ffma2_rn(p11, wv[o2], acc[o2][3]);
                }
            }
        uint32_t hb[8], lb[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const float a0 = (o & 1) ? acc[o >> 1][0].y : acc[o >> 1][0].x, a1 = (o & 1) ? acc[o >> 1][1].y : acc[o >> 1][1].x;
            const float a2 = (o & 1) ? acc[o >> 1][2].y : acc[o >> 1][2].x, a3 = (o & 1) ? acc[o >> 1][3].y : acc[o >> 1][3].x;
            const float best = fmaxf(fmaxf(cnn2_act<ACT>(a0), cnn2_act<ACT>(a1)), fmaxf(cnn2_act<ACT>(a2), cnn2_act<ACT>(a3)));
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

    const bool from_mel = msrc.ring != nullptr;
    long long w = blockIdx.x;
    if (!from_mel && w < n_windows) stager.issue(0, src.at(w), tid);
    for (int it = 0; w < n_windows; w += gridDim.x, ++it) {
        if (from_mel) {
            // ---- stream mode: the window's 98 frames are contiguous per mel row in the mirrored ring ---------
            // (count == nullptr: a plain (n, F, T) log-mel buffer — float feeds, nww_run_windows_f32)
            const long long s = msrc.count ? msrc.stream(w) : msrc.s0 + w;
            const bool plain = msrc.count == nullptr;
            const int row = plain ? D::TT : SMel::ROW;
            const float* ring = plain ? msrc.ring + s * (long long)(D::F * D::TT)
                                      : msrc.ring + s * SMel::STREAM_FLOATS + smel_slot(msrc.count[s] / SMel::HOP - 3 + 1);
            for (int i = tid; i < D::F * D::TT; i += D::NT) {
                const int m = i / D::TT, t = i - m * D::TT;
                melp[(m + 1) * D::MEL_P + t + 1] = ring[m * row + t];
            }
            __syncthreads();
        } else {
            const int16_t* x = stager.wait(0, it & 1, src.at(w));
            // ---- log-mel into the zero-bordered (F+2, T+2) plane (ends with a CTA barrier) ----------------
            fe3_logmel_window(x, smem, win_s, tw, tab, melp + D::MEL_P + 1, D::MEL_P, 1, tid);
            // the PCM slot is free: fetch the next window behind conv1 / conv2 / the epilogue
            const long long wn = w + gridDim.x;
            if (wn < n_windows) stager.issue(0, src.at(wn), tid);
        }
        if (mel_dump != nullptr) {
            float* md = mel_dump + w * (long long)(D::F * D::TT);
            for (int i = tid; i < D::F * D::TT; i += D::NT) md[i] = melp[(i / D::TT + 1) * D::MEL_P + (i % D::TT) + 1];
        }

        // ---- the FFT scratch becomes the parity planes.  conv1 writes every position that holds a real
        // pixel; the positions that stand for the zero border are cleared here (disjoint from conv1's, so no
        // barrier in between): position rows 0 and 11, and in the odd-x planes the column shared by x = -1
        // and x = 49 (s = 25 r).
        for (int i = tid; i < 16 * D::ZSLOTS; i += D::NT) {
            const int copy = i / D::ZSLOTS, z = i - copy * D::ZSLOTS;      // copy = (plane, hi|lo, kg)
            const int plane = copy >> 2;
            int s;
            if (z < 26) s = z;
            else if (z < 52) s = 250 + z;                                   // 276 .. 301
            else if (plane & 1) s = 25 * (z - 50);                          // 50, 75, .., 275
            else continue;
            *reinterpret_cast<uint4*>(a1b + copy * D::KG_BYTES + s * 16) = make_uint4(0, 0, 0, 0);
        }

        // ---- conv1, first three rounds; then M tile 0 of conv2 can start on the tensor core ----------------
#pragma unroll 1
        for (int rd = 0; rd < D::CONV1_SPLIT_ROUNDS; ++rd) conv1_task(rd * D::NT + tid);
        fence_proxy_async();
        __syncthreads();
        if (tid == D::NT - 32) issue_tile(0);                        // warp 15 has no conv1 work in the last round
        if (D::CONV1_SPLIT_ROUNDS * D::NT + tid < D::CONV1_TASKS) conv1_task(D::CONV1_SPLIT_ROUNDS * D::NT + tid);
        fence_proxy_async();
        __syncthreads();
        if (tid == D::NT - 32) issue_tile(1);

        // ---- epilogue: warp -> (TMEM lane quarter, M tile, channel half) ------------------------------------
        {
            const int q = warp & 3, tile = (warp >> 2) & 1, half = warp >> 3;
            mbar_wait(&mma_bar[tile], it & 1);
            tc_fence_after();
            uint32_t r[4][16];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tile * 4 * 32 + half * 16);
#pragma unroll
            for (int quad = 0; quad < 4; ++quad) tmem_ld_32x32b_x16_nowait(taddr + quad * 32, r[quad]);
            tmem_ld_wait();
            tc_fence_before();
            const int m = tile * 128 + q * 32 + lane;
            const int ph = m / D::PITCH, pw = m - ph * D::PITCH;
            if (m < D::H2 * D::PITCH && pw < D::W2) {
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
                float4* dl = reinterpret_cast<float4*>(feat_lo + off);
#pragma unroll
                for (int j = 0; j < 4; ++j) dh[j] = make_float4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                if (feat_lo != nullptr) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) dl[j] = make_float4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                }
            }
        }
        __syncthreads();      // planes (= FFT scratch) and TMEM are free again
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, D::TMEM_COLS);
    }
}

}  // namespace nww
