// nww_bc.cuh — BcResNet head on channel-last activations, with the pointwise + shortcut 1x1
// convolutions (72 % of the head's time as scalar layer kernels) as shared-memory row GEMMs.
//
// Reference: BcResNetModel / BcResNetBlock, nanowakeword/modules/architectures.py:620-687
//   init: conv3x3(1 -> 32, no bias) + BN + act + MaxPool2d(2)                (40,98) -> (32,20,49)
//   block(Cin -> Cout, stride s): o = act(BN(pointwise(depthwise3x3_s(x))))  (activation BEFORE the add, :646-647)
//                                 out = o + BN(shortcut1x1_s(x))
//   blocks (32->64, s(2,2)), (64->128, s(2,2)), (128->256, s(2,1)); global average pool; fc.
// BatchNorm is folded by the packer (weights.py).
//
// Layout: activations are [window][pixel][channel] ("NHWC"), so
//   * a pixel's channels are one contiguous GEMM row: pointwise and shortcut are plain
//     [pixels x Cin] x [Cin x Cout] products, done by the register-tiled row GEMM of nww_tcn.cuh with the
//     weights streamed through shared memory (cp.async) — FP32, exact;
//   * the depthwise kernel reads / writes 128-bit channel quads and also emits the centre tap of its
//     window, which IS the strided 1x1 shortcut's input pixel (stride s, pad 1: centre = (s y, s x)).
#pragma once

#include <string.h>
#include <vector>

#include "nww_tc.cuh"
#include "nww_tcn.cuh"
#ifndef NWW_CPUSIM
#include "nww_fe3.cuh"
#endif

namespace nww {

// init conv (Cin = 1) + folded BN + act + 2x2 max pool.  mel (F, T) per window -> out [n][(F/2)*(T/2)][C0].
// w [9][C0] (tap-major), one thread = one pooled pixel x 8 output channels.
__global__ void __launch_bounds__(256)
bc_init_conv_kernel(const float* __restrict__ mel, const float* __restrict__ w, const float* __restrict__ bias,
                    float* __restrict__ out, long long n, int F, int T, int C0, int act) {
    const int H1 = F / 2, W1 = T / 2, groups = C0 / 8;
    const long long total = n * H1 * W1 * groups;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(t % groups);
        const long long pix = t / groups;
        const int x = (int)(pix % W1), y = (int)((pix / W1) % H1);
        const long long b = pix / ((long long)W1 * H1);
        const float* m = mel + b * (long long)F * T;
        float in[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int yy = 2 * y - 1 + r, xx = 2 * x - 1 + c;
                in[r][c] = (yy >= 0 && yy < F && xx >= 0 && xx < T) ? __ldg(m + yy * T + xx) : 0.0f;
            }
        float acc[8][4];
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const float bv = __ldg(bias + g * 8 + o);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[o][q] = bv;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 wa = __ldg(reinterpret_cast<const float4*>(w + (r * 3 + c) * C0 + g * 8));
                const float4 wb = __ldg(reinterpret_cast<const float4*>(w + (r * 3 + c) * C0 + g * 8) + 1);
                const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    acc[o][0] = fmaf(in[r][c], wv[o], acc[o][0]);
                    acc[o][1] = fmaf(in[r][c + 1], wv[o], acc[o][1]);
                    acc[o][2] = fmaf(in[r + 1][c], wv[o], acc[o][2]);
                    acc[o][3] = fmaf(in[r + 1][c + 1], wv[o], acc[o][3]);
                }
            }
        float v[8];
#pragma unroll
        for (int o = 0; o < 8; ++o)
            v[o] = fmaxf(fmaxf(apply_act(acc[o][0], act), apply_act(acc[o][1], act)),
                         fmaxf(apply_act(acc[o][2], act), apply_act(acc[o][3], act)));
        float4* dst = reinterpret_cast<float4*>(out + pix * C0 + g * 8);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
}

// ---------------------------------------------------------------------------------------
// Stage kernel of the BcResNet head (batch mode, int16 PCM): PCM -> log-mel (the warp-private FP64 front end of
// nww_fe3.cuh, into a zero-bordered shared-memory plane) -> init conv + folded BN + act + 2x2 max pool -> channel-last
// activations [n][20 * 49][32], one window per CTA iteration.  The (40, 98) log-mel never goes to HBM and back, and the
// front end's shared-memory-bound phase and the conv's FMA-bound phase run in one launch.  Same arithmetic, FMA for
// FMA, as frontend3_kernel + bc_init_conv_kernel (which the stream / float-feed paths keep using): bit-identical.
// Shared memory as in cnn2_stage_kernel: FFT scratch | twiddles + Hann | mel plane | weights | one PCM slot (the next
// window's PCM is fetched by TMA as soon as the FFT phase has consumed this one).
// ---------------------------------------------------------------------------------------
#ifndef NWW_CPUSIM
struct BcStage {
    using G = GeoNS40x98;
    static constexpr int NT = 512, F = 40, TT = 98, C0 = 32, H1 = 20, W1 = 49;
    static constexpr int MEL_P = 100, MEL_ROWS = 42;
    static constexpr size_t oTw = align_up(Fe3::kWorkBytes, 1024);
    static constexpr size_t oMel = oTw + Fe3::kTwBytes + Fe3::kWinBytes;
    static constexpr size_t oW = oMel + align_up(sizeof(float) * MEL_ROWS * MEL_P, 128);
    static constexpr size_t oPcm = oW + align_up(sizeof(float) * (9 * C0 + C0), 128);
    using Stager = PcmStager<G::CLIP, 1>;
    static constexpr size_t kTotal = oPcm + Stager::kBytes;
};

// The conv phase of bc_stage_kernel as a function of its own (not inlined: inside the kernel the register allocator kept
// the front end's loop-invariant values live through this loop and spilled ~20 values per task — 15 % of the kernel's
// stall samples, profiles/r02_bc_stage_*; here the loop has the whole register file to itself).
template <int ACT>
__device__ __noinline__ void bc_stage_conv(const float* __restrict__ melp, const float* __restrict__ ws, const float* __restrict__ bs,
                                           float* __restrict__ ow, int tid) {
    using D = BcStage;
    // task = (pooled pixel, 8 channels): the 4 x 4 log-mel patch, 9 taps x 8 channels x 4 conv outputs, act, max
#pragma unroll 1
    for (int T = tid; T < D::H1 * D::W1 * (D::C0 / 8); T += D::NT) {
        const int pix = T >> 2, g = T & 3;
        // (g is the same for every task of a thread, and left alone the compiler keeps all 72 weights of the group in registers
        // across the loop — and then spills the accumulators around every task, reloads that miss the small L1 this kernel
        // leaves.  Laundering the pointer keeps the weight loads, 18 cheap shared-memory broadcasts per task, inside the loop.)
        const float* wsl = ws;
        asm volatile("" : "+l"(wsl));
        const int y = pix / D::W1, xx = pix - y * D::W1;
        float in[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float2 lo = *reinterpret_cast<const float2*>(melp + (2 * y + r) * D::MEL_P + 2 * xx);
            const float2 hi = *reinterpret_cast<const float2*>(melp + (2 * y + r) * D::MEL_P + 2 * xx + 2);
            in[r][0] = lo.x; in[r][1] = lo.y; in[r][2] = hi.x; in[r][3] = hi.y;
        }
        // packed FP32 FMAs (FFMA2: two IEEE FMAs per instruction — channels o, o + 1 of one conv output — so the 288
        // FMAs of a task cost 144 issue slots; results are bit-identical to the scalar form)
        float2 acc[4][4];                                          // [channel pair][pooling quad]
#pragma unroll
        for (int o2 = 0; o2 < 4; ++o2) {
            const float2 bv = *reinterpret_cast<const float2*>(bs + g * 8 + 2 * o2);
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[o2][q] = bv;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 wa = *reinterpret_cast<const float4*>(wsl + (r * 3 + c) * D::C0 + g * 8);
                const float4 wb = *reinterpret_cast<const float4*>(wsl + (r * 3 + c) * D::C0 + g * 8 + 4);
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
        float v[8];
#pragma unroll
        for (int o2 = 0; o2 < 4; ++o2) {
            v[2 * o2] = fmaxf(fmaxf(apply_act(acc[o2][0].x, ACT), apply_act(acc[o2][1].x, ACT)),
                              fmaxf(apply_act(acc[o2][2].x, ACT), apply_act(acc[o2][3].x, ACT)));
            v[2 * o2 + 1] = fmaxf(fmaxf(apply_act(acc[o2][0].y, ACT), apply_act(acc[o2][1].y, ACT)),
                                  fmaxf(apply_act(acc[o2][2].y, ACT), apply_act(acc[o2][3].y, ACT)));
        }
        float4* dst = reinterpret_cast<float4*>(ow + (size_t)pix * D::C0 + g * 8);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
}

template <int ACT>
__global__ void __launch_bounds__(BcStage::NT, 1)
bc_stage_kernel(WindowSource src, long long n_windows, FrontendTables<double> tab, const float* __restrict__ w /* [9][32] */,
                const float* __restrict__ bias, float* __restrict__ out, float* __restrict__ mel_dump /* nullable, (F, T) */) {
    using D = BcStage;
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x;
    cplx<double>* tw = reinterpret_cast<cplx<double>*>(smem + D::oTw);
    double* win_s = reinterpret_cast<double*>(smem + D::oTw + Fe3::kTwBytes);
    float* melp = reinterpret_cast<float*>(smem + D::oMel);
    float* ws = reinterpret_cast<float*>(smem + D::oW);
    float* bs = ws + 9 * D::C0;
    typename D::Stager stager;
    stager.carve(smem + D::oPcm);
    stager.init(tid);
    fe2_build_tables(tw, tab, tid, D::NT);
    fe3_build_window(win_s, tab.window, tid, D::NT);
    for (int i = tid; i < 9 * D::C0; i += D::NT) ws[i] = w[i];
    for (int i = tid; i < D::C0; i += D::NT) bs[i] = bias[i];
    for (int i = tid; i < D::MEL_ROWS * D::MEL_P; i += D::NT) melp[i] = 0.0f;    // zero border, written once
    __syncthreads();

    long long wi = blockIdx.x;
    if (wi < n_windows) stager.issue(0, src.at(wi), tid);
    for (int it = 0; wi < n_windows; wi += gridDim.x, ++it) {
        const int16_t* x = stager.wait(0, it & 1, src.at(wi));
        fe3_logmel_window(x, smem, win_s, tw, tab, melp + D::MEL_P + 1, D::MEL_P, 1, tid);       // ends with a CTA barrier
        const long long wn = wi + gridDim.x;
        if (wn < n_windows) stager.issue(0, src.at(wn), tid);
        if (mel_dump != nullptr) {
            float* md = mel_dump + wi * (long long)(D::F * D::TT);
            for (int i = tid; i < D::F * D::TT; i += D::NT) md[i] = melp[(i / D::TT + 1) * D::MEL_P + (i % D::TT) + 1];
        }
        bc_stage_conv<ACT>(melp, ws, bs, out + wi * (long long)(D::H1 * D::W1 * D::C0), tid);
        __syncthreads();                 // the mel plane and the FFT scratch are free for the next window
    }
}
#endif

// depthwise 3x3, stride (sh, sw), pad 1, no bias / activation, channel-last.  w [C][9].
// in [n][H*W][C] -> dwo [n][Ho*Wo][C] and ctr [n][Ho*Wo][C] = in at (sh y, sw x) (the shortcut's input).
__global__ void __launch_bounds__(256)
bc_dw_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ dwo, float* __restrict__ ctr,
             long long n, int C, int H, int W, int sh, int sw) {
    const int Ho = (H - 1) / sh + 1, Wo = (W - 1) / sw + 1, c4n = C / 4;
    const long long total = n * Ho * Wo * c4n;
    // the filter taps, transposed to [tap][C] in shared memory: a thread's four channels of a tap are ONE 128-bit load
    // (thirty-six scalar weight loads per output quad made the kernel load-instruction bound)
    __shared__ __align__(16) float ws[9 * 256];
    for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) ws[i] = __ldg(w + (i % C) * 9 + i / C);
    __syncthreads();
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(t % c4n);
        const long long pix = t / c4n;
        const int x = (int)(pix % Wo), y = (int)((pix / Wo) % Ho);
        const long long b = pix / ((long long)Wo * Ho);
        const float* src = in + b * (long long)H * W * C + 4 * c4;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f), centre = s;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int yy = y * sh - 1 + r, xx = x * sw - 1 + q;
                if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                const float4 v = __ldg(reinterpret_cast<const float4*>(src + ((long long)yy * W + xx) * C));
                const int tap = r * 3 + q;
                const float4 wt = *reinterpret_cast<const float4*>(ws + tap * C + 4 * c4);
                s.x = fmaf(v.x, wt.x, s.x);
                s.y = fmaf(v.y, wt.y, s.y);
                s.z = fmaf(v.z, wt.z, s.z);
                s.w = fmaf(v.w, wt.w, s.w);
                if (r == 1 && q == 1) centre = v;
            }
        reinterpret_cast<float4*>(dwo + pix * C)[c4] = s;
        reinterpret_cast<float4*>(ctr + pix * C)[c4] = centre;
    }
}

// Block tail as two row GEMMs per tile of rows:  out = act(dwo Wpw + bpw) + (ctr Wsc + bsc).
constexpr int kBcRows = 112;          // rows (pixels) per CTA tile
inline size_t bc_block_smem_bytes(int Cin) {
    return sizeof(float) * ((size_t)2 * kTcnWBuf + (size_t)2 * kBcRows * Cin + (size_t)kBcRows * 128);
}

__global__ void __launch_bounds__(kTcnNT, 1)
bc_block_gemm_kernel(const float* __restrict__ dwo, const float* __restrict__ ctr, const float* __restrict__ Wpw,
                     const float* __restrict__ bpw, const float* __restrict__ Wsc, const float* __restrict__ bsc,
                     float* __restrict__ out, long long rows, int Cin, int Cout, int act) {
    NWW_DYN_SMEM(smem);
    float* wbuf = reinterpret_cast<float*>(smem);
    float* a_dw = wbuf + 2 * kTcnWBuf;
    float* a_ct = a_dw + (size_t)kBcRows * Cin;
    float* tmp = a_ct + (size_t)kBcRows * Cin;           // [rows][<= 128] shortcut result of the current column half
    const int tid = threadIdx.x;
    for (long long r0 = (long long)blockIdx.x * kBcRows; r0 < rows; r0 += (long long)gridDim.x * kBcRows) {
        const int nr = (int)((rows - r0 < kBcRows) ? (rows - r0) : kBcRows);
        __syncthreads();
        const int n4 = nr * Cin / 4;
        const float4* gd = reinterpret_cast<const float4*>(dwo + r0 * Cin);
        const float4* gc = reinterpret_cast<const float4*>(ctr + r0 * Cin);
        for (int i = tid; i < n4; i += kTcnNT) {
            reinterpret_cast<float4*>(a_dw)[i] = __ldg(gd + i);
            reinterpret_cast<float4*>(a_ct)[i] = __ldg(gc + i);
        }
        __syncthreads();
        for (int n0 = 0; n0 < Cout; n0 += 128) {
            const int nc = (Cout - n0 < 128) ? (Cout - n0) : 128;
            // shortcut: tmp = ctr Wsc + bsc
            tcn_layer_any(TcnLayerArgs{a_ct, 0, 1, 0, Cin, Cin, Wsc + n0, bsc + n0, tmp, 0, nc, nr, 1, 0, nullptr, 0, 0, 0, Cout, nc, 0, 0},
                          wbuf, tid);
            // pointwise: out = act(dwo Wpw + bpw) + tmp
            tcn_layer_any(TcnLayerArgs{a_dw, 0, 1, 0, Cin, Cin, Wpw + n0, bpw + n0, out + r0 * Cout + n0, 0, nc, nr, 1, 2 + act, tmp, 0,
                                       1, 0, Cout, Cout, nc, 0},
                          wbuf, tid);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// The same block tail on the 5th-generation tensor cores (default).  One CTA owns tiles of 128 pixels:
//   * the tile's depthwise output and centre-tap rows ([128][Cin] FP32 each) are converted to bf16 hi / lo
//     and laid out as un-swizzled K-major UMMA operands  [K-group][row][8 channels]  (SBO 128 B, LBO 2 KB);
//   * for every chunk of 64 output channels the folded weights of both 1x1 convolutions arrive from the
//     engine's pre-split bf16 copy ([chunk][pw|sc][hi|lo][K-group][64][8], built once at create time);
//   * 2 x (Cin / 16) x 3 tcgen05.mma (128 x 64 x 16, bf16 split products a_hi w_hi + a_lo w_hi + a_hi w_lo)
//     accumulate the two products side by side in 128 TMEM columns;
//   * epilogue: act(pw + b_pw) + (sc + b_sc) straight from TMEM to the channel-last output.
// GELU / SiLU out of line: one copy of the erf / exp code in the kernel instead of one per unrolled epilogue element
__device__ __noinline__ float bc_act_slow(float x, int act) { return apply_act(x, act); }

constexpr int kBcuRows = 128, kBcuNC = 64, kBcuNT = 256;
inline size_t bcu_smem_bytes(int Cin) {
    return (size_t)4 * kBcuRows * Cin * 2 /* A: dw, ctr x hi, lo */ + (size_t)4 * kBcuNC * Cin * 2 /* B chunk */ + 128;
}

__global__ void __launch_bounds__(kBcuNT, 3)
bc_block_umma_kernel(const float* __restrict__ dwo, const float* __restrict__ ctr, const uint4* __restrict__ wq /* see above */,
                     const float* __restrict__ bpw, const float* __restrict__ bsc, float* __restrict__ out, long long rows,
                     int Cin, int Cout, int act) {
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int a_op = kBcuRows * Cin * 2;                 // bytes of one A operand (dw_hi, dw_lo, ct_hi, ct_lo)
    const int b_op = kBcuNC * Cin * 2;                   // bytes of one B operand (pw_hi, pw_lo, sc_hi, sc_lo)
    unsigned char* a_s = smem;
    unsigned char* b_s = smem + 4 * a_op;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * a_op + 4 * b_op);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint64_t da0 = umma_desc_noswz(smem_u32(a_s), kBcuRows * 16, 128), db0 = umma_desc_noswz(smem_u32(b_s), kBcuNC * 16, 128);
    const uint32_t idesc = umma_idesc_bf16(128, kBcuNC);
    const int kg_n = Cin / 8;                            // 16-byte K groups
    const int n_chunks = Cout / kBcuNC;
    uint32_t phase = 0;
    int b_loaded = -1;                                   // weight chunk currently in shared memory

    for (long long r0 = (long long)blockIdx.x * kBcuRows; r0 < rows; r0 += (long long)gridDim.x * kBcuRows) {
        // ---- A: FP32 rows -> bf16 hi / lo UMMA operands --------------------------------------------------------
        for (int i = tid; i < kBcuRows * kg_n; i += kBcuNT) {
            const int g = i / kBcuRows, r = i - g * kBcuRows;           // consecutive threads -> consecutive rows
            uint4 dh = make_uint4(0, 0, 0, 0), dl = dh, ch = dh, cl = dh;
            if (r0 + r < rows) {
                const float4* pd = reinterpret_cast<const float4*>(dwo + (r0 + r) * Cin + 8 * g);
                const float4* pc = reinterpret_cast<const float4*>(ctr + (r0 + r) * Cin + 8 * g);
                const float4 d0 = __ldg(pd), d1 = __ldg(pd + 1), c0 = __ldg(pc), c1 = __ldg(pc + 1);
                const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
                const float cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                uint32_t h[8], l[8], hc[8], lc[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    h[k] = float_to_bf16_bits(dv[k]);
                    l[k] = float_to_bf16_bits(dv[k] - bf16_bits_to_float(h[k]));
                    hc[k] = float_to_bf16_bits(cv[k]);
                    lc[k] = float_to_bf16_bits(cv[k] - bf16_bits_to_float(hc[k]));
                }
                dh = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
                dl = make_uint4(l[0] | (l[1] << 16), l[2] | (l[3] << 16), l[4] | (l[5] << 16), l[6] | (l[7] << 16));
                ch = make_uint4(hc[0] | (hc[1] << 16), hc[2] | (hc[3] << 16), hc[4] | (hc[5] << 16), hc[6] | (hc[7] << 16));
                cl = make_uint4(lc[0] | (lc[1] << 16), lc[2] | (lc[3] << 16), lc[4] | (lc[5] << 16), lc[6] | (lc[7] << 16));
            }
            const int off = (g * kBcuRows + r) * 16;
            *reinterpret_cast<uint4*>(a_s + 0 * a_op + off) = dh;
            *reinterpret_cast<uint4*>(a_s + 1 * a_op + off) = dl;
            *reinterpret_cast<uint4*>(a_s + 2 * a_op + off) = ch;
            *reinterpret_cast<uint4*>(a_s + 3 * a_op + off) = cl;
        }
        for (int nc = 0; nc < n_chunks; ++nc) {
            // ---- B: this chunk's pre-split weights (skipped when the single chunk is already resident) ---------
            if (b_loaded != nc) {
                const uint4* src = wq + (size_t)nc * (4 * b_op / 16);
                for (int i = tid; i < 4 * b_op / 16; i += kBcuNT) reinterpret_cast<uint4*>(b_s)[i] = __ldg(src + i);
                b_loaded = nc;
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                // (descriptors = two kernel-constant bases + offsets in 16-byte units: the issuing thread is a single dependent
                // chain on the tile's critical path, so nothing is re-encoded per MMA)
                for (int gemm = 0; gemm < 2; ++gemm) {                  // 0: pointwise (A = dw), 1: shortcut (A = ctr)
                    const uint32_t d_tmem = tmem_base + (uint32_t)(gemm * kBcuNC);
                    uint64_t dah = da0 + (uint64_t)(((2 * gemm) * a_op) >> 4), dbh = db0 + (uint64_t)(((2 * gemm) * b_op) >> 4);
                    const uint64_t a_lo = (uint64_t)(a_op >> 4), b_lo = (uint64_t)(b_op >> 4);
                    for (int ks = 0; ks < Cin / 16; ++ks, dah += 2 * kBcuRows, dbh += 2 * kBcuNC) {   // one MMA = 16 channels = 2 K groups
                        umma_bf16(d_tmem, dah, dbh, idesc, ks != 0);
                        umma_bf16(d_tmem, dah + a_lo, dbh, idesc, 1);
                        umma_bf16(d_tmem, dah, dbh + b_lo, idesc, 1);
                    }
                }
                umma_commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
            tc_fence_after();
            // ---- epilogue: warp -> (TMEM lane quarter, 32-column half) ---------------------------------------------
            {
                const int q = warp & 3, hcol = warp >> 2;
                float pv[32], sv[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hcol * 32), pv);
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kBcuNC + hcol * 32), sv);
                tc_fence_before();
                const long long r = r0 + q * 32 + lane;
                if (r < rows) {
                    const int n0 = nc * kBcuNC + hcol * 32;
                    float4* dst = reinterpret_cast<float4*>(out + r * Cout + n0);
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        float v[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int n = n0 + 4 * j4 + j;
                            const float pre = pv[4 * j4 + j] + __ldg(bpw + n);
                            v[j] = (act == ACT_RELU ? fmaxf(pre, 0.0f) : bc_act_slow(pre, act)) + (sv[4 * j4 + j] + __ldg(bsc + n));
                        }
                        dst[j4] = make_float4(v[0], v[1], v[2], v[3]);
                    }
                }
            }
            __syncthreads();          // TMEM and (for multi-chunk layers) the weight buffer are free again
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 128);
    }
}

// host: folded (Cin, Cout) FP32 weights of both 1x1 convolutions -> [chunk][pw|sc][hi|lo][K-group][64][8] bf16
inline void bcu_pack_weights(const float* pw, const float* sc, int Cin, int Cout, std::vector<uint16_t>* out) {
    auto bf16_rn = [](float x) {
        uint32_t u;
        memcpy(&u, &x, 4);
        u += 0x7FFFu + ((u >> 16) & 1u);
        return (uint16_t)(u >> 16);
    };
    auto bf16_f = [](uint16_t b) {
        uint32_t u = (uint32_t)b << 16;
        float f;
        memcpy(&f, &u, 4);
        return f;
    };
    const int n_chunks = Cout / kBcuNC, op = kBcuNC * Cin;          // elements per operand
    out->assign((size_t)n_chunks * 4 * op, 0);
    for (int nc = 0; nc < n_chunks; ++nc)
        for (int gemm = 0; gemm < 2; ++gemm) {
            const float* w = gemm ? sc : pw;
            for (int k = 0; k < Cin; ++k)
                for (int n = 0; n < kBcuNC; ++n) {
                    const float v = w[(size_t)k * Cout + nc * kBcuNC + n];
                    const uint16_t hi = bf16_rn(v), lo = bf16_rn(v - bf16_f(hi));
                    const size_t base = ((size_t)nc * 4 + 2 * gemm) * op + (size_t)(k >> 3) * kBcuNC * 8 + n * 8 + (k & 7);
                    (*out)[base] = hi;
                    (*out)[base + op] = lo;
                }
        }
}

// global average pool, channel-last: in [n][P][C] -> out [n][C]   (C % 4 == 0)
// One thread = one (window, channel quad), eight 128-bit loads in flight; pixels are added in order.
__global__ void __launch_bounds__(256) bc_gap_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, int P, int C) {
    const int c4n = C / 4;
    const long long total = n * c4n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long b = t / c4n;
        const int c4 = (int)(t - b * c4n);
        const float4* src = reinterpret_cast<const float4*>(in + b * (long long)P * C) + c4;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        int p = 0;
        for (; p + 8 <= P; p += 8) {
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __ldg(src + (long long)(p + i) * c4n);
#pragma unroll
            for (int i = 0; i < 8; ++i) { s.x += v[i].x; s.y += v[i].y; s.z += v[i].z; s.w += v[i].w; }
        }
        for (; p < P; ++p) {
            const float4 v = __ldg(src + (long long)p * c4n);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        reinterpret_cast<float4*>(out + b * C)[c4] = make_float4(s.x / (float)P, s.y / (float)P, s.z / (float)P, s.w / (float)P);
    }
}

}  // namespace nww
