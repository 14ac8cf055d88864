// nww_rnn.cuh — the recurrent heads on a (T, F) log-mel sequence: GRUModel, LSTMModel and RNNModel
// (reference modules/architectures.py:129-146, 83-99, 149-161: one bidirectional torch.nn.GRU / LSTM layer,
// `out[:, -1, :]`, then a Linear).  At the last time index the forward direction has run all S steps and the
// reverse direction exactly one (on x[S-1], from a zero state), so a window costs S + 1 cell evaluations.
//
// One CTA owns 128 windows = the 128 rows of a tcgen05 tile and keeps their state on chip for the whole sequence.
// Per step the pre-activations of ALL gates are one GEMM,  [x_t | 1 | h] (128 x K)  x  Wcat (K x 4H):
//   LSTM columns  [i | f | g | o],            Wcat = [W_ih^T ; b_ih + b_hh ; W_hh^T]
//   GRU  columns  [r | z | n_x | n_h],        n_x gets only the x rows (and b_in), n_h only the h rows (and b_hn),
//                                             so that n = tanh(n_x + r * n_h) keeps the reference's form
// (the bias rides on a constant-1 input column, which the padding of F = 40 to the MMA K leaves free).
// Operands are two-term IEEE-half splits of the FP32 values (hi = half(v), lo = half(v - hi): 22 significant bits,
// which the 100 dB range of the log-mel inputs needs; bf16 pairs carry 16) with three products per K step
// (hi hi + lo hi + hi lo) and FP32 accumulation in TMEM, un-swizzled K-major core matrices as everywhere in
// this engine.  The weights (L2-resident, ~350 KB per step for H = 128) stream through a ring of six 16 KB
// shared-memory slots by cp.async.bulk in sub-slices of (16 K rows) x (256 columns) x (both half terms): each
// sub-slice feeds the three split products as MMAs of N = 256, so the A operand is fetched once per 256 columns
// (with N = 64 blocks the MMAs were shared-memory bound at 54 cycles each, measured) and several copies are in
// flight while earlier ones are consumed.  The gate math reads the accumulators straight out of
// TMEM (thread = one window x H/2 hidden units, its c / h state lives in registers), with the sigmoid / tanh
// quotients of a gate update merged so that it costs 7 (LSTM) or 5 (GRU) MUFU operations instead of 10 / 6, and
// writes the new h back into the A operand as half terms.  With H = 128 the gate columns are grouped by hidden-unit
// half ([i|f|g|o] of units 0..63, then of units 64..127) and the MMAs run half by half.  kRnnSplit lets the warps
// that own the first half do their gate math while the second half's MMAs execute (their h waits in registers until
// every MMA of the step has read the old one); measured on B200 it is SLOWER (LSTM 2.23 vs 2.08 ms per 4096 windows:
// the gate math of the late half is the critical path either way and the extra registers cost more than the overlap
// saves), so it is compiled out.
#pragma once

#ifndef NWW_CPUSIM
#include <cuda_fp16.h>
#endif

#include <cstring>
#include <vector>

#include "nww_tc.cuh"
#include "nww_tcn.cuh"

namespace nww {

constexpr int kRnnTM = 128, kRnnNW = 256, kRnnNT = 256;   // windows per CTA, worker threads, threads
constexpr int kRnnIssuer = 128;                              // the thread that issues copies and MMAs: lane 0 of a warp of the LATE half
constexpr bool kRnnSplit = false;                            // overlap the first half's gate math with the second half's MMAs (measured slower)
constexpr int kRnnRingBytes = 96 * 1024;                     // weight sub-slices in flight in shared memory
enum { RNN_GRU = 0, RNN_LSTM = 1 };

template <int H, int IN> struct RnnDims {
    static constexpr int KX = (IN + 1 + 15) / 16 * 16;           // x columns + the constant-1 column, padded to the MMA K
    static constexpr int K = KX + H;
    static constexpr int NCOL = 4 * H;
    static constexpr int NMMA = NCOL > 256 ? 256 : NCOL, NHALF = NCOL / NMMA;     // MMA N; column halves (by hidden unit)
    static constexpr int UH = H / NHALF;                          // hidden units per column half
    static constexpr int A_PLANE = (K / 8) * kRnnTM * 16;         // bytes of one half term of [x | h]
    static constexpr int A_BYTES = 2 * A_PLANE;
    static constexpr int TERM = 2 * NMMA * 16;                    // one half term of a sub-slice
    static constexpr int SUB = 2 * TERM;                          // one weight sub-slice: 16 K rows x NMMA columns, both terms
    static constexpr int RING = kRnnRingBytes / SUB;
    static constexpr int NSUB_F = NHALF * (K / 16);               // sub-slices per forward step
    static constexpr int NSUB_B = NHALF * (KX / 16);              // reverse direction: x rows only
    static constexpr int X_BYTES = kRnnTM * IN * 4;
    static constexpr size_t SMEM = (size_t)A_BYTES + (size_t)RING * SUB + X_BYTES + 512;
    // accumulator (TMEM) column of gate g, hidden unit j
    __host__ __device__ static constexpr int col(int g, int j) { return (j / UH) * (4 * UH) + g * UH + j % UH; }
};

// IEEE half <-> float on the host (round to nearest even, subnormals kept)
inline uint16_t rnn_f2h(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    const uint32_t sign = (u >> 16) & 0x8000u, au = u & 0x7FFFFFFFu;
    if (au >= 0x47800000u) return (uint16_t)(sign | 0x7C00u);                 // >= 65536 (or inf / nan): inf
    if (au < 0x38800000u) {                                                    // < 2^-14: subnormal half, ulp 2^-24
        float a;
        memcpy(&a, &au, 4);
        const float r = a * 16777216.0f;                                       // exact scaling
        uint32_t m = (uint32_t)r;
        const float frac = r - (float)m;
        if (frac > 0.5f || (frac == 0.5f && (m & 1u))) ++m;
        return (uint16_t)(sign | m);
    }
    uint32_t v = au + 0xFFFu + ((au >> 13) & 1u);                              // round the 13 dropped mantissa bits
    return (uint16_t)(sign | ((v - 0x38000000u) >> 13));
}
inline float rnn_h2f(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu;
    float f;
    if (e == 0) {
        f = (float)m * (1.0f / 16777216.0f);
        uint32_t u;
        memcpy(&u, &f, 4);
        u |= sign;
        memcpy(&f, &u, 4);
        return f;
    }
    const uint32_t u = sign | ((e + 112u) << 23) | (m << 13);
    memcpy(&f, &u, 4);
    return f;
}

// host: w [K][4H] FP32 (columns gate-major: g * H + j) -> the stream of sub-slices the kernel consumes: for every column
// half (hidden units [hf * UH, +UH) of all four gates), for every 16 K rows:  [half term 2][K group 2][NMMA columns][8 k]
// (un-swizzled K-major core matrices, LBO = NMMA * 16 B, SBO = 128 B)
inline void rnn_pack_weights(const float* w, int K, int H, std::vector<uint16_t>* out) {
    const int ncol = 4 * H, nmma = ncol > 256 ? 256 : ncol, nhalf = ncol / nmma, uh = H / nhalf;
    const size_t term = (size_t)2 * nmma * 8, sub = 2 * term;      // elements per term / sub-slice
    out->assign((size_t)nhalf * (K / 16) * sub, 0);
    for (int k = 0; k < K; ++k)
        for (int g = 0; g < 4; ++g)
            for (int j = 0; j < H; ++j) {
                const float v = w[(size_t)k * ncol + g * H + j];
                const uint16_t t0 = rnn_f2h(v), t1 = rnn_f2h(v - rnn_h2f(t0));
                const int hf = j / uh, c = g * uh + j % uh;        // column inside the half
                const size_t first = (size_t)hf * (K / 16) + k / 16;
                const size_t at = (size_t)((k >> 3) & 1) * nmma * 8 + (size_t)c * 8 + (k & 7);
                (*out)[first * sub + at] = t0;
                (*out)[first * sub + term + at] = t1;
            }
}

// exp(-x) and exp(-2x) with the argument clamped so that products of three (1 + e) terms stay finite in FP32;
// sigmoid(28) and tanh(14) differ from 1 by < 2e-12
__device__ __forceinline__ float rnn_e1(float x) { return __expf(-fminf(fmaxf(x, -28.0f), 28.0f)); }
__device__ __forceinline__ float rnn_e2(float x) { return __expf(-2.0f * fminf(fmaxf(x, -14.0f), 14.0f)); }

#ifndef NWW_CPUSIM
// IEEE half bits of x, round to nearest even: the packed conversion runs on the ALU pipe (F2FP), the scalar one on the XU pipe
// that the gate math already saturates
__device__ __forceinline__ uint32_t rnn_half_bits(float x) {
    uint32_t d;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(0.0f), "f"(x));
    return d & 0xFFFFu;
}
__device__ __forceinline__ float rnn_half_value(uint32_t b) { return __half2float(__ushort_as_half((unsigned short)b)); }
#else
inline uint32_t rnn_half_bits(float x) { return rnn_f2h(x); }
inline float rnn_half_value(uint32_t b) { return rnn_h2f((uint16_t)b); }
#endif
__device__ __forceinline__ uint4 rnn_pack8(const uint32_t* b) {
    return make_uint4(b[0] | (b[1] << 16), b[2] | (b[3] << 16), b[4] | (b[5] << 16), b[6] | (b[7] << 16));
}

// x_tm: window w's sequence starts at x_tm + w * win_stride, step t at + t * IN  (IN % 4 == 0)
// tm: windows per CTA tile (a multiple of 32, <= 128): small batches take fewer rows per CTA so that every SM gets
//     a tile; the MMAs always span 128 rows, the unused ones hold zeros / stale finite data and are never read.
// wq_f / wq_b: rnn_pack_weights of the forward [K][4H] / reverse [KX][4H] matrices.  feat [n][2H] = [h_fwd(S-1) | h_bwd]
template <int CELL, int H, int IN>
__global__ void __launch_bounds__(kRnnNT, 1)
rnn_seq_kernel(const float* __restrict__ x_tm, long long win_stride, int S, long long n, int tm, const uint4* __restrict__ wq_f,
               const uint4* __restrict__ wq_b, float* __restrict__ feat) {
    using D = RnnDims<H, IN>;
    constexpr int KX = D::KX, HS = H / 2, NCH = HS / 16, TM = kRnnTM, R = D::RING, NW = kRnnNW;
    constexpr bool SPLIT = kRnnSplit && D::NHALF == 2;              // gate math of half 0 overlaps the MMAs of half 1
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool worker = tid < NW, issuer = tid == kRnnIssuer;
    unsigned char* a_s = smem;                                       // [term][K group][128 rows][16 B]
    unsigned char* b_s = smem + D::A_BYTES;                          // ring of R weight sub-slices
    float* xbuf = reinterpret_cast<float*>(b_s + R * D::SUB);        // [128][IN]: the next step's inputs
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(xbuf) + D::X_BYTES);
    uint64_t* bar_empty = bar_full + R;
    uint64_t* bar_done = bar_full + 2 * R;                           // all of the step's MMAs
    uint64_t* bar_done0 = bar_full + 2 * R + 1;                      // the first column half's MMAs
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 2 * R + 2);
    if (tid == 0) {
        for (int i = 0; i < 2 * R + 2; ++i) mbar_init(bar_full + i, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, D::NCOL);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = umma_idesc_f16(128, D::NMMA);
    const uint64_t da0 = umma_desc_noswz(smem_u32(a_s), TM * 16, 128);
    const uint64_t da1 = da0 + (uint64_t)(D::A_PLANE >> 4);
    const uint64_t db_ring = umma_desc_noswz(smem_u32(b_s), D::NMMA * 16, 128);
    const int q = warp & 3, hc = (warp >> 2) & 1, m_own = q * 32 + lane;
    // the issuer's view of the weight stream: consumer slot / phase, producer slot and position inside the tile
    uint32_t c_slot = 0, c_par = 0, p_slot = 0;
    uint32_t done_phase = 0;
    const int tile_subs = S * D::NSUB_F + D::NSUB_B;
    int p_pos = 0, p_in_step = 0;

    auto produce = [&]() {                                            // issuer: next sub-slice of the tile -> slot p_slot
        const bool bwd = p_pos >= S * D::NSUB_F;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(bwd ? wq_b : wq_f) + (size_t)p_in_step * D::SUB;
        mbar_expect_tx(bar_full + p_slot, D::SUB);
        bulk_g2s(b_s + (size_t)p_slot * D::SUB, src, D::SUB, bar_full + p_slot);
        p_slot = p_slot + 1 == R ? 0 : p_slot + 1;
        ++p_pos;
        if (++p_in_step == D::NSUB_F && !bwd) p_in_step = 0;
    };
    auto load_x = [&](long long w0, int mt, int t) {                  // workers: x_t of the tile -> xbuf (cp.async)
        if (worker) {
            for (int i = tid; i < TM * (IN / 4); i += NW) {
                const int m = i / (IN / 4), c4 = i - m * (IN / 4);
                if (m < mt) tcn_cp_async16(xbuf + m * IN + 4 * c4, x_tm + (w0 + m) * win_stride + (long long)t * IN + 4 * c4);
            }
        }
        tcn_cp_commit();
    };

    for (int i = tid; i < D::A_BYTES / 16; i += kRnnNT) reinterpret_cast<uint4*>(a_s)[i] = make_uint4(0, 0, 0, 0);
    for (long long w0 = (long long)blockIdx.x * tm; w0 < n; w0 += (long long)gridDim.x * tm) {
        const int mt = (n - w0 < tm) ? (int)(n - w0) : tm;
        const int mt32 = (mt + 31) & ~31;                            // rows that are converted / updated (whole warps)
        __syncthreads();
        for (int i = tid; i < (H / 8) * TM; i += kRnnNT) {            // h = 0
            const int off = (KX / 8) * TM * 16 + i * 16;
            *reinterpret_cast<uint4*>(a_s + off) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(a_s + D::A_PLANE + off) = make_uint4(0, 0, 0, 0);
        }
        for (int i = tid; i < TM * IN; i += kRnnNT) xbuf[i] = 0.0f;  // rows past the end of the batch stay zero
        float st[HS];                                                 // LSTM: c, GRU: h  (this thread's window, HS units)
#pragma unroll
        for (int i = 0; i < HS; ++i) st[i] = 0.0f;
        __syncthreads();
        load_x(w0, mt, 0);
        if (issuer) {
            p_pos = 0;
            p_in_step = 0;
            for (int i = 0; i < R && p_pos < tile_subs; ++i) produce();   // every slot is free: the last tile's MMAs are done
        }
        tcn_cp_wait<0>();
        __syncthreads();

        for (int s = 0; s <= S; ++s) {
            const bool bwd = s == S;                                  // the reverse direction's first (and only needed) step
            if (!bwd) {
                // x_t (+ the constant 1) -> two half terms in the first KX / 8 K groups of A
                if (worker) {
                    for (int i = tid; i < (KX / 8) * mt32; i += NW) {
                        const int g = i / mt32, m = i - g * mt32;
                        uint32_t t0[8], t1[8];
                        float v8[8];
                        if (8 * g + 8 <= IN) {
                            const float4 v0 = *reinterpret_cast<const float4*>(xbuf + m * IN + 8 * g);
                            const float4 v1 = *reinterpret_cast<const float4*>(xbuf + m * IN + 8 * g + 4);
                            v8[0] = v0.x; v8[1] = v0.y; v8[2] = v0.z; v8[3] = v0.w;
                            v8[4] = v1.x; v8[5] = v1.y; v8[6] = v1.z; v8[7] = v1.w;
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const int col = 8 * g + e;
                                v8[e] = col < IN ? xbuf[m * IN + col] : (col == IN ? 1.0f : 0.0f);
                            }
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            t0[e] = rnn_half_bits(v8[e]);
                            t1[e] = rnn_half_bits(v8[e] - rnn_half_value(t0[e]));
                        }
                        const int off = (g * TM + m) * 16;
                        *reinterpret_cast<uint4*>(a_s + off) = rnn_pack8(t0);
                        *reinterpret_cast<uint4*>(a_s + D::A_PLANE + off) = rnn_pack8(t1);
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < HS; ++i) st[i] = 0.0f;           // zero state; A still holds x_{S-1}
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (s + 1 < S) load_x(w0, mt, s + 1);
            if (issuer) {
                tc_fence_after();
                const int nks = bwd ? KX / 16 : D::K / 16;
                int prev_slot = -1;
                uint32_t prev_par = 0;
                for (int hf = 0; hf < D::NHALF; ++hf) {
                    const uint32_t d = tmem_base + (uint32_t)(hf * D::NMMA);
                    for (int ks = 0; ks < nks; ++ks) {
                        const uint64_t ao = (uint64_t)((ks * 2 * TM * 16) >> 4);
                        mbar_wait_one(bar_full + c_slot, c_par);
                        tc_fence_after();
                        const uint64_t db0 = db_ring + (uint64_t)((c_slot * (uint32_t)D::SUB) >> 4), db1 = db0 + (uint64_t)(D::TERM >> 4);
                        // hi hi + lo hi + hi lo   (umma_bf16 = kind::f16; the instruction descriptor says IEEE half)
                        umma_bf16(d, da0 + ao, db0, idesc, ks != 0);
                        umma_bf16(d, da1 + ao, db0, idesc, 1);
                        umma_bf16(d, da0 + ao, db1, idesc, 1);
                        umma_commit(bar_empty + c_slot);
                        // refill the slot of the PREVIOUS sub-slice (its MMAs finish while this one's run)
                        if (prev_slot >= 0) {
                            mbar_wait_one(bar_empty + prev_slot, prev_par);
                            if (p_pos < tile_subs) produce();
                        }
                        prev_slot = (int)c_slot;
                        prev_par = c_par;
                        if (++c_slot == R) { c_slot = 0; c_par ^= 1u; }
                    }
                    if (SPLIT && hf == 0) umma_commit(bar_done0);
                }
                umma_commit(bar_done);
                mbar_wait_one(bar_empty + prev_slot, prev_par);       // == all of this step's MMAs are complete
                if (p_pos < tile_subs) produce();
            }
            // gates: this thread = window m_own, hidden units [hc * HS, +HS)
            const bool early = SPLIT && hc == 0;                      // this thread's columns are complete after half 0
            if (SPLIT && worker && early) {
                mbar_wait(bar_done0, done_phase);
                tc_fence_after();
            } else if (worker || !SPLIT) {
                if (worker) mbar_wait(bar_done, done_phase);
                tc_fence_after();
            }
            float* out = nullptr;
            if (worker && m_own < mt && (bwd || s == S - 1)) out = feat + (w0 + m_own) * (long long)(2 * H) + (bwd ? H : 0) + hc * HS;
            uint4 pk[SPLIT ? NCH : 1][4];                             // early threads: new h, held until every MMA has read the old one
            if (worker && q * 32 < mt32) {                            // else: this warp's 32 windows are past the tile
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int j0 = hc * HS + 16 * c;
                    uint32_t g0[16], g1[16], g2[16], g3[16];
                    const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16);
                    tmem_ld_32x32b_x16_nowait(ta + (uint32_t)D::col(0, j0), g0);
                    tmem_ld_32x32b_x16_nowait(ta + (uint32_t)D::col(1, j0), g1);
                    tmem_ld_32x32b_x16_nowait(ta + (uint32_t)D::col(2, j0), g2);
                    tmem_ld_32x32b_x16_nowait(ta + (uint32_t)D::col(3, j0), g3);
                    tmem_ld_wait();
                    float hn[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float a0 = __uint_as_float(g0[e]), a1 = __uint_as_float(g1[e]);
                        const float a2 = __uint_as_float(g2[e]), a3 = __uint_as_float(g3[e]);
                        if (CELL == RNN_LSTM) {
                            // c' = sig(f) c + sig(i) tanh(g),  h' = sig(o) tanh(c'):  one reciprocal per line
                            const float pi = 1.0f + rnn_e1(a0), pf = 1.0f + rnn_e1(a1), eg = rnn_e2(a2);
                            const float pg = 1.0f + eg;
                            const float cs = __fdividef(st[16 * c + e] * (pi * pg) + (1.0f - eg) * pf, pf * (pi * pg));
                            st[16 * c + e] = cs;
                            const float ec = rnn_e2(cs);
                            hn[e] = __fdividef(1.0f - ec, (1.0f + rnn_e1(a3)) * (1.0f + ec));
                        } else {
                            // r = sig(a0);  n = tanh(a2 + r a3);  h' = (1 - z) n + z h  with z = sig(a1)
                            const float r = __fdividef(1.0f, 1.0f + rnn_e1(a0));
                            const float ez = rnn_e1(a1), en = rnn_e2(a2 + r * a3);
                            hn[e] = __fdividef(ez * (1.0f - en) + st[16 * c + e] * (1.0f + en), (1.0f + ez) * (1.0f + en));
                            st[16 * c + e] = hn[e];
                        }
                    }
                    if (!bwd) {
                        uint32_t t0[16], t1[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            t0[e] = rnn_half_bits(hn[e]);
                            t1[e] = rnn_half_bits(hn[e] - rnn_half_value(t0[e]));
                        }
                        if (early) {
                            pk[SPLIT ? c : 0][0] = rnn_pack8(t0); pk[SPLIT ? c : 0][1] = rnn_pack8(t0 + 8);
                            pk[SPLIT ? c : 0][2] = rnn_pack8(t1); pk[SPLIT ? c : 0][3] = rnn_pack8(t1 + 8);
                        } else {
                            const int off = ((KX + j0) / 8 * TM + m_own) * 16;
                            *reinterpret_cast<uint4*>(a_s + off) = rnn_pack8(t0);
                            *reinterpret_cast<uint4*>(a_s + off + TM * 16) = rnn_pack8(t0 + 8);
                            *reinterpret_cast<uint4*>(a_s + D::A_PLANE + off) = rnn_pack8(t1);
                            *reinterpret_cast<uint4*>(a_s + D::A_PLANE + off + TM * 16) = rnn_pack8(t1 + 8);
                        }
                    }
                    if (out != nullptr) {
#pragma unroll
                        for (int e4 = 0; e4 < 4; ++e4)
                            reinterpret_cast<float4*>(out + 16 * c)[e4] = make_float4(hn[4 * e4], hn[4 * e4 + 1], hn[4 * e4 + 2], hn[4 * e4 + 3]);
                    }
                }
            }
            if (SPLIT && worker && early) {
                mbar_wait(bar_done, done_phase);                      // the second half's MMAs have read the old h
                if (!bwd && q * 32 < mt32) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        const int off = ((KX + hc * HS + 16 * c) / 8 * TM + m_own) * 16;
                        *reinterpret_cast<uint4*>(a_s + off) = pk[SPLIT ? c : 0][0];
                        *reinterpret_cast<uint4*>(a_s + off + TM * 16) = pk[SPLIT ? c : 0][1];
                        *reinterpret_cast<uint4*>(a_s + D::A_PLANE + off) = pk[SPLIT ? c : 0][2];
                        *reinterpret_cast<uint4*>(a_s + D::A_PLANE + off + TM * 16) = pk[SPLIT ? c : 0][3];
                    }
                }
            }
            done_phase ^= 1u;
            tc_fence_before();
            tcn_cp_wait<0>();
            __syncthreads();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, D::NCOL);
    }
}

}  // namespace nww
