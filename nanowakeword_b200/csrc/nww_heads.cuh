// nww_heads.cuh — host-side orchestration of stage A for the layer-kernel heads
// (TCN, BcResNet, CRNN-GRU, E2E mel-CNN): front end -> log-mel in the scratch arena ->
// the head's layers (nww_layers.cuh) -> one feature row per window for the dense tail.
#pragma once

#include <algorithm>
#include <functional>
#include <string>
#include <type_traits>

#include "../../include/nww_b200.h"
#include "nww_layers.cuh"
#include "nww_fe3.cuh"
#include "nww_fe5.cuh"
#include "nww_stage.cuh"
#include "nww_tail.cuh"
#include "nww_tcn.cuh"
#include "nww_tcn_umma.cuh"
#include "nww_bc.cuh"
#include "nww_rowgemm.cuh"
#include "nww_conv_umma.cuh"
#include "nww_rnn.cuh"

namespace nww {

constexpr int kStageNT = 512;       // threads per stage-A CTA
constexpr int kNfb64 = 7;           // FFTs per batch, double
// REF64x101 (51 packed FFT-400 per window, radix 5 x 5 x 4 x 4): a pass has 80 (radix 5) or 100 (radix 4) butterflies per
// FFT, so 7 FFTs per batch leave 512 threads 55 % / 68 % busy in their second round; 17 FFTs per batch = exactly three
// batches per window, 89 % / 83 % busy, a third of the CTA barriers (109 KB of work buffer).
constexpr int kNfbRef = 17;
constexpr long long kTcnRowsMinWindows = 2048;     // TCN: windows per launch from which one row GEMM per layer beats the cone kernel
constexpr int kNfb32 = 13;          // FFTs per batch, float

struct ConvW { const float* w = nullptr; const float* b = nullptr; };

struct HeadWeights {
    int arch = -1;
    // TCN (default channels [64, 64, 128], k = 3)
    int tcn_levels = 0, tcn_k = 3, tcn_in = 0;
    int tcn_ch[8] = {0};
    ConvW tcn_c1[8], tcn_c2[8], tcn_down[8];
    bool crnn_cnn2 = false;           // conv1 + conv2 through cnn2_stage_kernel (default channel counts 16, 32, 32)
    bool tcn_cone = false;            // fused dependency-cone kernel (nww_tcn.cuh) instead of the layer kernels
    TcnConeParams tcn_plan{};
    // the cone as one K-chunked tcgen05 row GEMM per layer over ALL windows of a launch group (nww_rowgemm.cuh): the
    // weights of a layer stream from L2 once per 128-row tile instead of once per 4 windows, activations go through a
    // per-window scratch in global memory (L2).  Default for the TCN head; tcn_umma / the FP32 cone are the A/B variants.
    struct TcnRowLayer {
        long long a_off, o_off, r_off;            // float offsets: A / out / residual base inside mel (level 0 input) or act scratch
        int a_in_mel, o_in_feat, has_res;
        int a_row_stride, r_row_stride, n_pos, K, k_valid, N, act, pre_relu;
        const uint4* wq;
        const float* bias;
    };
    int tcn_rows_pw = 0;              // floats of activation scratch per window in the row-GEMM form (cone + per-level residuals)
    int tcn_res_off[8] = {0};
    bool tcn_rows = false;
    bool tcn_fused = false;           // all layers (and the stream-mode gather) in one cooperative launch (opt-in)
    int tcn_n_row_layers = 0;
    TcnRowLayer tcn_row[12];
    bool tcn_umma = false;            // ... with the layer GEMMs on tcgen05 (nww_tcn_umma.cuh)
    TcnUmmaParams tcn_uplan{};
    const uint4* tcn_wq = nullptr;
    // BcResNet
    ConvW bc_init, bc_pw[3], bc_sc[3];
    const float* bc_dw[3] = {nullptr, nullptr, nullptr};
    const uint4* bc_wq[3] = {nullptr, nullptr, nullptr};     // pre-split bf16 UMMA weights (nww_bc.cuh); null: FP32 row GEMM
    // CRNN
    int crnn_levels = 0, crnn_ch[4] = {0}, gru_hidden = 0, gru_in = 0;
    ConvW crnn_conv[4];
    const float *gru_wih_f = nullptr, *gru_whh_f = nullptr, *gru_bih_f = nullptr, *gru_bhh_f = nullptr;
    const float *gru_wih_b = nullptr, *gru_bih_b = nullptr, *gru_bhh_b = nullptr;
    const float *gru_wih_f_nk = nullptr, *gru_wih_b_nk = nullptr;    // [3H][In] copies for the dense kernel
    const float *gru_wih_f_kn = nullptr, *gru_wih_b_kn = nullptr;    // [In][3H] copies for the row GEMM (preferred)
    const uint4* gru_whh_q = nullptr;                                // W_hh^T pre-split for gru_tc_kernel (H = 128)
    const uint4 *gru_wih_f_q = nullptr, *gru_wih_b_q = nullptr;      // W_ih^T pre-split for rowgemm_umma_kernel
    // E2E mel-CNN
    ConvW e2e_conv[3];
    const uint4* e2e_wq[3] = {nullptr, nullptr, nullptr};    // conv2 / conv3 weights as bf16 UMMA operands (index 1, 2)
    ConvUmmaPlan e2e_plan[3];
    bool bc_stage = true;             // BcResNet batch path: front end + init conv as one stage kernel (reserved[0] bit 9 clears it)
    // CRNN conv3 on the same kernel
    const uint4* crnn_wq3 = nullptr;
    ConvUmmaPlan crnn_plan3;
    // GRU / LSTM / RNN heads (nww_rnn.cuh)
    int rnn_cell = 0, rnn_hidden = 0;
    const uint4 *rnn_wq_f = nullptr, *rnn_wq_b = nullptr;
    // QuartzNet (qn_dw_kernel + rowgemm_kc_umma_kernel)
    struct QnBlock { int C = 0, Cp = 0, N = 0, k = 0, K = 0, has_res = 0; const float *dw = nullptr, *b = nullptr; const uint4* wq = nullptr; };
    int qn_blocks = 0, qn_max_k = 0, qn_max_n = 0, qn_t = 0, qn_cin = 0;    // sequence length / channels the blocks see
    QnBlock qn[16];
    // raw-audio front end (E2ERawQuartzNet): strided Conv1d layers as row GEMMs over channel-last buffers
    struct RawLayer {
        int cin = 0, cout = 0, k = 0, stride = 0, pad = 0, K = 0, Npad = 0, t_in = 0, t_out = 0;
        long long in_len = 0;          // floats per window of the layer's (zero-padded) input buffer
        const float* b = nullptr;
        const uint4* wq = nullptr;
    };
    int raw_layers = 0;
    RawLayer raw[4];
    // E2ERawCNN backbone: conv1 direct, conv2..4 as row GEMMs on zero-padded NHWC images
    struct RcLayer { int cin = 0, cout = 0, K = 0, Npad = 0, s = 0, Hin = 0, Win = 0, Hout = 0, Wout = 0; const float* b = nullptr; const uint4* wq = nullptr; };
    const float *rc_w1 = nullptr, *rc_b1 = nullptr;
    int rc_c1 = 0, rc_h1 = 0, rc_w1o = 0;       // conv1 output: channels, height (= front-end channels), width
    RcLayer rc[3];
    // scratch layout (floats per window)
    size_t scratch_floats = 0;
};

using BlobLookup = std::function<const float*(const char*, size_t)>;
using BlobDims = std::function<std::vector<uint32_t>(const char*)>;

// int16 grid recovery for float PCM that was produced as int16 / 32768 (nanointerpreter.py:750).
__global__ void f32_to_i16_kernel(const float* __restrict__ x, int16_t* __restrict__ y, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = rintf(x[i] * 32768.0f);
        v = fminf(fmaxf(v, -32768.0f), 32767.0f);
        y[i] = (int16_t)v;
    }
}

template <typename K> static cudaError_t set_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

#define NWW_HCUDA(call)                                                               \
    do {                                                                              \
        cudaError_t _e = (call);                                                      \
        if (_e != cudaSuccess) {                                                      \
            *err = std::string(#call) + ": " + cudaGetErrorString(_e);                \
            return NWW_ECUDA;                                                         \
        }                                                                             \
    } while (0)

template <typename G>
static int launch_frontend_f64(const FrontendTables<double>& tab, int sm_count, WindowSource pcm, long long n, float* mel,
                               int time_major, cudaStream_t st, int64_t* launches, std::string* err, int frame_lo = 0) {
    const int grid = (int)std::min<long long>(n, sm_count);
    if (pcm.fbase != nullptr) {
        // float feed: generic FP64 front end reading float32 samples straight from global memory (all frames)
        constexpr int NFB = std::is_same<G, GeoNS40x98>::value ? kNfb64 : kNfbRef;
        auto k = frontend_f32_kernel<double, G, NFB, kStageNT>;
        NWW_HCUDA(set_smem(k, FrontendSmem<double, G, NFB>::kWork));
        k<<<grid, kStageNT, FrontendSmem<double, G, NFB>::kWork, st>>>(pcm.fbase, n, tab, mel, time_major);
    } else if constexpr (std::is_same<G, GeoNS40x98>::value) {
        NWW_HCUDA(set_smem(frontend3_kernel, Fe3KernelSmem::kTotal));
        frontend3_kernel<<<grid, Fe3::NT, Fe3KernelSmem::kTotal, st>>>(pcm, n, tab, mel, time_major, frame_lo);
    } else if (tab.mel_vec_ok) {
        NWW_HCUDA(set_smem(frontend5_kernel, Fe5::kTotal));
        frontend5_kernel<<<grid, Fe5::NT, Fe5::kTotal, st>>>(pcm, n, tab, mel, time_major);
    } else {
        auto k = frontend_kernel<double, G, kNfbRef, kStageNT>;
        NWW_HCUDA(set_smem(k, FrontendSmem<double, G, kNfbRef>::kTotal));
        k<<<grid, kStageNT, FrontendSmem<double, G, kNfbRef>::kTotal, st>>>(pcm, n, tab, mel, time_major);
    }
    (*launches)++;
    NWW_HCUDA(cudaGetLastError());
    return NWW_OK;
}

static inline int ew_grid(long long total, int sm_count) {
    return (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)sm_count * 8));
}

// ------------------------------------------------------------------------------ setup
inline int setup_head_weights(int arch, int geometry, const BlobLookup& get, const BlobDims& dims, HeadWeights* hw,
                              int* feat_dim, std::string* err) {
    hw->arch = arch;
    auto need = [&](const std::string& name, size_t numel) -> const float* {
        const float* p = get(name.c_str(), numel);
        if (!p && err->empty()) *err = "weight blob: tensor '" + name + "' missing or of unexpected size";
        return p;
    };
    err->clear();
    if (arch == NWW_ARCH_TCN) {
        if (geometry != NWW_GEOM_NS40X98) { *err = "tcn head is built for the NS40x98 geometry"; return NWW_EUNSUPPORTED; }
        int cin = GeoNS40x98::N_MELS;
        hw->tcn_in = cin;
        int lv = 0;
        for (; lv < 8; ++lv) {
            const std::string p = "tcn." + std::to_string(lv);
            auto d = dims((p + ".conv1.w").c_str());
            if (d.size() != 3) break;
            const int k = (int)d[0], c = (int)d[2];
            if ((int)d[1] != cin) { *err = p + ".conv1.w: input channels do not chain"; return NWW_EINVAL; }
            hw->tcn_k = k;
            hw->tcn_ch[lv] = c;
            hw->tcn_c1[lv] = {need(p + ".conv1.w", (size_t)k * cin * c), need(p + ".conv1.b", c)};
            hw->tcn_c2[lv] = {need(p + ".conv2.w", (size_t)k * c * c), need(p + ".conv2.b", c)};
            if (cin != c) hw->tcn_down[lv] = {need(p + ".down.w", (size_t)cin * c), need(p + ".down.b", c)};
            cin = c;
        }
        if (lv == 0) { *err = "weight blob: tcn.* missing"; return NWW_EINVAL; }
        hw->tcn_levels = lv;
        *feat_dim = cin;
        size_t maxc = hw->tcn_in;
        for (int i = 0; i < lv; ++i) maxc = std::max<size_t>(maxc, hw->tcn_ch[i]);
        hw->scratch_floats = (size_t)GeoNS40x98::N_FRAMES * (hw->tcn_in + 3 * maxc);     // mel + 3 rotating planes
        // the fused cone kernel covers the reference's default shape family: k = 3, <= 4 levels, channel counts
        // that are multiples of 8 up to 128, and a cone that fits in the window
        bool cone_ok = hw->tcn_k == 3 && lv <= kTcnMaxLevels && hw->tcn_in % 4 == 0;
        for (int i = 0; i < lv; ++i) cone_ok = cone_ok && hw->tcn_ch[i] % 8 == 0 && hw->tcn_ch[i] <= 128;
        if (cone_ok) {
            TcnConeParams& P = hw->tcn_plan;
            P.levels = lv;
            P.c_in = hw->tcn_in;
            P.T = GeoNS40x98::N_FRAMES;
            for (int i = 0; i < lv; ++i) {
                P.ch[i] = hw->tcn_ch[i];
                P.w1[i] = hw->tcn_c1[i].w; P.b1[i] = hw->tcn_c1[i].b;
                P.w2[i] = hw->tcn_c2[i].w; P.b2[i] = hw->tcn_c2[i].b;
                P.wd[i] = hw->tcn_down[i].w; P.bd[i] = hw->tcn_down[i].b;
            }
            cone_ok = tcn_plan(&P);
        }
        hw->tcn_cone = cone_ok;
        // the (T, F) log-mel + (row-GEMM layers) the cone's activations of one window
        if (cone_ok) {
            // row-GEMM layers: every level that has a 1x1 downsample gets its OWN residual buffer behind the cone's
            // activations (a buffer written twice in one launch could be stale in another SM's L1 in the fused kernel)
            int pw = hw->tcn_plan.per_window;
            int cin_l = hw->tcn_in;
            for (int i = 0; i < lv; ++i) {
                hw->tcn_res_off[i] = -1;
                if (cin_l != hw->tcn_ch[i]) {
                    hw->tcn_res_off[i] = pw;
                    pw += hw->tcn_plan.n_out[i] * hw->tcn_ch[i];
                }
                cin_l = hw->tcn_ch[i];
            }
            hw->tcn_rows_pw = (pw + 3) & ~3;
            hw->scratch_floats = (size_t)GeoNS40x98::N_FRAMES * hw->tcn_in + hw->tcn_rows_pw;
        }
    } else if (arch == NWW_ARCH_BCRESNET) {
        if (geometry != NWW_GEOM_NS40X98) { *err = "bcresnet head is built for the NS40x98 geometry"; return NWW_EUNSUPPORTED; }
        hw->bc_init = {need("bc.init.w", 32 * 9), need("bc.init.b", 32)};
        const int ch[4] = {32, 64, 128, 256};
        for (int j = 0; j < 3; ++j) {
            const std::string p = "bc." + std::to_string(j);
            hw->bc_dw[j] = need(p + ".dw", (size_t)ch[j] * 9);
            hw->bc_pw[j] = {need(p + ".pw.w", (size_t)ch[j] * ch[j + 1]), need(p + ".pw.b", ch[j + 1])};
            hw->bc_sc[j] = {need(p + ".sc.w", (size_t)ch[j] * ch[j + 1]), need(p + ".sc.b", ch[j + 1])};
        }
        *feat_dim = 256;
        // mel 40*98, init 32*20*49, then per block: dw + out
        hw->scratch_floats = 3920 + 31360 + (2 * 32 * 250 + 64 * 250) + (2 * 64 * 65 + 128 * 65) + (2 * 128 * 39 + 256 * 39);
    } else if (arch == NWW_ARCH_CRNN_GRU) {
        if (geometry != NWW_GEOM_NS40X98) { *err = "crnn head is built for the NS40x98 geometry"; return NWW_EUNSUPPORTED; }
        int cin = 1, lv = 0, h = 40, w = 98;
        size_t act_floats = 0;
        for (; lv < 4; ++lv) {
            const std::string p = "crnn.conv" + std::to_string(lv);
            auto d = dims((p + ".w").c_str());
            if (d.size() != 3) break;
            const int c = (int)d[2];
            if ((int)d[0] != cin || c % kOCT) { *err = p + ".w: unsupported channel counts"; return NWW_EINVAL; }
            hw->crnn_ch[lv] = c;
            hw->crnn_conv[lv] = {need(p + ".w", (size_t)cin * 9 * c), need(p + ".b", c)};
            cin = c; h /= 2; w /= 2;
            act_floats += (size_t)c * h * w;
        }
        if (lv == 0) { *err = "weight blob: crnn.conv* missing"; return NWW_EINVAL; }
        hw->crnn_levels = lv;
        hw->gru_in = cin * h;
        hw->crnn_cnn2 = lv == 3 && hw->crnn_ch[0] == 16 && hw->crnn_ch[1] == 32 && hw->crnn_ch[2] % kOCT == 0;
        auto dh = dims("crnn.gru.fwd.w_hh");
        if (dh.size() != 2) { *err = "weight blob: crnn.gru.fwd.w_hh missing"; return NWW_EINVAL; }
        const int H = (int)dh[0];
        hw->gru_hidden = H;
        hw->gru_whh_f = need("crnn.gru.fwd.w_hh", (size_t)H * 3 * H);
        hw->gru_bhh_f = need("crnn.gru.fwd.b_hh", 3 * H);
        hw->gru_bih_f = need("crnn.gru.fwd.b_ih", 3 * H);
        hw->gru_bih_b = need("crnn.gru.bwd.b_ih", 3 * H);
        hw->gru_bhh_b = need("crnn.gru.bwd.b_hh", 3 * H);
        hw->gru_wih_f_nk = need("crnn.gru.fwd.w_ih_nk", (size_t)3 * H * hw->gru_in);
        hw->gru_wih_b_nk = need("crnn.gru.bwd.w_ih_nk", (size_t)3 * H * hw->gru_in);
        // row-GEMM layout (newer blobs); eligible shapes: In, H multiples of 4, 3H a multiple of 8
        if (hw->gru_in % 4 == 0 && H % 8 == 0 && gru2_smem_bytes(H) <= 200 * 1024 && rowgemm_smem_bytes(hw->gru_in) <= 200 * 1024) {
            hw->gru_wih_f_kn = get("crnn.gru.fwd.w_ih_kn", (size_t)3 * H * hw->gru_in);
            hw->gru_wih_b_kn = get("crnn.gru.bwd.w_ih_kn", (size_t)3 * H * hw->gru_in);
        }
        *feat_dim = 2 * H;
        // mel + conv activations + seq [w][gru_in] + gi_f [w][3H] + gi_b [3H]
        hw->scratch_floats = 3920 + act_floats + (size_t)w * hw->gru_in + (size_t)w * 3 * H + 3 * H;
        if (hw->crnn_cnn2) hw->scratch_floats = 3920 + 7680 + 2 * (size_t)w * hw->gru_in + (size_t)w * 3 * H + 3 * H;
        if (gru_smem_bytes(H) > 200 * 1024) { *err = "GRU hidden size too large for shared memory"; return NWW_EUNSUPPORTED; }
    } else if (arch == NWW_ARCH_GRU || arch == NWW_ARCH_LSTM) {
        if (geometry != NWW_GEOM_NS40X98) { *err = "recurrent heads are built for the NS40x98 geometry"; return NWW_EUNSUPPORTED; }
        auto df = dims("rnn.fwd.w"), db = dims("rnn.bwd.w");
        if (df.size() != 2 || db.size() != 2) { *err = "weight blob: rnn.fwd.w / rnn.bwd.w missing"; return NWW_EINVAL; }
        const int H = (int)df[1] / 4;
        constexpr int KX = RnnDims<128, GeoNS40x98::N_MELS>::KX;
        if ((H != 64 && H != 128) || (int)df[0] != KX + H || (int)db[0] != KX || db[1] != df[1]) {
            *err = "recurrent heads are built for 64 or 128 hidden units on 40 mel bands";
            return NWW_EUNSUPPORTED;
        }
        if (!need("rnn.fwd.w", (size_t)(KX + H) * 4 * H) || !need("rnn.bwd.w", (size_t)KX * 4 * H)) return NWW_EINVAL;
        hw->rnn_cell = arch == NWW_ARCH_GRU ? RNN_GRU : RNN_LSTM;
        hw->rnn_hidden = H;
        *feat_dim = 2 * H;
        hw->scratch_floats = (size_t)GeoNS40x98::N_FRAMES * GeoNS40x98::N_MELS;      // the (T, F) log-mel
    } else if (arch == NWW_ARCH_QUARTZNET || arch == NWW_ARCH_E2E_QUARTZNET || arch == NWW_ARCH_E2E_CNN) {
        int cin = GeoNS40x98::N_MELS, tq = GeoNS40x98::N_FRAMES;
        size_t raw_floats = 0;
        if (arch == NWW_ARCH_QUARTZNET) {
            if (geometry != NWW_GEOM_NS40X98) { *err = "quartznet head is built for the NS40x98 geometry"; return NWW_EUNSUPPORTED; }
        } else {
            // RawAudioFrontend (architectures.py:695-714): k = 41 / stride 16 first, then k = 13 / stride 4, padding k / 2
            cin = 1;
            tq = 16000;
            int nl = 0;
            for (; nl < 4; ++nl) {
                const std::string p = "raw." + std::to_string(nl);
                auto dw = dims((p + ".w").c_str()), db = dims((p + ".b").c_str());
                if (dw.size() != 2 || db.size() != 1) break;
                HeadWeights::RawLayer& L = hw->raw[nl];
                L.cin = cin; L.cout = (int)db[0]; L.k = nl == 0 ? 41 : 13; L.stride = nl == 0 ? 16 : 4; L.pad = L.k / 2;
                L.K = (int)dw[0]; L.Npad = (int)dw[1];
                L.t_in = tq; L.t_out = (tq + 2 * L.pad - L.k) / L.stride + 1;
                if (L.K != (L.k * cin + 63) / 64 * 64 || L.Npad != (L.cout + 63) / 64 * 64 || L.cout % 32 || L.Npad > 512 ||
                    (cin != 1 && cin % 4) || L.t_out < 1) {
                    *err = p + ": unsupported raw-audio front-end layer shape";
                    return NWW_EUNSUPPORTED;
                }
                // input buffer: pad rows, the data, and whatever the last GEMM row reads past it (zeros)
                const long long need_rows = (long long)L.stride * (L.t_out - 1) * cin + L.K;
                L.in_len = (std::max<long long>((long long)(L.pad + L.t_in + L.pad) * cin, need_rows) + 3) / 4 * 4;
                L.b = need(p + ".b", (size_t)L.cout);
                if (!need(p + ".w", (size_t)L.K * L.Npad)) return NWW_EINVAL;
                raw_floats += (size_t)L.in_len;
                cin = L.cout;
                tq = L.t_out;
            }
            if (nl == 0) { *err = "weight blob: raw.* missing"; return NWW_EINVAL; }
            hw->raw_layers = nl;
            raw_floats += (size_t)tq * cin;                                        // the last layer's plain output
        }
        hw->qn_t = tq;
        hw->qn_cin = cin;
        if (arch == NWW_ARCH_E2E_CNN) {
            // RawAudioBackbone (architectures.py:738-774) on the (H = cin, W = tq) one-channel image
            auto d1 = dims("rawcnn.conv1.w");
            if (d1.size() != 2 || d1[0] != 9 || d1[1] % 4) { *err = "weight blob: rawcnn.conv1.w missing or malformed"; return NWW_EINVAL; }
            hw->rc_c1 = (int)d1[1];
            hw->rc_w1 = need("rawcnn.conv1.w", (size_t)9 * hw->rc_c1);
            hw->rc_b1 = need("rawcnn.conv1.b", (size_t)hw->rc_c1);
            hw->rc_h1 = cin;
            hw->rc_w1o = (tq + 2 - 3) / 2 + 1;
            int c = hw->rc_c1, H = hw->rc_h1, W = hw->rc_w1o;
            size_t img_floats = (size_t)(H + 2) * (W + 2) * c;
            const int strides[3] = {2, 2, 1};
            for (int j = 0; j < 3; ++j) {
                const std::string p = "rawcnn.conv" + std::to_string(j + 2);
                auto dw = dims((p + ".w").c_str()), db = dims((p + ".b").c_str());
                if (dw.size() != 2 || db.size() != 1) { *err = "weight blob: " + p + " missing"; return NWW_EINVAL; }
                HeadWeights::RcLayer& L = hw->rc[j];
                L.cin = c; L.cout = (int)db[0]; L.K = (int)dw[0]; L.Npad = (int)dw[1]; L.s = strides[j];
                L.Hin = H; L.Win = W; L.Hout = (H + 2 - 3) / L.s + 1; L.Wout = (W + 2 - 3) / L.s + 1;
                if (L.K != (9 * c + 63) / 64 * 64 || L.Npad != (L.cout + 63) / 64 * 64 || L.cout % 4 || (3 * c) % 8 || L.Npad > 512) {
                    *err = p + ": unsupported RawAudioBackbone layer shape";
                    return NWW_EUNSUPPORTED;
                }
                L.b = need(p + ".b", (size_t)L.cout);
                if (!need(p + ".w", (size_t)L.K * L.Npad)) return NWW_EINVAL;
                c = L.cout; H = L.Hout; W = L.Wout;
                img_floats += j < 2 ? (size_t)(H + 2) * (W + 2) * c : (size_t)H * W * c;
            }
            *feat_dim = c;
            hw->scratch_floats = raw_floats + img_floats;
            return err->empty() ? NWW_OK : NWW_EINVAL;
        }
        int nb = 0;
        for (; nb < 16; ++nb) {
            const std::string p = "qn." + std::to_string(nb);
            auto dw = dims((p + ".w").c_str()), dd = dims((p + ".dw").c_str());
            if (dw.size() != 2 || dd.size() != 2) break;
            HeadWeights::QnBlock& B = hw->qn[nb];
            B.K = (int)dw[0]; B.N = (int)dw[1]; B.k = (int)dd[0]; B.Cp = (int)dd[1];
            B.C = cin;
            B.has_res = B.K == 2 * B.Cp;
            if ((B.K != B.Cp && !B.has_res) || B.K % kKcKC || B.N % kKcNC || B.N > 512 || B.C % 4 || B.C > B.Cp ||
                (!B.has_res && B.C != B.N)) {
                *err = p + ": unsupported QuartzNet block shape (channels must be multiples of 4, outputs multiples of 64 up to 512)";
                return NWW_EUNSUPPORTED;
            }
            B.dw = need(p + ".dw", (size_t)B.k * B.Cp);
            B.b = need(p + ".b", (size_t)B.N);
            if (!need(p + ".w", (size_t)B.K * B.N)) return NWW_EINVAL;
            hw->qn_max_k = std::max(hw->qn_max_k, B.K);
            hw->qn_max_n = std::max(hw->qn_max_n, B.N);
            cin = B.N;
        }
        if (nb == 0) { *err = "weight blob: qn.* missing"; return NWW_EINVAL; }
        hw->qn_blocks = nb;
        *feat_dim = cin;
        // input (log-mel, or the raw front end's buffers) + the GEMM operand rows + two activation planes
        hw->scratch_floats = (arch == NWW_ARCH_QUARTZNET ? (size_t)tq * GeoNS40x98::N_MELS : raw_floats) +
                             (size_t)tq * (hw->qn_max_k + 2 * (size_t)hw->qn_max_n);
    } else if (arch == NWW_ARCH_E2E_MELCNN) {
        if (geometry != NWW_GEOM_REF64X101) { *err = "e2e mel-CNN is built for the REF64x101 geometry"; return NWW_EUNSUPPORTED; }
        const int ch[4] = {1, 16, 32, 64};
        for (int j = 0; j < 3; ++j) {
            const std::string p = "e2e.conv" + std::to_string(j);
            hw->e2e_conv[j] = {need(p + ".w", (size_t)ch[j] * 9 * ch[j + 1]), need(p + ".b", ch[j + 1])};
        }
        *feat_dim = 256;
        hw->scratch_floats = 64 * 101 + 16 * 32 * 50 + 32 * 16 * 25 + 64 * 16 * 25;
    } else {
        *err = "architecture id " + std::to_string(arch) + " is not built into this library";
        return NWW_EUNSUPPORTED;
    }
    return err->empty() ? NWW_OK : NWW_EINVAL;
}

// QuartzNet blocks + mean over time on channel-last rows x [n][hw.qn_t][pitch]; `arena` provides the GEMM operand rows
// and two activation planes (n windows each)
inline int launch_quartznet_blocks(const HeadWeights& hw, const float* x, int pitch, long long n, float* arena, float* feat,
                                   int sm_count, cudaStream_t st, int64_t* launches, std::string* err) {
    auto done = [&]() -> int {
        (*launches)++;
        NWW_HCUDA(cudaGetLastError());
        return NWW_OK;
    };
    int rc;
    const int T = hw.qn_t;
    float* a = arena;
    float* plane[2] = {a + (size_t)n * T * hw.qn_max_k, a + (size_t)n * T * hw.qn_max_k + (size_t)n * T * hw.qn_max_n};
    const long long rows = n * T;
    NWW_HCUDA(set_smem(rowgemm_kc_umma_kernel<false>, rowgemm_kc_smem_bytes()));
    NWW_HCUDA(cudaFuncSetAttribute(rowgemm_kc_umma_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    for (int i = 0; i < hw.qn_blocks; ++i) {
        const HeadWeights::QnBlock& B = hw.qn[i];
        const int tgrid = (int)std::min<long long>(n * (B.Cp / 32), (long long)sm_count * 8);
        if (B.Cp % 32 == 0 && T <= kQnSeg * kQnSegs && (B.k == 33 || B.k == 39 || B.k == 11 || B.k == 13 || B.k == 17)) {
            if (B.k == 33) qn_dw_tile_kernel<33><<<tgrid, 256, 0, st>>>(x, pitch, B.dw, a, n, T, B.C, B.Cp, B.K);
            else if (B.k == 39) qn_dw_tile_kernel<39><<<tgrid, 256, 0, st>>>(x, pitch, B.dw, a, n, T, B.C, B.Cp, B.K);
            else if (B.k == 11) qn_dw_tile_kernel<11><<<tgrid, 256, 0, st>>>(x, pitch, B.dw, a, n, T, B.C, B.Cp, B.K);
            else if (B.k == 13) qn_dw_tile_kernel<13><<<tgrid, 256, 0, st>>>(x, pitch, B.dw, a, n, T, B.C, B.Cp, B.K);
            else qn_dw_tile_kernel<17><<<tgrid, 256, 0, st>>>(x, pitch, B.dw, a, n, T, B.C, B.Cp, B.K);
        } else {
            qn_dw_kernel<<<ew_grid(rows * (B.Cp / 4), sm_count), 256, 0, st>>>(x, pitch, B.dw, a, n, T, B.C, B.Cp, B.k, B.K);
        }
        if ((rc = done())) return rc;
        float* y = plane[i & 1];
        // layers wider than 256 columns run as 256-column slices: a slice leaves room for two CTAs per SM (512 TMEM columns,
        // 101 KB of shared memory each), whose operand conversion / MMA / epilogue phases then overlap
        const int n_slice = B.N > 256 && B.N % 256 == 0 ? 256 : B.N;
        const KcLaunch kl = rowgemm_kc_launch(rows, n_slice, sm_count);
        for (int n_off = 0; n_off < B.N; n_off += n_slice) {
            rowgemm_kc_umma_kernel<false><<<kl.grid, kKcBlock, kl.smem, st>>>(
                a, kc_plain(rows, B.K), kc_one_seg(B.K), B.K, B.wq, B.b, B.has_res ? nullptr : x, y, kc_plain(rows, B.N), rows, n_slice, n_slice, 1,
                kl.ring, KcView{0, 0, 0, 0, 0, 0}, 0, n_off, B.N);
            if ((rc = done())) return rc;
        }
        x = y;
        pitch = B.N;
    }
    bc_gap_kernel<<<ew_grid(n * (pitch / 4), sm_count), 256, 0, st>>>(x, feat, n, T, pitch);
    return done();
}

// RawAudioFrontend: audio -> float -> strided Conv1d layers (row GEMMs over overlapping rows).  *last_out = the last
// layer's plain [n][t_out][c_out] output, *next_free = the scratch after it.
inline int launch_raw_frontend(const HeadWeights& hw, int sm_count, WindowSource pcm, long long n, float* scratch, const float** last_out,
                               float** next_free, cudaStream_t st, int64_t* launches, std::string* err) {
    auto done = [&]() -> int {
        (*launches)++;
        NWW_HCUDA(cudaGetLastError());
        return NWW_OK;
    };
    int rc;
    float* p = scratch;
    const HeadWeights::RawLayer& L0 = hw.raw[0];
    float* in = p;
    p += (size_t)n * L0.in_len;
    raw_pcm_kernel<<<ew_grid(n * (L0.in_len / 4), sm_count), 256, 0, st>>>(pcm, n, in, (int)L0.in_len, L0.pad);
    if ((rc = done())) return rc;
    NWW_HCUDA(set_smem(rowgemm_kc_umma_kernel<true>, rowgemm_kc_smem_bytes()));
    NWW_HCUDA(cudaFuncSetAttribute(rowgemm_kc_umma_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    for (int i = 0; i < hw.raw_layers; ++i) {
        const HeadWeights::RawLayer& L = hw.raw[i];
        const bool last = i + 1 == hw.raw_layers;
        // output: the next layer's padded input buffer (data rows start after its pad rows), or a plain matrix
        const long long out_len = last ? (long long)L.t_out * L.cout : hw.raw[i + 1].in_len;
        const int out_pad = last ? 0 : hw.raw[i + 1].pad;
        float* out = p;
        p += (size_t)n * out_len;
        if (!last) {
            const int head = out_pad * L.cout, tail_off = (out_pad + L.t_out) * L.cout;
            zero_pads_kernel<<<ew_grid(n * (head + (out_len - tail_off)), sm_count), 256, 0, st>>>(out, n, out_len, head, tail_off,
                                                                                               (int)(out_len - tail_off));
            if ((rc = done())) return rc;
        }
        const long long rows = n * L.t_out;
        const KcLaunch kl = rowgemm_kc_launch(rows, L.Npad, sm_count);
        rowgemm_kc_umma_kernel<true><<<kl.grid, kKcBlock, kl.smem, st>>>(
            in, kc_seq(L.t_out, L.in_len, (long long)L.stride * L.cin, 0), kc_one_seg(L.k * L.cin), L.K, L.wq, L.b, nullptr, out,
            kc_seq(L.t_out, out_len, L.cout, out_pad), rows, L.Npad, L.cout, 1, kl.ring);
        if ((rc = done())) return rc;
        in = out;
    }
    *last_out = in;
    *next_free = p;
    return NWW_OK;
}

// E2ERawQuartzNet: raw front end -> QuartzNet blocks
inline int launch_raw_quartznet(const HeadWeights& hw, int sm_count, WindowSource pcm, long long n, float* feat, float* scratch,
                                cudaStream_t st, int64_t* launches, std::string* err) {
    const float* x = nullptr;
    float* p = nullptr;
    int rc = launch_raw_frontend(hw, sm_count, pcm, n, scratch, &x, &p, st, launches, err);
    if (rc) return rc;
    return launch_quartznet_blocks(hw, x, hw.qn_cin, n, p, feat, sm_count, st, launches, err);
}

// E2ERawCNN: raw front end -> conv1 (direct) -> conv2..4 (row GEMMs on padded NHWC images) -> global average pool
inline int launch_raw_cnn(const HeadWeights& hw, int act, int sm_count, WindowSource pcm, long long n, float* feat, float* scratch,
                          cudaStream_t st, int64_t* launches, std::string* err) {
    auto done = [&]() -> int {
        (*launches)++;
        NWW_HCUDA(cudaGetLastError());
        return NWW_OK;
    };
    const float* bf = nullptr;
    float* p = nullptr;
    int rc = launch_raw_frontend(hw, sm_count, pcm, n, scratch, &bf, &p, st, launches, err);
    if (rc) return rc;
    int C = hw.rc_c1, H = hw.rc_h1, W = hw.rc_w1o;
    float* img = p;
    p += (size_t)n * (H + 2) * (W + 2) * C;
    zero_border_kernel<<<ew_grid(n * 2 * (H + W + 2) * (C / 4), sm_count), 256, 0, st>>>(img, n, H + 2, W + 2, C, H, W);
    if ((rc = done())) return rc;
    if (C == 24)
        rawcnn_conv1_px_kernel<6><<<ew_grid(n * H * W, sm_count), 256, 0, st>>>(bf, hw.rc_w1, hw.rc_b1, img, n, hw.qn_cin, hw.qn_t, W, act);
    else
        rawcnn_conv1_kernel<<<ew_grid(n * H * W * (C / 4), sm_count), 256, 0, st>>>(bf, hw.rc_w1, hw.rc_b1, img, n, hw.qn_cin, hw.qn_t, W, C, act);
    if ((rc = done())) return rc;
    for (int j = 0; j < 3; ++j) {
        const HeadWeights::RcLayer& L = hw.rc[j];
        const bool last = j == 2;
        const int Hp = L.Hin + 2, Wp = L.Win + 2;
        const int Hp2 = last ? L.Hout : L.Hout + 2, Wp2 = last ? L.Wout : L.Wout + 2;
        float* out = p;
        p += (size_t)n * Hp2 * Wp2 * L.cout;
        if (!last) {
            zero_border_kernel<<<ew_grid(n * 2 * (Hp2 + Wp2) * (L.cout / 4), sm_count), 256, 0, st>>>(out, n, Hp2, Wp2, L.cout, L.Hout, L.Wout);
            if ((rc = done())) return rc;
        }
        const long long rpw = (long long)L.Hout * L.Wout, rows = n * rpw;
        const KcView av{rpw, (long long)Hp * Wp * L.cin, (long long)L.s * L.cin, 0, L.Wout, (long long)L.s * Wp * L.cin};
        const KcView ov{rpw, (long long)Hp2 * Wp2 * L.cout, L.cout, 0, L.Wout, (long long)Wp2 * L.cout};
        const KcLaunch kl = rowgemm_kc_launch(rows, L.Npad, sm_count);
        rowgemm_kc_umma_kernel<true><<<kl.grid, kKcBlock, kl.smem, st>>>(
            img, av, KcSegs{3 * L.cin, (long long)Wp * L.cin, 9 * L.cin}, L.K, L.wq, L.b, nullptr,
            out + (last ? 0 : (size_t)(Wp2 + 1) * L.cout), ov, rows, L.Npad, L.cout, act + 1, kl.ring);
        if ((rc = done())) return rc;
        img = out;
        C = L.cout; H = L.Hout; W = L.Wout;
    }
    bc_gap_kernel<<<ew_grid(n * (C / 4), sm_count), 256, 0, st>>>(img, feat, n, H * W, C);
    return done();
}

// ------------------------------------------------------------------------------ launches
inline int launch_head_stage_a(const HeadWeights& hw, const FrontendTables<double>& tab, int act, int sm_count,
                               WindowSource pcm, long long n, float* feat, float* scratch, float* mel_dump,
                               cudaStream_t st, int64_t* launches, std::string* err, bool mel_ready = false,
                               const float* conv2_nhwc = nullptr /* CRNN: conv1 + conv2 output of cnn2_stage_kernel */,
                               MelRingRef ring = MelRingRef{nullptr, nullptr, 0} /* TCN cone in stream mode */) {
    // mel_ready: the caller already placed the (n, F, T) log-mel at the start of `scratch` (stream mode)
    float* p = scratch;
    auto take = [&](size_t floats_per_window) {
        float* r = p;
        p += floats_per_window * (size_t)n;
        return r;
    };
    auto done = [&]() -> int {
        (*launches)++;
        NWW_HCUDA(cudaGetLastError());
        return NWW_OK;
    };
    int rc;
    if (hw.arch == NWW_ARCH_E2E_QUARTZNET || hw.arch == NWW_ARCH_E2E_CNN) {
        if (mel_dump != nullptr) { *err = "a raw-audio model has no log-mel to return"; return NWW_EINVAL; }
        return hw.arch == NWW_ARCH_E2E_CNN ? launch_raw_cnn(hw, act, sm_count, pcm, n, feat, scratch, st, launches, err)
                                           : launch_raw_quartznet(hw, sm_count, pcm, n, feat, scratch, st, launches, err);
    }
    if (hw.arch == NWW_ARCH_E2E_MELCNN) {
        using G = GeoREF64x101;
        float* mel = take(64 * 101);
        float* a1 = take(16 * 32 * 50);
        float* a2 = take(32 * 16 * 25);
        float* a3 = take(64 * 16 * 25);
        if ((rc = launch_frontend_f64<G>(tab, sm_count, pcm, n, mel, 0, st, launches, err))) return rc;
        if (hw.e2e_wq[1] && hw.e2e_wq[2]) {
            // channel-last pipeline: conv1 (FP32, pool) -> conv2 (tcgen05, pool) -> conv3 (tcgen05) -> avg pool
            const int grid = (int)std::min<long long>(n, sm_count);
            if (hw.e2e_plan[1].f1_hm) {
                // conv1 inside conv2's loader: the (16, 32, 50) activations never exist in HBM
                const ConvUmmaPlan& P = hw.e2e_plan[1];
                NWW_HCUDA(conv3x3_umma_launch<true>(act, grid, st, mel, hw.e2e_wq[1], hw.e2e_conv[1].b, a2, n, P, hw.e2e_conv[0].w, hw.e2e_conv[0].b));
                if ((rc = done())) return rc;
            } else {
                bc_init_conv_kernel<<<ew_grid(n * 2 * 32 * 50, sm_count), 256, 0, st>>>(mel, hw.e2e_conv[0].w, hw.e2e_conv[0].b, a1, n, 64, 101, 16, act);
                if ((rc = done())) return rc;
            }
            for (int j = hw.e2e_plan[1].f1_hm ? 2 : 1; j <= 2; ++j) {
                const ConvUmmaPlan& P = hw.e2e_plan[j];
                NWW_HCUDA(conv3x3_umma_launch<false>(act, grid, st, j == 1 ? a1 : a2, hw.e2e_wq[j], hw.e2e_conv[j].b, j == 1 ? a2 : a3, n, P));
                if ((rc = done())) return rc;
            }
            avgpool_row_nhwc_kernel<25, 4><<<ew_grid(n * 16, sm_count), 256, 0, st>>>(a3, feat, n, 64, 16);
            if ((rc = done())) return rc;
            if (mel_dump) NWW_HCUDA(cudaMemcpyAsync(mel_dump, mel, (size_t)n * 64 * 101 * sizeof(float), cudaMemcpyDeviceToDevice, st));
            return NWW_OK;
        }
        conv3x3_kernel<true><<<ew_grid(n * 2 * 32 * 50, sm_count), 256, 0, st>>>(mel, hw.e2e_conv[0].w, hw.e2e_conv[0].b, a1, n, 1, 16, 64, 101, act);
        if ((rc = done())) return rc;
        conv3x3_kernel<true><<<ew_grid(n * 4 * 16 * 25, sm_count), 256, 0, st>>>(a1, hw.e2e_conv[1].w, hw.e2e_conv[1].b, a2, n, 16, 32, 32, 50, act);
        if ((rc = done())) return rc;
        conv3x3_kernel<false><<<ew_grid(n * 8 * 16 * 25, sm_count), 256, 0, st>>>(a2, hw.e2e_conv[2].w, hw.e2e_conv[2].b, a3, n, 32, 64, 16, 25, act);
        if ((rc = done())) return rc;
        avgpool_row_kernel<<<ew_grid(n * 256, sm_count), 256, 0, st>>>(a3, feat, n, 64, 16, 25, 4);
        if ((rc = done())) return rc;
        if (mel_dump) NWW_HCUDA(cudaMemcpyAsync(mel_dump, mel, (size_t)n * 64 * 101 * sizeof(float), cudaMemcpyDeviceToDevice, st));
        return NWW_OK;
    }
    using G = GeoNS40x98;
    constexpr int F = G::N_MELS, T = G::N_FRAMES;
    float* mel = take((size_t)F * T);
    if (hw.arch == NWW_ARCH_TCN && hw.tcn_cone) {
        // time-major log-mel of the cone's frames only (or, in stream mode, gathered by the caller), then one kernel
        const TcnConeParams& P = hw.tcn_plan;
        const int frame_lo = (T - P.n_in) & ~1;
        if (!mel_ready && (rc = launch_frontend_f64<G>(tab, sm_count, pcm, n, mel, 1, st, launches, err, frame_lo))) return rc;
        if (mel_dump && (rc = launch_frontend_f64<G>(tab, sm_count, pcm, n, mel_dump, 0, st, launches, err))) return rc;
        if (hw.tcn_rows && hw.tcn_fused && hw.tcn_n_row_layers <= kTcnFusedMaxLayers && P.n_in <= kMelTailMax) {
            // ONE cooperative launch: [gather of the cone's frames out of the mel ring] + all layers, grid barriers in between
            float* act = take((size_t)hw.tcn_rows_pw);
            TcnFusedParams FP{};
            FP.n_layers = hw.tcn_n_row_layers;
            FP.ring = ring;
            FP.n = n;
            FP.mel = mel;
            FP.t0 = T - P.n_in;
            FP.n_tail = P.n_in;
            for (int li = 0; li < hw.tcn_n_row_layers; ++li) {
                const HeadWeights::TcnRowLayer& L = hw.tcn_row[li];
                const long long a_ws = L.a_in_mel ? (long long)F * T : hw.tcn_rows_pw, o_ws = L.o_in_feat ? L.N : hw.tcn_rows_pw;
                TcnFusedLayer& D = FP.L[li];
                D.A = (L.a_in_mel ? mel : act) + L.a_off;
                D.av = kc_seq(L.n_pos, a_ws, L.a_row_stride, 0);
                D.k_valid = L.k_valid; D.K = L.K; D.wq = L.wq; D.bias = L.bias;
                D.res = L.has_res ? act + L.r_off : nullptr;
                D.rv = kc_seq(L.n_pos, hw.tcn_rows_pw, L.r_row_stride, 0);
                D.out = (L.o_in_feat ? feat : act) + L.o_off;
                D.ov = kc_seq(L.n_pos, o_ws, L.N, 0);
                D.rows = n * L.n_pos; D.N = L.N; D.act = L.act; D.pre_relu = L.pre_relu;
            }
            const size_t smem = rowgemm_kc_smem_bytes(2);
            NWW_HCUDA(set_smem(tcn_rows_fused_kernel, smem));
            NWW_HCUDA(cudaFuncSetAttribute(tcn_rows_fused_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            int per_sm = 0;
            NWW_HCUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tcn_rows_fused_kernel, kKcBlock, smem));
            if (per_sm >= 1) {
                const long long tiles0 = (n * hw.tcn_row[0].n_pos + kKcRows - 1) / kKcRows;
                const int grid = (int)std::min<long long>(std::max<long long>(tiles0, 1), (long long)sm_count * std::min(per_sm, 2));
                void* args[] = {(void*)&FP};
                NWW_HCUDA(cudaLaunchCooperativeKernel((const void*)tcn_rows_fused_kernel, dim3(grid), dim3(kKcBlock), args, smem, st));
                return done();
            }
            p -= (size_t)hw.tcn_rows_pw * (size_t)n;        // no co-residency: fall through to one launch per layer
        }
        // Small launches take the per-tile cone kernel (one launch, weights re-read per 4 windows: 86 us for one window against
        // 150 us for the eight row-GEMM launches, 256 vs 335 us for 256); from a couple of thousand windows on the row GEMMs over
        // the whole launch group win.  Same operand splits and accumulation order: the two routes are bit-identical.
        if (hw.tcn_rows && (n >= kTcnRowsMinWindows || !hw.tcn_umma)) {
            // stream mode: the cone's frames out of the mel ring, time-major, into the same place the front end writes them
            if (ring.ring != nullptr) {
                if (P.n_in > kMelTailMax) { *err = "tcn: dependency cone longer than 32 frames is not built for stream mode"; return NWW_EUNSUPPORTED; }
                stream_mel_tail_kernel<<<(int)std::min<long long>((n + kMelTailWarps - 1) / kMelTailWarps, (long long)sm_count * 8),
                                         kMelTailWarps * 32, 0, st>>>(ring, n, mel, T - P.n_in, P.n_in);
                if ((rc = done())) return rc;
            }
            float* act = take((size_t)hw.tcn_rows_pw);
            NWW_HCUDA(set_smem(rowgemm_kc_umma_kernel<true>, rowgemm_kc_smem_bytes()));
            NWW_HCUDA(cudaFuncSetAttribute(rowgemm_kc_umma_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            for (int li = 0; li < hw.tcn_n_row_layers; ++li) {
                const HeadWeights::TcnRowLayer& L = hw.tcn_row[li];
                const long long rows = n * L.n_pos;
                const long long a_ws = L.a_in_mel ? (long long)F * T : hw.tcn_rows_pw, o_ws = L.o_in_feat ? L.N : hw.tcn_rows_pw;
                const float* A = (L.a_in_mel ? mel : act) + L.a_off;
                float* O = (L.o_in_feat ? feat : act) + L.o_off;
                const KcLaunch kl = rowgemm_kc_launch(rows, L.N, sm_count);
                rowgemm_kc_umma_kernel<true><<<kl.grid, kKcBlock, kl.smem, st>>>(
                    A, kc_seq(L.n_pos, a_ws, L.a_row_stride, 0), kc_one_seg(L.k_valid), L.K, L.wq, L.bias,
                    L.has_res ? act + L.r_off : nullptr, O, kc_seq(L.n_pos, o_ws, L.N, 0), rows, L.N, L.N, L.act, kl.ring,
                    kc_seq(L.n_pos, hw.tcn_rows_pw, L.r_row_stride, 0), L.pre_relu);
                if ((rc = done())) return rc;
            }
            return NWW_OK;
        }
        if (hw.tcn_umma) {
            const size_t usm = tcn_umma_smem_bytes(hw.tcn_uplan);
            NWW_HCUDA(set_smem(tcn_cone_umma_kernel, usm));
            const long long utiles = (n + kTuWT - 1) / kTuWT;
            tcn_cone_umma_kernel<<<(int)std::min<long long>(utiles, sm_count), kTuNT, usm, st>>>(mel, (long long)F * T, ring, n,
                                                                                              hw.tcn_uplan, hw.tcn_wq, feat);
            return done();
        }
        const size_t smem = tcn_cone_smem_bytes(P);
        NWW_HCUDA(set_smem(tcn_cone_kernel, smem));
        const long long tiles = (n + P.wt - 1) / P.wt;
        tcn_cone_kernel<<<(int)std::min<long long>(tiles, sm_count), kTcnNT, smem, st>>>(mel, (long long)F * T, ring, n, P, feat);
        return done();
    }
    if (hw.arch == NWW_ARCH_GRU || hw.arch == NWW_ARCH_LSTM) {
        // time-major log-mel (or, in stream mode, gathered that way by the caller), then the whole sequence in one kernel
        if (!mel_ready && (rc = launch_frontend_f64<G>(tab, sm_count, pcm, n, mel, 1, st, launches, err))) return rc;
        if (mel_dump && (rc = launch_frontend_f64<G>(tab, sm_count, pcm, n, mel_dump, 0, st, launches, err))) return rc;
        // rows per CTA: a full 128-row tile when there is a wave of them, fewer when the batch is small
        const int tm = (int)std::min<long long>(kRnnTM, std::max<long long>(32, ((n + sm_count - 1) / sm_count + 31) / 32 * 32));
        const int grid = (int)std::min<long long>((n + tm - 1) / tm, sm_count);
        auto go = [&](auto kernel, size_t smem) -> int {
            NWW_HCUDA(set_smem(kernel, smem));
            kernel<<<grid, kRnnNT, smem, st>>>(mel, (long long)F * T, T, n, tm, hw.rnn_wq_f, hw.rnn_wq_b, feat);
            return done();
        };
        if (hw.rnn_cell == RNN_GRU)
            return hw.rnn_hidden == 128 ? go(rnn_seq_kernel<RNN_GRU, 128, F>, RnnDims<128, F>::SMEM)
                                        : go(rnn_seq_kernel<RNN_GRU, 64, F>, RnnDims<64, F>::SMEM);
        return hw.rnn_hidden == 128 ? go(rnn_seq_kernel<RNN_LSTM, 128, F>, RnnDims<128, F>::SMEM)
                                    : go(rnn_seq_kernel<RNN_LSTM, 64, F>, RnnDims<64, F>::SMEM);
    }
    if (hw.arch == NWW_ARCH_QUARTZNET) {
        if (!mel_ready && (rc = launch_frontend_f64<G>(tab, sm_count, pcm, n, mel, 1, st, launches, err))) return rc;
        if (mel_dump && (rc = launch_frontend_f64<G>(tab, sm_count, pcm, n, mel_dump, 0, st, launches, err))) return rc;
        return launch_quartznet_blocks(hw, mel, F, n, p, feat, sm_count, st, launches, err);
    }
    // BcResNet from int16 PCM: front end + init conv in one stage kernel (the log-mel stays in shared memory)
    const bool bc_stage = hw.arch == NWW_ARCH_BCRESNET && hw.bc_stage && !mel_ready && pcm.fbase == nullptr;
    if (!bc_stage) {
        if (!mel_ready && !conv2_nhwc && (rc = launch_frontend_f64<G>(tab, sm_count, pcm, n, mel, 0, st, launches, err))) return rc;
        if (mel_dump && !conv2_nhwc) NWW_HCUDA(cudaMemcpyAsync(mel_dump, mel, (size_t)n * F * T * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }

    if (hw.arch == NWW_ARCH_TCN) {
        // positions each layer output must cover so that the final step (T-1) is exact
        const int L = hw.tcn_levels, k = hw.tcn_k;
        int lo_out[8], lo_mid[8];
        int need_lo = T - 1;
        for (int i = L - 1; i >= 0; --i) {
            const int d = 1 << i;
            lo_out[i] = need_lo;                                  // block output (conv2 + residual)
            lo_mid[i] = std::max(0, lo_out[i] - (k - 1) * d);     // conv1 output
            need_lo = std::max(0, lo_mid[i] - (k - 1) * d);       // block input
        }
        size_t maxc = hw.tcn_in;
        for (int i = 0; i < L; ++i) maxc = std::max<size_t>(maxc, hw.tcn_ch[i]);
        float* buf[3] = {take(maxc * T), take(maxc * T), take(maxc * T)};
        const float* x = mel;                                     // (B, F, T) == permute(0, 2, 1) of the (T, F) input
        int cin = hw.tcn_in, cur = 0;
        for (int i = 0; i < L; ++i) {
            const int c = hw.tcn_ch[i], d = 1 << i;
            float* mid = buf[cur];
            float* out = buf[(cur + 1) % 3];
            tcn_conv_kernel<<<ew_grid(n * (T - lo_mid[i]) * c, sm_count), 256, 0, st>>>(
                x, hw.tcn_c1[i].w, hw.tcn_c1[i].b, nullptr, nullptr, nullptr, mid, n, cin, c, 0, T, k, d, lo_mid[i], 0);
            if ((rc = done())) return rc;
            tcn_conv_kernel<<<ew_grid(n * (T - lo_out[i]) * c, sm_count), 256, 0, st>>>(
                mid, hw.tcn_c2[i].w, hw.tcn_c2[i].b, x, hw.tcn_down[i].w, hw.tcn_down[i].b, out, n, c, c, cin, T, k, d, lo_out[i], 1);
            if ((rc = done())) return rc;
            x = out;
            cin = c;
            cur = (cur + 2) % 3;       // next block must not overwrite its own input (= out)
        }
        last_step_kernel<<<ew_grid(n * cin, sm_count), 256, 0, st>>>(x, feat, n, cin, T);
        return done();
    }
    if (hw.arch == NWW_ARCH_BCRESNET) {
        // channel-last pipeline (nww_bc.cuh): init conv -> 3 x (depthwise + centre tap, pointwise/shortcut row GEMMs) -> GAP
        float* a0 = take(32 * 20 * 49);
        if (bc_stage) {
            auto k = act == ACT_RELU ? bc_stage_kernel<ACT_RELU> : act == ACT_GELU ? bc_stage_kernel<ACT_GELU> : bc_stage_kernel<ACT_SILU>;
            NWW_HCUDA(set_smem(k, BcStage::kTotal));
            k<<<(int)std::min<long long>(n, sm_count), BcStage::NT, BcStage::kTotal, st>>>(pcm, n, tab, hw.bc_init.w, hw.bc_init.b, a0, mel_dump);
        } else {
            bc_init_conv_kernel<<<ew_grid(n * 4 * 20 * 49, sm_count), 256, 0, st>>>(mel, hw.bc_init.w, hw.bc_init.b, a0, n, F, T, 32, act);
        }
        if ((rc = done())) return rc;
        const int ch[4] = {32, 64, 128, 256};
        const int sh[3] = {2, 2, 2}, sw[3] = {2, 2, 1};
        int H = 20, W = 49;
        const float* x = a0;
        for (int j = 0; j < 3; ++j) {
            const int Ho = (H - 1) / sh[j] + 1, Wo = (W - 1) / sw[j] + 1;
            float* dwo = take((size_t)ch[j] * Ho * Wo);
            float* ctr = take((size_t)ch[j] * Ho * Wo);
            float* out = take((size_t)ch[j + 1] * Ho * Wo);
            bc_dw_kernel<<<ew_grid(n * (ch[j] / 4) * Ho * Wo, sm_count), 256, 0, st>>>(x, hw.bc_dw[j], dwo, ctr, n, ch[j], H, W, sh[j], sw[j]);
            if ((rc = done())) return rc;
            const long long rows = n * Ho * Wo;
            if (hw.bc_wq[j] != nullptr) {
                const size_t smem = bcu_smem_bytes(ch[j]);
                NWW_HCUDA(set_smem(bc_block_umma_kernel, smem));
                const long long tiles = (rows + kBcuRows - 1) / kBcuRows;
                // two CTAs per SM where they fit (Cin <= 64: <= 98 KB of shared memory, 102 registers, 128 TMEM columns each):
                // one CTA's operand conversion / epilogue runs under the other's MMAs
                const int per_sm = smem <= 70 * 1024 ? 3 : smem <= 110 * 1024 ? 2 : 1;
                bc_block_umma_kernel<<<(int)std::min<long long>(tiles, (long long)sm_count * per_sm), kBcuNT, smem, st>>>(
                    dwo, ctr, hw.bc_wq[j], hw.bc_pw[j].b, hw.bc_sc[j].b, out, rows, ch[j], ch[j + 1], act);
            } else {
                const size_t smem = bc_block_smem_bytes(ch[j]);
                NWW_HCUDA(set_smem(bc_block_gemm_kernel, smem));
                const long long tiles = (rows + kBcRows - 1) / kBcRows;
                bc_block_gemm_kernel<<<(int)std::min<long long>(tiles, (long long)sm_count), kTcnNT, smem, st>>>(
                    dwo, ctr, hw.bc_pw[j].w, hw.bc_pw[j].b, hw.bc_sc[j].w, hw.bc_sc[j].b, out, rows, ch[j], ch[j + 1], act);
            }
            if ((rc = done())) return rc;
            x = out; H = Ho; W = Wo;
        }
        bc_gap_kernel<<<ew_grid(n * 64, sm_count), 256, 0, st>>>(x, feat, n, H * W, 256);
        return done();
    }
    if (hw.arch == NWW_ARCH_CRNN_GRU) {
        int cin = 1, H = F, W = T;
        const float* x = mel;
        const int S0 = 12, In0 = hw.gru_in;
        float* seq_direct = nullptr;
        if (conv2_nhwc != nullptr) {
            // conv1 + conv2 already done by cnn2_stage_kernel (channel-last, 10 x 24 x 32): third conv -> sequence
            seq_direct = take((size_t)S0 * In0);
            if (hw.crnn_wq3 != nullptr) {
                // third conv on tcgen05 (nww_conv_umma.cuh), then the (c, h) -> feature repack
                float* a3 = take((size_t)S0 * In0);
                const ConvUmmaPlan& P = hw.crnn_plan3;
                NWW_HCUDA(conv3x3_umma_launch<false>(act, (int)std::min<long long>(n, sm_count), st, conv2_nhwc, hw.crnn_wq3, hw.crnn_conv[2].b, a3, n, P));
                if ((rc = done())) return rc;
                seq_pack_nhwc_kernel<<<ew_grid(n * S0 * In0, sm_count), 256, 0, st>>>(a3, seq_direct, n, hw.crnn_ch[2], 5, 12);
                if ((rc = done())) return rc;
            } else {
                crnn_conv3_seq_kernel<<<ew_grid(n * 5 * 12 * (hw.crnn_ch[2] / kOCT), sm_count), 256, 0, st>>>(
                    conv2_nhwc, hw.crnn_conv[2].w, hw.crnn_conv[2].b, seq_direct, n, hw.crnn_ch[1], hw.crnn_ch[2], 10, 24, act);
                if ((rc = done())) return rc;
            }
            cin = hw.crnn_ch[2]; H = 5; W = 12;
        }
        for (int i = 0; i < hw.crnn_levels && conv2_nhwc == nullptr; ++i) {
            const int c = hw.crnn_ch[i];
            float* out = take((size_t)c * (H / 2) * (W / 2));
            conv3x3_kernel<true><<<ew_grid(n * (c / kOCT) * (H / 2) * (W / 2), sm_count), 256, 0, st>>>(
                x, hw.crnn_conv[i].w, hw.crnn_conv[i].b, out, n, cin, c, H, W, act);
            if ((rc = done())) return rc;
            x = out; cin = c; H /= 2; W /= 2;
        }
        const int S = W, In = hw.gru_in, Hd = hw.gru_hidden, G3 = 3 * Hd;
        float* seq = seq_direct ? seq_direct : take((size_t)S * In);
        float* gi_f = take((size_t)S * G3);
        float* gi_b = take((size_t)G3);
        if (!seq_direct) {
            seq_pack_kernel<<<ew_grid(n * S * In, sm_count), 256, 0, st>>>(x, seq, n, cin, H, S);
            if ((rc = done())) return rc;
        }
        // input projections over rows (all S steps forward, the last step only backward), then the recurrence
        if (hw.gru_wih_f_kn && hw.gru_wih_b_kn) {
            const long long rows = n * S;
            if (hw.gru_wih_f_q && hw.gru_wih_b_q) {
                const size_t usm = rowgemm_umma_smem_bytes(In);
                NWW_HCUDA(set_smem(rowgemm_umma_kernel, usm));
                rowgemm_umma_kernel<<<(int)std::min<long long>((rows + kRuRows - 1) / kRuRows, sm_count), kRuNT, usm, st>>>(
                    seq, 1, 0, hw.gru_wih_f_q, hw.gru_bih_f, gi_f, rows, In, G3);
                if ((rc = done())) return rc;
                rowgemm_umma_kernel<<<(int)std::min<long long>((n + kRuRows - 1) / kRuRows, sm_count), kRuNT, usm, st>>>(
                    seq, S, S - 1, hw.gru_wih_b_q, hw.gru_bih_b, gi_b, n, In, G3);
                if ((rc = done())) return rc;
            } else {
                const size_t rsm = rowgemm_smem_bytes(In);
                NWW_HCUDA(set_smem(rowgemm_kernel, rsm));
                rowgemm_kernel<<<(int)std::min<long long>((rows + kRgRows - 1) / kRgRows, sm_count), kTcnNT, rsm, st>>>(
                    seq, 1, 0, hw.gru_wih_f_kn, hw.gru_bih_f, gi_f, rows, In, G3);
                if ((rc = done())) return rc;
                rowgemm_kernel<<<(int)std::min<long long>((n + kRgRows - 1) / kRgRows, sm_count), kTcnNT, rsm, st>>>(
                    seq, S, S - 1, hw.gru_wih_b_kn, hw.gru_bih_b, gi_b, n, In, G3);
                if ((rc = done())) return rc;
            }
            if (hw.gru_whh_q != nullptr && Hd == kGruTcH) {
                const size_t gsm = gru_tc_smem_bytes();
                NWW_HCUDA(set_smem(gru_tc_kernel, gsm));
                gru_tc_kernel<<<(int)std::min<long long>((n + kGruTcTM - 1) / kGruTcTM, (long long)sm_count), kGruTcNT, gsm, st>>>(
                    gi_f, gi_b, hw.gru_whh_q, hw.gru_bhh_f, hw.gru_bhh_b, feat, n, S);
                return done();
            }
            const size_t gsm = gru2_smem_bytes(Hd);
            NWW_HCUDA(set_smem(gru2_kernel, gsm));
            gru2_kernel<<<(int)std::min<long long>((n + kGru2TM - 1) / kGru2TM, (long long)sm_count), kTcnNT, gsm, st>>>(
                gi_f, gi_b, hw.gru_whh_f, hw.gru_bhh_f, hw.gru_bhh_b, feat, n, S, Hd);
            return done();
        }
        TailParams P{};
        P.n_layers = 1; P.act = act; P.max_width = G3; P.raw_out = 1;
        P.layers[0] = TailLayer{hw.gru_wih_f_nk, hw.gru_bih_f, nullptr, nullptr, In, G3, POST_NONE};
        const size_t smem = tail_smem_bytes(G3);
        NWW_HCUDA(set_smem(tail_kernel, smem));
        const long long rows = n * S;
        tail_kernel<<<(int)std::min<long long>((rows + kTailTM - 1) / kTailTM, (long long)sm_count * 2), kTailNT, smem, st>>>(
            seq, rows, P, gi_f, nullptr, nullptr);
        if ((rc = done())) return rc;
        P.layers[0] = TailLayer{hw.gru_wih_b_nk, hw.gru_bih_b, nullptr, nullptr, In, G3, POST_NONE};
        P.x_row_mul = S; P.x_row_off = S - 1;
        tail_kernel<<<(int)std::min<long long>((n + kTailTM - 1) / kTailTM, (long long)sm_count * 2), kTailNT, smem, st>>>(
            seq, n, P, gi_b, nullptr, nullptr);
        if ((rc = done())) return rc;
        const size_t gsmem = gru_smem_bytes(Hd);
        NWW_HCUDA(set_smem(gru_kernel, gsmem));
        gru_kernel<<<(int)std::min<long long>((n + kGruTM - 1) / kGruTM, (long long)sm_count), kGruNT, gsmem, st>>>(
            gi_f, gi_b, hw.gru_whh_f, hw.gru_bhh_f, hw.gru_bhh_b, feat, n, S, Hd);
        return done();
    }
    *err = "architecture id " + std::to_string(hw.arch) + " is not built into this library";
    return NWW_EUNSUPPORTED;
}

}  // namespace nww
