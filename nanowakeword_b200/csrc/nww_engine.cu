// nww_engine.cu — engine object and the C ABI of libnwwb200.so (see include/nww_b200.h).
//
// Host-side responsibilities: parse the weight blob, build the front-end tables, keep
// weights/tables/workspaces resident in HBM, size grids from the SM count, and enqueue
// stage A (per-window fused kernels) + stage B (dense tail) in L2-sized chunks.
#include <cuda_runtime.h>
#include <stdio.h>
#include <algorithm>
#include <memory>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/nww_b200.h"
#include "nww_blob.h"
#include "nww_cnn.cuh"
#include "nww_cnn2.cuh"
#include "nww_cnn3.cuh"
#include "nww_cnn4.cuh"
#include "nww_gemm_tc.cuh"
#include "nww_heads.cuh"
#include "nww_stage.cuh"
#include "nww_stream.cuh"
#include "nww_stream_mel.cuh"
#include "nww_tables.h"
#include "nww_tail.cuh"

using namespace nww;

static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
#define NWW_CUDA(call)                                                                                 \
    do {                                                                                               \
        cudaError_t _e = (call);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return fail(NWW_ECUDA, std::string(#call) + ": " + cudaGetErrorString(_e));                \
    } while (0)

namespace {

struct DeviceArena {                // one allocation for all small constant tables
    std::vector<unsigned char> host;
    unsigned char* dev = nullptr;
    size_t add(const void* p, size_t bytes) {
        size_t off = (host.size() + 255) & ~(size_t)255;
        host.resize(off + bytes);
        memcpy(host.data() + off, p, bytes);
        return off;
    }
};

}  // namespace

struct nww_engine {
    int device = 0;
    int sm_count = 0;
    nww_spec spec{};
    std::mutex mu;

    Blob blob;                       // parsed view over host copy
    std::vector<unsigned char> blob_host;
    unsigned char* d_blob = nullptr;
    DeviceArena arena;
    FrontendTables<float> tab32{};
    FrontendTables<double> tab64{};

    int n_mels = 0, n_frames = 0, clip = 0;
    int feat_dim = 0, emb_dim = 0;
    TailParams tail{};
    CnnWeights cnn{};
    HeadWeights heads{};

    // tensor-core path for the first (wide) dense layer
    bool tc_enabled = false;
    float *d_w_hi = nullptr, *d_w_lo = nullptr;      // [N][K] split weights
    float *d_feat_hi = nullptr, *d_feat_lo = nullptr; // [chunk][K] split feature rows
    float* d_part = nullptr;                          // [kTcMaxSplits][rows][N] split-K partial sums of that layer
    int tc_rows = 0;                                  // rows allocated per operand / partial slab (multiple of 128)
    int tc_kp = 0;                                    // K of that layer padded to a multiple of the 32-float K tile
    // CNN stage v2 (tcgen05 conv2): conv2 weights as UMMA operands; features written pre-split by the stage kernel
    bool cnn2_enabled = false;
    bool crnn_cnn2 = false;                           // CRNN head: conv1 + conv2 through cnn2_stage_kernel
    float* d_nhwc = nullptr;                          // [chunk][7680] channel-last conv2 output for that path
    uint4* d_w2_umma = nullptr;
    void* d_bc_wq[3] = {nullptr, nullptr, nullptr};   // BcResNet 1x1 weights as bf16 UMMA operands
    std::vector<void*> d_extra;                       // further pre-split weight buffers (QuartzNet blocks)
    void* d_conv_wq[3] = {nullptr, nullptr, nullptr}; // 3x3 conv weights as bf16 UMMA operands (E2E mel-CNN, CRNN conv3)
    Cnn2Weights cnn2{};
    CUtensorMap tm_xhi{}, tm_xlo{}, tm_whi{}, tm_wlo{};
    TailParams tail_rest{};                           // layers 1.. (after the tensor-core layer)

    int chunk = 0;
    int split_per_sm = 4;            // windows per SM and sub-chunk of the split CNN stage (reserved[1] overrides)
    float* d_feat = nullptr;         // [chunk][feat_dim]
    float* d_scratch = nullptr;      // per-head scratch (e.g. CRNN sequence), may be null
    size_t scratch_per_window = 0;

    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    int16_t* d_pcm[2] = {nullptr, nullptr};
    float* d_scores = nullptr;
    int64_t host_chunk = 0, scores_cap = 0;

    int64_t launches = 0, windows = 0;

    // Cross-stream ordering of the shared workspaces (d_feat, d_scratch, d_part, d_pcm, the stream rings): every
    // enqueueing entry point makes its stream wait for the previous call's last event and records a new one, so calls
    // on different CUDA streams are ordered like calls on one stream (the mutex only serialises the enqueueing).
    cudaEvent_t ev_last = nullptr;
    bool last_valid = false;
    cudaStream_t last_stream = nullptr;
    float* d_melf = nullptr;             // [chunk][F][T] log-mel of float feeds for the stage kernels that start from mel
    // split CNN stage (nww_cnn4.cuh): the front-end kernel runs on its own stream beside the convolution kernel
    cudaStream_t fe_stream = nullptr;
    cudaEvent_t ev_piece[8] = {};            // nww_stream_push_host: piece p of the bank has arrived (recorded on copy_stream)
    cudaEvent_t ev_fork = nullptr;
    std::vector<cudaEvent_t> ev_fe;
    // selective push (nww_stream_push_select): the streams to score, their window offsets and compact scores
    const long long* sel_ids = nullptr;  // non-null only while a selective push is being enqueued
    long long* d_sel_ids = nullptr;      // staging for the host variant
    long long* d_sel_off = nullptr;
    float* d_sel_scores = nullptr;
    int64_t sel_cap = 0;

    // multi-stream mode (nww_stream_*): mirrored int16 rings in HBM
    StreamState streams{};
    float* d_mel_ring = nullptr;         // [n_streams][40][2 x 98] incremental log-mel (NS40x98 only)
    bool mel_inc = false;                // every chunk since open / full reset was a multiple of the hop
    int16_t* d_chunk = nullptr;          // staging for nww_stream_push_host
    size_t chunk_cap = 0;
    long long* d_ids = nullptr;
    int64_t ids_cap = 0;

    // optional per-stage event timing
    bool profiling = false;
    struct Span { cudaEvent_t a, b; int stage; int64_t windows; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t get_event() {
        if (!event_pool.empty()) { cudaEvent_t ev = event_pool.back(); event_pool.pop_back(); return ev; }
        cudaEvent_t ev = nullptr;
        cudaEventCreate(&ev);
        return ev;
    }

    const float* dptr(const std::string& name) const {
        const BlobTensor* t = blob.find(name);
        return t ? reinterpret_cast<const float*>(d_blob + t->offset) : nullptr;
    }
    ~nww_engine();                       // frees every device buffer: a failing nww_create leaks nothing
};

// Order this call after whatever the previous call enqueued (possibly on another stream).
static cudaError_t order_enter(nww_engine* e, cudaStream_t st) {
    if (e->last_valid && e->last_stream != st) return cudaStreamWaitEvent(st, e->ev_last, 0);
    return cudaSuccess;
}
static cudaError_t order_leave(nww_engine* e, cudaStream_t st) {
    if (!e->ev_last) {
        cudaError_t ce = cudaEventCreateWithFlags(&e->ev_last, cudaEventDisableTiming);
        if (ce != cudaSuccess) return ce;
    }
    e->last_valid = true;
    e->last_stream = st;
    return cudaEventRecord(e->ev_last, st);
}

// ------------------------------------------------------------------------------ helpers
template <typename T>
static void fill_tables(nww_engine* e, const HostFrontendTables& h, FrontendTables<T>* out_offsets_as_ptrs) {
    std::vector<T> ws(h.window_scaled.begin(), h.window_scaled.end());
    std::vector<T> wu(h.window_unscaled.begin(), h.window_unscaled.end());
    std::vector<cplx<T>> tw(h.n_fft);
    for (int i = 0; i < h.n_fft; ++i) tw[i] = {(T)h.tw_re[i], (T)h.tw_im[i]};
    FrontendTables<T>& t = *out_offsets_as_ptrs;
    // store offsets in the pointer fields first; rebased after the arena is uploaded
    t.window = reinterpret_cast<const T*>(e->arena.add(ws.data(), ws.size() * sizeof(T)));
    t.window_unscaled = reinterpret_cast<const T*>(e->arena.add(wu.data(), wu.size() * sizeof(T)));
    t.twiddle = reinterpret_cast<const cplx<T>*>(e->arena.add(tw.data(), tw.size() * sizeof(cplx<T>)));
    t.binpos = reinterpret_cast<const uint16_t*>(e->arena.add(h.binpos.data(), h.binpos.size() * sizeof(uint16_t)));
    t.mel_start = reinterpret_cast<const int*>(e->arena.add(h.mel_start.data(), h.mel_start.size() * sizeof(int)));
    t.mel_count = reinterpret_cast<const int*>(e->arena.add(h.mel_count.data(), h.mel_count.size() * sizeof(int)));
    t.mel_woff = reinterpret_cast<const int*>(e->arena.add(h.mel_woff.data(), h.mel_woff.size() * sizeof(int)));
    t.mel_w = reinterpret_cast<const float*>(e->arena.add(h.mel_w.data(), h.mel_w.size() * sizeof(float)));
    t.amin = 1e-10f;
    t.floor_db = -100.0f;            // 10*log10(1e-10)
    t.mel_vec_ok = h.mel_vec_ok;
}

template <typename T> static void rebase_tables(FrontendTables<T>* t, unsigned char* base) {
    auto fix = [&](auto& p) {
        using P = std::remove_reference_t<decltype(p)>;
        p = reinterpret_cast<P>(base + reinterpret_cast<size_t>(p));
    };
    fix(t->window);
    fix(t->window_unscaled);
    fix(t->twiddle);
    fix(t->binpos);
    fix(t->mel_start);
    fix(t->mel_count);
    fix(t->mel_woff);
    fix(t->mel_w);
}

static int build_tail(nww_engine* e) {
    const BlobTensor* nl = e->blob.find("tail.n_layers");
    if (!nl) return fail(NWW_EINVAL, "weight blob: tail.n_layers missing");
    const int n = *reinterpret_cast<const int*>(e->blob.base + nl->offset);
    if (n < 2 || n > kMaxTailLayers) return fail(NWW_EINVAL, "weight blob: bad tail.n_layers");
    TailParams& P = e->tail;
    P.n_layers = n;
    P.act = e->spec.activation;
    P.max_width = 1;
    int prev_n = e->feat_dim;
    for (int i = 0; i < n; ++i) {
        const std::string p = "tail." + std::to_string(i);
        const BlobTensor* w = e->blob.find(p + ".W");
        const BlobTensor* b = e->blob.find(p + ".b");
        const BlobTensor* post = e->blob.find(p + ".post");
        if (!w || !b || !post || w->dims.size() != 2) return fail(NWW_EINVAL, "weight blob: " + p + " incomplete");
        TailLayer& L = P.layers[i];
        L.N = (int)w->dims[0];
        L.K = (int)w->dims[1];
        if (L.K != prev_n)
            return fail(NWW_EINVAL, p + ": input width " + std::to_string(L.K) + " does not match producer width " +
                                        std::to_string(prev_n));
        if (b->numel() != (size_t)L.N) return fail(NWW_EINVAL, p + ": bias size mismatch");
        L.W = e->dptr(p + ".W");
        L.b = e->dptr(p + ".b");
        L.ln_g = e->dptr(p + ".ln_g");
        L.ln_b = e->dptr(p + ".ln_b");
        L.post = *reinterpret_cast<const int*>(e->blob.base + post->offset);
        if (L.post == POST_LN_ACT && (!L.ln_g || !L.ln_b)) return fail(NWW_EINVAL, p + ": LayerNorm terms missing");
        P.max_width = std::max(P.max_width, L.N);
        if (i > 0) P.max_width = std::max(P.max_width, L.K);
        prev_n = L.N;
    }
    if (P.layers[n - 1].N != 1) return fail(NWW_EINVAL, "tail: last layer must have one output");
    e->emb_dim = P.layers[n - 3 >= 0 ? n - 3 : 0].N;
    if (tail_smem_bytes(P.max_width) > 200 * 1024) return fail(NWW_EUNSUPPORTED, "tail: layer too wide for shared memory");
    return NWW_OK;
}

template <typename G> static int check_geometry(const nww_spec& s) {
    if (s.n_fft != G::N_FFT || s.win_length != G::WIN || s.hop_length != G::HOP || s.n_mels != G::N_MELS ||
        (s.center != 0) != (G::CENTER != 0) || s.clip_samples != G::CLIP)
        return fail(NWW_EUNSUPPORTED,
                    "front-end parameters do not match a built-in geometry (NS40x98: 400/512/160/40 not centred; "
                    "REF64x101: 400/400/160/64 centred; clip 16000)");
    return NWW_OK;
}

template <typename G> static int setup_frontend(nww_engine* e) {
    int rc = check_geometry<G>(e->spec);
    if (rc) return rc;
    const BlobTensor* win = e->blob.find("frontend.window");
    const BlobTensor* fb = e->blob.find("frontend.fb");
    if (!win || !fb || win->numel() != (size_t)G::WIN || fb->numel() != (size_t)G::N_FREQS * G::N_MELS)
        return fail(NWW_EINVAL, "weight blob: frontend.window / frontend.fb missing or wrong shape");
    HostFrontendTables h;
    std::string err;
    const int rad[4] = {G::R0, G::R1, G::R2, G::R3};
    if (!build_frontend_tables(G::N_FFT, G::WIN, G::N_MELS, rad, G::N_PASS, e->blob.f32("frontend.window"),
                               e->blob.f32("frontend.fb"), &h, &err))
        return fail(NWW_EINVAL, err);
    fill_tables<float>(e, h, &e->tab32);
    fill_tables<double>(e, h, &e->tab64);
    e->n_mels = G::N_MELS;
    e->n_frames = G::N_FRAMES;
    e->clip = G::CLIP;
    return NWW_OK;
}


// conv weights of a (1 -> 16, 16 -> 32) 3x3 pair in the layouts cnn2_stage_kernel wants:
//   conv1 [cg][tap][8 oc] FP32 and conv2 [tap][hi|lo][kg][oc][8 ic] bf16 (un-swizzled K-major UMMA operand).
// w1_at(oc, tap) reads the first conv's weight; w2 is [ic 16][tap 9][oc 32].
template <typename W1At>
static int build_cnn2_weights(nww_engine* e, W1At w1_at, const float* w2, const float* d_b1, const float* d_b2) {
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
    std::vector<uint16_t> wb(Cnn2::W2_BYTES / 2);
    for (int tap = 0; tap < 9; ++tap)
        for (int ic = 0; ic < 16; ++ic)
            for (int oc = 0; oc < 32; ++oc) {
                const float v = w2[(ic * 9 + tap) * 32 + oc];
                const uint16_t hi = bf16_rn(v), lo = bf16_rn(v - bf16_f(hi));
                const size_t base = (size_t)tap * 2 * (Cnn2::W2_TAP_BYTES / 2) + (size_t)(ic >> 3) * 256 + oc * 8 + (ic & 7);
                wb[base] = hi;
                wb[base + Cnn2::W2_TAP_BYTES / 2] = lo;
            }
    std::vector<float> w1v(144);
    for (int oc = 0; oc < 16; ++oc)
        for (int tap = 0; tap < 9; ++tap) w1v[(oc >> 3) * 72 + tap * 8 + (oc & 7)] = w1_at(oc, tap);
    NWW_CUDA(cudaMalloc(&e->d_w2_umma, Cnn2::W2_BYTES + 144 * sizeof(float)));
    NWW_CUDA(cudaMemcpy(e->d_w2_umma, wb.data(), Cnn2::W2_BYTES, cudaMemcpyHostToDevice));
    float* d_w1v = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(e->d_w2_umma) + Cnn2::W2_BYTES);
    NWW_CUDA(cudaMemcpy(d_w1v, w1v.data(), 144 * sizeof(float), cudaMemcpyHostToDevice));
    e->cnn2 = Cnn2Weights{d_w1v, d_b1, e->d_w2_umma, d_b2};
    return NWW_OK;
}

// ------------------------------------------------------------------------------ launches
static int grid_for(const nww_engine* e, int64_t n, int per_sm = 1) {
    return (int)std::min<int64_t>(n, (int64_t)e->sm_count * per_sm);
}

template <typename G>
static int launch_frontend(nww_engine* e, WindowSource pcm, int64_t n, float* mel, int time_major, cudaStream_t st) {
    if (pcm.fbase != nullptr)            // float feed: the generic FP64 front end on float samples
        return launch_frontend_f64<G>(e->tab64, e->sm_count, pcm, n, mel, time_major, st, &e->launches, &g_last_error);
    if (e->spec.frontend_precision == NWW_FRONTEND_FP32) {
        auto k = frontend_kernel<float, G, kNfb32, kStageNT>;
        NWW_CUDA(set_smem(k, FrontendSmem<float, G, kNfb32>::kTotal));
        k<<<grid_for(e, n), kStageNT, FrontendSmem<float, G, kNfb32>::kTotal, st>>>(pcm, n, e->tab32, mel, time_major);
    } else if constexpr (std::is_same<G, GeoNS40x98>::value) {
        NWW_CUDA(set_smem(frontend3_kernel, Fe3KernelSmem::kTotal));
        frontend3_kernel<<<grid_for(e, n), Fe3::NT, Fe3KernelSmem::kTotal, st>>>(pcm, n, e->tab64, mel, time_major, 0);
    } else {
        auto k = frontend_kernel<double, G, kNfbRef, kStageNT>;
        NWW_CUDA(set_smem(k, FrontendSmem<double, G, kNfbRef>::kTotal));
        k<<<grid_for(e, n), kStageNT, FrontendSmem<double, G, kNfbRef>::kTotal, st>>>(pcm, n, e->tab64, mel, time_major);
    }
    e->launches++;
    NWW_CUDA(cudaGetLastError());
    return NWW_OK;
}

static int launch_tail_tc(nww_engine* e, int64_t n, float* scores, float* logits, float* emb, cudaStream_t st) {
    const TailLayer& L0 = e->tail.layers[0];
    if (!e->cnn2_enabled) {                                     // the v2 CNN stage writes the hi / lo pair itself
        const long long n4 = n * (long long)L0.K / 4;
        split_tf32_kernel<<<(int)std::min<long long>((n4 + 255) / 256, (long long)e->sm_count * 8), 256, 0, st>>>(
            e->d_feat, e->d_feat_hi, e->d_feat_lo, n, L0.K, e->tc_kp);
        e->launches++;
        NWW_CUDA(cudaGetLastError());
    }
    // split-K with a FIXED number of K blocks per split (so a window's result does not depend on the batch
    // it is scored in): 7680 / 32 = 240 K blocks -> 12 splits x 10 row tiles = 120 CTAs for a full chunk;
    // partial sums are reduced in a fixed order by the tail's pre-stage
    const int m_tiles = (int)((n + kTcBM - 1) / kTcBM);
    const int nkb = e->tc_kp / kTcBK;
    const int kbps = std::max(kTcKbPerSplit, (nkb + kTcMaxSplits - 1) / kTcMaxSplits);
    const int splits = (nkb + kbps - 1) / kbps;
    GemmTcArgs a{L0.b, L0.ln_g, L0.ln_b, nullptr, (int)n, L0.N, e->tc_kp, L0.post, e->spec.activation, e->d_part, kbps, e->tc_rows};
    const size_t smem_tc = tc_smem_bytes(L0.N);
    NWW_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc));
    gemm_tf32x3_kernel<<<dim3(m_tiles, splits), kTcThreads, smem_tc, st>>>(e->tm_xhi, e->tm_xlo, e->tm_whi, e->tm_wlo, a);
    e->launches++;
    NWW_CUDA(cudaGetLastError());
    TailParams P = e->tail_rest;
    P.pre_part = e->d_part;
    P.pre_b = L0.b;
    P.pre_g = L0.ln_g;
    P.pre_beta = L0.ln_b;
    P.pre_splits = splits;
    P.pre_mpad = e->tc_rows;
    P.pre_N = L0.N;
    P.pre_post = L0.post;
    // what is left is a chain of small layers: 8-window tiles give one CTA per SM at chunk size
    const size_t smem = tail_smem_bytes(P.max_width, kTailTMSmall);
    NWW_CUDA(set_smem(tail_kernel_t<kTailTMSmall>, smem));
    const int64_t tiles = (n + kTailTMSmall - 1) / kTailTMSmall;
    tail_kernel_t<kTailTMSmall><<<grid_for(e, tiles, 2), kTailNT, smem, st>>>(nullptr, n, P, scores, logits, emb);
    e->launches++;
    NWW_CUDA(cudaGetLastError());
    return NWW_OK;
}

static int launch_tail(nww_engine* e, const float* feat, int64_t n, float* scores, float* logits, float* emb,
                       cudaStream_t st) {
    if (e->tc_enabled) return launch_tail_tc(e, n, scores, logits, emb, st);
    // small layers only: favour CTA count — until the launch has enough rows to fill the GPU with 32-window tiles too
    // (a stream bank of 65 536 rows: the weights are then fetched by a quarter of the CTAs)
    if ((size_t)e->tail.layers[0].K * e->tail.layers[0].N <= 65536 && n < (int64_t)kTailTM * 4 * e->sm_count) {
        const size_t smem = tail_smem_bytes(e->tail.max_width, kTailTMSmall);
        NWW_CUDA(set_smem(tail_kernel_t<kTailTMSmall>, smem));
        const int64_t tiles = (n + kTailTMSmall - 1) / kTailTMSmall;
        tail_kernel_t<kTailTMSmall><<<grid_for(e, tiles, 2), kTailNT, smem, st>>>(feat, n, e->tail, scores, logits, emb);
        e->launches++;
        NWW_CUDA(cudaGetLastError());
        return NWW_OK;
    }
    const size_t smem = tail_smem_bytes(e->tail.max_width);
    NWW_CUDA(set_smem(tail_kernel, smem));
    const int64_t tiles = (n + kTailTM - 1) / kTailTM;
    tail_kernel<<<grid_for(e, tiles, 2), kTailNT, smem, st>>>(feat, n, e->tail, scores, logits, emb);
    e->launches++;
    NWW_CUDA(cudaGetLastError());
    return NWW_OK;
}

// The fused CNN stage (PCM or log-mel -> conv1 -> tcgen05 conv2 -> feature rows): v3 = warp-specialised pipeline
// (nww_cnn3.cuh, default), v2 = the phase-serial kernel it replaces (nww_cnn2.cuh; reserved[0] bit 3, A/B measurements).
static int launch_cnn_stage(nww_engine* e, WindowSource pcm, Cnn2MelSource ms, int64_t n, float* feat_hi, float* feat_lo, float* mel,
                            cudaStream_t st) {
    const int act = e->spec.activation;
    if ((e->spec.reserved[0] & 64) && pcm.base != nullptr && pcm.offsets == nullptr && pcm.fbase == nullptr && ms.ring == nullptr) {
        // two co-resident kernels: front end of sub-chunk k + 1 on the engine's front-end stream beside the convolution of
        // sub-chunk k on the caller's stream; the log-mel goes through a (n, F, T) buffer in L2
        const int64_t fe_floats = (int64_t)e->n_mels * e->n_frames;
        if (!mel && !e->d_melf) NWW_CUDA(cudaMalloc(&e->d_melf, (size_t)e->chunk * fe_floats * sizeof(float)));
        float* melbuf = mel ? mel : e->d_melf;
        if (!e->fe_stream) {
            NWW_CUDA(cudaStreamCreateWithFlags(&e->fe_stream, cudaStreamNonBlocking));
            NWW_CUDA(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
        }
        const int64_t S = (int64_t)e->sm_count * e->split_per_sm;
        const int n_sub = (int)((n + S - 1) / S);
        while ((int)e->ev_fe.size() < n_sub) {
            cudaEvent_t ev = nullptr;
            NWW_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            e->ev_fe.push_back(ev);
        }
        auto kc = act == NWW_ACT_RELU ? conv4_mel_kernel<ACT_RELU> : act == NWW_ACT_GELU ? conv4_mel_kernel<ACT_GELU>
                                                                                           : conv4_mel_kernel<ACT_SILU>;
        NWW_CUDA(set_smem(fe4_mel_kernel, Fe4::kTotal));
        NWW_CUDA(set_smem(kc, Conv4::kTotal));
        NWW_CUDA(cudaFuncSetAttribute(fe4_mel_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        NWW_CUDA(cudaFuncSetAttribute(kc, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        NWW_CUDA(cudaEventRecord(e->ev_fork, st));                          // inputs ready, workspaces free: in the caller's stream order
        NWW_CUDA(cudaStreamWaitEvent(e->fe_stream, e->ev_fork, 0));
        for (int k = 0; k < n_sub; ++k) {
            const int64_t w0 = k * S, m = std::min<int64_t>(S, n - w0);
            fe4_mel_kernel<<<grid_for(e, m), Fe4::NT, Fe4::kTotal, e->fe_stream>>>(pcm.base + w0 * e->clip, m, e->tab64, melbuf + w0 * fe_floats);
            NWW_CUDA(cudaGetLastError());
            NWW_CUDA(cudaEventRecord(e->ev_fe[k], e->fe_stream));
            NWW_CUDA(cudaStreamWaitEvent(st, e->ev_fe[k], 0));
            kc<<<grid_for(e, m), Conv4::NT, Conv4::kTotal, st>>>(melbuf + w0 * fe_floats, m, e->cnn2, feat_hi + w0 * Cnn2::FEAT,
                                                                 feat_lo ? feat_lo + w0 * Cnn2::FEAT : nullptr);
            NWW_CUDA(cudaGetLastError());
            e->launches += 2;
        }
        return NWW_OK;
    }
    if (e->spec.reserved[0] & 8) {
        auto k = act == NWW_ACT_RELU ? cnn2_stage_kernel<ACT_RELU> : act == NWW_ACT_GELU ? cnn2_stage_kernel<ACT_GELU>
                                                                                           : cnn2_stage_kernel<ACT_SILU>;
        NWW_CUDA(set_smem(k, Cnn2::kTotal));
        k<<<grid_for(e, n), Cnn2::NT, Cnn2::kTotal, st>>>(pcm, ms, n, e->tab64, e->cnn2, feat_hi, feat_lo, mel);
    } else {
        auto k = act == NWW_ACT_RELU ? cnn3_stage_kernel<ACT_RELU> : act == NWW_ACT_GELU ? cnn3_stage_kernel<ACT_GELU>
                                                                                           : cnn3_stage_kernel<ACT_SILU>;
        NWW_CUDA(set_smem(k, Cnn3::kTotal));
        k<<<grid_for(e, n), Cnn3::NT, Cnn3::kTotal, st>>>(pcm, ms, n, e->tab64, e->cnn2, feat_hi, feat_lo, mel);
    }
    e->launches++;
    NWW_CUDA(cudaGetLastError());
    return NWW_OK;
}

// Stage A for one chunk: PCM (int16, device) -> feature rows in e->d_feat.
// stream_s0 >= 0: stream mode, the log-mel of window i is that of stream stream_s0 + i in the mel ring.
static int launch_stage_a(nww_engine* e, WindowSource pcm, int64_t n, float* mel, cudaStream_t st, int64_t stream_s0 = -1) {
    const bool from_ring = stream_s0 >= 0;
    // window i of this launch group is stream stream_s0 + i of the bank, or, in a selective push, stream sel_ids[stream_s0 + i]
    const MelRingRef ring_ref{from_ring ? e->d_mel_ring : nullptr, e->streams.count, from_ring ? stream_s0 : 0, e->sel_ids};
    // Float feed into a stage kernel that stages int16 PCM itself (cnn2_stage_kernel): log-mel by the float front end
    // into d_melf first, then the kernel starts from that plain (n, F, T) buffer (count == nullptr marks the layout).
    Cnn2MelSource cnn2_src = ring_ref;
    if (pcm.fbase != nullptr && (e->cnn2_enabled || e->crnn_cnn2)) {
        if (!e->d_melf) NWW_CUDA(cudaMalloc(&e->d_melf, (size_t)e->chunk * e->n_mels * e->n_frames * sizeof(float)));
        int rc = launch_frontend<GeoNS40x98>(e, pcm, n, e->d_melf, 0, st);
        if (rc) return rc;
        if (mel) NWW_CUDA(cudaMemcpyAsync(mel, e->d_melf, (size_t)n * e->n_mels * e->n_frames * sizeof(float), cudaMemcpyDeviceToDevice, st));
        mel = nullptr;
        cnn2_src = Cnn2MelSource{e->d_melf, nullptr, 0, nullptr};
    }
    switch (e->spec.arch) {
        case NWW_ARCH_DNN: {
            if (from_ring) {
                stream_mel_gather_kernel<<<ew_grid(n * 3920, e->sm_count), 256, 0, st>>>(ring_ref, n, e->d_feat, 1);
                e->launches++;
                NWW_CUDA(cudaGetLastError());
                return NWW_OK;
            }
            // the DNN body is the identity on the (T, F) log-mel: features = flattened mel
            int rc = launch_frontend<GeoNS40x98>(e, pcm, n, e->d_feat, /*time_major=*/1, st);
            if (rc) return rc;
            if (mel) return launch_frontend<GeoNS40x98>(e, pcm, n, mel, 0, st);
            return NWW_OK;
        }
        case NWW_ARCH_CNN: {
            using G = GeoNS40x98;
            if (e->cnn2_enabled) {
                int rc = launch_cnn_stage(e, pcm, cnn2_src, n, e->d_feat_hi, e->d_feat_lo, mel, st);
                return rc;
            }
            if (pcm.fbase != nullptr) return fail(NWW_EUNSUPPORTED, "float feeds need the default (v2) CNN stage");
            auto k = cnn_stage_kernel<double, G, kNfb64, kStageNT>;
            const size_t smem = CnnSmem<double, G, kNfb64>::kTotal;
            NWW_CUDA(set_smem(k, smem));
            k<<<grid_for(e, n), kStageNT, smem, st>>>(pcm, n, e->tab64, e->cnn, e->spec.activation, e->d_feat, mel);
            e->launches++;
            NWW_CUDA(cudaGetLastError());
            return NWW_OK;
        }
        case NWW_ARCH_CRNN_GRU:
            if (e->crnn_cnn2) {
                int rc = launch_cnn_stage(e, pcm, cnn2_src, n, e->d_nhwc, nullptr, mel, st);
                if (rc) return rc;
                return launch_head_stage_a(e->heads, e->tab64, e->spec.activation, e->sm_count, pcm, n, e->d_feat, e->d_scratch,
                                           nullptr, st, &e->launches, &g_last_error, true, e->d_nhwc);
            }
            // fall through
        default:
            if (from_ring && e->spec.arch == NWW_ARCH_TCN && e->heads.tcn_cone)     // the cone kernel reads the ring itself
                return launch_head_stage_a(e->heads, e->tab64, e->spec.activation, e->sm_count, pcm, n, e->d_feat, e->d_scratch,
                                           mel, st, &e->launches, &g_last_error, true, nullptr,
                                           ring_ref);
            if (from_ring) {
                const int tm = e->spec.arch == NWW_ARCH_GRU || e->spec.arch == NWW_ARCH_LSTM ||
                               e->spec.arch == NWW_ARCH_QUARTZNET;                               // sequence heads read (T, F)
                stream_mel_gather_kernel<<<ew_grid(n * 3920, e->sm_count), 256, 0, st>>>(ring_ref, n, e->d_scratch, tm);
                e->launches++;
                NWW_CUDA(cudaGetLastError());
            }
            return launch_head_stage_a(e->heads, e->tab64, e->spec.activation, e->sm_count, pcm, n, e->d_feat, e->d_scratch,
                                       mel, st, &e->launches, &g_last_error, from_ring);
    }
}

static int run_device(nww_engine* e, const int16_t* pcm, int64_t n, float* scores, float* mel, float* logits, float* emb,
                      cudaStream_t st, const long long* win_off = nullptr, bool from_mel_ring = false,
                      const float* pcm_f32 = nullptr, int64_t stream_base = 0 /* window 0 of the call is this stream of the bank */) {
    const int64_t mel_stride = (int64_t)e->n_mels * e->n_frames;
    for (int64_t w0 = 0; w0 < n; w0 += e->chunk) {
        const int64_t m = std::min<int64_t>(e->chunk, n - w0);
        cudaEvent_t ea = nullptr, eb = nullptr, ec = nullptr;
        if (e->profiling) {
            ea = e->get_event(); eb = e->get_event(); ec = e->get_event();
            cudaEventRecord(ea, st);
        }
        const WindowSource src = pcm_f32  ? WindowSource{nullptr, nullptr, e->clip, pcm_f32 + w0 * e->clip}
                                 : win_off ? WindowSource{pcm, win_off + w0, e->clip}
                                           : WindowSource{pcm + w0 * e->clip, nullptr, e->clip};
        int rc = launch_stage_a(e, src, m, mel ? mel + w0 * mel_stride : nullptr, st, from_mel_ring ? stream_base + w0 : -1);
        if (rc) return rc;
        if (e->profiling) cudaEventRecord(eb, st);
        rc = launch_tail(e, e->d_feat, m, scores + w0, logits ? logits + w0 : nullptr,
                         emb ? emb + w0 * e->emb_dim : nullptr, st);
        if (rc) return rc;
        if (e->profiling) {
            cudaEventRecord(ec, st);
            e->spans.push_back({ea, eb, 0, m});
            e->spans.push_back({eb, ec, 1, m});
        }
    }
    e->windows += n;
    return NWW_OK;
}

static void stream_free(nww_engine* e) {
    cudaFree(e->streams.ring);
    cudaFree(e->streams.wpos);
    cudaFree(e->streams.count);
    cudaFree(e->streams.win_off);
    cudaFree(e->d_mel_ring);
    e->d_mel_ring = nullptr;
    e->mel_inc = false;
    e->streams = StreamState{};
    // the selective-push staging is sized by the bank: a re-opened (larger) bank allocates it again
    cudaFree(e->d_sel_ids);
    cudaFree(e->d_sel_off);
    cudaFree(e->d_sel_scores);
    e->d_sel_ids = nullptr;
    e->d_sel_off = nullptr;
    e->d_sel_scores = nullptr;
    e->sel_cap = 0;
}

template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters) {
    // 16 independent accumulator chains per thread: enough ILP to saturate the pipe at 8 resident warps per scheduler
    T a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (T)(threadIdx.x + i) * (T)1e-3;
    const T x = (T)1.0000001, y = (T)1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = a[i] * x + y;
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == (T)-1.0) out[0] = s;           // never true: keeps the chains alive
}

// ------------------------------------------------------------------------------ C ABI
extern "C" {

const char* nww_last_error(void) { return g_last_error.c_str(); }

int nww_create(const nww_spec* spec, const void* weights, size_t weights_size, int device, nww_engine** out) {
    if (!spec || !weights || !out) return fail(NWW_EINVAL, "nww_create: null argument");
    if (spec->struct_size != sizeof(nww_spec)) return fail(NWW_EINVAL, "nww_create: nww_spec size mismatch");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(NWW_ECUDA, std::string("no CUDA device available (this engine has no CPU fallback): ") +
                                   cudaGetErrorString(ce));
    if (device < 0 || device >= ndev) return fail(NWW_EINVAL, "nww_create: bad device index");
    NWW_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    NWW_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(NWW_EUNSUPPORTED, "this library is built for sm_100a (B200) only; found sm_" +
                                          std::to_string(prop.major) + std::to_string(prop.minor));

    std::unique_ptr<nww_engine> e(new nww_engine);
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    e->spec = *spec;
    if (spec->reserved[1] > 0) e->split_per_sm = spec->reserved[1];
    e->blob_host.assign(static_cast<const unsigned char*>(weights), static_cast<const unsigned char*>(weights) + weights_size);
    std::string err;
    if (!parse_blob(e->blob_host.data(), e->blob_host.size(), &e->blob, &err)) return fail(NWW_EINVAL, err);
    if (spec->frontend_precision != NWW_FRONTEND_FP64 && spec->frontend_precision != NWW_FRONTEND_FP32)
        return fail(NWW_EINVAL, "bad frontend_precision");

    int rc = (spec->geometry == NWW_GEOM_NS40X98)    ? setup_frontend<GeoNS40x98>(e.get())
             : (spec->geometry == NWW_GEOM_REF64X101) ? setup_frontend<GeoREF64x101>(e.get())
                                                      : fail(NWW_EUNSUPPORTED, "unknown geometry id");
    if (rc) return rc;

    NWW_CUDA(cudaMalloc(&e->d_blob, e->blob_host.size()));
    NWW_CUDA(cudaMemcpy(e->d_blob, e->blob_host.data(), e->blob_host.size(), cudaMemcpyHostToDevice));
    NWW_CUDA(cudaMalloc(&e->arena.dev, e->arena.host.size()));
    NWW_CUDA(cudaMemcpy(e->arena.dev, e->arena.host.data(), e->arena.host.size(), cudaMemcpyHostToDevice));
    rebase_tables(&e->tab32, e->arena.dev);
    rebase_tables(&e->tab64, e->arena.dev);

    const bool ns = spec->geometry == NWW_GEOM_NS40X98;
    switch (spec->arch) {
        case NWW_ARCH_DNN:
            if (!ns) return fail(NWW_EUNSUPPORTED, "dnn head is built for the NS40x98 geometry");
            e->feat_dim = e->n_mels * e->n_frames;
            break;
        case NWW_ARCH_CNN: {
            if (!ns) return fail(NWW_EUNSUPPORTED, "cnn head is built for the NS40x98 geometry");
            e->feat_dim = CnnDims<GeoNS40x98>::FEAT;
            e->cnn = CnnWeights{e->dptr("cnn.w1"), e->dptr("cnn.b1"), e->dptr("cnn.w2"), e->dptr("cnn.b2")};
            if (!e->cnn.w1 || !e->cnn.b1 || !e->cnn.w2 || !e->cnn.b2) return fail(NWW_EINVAL, "weight blob: cnn.* missing");
            break;
        }
        default: {
            auto lookup = [&](const char* name, size_t expect) -> const float* {
                const BlobTensor* t = e->blob.find(name);
                if (!t || (expect && t->numel() != expect)) return nullptr;
                return e->dptr(name);
            };
            auto dims = [&](const char* name) -> std::vector<uint32_t> {
                const BlobTensor* t = e->blob.find(name);
                return t ? t->dims : std::vector<uint32_t>();
            };
            rc = setup_head_weights(spec->arch, spec->geometry, lookup, dims, &e->heads, &e->feat_dim, &g_last_error);
            if (rc) return rc;
            if (spec->arch == NWW_ARCH_CRNN_GRU && e->heads.crnn_cnn2 && !(spec->reserved[0] & 1)) {
                // conv1 + conv2 (same shapes as the CNN head's) through cnn2_stage_kernel, channel-last FP32 output
                const float* w0 = e->blob.f32("crnn.conv0.w");    // [1][tap 9][oc 16]
                rc = build_cnn2_weights(e.get(), [&](int oc, int tap) { return w0[tap * 16 + oc]; }, e->blob.f32("crnn.conv1.w"),
                                        e->heads.crnn_conv[0].b, e->heads.crnn_conv[1].b);
                if (rc) return rc;
                e->crnn_cnn2 = true;
                if (conv_umma_plan(10, 24, e->heads.crnn_ch[1], e->heads.crnn_ch[2], 1, &e->heads.crnn_plan3)) {
                    std::vector<uint16_t> wq;
                    conv_umma_pack_weights(e->blob.f32("crnn.conv2.w"), e->heads.crnn_ch[1], e->heads.crnn_ch[2], &wq);
                    NWW_CUDA(cudaMalloc(&e->d_conv_wq[0], wq.size() * sizeof(uint16_t)));
                    NWW_CUDA(cudaMemcpy(e->d_conv_wq[0], wq.data(), wq.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                    e->heads.crnn_wq3 = reinterpret_cast<const uint4*>(e->d_conv_wq[0]);
                }
            }
            if (spec->arch == NWW_ARCH_TCN && e->heads.tcn_cone && !(spec->reserved[0] & 1)) {
                // every layer of the cone on tcgen05: layer program + weights pre-split into the chunks the kernel eats
                std::vector<TuHostLayer> hl;
                int cin = e->heads.tcn_in;
                for (int l = 0; l < e->heads.tcn_levels; ++l) {
                    const std::string p = "tcn." + std::to_string(l);
                    const int c = e->heads.tcn_ch[l];
                    hl.push_back(TuHostLayer{e->blob.f32(p + ".conv1.w"), e->heads.tcn_c1[l].b, cin, 3, c});
                    if (cin != c) hl.push_back(TuHostLayer{e->blob.f32(p + ".down.w"), e->heads.tcn_down[l].b, cin, 1, c});
                    hl.push_back(TuHostLayer{e->blob.f32(p + ".conv2.w"), e->heads.tcn_c2[l].b, c, 3, c});
                    cin = c;
                }
                std::vector<uint16_t> wq;
                if (tcn_umma_build(e->heads.tcn_plan, hl, &e->heads.tcn_uplan, &wq)) {
                    NWW_CUDA(cudaMalloc(&e->d_conv_wq[2], wq.size() * sizeof(uint16_t)));
                    NWW_CUDA(cudaMemcpy(e->d_conv_wq[2], wq.data(), wq.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                    e->heads.tcn_wq = reinterpret_cast<const uint4*>(e->d_conv_wq[2]);
                    e->heads.tcn_umma = true;
                }
            }
            if (spec->arch == NWW_ARCH_TCN && e->heads.tcn_cone && !(spec->reserved[0] & 1) && !(spec->reserved[0] & 16)) {
                // the cone as one tcgen05 row GEMM per layer over all windows of a launch group (reserved[0] bit 4 keeps
                // the per-tile cone kernel): layer program + weights as K-chunked operand streams
                const TcnConeParams& C = e->heads.tcn_plan;
                HeadWeights& H = e->heads;
                int li = 0;
                bool ok = true;
                auto add = [&](const float* w_host, const float* bias_dev, int Cin, int taps, int Cout, int a_in_mel, long long a_off,
                               int a_rs, int n_pos, int o_in_feat, long long o_off, int act, int pre_relu, int has_res, long long r_off,
                               int r_rs) -> int {
                    if (li >= 12 || Cout % 64 || Cout > 512 || Cin % 4) { ok = false; return NWW_OK; }
                    const int K = taps * Cin, Kp = (K + kKcKC - 1) / kKcKC * kKcKC;
                    std::vector<float> wp((size_t)Kp * Cout, 0.0f);
                    memcpy(wp.data(), w_host, (size_t)K * Cout * sizeof(float));
                    std::vector<uint16_t> wq;
                    rowgemm_kc_pack(wp.data(), Kp, Cout, &wq);
                    void* d = nullptr;
                    NWW_CUDA(cudaMalloc(&d, wq.size() * sizeof(uint16_t)));
                    e->d_extra.push_back(d);
                    NWW_CUDA(cudaMemcpy(d, wq.data(), wq.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                    H.tcn_row[li++] = HeadWeights::TcnRowLayer{a_off, o_off, r_off, a_in_mel, o_in_feat, has_res, a_rs, r_rs, n_pos, Kp, K,
                                                               Cout, act, pre_relu, reinterpret_cast<const uint4*>(d), bias_dev};
                    return NWW_OK;
                };
                int cin = C.c_in;
                long long x_off = (long long)(C.T - C.n_in) * C.c_in;       // level 0 reads the time-major log-mel itself
                int x_in_mel = 1;
                for (int l = 0; l < C.levels && ok; ++l) {
                    const std::string p = "tcn." + std::to_string(l);
                    const int Cc = C.ch[l];
                    const bool last = l + 1 == C.levels;
                    rc = add(e->blob.f32(p + ".conv1.w"), H.tcn_c1[l].b, cin, 3, Cc, x_in_mel, x_off, cin, C.n_mid[l], 0, C.off_mid[l], 1, 0,
                             0, 0, 0);
                    if (rc) return rc;
                    long long r_off = x_off + 4ll * cin;                     // identity residual of output p = block input 2 p + 4
                    int r_rs = 2 * cin;
                    if (cin != Cc) {
                        rc = add(e->blob.f32(p + ".down.w"), H.tcn_down[l].b, cin, 1, Cc, x_in_mel, x_off + 4ll * cin, 2 * cin, C.n_out[l], 0,
                                 H.tcn_res_off[l], 0, 0, 0, 0, 0);
                        if (rc) return rc;
                        r_off = H.tcn_res_off[l];
                        r_rs = Cc;
                    } else if (x_in_mel) {
                        ok = false;                                          // an identity residual straight from the log-mel: not built
                    }
                    rc = add(e->blob.f32(p + ".conv2.w"), H.tcn_c2[l].b, Cc, 3, Cc, 0, C.off_mid[l], 2 * Cc, C.n_out[l], last ? 1 : 0,
                             last ? 0 : C.off_out[l], 1, 1, 1, r_off, r_rs);
                    if (rc) return rc;
                    x_off = C.off_out[l];
                    x_in_mel = 0;
                    cin = Cc;
                }
                H.tcn_n_row_layers = li;
                H.tcn_rows = ok;
                // bit 7: all layers (+ the stream-mode gather) in ONE cooperative launch.  Measured 15 % slower than one launch
                // per layer (14.7 vs 18.2 M stream-steps/s at 65 536 streams: the grid barriers cost more than the launches
                // they replace), so it is opt-in: 4 launches per push instead of 12
                H.tcn_fused = (spec->reserved[0] & 128) != 0;
            }
            if (spec->arch == NWW_ARCH_CRNN_GRU && e->heads.gru_hidden == kGruTcH && e->heads.gru_wih_f_kn &&
                !(spec->reserved[0] & 1)) {
                std::vector<uint16_t> wq;                         // blob w_hh is (H, 3H) = [k][n]
                gru_tc_pack_whh(e->blob.f32("crnn.gru.fwd.w_hh"), &wq);
                NWW_CUDA(cudaMalloc(&e->d_conv_wq[1], wq.size() * sizeof(uint16_t)));
                NWW_CUDA(cudaMemcpy(e->d_conv_wq[1], wq.data(), wq.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                e->heads.gru_whh_q = reinterpret_cast<const uint4*>(e->d_conv_wq[1]);
                // input projections on tcgen05 too: W_ih^T (In, 3H) in 64-column chunks, both directions in one buffer
                const int In = e->heads.gru_in, G3 = 3 * kGruTcH;
                if (In % 16 == 0 && rowgemm_umma_smem_bytes(In) <= 220 * 1024) {
                    std::vector<uint16_t> qf, qb;
                    rowgemm_umma_pack(e->blob.f32("crnn.gru.fwd.w_ih_kn"), In, G3, &qf);
                    rowgemm_umma_pack(e->blob.f32("crnn.gru.bwd.w_ih_kn"), In, G3, &qb);
                    NWW_CUDA(cudaMalloc(&e->d_conv_wq[2], (qf.size() + qb.size()) * sizeof(uint16_t)));
                    NWW_CUDA(cudaMemcpy(e->d_conv_wq[2], qf.data(), qf.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                    NWW_CUDA(cudaMemcpy(static_cast<uint16_t*>(e->d_conv_wq[2]) + qf.size(), qb.data(), qb.size() * sizeof(uint16_t),
                                        cudaMemcpyHostToDevice));
                    e->heads.gru_wih_f_q = reinterpret_cast<const uint4*>(e->d_conv_wq[2]);
                    e->heads.gru_wih_b_q = reinterpret_cast<const uint4*>(static_cast<uint16_t*>(e->d_conv_wq[2]) + qf.size());
                }
            }
            if (spec->arch == NWW_ARCH_GRU || spec->arch == NWW_ARCH_LSTM) {
                // gate matrices as two-term half UMMA operands (nww_rnn.cuh); this head has no CUDA-core variant
                const int H = e->heads.rnn_hidden, KX = RnnDims<128, GeoNS40x98::N_MELS>::KX;
                std::vector<uint16_t> qf, qb;
                rnn_pack_weights(e->blob.f32("rnn.fwd.w"), KX + H, H, &qf);
                rnn_pack_weights(e->blob.f32("rnn.bwd.w"), KX, H, &qb);
                NWW_CUDA(cudaMalloc(&e->d_conv_wq[1], qf.size() * sizeof(uint16_t)));
                NWW_CUDA(cudaMemcpy(e->d_conv_wq[1], qf.data(), qf.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                NWW_CUDA(cudaMalloc(&e->d_conv_wq[2], qb.size() * sizeof(uint16_t)));
                NWW_CUDA(cudaMemcpy(e->d_conv_wq[2], qb.data(), qb.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                e->heads.rnn_wq_f = reinterpret_cast<const uint4*>(e->d_conv_wq[1]);
                e->heads.rnn_wq_b = reinterpret_cast<const uint4*>(e->d_conv_wq[2]);
            }
            if (spec->arch == NWW_ARCH_E2E_CNN) {
                for (int j = 0; j < 3; ++j) {                          // conv2..4 of RawAudioBackbone as row-GEMM weight streams
                    auto& L = e->heads.rc[j];
                    std::vector<uint16_t> wq;
                    rowgemm_kc_pack(e->blob.f32("rawcnn.conv" + std::to_string(j + 2) + ".w"), L.K, L.Npad, &wq);
                    void* d = nullptr;
                    NWW_CUDA(cudaMalloc(&d, wq.size() * sizeof(uint16_t)));
                    e->d_extra.push_back(d);
                    NWW_CUDA(cudaMemcpy(d, wq.data(), wq.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                    L.wq = reinterpret_cast<const uint4*>(d);
                }
            }
            if (spec->arch == NWW_ARCH_E2E_QUARTZNET || spec->arch == NWW_ARCH_E2E_CNN) {
                for (int i = 0; i < e->heads.raw_layers; ++i) {        // strided Conv1d layers as row-GEMM weight streams
                    auto& L = e->heads.raw[i];
                    std::vector<uint16_t> wq;
                    rowgemm_kc_pack(e->blob.f32("raw." + std::to_string(i) + ".w"), L.K, L.Npad, &wq);
                    void* d = nullptr;
                    NWW_CUDA(cudaMalloc(&d, wq.size() * sizeof(uint16_t)));
                    e->d_extra.push_back(d);
                    NWW_CUDA(cudaMemcpy(d, wq.data(), wq.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                    L.wq = reinterpret_cast<const uint4*>(d);
                }
            }
            if (spec->arch == NWW_ARCH_QUARTZNET || spec->arch == NWW_ARCH_E2E_QUARTZNET) {
                // folded pointwise (+ residual) weights of every block as bf16 UMMA operand streams (nww_rowgemm.cuh)
                for (int i = 0; i < e->heads.qn_blocks; ++i) {
                    auto& B = e->heads.qn[i];
                    std::vector<uint16_t> wq;
                    rowgemm_kc_pack(e->blob.f32("qn." + std::to_string(i) + ".w"), B.K, B.N, &wq);
                    void* d = nullptr;
                    NWW_CUDA(cudaMalloc(&d, wq.size() * sizeof(uint16_t)));
                    e->d_extra.push_back(d);
                    NWW_CUDA(cudaMemcpy(d, wq.data(), wq.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                    B.wq = reinterpret_cast<const uint4*>(d);
                }
            }
            if (spec->arch == NWW_ARCH_E2E_MELCNN && !(spec->reserved[0] & 1)) {
                // conv2 (16 -> 32 on 32 x 50, pool) and conv3 (32 -> 64 on 16 x 25) as tcgen05 implicit GEMMs
                const int cin[3] = {1, 16, 32}, cout[3] = {16, 32, 64}, hh[3] = {64, 32, 16}, ww[3] = {101, 50, 25}, pool[3] = {1, 1, 0};
                bool ok = true;
                for (int j = 1; j <= 2; ++j) ok = ok && conv_umma_plan(hh[j], ww[j], cin[j], cout[j], pool[j], &e->heads.e2e_plan[j]);
                if (ok && !(spec->reserved[0] & 256)) conv_umma_plan_front1(&e->heads.e2e_plan[1], hh[0], ww[0]);   // bit 8: keep conv1 a kernel of its own
                for (int j = 1; j <= 2 && ok; ++j) {
                    std::vector<uint16_t> wq;
                    conv_umma_pack_weights(e->blob.f32("e2e.conv" + std::to_string(j) + ".w"), cin[j], cout[j], &wq);
                    NWW_CUDA(cudaMalloc(&e->d_conv_wq[j], wq.size() * sizeof(uint16_t)));
                    NWW_CUDA(cudaMemcpy(e->d_conv_wq[j], wq.data(), wq.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                    e->heads.e2e_wq[j] = reinterpret_cast<const uint4*>(e->d_conv_wq[j]);
                }
            }
            if (spec->arch == NWW_ARCH_BCRESNET) e->heads.bc_stage = !(spec->reserved[0] & 512);
            if (spec->arch == NWW_ARCH_BCRESNET && !(spec->reserved[0] & 1)) {
                // pointwise + shortcut weights as pre-split bf16 UMMA operands (reserved[0] bit 0 keeps the FP32 row GEMM)
                const int ch[4] = {32, 64, 128, 256};
                for (int j = 0; j < 3; ++j) {
                    const std::string p = "bc." + std::to_string(j);
                    std::vector<uint16_t> wq;
                    bcu_pack_weights(e->blob.f32(p + ".pw.w"), e->blob.f32(p + ".sc.w"), ch[j], ch[j + 1], &wq);
                    NWW_CUDA(cudaMalloc(&e->d_bc_wq[j], wq.size() * sizeof(uint16_t)));
                    NWW_CUDA(cudaMemcpy(e->d_bc_wq[j], wq.data(), wq.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                    e->heads.bc_wq[j] = reinterpret_cast<const uint4*>(e->d_bc_wq[j]);
                }
            }
            e->scratch_per_window = e->heads.scratch_floats * sizeof(float);
        }
    }
    rc = build_tail(e.get());
    if (rc) return rc;

    // chunk: windows per internal launch group, a multiple of the SM count.  Measured on B200
    // (tools/chunk_bench.py): every head gains from large chunks (more CTAs / rows per launch; the
    // intermediates need not stay L2-resident), so take up to 28 windows per SM within a 2 GB budget.
    {
        const size_t per_window = (size_t)e->feat_dim * sizeof(float) * 3 + e->scratch_per_window;
        // the recurrent heads work on 128-window tiles, one per SM: two waves of those per chunk
        const bool rnn = spec->arch == NWW_ARCH_GRU || spec->arch == NWW_ARCH_LSTM;
        // the TCN's row-GEMM layers are per-launch-group kernels over (window, position) rows: one group should hold a whole
        // bank of streams (BASELINE config #3: 65 536), within a 3 GB workspace
        const bool tcn_rows = spec->arch == NWW_ARCH_TCN && e->heads.tcn_rows;
        int mult = (int)std::max<size_t>(1, std::min<size_t>(rnn ? 2 * kRnnTM : tcn_rows ? 448 : 28,
                                                              ((size_t)(tcn_rows ? 3 : 2) << 30) / (per_window * e->sm_count)));
        e->chunk = spec->chunk_windows > 0 ? spec->chunk_windows : e->sm_count * mult;
    }
    NWW_CUDA(cudaMalloc(&e->d_feat, (size_t)e->chunk * e->feat_dim * sizeof(float)));
    if (e->crnn_cnn2) NWW_CUDA(cudaMalloc(&e->d_nhwc, (size_t)e->chunk * Cnn2::FEAT * sizeof(float)));
    if (e->scratch_per_window) NWW_CUDA(cudaMalloc(&e->d_scratch, (size_t)e->chunk * e->scratch_per_window));
    {
        // First dense layer on tcgen05 (3xTF32) when its shape fits the tile constraints.
        const TailLayer& L0 = e->tail.layers[0];
        const bool want_tc = !(spec->reserved[0] & 1);          // reserved[0] bit 0: force the CUDA-core tail
        if (want_tc && e->tail.n_layers >= 3 && tc_layer_eligible(L0.N, L0.K) && get_tmap_encoder() != nullptr) {
            const int Kp = tc_padded_k(L0.K);
            e->tc_kp = Kp;
            const size_t wn = (size_t)L0.N * L0.K, wnp = (size_t)L0.N * Kp;
            std::vector<float> whi(wnp, 0.0f), wlo(wnp, 0.0f), wperm;
            const float* w = e->blob.f32("tail.0.W");
            // CNN stage v2 (tcgen05 conv2) emits features in the K order (ph, pw, oc): permute fc1's columns to match
            const bool want_cnn2 = spec->arch == NWW_ARCH_CNN && !(spec->reserved[0] & 2) && L0.K == Cnn2::FEAT;
            if (want_cnn2) {
                wperm.resize(wn);
                for (int nn = 0; nn < L0.N; ++nn)
                    for (int oc = 0; oc < Cnn2::C2; ++oc)
                        for (int ph = 0; ph < Cnn2::H2; ++ph)
                            for (int pw = 0; pw < Cnn2::W2; ++pw)
                                wperm[(size_t)nn * L0.K + (ph * Cnn2::W2 + pw) * Cnn2::C2 + oc] =
                                    w[(size_t)nn * L0.K + (oc * Cnn2::H2 + ph) * Cnn2::W2 + pw];
                w = wperm.data();
            }
            auto rn_tf32 = [](float x) {
                uint32_t u;
                memcpy(&u, &x, 4);
                u = (u + 0x1000u) & 0xFFFFE000u;                // round to nearest (ties away), like cvt.rna.tf32
                float y;
                memcpy(&y, &u, 4);
                return y;
            };
            for (int nn = 0; nn < L0.N; ++nn)
                for (int k = 0; k < L0.K; ++k) {
                    const float v = w[(size_t)nn * L0.K + k];
                    const float h = rn_tf32(v);
                    whi[(size_t)nn * Kp + k] = h;
                    wlo[(size_t)nn * Kp + k] = rn_tf32(v - h);
                }
            (void)wn;
            NWW_CUDA(cudaMalloc(&e->d_w_hi, wnp * sizeof(float)));
            NWW_CUDA(cudaMalloc(&e->d_w_lo, wnp * sizeof(float)));
            NWW_CUDA(cudaMemcpy(e->d_w_hi, whi.data(), wnp * sizeof(float), cudaMemcpyHostToDevice));
            NWW_CUDA(cudaMemcpy(e->d_w_lo, wlo.data(), wnp * sizeof(float), cudaMemcpyHostToDevice));
            const size_t rows = ((size_t)e->chunk + kTcBM - 1) / kTcBM * kTcBM;
            e->tc_rows = (int)rows;
            NWW_CUDA(cudaMalloc(&e->d_feat_hi, rows * Kp * sizeof(float)));
            NWW_CUDA(cudaMalloc(&e->d_feat_lo, rows * Kp * sizeof(float)));
            NWW_CUDA(cudaMemset(e->d_feat_hi, 0, rows * Kp * sizeof(float)));
            NWW_CUDA(cudaMemset(e->d_feat_lo, 0, rows * Kp * sizeof(float)));
            NWW_CUDA(cudaMalloc(&e->d_part, (size_t)kTcMaxSplits * rows * L0.N * sizeof(float)));
            const bool ok = make_tmap_2d(&e->tm_xhi, e->d_feat_hi, rows, Kp, kTcBM) &&
                            make_tmap_2d(&e->tm_xlo, e->d_feat_lo, rows, Kp, kTcBM) &&
                            make_tmap_2d(&e->tm_whi, e->d_w_hi, L0.N, Kp, L0.N) &&
                            make_tmap_2d(&e->tm_wlo, e->d_w_lo, L0.N, Kp, L0.N);
            if (!ok) return fail(NWW_ECUDA, "cuTensorMapEncodeTiled failed for the dense-layer operands");
            if (want_cnn2) {
                const float* w1 = e->blob.f32("cnn.w1");          // [oc 16][tap 9]
                int rcw = build_cnn2_weights(e.get(), [&](int oc, int tap) { return w1[oc * 9 + tap]; }, e->blob.f32("cnn.w2"),
                                             e->cnn.b1, e->cnn.b2);
                if (rcw) return rcw;
                e->cnn2_enabled = true;
            }
            e->tail_rest = e->tail;
            e->tail_rest.n_layers = e->tail.n_layers - 1;
            for (int i = 0; i < e->tail_rest.n_layers; ++i) e->tail_rest.layers[i] = e->tail.layers[i + 1];
            e->tail_rest.max_width = 1;
            for (int i = 0; i < e->tail_rest.n_layers; ++i)
                e->tail_rest.max_width = std::max(e->tail_rest.max_width, std::max(e->tail_rest.layers[i].N, e->tail_rest.layers[i].K));
            e->tc_enabled = true;
        }
    }
    NWW_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    NWW_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        NWW_CUDA(cudaEventCreateWithFlags(&e->ev_copy[i], cudaEventDisableTiming));
        NWW_CUDA(cudaEventCreateWithFlags(&e->ev_done[i], cudaEventDisableTiming));
    }
    *out = e.release();
    return NWW_OK;
}

void nww_destroy(nww_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    delete e;
}

}  // extern "C"

nww_engine::~nww_engine() {
    nww_engine* e = this;
    cudaFree(e->d_melf);
    if (e->fe_stream) cudaStreamDestroy(e->fe_stream);
    for (auto ev : e->ev_piece) if (ev) cudaEventDestroy(ev);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    for (auto ev : e->ev_fe) cudaEventDestroy(ev);
    cudaFree(e->d_sel_ids);
    cudaFree(e->d_sel_off);
    cudaFree(e->d_sel_scores);
    if (e->ev_last) cudaEventDestroy(e->ev_last);
    cudaFree(e->d_blob);
    cudaFree(e->arena.dev);
    cudaFree(e->d_feat);
    cudaFree(e->d_scratch);
    cudaFree(e->d_w_hi);
    cudaFree(e->d_w_lo);
    cudaFree(e->d_feat_hi);
    cudaFree(e->d_feat_lo);
    cudaFree(e->d_part);
    cudaFree(e->d_w2_umma);
    cudaFree(e->d_nhwc);
    for (int j = 0; j < 3; ++j) cudaFree(e->d_bc_wq[j]);
    for (int j = 0; j < 3; ++j) cudaFree(e->d_conv_wq[j]);
    for (void* p : e->d_extra) cudaFree(p);
    cudaFree(e->d_pcm[0]);
    cudaFree(e->d_pcm[1]);
    cudaFree(e->d_scores);
    cudaFree(e->d_chunk);
    cudaFree(e->d_ids);
    stream_free(e);
    for (auto& sp : e->spans) { if (sp.stage == 0) cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto ev : e->event_pool) cudaEventDestroy(ev);
    for (int i = 0; i < 2; ++i) {
        if (e->ev_copy[i]) cudaEventDestroy(e->ev_copy[i]);
        if (e->ev_done[i]) cudaEventDestroy(e->ev_done[i]);
    }
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
}

extern "C" {

int nww_get_info(nww_engine* e, nww_info_t* info) {
    if (!e || !info) return fail(NWW_EINVAL, "nww_get_info: null argument");
    info->device = e->device;
    info->sm_count = e->sm_count;
    info->n_mels = e->n_mels;
    info->n_frames = e->n_frames;
    info->clip_samples = e->clip;
    info->feature_dim = e->feat_dim;
    info->embedding_dim = e->emb_dim;
    info->chunk_windows = e->chunk;
    info->kernel_launches = e->launches;
    info->windows_scored = e->windows;
    return NWW_OK;
}

int nww_run_windows(nww_engine* e, const int16_t* pcm_dev, int64_t n, float* scores_dev, float* mel_dev, float* logits_dev,
                    float* emb_dev, void* stream) {
    if (!e || (n > 0 && (!pcm_dev || !scores_dev))) return fail(NWW_EINVAL, "nww_run_windows: null argument");
    if (n < 0) return fail(NWW_EINVAL, "nww_run_windows: negative window count");
    if (n == 0) return NWW_OK;
    if (reinterpret_cast<uintptr_t>(pcm_dev) & 15) return fail(NWW_EINVAL, "nww_run_windows: pcm_dev must be 16-byte aligned");
    std::lock_guard<std::mutex> lock(e->mu);
    NWW_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);   // NULL = the legacy default stream, as in CUDA
    NWW_CUDA(order_enter(e, st));
    int rc = run_device(e, pcm_dev, n, scores_dev, mel_dev, logits_dev, emb_dev, st);
    NWW_CUDA(order_leave(e, st));
    return rc;
}

int nww_logmel(nww_engine* e, const int16_t* pcm_dev, int64_t n, float* mel_dev, int time_major, void* stream) {
    if (!e || (n > 0 && (!pcm_dev || !mel_dev))) return fail(NWW_EINVAL, "nww_logmel: null argument");
    if (n <= 0) return n == 0 ? NWW_OK : fail(NWW_EINVAL, "nww_logmel: negative window count");
    if (reinterpret_cast<uintptr_t>(pcm_dev) & 15) return fail(NWW_EINVAL, "nww_logmel: pcm_dev must be 16-byte aligned");
    std::lock_guard<std::mutex> lock(e->mu);
    NWW_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);   // NULL = the legacy default stream, as in CUDA
    const WindowSource src{pcm_dev, nullptr, e->clip};
    NWW_CUDA(order_enter(e, st));
    int rc = e->spec.geometry == NWW_GEOM_NS40X98 ? launch_frontend<GeoNS40x98>(e, src, n, mel_dev, time_major, st)
                                                  : launch_frontend<GeoREF64x101>(e, src, n, mel_dev, time_major, st);
    NWW_CUDA(order_leave(e, st));
    return rc;
}

int nww_run_windows_f32(nww_engine* e, const float* pcm_dev, int64_t n, float* scores_dev, float* mel_dev, float* logits_dev,
                        float* emb_dev, void* stream) {
    // Float PCM as the reference feeds its session (x = int16 / 32768, nanointerpreter.py:750, 771-775) — but ANY float32
    // audio is taken as it is: the samples go through the FP64 front end (or the raw-audio layers) unquantised.  For
    // samples on the int16 grid the result is bit-identical to the generic int16 front end (x * Hann in FP64 is exact
    // either way); host callers that know their feed is on the grid should prefer the int16 entry points (half the bytes).
    if (!e || (n > 0 && (!pcm_dev || !scores_dev))) return fail(NWW_EINVAL, "nww_run_windows_f32: null argument");
    if (n <= 0) return n == 0 ? NWW_OK : fail(NWW_EINVAL, "nww_run_windows_f32: negative window count");
    if (reinterpret_cast<uintptr_t>(pcm_dev) & 15) return fail(NWW_EINVAL, "nww_run_windows_f32: pcm_dev must be 16-byte aligned");
    std::lock_guard<std::mutex> lock(e->mu);
    NWW_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);   // NULL = the legacy default stream, as in CUDA
    NWW_CUDA(order_enter(e, st));
    int rc = run_device(e, nullptr, n, scores_dev, mel_dev, logits_dev, emb_dev, st, nullptr, false, pcm_dev);
    NWW_CUDA(order_leave(e, st));
    return rc;
}

int nww_run_windows_host(nww_engine* e, const int16_t* pcm_host, int64_t n, float* scores_host) {
    if (!e || (n > 0 && (!pcm_host || !scores_host))) return fail(NWW_EINVAL, "nww_run_windows_host: null argument");
    if (n <= 0) return n == 0 ? NWW_OK : fail(NWW_EINVAL, "nww_run_windows_host: negative window count");
    std::lock_guard<std::mutex> lock(e->mu);
    NWW_CUDA(cudaSetDevice(e->device));
    if (!e->d_pcm[0]) {
        // copy / compute overlap wants several chunks per call: 8 windows per SM (38 MB) per copy
        e->host_chunk = std::min<int64_t>(e->chunk, (int64_t)e->sm_count * (e->heads.rnn_hidden ? 32 : 8));
        for (int i = 0; i < 2; ++i) NWW_CUDA(cudaMalloc(&e->d_pcm[i], (size_t)e->host_chunk * e->clip * sizeof(int16_t)));
    }
    if (e->scores_cap < n) {
        cudaFree(e->d_scores);
        e->d_scores = nullptr;
        NWW_CUDA(cudaMalloc(&e->d_scores, (size_t)n * sizeof(float)));
        e->scores_cap = n;
    }
    // copy(c) on copy_stream -> ev_copy[slot]; compute(c) on stream waits for it and records
    // ev_done[slot], which copy(c + 2) waits for before overwriting the slot.
    NWW_CUDA(order_enter(e, e->stream));
    int64_t c = 0;
    for (int64_t w0 = 0; w0 < n; w0 += e->host_chunk, ++c) {
        const int slot = (int)(c & 1);
        const int64_t m = std::min<int64_t>(e->host_chunk, n - w0);
        if (c >= 2) NWW_CUDA(cudaStreamWaitEvent(e->copy_stream, e->ev_done[slot], 0));
        NWW_CUDA(cudaMemcpyAsync(e->d_pcm[slot], pcm_host + w0 * e->clip, (size_t)m * e->clip * sizeof(int16_t),
                                 cudaMemcpyHostToDevice, e->copy_stream));
        NWW_CUDA(cudaEventRecord(e->ev_copy[slot], e->copy_stream));
        NWW_CUDA(cudaStreamWaitEvent(e->stream, e->ev_copy[slot], 0));
        int rc = run_device(e, e->d_pcm[slot], m, e->d_scores + w0, nullptr, nullptr, nullptr, e->stream);
        if (rc) return rc;
        NWW_CUDA(cudaEventRecord(e->ev_done[slot], e->stream));
    }
    NWW_CUDA(cudaMemcpyAsync(scores_host, e->d_scores, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    NWW_CUDA(order_leave(e, e->stream));
    NWW_CUDA(cudaStreamSynchronize(e->stream));
    return NWW_OK;
}

// ------------------------------------------------------------------------------ streams

int nww_stream_open(nww_engine* e, int64_t n_streams) {
    if (!e || n_streams <= 0) return fail(NWW_EINVAL, "nww_stream_open: need an engine and n_streams > 0");
    std::lock_guard<std::mutex> lock(e->mu);
    NWW_CUDA(cudaSetDevice(e->device));
    NWW_CUDA(cudaDeviceSynchronize());          // pushes may be in flight on any caller stream
    stream_free(e);
    StreamState st{};
    st.n_streams = n_streams;
    st.R = e->clip;
    const size_t ring_bytes = (size_t)n_streams * st.pitch() * sizeof(int16_t);
    NWW_CUDA(cudaMalloc(&st.ring, ring_bytes));
    NWW_CUDA(cudaMalloc(&st.wpos, (size_t)n_streams * sizeof(int)));
    NWW_CUDA(cudaMalloc(&st.count, (size_t)n_streams * sizeof(long long)));
    NWW_CUDA(cudaMalloc(&st.win_off, (size_t)n_streams * sizeof(long long)));
    e->streams = st;
    // incremental log-mel (nww_stream_mel.cuh): un-centred geometry, and a stage A that can start from mel
    const bool v1_cnn = e->spec.arch == NWW_ARCH_CNN && !e->cnn2_enabled;
    if (e->spec.geometry == NWW_GEOM_NS40X98 && e->spec.frontend_precision == NWW_FRONTEND_FP64 && !v1_cnn &&
        e->spec.arch != NWW_ARCH_E2E_QUARTZNET && e->spec.arch != NWW_ARCH_E2E_CNN /* raw-audio models: no log-mel */ &&
        !(e->spec.reserved[0] & 4)) {
        NWW_CUDA(cudaMalloc(&e->d_mel_ring, (size_t)n_streams * SMel::STREAM_FLOATS * sizeof(float)));
        NWW_CUDA(cudaMemsetAsync(e->d_mel_ring, 0, (size_t)n_streams * SMel::STREAM_FLOATS * sizeof(float), e->stream));
        e->mel_inc = true;
    }
    stream_reset_kernel<<<(unsigned)n_streams, 256, 0, e->stream>>>(st, nullptr, n_streams);
    e->launches++;
    NWW_CUDA(cudaGetLastError());
    NWW_CUDA(cudaStreamSynchronize(e->stream));
    return NWW_OK;
}

int nww_stream_close(nww_engine* e) {
    if (!e) return fail(NWW_EINVAL, "nww_stream_close: null engine");
    std::lock_guard<std::mutex> lock(e->mu);
    NWW_CUDA(cudaSetDevice(e->device));
    NWW_CUDA(cudaDeviceSynchronize());
    stream_free(e);
    return NWW_OK;
}

// Score the streams of the bank after an ingest step: all of them, or (ids_dev != nullptr) only the n_ids listed ones —
// the others report 0 and cost nothing beyond the ingest.
static int stream_score_locked(nww_engine* e, float* scores_dev, cudaStream_t st, bool from_mel, const long long* ids_dev, int64_t n_ids) {
    const StreamState& S = e->streams;
    int rc;
    if (ids_dev == nullptr) {
        rc = run_device(e, S.ring, S.n_streams, scores_dev, nullptr, nullptr, nullptr, st, S.win_off, from_mel);
        if (rc) return rc;
        stream_mask_kernel<<<(unsigned)((S.n_streams + 255) / 256), 256, 0, st>>>(S, scores_dev);
        e->launches++;
        NWW_CUDA(cudaGetLastError());
        return NWW_OK;
    }
    NWW_CUDA(cudaMemsetAsync(scores_dev, 0, (size_t)S.n_streams * sizeof(float), st));
    if (n_ids == 0) return NWW_OK;
    if (e->sel_cap < n_ids) {
        cudaFree(e->d_sel_off);
        cudaFree(e->d_sel_scores);
        e->d_sel_off = nullptr;
        e->d_sel_scores = nullptr;
        e->sel_cap = 0;
        NWW_CUDA(cudaMalloc(&e->d_sel_off, (size_t)S.n_streams * sizeof(long long)));
        NWW_CUDA(cudaMalloc(&e->d_sel_scores, (size_t)S.n_streams * sizeof(float)));
        e->sel_cap = S.n_streams;
    }
    const unsigned g = (unsigned)((n_ids + 255) / 256);
    stream_select_offsets_kernel<<<g, 256, 0, st>>>(S, ids_dev, n_ids, e->d_sel_off);
    e->launches++;
    NWW_CUDA(cudaGetLastError());
    e->sel_ids = ids_dev;
    rc = run_device(e, S.ring, n_ids, e->d_sel_scores, nullptr, nullptr, nullptr, st, e->d_sel_off, from_mel);
    e->sel_ids = nullptr;
    if (rc) return rc;
    stream_select_scatter_kernel<<<g, 256, 0, st>>>(S, ids_dev, n_ids, e->d_sel_scores, scores_dev);
    e->launches++;
    NWW_CUDA(cudaGetLastError());
    return NWW_OK;
}

static int stream_push_locked(nww_engine* e, const int16_t* chunks_dev, int chunk_len, float* scores_dev, cudaStream_t st,
                              const long long* ids_dev = nullptr, int64_t n_ids = 0) {
    const StreamState& S = e->streams;
    const int n_new = chunk_len / SMel::HOP;
    if (e->mel_inc && (chunk_len % SMel::HOP != 0 || n_new > SMel::MAX_NEW)) e->mel_inc = false;   // until the next full reset
    if (e->mel_inc && !(e->spec.reserved[0] & 32) && (reinterpret_cast<uintptr_t>(chunks_dev) & 15) == 0) {
        // fused ingest: ring append + the frames the chunk completes, one warp per stream (reserved[0] bit 5: the two-kernel form)
        NWW_CUDA(set_smem(stream_push_mel_kernel, SPush::kTotal));
        const int64_t groups = (S.n_streams + SPush::NW - 1) / SPush::NW;
        stream_push_mel_kernel<<<grid_for(e, groups), SPush::NT, SPush::kTotal, st>>>(S, chunks_dev, chunk_len, e->d_mel_ring, e->tab64, 0,
                                                                                    S.n_streams);
        e->launches++;
        NWW_CUDA(cudaGetLastError());
        return stream_score_locked(e, scores_dev, st, true, ids_dev, n_ids);
    }
    stream_append_kernel<<<(unsigned)S.n_streams, 256, 0, st>>>(S, chunks_dev, chunk_len);
    e->launches++;
    NWW_CUDA(cudaGetLastError());
    if (e->mel_inc) {
        NWW_CUDA(set_smem(stream_mel_update_kernel, SMel::kTotal));
        const int64_t groups = (S.n_streams + SMel::SPB - 1) / SMel::SPB;
        stream_mel_update_kernel<<<grid_for(e, groups), Fe2::NT, SMel::kTotal, st>>>(S, e->d_mel_ring, e->tab64, n_new);
        e->launches++;
        NWW_CUDA(cudaGetLastError());
        return stream_score_locked(e, scores_dev, st, true, ids_dev, n_ids);
    }
    return stream_score_locked(e, scores_dev, st, false, ids_dev, n_ids);
}

int nww_stream_push(nww_engine* e, const int16_t* chunks_dev, int32_t chunk_len, float* scores_dev, void* stream) {
    if (!e || !chunks_dev || !scores_dev) return fail(NWW_EINVAL, "nww_stream_push: null argument");
    if (chunk_len <= 0) return fail(NWW_EINVAL, "nww_stream_push: chunk_len must be positive");
    std::lock_guard<std::mutex> lock(e->mu);
    if (!e->streams.ring) return fail(NWW_EINVAL, "nww_stream_push: no streams are open (call nww_stream_open)");
    NWW_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    NWW_CUDA(order_enter(e, st));
    int rc = stream_push_locked(e, chunks_dev, chunk_len, scores_dev, st);
    NWW_CUDA(order_leave(e, st));
    return rc;
}

int nww_stream_push_host(nww_engine* e, const int16_t* chunks_host, int32_t chunk_len, float* scores_host) {
    if (!e || !chunks_host || !scores_host) return fail(NWW_EINVAL, "nww_stream_push_host: null argument");
    if (chunk_len <= 0) return fail(NWW_EINVAL, "nww_stream_push_host: chunk_len must be positive");
    std::lock_guard<std::mutex> lock(e->mu);
    if (!e->streams.ring) return fail(NWW_EINVAL, "nww_stream_push_host: no streams are open (call nww_stream_open)");
    NWW_CUDA(cudaSetDevice(e->device));
    const int64_t n = e->streams.n_streams;
    const size_t bytes = (size_t)n * chunk_len * sizeof(int16_t);
    if (e->chunk_cap < bytes) {
        cudaFree(e->d_chunk);
        e->d_chunk = nullptr;
        NWW_CUDA(cudaMalloc(&e->d_chunk, bytes));
        e->chunk_cap = bytes;
    }
    if (e->scores_cap < n) {
        cudaFree(e->d_scores);
        e->d_scores = nullptr;
        NWW_CUDA(cudaMalloc(&e->d_scores, (size_t)n * sizeof(float)));
        e->scores_cap = n;
    }
    NWW_CUDA(order_enter(e, e->stream));
    const StreamState& S = e->streams;
    const int n_new = chunk_len / SMel::HOP;
    constexpr int64_t kPieceMin = 4096;                        // streams per piece below which splitting costs more than it hides
    const bool piecewise = e->mel_inc && chunk_len % SMel::HOP == 0 && n_new <= SMel::MAX_NEW && !(e->spec.reserved[0] & 32) && n >= 2 * kPieceMin;
    int rc = NWW_OK;
    if (piecewise) {
        // The bank in up to four pieces: piece p + 1 crosses PCIe while piece p is ingested AND scored (a stream's step
        // depends on nothing but its own chunk), so a push costs max(copy, kernels) + one piece instead of their sum.
        // (measured, 65 536 streams x TCN: 1 piece 6.14 ms, 2: 4.82, 3: 4.44, 4: 4.45, 8: 4.63 per push; reserved[2] overrides for A/B runs)
        const int max_pieces = (e->spec.reserved[2] > 0 && e->spec.reserved[2] <= 8) ? e->spec.reserved[2] : 4;
        const int pieces = (int)std::max<int64_t>(1, std::min<int64_t>(max_pieces, n / kPieceMin));
        for (int p = 0; p < pieces; ++p)
            if (!e->ev_piece[p]) NWW_CUDA(cudaEventCreateWithFlags(&e->ev_piece[p], cudaEventDisableTiming));
        NWW_CUDA(cudaEventRecord(e->ev_piece[0], e->stream));                  // the staging buffer is free (earlier work on our stream)
        NWW_CUDA(cudaStreamWaitEvent(e->copy_stream, e->ev_piece[0], 0));
        NWW_CUDA(set_smem(stream_push_mel_kernel, SPush::kTotal));
        for (int p = 0; p < pieces && rc == NWW_OK; ++p) {
            const int64_t sb = n * p / pieces, se = n * (p + 1) / pieces;
            NWW_CUDA(cudaMemcpyAsync(e->d_chunk + sb * chunk_len, chunks_host + sb * chunk_len, (size_t)(se - sb) * chunk_len * sizeof(int16_t),
                                     cudaMemcpyHostToDevice, e->copy_stream));
            NWW_CUDA(cudaEventRecord(e->ev_piece[p], e->copy_stream));
            NWW_CUDA(cudaStreamWaitEvent(e->stream, e->ev_piece[p], 0));
            const int64_t groups = (se - sb + SPush::NW - 1) / SPush::NW;
            stream_push_mel_kernel<<<grid_for(e, groups), SPush::NT, SPush::kTotal, e->stream>>>(S, e->d_chunk, chunk_len, e->d_mel_ring, e->tab64, sb, se);
            e->launches++;
            NWW_CUDA(cudaGetLastError());
            rc = run_device(e, S.ring, se - sb, e->d_scores + sb, nullptr, nullptr, nullptr, e->stream, S.win_off + sb, true, nullptr, sb);
        }
        if (rc) return rc;
        stream_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(S, e->d_scores);
        e->launches++;
        NWW_CUDA(cudaGetLastError());
    } else {
        NWW_CUDA(cudaMemcpyAsync(e->d_chunk, chunks_host, bytes, cudaMemcpyHostToDevice, e->stream));
        rc = stream_push_locked(e, e->d_chunk, chunk_len, e->d_scores, e->stream);
        if (rc) return rc;
    }
    NWW_CUDA(cudaMemcpyAsync(scores_host, e->d_scores, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    NWW_CUDA(order_leave(e, e->stream));
    NWW_CUDA(cudaStreamSynchronize(e->stream));
    return NWW_OK;
}

int nww_stream_push_select(nww_engine* e, const int16_t* chunks_dev, int32_t chunk_len, const int64_t* ids_dev, int64_t n_ids,
                           float* scores_dev, void* stream) {
    if (!e || !chunks_dev || !scores_dev || (n_ids > 0 && !ids_dev)) return fail(NWW_EINVAL, "nww_stream_push_select: null argument");
    if (chunk_len <= 0 || n_ids < 0) return fail(NWW_EINVAL, "nww_stream_push_select: chunk_len must be positive, n_ids >= 0");
    std::lock_guard<std::mutex> lock(e->mu);
    if (!e->streams.ring) return fail(NWW_EINVAL, "nww_stream_push_select: no streams are open (call nww_stream_open)");
    if (n_ids > e->streams.n_streams) return fail(NWW_EINVAL, "nww_stream_push_select: more ids than streams");
    NWW_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static_assert(sizeof(long long) == sizeof(int64_t), "id width");
    static const long long kNone = 0;
    NWW_CUDA(order_enter(e, st));
    int rc = stream_push_locked(e, chunks_dev, chunk_len, scores_dev, st, n_ids ? reinterpret_cast<const long long*>(ids_dev) : &kNone, n_ids);
    NWW_CUDA(order_leave(e, st));
    return rc;
}

int nww_stream_push_select_host(nww_engine* e, const int16_t* chunks_host, int32_t chunk_len, const int64_t* ids_host, int64_t n_ids,
                                float* scores_host) {
    if (!e || !chunks_host || !scores_host || (n_ids > 0 && !ids_host)) return fail(NWW_EINVAL, "nww_stream_push_select_host: null argument");
    if (chunk_len <= 0 || n_ids < 0) return fail(NWW_EINVAL, "nww_stream_push_select_host: chunk_len must be positive, n_ids >= 0");
    std::lock_guard<std::mutex> lock(e->mu);
    if (!e->streams.ring) return fail(NWW_EINVAL, "nww_stream_push_select_host: no streams are open (call nww_stream_open)");
    const int64_t n = e->streams.n_streams;
    for (int64_t i = 0; i < n_ids; ++i)
        if (ids_host[i] < 0 || ids_host[i] >= n) return fail(NWW_EINVAL, "nww_stream_push_select_host: stream id out of range");
    if (n_ids > n) return fail(NWW_EINVAL, "nww_stream_push_select_host: more ids than streams");
    NWW_CUDA(cudaSetDevice(e->device));
    const size_t bytes = (size_t)n * chunk_len * sizeof(int16_t);
    if (e->chunk_cap < bytes) {
        cudaFree(e->d_chunk);
        e->d_chunk = nullptr;
        NWW_CUDA(cudaMalloc(&e->d_chunk, bytes));
        e->chunk_cap = bytes;
    }
    if (e->scores_cap < n) {
        cudaFree(e->d_scores);
        e->d_scores = nullptr;
        NWW_CUDA(cudaMalloc(&e->d_scores, (size_t)n * sizeof(float)));
        e->scores_cap = n;
    }
    if (!e->d_sel_ids) NWW_CUDA(cudaMalloc(&e->d_sel_ids, (size_t)n * sizeof(long long)));
    static const long long kNone = 0;
    NWW_CUDA(order_enter(e, e->stream));
    NWW_CUDA(cudaMemcpyAsync(e->d_chunk, chunks_host, bytes, cudaMemcpyHostToDevice, e->stream));
    if (n_ids) NWW_CUDA(cudaMemcpyAsync(e->d_sel_ids, ids_host, (size_t)n_ids * sizeof(int64_t), cudaMemcpyHostToDevice, e->stream));
    int rc = stream_push_locked(e, e->d_chunk, chunk_len, e->d_scores, e->stream, n_ids ? e->d_sel_ids : &kNone, n_ids);
    if (rc) return rc;
    NWW_CUDA(cudaMemcpyAsync(scores_host, e->d_scores, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    NWW_CUDA(order_leave(e, e->stream));
    NWW_CUDA(cudaStreamSynchronize(e->stream));
    return NWW_OK;
}

int nww_stream_reset(nww_engine* e, const int64_t* ids_host, int64_t n_ids) {
    if (!e) return fail(NWW_EINVAL, "nww_stream_reset: null engine");
    std::lock_guard<std::mutex> lock(e->mu);
    if (!e->streams.ring) return fail(NWW_EINVAL, "nww_stream_reset: no streams are open");
    NWW_CUDA(cudaSetDevice(e->device));
    const StreamState& S = e->streams;
    NWW_CUDA(order_enter(e, e->stream));        // after the pushes, whichever stream they were enqueued on
    if (!ids_host) {
        if (e->d_mel_ring) e->mel_inc = true;      // all counters return to 0: frame numbering restarts
        stream_reset_kernel<<<(unsigned)S.n_streams, 256, 0, e->stream>>>(S, nullptr, S.n_streams);
    } else {
        if (n_ids <= 0) return NWW_OK;
        for (int64_t i = 0; i < n_ids; ++i)
            if (ids_host[i] < 0 || ids_host[i] >= S.n_streams) return fail(NWW_EINVAL, "nww_stream_reset: stream id out of range");
        if (e->ids_cap < n_ids) {
            cudaFree(e->d_ids);
            e->d_ids = nullptr;
            NWW_CUDA(cudaMalloc(&e->d_ids, (size_t)n_ids * sizeof(long long)));
            e->ids_cap = n_ids;
        }
        static_assert(sizeof(long long) == sizeof(int64_t), "id width");
        NWW_CUDA(cudaMemcpyAsync(e->d_ids, ids_host, (size_t)n_ids * sizeof(int64_t), cudaMemcpyHostToDevice, e->stream));
        stream_reset_kernel<<<(unsigned)n_ids, 256, 0, e->stream>>>(S, e->d_ids, n_ids);
    }
    e->launches++;
    NWW_CUDA(cudaGetLastError());
    NWW_CUDA(order_leave(e, e->stream));
    NWW_CUDA(cudaStreamSynchronize(e->stream));
    return NWW_OK;
}

int nww_set_profiling(nww_engine* e, int enable) {
    if (!e) return fail(NWW_EINVAL, "nww_set_profiling: null engine");
    std::lock_guard<std::mutex> lock(e->mu);
    e->profiling = enable != 0;
    return NWW_OK;
}

int nww_get_profile(nww_engine* e, nww_profile_t* out) {
    if (!e || !out) return fail(NWW_EINVAL, "nww_get_profile: null argument");
    std::lock_guard<std::mutex> lock(e->mu);
    NWW_CUDA(cudaSetDevice(e->device));
    *out = nww_profile_t{};
    // spans come in (a, b) pairs that share the middle event: release each event once
    for (size_t i = 0; i < e->spans.size(); ++i) {
        const auto& sp = e->spans[i];
        NWW_CUDA(cudaEventSynchronize(sp.b));
        float ms = 0.f;
        NWW_CUDA(cudaEventElapsedTime(&ms, sp.a, sp.b));
        if (sp.stage == 0) { out->stage_a_ms += ms; out->stage_a_spans++; out->stage_a_windows += sp.windows; }
        else { out->stage_b_ms += ms; out->stage_b_spans++; out->stage_b_windows += sp.windows; }
    }
    for (size_t i = 0; i < e->spans.size(); ++i) {
        if (e->spans[i].stage == 0) e->event_pool.push_back(e->spans[i].a);
        e->event_pool.push_back(e->spans[i].b);
    }
    e->spans.clear();
    return NWW_OK;
}

// ---- pipe micro-benchmarks: the on-chip roofs bench.py quotes next to the HBM / tensor peaks ----------------------
int nww_microbench(int device, int kind, double* tflops) {
    // kind 0: FP32 FFMA, 1: FP64 DFMA.  Measured with CUDA events, best of 5; result in TFLOP/s (1 FMA = 2 flop).
    if (!tflops || (kind != 0 && kind != 1)) return fail(NWW_EINVAL, "nww_microbench: bad argument");
    NWW_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    NWW_CUDA(cudaGetDeviceProperties(&prop, device));
    void* d = nullptr;
    NWW_CUDA(cudaMalloc(&d, 64));
    cudaEvent_t a, b;
    NWW_CUDA(cudaEventCreate(&a));
    NWW_CUDA(cudaEventCreate(&b));
    const int grid = prop.multiProcessorCount * 8, iters = kind == 0 ? 8192 : 2048;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        NWW_CUDA(cudaEventRecord(a, 0));
        if (kind == 0) fma_peak_kernel<float><<<grid, 256>>>(static_cast<float*>(d), iters);
        else fma_peak_kernel<double><<<grid, 256>>>(static_cast<double*>(d), iters);
        NWW_CUDA(cudaEventRecord(b, 0));
        NWW_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        NWW_CUDA(cudaEventElapsedTime(&ms, a, b));
        const double fl = 2.0 * 16.0 * iters * 256.0 * grid;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);     // first launch = warm-up
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    *tflops = best;
    return NWW_OK;
}

int nww_synchronize(nww_engine* e) {
    if (!e) return fail(NWW_EINVAL, "nww_synchronize: null engine");
    NWW_CUDA(cudaSetDevice(e->device));
    NWW_CUDA(cudaStreamSynchronize(e->copy_stream));
    NWW_CUDA(cudaStreamSynchronize(e->stream));
    return NWW_OK;
}

}  // extern "C"
