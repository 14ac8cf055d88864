// DEVELOPER TOOL: run stage A + dense tail of one head on host threads.
// usage: sim_model <arch> <dir>   reads <dir>/blob.bin, pcm.i16 ; writes mel.f32 feat.f32 emb.f32 logits.f32
#define NWW_CPUSIM 1
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "../../nanowakeword_b200/csrc/nww_cnn.cuh"
#include "../../nanowakeword_b200/csrc/nww_tail.cuh"
#include "../../nanowakeword_b200/csrc/nww_tables.h"
#include "../../nanowakeword_b200/csrc/nww_blob.h"
using namespace nww;

template <typename V> std::vector<V> slurp(const std::string& p) {
    FILE* f = fopen(p.c_str(), "rb");
    if (!f) { perror(p.c_str()); exit(1); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<V> v(n / sizeof(V));
    if (fread(v.data(), 1, n, f) != (size_t)n) exit(1);
    fclose(f);
    return v;
}
template <typename V> void dump(const std::string& p, const std::vector<V>& v) {
    FILE* f = fopen(p.c_str(), "wb"); fwrite(v.data(), sizeof(V), v.size(), f); fclose(f);
}

template <typename T, typename G> struct Tabs {
    HostFrontendTables h; std::vector<T> ws, wu; std::vector<cplx<T>> tw; FrontendTables<T> dev;
    Tabs(const Blob& b) {
        std::string err; int rad[4] = {G::R0, G::R1, G::R2, G::R3};
        if (!build_frontend_tables(G::N_FFT, G::WIN, G::N_MELS, rad, G::N_PASS, b.f32("frontend.window"), b.f32("frontend.fb"), &h, &err)) { fprintf(stderr, "%s\n", err.c_str()); exit(1); }
        ws.assign(h.window_scaled.begin(), h.window_scaled.end()); wu.assign(h.window_unscaled.begin(), h.window_unscaled.end());
        tw.resize(G::N_FFT); for (int i = 0; i < G::N_FFT; ++i) tw[i] = {(T)h.tw_re[i], (T)h.tw_im[i]};
        dev = FrontendTables<T>{ws.data(), wu.data(), tw.data(), h.binpos.data(), h.mel_start.data(), h.mel_count.data(), h.mel_woff.data(), h.mel_w.data(), 1e-10f, -100.0f, h.mel_vec_ok};
    }
};

TailParams make_tail(const Blob& b) {
    TailParams P{}; 
    const int n = *reinterpret_cast<const int*>(b.base + b.find("tail.n_layers")->offset);
    P.n_layers = n; P.act = 0; P.max_width = 1;
    for (int i = 0; i < n; ++i) {
        std::string p = "tail." + std::to_string(i);
        const BlobTensor* w = b.find(p + ".W");
        TailLayer& L = P.layers[i];
        L.W = b.f32(p + ".W"); L.b = b.f32(p + ".b"); L.ln_g = b.f32(p + ".ln_g"); L.ln_b = b.f32(p + ".ln_b");
        L.N = w->dims[0]; L.K = w->dims[1];
        L.post = *reinterpret_cast<const int*>(b.base + b.find(p + ".post")->offset);
        if (L.N > P.max_width) P.max_width = L.N;
        if (i > 0 && L.K > P.max_width) P.max_width = L.K;
    }
    return P;
}

int main(int argc, char** argv) {
    std::string arch = argv[1], dir = argv[2];
    auto blobv = slurp<unsigned char>(dir + "/blob.bin");
    auto pcm = slurp<int16_t>(dir + "/pcm.i16");
    Blob b; std::string err;
    if (!parse_blob(blobv.data(), blobv.size(), &b, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    const long long nw = pcm.size() / 16000;
    std::vector<float> feat, mel;
    if (arch == "cnn") {
        using G = GeoNS40x98; using T = double; constexpr int NFB = 7, NT = 128;
        Tabs<T, G> tabs(b);
        CnnWeights wt{b.f32("cnn.w1"), b.f32("cnn.b1"), b.f32("cnn.w2"), b.f32("cnn.b2")};
        feat.assign(nw * CnnDims<G>::FEAT, -7777.f); mel.assign(nw * 40 * 98, -7777.f);
        cudasim::launch(dim3(3), dim3(NT), CnnSmem<T, G, NFB>::kTotal, [&] {
            cnn_stage_kernel<T, G, NFB, NT>(WindowSource{pcm.data(), nullptr, 16000}, nw, tabs.dev, wt, 0, feat.data(), mel.data());
        });
    } else { fprintf(stderr, "arch?\n"); return 1; }
    TailParams P = make_tail(b);
    std::vector<float> scores(nw), logits(nw), emb(nw * P.layers[P.n_layers - 3].N);
    cudasim::launch(dim3(2), dim3(kTailNT), tail_smem_bytes(P.max_width), [&] {
        tail_kernel(feat.data(), nw, P, scores.data(), logits.data(), emb.data());
    });
    dump(dir + "/mel.f32", mel); dump(dir + "/feat.f32", feat); dump(dir + "/emb.f32", emb);
    dump(dir + "/logits.f32", logits); dump(dir + "/scores.f32", scores);
    return 0;
}
