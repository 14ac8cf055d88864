// DEVELOPER TOOL: run the v2 CNN stage (fe2 front end + conv1 + tcgen05-modelled conv2) and the dense
// tail on host threads.  usage: sim_cnn2 <dir>   reads <dir>/blob.bin, pcm.i16 ; writes mel/feat/emb/logits/scores .f32
#define NWW_CPUSIM 1
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "../../nanowakeword_b200/csrc/nww_cnn2.cuh"
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

int main(int argc, char** argv) {
    std::string dir = argv[1];
    auto blobv = slurp<unsigned char>(dir + "/blob.bin");
    auto pcm = slurp<int16_t>(dir + "/pcm.i16");
    Blob b; std::string err;
    if (!parse_blob(blobv.data(), blobv.size(), &b, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    const long long nw = pcm.size() / 16000;
    using G = GeoNS40x98;
    HostFrontendTables h; int rad[4] = {8, 8, 8, 1};
    if (!build_frontend_tables(G::N_FFT, G::WIN, G::N_MELS, rad, 3, b.f32("frontend.window"), b.f32("frontend.fb"), &h, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    std::vector<double> ws(h.window_scaled.begin(), h.window_scaled.end()), wu(h.window_unscaled.begin(), h.window_unscaled.end());
    std::vector<cplx<double>> tw(512);
    for (int i = 0; i < 512; ++i) tw[i] = {h.tw_re[i], h.tw_im[i]};
    FrontendTables<double> tab{ws.data(), wu.data(), tw.data(), h.binpos.data(), h.mel_start.data(), h.mel_count.data(), h.mel_woff.data(), h.mel_w.data(), 1e-10f, -100.0f, h.mel_vec_ok};

    // conv2 weights as UMMA operands (same code as nww_create)
    const float* w2 = b.f32("cnn.w2");
    auto bf16_rn = [](float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x7FFFu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); };
    auto bf16_f = [](uint16_t v) { uint32_t u = (uint32_t)v << 16; float f; memcpy(&f, &u, 4); return f; };
    std::vector<uint16_t> wb(Cnn2::W2_BYTES / 2);
    for (int tap = 0; tap < 9; ++tap)
        for (int ic = 0; ic < 16; ++ic)
            for (int oc = 0; oc < 32; ++oc) {
                const float v = w2[(ic * 9 + tap) * 32 + oc];
                const uint16_t hi = bf16_rn(v), lo = bf16_rn(v - bf16_f(hi));
                const size_t base = (size_t)tap * 2 * (Cnn2::W2_TAP_BYTES / 2) + (size_t)(ic >> 3) * 256 + oc * 8 + (ic & 7);
                wb[base] = hi; wb[base + Cnn2::W2_TAP_BYTES / 2] = lo;
            }
    const float* w1 = b.f32("cnn.w1");
    std::vector<float> w1v(144);
    for (int oc = 0; oc < 16; ++oc)
        for (int tap = 0; tap < 9; ++tap) w1v[(oc >> 3) * 72 + tap * 8 + (oc & 7)] = w1[oc * 9 + tap];
    Cnn2Weights wt{w1v.data(), b.f32("cnn.b1"), reinterpret_cast<const uint4*>(wb.data()), b.f32("cnn.b2")};
    std::vector<float> fhi(nw * 7680, -7777.f), flo(nw * 7680, -7777.f), mel(nw * 40 * 98, -7777.f);
    cudasim::launch(dim3(3), dim3(Cnn2::NT), Cnn2::kTotal, [&] {
        cnn2_stage_kernel<ACT_RELU>(WindowSource{pcm.data(), nullptr, 16000}, Cnn2MelSource{nullptr, nullptr, 0}, nw, tab, wt, fhi.data(), flo.data(), mel.data());
    });
    // back to the reference's (oc, ph, pw) flatten order
    std::vector<float> feat(nw * 7680);
    for (long long w = 0; w < nw; ++w)
        for (int oc = 0; oc < 32; ++oc)
            for (int ph = 0; ph < 10; ++ph)
                for (int pw = 0; pw < 24; ++pw) {
                    const size_t k2 = (size_t)w * 7680 + (ph * 24 + pw) * 32 + oc;
                    feat[(size_t)w * 7680 + (oc * 10 + ph) * 24 + pw] = fhi[k2] + flo[k2];
                }
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
    std::vector<float> scores(nw), logits(nw), emb(nw * P.layers[P.n_layers - 3].N);
    cudasim::launch(dim3(2), dim3(kTailNT), tail_smem_bytes(P.max_width), [&] {
        tail_kernel(feat.data(), nw, P, scores.data(), logits.data(), emb.data());
    });
    dump(dir + "/mel.f32", mel); dump(dir + "/feat.f32", feat); dump(dir + "/emb.f32", emb);
    dump(dir + "/logits.f32", logits); dump(dir + "/scores.f32", scores);
    return 0;
}
