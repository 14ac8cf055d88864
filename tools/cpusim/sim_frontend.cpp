// DEVELOPER TOOL: run frontend_kernel on host threads.  usage: sim_frontend <geom NS|REF> <f32|f64> <dir>
// reads <dir>/pcm.i16, window.f32, fb.f32 ; writes <dir>/mel.f32  (F,T layout)
#define NWW_CPUSIM 1
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "../../nanowakeword_b200/csrc/nww_stage.cuh"
#include "../../nanowakeword_b200/csrc/nww_tables.h"
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

template <typename T, typename G, int NFB> void run(const std::string& dir) {
    auto pcm = slurp<int16_t>(dir + "/pcm.i16");
    auto win = slurp<float>(dir + "/window.f32");
    auto fb = slurp<float>(dir + "/fb.f32");
    const long long nw = pcm.size() / G::CLIP;
    HostFrontendTables h;
    std::string err;
    int rad[4] = {G::R0, G::R1, G::R2, G::R3};
    if (!build_frontend_tables(G::N_FFT, G::WIN, G::N_MELS, rad, G::N_PASS, win.data(), fb.data(), &h, &err)) {
        fprintf(stderr, "%s\n", err.c_str()); exit(1);
    }
    std::vector<T> ws(h.window_scaled.begin(), h.window_scaled.end()), wu(h.window_unscaled.begin(), h.window_unscaled.end());
    std::vector<cplx<T>> tw(G::N_FFT);
    for (int i = 0; i < G::N_FFT; ++i) tw[i] = {(T)h.tw_re[i], (T)h.tw_im[i]};
    FrontendTables<T> tab{ws.data(), wu.data(), tw.data(), h.binpos.data(), h.mel_start.data(), h.mel_count.data(),
                          h.mel_woff.data(), h.mel_w.data(), 1e-10f, -100.0f, h.mel_vec_ok};
    std::vector<float> mel((size_t)nw * G::N_MELS * G::N_FRAMES, -7777.f);
    constexpr int NT = 128;
    cudasim::launch(dim3(2), dim3(NT), FrontendSmem<T, G, NFB>::kTotal, [&] {
        frontend_kernel<T, G, NFB, NT>(WindowSource{pcm.data(), nullptr, G::CLIP}, nw, tab, mel.data(), 0);
    });
    FILE* f = fopen((dir + "/mel.f32").c_str(), "wb");
    fwrite(mel.data(), sizeof(float), mel.size(), f);
    fclose(f);
}

int main(int argc, char** argv) {
    std::string g = argv[1], p = argv[2], dir = argv[3];
    if (g == "NS" && p == "f32") run<float, GeoNS40x98, 13>(dir);
    else if (g == "NS") run<double, GeoNS40x98, 7>(dir);
    else if (p == "f32") run<float, GeoREF64x101, 13>(dir);
    else run<double, GeoREF64x101, 9>(dir);
    return 0;
}
