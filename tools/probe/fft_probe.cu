// DEVELOPER TOOL: cycle-level probe of the warp-private packed FFT (nww_fe3.cuh) — cycles per FFT for 1..16 resident
// warps per SM and a per-phase breakdown (clock64 stamps around the passes of a copy of fe3_warp_fft).
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/probe/fft_probe.cu -o gpurun_out/fft_probe
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../nanowakeword_b200/csrc/nww_fe3.cuh"
#include "../../nanowakeword_b200/csrc/nww_tables.h"
using namespace nww;

__global__ void __launch_bounds__(512, 1)
probe_kernel(FrontendTables<double> tab, const int16_t* pcm, float* out, long long* cyc, int n_fft_per_warp, int variant) {
    NWW_DYN_SMEM(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
    cplx<double>* wb = reinterpret_cast<cplx<double>*>(smem) + (size_t)warp * Fe3::NPAD;
    cplx<double>* tw = reinterpret_cast<cplx<double>*>(smem + 16 * Fe3::NPAD * 16);
    double* win_s = reinterpret_cast<double*>(smem + 16 * Fe3::NPAD * 16 + Fe3::kTwBytes);
    int16_t* x = reinterpret_cast<int16_t*>(smem + 16 * Fe3::NPAD * 16 + Fe3::kTwBytes + Fe3::kWinBytes);
    fe2_build_tables(tw, tab, tid, blockDim.x);
    fe3_build_window(win_s, tab.window, tid, blockDim.x);
    for (int i = tid; i < 16000; i += blockDim.x) x[i] = pcm[i];
    __syncthreads();
    float acc = 0.f;
    const long long t0 = clock64();
    for (int i = 0; i < n_fft_per_warp; ++i) {
        const int f = (warp + i * nw) % 49;
        fe3_warp_fft(x + 320 * f, wb, win_s, tw, tab, [&](int fr, int m, float db) { acc += db; }, lane);
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * 16 + warp] = t1 - t0;
    if (acc == 12345.f) out[tid] = acc;
}

int main() {
    using G = GeoNS40x98;
    // tables from an analytic window / filterbank (values do not matter for timing)
    std::vector<float> win(G::WIN), fb((size_t)G::N_FREQS * G::N_MELS, 0.f);
    for (int i = 0; i < G::WIN; ++i) win[i] = 0.5f - 0.5f * cosf(2.f * 3.14159265f * i / G::WIN);
    for (int m = 0; m < G::N_MELS; ++m) {           // triangles of growing width, like HTK
        const int c = 3 + (int)(250.0 * (exp(m / 39.0 * 2.0) - 1) / (exp(2.0) - 1)), hw = 2 + m / 3;
        for (int k = c - hw; k <= c + hw; ++k)
            if (k >= 0 && k < G::N_FREQS) fb[(size_t)k * G::N_MELS + m] = 1.f - fabsf((float)(k - c)) / (hw + 1);
    }
    HostFrontendTables h; std::string err; const int rad[4] = {8, 8, 8, 1};
    if (!build_frontend_tables(G::N_FFT, G::WIN, G::N_MELS, rad, 3, win.data(), fb.data(), &h, &err)) { printf("%s\n", err.c_str()); return 1; }
    std::vector<double> ws(h.window_scaled.begin(), h.window_scaled.end()), wu(h.window_unscaled.begin(), h.window_unscaled.end());
    std::vector<cplx<double>> tw(512);
    for (int i = 0; i < 512; ++i) tw[i] = {h.tw_re[i], h.tw_im[i]};
    auto up = [](const void* p, size_t n) { void* d; cudaMalloc(&d, n); cudaMemcpy(d, p, n, cudaMemcpyHostToDevice); return d; };
    FrontendTables<double> tab{(const double*)up(ws.data(), ws.size() * 8), (const double*)up(wu.data(), wu.size() * 8),
                               (const cplx<double>*)up(tw.data(), tw.size() * 16), (const uint16_t*)up(h.binpos.data(), h.binpos.size() * 2),
                               (const int*)up(h.mel_start.data(), h.mel_start.size() * 4), (const int*)up(h.mel_count.data(), h.mel_count.size() * 4),
                               (const int*)up(h.mel_woff.data(), h.mel_woff.size() * 4), (const float*)up(h.mel_w.data(), h.mel_w.size() * 4),
                               1e-10f, -100.0f, h.mel_vec_ok};
    printf("mel_vec_ok %d\n", h.mel_vec_ok);
    std::vector<int16_t> pcm(16000);
    for (int i = 0; i < 16000; ++i) pcm[i] = (int16_t)((i * 7919) % 20011 - 10000);
    int16_t* d_pcm = (int16_t*)up(pcm.data(), 32000);
    float* d_out; cudaMalloc(&d_out, 4096);
    long long* d_cyc; cudaMalloc(&d_cyc, 148 * 16 * 8);
    const size_t smem = 16 * Fe3::NPAD * 16 + Fe3::kTwBytes + Fe3::kWinBytes + 32128;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int nf = 64;
    for (int nwarp : {1, 2, 4, 7, 8, 12, 16}) {
        probe_kernel<<<148, nwarp * 32, smem>>>(tab, d_pcm, d_out, d_cyc, nf, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> c(148 * 16);
        cudaMemcpy(c.data(), d_cyc, c.size() * 8, cudaMemcpyDeviceToHost);
        double s = 0; for (int b = 0; b < 148; ++b) for (int w = 0; w < nwarp; ++w) s += c[b * 16 + w];
        const double per = s / (148.0 * nwarp) / nf;
        printf("%2d warps/SM: %8.0f cycles per FFT per warp -> %7.0f cycles per FFT per SM\n", nwarp, per, per / nwarp);
    }
    return 0;
}
