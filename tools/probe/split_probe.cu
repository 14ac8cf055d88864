// DEVELOPER TOOL: do fe4_mel_kernel and conv4_mel_kernel (nww_cnn4.cuh) really share an SM?  Times each alone and both
// on two streams.   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a tools/probe/split_probe.cu -o tools/probe/split_probe.bin
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../nanowakeword_b200/csrc/nww_cnn4.cuh"
#include "../../nanowakeword_b200/csrc/nww_tables.h"
using namespace nww;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
int main() {
    using G = GeoNS40x98;
    std::vector<float> win(G::WIN), fb((size_t)G::N_FREQS * G::N_MELS, 0.f);
    for (int i = 0; i < G::WIN; ++i) win[i] = 0.5f - 0.5f * cosf(2.f * 3.14159265f * i / G::WIN);
    for (int m = 0; m < G::N_MELS; ++m) {
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
    const long long n = 148 * 14;
    int16_t* pcm; CK(cudaMalloc(&pcm, n * 32000)); CK(cudaMemset(pcm, 1, n * 32000));
    float *mel, *fh, *fl; CK(cudaMalloc(&mel, n * 3920 * 4)); CK(cudaMalloc(&fh, n * 7680 * 4)); CK(cudaMalloc(&fl, n * 7680 * 4));
    CK(cudaMemset(mel, 0, n * 3920 * 4));
    void* wz; CK(cudaMalloc(&wz, 65536)); CK(cudaMemset(wz, 0, 65536));
    Cnn2Weights wt{(const float*)wz, (const float*)wz, (const uint4*)wz, (const float*)wz};
    CK(cudaFuncSetAttribute(fe4_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Fe4::kTotal));
    CK(cudaFuncSetAttribute(conv4_mel_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Conv4::kTotal));
    CK(cudaFuncSetAttribute(fe4_mel_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute(conv4_mel_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int occ_f = 0, occ_c = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, fe4_mel_kernel, Fe4::NT, Fe4::kTotal);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, conv4_mel_kernel<0>, Conv4::NT, Conv4::kTotal);
    cudaFuncAttributes af, ac; cudaFuncGetAttributes(&af, fe4_mel_kernel); cudaFuncGetAttributes(&ac, conv4_mel_kernel<0>);
    printf("fe4: %d regs, %zu B smem, occupancy %d | conv4: %d regs, %zu B smem, occupancy %d\n", af.numRegs, (size_t)Fe4::kTotal, occ_f,
           ac.numRegs, (size_t)Conv4::kTotal, occ_c);
    {
        int v1 = 0, v2 = 0, v3 = 0, v4 = 0;
        cudaDeviceGetAttribute(&v1, cudaDevAttrMaxSharedMemoryPerMultiprocessor, 0);
        cudaDeviceGetAttribute(&v2, cudaDevAttrReservedSharedMemoryPerBlock, 0);
        cudaDeviceGetAttribute(&v3, cudaDevAttrMaxRegistersPerMultiprocessor, 0);
        cudaDeviceGetAttribute(&v4, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0);
        printf("smem per SM %d, reserved per block %d, regs per SM %d, optin per block %d\n", v1, v2, v3, v4);
        for (int kb = 40; kb <= 116; kb += 4) {
            int o = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, conv4_mel_kernel<0>, Conv4::NT, (size_t)kb * 1024);
            printf("conv4 occupancy at %d KB dynamic smem: %d\n", kb, o);
        }
    }
    cudaStream_t s1, s2; CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](int mode) -> float {      // 1 = fe only, 2 = conv only, 3 = both
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaDeviceSynchronize());
            cudaEventRecord(a, s1);
            cudaStreamWaitEvent(s2, a, 0);
            if (mode & 1) fe4_mel_kernel<<<148, Fe4::NT, Fe4::kTotal, s1>>>(pcm, n, tab, mel);
            if (mode & 2) conv4_mel_kernel<0><<<148, Conv4::NT, Conv4::kTotal, s2>>>(mel, n, wt, fh, fl);
            cudaEventRecord(b, s2);
            cudaStreamWaitEvent(s1, b, 0);
            cudaEventRecord(b, s1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep) best = fminf(best, ms);
        }
        return best;
    };
    const float tf = run(1), tc = run(2), tb = run(3);
    printf("%lld windows: fe4 alone %.3f ms (%.0f cycles/window/SM), conv4 alone %.3f ms (%.0f), both %.3f ms\n", n, tf,
           tf * 1e-3 * 1.965e9 / 14, tc, tc * 1e-3 * 1.965e9 / 14, tb);
    return 0;
}
