// nww_heads.cuh — stage A of the remaining heads (TCN, BcResNet, CRNN-GRU, E2E mel-CNN).
#pragma once

#include <functional>
#include <string>

#include "../../include/nww_b200.h"
#include "nww_stage.cuh"

namespace nww {

struct HeadWeights {
    int dummy = 0;
};

// int16 grid recovery for float PCM that was produced as int16 / 32768 (nanointerpreter.py:750).
__global__ void f32_to_i16_kernel(const float* __restrict__ x, int16_t* __restrict__ y, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = rintf(x[i] * 32768.0f);
        v = fminf(fmaxf(v, -32768.0f), 32767.0f);
        y[i] = (int16_t)v;
    }
}

inline int setup_head_weights(int arch, int geometry, const std::function<const float*(const char*, size_t)>& lookup,
                              HeadWeights* hw, int* feat_dim, size_t* scratch_per_window, std::string* err) {
    (void)geometry; (void)lookup; (void)hw; (void)feat_dim; (void)scratch_per_window;
    *err = "architecture id " + std::to_string(arch) + " is not built into this library yet";
    return NWW_EUNSUPPORTED;
}

inline int launch_head_stage_a(int arch, const HeadWeights& hw, const FrontendTables<double>& tab, int act, int sm_count,
                               const int16_t* pcm, long long n, float* feat, float* scratch, float* mel, cudaStream_t st,
                               int64_t* launches, std::string* err) {
    (void)hw; (void)tab; (void)act; (void)sm_count; (void)pcm; (void)n; (void)feat; (void)scratch; (void)mel; (void)st; (void)launches;
    *err = "architecture id " + std::to_string(arch) + " is not built into this library yet";
    return NWW_EUNSUPPORTED;
}

}  // namespace nww
