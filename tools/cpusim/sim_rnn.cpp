// DEVELOPER TOOL: run rnn_seq_kernel (the fused GRU / LSTM sequence kernel, nww_rnn.cuh) on host threads with the
// functional UMMA / TMEM / bulk-copy model and compare [h_fwd(S-1) | h_bwd] with a float64 evaluation of the same
// packed gate matrices.  usage: sim_rnn
#define NWW_CPUSIM 1
#include <stdio.h>
#include <stdlib.h>
#include <random>
#include <vector>
#include "cuda_sim.h"
inline float __expf(float x) { return expf(x); }
#include "../../nanowakeword_b200/csrc/nww_rnn.cuh"
using namespace nww;

template <int CELL, int H> static double run_case(const char* name, int n, int S, int tm, int grid) {
    constexpr int IN = 40;
    using D = RnnDims<H, IN>;
    std::mt19937 rng(11);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> x((size_t)n * S * IN), wf((size_t)D::K * 4 * H, 0.f), wb((size_t)D::KX * 4 * H, 0.f), feat((size_t)n * 2 * H, -7.f);
    for (auto& v : x) v = 30.f * nd(rng) - 20.f;                   // log-mel-like magnitudes
    const float sc = 1.0f / sqrtf((float)H);
    auto fill = [&](std::vector<float>& w, int rows_h) {           // rows: x (IN), the constant-1 row, zero padding, h (rows_h)
        for (int k = 0; k <= IN; ++k)
            for (int c = 0; c < 4 * H; ++c) w[(size_t)k * 4 * H + c] = sc * nd(rng) * (k == IN ? 1.0f : 0.1f);
        for (int k = 0; k < rows_h; ++k)
            for (int c = 0; c < 4 * H; ++c) w[(size_t)(D::KX + k) * 4 * H + c] = sc * nd(rng);
        if (CELL == RNN_GRU) {                                     // n_x columns take no h rows, n_h columns no x rows
            for (int k = 0; k < IN; ++k)
                for (int j = 0; j < H; ++j) w[(size_t)k * 4 * H + 3 * H + j] = 0.f;
            for (int k = 0; k < rows_h; ++k)
                for (int j = 0; j < H; ++j) w[(size_t)(D::KX + k) * 4 * H + 2 * H + j] = 0.f;
        }
    };
    fill(wf, H);
    fill(wb, 0);
    std::vector<uint16_t> qf, qb;
    rnn_pack_weights(wf.data(), D::K, H, &qf);
    rnn_pack_weights(wb.data(), D::KX, H, &qb);
    cudasim::launch(dim3(grid), dim3(kRnnNT), D::SMEM, [&] {
        rnn_seq_kernel<CELL, H, IN>(x.data(), (long long)S * IN, S, n, tm, reinterpret_cast<const uint4*>(qf.data()),
                                    reinterpret_cast<const uint4*>(qb.data()), feat.data());
    });
    auto sig = [](double v) { return 1.0 / (1.0 + exp(-v)); };
    double worst = 0;
    for (int w = 0; w < n; ++w) {
        std::vector<double> h(H, 0.0), c(H, 0.0), out(2 * H);
        auto step = [&](const float* xt, const std::vector<float>& W, bool use_h, std::vector<double>& hh, std::vector<double>& cc) {
            std::vector<double> g(4 * H, 0.0);
            for (int col = 0; col < 4 * H; ++col) {
                double s = W[(size_t)IN * 4 * H + col];
                for (int k = 0; k < IN; ++k) s += (double)xt[k] * W[(size_t)k * 4 * H + col];
                if (use_h)
                    for (int k = 0; k < H; ++k) s += hh[k] * W[(size_t)(D::KX + k) * 4 * H + col];
                g[col] = s;
            }
            for (int j = 0; j < H; ++j) {
                if (CELL == RNN_LSTM) {
                    cc[j] = sig(g[H + j]) * cc[j] + sig(g[j]) * tanh(g[2 * H + j]);
                    hh[j] = sig(g[3 * H + j]) * tanh(cc[j]);
                } else {
                    const double r = sig(g[j]), z = sig(g[H + j]);
                    hh[j] = (1 - z) * tanh(g[2 * H + j] + r * g[3 * H + j]) + z * hh[j];
                }
            }
        };
        for (int s = 0; s < S; ++s) step(&x[((size_t)w * S + s) * IN], wf, true, h, c);
        std::vector<double> hb(H, 0.0), cb(H, 0.0);
        step(&x[((size_t)w * S + S - 1) * IN], wb, false, hb, cb);
        for (int j = 0; j < H; ++j) {
            worst = std::max(worst, fabs(feat[(size_t)w * 2 * H + j] - h[j]));
            worst = std::max(worst, fabs(feat[(size_t)w * 2 * H + H + j] - hb[j]));
        }
    }
    printf("%-22s n %3d S %2d tile %3d  max |err| %.3e\n", name, n, S, tm, worst);
    return worst;
}

int main() {
    double bad = 0;
    bad += run_case<RNN_LSTM, 128>("LSTM H=128", 70, 4, 64, 2);
    bad += run_case<RNN_GRU, 128>("GRU H=128", 40, 3, 32, 2);
    bad += run_case<RNN_LSTM, 64>("LSTM H=64 (RNNModel)", 33, 3, 32, 1);
    return bad < 3e-4 ? 0 : 1;
}
