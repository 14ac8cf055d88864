// DEVELOPER TOOL: run rowgemm_kc_umma_kernel (K-chunked tcgen05 row GEMM, nww_rowgemm.cuh) on host threads with the
// functional UMMA / TMEM model of nww_tc.cuh, for (1) a plain matrix with residual + ReLU and (2) the overlapping-row /
// padded-output views the raw-audio front end uses, and compare with a float64 evaluation.  usage: sim_rowgemm
#define NWW_CPUSIM 1
#include <stdio.h>
#include <stdlib.h>
#include <random>
#include <vector>
#include "cuda_sim.h"
#define __shared__ static      // one block runs at a time in the host model
#include "../../nanowakeword_b200/csrc/nww_rowgemm.cuh"
using namespace nww;

static double run_case(const char* name, long long n_win, int rpw, long long a_win, long long a_row, int K, int N, int n_valid,
                       long long o_win, int o_pitch, int o_off, bool with_res, int grid) {
    std::mt19937 rng(7);
    std::normal_distribution<float> nd(0.f, 1.f);
    const long long rows = n_win * rpw;
    std::vector<float> A((size_t)(n_win * a_win + rpw * a_row + K + 64)), W((size_t)K * N), bias(N), res,
        out((size_t)(n_win * o_win + (long long)(rpw + o_off) * o_pitch + 64), -7777.f);
    for (auto& v : A) v = 3.f * nd(rng);
    for (auto& v : W) v = 0.1f * nd(rng);
    for (auto& v : bias) v = nd(rng);
    if (with_res) { res.resize((size_t)rows * N); for (auto& v : res) v = nd(rng); }
    std::vector<uint16_t> wq;
    rowgemm_kc_pack(W.data(), K, N, &wq);
    const KcView av{rpw, a_win, a_row, 0}, ov{rpw, o_win, o_pitch, o_off};
    cudasim::launch(dim3(grid), dim3(kKcNT), rowgemm_kc_smem_bytes(), [&] {
        rowgemm_kc_umma_kernel(A.data(), av, K, reinterpret_cast<const uint4*>(wq.data()), bias.data(), with_res ? res.data() : nullptr,
                               out.data(), ov, rows, N, n_valid, 1);
    });
    double worst = 0;
    for (long long r = 0; r < rows; ++r) {
        const long long w = r / rpw, t = r % rpw;
        for (int c = 0; c < n_valid; ++c) {
            double s = bias[c];
            for (int k = 0; k < K; ++k) s += (double)A[w * a_win + t * a_row + k] * W[(size_t)k * N + c];
            if (with_res) s += res[r * N + c];
            s = s > 0 ? s : 0;
            const double got = out[w * o_win + (t + o_off) * o_pitch + c];
            worst = std::max(worst, fabs(got - s));
        }
    }
    // untouched: pad rows of the output view
    long long dirty = 0;
    for (long long w = 0; w < n_win; ++w)
        for (long long i = 0; i < (long long)o_off * o_pitch; ++i) dirty += out[w * o_win + i] != -7777.f;
    printf("%-28s rows %lld K %d N %d  max |err| %.3e  pad rows touched %lld\n", name, rows, K, N, worst, dirty);
    return worst + (double)dirty;
}

int main() {
    double bad = 0;
    bad += run_case("plain + residual", 1, 300, 0, 128, 128, 128, 128, 0, 128, 0, true, 2);
    bad += run_case("conv view k13 s4 c32 -> 64", 3, 25, 4000, 128, 448, 64, 64, 2400, 64, 6, false, 2);
    bad += run_case("conv view, 32 of 64 columns", 2, 40, 1000, 16, 64, 64, 32, 1500, 32, 6, false, 1);
    bad += run_case("wide N 512, K 256", 1, 130, 0, 256, 256, 512, 512, 0, 512, 0, false, 1);
    return bad < 5e-3 ? 0 : 1;
}
