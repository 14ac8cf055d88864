// DEVELOPER TOOL: run rowgemm_kc_umma_kernel (K-chunked tcgen05 row GEMM, nww_rowgemm.cuh) on host threads with the
// functional UMMA / TMEM model of nww_tc.cuh — plain matrices, the overlapping-row views of the strided Conv1d layers
// and the segmented / two-level views of the strided 3x3 Conv2d layers — and compare with a float64 evaluation.
// usage: sim_rowgemm
#define NWW_CPUSIM 1
#include <stdio.h>
#include <stdlib.h>
#include <random>
#include <vector>
#include "cuda_sim.h"
#include "../../nanowakeword_b200/csrc/nww_rowgemm.cuh"
using namespace nww;

static double run_case(const char* name, long long n_win, KcView av, KcSegs sg, int K, int N, int n_valid, KcView ov, bool with_res,
                       int act, int grid) {
    std::mt19937 rng(7);
    std::normal_distribution<float> nd(0.f, 1.f);
    const long long rows = n_win * av.rpw;
    const long long a_len = av.at(rows - 1) + (sg.seg_len < K ? (K / sg.seg_len + 1) * sg.seg_stride : 0) + K + 64;
    const long long o_len = ov.at(rows - 1) + N + 64;
    std::vector<float> A((size_t)a_len), W((size_t)K * N), bias(N), res, out((size_t)o_len, -7777.f);
    for (auto& v : A) v = 3.f * nd(rng);
    for (auto& v : W) v = 0.1f * nd(rng);
    for (auto& v : bias) v = nd(rng);
    if (with_res) { res.resize((size_t)rows * N); for (auto& v : res) v = nd(rng); }
    std::vector<uint16_t> wq;
    rowgemm_kc_pack(W.data(), K, N, &wq);
    cudasim::launch(dim3(grid), dim3(kKcBlock), rowgemm_kc_smem_bytes(), [&] {
        rowgemm_kc_umma_kernel<true>(A.data(), av, sg, K, reinterpret_cast<const uint4*>(wq.data()), bias.data(), with_res ? res.data() : nullptr,
                               out.data(), ov, rows, N, n_valid, act, N <= 128 ? 2 : kKcRing);
    });
    double worst = 0;
    std::vector<char> written(out.size(), 0);
    for (long long r = 0; r < rows; ++r) {
        const long long w = r / av.rpw, t = r % av.rpw, o = t / av.inner, i = t % av.inner;
        const long long a0 = w * av.win_stride + o * av.outer_stride + (i + av.row_off) * av.row_stride;
        const long long oo = t / ov.inner, oi = t % ov.inner;
        const long long o0 = w * ov.win_stride + oo * ov.outer_stride + (oi + ov.row_off) * ov.row_stride;
        for (int c = 0; c < n_valid; ++c) {
            double s = bias[c];
            for (int k = 0; k < sg.k_valid; ++k) {
                const int seg = k / sg.seg_len;
                s += (double)A[a0 + seg * sg.seg_stride + (k - seg * sg.seg_len)] * W[(size_t)k * N + c];
            }
            if (with_res) s += res[r * N + c];
            if (act == 1) s = s > 0 ? s : 0;
            if (act == 3) s = s / (1.0 + exp(-s));
            worst = std::max(worst, fabs((double)out[o0 + c] - s));
            written[o0 + c] = 1;
        }
    }
    long long dirty = 0;                                           // nothing outside the view's valid columns may be touched
    for (size_t i = 0; i < out.size(); ++i) dirty += !written[i] && out[i] != -7777.f;
    printf("%-34s rows %6lld K %4d N %3d (%3d stored)  max |err| %.3e  stray writes %lld\n", name, rows, K, N, n_valid, worst, dirty);
    return worst + (double)dirty;
}

int main() {
    double bad = 0;
    bad += run_case("plain + residual + ReLU", 1, kc_plain(300, 128), kc_one_seg(128), 128, 128, 128, kc_plain(300, 128), true, 1, 2);
    bad += run_case("conv1d k13 s4 c32 -> 64", 3, kc_seq(25, 4000, 128, 0), kc_one_seg(416), 448, 64, 64, kc_seq(25, 2400, 64, 6), false, 1, 2);
    bad += run_case("conv1d k41 s16, 32 of 64 columns", 2, kc_seq(40, 1000, 16, 0), kc_one_seg(41), 64, 64, 32, kc_seq(40, 1500, 32, 6), false, 1, 1);
    bad += run_case("wide N 512, K 256, no act", 1, kc_plain(130, 256), kc_one_seg(256), 256, 512, 512, kc_plain(130, 512), false, 0, 1);
    {   // 3x3 conv, stride 2, C_in 24 on a padded (13 x 17) image -> (5 x 7) outputs of 48 channels into a padded (7 x 9) image
        const int C = 24, Wp = 17, Hp = 13, Ho = 5, Wo = 7, Co = 48, Wp2 = 9, Hp2 = 7;
        KcView av{Ho * Wo, (long long)Hp * Wp * C, 2 * C, 0, Wo, 2LL * Wp * C};
        KcView ov{Ho * Wo, (long long)Hp2 * Wp2 * Co, Co, 0, Wo, (long long)Wp2 * Co};
        bad += run_case("conv2d 3x3 s2 c24 -> 48, SiLU", 3, av, KcSegs{3 * C, (long long)Wp * C, 9 * C}, 256, 64, Co, ov, false, 3, 2);
    }
    {   // the plain fast path (no views) against the same reference
        const long long rows = 300;
        const int K = 192, N = 256;
        std::mt19937 rng(9);
        std::normal_distribution<float> nd(0.f, 1.f);
        std::vector<float> A(rows * K), W((size_t)K * N), bias(N), res(rows * N), out(rows * N, -7777.f);
        for (auto& v : A) v = 3.f * nd(rng);
        for (auto& v : W) v = 0.1f * nd(rng);
        for (auto& v : bias) v = nd(rng);
        for (auto& v : res) v = nd(rng);
        std::vector<uint16_t> wq;
        rowgemm_kc_pack(W.data(), K, N, &wq);
        cudasim::launch(dim3(2), dim3(kKcBlock), rowgemm_kc_smem_bytes(), [&] {
            rowgemm_kc_umma_kernel<false>(A.data(), kc_plain(rows, K), kc_one_seg(K), K, reinterpret_cast<const uint4*>(wq.data()), bias.data(),
                                          res.data(), out.data(), kc_plain(rows, N), rows, N, N, 1, kKcRing);
        });
        double worst = 0;
        for (long long r = 0; r < rows; ++r)
            for (int c = 0; c < N; ++c) {
                double s = bias[c] + res[r * N + c];
                for (int k = 0; k < K; ++k) s += (double)A[r * K + k] * W[(size_t)k * N + c];
                worst = std::max(worst, fabs((double)out[r * N + c] - (s > 0 ? s : 0)));
            }
        printf("%-34s rows %6lld K %4d N %3d               max |err| %.3e\n", "plain fast path", rows, K, N, worst);
        bad += worst;
    }
    return bad < 5e-3 ? 0 : 1;
}
