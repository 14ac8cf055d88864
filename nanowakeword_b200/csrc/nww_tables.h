// nww_tables.h — host-side construction of the front end's constant tables (plain C++).
//
// The Hann window and the mel filterbank come from the caller as the float32 tables the
// reference model carries (torchaudio buffers, reference modules/architectures.py:830-836);
// twiddles, the digit-reversal map of the in-place FFT and the sparse view of the filterbank
// are derived here in double precision.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string>
#include <vector>

namespace nww {

struct HostFrontendTables {
    int n_fft = 0, win = 0, n_mels = 0, n_freqs = 0;
    std::vector<double> window_scaled;     // window * 2^-15
    std::vector<double> window_unscaled;
    std::vector<double> tw_re, tw_im;      // exp(-2 pi i k / N)
    std::vector<uint16_t> binpos;
    std::vector<int> mel_start, mel_count, mel_woff;
    std::vector<float> mel_w;
    int mel_vec_ok = 0;                    // all filters fit the padded row of the vectorised mel (see below)
};

// radices[] lists the DIF pass radices in execution order; their product must be n_fft.
inline bool build_frontend_tables(int n_fft, int win, int n_mels, const int* radices, int n_radices,
                                  const float* window_f32, const float* fb_f32 /*[n_freqs][n_mels]*/,
                                  HostFrontendTables* out, std::string* err) {
    const int n_freqs = n_fft / 2 + 1;
    int prod = 1;
    for (int i = 0; i < n_radices; ++i) prod *= radices[i];
    if (prod != n_fft) {
        if (err) *err = "radix product does not match n_fft";
        return false;
    }
    out->n_fft = n_fft;
    out->win = win;
    out->n_mels = n_mels;
    out->n_freqs = n_freqs;
    out->window_scaled.resize(win);
    out->window_unscaled.resize(win);
    for (int i = 0; i < win; ++i) {
        out->window_unscaled[i] = (double)window_f32[i];
        out->window_scaled[i] = (double)window_f32[i] / 32768.0;
    }
    out->tw_re.resize(n_fft);
    out->tw_im.resize(n_fft);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k = 0; k < n_fft; ++k) {
        const double a = -two_pi * (double)k / (double)n_fft;
        out->tw_re[k] = cos(a);
        out->tw_im[k] = sin(a);
    }
    // k = q0 + R0*(q1 + R1*(q2 + ...))  sits at  q0*N/R0 + q1*N/(R0 R1) + ...
    out->binpos.resize(n_fft);
    for (int k = 0; k < n_fft; ++k) {
        int rem = k, pos = 0, block = n_fft;
        for (int i = 0; i < n_radices; ++i) {
            const int q = rem % radices[i];
            rem /= radices[i];
            block /= radices[i];
            pos += q * block;
        }
        out->binpos[k] = (uint16_t)pos;
    }
    out->mel_start.assign(n_mels, 0);
    out->mel_count.assign(n_mels, 0);
    out->mel_woff.assign(n_mels, 0);
    out->mel_w.clear();
    for (int m = 0; m < n_mels; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < n_freqs; ++k)
            if (fb_f32[(size_t)k * n_mels + m] != 0.0f) {
                if (first < 0) first = k;
                last = k;
            }
        out->mel_woff[m] = (int)out->mel_w.size();
        if (first < 0) continue;     // an all-zero filter (torchaudio warns about these) stays empty
        out->mel_start[m] = first;
        out->mel_count[m] = last - first + 1;
        for (int k = first; k <= last; ++k) out->mel_w.push_back(fb_f32[(size_t)k * n_mels + m]);
    }
    // the padded-row mel of the warp-private front ends: 40 filters x 32 bins (nww_fe2.cuh, n_fft 512) or
    // 64 filters x 20 bins (nww_fe5.cuh, n_fft 400); other tables take the generic front end
    const int row = (n_fft == 400) ? 20 : 32;
    out->mel_vec_ok = (n_fft == 400) ? (n_mels == 64) : (n_mels <= 40);
    for (int m = 0; m < n_mels; ++m)
        if ((out->mel_start[m] & 3) + out->mel_count[m] > row) out->mel_vec_ok = 0;
    return true;
}

}  // namespace nww
