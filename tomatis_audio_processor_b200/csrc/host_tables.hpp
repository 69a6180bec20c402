// Host-side table builders shared by the CUDA library and the CPU emulation used in tests.
#pragma once
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "fft4096.cuh"

namespace tmt {

// Per-thread twiddle bases (see fft4096.cuh tw_pow), double precision rounded once:
//   base[4*t + 0] = W4096^t      base[4*t + 1] = W4096^(4t)      (stage A, thread t)
//   base[4*t + 2] = W256^n3      base[4*t + 3] = W256^(4*n3)     (stage B, thread t holds n3 = b_n3(t))
inline std::vector<float2> build_tw_bases() {
    std::vector<float2> b(4 * 256);
    const double two_pi = 6.283185307179586476925286766559;
    auto w = [&](int num, int den) {
        const double a = -two_pi * double(num % den) / double(den);
        return make_float2((float)std::cos(a), (float)std::sin(a));
    };
    for (int t = 0; t < 256; ++t) {
        b[4 * t + 0] = w(t, 4096);
        b[4 * t + 1] = w(4 * t, 4096);
        b[4 * t + 2] = w(b_n3(t), 256);
        b[4 * t + 3] = w(4 * b_n3(t), 256);
    }
    return b;
}

// Stage-A twiddles, exact: tw[16*t + k] = W4096^(t*k), k = 0..15 (double precision rounded once).
inline std::vector<float2> build_tw_stage_a() {
    std::vector<float2> b(16 * 256);
    const double two_pi = 6.283185307179586476925286766559;
    for (int t = 0; t < 256; ++t)
        for (int k = 0; k < 16; ++k) {
            const double a = -two_pi * double((t * k) % 4096) / 4096.0;
            b[16 * t + k] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    return b;
}

// Natural-order half-spectrum gain row g[0..2048] -> register-order full-spectrum row:
//   out[t*16 + j] = G[bin_of(t, j)] / 4096,  G[k] = g[k] (k<=2048) else g[4096-k].
inline void permute_gain_row(const float* g_half, float* out) {
    for (int t = 0; t < 256; ++t)
        for (int j = 0; j < 16; ++j) {
            const int k = bin_of(t, j);
            const float g = (k <= 2048) ? g_half[k] : g_half[4096 - k];
            out[t * 16 + j] = g * (1.0f / 4096.0f);   // power of two: exact
        }
}

}  // namespace tmt
