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
inline std::vector<float2> build_tw_bases(bool pair = kPair) {
    std::vector<float2> b(4 * 256);
    const double two_pi = 6.283185307179586476925286766559;
    auto w = [&](int num, int den) {
        const double a = -two_pi * double(num % den) / double(den);
        return make_float2((float)std::cos(a), (float)std::sin(a));
    };
    for (int t = 0; t < 256; ++t) {             // pair mode (two 2048-point frames per pass): the parity bit drops out of both exponents
        const int ta = pair ? (t & ~1) : t, n3 = pair ? (b_n3(t) & ~1) : b_n3(t);
        b[4 * t + 0] = w(ta, 4096);
        b[4 * t + 1] = w(4 * ta, 4096);
        b[4 * t + 2] = w(n3, 256);
        b[4 * t + 3] = w(4 * n3, 256);
    }
    return b;
}

// Stage-A twiddles, exact: tw[16*t + k] = W4096^(t*k), k = 0..15 (double precision rounded once).
inline std::vector<float2> build_tw_stage_a(bool pair = kPair) {
    std::vector<float2> b(16 * 256);
    const double two_pi = 6.283185307179586476925286766559;
    for (int t = 0; t < 256; ++t)
        for (int k = 0; k < 16; ++k) {
            const double a = -two_pi * double(((pair ? (t & ~1) : t) * k) % 4096) / 4096.0;
            b[16 * t + k] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    return b;
}

// Natural-order half-spectrum gain row g[0..2048] -> register-order full-spectrum row:
//   out[t*16 + j] = G[bin_of(t, j)] / 4096,  G[k] = g[k] (k<=2048) else g[4096-k].
// Pair mode: g_half[0..1024] of a 2048-point frame; registers j and j + 8 hold the same bins of the pass's two frames, so
//   out[t*16 + j] = G[bin_of_pair(t, j)] / 2048 (the kernel takes j < 8 from the first frame's row and j >= 8 from the second's).
inline void permute_gain_row(const float* g_half, float* out, bool pair = kPair) {
    if (pair) {
        for (int t = 0; t < 256; ++t)
            for (int j = 0; j < 16; ++j) {
                const int k = bin_of_pair(t, j);
                out[t * 16 + j] = ((k <= 1024) ? g_half[k] : g_half[2048 - k]) * (1.0f / 2048.0f);
            }
        return;
    }
    for (int t = 0; t < 256; ++t)
        for (int j = 0; j < 16; ++j) {
            const int k = bin_of(t, j);
            const float g = (k <= 2048) ? g_half[k] : g_half[4096 - k];
            out[t * 16 + j] = g * (1.0f / 4096.0f);   // power of two: exact
        }
}

}  // namespace tmt
