// Host-side table builders shared by the CUDA library and the CPU emulation used in tests.
#pragma once
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

namespace tmt {

// twA[k1*16 + n2] = W256^(n2*k1), W = exp(-2*pi*i/256); double precision, rounded once.
inline std::vector<float2> build_twA() {
    std::vector<float2> t(256);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k1 = 0; k1 < 16; ++k1)
        for (int n2 = 0; n2 < 16; ++n2) {
            const double a = -two_pi * double((n2 * k1) % 256) / 256.0;
            t[k1 * 16 + n2] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    return t;
}

// twB[k2*256 + t] = W4096^(n3*(k1 + 16*k2)), k1 = t>>4, n3 = t&15.
inline std::vector<float2> build_twB() {
    std::vector<float2> t(4096);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k2 = 0; k2 < 16; ++k2)
        for (int th = 0; th < 256; ++th) {
            const int k1 = th >> 4, n3 = th & 15;
            const double a = -two_pi * double((n3 * (k1 + 16 * k2)) % 4096) / 4096.0;
            t[k2 * 256 + th] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    return t;
}

// Natural-order half-spectrum gain row g[0..2048] -> register-order full-spectrum row:
//   out[t*16 + j] = G[(t>>4) + 16*(t&15) + 256*j] / 4096,  G[k] = g[k] (k<=2048) else g[4096-k].
inline void permute_gain_row(const float* g_half, float* out) {
    for (int t = 0; t < 256; ++t)
        for (int j = 0; j < 16; ++j) {
            const int k = (t >> 4) + 16 * (t & 15) + 256 * j;
            const float g = (k <= 2048) ? g_half[k] : g_half[4096 - k];
            out[t * 16 + j] = g * (1.0f / 4096.0f);   // power of two: exact
        }
}

}  // namespace tmt
