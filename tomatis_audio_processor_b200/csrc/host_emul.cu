// CPU emulation of one CTA of the fused STFT kernel (TEST SUPPORT, not shipped in the product .so).
// Runs the very same __host__ __device__ stage functions as the GPU kernel, thread by thread,
// so index maps / twiddles / layouts can be verified in the GPU-less build container.
#include <vector>
#include <cstring>
#include "fft4096.cuh"
#include "host_tables.hpp"

using namespace tmt;

extern "C" {

// z_in: 4096 complex (interleaved re,im); spec_out: 4096 complex in NATURAL bin order (forward only)
int tmt_emul_forward(const float* z_in, float* spec_out) {
    auto tb = build_tw_bases();
    auto WA = [&](int t) { return TwBase{tb[4 * t], tb[4 * t + 1]}; };
    auto WB = [&](int t) { return TwBase{tb[4 * t + 2], tb[4 * t + 3]}; };
    std::vector<float2> P(kExchFloat2), Q(kExchFloat2);
    std::vector<float2> regs(256 * 16);
    float2 v[16];
    for (int t = 0; t < 256; ++t) {
        for (int j = 0; j < 16; ++j) v[j] = make_float2(z_in[2 * (256 * j + t)], z_in[2 * (256 * j + t) + 1]);
        fwd_a(v, t, WA(t), P.data());
    }
    for (int t = 0; t < 256; ++t) fwd_b(v, t, WB(t), P.data(), Q.data());
    for (int t = 0; t < 256; ++t) {
        fwd_c(v, t, Q.data());
        for (int j = 0; j < 16; ++j) {
            const int k = bin_of(t, j);
            spec_out[2 * k] = v[j].x;
            spec_out[2 * k + 1] = v[j].y;
        }
    }
    return 0;
}

// Full per-frame operator: out = IFFT(gain * FFT(z_in)) with gain given as a natural-order
// half spectrum g[0..2048] (the 1/4096 is applied through the permuted gain row, as on the GPU).
int tmt_emul_filter(const float* z_in, const float* g_half, float* z_out) {
    auto tb = build_tw_bases();
    auto WA = [&](int t) { return TwBase{tb[4 * t], tb[4 * t + 1]}; };
    auto WB = [&](int t) { return TwBase{tb[4 * t + 2], tb[4 * t + 3]}; };
    std::vector<float> gperm(4096);
    permute_gain_row(g_half, gperm.data());
    std::vector<float2> P(kExchFloat2), Q(kExchFloat2);
    std::vector<float2> regs(256 * 16);
    float2 v[16];
    for (int t = 0; t < 256; ++t) {
        for (int j = 0; j < 16; ++j) v[j] = make_float2(z_in[2 * (256 * j + t)], z_in[2 * (256 * j + t) + 1]);
        fwd_a(v, t, WA(t), P.data());
    }
    for (int t = 0; t < 256; ++t) fwd_b(v, t, WB(t), P.data(), Q.data());
    for (int t = 0; t < 256; ++t) {           // C, gain, C' : registers only; C' writes P (padded layout)
        fwd_c(v, t, Q.data());
        for (int j = 0; j < 16; ++j) {
            const float g = gperm[t * 16 + j];
            v[j].x *= g;
            v[j].y *= g;
        }
        inv_c(v, t, P.data());
    }
    for (int t = 0; t < 256; ++t) inv_b(v, t, WB(t), P.data(), Q.data());
    for (int t = 0; t < 256; ++t) {
        inv_a(v, t, WA(t), Q.data());
        for (int j = 0; j < 16; ++j) {
            z_out[2 * (256 * j + t)] = v[j].x;
            z_out[2 * (256 * j + t) + 1] = v[j].y;
        }
    }
    return 0;
}

}  // extern "C"
