// Conditional-spectrum validator: per-frame magnitude ratio |Y| / |X| of an output / input
// file pair, as src/validate_layer1.py:261-389 and src/verify_tomatis_15db_v2.py:270-369 compute it
// (Hann window, rfft per channel, channel-averaged magnitudes, X floored at 1e-10, optional anchor-band gain).
// The transform is the fp64 shared-memory FFT of edge_kernel: an output file of the tilt filter spans ~70 dB inside one
// frame, where a float32 FFT (3 ulp of the largest bin) is 1e-2 dB off in the quiet bins, while NumPy's rfft of float32
// data is accurate to the final rounding of its complex64 result -- which is what is reproduced here: double spectrum,
// rounded once to float32 per component, float32 arithmetic from there on, in the reference's order.
// The per-thread pieces are __host__ __device__: csrc/host_emul.cu runs them on the CPU.
#pragma once
#include "fft4096.cuh"

namespace tmt {

constexpr int kBins = kNfft / 2 + 1;     // 2049 rfft bins

struct cplx64 { double x, y; };          // layout of double2 without needing vector types on the host side

// (frame[:, c] * win) in float32 for both channels, packed as L + iR and widened for the transform
TMT_HD cplx64 spec_window_sample(float2 s, float w) { return cplx64{(double)(s.x * w), (double)(s.y * w)}; }

// Z = FFT(L + iR)  ->  (|rfft(L)[k]| + |rfft(R)[k]|) / 2   (X += np.abs(...) per channel; X /= ch)
TMT_HD float spec_mean_mag(const cplx64* Z, int k) {
    const cplx64 a = Z[k], b = Z[(kNfft - k) & (kNfft - 1)];
    const float lx = (float)(0.5 * (a.x + b.x)), ly = (float)(0.5 * (a.y - b.y));       // L[k] = (Z[k] + conj(Z[N-k])) / 2
    const float rx = (float)(0.5 * (a.y + b.y)), ry = (float)(0.5 * (b.x - a.x));       // R[k] = (Z[k] - conj(Z[N-k])) / 2i
    const float ml = (float)sqrt((double)lx * (double)lx + (double)ly * (double)ly);    // np.abs(complex64)
    const float mr = (float)sqrt((double)rx * (double)rx + (double)ry * (double)ry);
    return (ml + mr) * 0.5f;
}

// bins of thread t: k = t + 256*i, i < n_fft/512 (8 at 4096, 4 at 2048), and the Nyquist bin n_fft/2 for t == 0
TMT_HD int spec_bin(int t, int i) { return t + 256 * i; }
TMT_HD int spec_bins_of_thread(int t) { return kNfft / 512 + (t == 0 ? 1 : 0); }

TMT_HD float spec_ratio(float ymag, float xmag) { return ymag / fmaxf(xmag, 1e-10f); }   // X = np.maximum(X, 1e-10); Y / X

// np.mean(ratio[anchor_mask]) over bins [a0, a1] (verify_tomatis_15db_v2.py:326)
TMT_HD float spec_anchor_gain(const float* ratio, int a0, int a1) {
    float s = 0.f;
    for (int k = a0; k <= a1; ++k) s += ratio[k];
    return s / (float)(a1 - a0 + 1);
}

}  // namespace tmt
