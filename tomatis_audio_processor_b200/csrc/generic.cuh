// General FFT sizes.  The reference exposes --n_fft / --hop
// (src/process_tomatis.py:509-510); the fused kernels are specialised for 4096 / 2048.  This is the plain path for every
// other power-of-two n_fft in [128, 8192] and any hop in [1, n_fft]: per-frame levels in NumPy's pairwise order, one CTA per
// frame for window -> FFT -> gain -> IFFT -> window in double precision with the reference's float32 roundings on either side
// (the arithmetic of edge_kernel), frames spilled to a scratch buffer, a gather overlap-add that adds the frames in the
// reference's order, and the limiter.  Built for coverage, not for speed: ~3x the HBM traffic of the fused kernel.
// Per-thread pieces are __host__ __device__ so that csrc/host_emul.cu runs them on the CPU.
#pragma once
#include "fft4096.cuh"

namespace tmt {

template <typename T> struct GenArith;
template <> struct GenArith<float> {
    static TMT_HD float add(float a, float b) {
#ifdef __CUDA_ARCH__
        return __fadd_rn(a, b);
#else
        return a + b;
#endif
    }
    // mono^2 of one sample-frame: mono = sqrt(mean(frame**2, axis=1)) (src/process_tomatis.py:370), all float32
    static TMT_HD float msq(float2 x, float sc, bool mono_file) {
#ifdef __CUDA_ARCH__
        const float l = __fmul_rn(x.x, sc), r = __fmul_rn(x.y, sc);
        const float h = mono_file ? __fmul_rn(l, l) : __fmul_rn(__fadd_rn(__fmul_rn(l, l), __fmul_rn(r, r)), 0.5f);
        const float m = __fsqrt_rn(h);
        return __fmul_rn(m, m);
#else
        const float l = x.x * sc, r = x.y * sc;
        const float ll = l * l, rr = r * r;
        const float h = mono_file ? ll : (ll + rr) * 0.5f;
        const float m = sqrtf(h);
        return m * m;
#endif
    }
    static TMT_HD float mean(float s, int n) { return s / (float)n; }      // n is a power of two: exact scaling
};
template <> struct GenArith<double> {
    static TMT_HD double add(double a, double b) {
#ifdef __CUDA_ARCH__
        return __dadd_rn(a, b);
#else
        return a + b;
#endif
    }
    static TMT_HD double msq(float2 x, float sc, bool mono_file) {
#ifdef __CUDA_ARCH__
        const double l = __dmul_rn((double)x.x, (double)sc), r = __dmul_rn((double)x.y, (double)sc);
        const double h = mono_file ? __dmul_rn(l, l) : __dmul_rn(__dadd_rn(__dmul_rn(l, l), __dmul_rn(r, r)), 0.5);
        const double m = __dsqrt_rn(h);
        return __dmul_rn(m, m);
#else
        const double l = (double)x.x * (double)sc, r = (double)x.y * (double)sc;
        const double ll = l * l, rr = r * r;
        const double h = mono_file ? ll : (ll + rr) * 0.5;
        const double m = sqrt(h);
        return m * m;
#endif
    }
    static TMT_HD double mean(double s, int n) { return s / (double)n; }
};

TMT_HD float2 gen_sample(const float2* x, long long total, long long p) {
    return (p >= 0 && p < total) ? x[p] : make_float2(0.f, 0.f);
}

// np.mean(mono * mono) of the frame [pos0, pos0 + n) (zeros outside the file) in NumPy's pairwise order: 128-element leaves
// with 8 strided accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), leaves combined by a balanced binary tree
// (n a power of two >= 128, so every split is an exact halving).
template <typename T>
TMT_HD T gen_frame_meansq(const float2* x, long long total, long long pos0, int n, float sc, bool mono_file) {
    T st[8];
    int depth = 0;
    for (int leaf = 0; leaf < n / 128; ++leaf) {
        const long long base = pos0 + (long long)leaf * 128;
        T r[8];
        for (int j = 0; j < 8; ++j) r[j] = GenArith<T>::msq(gen_sample(x, total, base + j), sc, mono_file);
        for (int i = 8; i < 128; i += 8)
            for (int j = 0; j < 8; ++j) r[j] = GenArith<T>::add(r[j], GenArith<T>::msq(gen_sample(x, total, base + i + j), sc, mono_file));
        T s = GenArith<T>::add(GenArith<T>::add(GenArith<T>::add(r[0], r[1]), GenArith<T>::add(r[2], r[3])),
                               GenArith<T>::add(GenArith<T>::add(r[4], r[5]), GenArith<T>::add(r[6], r[7])));
        for (int m = leaf; m & 1; m >>= 1) s = GenArith<T>::add(st[--depth], s);      // binary-counter merge = balanced tree
        st[depth++] = s;
    }
    return GenArith<T>::mean(st[0], n);
}

// flavours of the frame arithmetic
constexpr int kGenStreaming = 0;     // standard / xfade: float32 frames, irfft(...).astype(float32) * win, out / (w2 + 1e-12)
constexpr int kGenAdaptiveF32 = 1;   // adaptive with pre-attenuation: float32 frames, fl32(irfft * win), y / max(norm, 1e-8)
constexpr int kGenAdaptiveF64 = 2;   // adaptive without: the whole pipeline in double

// windowed input of the transform: frame * win with the reference's roundings (src/process_tomatis.py:396, _adaptive.py:215,311)
TMT_HD void gen_input(float2 x, float sc, float w, int flavour, double* re, double* im) {
    if (flavour == kGenAdaptiveF64) {
        *re = (double)x.x * (double)sc * (double)w;
        *im = (double)x.y * (double)sc * (double)w;
    } else {
#ifdef __CUDA_ARCH__
        *re = (double)__fmul_rn(__fmul_rn(x.x, sc), w);
        *im = (double)__fmul_rn(__fmul_rn(x.y, sc), w);
#else
        const float a = x.x * sc, b = x.y * sc;
        const float aw = a * w, bw = b * w;
        *re = (double)aw;
        *im = (double)bw;
#endif
    }
}

// one output tap of a frame: (yr, yi) = the inverse transform (already divided by n) at tap i, w = win[i]
TMT_HD void gen_output_f32(double yr, double yi, float w, int flavour, float* ox, float* oy) {
    if (flavour == kGenAdaptiveF32) {           // float64 irfft * float32 win assigned into a float32 frame: one rounding
        *ox = (float)(yr * (double)w);
        *oy = (float)(yi * (double)w);
    } else {                                    // irfft(...).astype(float32) * win
#ifdef __CUDA_ARCH__
        *ox = __fmul_rn((float)yr, w);
        *oy = __fmul_rn((float)yi, w);
#else
        const float a = (float)yr, b = (float)yi;
        *ox = a * w;
        *oy = b * w;
#endif
    }
}

// Overlap-add of one output position s: the frames that cover it are added in increasing frame order, like
// out_buf[start:start+n_fft] += y frame after frame (src/process_tomatis.py:400-406, _adaptive.py:316-323); the window
// energy accumulates beside it in float32.  F = float2 / double2 frames; returns the normalised sample in A (float / double).
template <typename A, typename F>
TMT_HD void gen_ola_sample(const F* frames, const float* win, long long s, long long first_start, int n_fft, int hop, int n_frames,
                           bool clamp_norm, A* ox, A* oy) {
    const long long rel = s - first_start;                                        // >= 0 for every file position of interest
    long long k_hi = rel >= 0 ? rel / hop : -1;
    if (k_hi > n_frames - 1) k_hi = n_frames - 1;
    long long k_lo = rel - n_fft + 1 <= 0 ? 0 : (rel - n_fft + 1 + hop - 1) / hop;
    A ax = 0, ay = 0;
    float wsum = 0.f;
    for (long long k = k_lo; k <= k_hi; ++k) {
        const int i = (int)(rel - k * hop);
        const F v = frames[(size_t)k * n_fft + i];
        const float w = win[i];
#ifdef __CUDA_ARCH__
        wsum = __fadd_rn(wsum, __fmul_rn(w, w));
#else
        const float w2 = w * w;
        wsum = wsum + w2;
#endif
        ax = ax + (A)v.x;                                                         // A == element type of F: plain IEEE add
        ay = ay + (A)v.y;
    }
    if (clamp_norm) {                                                             // y / np.maximum(norm, 1e-8)
        const float nrm = wsum > 1e-8f ? wsum : 1e-8f;
        *ox = ax / (A)nrm;
        *oy = ay / (A)nrm;
    } else {                                                                      // out_buf / (w_buf + 1e-12), float32
#ifdef __CUDA_ARCH__
        const float den = __fadd_rn(wsum, 1e-12f);
#else
        const float den = wsum + 1e-12f;
#endif
        *ox = ax / (A)den;
        *oy = ay / (A)den;
    }
}

}  // namespace tmt
