// Calibration front end (src/calibrate_to_baseline_v2.py): per-thread pieces of the four kernels,
// __host__ __device__ so that csrc/host_emul.cu can run them on the CPU.
//   envelope + polyphase decimation   power_mono (:8-11) and scipy.signal.resample_poly as find_delay_by_corr uses it (:60,73)
//   valid cross-correlation           fftconvolve(mo_ds, mb_ds[::-1], mode="valid") (:77-78), computed directly
//   band energies                     stft_band_tilt (:17-30)
//   gate grid                         simulate_state (:88-112) for one (threshold, hysteresis, delay) combination
#pragma once
#include "fft4096.cuh"
#include "spectrum.cuh"

namespace tmt {

// mono amplitude by power average, float32 like the reference: sqrt(0.5*(l*l + r*r) + 1e-12)
TMT_HD float power_mono(float2 s) {
    const float ll = s.x * s.x, rr = s.y * s.y;          // separate roundings: no fused multiply-add
#ifdef __CUDA_ARCH__
    const float p = __fmul_rn(0.5f, __fadd_rn(ll, rr));
    return __fsqrt_rn(__fadd_rn(p, 1e-12f));
#else
    const float p = 0.5f * (ll + rr);
    return sqrtf(p + 1e-12f);
#endif
}

// One output sample of upfirdn(h, env, up, down)[m] = sum_i h[m*down - i*up] * env[i], env = power_mono(x), zero extension
// (scipy's 'constant' mode).  h is the zero-padded filter resample_poly builds; accumulation in double.
TMT_HD float decimate_sample(const float2* x, long long n_in, const float* h, int len_h, int up, int down, long long m) {
    const long long q_top = m * (long long)down;           // h index of input sample i is q_top - i*up
    long long i_lo = (q_top - (len_h - 1) + up - 1) / up;   // ceil((q_top - len_h + 1) / up), numerator may be negative
    if (q_top - (len_h - 1) < 0) i_lo = 0;
    long long i_hi = q_top / up;
    if (i_hi > n_in - 1) i_hi = n_in - 1;
    double acc = 0.0;
    for (long long i = i_lo; i <= i_hi; ++i) acc += (double)h[q_top - i * up] * (double)power_mono(x[i]);
    return (float)acc;
}

// |rfft(mono * win)|^2 of one bin as the reference forms it from the complex64 spectrum: fl(fl(re*re) + fl(im*im))
TMT_HD float power_bin(const cplx64* Z, int k) {
    const float re = (float)Z[k].x, im = (float)Z[k].y;
    const float a = re * re, b = im * im;
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}

// simulate_state for one parameter combination over irregular frame positions; comparisons in float32 (NumPy compares a
// float32 level with a Python-float threshold in float32).  Returns mismatches against `want` and the number of switches;
// `states` (optional) receives the state of every frame.
TMT_HD void gate_grid_combo(const float* level, const long long* start, const unsigned char* want, int n, float on, float off,
                            long long delay, int* mismatches, int* switches, unsigned char* states = nullptr) {
    int state = 1, prev = 0, mis = 0, sw = 0;
    bool armed = false;
    long long pending = 0;
    for (int i = 0; i < n; ++i) {
        const float lv = level[i];
        const long long st = start[i];
        if (state == 1) {
            if (lv >= on) {
                if (!armed) { armed = true; pending = st + delay; }
            } else {
                armed = false;
            }
            if (armed && st >= pending) { state = 2; armed = false; }
        } else if (lv <= off) {
            state = 1;
            armed = false;
        }
        mis += (state != (int)want[i]);
        sw += (i > 0 && state != prev);
        prev = state;
        if (states) states[i] = (unsigned char)state;
    }
    *mismatches = mis;
    *switches = sw;
}

}  // namespace tmt
