// CPU emulation of one CTA of the fused STFT kernel (TEST SUPPORT, not shipped in the product .so).
// Runs the very same __host__ __device__ stage functions as the GPU kernel, thread by thread,
// so index maps / twiddles / layouts can be verified in the GPU-less build container.
#include <vector>
#include <cmath>
#include <cstring>
#include "fft4096.cuh"
#include "spectrum.cuh"
#include "calib.cuh"
#include "generic.cuh"
#include "host_tables.hpp"

using namespace tmt;

extern "C" {

// The stage sequence of stft_kernel for one frame, thread by thread, with the tensor-memory round trips of exchange E2 replaced
// by their data movement (emul_trip_fwd / emul_trip_inv, verified on hardware by tools/mb_tmem_xpose.cu).
// spec: if non-null, receives the spectrum in NATURAL bin order (before the gain); g_half: natural-order half-spectrum gain or
// null (forward only).
// pair: two 2048-point frames as the even / odd samples of z (fft4096.cuh, pair mode); spec then holds frame 0's 2048 bins
// followed by frame 1's, g_half = two half spectra of 1025 gains (frame 0, frame 1)
static void emul_frame(const float* z_in, const float* g_half, float* spec, float* z_out, bool pair = false) {
    auto tb = build_tw_bases(pair);
    auto WA = [&](int t) { return TwBase{tb[4 * t], tb[4 * t + 1]}; };
    auto WB = [&](int t) { return TwBase{tb[4 * t + 2], tb[4 * t + 3]}; };
    std::vector<float> gperm(4096, pair ? 1.0f / 2048.0f : 1.0f);
    if (g_half && !pair) permute_gain_row(g_half, gperm.data(), false);
    if (g_half && pair) {                         // registers 0-7 from the first frame's row, 8-15 from the second's
        std::vector<float> g0(4096), g1(4096);
        permute_gain_row(g_half, g0.data(), true);
        permute_gain_row(g_half + 1025, g1.data(), true);
        for (int t = 0; t < 256; ++t)
            for (int j = 0; j < 16; ++j) gperm[t * 16 + j] = (j < 8 ? g0 : g1)[t * 16 + j];
    }
    std::vector<float2> P(kE1Float2);
    std::vector<float> R(256 * 32), S(256 * 32);
    float2 v[16];
    for (int t = 0; t < 256; ++t) {                                   // A
        for (int j = 0; j < 16; ++j) v[j] = make_float2(z_in[2 * (256 * j + t)], z_in[2 * (256 * j + t) + 1]);
        dft16<false>(v);
        tw_pow<false>(v, WA(t));
        st_e1a(v, t, P.data());
    }
    for (int t = 0; t < 256; ++t) {                                   // B
        ld_e1b(v, t, P.data());
        dft16<false>(v);
        tw_pow<false>(v, WB(t));
        float r[32];
        x_fwd1_pack(v, r);
        for (int c = 0; c < 32; ++c) R[t * 32 + c] = r[c];
    }
    for (int w = 0; w < 8; ++w) emul_trip_fwd(R.data() + w * 1024, S.data() + w * 1024);
    for (int t = 0; t < 256; ++t) {                                   // first layer of C
        float r[32], s[32];
        for (int c = 0; c < 32; ++c) r[c] = S[t * 32 + c];
        x_layer_a<false>(r, s);
        for (int c = 0; c < 32; ++c) R[t * 32 + c] = s[c];
    }
    for (int w = 0; w < 8; ++w) emul_trip_fwd(R.data() + w * 1024, S.data() + w * 1024);
    for (int t = 0; t < 256; ++t) {                                   // second layer of C, gain, first layer of C'
        float s[32], r[32];
        for (int c = 0; c < 32; ++c) s[c] = S[t * 32 + c];
        if (pair) x_fwd2_finish_pair(s, v); else x_fwd2_finish(s, v);
        if (spec)
            for (int j = 0; j < 16; ++j) {
                const int k = pair ? 2048 * (j >> 3) + bin_of_pair(t, j) : bin_of(t, j);
                spec[2 * k] = v[j].x; spec[2 * k + 1] = v[j].y;
            }
        if (!z_out) continue;
        if (pair) {
            x_inv1_pack_pair(v, r, gperm.data() + t * 16);
        } else {
            for (int j = 0; j < 16; ++j) { v[j].x *= gperm[t * 16 + j]; v[j].y *= gperm[t * 16 + j]; }
            x_inv1_pack(v, r);
        }
        for (int c = 0; c < 32; ++c) R[t * 32 + c] = r[c];
    }
    if (!z_out) return;
    for (int w = 0; w < 8; ++w) emul_trip_inv(R.data() + w * 1024, S.data() + w * 1024);
    for (int t = 0; t < 256; ++t) {                                   // second layer of C'
        float r[32], s[32];
        for (int c = 0; c < 32; ++c) r[c] = S[t * 32 + c];
        x_layer_c_inv(r, s);
        for (int c = 0; c < 32; ++c) R[t * 32 + c] = s[c];
    }
    for (int w = 0; w < 8; ++w) emul_trip_inv(R.data() + w * 1024, S.data() + w * 1024);
    for (int t = 0; t < 256; ++t) {                                   // B'
        float s[32];
        for (int c = 0; c < 32; ++c) s[c] = S[t * 32 + c];
        x_inv2_unpack(s, v);
        float2 p[16];
        tw_table(p, WB(t));
        dft16_inv_tw(v, p);
        st_e1b(v, t, P.data());
    }
    for (int t = 0; t < 256; ++t) {                                   // A'
        ld_e1a(v, t, P.data());
        float2 p[16];
        tw_table(p, WA(t));
        dft16_inv_tw(v, p);
        for (int j = 0; j < 16; ++j) { z_out[2 * (256 * j + t)] = v[j].x; z_out[2 * (256 * j + t) + 1] = v[j].y; }
    }
}

// z_in: 4096 complex (interleaved re,im); spec_out: 4096 complex in NATURAL bin order (forward only)
int tmt_emul_forward(const float* z_in, float* spec_out) {
    emul_frame(z_in, nullptr, spec_out, nullptr);
    return 0;
}

// Full per-frame operator: out = IFFT(gain * FFT(z_in)) with gain given as a natural-order
// half spectrum g[0..2048] (the 1/4096 is applied through the permuted gain row, as on the GPU).
int tmt_emul_filter(const float* z_in, const float* g_half, float* z_out) {
    emul_frame(z_in, g_half, nullptr, z_out);
    return 0;
}

// Pair mode: z_in = two 2048-point complex frames interleaved sample by sample (z[2m + p] = frame_p[m]); spec_out = frame 0's
// 2048 bins followed by frame 1's
int tmt_emul_forward_pair(const float* z_in, float* spec_out) {
    emul_frame(z_in, nullptr, spec_out, nullptr, true);
    return 0;
}
// g_half = [2][1025]: one natural-order half spectrum per frame of the pair (1/2048 applied through the permuted rows)
int tmt_emul_filter_pair(const float* z_in, const float* g_half, float* z_out) {
    emul_frame(z_in, g_half, nullptr, z_out, true);
    return 0;
}

// One CTA of spectrum_ratio_kernel: x_frame / y_frame = 4096 interleaved stereo sample-frames, ratio_out[2049].  The
// device code's fp64 FFT (fft4096_f64, exercised on the GPU by the edge-frame parity tests) is replaced by a plain
// radix-2 transform in double; everything around it is the shared __host__ __device__ code of spectrum.cuh.
int tmt_emul_spectrum_ratio(const float* x_frame, const float* y_frame, const float* win, int a0, int a1, float* ratio_out) {
    std::vector<float> mag(2 * kBins);
    std::vector<cplx64> Z(kNfft);
    for (int side = 0; side < 2; ++side) {
        const float2* src = reinterpret_cast<const float2*>(side == 0 ? x_frame : y_frame);
        for (int n = 0; n < kNfft; ++n) {
            int r = 0;
            for (int b = 0; b < 12; ++b) r |= ((n >> b) & 1) << (11 - b);
            Z[r] = spec_window_sample(src[n], win[n]);
        }
        for (int len = 2; len <= kNfft; len <<= 1) {
            const int half = len >> 1;
            for (int i0 = 0; i0 < kNfft; i0 += len)
                for (int p = 0; p < half; ++p) {
                    const double ang = -3.14159265358979323846 * (double)p / (double)half;
                    const double cs = std::cos(ang), sn = std::sin(ang);
                    const cplx64 u = Z[i0 + p], w = Z[i0 + p + half];
                    const cplx64 v = {w.x * cs - w.y * sn, w.x * sn + w.y * cs};
                    Z[i0 + p] = cplx64{u.x + v.x, u.y + v.y};
                    Z[i0 + p + half] = cplx64{u.x - v.x, u.y - v.y};
                }
        }
        for (int t = 0; t < 256; ++t)
            for (int i = 0; i < spec_bins_of_thread(t); ++i) mag[side * kBins + spec_bin(t, i)] = spec_mean_mag(Z.data(), spec_bin(t, i));
    }
    for (int k = 0; k < kBins; ++k) ratio_out[k] = spec_ratio(mag[kBins + k], mag[k]);
    if (a1 >= a0) {
        const float g = spec_anchor_gain(ratio_out, a0, a1);
        if (g > 0.f) for (int k = 0; k < kBins; ++k) ratio_out[k] = ratio_out[k] / g;
    }
    return 0;
}

// calib.cuh on the CPU: the same per-thread functions the calibration kernels call.
int tmt_emul_decimate(const float* x, long long n_in, const float* h, int len_h, int up, int down, long long n_pre_remove,
                      long long n_out, float* out) {
    for (long long j = 0; j < n_out; ++j)
        out[j] = decimate_sample(reinterpret_cast<const float2*>(x), n_in, h, len_h, up, down, j + n_pre_remove);
    return 0;
}

int tmt_emul_power_levels(const float* x, long long n, float* mono_out) {
    for (long long i = 0; i < n; ++i) mono_out[i] = power_mono(reinterpret_cast<const float2*>(x)[i]);
    return 0;
}

// one CTA of calib_band_kernel (plain radix-2 double FFT in place of fft4096_f64, like tmt_emul_spectrum_ratio)
int tmt_emul_band_energies(const float* frame, const float* win, int lo0, int lo1, int hi0, int hi1, float* e_lo, float* e_hi) {
    std::vector<cplx64> Z(kNfft);
    const float2* src = reinterpret_cast<const float2*>(frame);
    for (int n = 0; n < kNfft; ++n) {
        int r = 0;
        for (int b = 0; b < 12; ++b) r |= ((n >> b) & 1) << (11 - b);
        Z[r] = cplx64{(double)(power_mono(src[n]) * win[n]), 0.0};
    }
    for (int len = 2; len <= kNfft; len <<= 1) {
        const int half = len >> 1;
        for (int i0 = 0; i0 < kNfft; i0 += len)
            for (int p = 0; p < half; ++p) {
                const double ang = -3.14159265358979323846 * (double)p / (double)half;
                const double cs = std::cos(ang), sn = std::sin(ang);
                const cplx64 u = Z[i0 + p], w = Z[i0 + p + half];
                const cplx64 v = {w.x * cs - w.y * sn, w.x * sn + w.y * cs};
                Z[i0 + p] = cplx64{u.x + v.x, u.y + v.y};
                Z[i0 + p + half] = cplx64{u.x - v.x, u.y - v.y};
            }
    }
    double a = 0.0, b = 0.0;
    for (int k = lo0; k < lo1; ++k) a += (double)power_bin(Z.data(), k);
    for (int k = hi0; k < hi1; ++k) b += (double)power_bin(Z.data(), k);
    *e_lo = (float)a;
    *e_hi = (float)b;
    return 0;
}

int tmt_emul_gate_grid(const float* level, const long long* start, const unsigned char* want, int n, const float* on, const float* off,
                       const long long* delay, int n_combos, int* mismatches, int* switches, unsigned char* states) {
    for (int c = 0; c < n_combos; ++c)
        gate_grid_combo(level, start, want, n, on[c], off[c], delay[c], &mismatches[c], &switches[c], states ? states + (size_t)c * n : nullptr);
    return 0;
}

// generic.cuh on the CPU: the per-thread functions of the general-FFT-size kernels; the cooperative transform of
// gen_frame_kernel is replaced by a plain radix-2 double FFT (same butterflies, twiddles from cos / sin).
int tmt_emul_gen_meansq(const float* x, long long total, long long first_start, int n_fft, int hop, int n_frames, int use_f64, float sc,
                        int mono_file, void* out) {
    const float2* src = reinterpret_cast<const float2*>(x);
    for (int k = 0; k < n_frames; ++k) {
        if (use_f64) reinterpret_cast<double*>(out)[k] = gen_frame_meansq<double>(src, total, first_start + (long long)k * hop, n_fft, sc, mono_file != 0);
        else reinterpret_cast<float*>(out)[k] = gen_frame_meansq<float>(src, total, first_start + (long long)k * hop, n_fft, sc, mono_file != 0);
    }
    return 0;
}

static void emul_fft_pow2(std::vector<cplx64>& Z, int n) {       // in: bit-reversed order, out: natural order, forward
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len >> 1;
        for (int i0 = 0; i0 < n; i0 += len)
            for (int p = 0; p < half; ++p) {
                const double ang = -3.14159265358979323846 * (double)p / (double)half;
                const double cs = std::cos(ang), sn = std::sin(ang);
                const cplx64 u = Z[i0 + p], w = Z[i0 + p + half];
                const cplx64 v = {w.x * cs - w.y * sn, w.x * sn + w.y * cs};
                Z[i0 + p] = cplx64{u.x + v.x, u.y + v.y};
                Z[i0 + p + half] = cplx64{u.x - v.x, u.y - v.y};
            }
    }
}

int tmt_emul_gen_frames(const float* x, long long total, long long first_start, int n_fft, int hop, int n_frames, const float* win,
                        const float* gains, const unsigned short* rows, float sc, int flavour, void* frames_out) {
    const float2* src = reinterpret_cast<const float2*>(x);
    int log2n = 0;
    while ((1 << log2n) < n_fft) ++log2n;
    auto rev = [&](int v) { int r = 0; for (int b = 0; b < log2n; ++b) r |= ((v >> b) & 1) << (log2n - 1 - b); return r; };
    std::vector<cplx64> Z(n_fft), Y(n_fft);
    for (int k = 0; k < n_frames; ++k) {
        const long long pos0 = first_start + (long long)k * hop;
        for (int n = 0; n < n_fft; ++n) {
            double re, im;
            gen_input(gen_sample(src, total, pos0 + n), sc, win[n], flavour, &re, &im);
            Z[rev(n)] = cplx64{re, im};
        }
        emul_fft_pow2(Z, n_fft);
        const float* g = gains + (size_t)rows[k] * (n_fft / 2 + 1);
        for (int q = 0; q < n_fft; ++q) {
            const double gg = (double)g[q <= n_fft / 2 ? q : n_fft - q];
            Y[rev(q)] = cplx64{Z[q].x * gg, -Z[q].y * gg};
        }
        emul_fft_pow2(Y, n_fft);
        for (int n = 0; n < n_fft; ++n) {
            const double yr = Y[n].x * (1.0 / n_fft), yi = -Y[n].y * (1.0 / n_fft);
            const size_t o = (size_t)k * n_fft + n;
            if (flavour == kGenAdaptiveF64) {
                reinterpret_cast<double2*>(frames_out)[o] = make_double2(yr * (double)win[n], yi * (double)win[n]);
            } else {
                float ox, oy;
                gen_output_f32(yr, yi, win[n], flavour, &ox, &oy);
                reinterpret_cast<float2*>(frames_out)[o] = make_float2(ox, oy);
            }
        }
    }
    return 0;
}

int tmt_emul_gen_ola(const void* frames, int flavour, long long total, long long first_start, int n_fft, int hop, int n_frames,
                     const float* win, float post, void* y_out) {
    for (long long s = 0; s < total; ++s) {
        if (flavour == kGenAdaptiveF64) {
            double ox, oy;
            gen_ola_sample<double, double2>(reinterpret_cast<const double2*>(frames), win, s, first_start, n_fft, hop, n_frames, true, &ox, &oy);
            reinterpret_cast<double2*>(y_out)[s] = make_double2(ox * (double)post, oy * (double)post);
        } else {
            float ox, oy;
            gen_ola_sample<float, float2>(reinterpret_cast<const float2*>(frames), win, s, first_start, n_fft, hop, n_frames,
                                          flavour == kGenAdaptiveF32, &ox, &oy);
            reinterpret_cast<float2*>(y_out)[s] = make_float2(ox * post, oy * post);
        }
    }
    return 0;
}

}  // extern "C"
