// libtomatis_b200.so -- hand-written sm_100a kernels for the Tomatis processing path and the C ABI
// declared in include/tomatis_b200.h.  See DESIGN.md for the data layout and the roofline of each
// kernel; reference citations (file:line under /root/reference) are next to each kernel.
//
// Kernels
//   input_peak_kernel   max|x| per track                                  (HBM stream, 8 B/sf read)
//   levels_kernel<T>    per hop-block pairwise sum of mono^2, bit-exact   (HBM stream, 8 B/sf read)
//   meansq_kernel<T>    m[k] = (H[k] + H[k+1]) / n_fft
//   gate_kernel<T,A,P>  gate automaton + crossfade counter, scan by map composition (one launch; three passes for long tracks)
//   stft_kernel         gather + window + FFT + gain + IFFT + window + OLA + normalise + chunk peaks + fused chunk limiter
//   edge_kernel         fp64 recomputation of the single-frame (ill-conditioned) edge blocks
//   limiter_kernel      per-chunk in-place rescale (whole-file chunks, chunks a shard owns only partly)
//   s16/s24_to_float, float_to_s24, requantise_scale     PCM edge
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <type_traits>
#include <algorithm>
#include <new>
#include <mutex>

#include "../../include/tomatis_b200.h"
#include "fft4096.cuh"
#include "spectrum.cuh"
#include "calib.cuh"
#include "generic.cuh"
#include "host_tables.hpp"

namespace {

using namespace tmt;

constexpr long long kFlushSafe = 48000LL * 5;   // src/process_tomatis.py:420 (a sample count)
constexpr int kLevelWarps = 8;                   // hop-blocks per CTA of levels_kernel
constexpr int kMaxGateStates = 255;
constexpr int kGateSegCap = 1024;                // gate-scan segments per plan (multi-segment scan of long tracks)

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                           \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return fail(TMT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                        __FILE__, __LINE__);                                                     \
    } while (0)

// ------------------------------------------------------------------------------------------------
// device-side descriptors
struct TrackDev {
    const float2* in;        // element 0 = file position in_origin
    float2* out;             // element 0 = file position out_origin
    long long in_lo, in_hi;  // readable file positions [in_lo, in_hi)  (already clipped to [0,total))
    long long in_origin;
    long long out_lo, out_hi;
    long long out_origin;
    long long total;
    long long first_start;   // position of frame 0: -n_fft/2 (streaming) or 0 (whole-file)
    int n_frames;
    int frame_base;          // offset of frame 0 in per-frame arrays
    int hs_base;             // offset of hop-block 0 in the hop-sum array (= frame_base + track index)
    int hb_lo, hb_hi;        // hop-blocks whose sums this plan computes
    int f_lo, f_hi;          // frames whose mean square this plan computes
    int chunk_base, n_chunks;
    int edge_lo, edge_hi;    // 1: block 0 / block n_frames has a single contributing frame (ill-conditioned edge)
};
struct EdgeDev { int track, frame, half, chunk; };   // fp64 edge job: half 0 -> block `frame`, half 1 -> block frame+1
struct UnitDev { int track, b0, b1, chunk; };        // STFT work unit: output blocks [b0,b1) of a track
struct ChunkDev { int track; long long s0, s1; int n_units; int fusable; };   // limiter chunk: file positions [s0,s1); fusable: the plan
                                                                            // produces the whole chunk, so the STFT kernel may limit it itself

// ------------------------------------------------------------------------------------------------
// streaming loads/stores: audio is touched once per pass, keep it out of L1
__device__ __forceinline__ float2 ld_stream(const float2* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
// read-only table load that stays where it is written (volatile: the compiler would hoist a plain __ldg to the top of the stage and
// hold its 4 registers across all of it)
__device__ __forceinline__ float4 ld_table4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(float2* p, float2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// ------------------------------------------------------------------------------------------------
// input peak (src/process_tomatis_adaptive.py:201  input_peak = np.max(np.abs(x)))
__device__ __forceinline__ float max_abs4(float m, float4 v) {
    return fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
}
// 128-bit loads, four independent loads in flight per thread (the 64-bit, one-load-per-iteration version reached 46 % of the
// DRAM peak under ncu); an odd leading / trailing sample-frame is peeled so that the body is 16-byte aligned; one atomic per CTA.
__global__ void __launch_bounds__(256) input_peak_kernel(const TrackDev* __restrict__ tracks, float* __restrict__ peaks) {
    __shared__ float warp_max[8];
    const TrackDev tr = tracks[blockIdx.y];
    const long long n = tr.in_hi - tr.in_lo;
    const float2* src = tr.in + (tr.in_lo - tr.in_origin);
    const long long head = (n > 0 && (reinterpret_cast<unsigned long long>(src) & 8ull)) ? 1 : 0;
    const float4* s4 = reinterpret_cast<const float4*>(src + head);
    const long long n4 = (n - head) >> 1;                      // float4 = two sample-frames
    const long long stride = (long long)gridDim.x * blockDim.x;
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        const float4 v0 = ld_stream4(s4 + i), v1 = ld_stream4(s4 + i + stride), v2 = ld_stream4(s4 + i + 2 * stride),
                     v3 = ld_stream4(s4 + i + 3 * stride);
        m0 = max_abs4(m0, v0); m1 = max_abs4(m1, v1); m2 = max_abs4(m2, v2); m3 = max_abs4(m3, v3);
    }
    for (; i < n4; i += stride) m0 = max_abs4(m0, ld_stream4(s4 + i));
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (head) { const float2 x = ld_stream(src); m1 = fmaxf(m1, fmaxf(fabsf(x.x), fabsf(x.y))); }
        if ((n - head) & 1) { const float2 x = ld_stream(src + n - 1); m2 = fmaxf(m2, fmaxf(fabsf(x.x), fabsf(x.y))); }
    }
    float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < 8 ? warp_max[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 4; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(peaks + blockIdx.y), __float_as_int(m));
    }
}

// ------------------------------------------------------------------------------------------------
// K2a levels.  The reference computes, per frame (src/process_tomatis.py:370, :51):
//     mono = np.sqrt(np.mean(frame**2, axis=1));  m = np.mean(mono*mono)
// in float32 (float64 in the adaptive no-attenuation branch) with NumPy's pairwise summation:
// 128-element leaves, 8 strided accumulators per leaf combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)),
// leaves combined by a balanced binary tree.  With hop = n_fft/2 = 16 leaves the frame sum is exactly
// H[k] + H[k+1] with H the pairwise sum of one 2048-sample hop block, so each hop block is reduced once.
// One warp per hop block: lane = (leaf & 7, accumulator pair); each accumulator is summed sequentially
// by one lane in NumPy's order; every add/mul/sqrt is an explicit round-to-nearest intrinsic (no FMA
// contraction), which makes m bit-identical to NumPy's.
template <typename T> struct Arith;
template <> struct Arith<float> {
    static __device__ __forceinline__ float msq(float2 x, float sc) {
        const float l = __fmul_rn(x.x, sc), r = __fmul_rn(x.y, sc);
        const float h = __fmul_rn(__fadd_rn(__fmul_rn(l, l), __fmul_rn(r, r)), 0.5f);
        const float mono = __fsqrt_rn(h);
        return __fmul_rn(mono, mono);
    }
    // single-channel file: mono = sqrt(mean(frame**2, axis=1)) = sqrt(x*x) (src/process_tomatis_adaptive.py:74 with ch == 1)
    static __device__ __forceinline__ float msq_mono(float2 x, float sc) {
        const float l = __fmul_rn(x.x, sc);
        const float mono = __fsqrt_rn(__fmul_rn(l, l));
        return __fmul_rn(mono, mono);
    }
    // calibration front end: mono = sqrt(0.5*(l*l + r*r) + 1e-12) (src/calibrate_to_baseline_v2.py:8-15)
    static __device__ __forceinline__ float msq_power(float2 x, float sc) {
        const float mono = power_mono(make_float2(__fmul_rn(x.x, sc), __fmul_rn(x.y, sc)));
        return __fmul_rn(mono, mono);
    }
    // one channel on its own: np.mean(x_mono * x_mono) (src/analyze_stereo_state.py:16-19,112-113)
    static __device__ __forceinline__ float sq(float x, float sc) {
        const float v = __fmul_rn(x, sc);
        return __fmul_rn(v, v);
    }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct Arith<double> {
    static __device__ __forceinline__ double msq(float2 x, float sc) {
        const double l = __dmul_rn((double)x.x, (double)sc), r = __dmul_rn((double)x.y, (double)sc);
        const double h = __dmul_rn(__dadd_rn(__dmul_rn(l, l), __dmul_rn(r, r)), 0.5);
        const double mono = __dsqrt_rn(h);
        return __dmul_rn(mono, mono);
    }
    static __device__ __forceinline__ double msq_mono(float2 x, float sc) {
        const double l = __dmul_rn((double)x.x, (double)sc);
        const double mono = __dsqrt_rn(__dmul_rn(l, l));
        return __dmul_rn(mono, mono);
    }
    static __device__ __forceinline__ double msq_power(float2 x, float sc) {
        const double l = __dmul_rn((double)x.x, (double)sc), r = __dmul_rn((double)x.y, (double)sc);
        const double mono = __dsqrt_rn(__dadd_rn(__dmul_rn(0.5, __dadd_rn(__dmul_rn(l, l), __dmul_rn(r, r))), 1e-12));
        return __dmul_rn(mono, mono);
    }
    static __device__ __forceinline__ double sq(float x, float sc) {
        const double v = __dmul_rn((double)x, (double)sc);
        return __dmul_rn(v, v);
    }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

template <typename T>
__global__ void __launch_bounds__(kLevelWarps * 32)
levels_kernel(const TrackDev* __restrict__ tracks, const float* __restrict__ in_scale, T* __restrict__ hsum, int mono) {
    const TrackDev tr = tracks[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = tr.hb_lo + blockIdx.x * kLevelWarps + warp;
    if (q >= tr.hb_hi) return;                       // whole warp exits together
    const float sc = in_scale ? in_scale[blockIdx.y] : 1.0f;
    const long long start = tr.first_start + (long long)q * kHop;
    const int pr = lane & 3, leaf = lane >> 2;
    const bool inside = (start >= tr.in_lo) && (start + kHop <= tr.in_hi);
    const bool aligned = (((start - tr.in_origin) & 1LL) == 0) && ((reinterpret_cast<uintptr_t>(tr.in) & 15u) == 0);
    T tot[2];
#pragma unroll
    for (int ps = 0; ps < kHop / 1024; ++ps) {
        const long long base = start + (ps * 8 + leaf) * 128 + 2 * pr;
        float2 x0[16], x1[16];
        if (inside && aligned) {
            const float4* src = reinterpret_cast<const float4*>(tr.in + (base - tr.in_origin));
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float4 v = ld_stream4(src + 4 * i);      // 8 sf = 4 float4 per step
                x0[i] = make_float2(v.x, v.y);
                x1[i] = make_float2(v.z, v.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const long long p0 = base + 8 * i, p1 = p0 + 1;
                x0[i] = (p0 >= tr.in_lo && p0 < tr.in_hi) ? tr.in[p0 - tr.in_origin] : make_float2(0.f, 0.f);
                x1[i] = (p1 >= tr.in_lo && p1 < tr.in_hi) ? tr.in[p1 - tr.in_origin] : make_float2(0.f, 0.f);
            }
        }
        T a0, a1;
        if (!mono) {
            a0 = Arith<T>::msq(x0[0], sc); a1 = Arith<T>::msq(x1[0], sc);
#pragma unroll
            for (int i = 1; i < 16; ++i) {
                a0 = Arith<T>::add(a0, Arith<T>::msq(x0[i], sc));
                a1 = Arith<T>::add(a1, Arith<T>::msq(x1[i], sc));
            }
        } else if (mono == 1) {
            a0 = Arith<T>::msq_mono(x0[0], sc); a1 = Arith<T>::msq_mono(x1[0], sc);
#pragma unroll
            for (int i = 1; i < 16; ++i) {
                a0 = Arith<T>::add(a0, Arith<T>::msq_mono(x0[i], sc));
                a1 = Arith<T>::add(a1, Arith<T>::msq_mono(x1[i], sc));
            }
        } else if (mono == 4) {                      // power-average mono with the epsilon inside the root (calibration)
            a0 = Arith<T>::msq_power(x0[0], sc); a1 = Arith<T>::msq_power(x1[0], sc);
#pragma unroll
            for (int i = 1; i < 16; ++i) {
                a0 = Arith<T>::add(a0, Arith<T>::msq_power(x0[i], sc));
                a1 = Arith<T>::add(a1, Arith<T>::msq_power(x1[i], sc));
            }
        } else {                                     // 2: left channel alone, 3: right channel alone
            const bool right = (mono == 3);
            a0 = Arith<T>::sq(right ? x0[0].y : x0[0].x, sc); a1 = Arith<T>::sq(right ? x1[0].y : x1[0].x, sc);
#pragma unroll
            for (int i = 1; i < 16; ++i) {
                a0 = Arith<T>::add(a0, Arith<T>::sq(right ? x0[i].y : x0[i].x, sc));
                a1 = Arith<T>::add(a1, Arith<T>::sq(right ? x1[i].y : x1[i].x, sc));
            }
        }
        T s = Arith<T>::add(a0, a1);                                         // r[2p] + r[2p+1]
        s = Arith<T>::add(s, __shfl_xor_sync(0xffffffffu, s, 1));            // (r0+r1)+(r2+r3) | (r4+r5)+(r6+r7)
        s = Arith<T>::add(s, __shfl_xor_sync(0xffffffffu, s, 2));            // leaf sum
        s = Arith<T>::add(s, __shfl_xor_sync(0xffffffffu, s, 4));            // 2 leaves
        s = Arith<T>::add(s, __shfl_xor_sync(0xffffffffu, s, 8));            // 4 leaves
        s = Arith<T>::add(s, __shfl_xor_sync(0xffffffffu, s, 16));           // 8 leaves = 1024 samples
        tot[ps] = s;
    }
    if (lane == 0) hsum[tr.hs_base + q] = (kHop == 2048) ? Arith<T>::add(tot[0], tot[1]) : tot[0];
}

template <typename T>
__global__ void __launch_bounds__(256)
meansq_kernel(const TrackDev* __restrict__ tracks, const T* __restrict__ hsum, T* __restrict__ msq) {
    const TrackDev tr = tracks[blockIdx.y];
    const int k = tr.f_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= tr.f_hi) return;
    const T s = Arith<T>::add(hsum[tr.hs_base + k], hsum[tr.hs_base + k + 1]);   // pairwise split at n/2
    msq[tr.frame_base + k] = Arith<T>::div(s, (T)kNfft);                          // np.mean: sum / count
}

// ------------------------------------------------------------------------------------------------
// More than two channels (adaptive mode only: the reference loops `for c in range(ch)`, src/process_tomatis_adaptive.py:307-313).
// The file is cut into channel pairs, each pair is one track of the plan and runs through the stereo kernels (an odd last
// channel rides with a silent partner); the level that drives the one shared gate comes from all channels:
//     mono = np.sqrt(np.mean(frame**2, axis=1))      (src/process_tomatis_adaptive.py:74)
// np.mean over the channel axis adds the squares in NumPy's pairwise order -- a plain left-to-right sum below 8 channels,
// 8 strided accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential remainder from 8 to 128 --
// and divides by the channel count in the array's dtype.
template <typename T> struct ArithX;
template <> struct ArithX<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float root(float a) { return __fsqrt_rn(a); }
};
template <> struct ArithX<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double root(double a) { return __dsqrt_rn(a); }
};
template <typename T>
__device__ __forceinline__ T msq_multi(const float* __restrict__ xs, int ch, float sc) {
    auto sq = [&](int c) { const T v = ArithX<T>::mul((T)__ldg(xs + c), (T)sc); return ArithX<T>::mul(v, v); };
    T s;
    if (ch < 8) {
        s = sq(0);
        for (int c = 1; c < ch; ++c) s = Arith<T>::add(s, sq(c));
    } else {
        T r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = sq(j);
        int i = 8;
        for (; i < ch - (ch % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = Arith<T>::add(r[j], sq(i + j));
        }
        s = Arith<T>::add(Arith<T>::add(Arith<T>::add(r[0], r[1]), Arith<T>::add(r[2], r[3])),
                          Arith<T>::add(Arith<T>::add(r[4], r[5]), Arith<T>::add(r[6], r[7])));
        for (; i < ch; ++i) s = Arith<T>::add(s, sq(i));
    }
    const T mono = ArithX<T>::root(Arith<T>::div(s, (T)ch));
    return ArithX<T>::mul(mono, mono);
}

// Hop-block sums of the all-channel level, same lane layout and summation order as levels_kernel; x = the interleaved file
// [total][ch]; the geometry is track 0's (every pair track of the plan has the same), the sums go to every track's slots.
template <typename T>
__global__ void __launch_bounds__(kLevelWarps * 32)
levels_multi_kernel(const TrackDev* __restrict__ tracks, int n_tracks, const float* __restrict__ in_scale, const float* __restrict__ x,
                    int ch, T* __restrict__ hsum) {
    const TrackDev tr = tracks[0];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = tr.hb_lo + blockIdx.x * kLevelWarps + warp;
    if (q >= tr.hb_hi) return;
    const float sc = in_scale ? in_scale[0] : 1.0f;
    const long long start = tr.first_start + (long long)q * kHop;
    const int pr = lane & 3, leaf = lane >> 2;
    T tot[2];
#pragma unroll
    for (int ps = 0; ps < kHop / 1024; ++ps) {
        const long long base = start + (ps * 8 + leaf) * 128 + 2 * pr;
        T a0 = (T)0, a1 = (T)0;
        for (int i = 0; i < 16; ++i) {
            const long long p0 = base + 8 * i, p1 = p0 + 1;
            const T m0 = (p0 >= tr.in_lo && p0 < tr.in_hi) ? msq_multi<T>(x + (p0 - tr.in_origin) * ch, ch, sc) : (T)0;
            const T m1 = (p1 >= tr.in_lo && p1 < tr.in_hi) ? msq_multi<T>(x + (p1 - tr.in_origin) * ch, ch, sc) : (T)0;
            a0 = i ? Arith<T>::add(a0, m0) : m0;
            a1 = i ? Arith<T>::add(a1, m1) : m1;
        }
        T s = Arith<T>::add(a0, a1);
        s = Arith<T>::add(s, __shfl_xor_sync(0xffffffffu, s, 1));
        s = Arith<T>::add(s, __shfl_xor_sync(0xffffffffu, s, 2));
        s = Arith<T>::add(s, __shfl_xor_sync(0xffffffffu, s, 4));
        s = Arith<T>::add(s, __shfl_xor_sync(0xffffffffu, s, 8));
        s = Arith<T>::add(s, __shfl_xor_sync(0xffffffffu, s, 16));
        tot[ps] = s;
    }
    if (lane == 0) {
        const T h = (kHop == 2048) ? Arith<T>::add(tot[0], tot[1]) : tot[0];
        for (int k = 0; k < n_tracks; ++k) hsum[tracks[k].hs_base + q] = h;
    }
}

// [total][ch] interleaved -> ceil(ch/2) stereo planes [pair][total][2] (a missing partner is silence), and back
__global__ void __launch_bounds__(256) channels_split_kernel(const float* __restrict__ in, long long total, int ch, float2* __restrict__ pairs) {
    const int np = (ch + 1) / 2;
    const long long n = total * np;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long s = i / np;
        const int p = (int)(i - s * np);
        const float* src = in + s * ch + 2 * p;
        pairs[(long long)p * total + s] = make_float2(src[0], (2 * p + 1 < ch) ? src[1] : 0.f);
    }
}
__global__ void __launch_bounds__(256) channels_merge_kernel(const float2* __restrict__ pairs, long long total, int ch, float* __restrict__ out) {
    const int np = (ch + 1) / 2;
    const long long n = total * np;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long s = i / np;
        const int p = (int)(i - s * np);
        const float2 v = pairs[(long long)p * total + s];
        float* dst = out + s * ch + 2 * p;
        dst[0] = v.x;
        if (2 * p + 1 < ch) dst[1] = v.y;
    }
}

// ------------------------------------------------------------------------------------------------
// Long file over several GPUs: the per-pass exchange through PEER MEMORY (NVLink) instead of collectives.
// Each rank owns one "exchange buffer" (cudaMalloc, shared with the other ranks' processes through CUDA IPC), same layout
// everywhere:   header (epoch, CTA counter, status) | flags[64] | halo staging [2 parities][2 sides][one hop] | hop sums [2][nb]
// publish_kernel (after the rank summed its own hop blocks): stores these sums straight into EVERY rank's hop-sum array,
// its first hop into the left neighbour's "from the right" slot and its last hop into the right neighbour's "from the left"
// slot, then -- last CTA, after a system-scope fence -- raises flag[rank] = epoch in every rank's buffer (release).
// wait_unpack_kernel (before the gate scan): spins until all flags of its own buffer reached the epoch (acquire), then copies
// the assembled hop sums into the plan and the two halos into the rank's input window.
// Two parities: a rank can be at most one pass ahead of a peer (it cannot leave pass k+1's wait before the peer published
// pass k+1, i.e. finished reading pass k), so the data of pass k is never overwritten before it was read.
constexpr int kPeerMaxWorld = 64;
constexpr size_t kPeerOffFlags = 256, kPeerOffHalo = 512, kPeerHaloBytes = 2 * 2 * (size_t)kHop * sizeof(float2);
constexpr size_t kPeerOffHsum = kPeerOffHalo + kPeerHaloBytes;
__host__ __device__ inline size_t peer_nb_pad(int nb) { return ((size_t)nb + 63) & ~size_t(63); }

struct PeerPublishParams {
    unsigned char* bases[kPeerMaxWorld];   // every rank's exchange buffer as this process sees it (own one included)
    int rank, world;
    const float* hsum;                     // the plan's hop sums of this track (index = hop block)
    int hb_lo, hb_hi, nb;
    const float2* first_hop;               // own samples [0, hop) and [len - hop, len)
    const float2* last_hop;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) peer_publish_kernel(const PeerPublishParams prm) {
    __shared__ int is_last;
    unsigned* hdr = reinterpret_cast<unsigned*>(prm.bases[prm.rank]);
    const unsigned epoch = hdr[0] + 1u;                               // hdr[0] is only rewritten by the last CTA, below
    const unsigned par = epoch & 1u;
    const size_t nbp = peer_nb_pad(prm.nb);
    const int gt = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
    for (int r = 0; r < prm.world; ++r) {
        float* dst = reinterpret_cast<float*>(prm.bases[r] + kPeerOffHsum) + par * nbp;
        for (int q = prm.hb_lo + gt; q < prm.hb_hi; q += gn) dst[q] = prm.hsum[q];
    }
    if (prm.rank + 1 < prm.world) {                                   // my last hop = the right neighbour's left halo
        float2* dst = reinterpret_cast<float2*>(prm.bases[prm.rank + 1] + kPeerOffHalo) + (par * 2 + 0) * kHop;
        for (int i = gt; i < kHop; i += gn) dst[i] = prm.last_hop[i];
    }
    if (prm.rank > 0) {                                               // my first hop = the left neighbour's right halo
        float2* dst = reinterpret_cast<float2*>(prm.bases[prm.rank - 1] + kPeerOffHalo) + (par * 2 + 1) * kHop;
        for (int i = gt; i < kHop; i += gn) dst[i] = prm.first_hop[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(hdr + 1, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (is_last) {
        __threadfence_system();
        for (int r = threadIdx.x; r < prm.world; r += blockDim.x)
            st_release_sys(reinterpret_cast<unsigned*>(prm.bases[r] + kPeerOffFlags) + prm.rank, epoch);
        if (threadIdx.x == 0) { hdr[1] = 0u; hdr[0] = epoch; }
    }
}

__global__ void __launch_bounds__(256)
peer_wait_unpack_kernel(unsigned char* base, int world, float* hsum_dst, int nb, float2* window, long long window_len, int left, int right,
                        unsigned long long timeout_ns) {
    unsigned* hdr = reinterpret_cast<unsigned*>(base);
    const unsigned epoch = hdr[0];                                    // raised by this rank's publish_kernel earlier on the stream
    if (threadIdx.x < world) {
        const unsigned* flag = reinterpret_cast<const unsigned*>(base + kPeerOffFlags) + threadIdx.x;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
            __nanosleep(200);
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) { atomicExch(hdr + 2, 1u + threadIdx.x); break; }     // a peer never arrived: say which, do not hang
        }
    }
    __syncthreads();
    const unsigned par = epoch & 1u;
    const int gt = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
    const float* src = reinterpret_cast<const float*>(base + kPeerOffHsum) + par * peer_nb_pad(nb);
    for (int q = gt; q < nb; q += gn) hsum_dst[q] = __ldcg(src + q);
    const float2* halo = reinterpret_cast<const float2*>(base + kPeerOffHalo) + (size_t)par * 2 * kHop;
    for (int i = gt; i < left; i += gn) window[i] = __ldcg(halo + (kHop - left) + i);
    for (int i = gt; i < right; i += gn) window[window_len - right + i] = __ldcg(halo + kHop + i);
}

// ------------------------------------------------------------------------------------------------
// K2b gate.  Both automata are finite-state machines driven by two bits per frame
// (hi: value >= on, lo: value <= off), so a run of frames is a map on the state set and maps compose
// associatively.  One CTA per track; each thread owns a contiguous segment of frames, builds the
// segment's map by simulating every start state, the maps are combined by a two-level (warp, CTA) chain,
// and each thread then replays its segment from its true start state.  The crossfade counter
// k in [0, X] (alpha = k / X) follows the state by clamped +-1 steps; its segment maps are clamp
// triples (d, lo, hi) combined the same way.
//   UPDELAY (src/process_tomatis.py:373-385): states 0..R-1 = C1 after r consecutive hi frames, R = C2.
//   MINHOLD (src/process_tomatis_adaptive.py:100-119): s = st*(H+1) + c, c = min(frames_since_switch, H).
template <int AUTO>
__device__ __forceinline__ int gate_next(int s, bool hi, bool lo, int param) {
    if (AUTO == TMT_GATE_UPDELAY) {
        if (s == param) return lo ? 0 : param;
        return hi ? s + 1 : 0;
    } else {
        const int st = (s > param) ? 1 : 0;
        const int c = s - st * (param + 1);
        const int c1 = min(c + 1, param);
        if (c1 >= param) {
            if (!st && hi) return param + 1;     // -> C2, counter 0
            if (st && lo) return 0;              // -> C1, counter 0
        }
        return st * (param + 1) + c1;
    }
}
template <int AUTO> __device__ __forceinline__ bool gate_is_c2(int s, int param) {
    return AUTO == TMT_GATE_UPDELAY ? (s == param) : (s > param);
}

struct Clamp3 { int d, lo, hi; };   // k -> min(max(k + d, lo), hi)
__device__ __forceinline__ int clamp3_apply(Clamp3 m, int k) { return min(max(k + m.d, m.lo), m.hi); }
// g after f
__device__ __forceinline__ Clamp3 clamp3_then(Clamp3 f, Clamp3 g) {
    Clamp3 r;
    r.d = f.d + g.d;
    r.lo = min(max(f.lo + g.d, g.lo), g.hi);
    r.hi = min(max(f.hi + g.d, g.lo), g.hi);
    return r;
}

constexpr int GATE_FUSED = 0, GATE_MAP = 1, GATE_STATES = 2, GATE_ROWS = 3;

// Multi-segment scan for long tracks: the frames of a track are cut into `nseg` segments, one CTA each.
//   GATE_MAP     segment transition map (S entries)                 -> segmap   [track][seg][S]
//   GATE_STATES  start state = chain of the earlier segments' maps; replay -> state[], C2 count,
//                segment clamp map of the crossfade counter          -> segclamp [track][seg][3]
//   GATE_ROWS    start counter = chain of the earlier clamp maps; replay -> rows[]
// GATE_FUSED (nseg == 1) does everything in one launch.
struct GateSeg {
    int nseg;
    uint8_t* segmap;
    int* segclamp;
};

template <typename T, int AUTO, int PASS>
__global__ void gate_kernel(const TrackDev* __restrict__ tracks, const T* __restrict__ vals,
                            const double* __restrict__ von, const double* __restrict__ voff, int param, int S,
                            int X, int alpha_init, int count_only, uint8_t* __restrict__ state,
                            uint16_t* __restrict__ rows, int* __restrict__ c2_count, GateSeg seg) {
    extern __shared__ __align__(16) unsigned char gsm[];
    const int NT = blockDim.x, NW = NT >> 5;
    const int i = threadIdx.x, w = i >> 5, lane = i & 31;
    uint8_t* maps = gsm;                     // [S][NT]
    uint8_t* wmap = maps + S * NT;           // [S][NW]
    uint8_t* segstart = wmap + S * NW;       // [NT]
    uint8_t* wstart = segstart + NT;         // [NW]
    int* ibase = reinterpret_cast<int*>(gsm + ((S * NT + S * NW + NT + NW + 15) & ~15));
    int* a_d = ibase;                        // [NT] segment clamp maps
    int* a_lo = a_d + NT;
    int* a_hi = a_lo + NT;
    int* aw = a_hi + NT;                     // [3][NW] warp clamp maps
    int* akstart = aw + 3 * NW;              // [NT]
    int* awstart = akstart + NT;             // [NW]
    int* red = awstart + NW;                 // [NW]
    int* cta_init = red + NW;                // [2] start state / start counter of this CTA
    unsigned char* prior = reinterpret_cast<unsigned char*>(cta_init + 2);   // earlier segments' maps (multi-segment passes)

    const int track = blockIdx.y, c = blockIdx.x, nseg = seg.nseg;
    const TrackDev tr = tracks[track];
    const int F = tr.n_frames;
    const T* v = vals + tr.frame_base;
    const T on = (T)von[track], off = (T)voff[track];
    const int Lc = (F + nseg - 1) / nseg;                       // frames per segment
    const int cbeg = min(F, c * Lc), cend = min(F, cbeg + Lc);
    const int L = (Lc + NT - 1) / NT;
    const int f0 = min(cend, cbeg + i * L), f1 = min(cend, f0 + L);
    const int Xe = max(X, 1);
    const int s_init = (AUTO == TMT_GATE_UPDELAY) ? 0 : param;

    if (PASS != GATE_ROWS) {
        // 1. per-thread maps
        for (int s = 0; s < S; ++s) maps[s * NT + i] = (uint8_t)s;
        for (int f = f0; f < f1; ++f) {
            const T x = v[f];
            const bool hi = x >= on, lo = x <= off;
            for (int s = 0; s < S; ++s) maps[s * NT + i] = (uint8_t)gate_next<AUTO>(maps[s * NT + i], hi, lo, param);
        }
        if (PASS == GATE_STATES) {               // earlier segments' maps -> shared memory
            const uint8_t* src = seg.segmap + (size_t)track * nseg * S;
            for (int q = i; q < c * S; q += NT) prior[q] = src[q];
        }
        __syncthreads();
        // 2. warp maps
        for (int s = lane; s < S; s += 32) {
            int cur = s;
            for (int j = 0; j < 32; ++j) cur = maps[cur * NT + w * 32 + j];
            wmap[s * NW + w] = (uint8_t)cur;
        }
        __syncthreads();
        if (PASS == GATE_MAP) {                  // segment map: chain of the warp maps for every start state
            for (int s = i; s < S; s += NT) {
                int cur = s;
                for (int ww = 0; ww < NW; ++ww) cur = wmap[cur * NW + ww];
                seg.segmap[((size_t)track * nseg + c) * S + s] = (uint8_t)cur;
            }
            return;
        }
        if (i == 0) {
            int cur = s_init;
            if (PASS == GATE_STATES)
                for (int q = 0; q < c; ++q) cur = prior[q * S + cur];
            for (int ww = 0; ww < NW; ++ww) { wstart[ww] = (uint8_t)cur; cur = wmap[cur * NW + ww]; }
        }
        __syncthreads();
        if (lane == 0) {
            int cur = wstart[w];
            for (int j = 0; j < 32; ++j) { segstart[w * 32 + j] = (uint8_t)cur; cur = maps[cur * NT + w * 32 + j]; }
        }
        __syncthreads();
    }
    // 3. replay: states, C2 count, crossfade clamp map of the thread's frames
    int c2 = 0;
    Clamp3 am = {0, 0, Xe};
    if (PASS != GATE_ROWS) {
        int cur = segstart[i];
        for (int f = f0; f < f1; ++f) {
            const T x = v[f];
            cur = gate_next<AUTO>(cur, x >= on, x <= off, param);
            const int t2 = gate_is_c2<AUTO>(cur, param) ? 1 : 0;
            c2 += t2;
            if (!count_only) {
                state[tr.frame_base + f] = (uint8_t)(1 + t2);
                Clamp3 g;
                if (alpha_init && f == 0) { g.d = 0; g.lo = g.hi = t2 * Xe; }
                else { g.d = t2 ? 1 : -1; g.lo = 0; g.hi = Xe; }
                am = clamp3_then(am, g);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) c2 += __shfl_xor_sync(0xffffffffu, c2, o);
        if (lane == 0) red[w] = c2;
    } else {
        for (int f = f0; f < f1; ++f) {
            const int t2 = state[tr.frame_base + f] == 2 ? 1 : 0;
            Clamp3 g;
            if (alpha_init && f == 0) { g.d = 0; g.lo = g.hi = t2 * Xe; }
            else { g.d = t2 ? 1 : -1; g.lo = 0; g.hi = Xe; }
            am = clamp3_then(am, g);
        }
        const int* src = seg.segclamp + (size_t)track * nseg * 3;
        int* pr = reinterpret_cast<int*>(prior);
        for (int q = i; q < c * 3; q += NT) pr[q] = src[q];
    }
    if (!count_only) { a_d[i] = am.d; a_lo[i] = am.lo; a_hi[i] = am.hi; }
    __syncthreads();
    if (PASS != GATE_ROWS && i == 0) {
        int tot = 0;
        for (int ww = 0; ww < NW; ++ww) tot += red[ww];
        if (PASS == GATE_FUSED) c2_count[track] = tot;
        else if (tot) atomicAdd(c2_count + track, tot);
    }
    if (count_only) return;
    // 4. chain the clamp maps
    if (lane == 0) {
        Clamp3 m = {0, 0, Xe};
        for (int j = 0; j < 32; ++j) {
            const int sidx = w * 32 + j;
            m = clamp3_then(m, Clamp3{a_d[sidx], a_lo[sidx], a_hi[sidx]});
        }
        aw[w] = m.d; aw[NW + w] = m.lo; aw[2 * NW + w] = m.hi;
    }
    __syncthreads();
    if (PASS == GATE_STATES) {                   // segment clamp map; rows come from the GATE_ROWS pass
        if (i == 0) {
            Clamp3 m = {0, 0, Xe};
            for (int ww = 0; ww < NW; ++ww) m = clamp3_then(m, Clamp3{aw[ww], aw[NW + ww], aw[2 * NW + ww]});
            int* dst = seg.segclamp + ((size_t)track * nseg + c) * 3;
            dst[0] = m.d; dst[1] = m.lo; dst[2] = m.hi;
        }
        return;
    }
    if (i == 0) {
        int k = 0;
        if (PASS == GATE_ROWS) {
            const int* pr = reinterpret_cast<const int*>(prior);
            for (int q = 0; q < c; ++q) k = clamp3_apply(Clamp3{pr[3 * q], pr[3 * q + 1], pr[3 * q + 2]}, k);
        }
        for (int ww = 0; ww < NW; ++ww) { awstart[ww] = k; k = clamp3_apply(Clamp3{aw[ww], aw[NW + ww], aw[2 * NW + ww]}, k); }
    }
    __syncthreads();
    if (lane == 0) {
        int k = awstart[w];
        for (int j = 0; j < 32; ++j) {
            const int sidx = w * 32 + j;
            akstart[sidx] = k;
            k = clamp3_apply(Clamp3{a_d[sidx], a_lo[sidx], a_hi[sidx]}, k);
        }
    }
    __syncthreads();
    // 5. replay the counter
    int k = akstart[i];
    for (int f = f0; f < f1; ++f) {
        const int t2 = state[tr.frame_base + f] == 2 ? 1 : 0;
        if (alpha_init && f == 0) k = t2 * Xe;
        else k = min(max(k + (t2 ? 1 : -1), 0), Xe);
        rows[tr.frame_base + f] = (uint16_t)k;
    }
}

// Fallback for automata with more states than the scan's byte-wide maps hold (up-delay beyond ~10 s, min-hold beyond ~5 s at
// 48 kHz; the reference accepts any value): one thread per track walks the frames in order.  337 500 frames take a few
// milliseconds -- a rarely used corner, kept exact rather than fast.
template <typename T, int AUTO>
__global__ void gate_serial_kernel(const TrackDev* __restrict__ tracks, const T* __restrict__ vals, const double* __restrict__ von,
                                   const double* __restrict__ voff, int param, int X, int alpha_init, int count_only,
                                   uint8_t* __restrict__ state, uint16_t* __restrict__ rows, int* __restrict__ c2_count) {
    if (threadIdx.x != 0) return;
    const int track = blockIdx.x;
    const TrackDev tr = tracks[track];
    const T* v = vals + tr.frame_base;
    const T on = (T)von[track], off = (T)voff[track];
    const int Xe = max(X, 1);
    int cur = (AUTO == TMT_GATE_UPDELAY) ? 0 : param, k = 0, c2 = 0;
    for (int f = 0; f < tr.n_frames; ++f) {
        const T x = v[f];
        cur = gate_next<AUTO>(cur, x >= on, x <= off, param);
        const int t2 = gate_is_c2<AUTO>(cur, param) ? 1 : 0;
        c2 += t2;
        if (!count_only) {
            state[tr.frame_base + f] = (uint8_t)(1 + t2);
            if (alpha_init && f == 0) k = t2 * Xe;
            else k = min(max(k + (t2 ? 1 : -1), 0), Xe);
            rows[tr.frame_base + f] = (uint16_t)k;
        }
    }
    c2_count[track] = c2;
}

// ------------------------------------------------------------------------------------------------
// Threshold search of adaptive mode (src/process_tomatis_adaptive.py:124-154) as ONE launch: one CTA per track runs the whole
// <= 30-step bisection, every step a count-only scan of the min-hold automaton over the track's levels (same map composition as
// gate_kernel, single segment).  The scalar bookkeeping is the reference's float64 arithmetic, operation for operation:
// T_mid = (T_low + T_high) / 2, thresholds T_mid +- hyst_db / 2, c2_ratio = count / n (a correctly rounded quotient like Python's
// int / int), diff = |c2_ratio - target|, best kept on strict improvement, stop below 0.01, halve towards the target.  The percentiles
// that start it (p5 / p95 / median of the valid levels) stay on the host, computed once.  Results: the final thresholds go straight
// into the plan's device-side on / off arrays for the final gate launch; best T, step count and the (T_mid, count) trace are read
// back with the other results -- no host round trip per step.
constexpr int kBisectMaxIter = 32;
struct BisectParams {
    const TrackDev* tracks;
    const double* levels;     // TMT_ARR_GATE_F64
    const double* t_low;      // [tracks] p5 of the valid levels
    const double* t_high;     // [tracks] p95
    const double* best0;      // [tracks] median (the answer when the search never improves on it, or the track is not searched)
    const int* active;        // [tracks] 0: no valid level / no frame -> best0 is the answer
    double half_hyst, target;
    int hold, S, max_iter;
    double* von;
    double* voff;
    double* best_T;           // [tracks]
    int* n_iter;              // [tracks]
    double* trace_T;          // [tracks][kBisectMaxIter]
    int* trace_c2;            // [tracks][kBisectMaxIter]
    // final gate with the thresholds found (tmt_plan_bisect_gate; emit = 0: search only)
    int emit, X, alpha_init;
    uint8_t* state;
    uint16_t* rows;
    int* c2_count;
};

// Fast path of the search for automata of at most 16 states (min-hold up to 7 frames: the reference default is 6): a transition
// map is 16 nibbles in one 64-bit register, a thread's levels stay in registers across the steps, and the scan by map composition
// runs on warp shuffles (gate_kernel's shared-memory maps cost 25 us per step on a 10-minute track, this costs ~3).
// All 16 nibbles, fully unrolled on the two 32-bit halves (constant shifts, one funnel shift per look-up): nibbles of states >= S
// carry values nobody reads (a valid state never maps to them).  The loop over S with run-time shifts was twice as long.
__device__ __forceinline__ unsigned long long nib_compose(unsigned long long a, unsigned long long b, int /*S*/) {   // s -> b[a[s]]
    const unsigned alo = (unsigned)a, ahi = (unsigned)(a >> 32);
    unsigned rlo = 0, rhi = 0;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const unsigned x = (alo >> (4 * s)) & 15u, y = (ahi >> (4 * s)) & 15u;
        rlo |= ((unsigned)(b >> (4 * x)) & 15u) << (4 * s);
        rhi |= ((unsigned)(b >> (4 * y)) & 15u) << (4 * s);
    }
    return ((unsigned long long)rhi << 32) | rlo;
}
constexpr int kBisectFastFrames = 16;        // frames per thread held in registers (1024 threads x 16 = the single-segment limit)

__global__ void __launch_bounds__(1024) bisect_fast_kernel(const BisectParams prm) {
    __shared__ unsigned long long wtot[32];
    __shared__ int red[33];
    const int i = threadIdx.x, w = i >> 5, lane = i & 31;
    const int S = prm.S, param = prm.hold;
    const int track = blockIdx.x;
    const TrackDev tr = prm.tracks[track];
    const int F = tr.n_frames;
    const double* v = prm.levels + tr.frame_base;
    const int L = (F + 1023) / 1024;                      // <= kBisectFastFrames (checked by the host)
    const int f0 = min(F, i * L), nf = min(F, f0 + L) - f0;
    double x[kBisectFastFrames];
#pragma unroll
    for (int k = 0; k < kBisectFastFrames; ++k) x[k] = (k < nf) ? v[f0 + k] : 0.0;
    // the four per-frame transition maps (class = hi | 2 * lo; both at once only when the hysteresis is zero) and the identity
    unsigned long long mc[4] = {0, 0, 0, 0}, ident = 0;
    for (int s = 0; s < S; ++s) {
#pragma unroll
        for (int c = 0; c < 4; ++c) mc[c] |= (unsigned long long)gate_next<TMT_GATE_MINHOLD>(s, (c & 1) != 0, (c & 2) != 0, param) << (4 * s);
        ident |= (unsigned long long)s << (4 * s);
    }
    auto map_of = [&](unsigned c) { return c == 0u ? mc[0] : (c == 1u ? mc[1] : (c == 2u ? mc[2] : mc[3])); };
    // The maps of four consecutive frames, one per combination of their classes (256 of them, fixed for the whole search): a
    // thread's segment map then costs one table read and one composition per four frames instead of four compositions -- the
    // compositions (S dependent nibble look-ups each) are what a search step is made of.
    __shared__ unsigned long long quad[256];
    if (i < 256)
        quad[i] = nib_compose(nib_compose(nib_compose(map_of(i & 3u), map_of((i >> 2) & 3u), S), map_of((i >> 4) & 3u), S), map_of((i >> 6) & 3u), S);
    __syncthreads();
    double T_low = prm.t_low[track], T_high = prm.t_high[track], best_T = prm.best0[track], best_diff = 1.0;
    int iters = 0;
    if (prm.active[track] && F > 0) {
        for (int it = 0; it < prm.max_iter; ++it) {
            const double T_mid = __dmul_rn(__dadd_rn(T_low, T_high), 0.5);
            const double on = __dadd_rn(T_mid, prm.half_hyst), off = __dsub_rn(T_mid, prm.half_hyst);
            // this thread's segment map
            unsigned long long seg = ident;
            unsigned cls = 0;                               // 2 bits per frame: hi | 2 * lo
#pragma unroll
            for (int k = 0; k < kBisectFastFrames; ++k) {
                if (k < nf) cls |= (((x[k] >= on) ? 1u : 0u) | ((x[k] <= off) ? 2u : 0u)) << (2 * k);
            }
#pragma unroll
            for (int k = 0; k < kBisectFastFrames; k += 4) {
                if (k + 4 <= nf) {
                    const unsigned long long q = quad[(cls >> (2 * k)) & 255u];
                    seg = k ? nib_compose(seg, q, S) : q;
                } else {
#pragma unroll
                    for (int j = k; j < k + 4; ++j)
                        if (j < nf) seg = nib_compose(seg, map_of((cls >> (2 * j)) & 3u), S);
                }
            }
            // inclusive scan of the maps over the warp, then over the warps
            unsigned long long inc = seg;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long prev = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc = nib_compose(prev, inc, S);
            }
            if (lane == 31) wtot[w] = inc;
            __syncthreads();
            if (w == 0) {
                unsigned long long t = wtot[lane];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned long long prev = __shfl_up_sync(0xffffffffu, t, o);
                    if (lane >= o) t = nib_compose(prev, t, S);
                }
                wtot[lane] = t;                              // inclusive over warps
            }
            __syncthreads();
            // start state of this thread: frames_since_switch starts at min_hold_frames in C1 (:100-103) = state `param`
            unsigned long long excl = __shfl_up_sync(0xffffffffu, inc, 1);
            if (lane == 0) excl = ident;
            int cur = param;
            if (w > 0) cur = (int)((wtot[w - 1] >> (4 * cur)) & 15ull);
            cur = (int)((excl >> (4 * cur)) & 15ull);
            int c2 = 0;
#pragma unroll
            for (int k = 0; k < kBisectFastFrames; ++k) {
                if (k < nf) {
                    const unsigned long long m = map_of((cls >> (2 * k)) & 3u);
                    cur = (int)((m >> (4 * cur)) & 15ull);
                    c2 += (cur > param) ? 1 : 0;
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) c2 += __shfl_xor_sync(0xffffffffu, c2, o);
            if (lane == 0) red[w] = c2;
            __syncthreads();
            if (w == 0) {
                int tsum = red[lane];
#pragma unroll
                for (int o = 16; o; o >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
                if (lane == 0) red[32] = tsum;
            }
            __syncthreads();
            const int total = red[32];
            const double ratio = __ddiv_rn((double)total, (double)F);
            if (i == 0) { prm.trace_T[track * kBisectMaxIter + it] = T_mid; prm.trace_c2[track * kBisectMaxIter + it] = total; }
            const double diff = fabs(__dsub_rn(ratio, prm.target));
            if (diff < best_diff) { best_diff = diff; best_T = T_mid; }
            ++iters;
            if (diff < 0.01) break;
            if (ratio < prm.target) T_high = T_mid; else T_low = T_mid;
        }
    }
    if (i == 0) {
        prm.best_T[track] = best_T;
        prm.n_iter[track] = iters;
        prm.von[track] = __dadd_rn(best_T, prm.half_hyst);
        prm.voff[track] = __dsub_rn(best_T, prm.half_hyst);
    }
    if (!prm.emit || F == 0) return;
    // Final gate with the thresholds found (simulate_gate + the alpha follower, src/process_tomatis_adaptive.py:226,253-265): one
    // more scan of the maps for the states, then a scan of the crossfade counter's clamp maps (gate_kernel's steps 3-5 on shuffles).
    {
        const double on = __dadd_rn(best_T, prm.half_hyst), off = __dsub_rn(best_T, prm.half_hyst);
        unsigned long long seg = ident;
        unsigned cls = 0;
#pragma unroll
        for (int k = 0; k < kBisectFastFrames; ++k) {
            if (k < nf) cls |= (((x[k] >= on) ? 1u : 0u) | ((x[k] <= off) ? 2u : 0u)) << (2 * k);
        }
#pragma unroll
        for (int k = 0; k < kBisectFastFrames; k += 4) {
            if (k + 4 <= nf) {
                const unsigned long long q = quad[(cls >> (2 * k)) & 255u];
                seg = k ? nib_compose(seg, q, S) : q;
            } else {
#pragma unroll
                for (int j = k; j < k + 4; ++j)
                    if (j < nf) seg = nib_compose(seg, map_of((cls >> (2 * j)) & 3u), S);
            }
        }
        unsigned long long inc = seg;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long prev = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc = nib_compose(prev, inc, S);
        }
        __syncthreads();                                   // the search's last readers of wtot / red are done
        if (lane == 31) wtot[w] = inc;
        __syncthreads();
        if (w == 0) {
            unsigned long long t = wtot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long prev = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t = nib_compose(prev, t, S);
            }
            wtot[lane] = t;
        }
        __syncthreads();
        unsigned long long excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = ident;
        int cur = param;
        if (w > 0) cur = (int)((wtot[w - 1] >> (4 * cur)) & 15ull);
        cur = (int)((excl >> (4 * cur)) & 15ull);
        // states of this thread's frames, C2 count, clamp map of the crossfade counter over them
        const int Xe = max(prm.X, 1);
        unsigned c2bits = 0;
        int c2 = 0;
        Clamp3 am = {0, 0, Xe};
        uint8_t* st = prm.state + tr.frame_base + f0;
#pragma unroll
        for (int k = 0; k < kBisectFastFrames; ++k) {
            if (k < nf) {
                const unsigned long long m = map_of((cls >> (2 * k)) & 3u);
                cur = (int)((m >> (4 * cur)) & 15ull);
                const int t2 = (cur > param) ? 1 : 0;
                c2 += t2;
                c2bits |= (unsigned)t2 << k;
                st[k] = (uint8_t)(1 + t2);
                Clamp3 g;
                if (prm.alpha_init && f0 + k == 0) { g.d = 0; g.lo = g.hi = t2 * Xe; }
                else { g.d = t2 ? 1 : -1; g.lo = 0; g.hi = Xe; }
                am = clamp3_then(am, g);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) c2 += __shfl_xor_sync(0xffffffffu, c2, o);
        __shared__ int cw[3 * 32];
        // inclusive scan of the clamp maps over the warp, then over the warps
        Clamp3 ci = am;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            Clamp3 pv;
            pv.d = __shfl_up_sync(0xffffffffu, ci.d, o); pv.lo = __shfl_up_sync(0xffffffffu, ci.lo, o); pv.hi = __shfl_up_sync(0xffffffffu, ci.hi, o);
            if (lane >= o) ci = clamp3_then(pv, ci);
        }
        __syncthreads();
        if (lane == 0) red[w] = c2;
        if (lane == 31) { cw[w] = ci.d; cw[32 + w] = ci.lo; cw[64 + w] = ci.hi; }
        __syncthreads();
        if (w == 0) {
            Clamp3 t = {cw[lane], cw[32 + lane], cw[64 + lane]};
            int tsum = red[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                Clamp3 pv;
                pv.d = __shfl_up_sync(0xffffffffu, t.d, o); pv.lo = __shfl_up_sync(0xffffffffu, t.lo, o); pv.hi = __shfl_up_sync(0xffffffffu, t.hi, o);
                if (lane >= o) t = clamp3_then(pv, t);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
            cw[lane] = t.d; cw[32 + lane] = t.lo; cw[64 + lane] = t.hi;
            if (lane == 0) prm.c2_count[track] = tsum;
        }
        __syncthreads();
        Clamp3 ce;
        ce.d = __shfl_up_sync(0xffffffffu, ci.d, 1); ce.lo = __shfl_up_sync(0xffffffffu, ci.lo, 1); ce.hi = __shfl_up_sync(0xffffffffu, ci.hi, 1);
        int kc = 0;                                         // the counter starts at 0 (alpha[0] = target_alpha[0] through alpha_init)
        if (w > 0) kc = clamp3_apply(Clamp3{cw[w - 1], cw[32 + w - 1], cw[64 + w - 1]}, kc);
        if (lane > 0) kc = clamp3_apply(ce, kc);
        uint16_t* rw = prm.rows + tr.frame_base + f0;
#pragma unroll
        for (int k = 0; k < kBisectFastFrames; ++k) {
            if (k < nf) {
                const int t2 = (int)((c2bits >> k) & 1u);
                if (prm.alpha_init && f0 + k == 0) kc = t2 * Xe;
                else kc = min(max(kc + (t2 ? 1 : -1), 0), Xe);
                rw[k] = (uint16_t)kc;
            }
        }
    }
}

__global__ void bisect_kernel(const BisectParams prm) {
    extern __shared__ __align__(16) unsigned char gsm[];
    const int NT = blockDim.x, NW = NT >> 5;
    const int i = threadIdx.x, w = i >> 5, lane = i & 31;
    const int S = prm.S, param = prm.hold;
    uint8_t* maps = gsm;                     // [S][NT]
    uint8_t* wmap = maps + S * NT;           // [S][NW]
    uint8_t* segstart = wmap + S * NW;       // [NT]
    uint8_t* wstart = segstart + NT;         // [NW]
    int* red = reinterpret_cast<int*>(gsm + ((S * NT + S * NW + NT + NW + 15) & ~15));   // [NW + 1]
    const int track = blockIdx.x;
    const TrackDev tr = prm.tracks[track];
    const int F = tr.n_frames;
    const double* v = prm.levels + tr.frame_base;
    const int L = (F + NT - 1) / NT;
    const int f0 = min(F, i * L), f1 = min(F, f0 + L);
    double T_low = prm.t_low[track], T_high = prm.t_high[track], best_T = prm.best0[track], best_diff = 1.0;
    int iters = 0;
    if (prm.active[track] && F > 0) {
        for (int it = 0; it < prm.max_iter; ++it) {
            const double T_mid = __dmul_rn(__dadd_rn(T_low, T_high), 0.5);
            const double on = __dadd_rn(T_mid, prm.half_hyst), off = __dsub_rn(T_mid, prm.half_hyst);
            // count-only scan (phases 1-3 of gate_kernel, one segment)
            for (int s = 0; s < S; ++s) maps[s * NT + i] = (uint8_t)s;
            for (int f = f0; f < f1; ++f) {
                const double x = v[f];
                const bool hi = x >= on, lo = x <= off;
                for (int s = 0; s < S; ++s) maps[s * NT + i] = (uint8_t)gate_next<TMT_GATE_MINHOLD>(maps[s * NT + i], hi, lo, param);
            }
            __syncthreads();
            for (int s = lane; s < S; s += 32) {
                int cur = s;
                for (int j = 0; j < 32; ++j) cur = maps[cur * NT + w * 32 + j];
                wmap[s * NW + w] = (uint8_t)cur;
            }
            __syncthreads();
            if (i == 0) {
                int cur = param;                                   // frames_since_switch starts at min_hold_frames, state C1 (:100-103)
                for (int ww = 0; ww < NW; ++ww) { wstart[ww] = (uint8_t)cur; cur = wmap[cur * NW + ww]; }
            }
            __syncthreads();
            if (lane == 0) {
                int cur = wstart[w];
                for (int j = 0; j < 32; ++j) { segstart[w * 32 + j] = (uint8_t)cur; cur = maps[cur * NT + w * 32 + j]; }
            }
            __syncthreads();
            int c2 = 0, cur = segstart[i];
            for (int f = f0; f < f1; ++f) {
                const double x = v[f];
                cur = gate_next<TMT_GATE_MINHOLD>(cur, x >= on, x <= off, param);
                c2 += gate_is_c2<TMT_GATE_MINHOLD>(cur, param) ? 1 : 0;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) c2 += __shfl_xor_sync(0xffffffffu, c2, o);
            if (lane == 0) red[w] = c2;
            __syncthreads();
            if (i == 0) {
                int tot = 0;
                for (int ww = 0; ww < NW; ++ww) tot += red[ww];
                red[NW] = tot;
            }
            __syncthreads();
            const int total = red[NW];
            const double ratio = __ddiv_rn((double)total, (double)F);
            if (i == 0) { prm.trace_T[track * kBisectMaxIter + it] = T_mid; prm.trace_c2[track * kBisectMaxIter + it] = total; }
            const double diff = fabs(__dsub_rn(ratio, prm.target));
            if (diff < best_diff) { best_diff = diff; best_T = T_mid; }
            ++iters;
            if (diff < 0.01) break;
            if (ratio < prm.target) T_high = T_mid; else T_low = T_mid;
        }
    }
    if (i == 0) {
        prm.best_T[track] = best_T;
        prm.n_iter[track] = iters;
        prm.von[track] = __dadd_rn(best_T, prm.half_hyst);
        prm.voff[track] = __dsub_rn(best_T, prm.half_hyst);
    }
}

// ------------------------------------------------------------------------------------------------
// K1+K3+K4 fused STFT kernel.  See fft4096.cuh for the FFT decomposition.
// Output block b of a track = positions [first_start + b*hop, +hop) = second half of frame b-1 plus
// first half of frame b.  A CTA owns a run of blocks [b0,b1) (one limiter chunk or a slice of one),
// processes frames b0-1 .. b1-1 in order and carries the running half frame from one frame to the next,
// so every output sample is written exactly once and there is no OLA buffer, no sum-of-w^2 buffer and no
// atomics on the audio path (reference: out_buf/w_buf accumulation src/process_tomatis.py:400-406,
// flush :419-426).  Thread t holds samples n = 256*j + t of the frame (j = 0..15): global loads/stores
// are coalesced 256-byte rows per warp.
//
// Register budget.  The three radix-16 stages want ~190 registers when the per-thread constants
// (analysis window 16, synthesis window x normalisation 16, overlap-add carry 16) also live in registers;
// at 128 (two CTAs per SM) ptxas spilled 43 of them to local memory, and those spills were 48 % of the
// L1 data-pipe wavefronts of v1 (profiles/r01).  Everything thread-private is therefore parked in TENSOR
// MEMORY: each warp owns a 32-lane x 128-column region of the CTA's allocation and moves 16 values per thread
// with one tcgen05.st / tcgen05.ld (.32x32b.x16).  TMEM is otherwise idle in this kernel (no MMA), its
// datapath is separate from the L1/shared-memory pipe, and nothing stays resident in registers across the
// butterflies.  The same region also carries the E2 exchange of the FFT (fft4096.cuh): the 16 x 16 transposes
// between stages B and C run as tcgen05.st/.ld round trips in two different shapes instead of through shared
// memory, which halves the shared-memory wavefronts of a frame (the pipe that limited v7, profiles/r01).
struct StftParams {
    const TrackDev* tracks;
    const UnitDev* units;
    int n_units;
    const uint16_t* rows;
    const float* gperm;     // [n_rows][4096] register-order gains, 1/4096 folded in
    const float* win;       // [4096] analysis window
    const float* swin;      // [4096] synthesis window x interior normalisation: w[n] / (w2[n%hop] + w2[n%hop + hop] (+eps | clamped))
    const float2* tw_bases; // [256][4]: per-thread twiddle bases (host_tables.hpp)
    const float2* tw_a;     // [256][16]: stage-A twiddles (W4096^t)^k, k = 0..15, exact (parked in tensor memory)
    float* chunk_peaks;
    float post_gain;
    const ChunkDev* chunks; // fused limiter (limit > 0): the CTA that finishes a chunk's last work unit rescales the chunk
    int* chunk_done;
    int* unit_counter;      // dynamic work distribution (zeroed before every launch)
    float limit;
    unsigned long long* dbg; // [2] dev counters: nanoseconds spent in chunk rescales (summed over CTAs), number of rescales
};

// tensor-memory columns of a warp: synthesis window 16 | carry 16 | raw input halves 2 x 16 | stage-A twiddles 32 | E2 exchange 32
constexpr int kPassHop = 2048;                       // sample-frames one pass of the kernel advances: one hop of 2048 or (pair mode) two of 1024
constexpr int kSampStride = kPair ? 128 : 256;       // distance between a thread's consecutive samples of its frame
// sample offset of thread t inside its frame (pair mode: lanes 2h and 2h + 1 hold sample h of the pass's first / second frame)
__device__ __forceinline__ int samp_ofs(int t) { return kPair ? (t >> 1) : t; }
constexpr int kTmemWarpCols = 128;
constexpr int kTcSwin = 0, kTcCarry = 16, kTcHalf = 32, kTcTwA = 64, kTcXchg = 96;
constexpr int kTmemCols = 256;                       // per CTA: 2 warps per lane quarter x 128 columns (2 CTAs = all 512)
// shared memory: two E1 exchange buffers | analysis window, thread-private float4 quads [4][256] | tail (TMEM slot, reduction,
// two mbarriers, queue)
constexpr int kStftSmemE1 = 2 * kE1Float2 * (int)sizeof(float2);     // two frames in flight (frame pipeline of stft_kernel)
constexpr int kStftSmemWin = kPassLen * (int)sizeof(float);
constexpr int kStftSmemWb = 0;
constexpr int kStftSmem = kStftSmemE1 + kStftSmemWin + kStftSmemWb + 96;

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

#define TMT_OPS16_IN(r, o) "f"(r[o+0]), "f"(r[o+1]), "f"(r[o+2]), "f"(r[o+3]), "f"(r[o+4]), "f"(r[o+5]), "f"(r[o+6]), "f"(r[o+7]), \
                           "f"(r[o+8]), "f"(r[o+9]), "f"(r[o+10]), "f"(r[o+11]), "f"(r[o+12]), "f"(r[o+13]), "f"(r[o+14]), "f"(r[o+15])
#define TMT_OPS16_OUT(r, o) "=f"(r[o+0]), "=f"(r[o+1]), "=f"(r[o+2]), "=f"(r[o+3]), "=f"(r[o+4]), "=f"(r[o+5]), "=f"(r[o+6]), "=f"(r[o+7]), \
                            "=f"(r[o+8]), "=f"(r[o+9]), "=f"(r[o+10]), "=f"(r[o+11]), "=f"(r[o+12]), "=f"(r[o+13]), "=f"(r[o+14]), "=f"(r[o+15])
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), TMT_OPS16_IN(r, 0) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : TMT_OPS16_OUT(r, 0) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// E2 round trips (fft4096.cuh).  forward: thread-major store (.32x32b), row-major load (.16x256b, lanes 0-15 then 16-31 of the
// warp's quarter); inverse: the same two shapes the other way round.  In place on the thread's 32-register image.
__device__ __forceinline__ void tmem_trip_fwd(uint32_t a, float (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(a), TMT_OPS16_IN(r, 0) : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(a + 16), TMT_OPS16_IN(r, 16) : "memory");
    tmem_wait_st();
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : TMT_OPS16_OUT(r, 0) : "r"(a) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : TMT_OPS16_OUT(r, 16) : "r"(a + (16u << 16)) : "memory");
    tmem_wait_ld();
}
__device__ __forceinline__ void tmem_trip_inv(uint32_t a, float (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(a), TMT_OPS16_IN(r, 0) : "memory");
    asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(a + (16u << 16)), TMT_OPS16_IN(r, 16) : "memory");
    tmem_wait_st();
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : TMT_OPS16_OUT(r, 0) : "r"(a) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : TMT_OPS16_OUT(r, 16) : "r"(a + 16) : "memory");
    tmem_wait_ld();
}

// Thread-private parking space in tensor memory (+ the analysis window in shared memory).
struct Park {
    uint32_t base;       // TMEM address of this warp's 128 columns (lane quarter in bits 31:16)
    const float4* aw;    // this thread's analysis-window quads in shared memory: aw[256 * g] = w[256 * (4g + 0..3) + t]
    __device__ __forceinline__ void init(unsigned char* smem_win, unsigned char* smem_tail, int t, const float* win, const float* swin,
                                         const float2* tw_a, float post_gain) {
        uint32_t* slot = reinterpret_cast<uint32_t*>(smem_tail);
        if ((t >> 5) == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"((uint32_t)__cvta_generic_to_shared(slot)), "n"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // the shuffle tells the compiler that the warp index is warp-uniform: the tensor-memory addresses then live in uniform
        // registers for the whole kernel instead of being converted (R2UR) in front of every access (-21 instructions per frame,
        // -0.6 .. -1.3 % kernel time, profiles/r02/ab_uniform_base.txt)
        const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
        base = *slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * kTmemWarpCols);
        float4* awr = reinterpret_cast<float4*>(smem_win) + t;
        const int so = samp_ofs(t);
#pragma unroll
        for (int g = 0; g < 4; ++g)
            awr[256 * g] = make_float4(__ldg(win + kSampStride * (4 * g) + so), __ldg(win + kSampStride * (4 * g + 1) + so),
                                       __ldg(win + kSampStride * (4 * g + 2) + so), __ldg(win + kSampStride * (4 * g + 3) + so));
        aw = awr;
        fill_tables(t, swin, tw_a, post_gain);
    }
    // thread-private constants -> this region's columns: synthesis window x normalisation x output gain | stage-A twiddles
    __device__ __forceinline__ void fill_tables(int t, const float* swin, const float2* tw_a, float post_gain) const {
        float r[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = __ldg(swin + kSampStride * j + samp_ofs(t)) * post_gain;   // output gain folded into the synthesis window
        tmem_st16(base + kTcSwin, r);
        const float4* ta = reinterpret_cast<const float4*>(tw_a + 16 * t);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 x = __ldg(ta + 4 * h + q);
                r[4 * q] = x.x; r[4 * q + 1] = x.y; r[4 * q + 2] = x.z; r[4 * q + 3] = x.w;
            }
            tmem_st16(base + kTcTwA + 16 * h, r);
        }
        tmem_wait_st();
    }
    // forward stage-A twiddles (powers of W4096^t read from tensor memory)
    __device__ __forceinline__ void twiddle_a_fwd(float2 (&v)[16], const float (&a)[16], const float (&b)[16]) const {
#pragma unroll
        for (int k = 1; k < 8; ++k) v[k] = cmul(v[k], make_float2(a[2 * k], a[2 * k + 1]));
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k + 8] = cmul(v[k + 8], make_float2(b[2 * k], b[2 * k + 1]));
    }
    // inverse stage-A twiddles are issued before the hand-off wait that precedes stage A' (the loads fly while the warp waits);
    // the operands of the frame's tail (synthesis window, carry) are issued after A's first butterfly layer, when the twiddle
    // registers are free again, and complete under the second layer
    __device__ __forceinline__ void inv_fetch_issue(float (&a)[16], float (&b)[16]) const {
        tmem_ld16(base + kTcTwA, a);
        tmem_ld16(base + kTcTwA + 16, b);
    }
    __device__ __forceinline__ void inv_fetch_apply(float2 (&v)[16], const float (&a)[16], const float (&b)[16], float (&s)[16], float2 (&c)[8]) const {
        tmem_wait_ld();
        {
            float2 p[16];
#pragma unroll
            for (int k = 0; k < 8; ++k) { p[k] = make_float2(a[2 * k], a[2 * k + 1]); p[k + 8] = make_float2(b[2 * k], b[2 * k + 1]); }
            radix4_tw<true, true, false>(v[0], v[4], v[8], v[12], p[0], p[4], p[8], p[12]);      // stage A', first layer with its twiddles fused in
            radix4_tw<true, true, true>(v[1], v[5], v[9], v[13], p[1], p[5], p[9], p[13]);
            radix4_tw<true, true, true>(v[2], v[6], v[10], v[14], p[2], p[6], p[10], p[14]);
            radix4_tw<true, true, true>(v[3], v[7], v[11], v[15], p[3], p[7], p[11], p[15]);
        }
        float cr[16];
        tmem_ld16(base + kTcSwin, s);
        tmem_ld16(base + kTcCarry, cr);
        dft16_layer2<true>(v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = make_float2(cr[2 * j], cr[2 * j + 1]);
    }
    __device__ __forceinline__ void fini(unsigned char* smem_tail, int t) {
        tmem_wait_st();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if ((t >> 5) == 0) {
            const uint32_t a = *reinterpret_cast<uint32_t*>(smem_tail);
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(a), "n"(kTmemCols) : "memory");
        }
    }
    __device__ __forceinline__ void load_tail(float (&s)[16], float2 (&c)[8]) const {
        float cr[16];
        tmem_ld16(base + kTcSwin, s);
        tmem_ld16(base + kTcCarry, cr);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = make_float2(cr[2 * j], cr[2 * j + 1]);
    }
    __device__ __forceinline__ void store_carry(const float2 (&c)[8]) const {
        float cr[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) { cr[2 * j] = c[j].x; cr[2 * j + 1] = c[j].y; }
        tmem_st16(base + kTcCarry, cr);
    }
    // raw (unwindowed) input half frames, two slots
    __device__ __forceinline__ void stage_put(int slot, const float2 (&x)[8]) const {
        float r[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) { r[2 * j] = x[j].x; r[2 * j + 1] = x[j].y; }
        tmem_st16(base + kTcHalf + 16 * slot, r);
    }
    // v = stage A's 16-point DFTs of [half `slot`, half `slot ^ 1`] x analysis window
    __device__ __forceinline__ void stage_get_windowed(int slot, float2 (&v)[16], float (&ta)[16], float (&tb2)[16]) const {
        float a[16], b[16];
        tmem_ld16(base + kTcHalf + 16 * slot, a);
        tmem_ld16(base + kTcHalf + 16 * (slot ^ 1), b);
        tmem_ld16(base + kTcTwA, ta);          // stage-A twiddles ride along: one wait for the whole frame prologue
        tmem_ld16(base + kTcTwA + 16, tb2);
        const float4 w0 = aw[0], w1 = aw[256], w2 = aw[512], w3 = aw[768];
        tmem_wait_ld();
        const float w[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
        // the window taps are fused into stage A's first butterflies (8 packed instructions fewer per frame)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v[j] = make_float2(a[2 * j], a[2 * j + 1]);
            v[j + 8] = make_float2(b[2 * j], b[2 * j + 1]);
        }
        dft16_scaled<false>(v, w);
    }
    __device__ __forceinline__ void sync_stores() const { tmem_wait_st(); }
    __device__ __forceinline__ void trip_fwd(float (&r)[32]) const { tmem_trip_fwd(base + kTcXchg, r); }
    __device__ __forceinline__ void trip_inv(float (&r)[32]) const { tmem_trip_inv(base + kTcXchg, r); }
};

// End of a work unit: publish the unit's peak; with the fused limiter, the CTA that completes a chunk's last unit rescales it.
// Called by all threads of the CTA (contains CTA barriers).  red: 9 floats of shared memory.
__device__ __forceinline__ void unit_epilogue(const StftParams& prm, int chunk, const TrackDev* trp, float peak, int t, float* red) {
    // per-chunk peak: one atomic per work unit (values are >= 0, so int ordering == float ordering)
#pragma unroll
    for (int o = 16; o; o >>= 1) peak = fmaxf(peak, __shfl_xor_sync(0xffffffffu, peak, o));
    if ((t & 31) == 0) red[t >> 5] = peak;
    __syncthreads();
    if (t == 0) {
        float m = red[0];
#pragma unroll
        for (int q = 1; q < 8; ++q) m = fmaxf(m, red[q]);
        if (m > 0.f) atomicMax(reinterpret_cast<int*>(prm.chunk_peaks + chunk), __float_as_int(m));
    }
    if (prm.limit > 0.f) {
        // Fused limiter (src/process_tomatis.py:352-355).  Every unit publishes its samples and its peak, then bumps
        // the chunk's counter; whoever bumps it last owns the finished chunk and rescales it if its peak exceeds the
        // limit.  Nobody waits for anybody, and the pass runs under the butterflies of the other resident CTAs
        // (the kernel is FP32-bound, HBM is 3/4 idle).
        const ChunkDev ch = prm.chunks[chunk];
        if (ch.fusable) {
            __threadfence();
            __syncthreads();
            if (t == 0) {
                const int prev = atomicAdd(prm.chunk_done + chunk, 1);
                red[8] = (prev == ch.n_units - 1) ? 1.f : 0.f;
            }
            __syncthreads();
            if (red[8] != 0.f) {
                __threadfence();
                const float pk = __int_as_float(atomicMax(reinterpret_cast<int*>(prm.chunk_peaks + chunk), 0));


                if (pk > prm.limit) {
                    unsigned long long t_begin = 0;
                    if (t == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
                    const float sc = __fdiv_rn(prm.limit, pk);
                    const long long s0 = max(ch.s0, trp->out_lo), s1 = min(ch.s1, trp->out_hi);
                    float2* base = trp->out + (s0 - trp->out_origin);
                    const long long n = s1 - s0;
                    long long q = t;
                    if ((reinterpret_cast<uintptr_t>(base) & 15u) == 0) {
                        // 64 KB in flight per CTA (16 x 16 B per thread).  What this pass costs is its traffic, not the time the
                        // CTA is held up (profiles/r02/limiter_study.md): 8 / 16 / 24 loads per thread in flight, the chunk
                        // prefetched into L2 (per line or through the bulk-copy engine), streaming cache policies and a loop twice
                        // as slow all leave the kernel time unchanged, while skipping the data movement alone makes the kernel
                        // 13 % faster -- the extra 1.4 TB/s of mixed read / write traffic lengthens everybody's memory latency.
                        float4* b4 = reinterpret_cast<float4*>(base);
                        const long long n4 = n >> 1;
                        long long r = t;
                        for (; r + 15 * kThreads < n4; r += 16 * kThreads) {
                            float4 x[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
                                             : "=f"(x[j].x), "=f"(x[j].y), "=f"(x[j].z), "=f"(x[j].w) : "l"(b4 + r + j * kThreads));
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(b4 + r + j * kThreads),
                                             "f"(x[j].x * sc), "f"(x[j].y * sc), "f"(x[j].z * sc), "f"(x[j].w * sc) : "memory");
                        }
                        for (; r + 3 * kThreads < n4; r += 4 * kThreads) {       // the rest of the chunk, four loads in flight
                            float4 x[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
                                             : "=f"(x[j].x), "=f"(x[j].y), "=f"(x[j].z), "=f"(x[j].w) : "l"(b4 + r + j * kThreads));
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(b4 + r + j * kThreads),
                                             "f"(x[j].x * sc), "f"(x[j].y * sc), "f"(x[j].z * sc), "f"(x[j].w * sc) : "memory");
                        }
                        for (; r < n4; r += kThreads) {
                            float4 x;
                            asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "l"(b4 + r));
                            asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(b4 + r),
                                         "f"(x.x * sc), "f"(x.y * sc), "f"(x.z * sc), "f"(x.w * sc) : "memory");
                        }
                        q = 2 * n4 + t;          // an odd last sample-frame is left to the scalar loop
                    }
                    for (; q < n; q += kThreads) {
                        float2 x;
                        asm volatile("ld.global.cg.v2.f32 {%0,%1}, [%2];" : "=f"(x.x), "=f"(x.y) : "l"(base + q));
                        st_stream(base + q, make_float2(x.x * sc, x.y * sc));
                    }
                    __syncthreads();
                    if (t == 0) {
                        unsigned long long t_end;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
                        atomicAdd(prm.dbg, t_end - t_begin);
                        atomicAdd(prm.dbg + 1, 1ull);
                    }
                }
            }
        }
    }
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tTMT_MBAR_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra TMT_MBAR_WAIT;\n\t}"
                 ::"r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(kThreads, 2) stft_kernel(const StftParams prm) {
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* bufP = reinterpret_cast<float2*>(smraw);         // E1 exchange (row pairs interleaved, 2 x 32 KB: frames i and i+1)
    unsigned char* tail = smraw + kStftSmemE1 + kStftSmemWin + kStftSmemWb;
    float* red = reinterpret_cast<float*>(tail + 16);
    int t;                                                    // read once and kept: the compiler otherwise re-reads %tid.x (S2R, a
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));           // long-latency instruction) in front of several address computations per frame

    Park park;
    park.init(smraw + kStftSmemE1, tail, t, prm.win, prm.swin, prm.tw_a, prm.post_gain);
    // stage-B twiddle bases stay in registers: the compiler hoists the powers b^k out of the frame loop; a 16 x 16 table in shared
    // memory (16 more LDS.128 per frame, 38 register copies fewer) measured the same (profiles/r02/ab_twb_smem.txt)
    const float4 bb = __ldg(reinterpret_cast<const float4*>(prm.tw_bases) + 2 * t + 1);
    const TwBase wb = {make_float2(bb.x, bb.y), make_float2(bb.z, bb.w)};

    // Work units are claimed from a global counter: unit lengths are deliberately uneven (see tmt_plan_create), so the
    // CTAs drift out of lock step and the chunk rescales of the fused limiter do not hit HBM all at once.
    int* next_unit = reinterpret_cast<int*>(tail + 56);
    // split-phase hand-offs of the frame pipeline: bar_y = "stage A stored" (R -> Q), bar_x = "stage B' stored" (Q -> P)
    const uint32_t bar_y = (uint32_t)__cvta_generic_to_shared(tail + 64), bar_x = (uint32_t)__cvta_generic_to_shared(tail + 72);
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_y), "n"(kThreads) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_x), "n"(kThreads) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t px = 0, py = 0;                                 // phase parities (uniform)
    for (;;) {
        if (t == 0) *next_unit = atomicAdd(prm.unit_counter, 1);
        __syncthreads();
        const int u = *next_unit;
        if (u >= prm.n_units) break;
        const UnitDev un = prm.units[u];
        // per-unit scalars, all relative to the unit's first frame so they fit 32 bits
        const TrackDev* trp = prm.tracks + un.track;
        const int n_frames = trp->n_frames;
        const long long upos = trp->first_start + (long long)(un.b0 - 1) * kHop;     // position of frame b0-1
        const int so = samp_ofs(t);                                                 // this thread's sample offset inside a frame
        const int tp = kPair ? (t & 1) : 0;                                         // pair mode: which frame of the pass this lane holds
        const float2* in_u = trp->in + (upos - trp->in_origin);                     // sample 0 of frame b0-1 (uniform; the thread adds its offset)
        float2* out_u = trp->out + (upos - trp->out_origin);
        const long long span = (long long)(un.b1 - un.b0 + 8) * kHop + 2 * kPassLen;
        const int in_lo = (int)max(-span, min(span, trp->in_lo - upos)), in_hi = (int)max(-span, min(span, trp->in_hi - upos));
        const int out_lo = (int)max(-span, min(span, trp->out_lo - upos)), out_hi = (int)max(-span, min(span, trp->out_hi - upos));
        const uint16_t* rows = prm.rows + trp->frame_base;
        const bool edge_lo = trp->edge_lo, edge_hi = trp->edge_hi;
        {
            float2 z[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) z[j] = make_float2(0.f, 0.f);
            park.store_carry(z);
        }
        float peak = 0.f;
        // passes of this unit: i = 0 .. last.  4096 mode: pass i = frame f = b0 - 1 + i.  Pair mode: pass i = frames f = b0 - 1 + 2i
        // (even lanes) and f + 1 (odd lanes); frames past the unit's last one (b1 - 1) are treated as absent.
        const int last = kPair ? ((un.b1 - un.b0) >> 1) : (un.b1 - un.b0);
        const int f_end = min(n_frames, un.b1);                                     // frames of this unit end here
        auto frame_exists = [&](int f) { return (f >= 0) && (f < f_end); };

        // 8 raw samples of the unit-relative hop block q (4096 mode: half frame q), zero outside the file or when `live` is false
        auto load_blk = [&](int q, bool live, float2 (&x)[8]) {
            const int p0 = q * kHop;
            const float2* src = in_u + p0 + so;
            if (live && p0 >= in_lo && p0 + kHop <= in_hi) {
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = ld_stream(src + kSampStride * j);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int p = p0 + kSampStride * j + so;
                    x[j] = (live && p >= in_lo && p < in_hi) ? __ldg(src + kSampStride * j) : make_float2(0.f, 0.f);
                }
            }
        };
        {                                  // prologue: the two staging slots of pass 0
            float2 x[8];
            if (kPair) {                   // slot 0 = first half, slot 1 = second half of this lane's frame b0 - 1 + tp
                const bool live = frame_exists(un.b0 - 1 + tp);
                load_blk(tp, live, x);
                park.stage_put(0, x);
                load_blk(tp + 1, live, x);
                park.stage_put(1, x);
            } else {                       // halves 0 and 1 of the unit
                load_blk(0, true, x);
                park.stage_put(0, x);
                load_blk(1, true, x);
                park.stage_put(1, x);
            }
        }
        // gain-row index of the next frame(s), fetched one pass ahead (it heads a dependent chain: index -> row address -> gains)
        auto row_of = [&](int f) { return (f >= 0 && f < n_frames) ? (int)__ldg(rows + f) : 0; };
        int row_next = row_of(un.b0 - 1);
        int row_next_b = kPair ? row_of(un.b0) : 0;
        auto have_pass = [&](int i) {
            const int f = un.b0 - 1 + (kPair ? 2 * i : i);
            return kPair ? (frame_exists(f) || frame_exists(f + 1)) : ((f >= 0) && (f < n_frames));
        };

        // Frame pipeline.  Per frame: R = window + stage A -> E1 buffer (A layout); Q = stages B, C, gain, C', B' on the same buffer
        // (B layout, in place); P = stage A' + overlap-add + output.  The loop runs Q(i), R(i+1), P(i) with frames i and i+1 in
        // the two E1 buffers: the stores of Q(i) are separated from their readers in P(i) by the whole of R(i+1), and the stores of
        // R(i+1) from their readers in Q(i+1) by the whole of P(i).  Both hand-offs are split-phase mbarriers (arrive after the
        // stores, wait before the loads), so a warp that is ahead does not stop, and the warps of a CTA drift out of lock step
        // instead of hitting the FP32 pipe and the exchanges all at the same time.  Hazards: R(i+2) overwrites exactly the words
        // this thread read in P(i); Q rewrites the words it read; everything else is ordered by the two barriers.
        auto stage_r = [&](int par) {               // window + A of a frame with parity par -> buf[par]
            float2 v[16];
            float fa[16], fb[16];                   // forward stage-A twiddles
            park.sync_stores();
            park.stage_get_windowed(kPair ? 0 : par, v, fa, fb);                  // window + A
            park.twiddle_a_fwd(v, fa, fb);
            st_e1a(v, t, bufP + par * kE1Float2);
            mbar_arrive(bar_y);
        };
        if (have_pass(0)) stage_r(0);

        // One pass of the pipeline.  STEADY = std::true_type for the interior iterations of a unit, where everything the general
        // body has to ask is known: passes i and i+1 exist (pair mode: all four frames), the input of passes i+1 and i+2 lies inside
        // the input window, the output blocks lie inside the output window and are no edge blocks.  The steady body has no branches
        // besides the two hand-off waits.  PAR = the pass's parity (which E1 buffer, which staging slot) when it is known at
        // compile time -- the steady loop runs two passes per iteration -- or -1.
        auto frame_iter = [&](const int i, auto steady_tag, auto par_tag) {
            constexpr bool STEADY = decltype(steady_tag)::value;
            constexpr int PAR = decltype(par_tag)::value;
            const int par = PAR < 0 ? (i & 1) : PAR;
            const int f = un.b0 - 1 + (kPair ? 2 * i : i);     // pair mode: the even lanes' frame, the odd lanes hold f + 1
            const bool have = STEADY ? true : have_pass(i);
            const int rel = i * kPassHop;                      // pass start relative to the unit
            float2* buf = bufP + par * kE1Float2;
            // every input sample is read from global memory ahead of its first use and waits in tensor memory (4096 mode: exactly
            // once; pair mode: the middle hop block of a pass by both lanes of a pair, the second time from L2);
            // the loads below belong to pass i+1 and complete under this pass's butterflies
            float2 pf[8];
            const bool do_pf = STEADY ? true : (i < last);
            const bool live_next = kPair ? frame_exists(f + 2 + tp) : true;      // pair mode, general body: this lane's frame of pass i+1
            const int row = row_next, row_b = row_next_b;
            if (STEADY) {
                if (kPair) {
                    row_next = (int)__ldg(rows + f + 2);
                    row_next_b = (int)__ldg(rows + f + 3);
                    const float2* src = in_u + (2 * i + 2 + tp) * kHop + so;      // first half of this lane's frame of pass i+1
#pragma unroll
                    for (int j = 0; j < 8; ++j) pf[j] = ld_stream(src + kSampStride * j);
                    if (t < 32) {                              // the two new hop blocks of pass i+2 -> L2 (128 lines)
                        const float2* q = in_u + (2 * i + 5) * kHop + t * 16;
#pragma unroll
                        for (int j = 0; j < 4; ++j) prefetch_l2(q + 512 * j);
                    }
                } else {
                    row_next = (int)__ldg(rows + f + 1);
                    const float2* src = in_u + (i + 2) * kHop + t;
#pragma unroll
                    for (int j = 0; j < 8; ++j) pf[j] = ld_stream(src + 256 * j);
                    if (t < 32) {                              // half i+3 -> L2: the first warp asks for all 128 lines of it
                        const float2* q = in_u + (i + 3) * kHop + t * 16;
#pragma unroll
                        for (int j = 0; j < 4; ++j) prefetch_l2(q + 512 * j);
                    }
                }
            } else if (kPair) {
                row_next = row_of(f + 2);
                row_next_b = row_of(f + 3);
                if (do_pf) load_blk(2 * i + 2 + tp, live_next, pf);
            } else {
                row_next = row_of(f + 1);
                if (do_pf) load_blk(i + 2, true, pf);
                if (i + 1 < last && (t & 15) == 0) {           // half i+3 -> L2 (one 128-byte line per 16 lanes), so that next frame's loads are short
                    const int p0 = (i + 3) * kHop;
                    if (p0 >= in_lo && p0 + kHop <= in_hi) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) prefetch_l2(in_u + p0 + t + 256 * j);
                    }
                }
            }
            // pair mode: the second half of this lane's frame of pass i+1, loaded into the registers the first half has just left
            auto load_second = [&]() {
                if (STEADY) {
                    const float2* src = in_u + (2 * i + 3 + tp) * kHop + so;
#pragma unroll
                    for (int j = 0; j < 8; ++j) pf[j] = ld_stream(src + kSampStride * j);
                } else {
                    load_blk(2 * i + 3 + tp, live_next, pf);
                }
            };
            if (have) {                                                           // ---- Q(i)
                float2 v[16];
                mbar_wait(bar_y, py);
                py ^= 1;
                ld_e1b(v, t, buf);
                dft16<false>(v);                                                  // B
                tw_pow<false>(v, wb);
                // E2: the 16 x 16 transposes between stages B and C run through tensor memory, one round trip on either side
                // of C's first radix-4 layer (fft4096.cuh); no shared-memory traffic, no barrier
                float r[32], q[32];
                if (do_pf) park.stage_put(kPair ? 0 : par, pf);   // 4096 mode: slot of the half frame i no longer needs (its stage A is long done)
                x_fwd1_pack(v, r);
                if (kPair && do_pf) load_second();             // completes under the first round trip and C's first layer (issued after
                park.trip_fwd(r);                              // the trip or after the layer: same time, profiles/r02/fused2048_ab_load_position.txt)
                // tilt gain x crossfade weight: one real row per frame, register order; issued before the second round trip
                // (the few rows in use stay in L1; one trip earlier measured the same).  Pair mode: registers 0-7 belong to the
                // pass's first frame, 8-15 to its second, each with its own row.
                const float4* g4 = reinterpret_cast<const float4*>(prm.gperm + (size_t)row * kPassLen + t * 16);
                const float4* g4b = kPair ? reinterpret_cast<const float4*>(prm.gperm + (size_t)row_b * kPassLen + t * 16) : g4;
                const float4 g0 = ld_table4(g4), g1 = ld_table4(g4 + 1), g2 = ld_table4(g4b + 2), g3 = ld_table4(g4b + 3);
                x_layer_a<false>(r, q);                                           // C, first layer
                if (kPair && do_pf) park.stage_put(1, pf);
                park.trip_fwd(q);
                if (kPair) x_fwd2_finish_pair(q, v); else x_fwd2_finish(q, v);    // C, second layer
                {       // the gains are fused into the first butterflies of C' (8 packed instructions fewer per frame)
                    const float g[16] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w, g2.x, g2.y, g2.z, g2.w, g3.x, g3.y, g3.z, g3.w};
                    if (kPair) x_inv1_pack_pair(v, r, g); else x_inv1_pack(v, r, g);   // gain + C', first layer + inner twiddles
                }
                park.trip_inv(r);
                x_layer_c_inv(r, q);                                              // C', second layer
                park.trip_inv(q);
                x_inv2_unpack(q, v);
                {
                    float2 pw[16];
                    tw_table(pw, wb);
                    dft16_inv_tw(v, pw);                                          // B' (twiddles fused into the first butterflies)
                }
                st_e1b(v, t, buf);             // the very words this thread read in ld_e1b
                mbar_arrive(bar_x);
            }
            if (do_pf) {
                if (!have) {
                    park.stage_put(kPair ? 0 : par, pf);
                    if (kPair) { load_second(); park.stage_put(1, pf); }
                }
                if (STEADY || have_pass(i + 1)) stage_r(par ^ 1);                // ---- R(i+1)
            }
            float2 v[16];                                                         // ---- P(i)
            float s[16];                                           // synthesis window x normalisation (x output gain)
            float2 c[8];                                           // carried half frame
            park.sync_stores();
            if (have) {
                float ta[16], tb2[16];
                park.inv_fetch_issue(ta, tb2);
                mbar_wait(bar_x, px);
                px ^= 1;
                ld_e1a(v, t, buf);             // stage A of frame i+2 overwrites exactly the words this thread reads here
                park.inv_fetch_apply(v, ta, tb2, s, c);                         // A' (twiddles fused into the first butterflies)
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = make_float2(0.f, 0.f);
                park.load_tail(s, c);
            }

            // synthesis window, overlap-add with the carried half, interior normalisation (folded into swin)
            if (kPair) {
                // Block f = carried half (second half of frame f-1, held by the ODD lanes) + first half of frame f (even lanes);
                // block f+1 = second half of frame f (even lanes) + first half of frame f+1 (odd lanes).  The lanes of a pair
                // swap their raw first halves (the synthesis-window taps are the same for both): the odd lane then emits block f,
                // the even lane block f+1, and the odd lane's second half is the next carry.
                float2 b2[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    v[j].x = __shfl_xor_sync(0xffffffffu, v[j].x, 1);
                    v[j].y = __shfl_xor_sync(0xffffffffu, v[j].y, 1);
                    b2[j] = cscale(v[j + 8], s[j + 8]);
                }
                const int blk = f + 1 - tp;                                    // the block this lane emits
                const int relb = rel + (1 - tp) * kHop;
                const bool edge_blk = STEADY ? false : ((blk == 0 && edge_lo) || (blk == n_frames && edge_hi));
                if (STEADY || (blk >= un.b0 && blk < un.b1 && !edge_blk)) {
                    float2* dst = out_u + relb + so;
                    if (STEADY || ((relb >= out_lo) && (relb + kHop <= out_hi))) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float2 o = __ffma2_rn(v[j], make_float2(s[j], s[j]), tp ? c[j] : b2[j]);
                            st_stream(dst + kSampStride * j, o);
                            peak = fmaxf(peak, fmaxf(fabsf(o.x), fabsf(o.y)));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float2 o = __ffma2_rn(v[j], make_float2(s[j], s[j]), tp ? c[j] : b2[j]);
                            const int p = relb + kSampStride * j + so;
                            if (p >= out_lo && p < out_hi) {
                                st_stream(dst + kSampStride * j, o);
                                peak = fmaxf(peak, fmaxf(fabsf(o.x), fabsf(o.y)));
                            }
                        }
                    }
                }
                park.store_carry(b2);
            } else {
            const bool edge_blk = STEADY ? false : ((f == 0 && edge_lo) || (f == n_frames && edge_hi));   // single-frame blocks: edge_kernel
            if (STEADY || (f >= un.b0 && !edge_blk)) {       // emit output block f
                float2* dst = out_u + rel + t;
                if (STEADY || ((rel >= out_lo) && (rel + kHop <= out_hi))) {          // whole block inside the output window
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float2 o = __ffma2_rn(v[j], make_float2(s[j], s[j]), c[j]);
                        st_stream(dst + 256 * j, o);
                        peak = fmaxf(peak, fmaxf(fabsf(o.x), fabsf(o.y)));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float2 o = __ffma2_rn(v[j], make_float2(s[j], s[j]), c[j]);
                        const int p = rel + 256 * j + t;
                        if (p >= out_lo && p < out_hi) {
                            st_stream(dst + 256 * j, o);
                            peak = fmaxf(peak, fmaxf(fabsf(o.x), fabsf(o.y)));
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) c[j] = cscale(v[j + 8], s[j + 8]);
            park.store_carry(c);
            }
        };
        // steady interval [s_lo, s_hi) of this unit (all bounds are monotone in i; positions are relative to the unit)
        int s_lo, s_hi;
        if (kPair) {
            // passes i and i+1 hold existing frames of the unit (f + 3 < f_end), i >= 1 (block f is the unit's own and no edge block:
            // f >= b0 >= 0, and f >= 1), the loads of pass i+1 ([2i+2, 2i+5) hop blocks) and the L2 prefetch of pass i+2 (up to
            // 2i+7) inside the input window, the output blocks [2i, 2i+2) inside the output window
            s_lo = max(1, max(((((in_lo + kHop - 1) >> 10) - 2) + 1) >> 1, (((out_lo + kHop - 1) >> 10) + 1) >> 1));
            s_hi = min((f_end - un.b0 - 1) >> 1, min(((in_hi >> 10) - 5) >> 1, (out_hi >> 10) >> 1));
            if (edge_lo && un.b0 == 0) s_lo = max(s_lo, 1);      // block 0 is pass 0's (never steady); block 1 = pass 1's f: no edge
        } else {
            s_lo = max(1, 1 - un.b0 + (edge_lo ? 1 : 0)), s_hi = min(last - 1, n_frames - un.b0);
            s_lo = max(s_lo, max(((in_lo + kHop - 1) >> 11) - 2, (out_lo + kHop - 1) >> 11));
            s_hi = min(s_hi, min((in_hi >> 11) - 3, out_hi >> 11));
        }
        using Even = std::integral_constant<int, 0>;
        using Odd = std::integral_constant<int, 1>;
        using Any = std::integral_constant<int, -1>;
        for (int i = 0; i <= last;) {                 // general iterations around the steady run, which starts on an even frame
            if (i >= s_lo && i + 1 < s_hi && !(i & 1)) {
                do {
                    frame_iter(i, std::true_type{}, Even{});
                    frame_iter(i + 1, std::true_type{}, Odd{});
                    i += 2;
                } while (i + 1 < s_hi);
            } else {
                frame_iter(i, std::false_type{}, Any{});
                ++i;
            }
        }
        unit_epilogue(prm, un.chunk, trp, peak, t, red);
        __syncthreads();
    }
    park.fini(tail, t);
}

// ------------------------------------------------------------------------------------------------
// fp64 edge frames.  Two output regions of the reference are ill-conditioned: the
// first hop of adaptive mode (one frame, y*w / max(w^2,1e-8) with w -> 0) and the tail block of all modes
// (y*w / (w^2 + 1e-12) up to the window's last tap).  Rounding noise of an fp32 FFT is amplified there by
// up to 1/w ~ 1e4..4e5 -- the reference's own float32 and float64 FFT runs differ by 1e-4 in those samples --
// so the single frame that feeds such a block is recomputed here in double precision (radix-2 in shared
// memory; one CTA per edge, cost irrelevant) with the reference's float32 roundings on either side of the
// FFT reproduced explicitly.
struct EdgeParams {
    const TrackDev* tracks;
    const EdgeDev* edges;
    const uint16_t* rows;
    const float* gnat;       // [n_rows][2049] natural-order gains
    const float* win;
    const float* in_scale;   // per track or NULL
    const float* out_scale;  // per track or NULL
    float* chunk_peaks;
    int norm_clamp;
    int pipeline_f64;        // adaptive float64 branch: no float32 roundings around the FFT
    float post_gain;
};

constexpr int kEdgeSmemBytes = kNfft * (int)sizeof(double2);
constexpr int kEdgeKernelSmemBytes = kEdgeSmemBytes + 2048 * (int)sizeof(double2);     // edge_kernel: + the twiddle table

constexpr int kLog2N = kPair ? 11 : 12;
__device__ __forceinline__ int bitrev12(int x) { return (int)(__brev((unsigned)x) >> (32 - kLog2N)); }   // bit reversal over log2(n_fft) bits

// exp(-i*pi*k/2048), k < 2048: filled once per device by tw64_init_kernel with the very sincospi values the butterflies used to
// compute on the fly (bit-identical results; the trigonometry was 3/4 of edge_kernel's 87 us)
__device__ double2 g_tw64[2048];
__global__ void tw64_init_kernel() {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < 2048) {
        double sn, cs;
        sincospi(-(double)k / 2048.0, &sn, &cs);
        g_tw64[k] = make_double2(cs, sn);
    }
}

// tw: the twiddle table, g_tw64 itself or a shared-memory copy of it (edge_kernel: a table read from global memory per
// stage was a third of that latency-bound kernel)
__device__ void fft4096_f64(double2* sm, int t, const double2* tw = g_tw64) {      // in: bit-reversed order, out: natural order, forward
    for (int s = 1; s <= kLog2N; ++s) {          // the table holds exp(-2*pi*i*k/4096): stage s reads every (4096 >> s)-th entry
        const int half = 1 << (s - 1);
        for (int b = t; b < kNfft / 2; b += 256) {
            const int pos = b & (half - 1);
            const int i = ((b >> (s - 1)) << s) + pos, j = i + half;
            const double2 w = tw[pos << (12 - s)];
            const double cs = w.x, sn = w.y;
            const double2 u = sm[i], x = sm[j];
            const double2 v = make_double2(x.x * cs - x.y * sn, x.x * sn + x.y * cs);
            sm[i] = make_double2(u.x + v.x, u.y + v.y);
            sm[j] = make_double2(u.x - v.x, u.y - v.y);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) edge_kernel(const EdgeParams prm) {
    extern __shared__ __align__(16) unsigned char smraw[];
    double2* sm = reinterpret_cast<double2*>(smraw);
    __shared__ float red[8];
    const int t = threadIdx.x;
    const EdgeDev ed = prm.edges[blockIdx.x];
    const TrackDev tr = prm.tracks[ed.track];
    const long long pos0 = tr.first_start + (long long)ed.frame * kHop;
    const float sc = prm.in_scale ? prm.in_scale[ed.track] : 1.0f;
    const float osc = prm.out_scale ? prm.out_scale[ed.track] : 1.0f;
    double2* tw = sm + kNfft;                              // twiddles staged once per CTA (same values, same arithmetic)
    for (int k = t; k < 2048; k += 256) tw[k] = g_tw64[k];
    for (int n = t; n < kNfft; n += 256) {
        const long long p = pos0 + n;
        float2 x = (p >= tr.in_lo && p < tr.in_hi) ? tr.in[p - tr.in_origin] : make_float2(0.f, 0.f);
        const float w = prm.win[n];
        double2 z;
        if (prm.pipeline_f64) {
            z = make_double2((double)x.x * (double)sc * (double)w, (double)x.y * (double)sc * (double)w);
        } else {   // x_atten = fl32(x*atten); frame*win in float32 (src/process_tomatis.py:396, _adaptive.py:215,311)
            z = make_double2((double)__fmul_rn(__fmul_rn(x.x, sc), w), (double)__fmul_rn(__fmul_rn(x.y, sc), w));
        }
        sm[bitrev12(n)] = z;
    }
    __syncthreads();
    fft4096_f64(sm, t, tw);
    {   // gain (real, symmetric), conjugate for the inverse transform, then bit-reverse in place
        const float* g = prm.gnat + (size_t)prm.rows[tr.frame_base + ed.frame] * (kNfft / 2 + 1);
        for (int k = t; k < kNfft; k += 256) {
            const double gg = (double)g[k <= kNfft / 2 ? k : kNfft - k];
            const double2 v = sm[k];
            sm[k] = make_double2(v.x * gg, -v.y * gg);
        }
        __syncthreads();
        for (int k = t; k < kNfft; k += 256) {
            const int r = bitrev12(k);
            if (k < r) { const double2 a = sm[k]; sm[k] = sm[r]; sm[r] = a; }
        }
        __syncthreads();
    }
    fft4096_f64(sm, t, tw);  // conj(FFT(conj(Y))) = N * IFFT(Y)
    float peak = 0.f;
    const int blk = ed.frame + ed.half;
    const long long out0 = tr.first_start + (long long)blk * kHop;
    for (int n = t; n < kHop; n += 256) {
        const int wi = n + ed.half * kHop;                 // window tap of this output sample
        const double2 v = sm[wi];
        const double yr = v.x * (1.0 / kNfft), yi = -v.y * (1.0 / kNfft);
        const float w = prm.win[wi];
        float ox, oy;
        if (prm.pipeline_f64) {          // everything in double, one final rounding
            const double nrm = (double)__fmul_rn(w, w);    // norm is float32 in the reference (_adaptive.py:293,323)
            const double den = fmax(nrm, (double)1e-8f);
            ox = (float)(yr * (double)w / den * (double)osc);
            oy = (float)(yi * (double)w / den * (double)osc);
        } else {
            float ax, ay;
            if (prm.norm_clamp) {        // adaptive: y_frame = fl32(irfft * win)
                ax = (float)(yr * (double)w); ay = (float)(yi * (double)w);
            } else {                     // streaming: irfft(...).astype(float32) * win
                ax = __fmul_rn((float)yr, w); ay = __fmul_rn((float)yi, w);
            }
            const float nrm = __fmul_rn(w, w);
            const float den = prm.norm_clamp ? fmaxf(nrm, 1e-8f) : __fadd_rn(nrm, 1e-12f);
            ox = __fdiv_rn(ax, den); oy = __fdiv_rn(ay, den);
            if (prm.out_scale) { ox = __fmul_rn(ox, osc); oy = __fmul_rn(oy, osc); }
        }
        if (prm.post_gain != 1.0f) { ox = __fmul_rn(ox, prm.post_gain); oy = __fmul_rn(oy, prm.post_gain); }
        const long long p = out0 + n;
        if (p >= tr.out_lo && p < tr.out_hi) {
            tr.out[p - tr.out_origin] = make_float2(ox, oy);
            peak = fmaxf(peak, fmaxf(fabsf(ox), fabsf(oy)));
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) peak = fmaxf(peak, __shfl_xor_sync(0xffffffffu, peak, o));
    if ((t & 31) == 0) red[t >> 5] = peak;
    __syncthreads();
    if (t == 0) {
        float m = red[0];
        for (int q = 1; q < 8; ++q) m = fmaxf(m, red[q]);
        if (m > 0.f) atomicMax(reinterpret_cast<int*>(prm.chunk_peaks + ed.chunk), __float_as_int(m));
    }
}

// ------------------------------------------------------------------------------------------------
// limiter (src/process_tomatis.py:352-355): if peak > limit: chunk *= limit / peak
__global__ void __launch_bounds__(256)
limiter_kernel(const TrackDev* __restrict__ tracks, const ChunkDev* __restrict__ chunks,
               const float* __restrict__ peaks, float limit, int skip_fusable) {
    const float peak = peaks[blockIdx.y];
    if (!(peak > limit)) return;
    const float scale = __fdiv_rn(limit, peak);
    const ChunkDev ch = chunks[blockIdx.y];
    if (skip_fusable && ch.fusable) return;            // already limited inside stft_kernel
    const TrackDev tr = tracks[ch.track];
    const long long s0 = max(ch.s0, tr.out_lo), s1 = min(ch.s1, tr.out_hi);
    float2* dst = tr.out + (s0 - tr.out_origin);
    const long long n = s1 - s0;
    if (n <= 0) return;
    // 128-bit accesses, two per thread in flight; an odd leading / trailing sample-frame is peeled (16-byte alignment)
    const long long head = (reinterpret_cast<unsigned long long>(dst) & 8ull) ? 1 : 0;
    float4* d4 = reinterpret_cast<float4*>(dst + head);
    const long long n4 = (n - head) >> 1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll 1
    for (; i + stride < n4; i += 2 * stride) {
        float4 a = __ldcs(d4 + i), b = __ldcs(d4 + i + stride);
        a.x = __fmul_rn(a.x, scale); a.y = __fmul_rn(a.y, scale); a.z = __fmul_rn(a.z, scale); a.w = __fmul_rn(a.w, scale);
        b.x = __fmul_rn(b.x, scale); b.y = __fmul_rn(b.y, scale); b.z = __fmul_rn(b.z, scale); b.w = __fmul_rn(b.w, scale);
        __stcs(d4 + i, a);
        __stcs(d4 + i + stride, b);
    }
    for (; i < n4; i += stride) {
        float4 a = __ldcs(d4 + i);
        a.x = __fmul_rn(a.x, scale); a.y = __fmul_rn(a.y, scale); a.z = __fmul_rn(a.z, scale); a.w = __fmul_rn(a.w, scale);
        __stcs(d4 + i, a);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (head) { float2 x = dst[0]; x.x = __fmul_rn(x.x, scale); x.y = __fmul_rn(x.y, scale); dst[0] = x; }
        if ((n - head) & 1) { float2 x = dst[n - 1]; x.x = __fmul_rn(x.x, scale); x.y = __fmul_rn(x.y, scale); dst[n - 1] = x; }
    }
}

// ------------------------------------------------------------------------------------------------
// PCM edge: the reference reads integer PCM files as float32 (soundfile: value / 2^(bits-1)) and
// writes PCM_24 (src/process_tomatis.py:243, _adaptive.py:351).  Doing both conversions on the device lets the host
// move 2-3 bytes per sample over PCIe instead of 4.
__global__ void __launch_bounds__(256) s16_to_float_kernel(const int4* __restrict__ in, float4* __restrict__ out, long long n8,
                                                            const short* __restrict__ in_s, float* __restrict__ out_s, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const int4 v = in[i];
        const int w[4] = {v.x, v.y, v.z, v.w};
        float f[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            f[2 * k] = (float)(short)(w[k] & 0xffff) * (1.0f / 32768.0f);
            f[2 * k + 1] = (float)(short)(w[k] >> 16) * (1.0f / 32768.0f);
        }
        out[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
        out[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    for (long long i = 8 * n8 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out_s[i] = (float)in_s[i] * (1.0f / 32768.0f);
}

// packed little-endian 24-bit: 4 samples = 12 bytes = 3 words
__global__ void __launch_bounds__(256) s24_to_float_kernel(const unsigned* __restrict__ in, float4* __restrict__ out, long long n4,
                                                            const unsigned char* __restrict__ in_b, float* __restrict__ out_s, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const unsigned a = in[3 * i], b = in[3 * i + 1], c = in[3 * i + 2];
        const int s0 = (int)(a << 8) >> 8;
        const int s1 = (int)(((a >> 24) | (b << 8)) << 8) >> 8;
        const int s2 = (int)(((b >> 16) | (c << 16)) << 8) >> 8;
        const int s3 = (int)c >> 8;
        const float k = 1.0f / 8388608.0f;
        out[i] = make_float4((float)s0 * k, (float)s1 * k, (float)s2 * k, (float)s3 * k);
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int v = (int)((unsigned)in_b[3 * i] | ((unsigned)in_b[3 * i + 1] << 8) | ((unsigned)in_b[3 * i + 2] << 16));
        out_s[i] = (float)((v << 8) >> 8) * (1.0f / 8388608.0f);
    }
}

// PCM edge fused with the levels pass (round 2): one warp per hop block converts the block's integer samples to float32 into the
// plan's input buffer (soundfile's read: value / 2^(bits-1), exactly s16_to_float_kernel / s24_to_float_kernel) and, on the way,
// sums the block in levels_kernel's lane layout and order -- the host-buffer pipeline with integer input then has no levels pass
// over the float samples at all.  The hop blocks of a track tile [first_start, first_start + (n_frames + 1) * hop), which covers
// the file, so every sample is converted exactly once.  FMT: 0 = int16, 1 = packed little-endian 24 bit.
template <int FMT>
__device__ __forceinline__ void pcm_pair(const unsigned char* __restrict__ src, long long p0, long long total, bool word_aligned,
                                         float2& x0, float2& x1) {
    x0 = x1 = make_float2(0.f, 0.f);
    if (p0 + 1 < total && p0 >= 0 && word_aligned) {             // both sample-frames inside the file: 8 or 12 aligned bytes (p0 is even)
        if (FMT == 0) {
            const int2 w = *reinterpret_cast<const int2*>(src + 4 * p0);
            x0 = make_float2((float)(short)(w.x & 0xffff) * (1.0f / 32768.0f), (float)(short)(w.x >> 16) * (1.0f / 32768.0f));
            x1 = make_float2((float)(short)(w.y & 0xffff) * (1.0f / 32768.0f), (float)(short)(w.y >> 16) * (1.0f / 32768.0f));
        } else {
            const unsigned* q = reinterpret_cast<const unsigned*>(src + 6 * p0);
            const unsigned a = q[0], b = q[1], c = q[2];
            const float k = 1.0f / 8388608.0f;
            x0 = make_float2((float)((int)(a << 8) >> 8) * k, (float)((int)(((a >> 24) | (b << 8)) << 8) >> 8) * k);
            x1 = make_float2((float)((int)(((b >> 16) | (c << 16)) << 8) >> 8) * k, (float)((int)c >> 8) * k);
        }
        return;
    }
    auto one = [&](long long p) {
        if (p < 0 || p >= total) return make_float2(0.f, 0.f);
        if (FMT == 0) {
            const short* q = reinterpret_cast<const short*>(src + 4 * p);
            return make_float2((float)q[0] * (1.0f / 32768.0f), (float)q[1] * (1.0f / 32768.0f));
        }
        const unsigned char* q = src + 6 * p;
        const int l = (int)((unsigned)q[0] | ((unsigned)q[1] << 8) | ((unsigned)q[2] << 16));
        const int r = (int)((unsigned)q[3] | ((unsigned)q[4] << 8) | ((unsigned)q[5] << 16));
        return make_float2((float)((l << 8) >> 8) * (1.0f / 8388608.0f), (float)((r << 8) >> 8) * (1.0f / 8388608.0f));
    };
    x0 = one(p0);
    x1 = one(p0 + 1);
}

template <int FMT>
__global__ void __launch_bounds__(kLevelWarps * 32)
pcm_levels_kernel(const TrackDev* __restrict__ tracks, const unsigned char* __restrict__ pcm, long long track_stride, float* __restrict__ hsum) {
    const TrackDev tr = tracks[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * kLevelWarps + warp;
    if (tr.n_frames <= 0 || q > tr.n_frames) return;      // whole warp exits together
    const unsigned char* src = pcm + (long long)blockIdx.y * track_stride;
    const bool word_aligned = (reinterpret_cast<uintptr_t>(src) & (FMT == 0 ? 7u : 3u)) == 0;
    float2* dst = const_cast<float2*>(tr.in);             // whole-track plans: in_origin = 0
    const bool dst16 = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
    const long long start = tr.first_start + (long long)q * kHop;
    const int pr = lane & 3, leaf = lane >> 2;
    float tot[2];
#pragma unroll
    for (int ps = 0; ps < kHop / 1024; ++ps) {
        const long long base = start + (ps * 8 + leaf) * 128 + 2 * pr;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const long long p0 = base + 8 * i;
            float2 x0, x1;
            pcm_pair<FMT>(src, p0, tr.total, word_aligned, x0, x1);
            if (p0 >= 0 && p0 + 1 < tr.total && dst16) {
                *reinterpret_cast<float4*>(dst + p0) = make_float4(x0.x, x0.y, x1.x, x1.y);
            } else {
                if (p0 >= 0 && p0 < tr.total) dst[p0] = x0;
                if (p0 + 1 >= 0 && p0 + 1 < tr.total) dst[p0 + 1] = x1;
            }
            const float m0 = Arith<float>::msq(x0, 1.0f), m1 = Arith<float>::msq(x1, 1.0f);
            a0 = i ? Arith<float>::add(a0, m0) : m0;
            a1 = i ? Arith<float>::add(a1, m1) : m1;
        }
        float s = Arith<float>::add(a0, a1);
        s = Arith<float>::add(s, __shfl_xor_sync(0xffffffffu, s, 1));
        s = Arith<float>::add(s, __shfl_xor_sync(0xffffffffu, s, 2));
        s = Arith<float>::add(s, __shfl_xor_sync(0xffffffffu, s, 4));
        s = Arith<float>::add(s, __shfl_xor_sync(0xffffffffu, s, 8));
        s = Arith<float>::add(s, __shfl_xor_sync(0xffffffffu, s, 16));
        tot[ps] = s;
    }
    if (lane == 0 && q >= tr.hb_lo && q < tr.hb_hi) hsum[tr.hs_base + q] = (kHop == 2048) ? Arith<float>::add(tot[0], tot[1]) : tot[0];
}

// float -> PCM_24 as libsndfile writes it into a FLAC with clipping switched on (what python-soundfile does; restated in
// audio_io.quantise_pcm24 from src/flac.c f2flac24_clip_array): lrintf(x * 2^23), pinned to [-2^23, 2^23 - 1].  The scale is a
// power of two, so the product is exact and the only rounding is the round-half-even conversion (which saturates).
__device__ __forceinline__ int quant24(float x) {
    const int v = __float2int_rn(__fmul_rn(x, 8388608.0f));
    return max(-8388608, min(8388607, v));
}
__global__ void __launch_bounds__(256) float_to_s24_kernel(const float4* __restrict__ in, unsigned* __restrict__ out, long long n4,
                                                            const float* __restrict__ in_s, unsigned char* __restrict__ out_b, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 x = in[i];
        const unsigned s0 = (unsigned)quant24(x.x) & 0xffffffu, s1 = (unsigned)quant24(x.y) & 0xffffffu;
        const unsigned s2 = (unsigned)quant24(x.z) & 0xffffffu, s3 = (unsigned)quant24(x.w) & 0xffffffu;
        out[3 * i] = s0 | (s1 << 24);
        out[3 * i + 1] = (s1 >> 8) | (s2 << 16);
        out[3 * i + 2] = (s2 >> 16) | (s3 << 8);
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned v = (unsigned)quant24(in_s[i]);
        out_b[3 * i] = (unsigned char)(v & 0xff);
        out_b[3 * i + 1] = (unsigned char)((v >> 8) & 0xff);
        out_b[3 * i + 2] = (unsigned char)((v >> 16) & 0xff);
    }
}

// what a PCM_24 file hands back after a write/read round trip, times a gain: the static-EQ gain-protect pass re-reads its
// own output file (src/layer2_apply_eq.py:220-234)
__global__ void __launch_bounds__(256) requantise_scale_kernel(float* __restrict__ y, long long n, float scale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = __fmul_rn((float)quant24(y[i]) * (1.0f / 8388608.0f), scale);
}

// ------------------------------------------------------------------------------------------------
// N3 validator: conditional spectrum (src/validate_layer1.py:261-389, src/verify_tomatis_15db_v2.py:270-369).
// One CTA per selected frame: Hann-windowed 4096-point transform of L + iR for the input and the output file with the fp64
// FFT of edge_kernel (see spectrum.cuh for why not the float32 stages of stft_kernel -- this is a report, not the hot
// path), channel spectra separated by symmetry, ratio[k] = mean_c|Y_c[k]| / max(mean_c|X_c[k]|, 1e-10), optionally divided
// by its mean over the anchor band.  Ratios are stored bin-major ([2049][ld]): the per-bin median reads a contiguous column.
constexpr int kSpecSmemBytes = kEdgeSmemBytes;

__global__ void __launch_bounds__(256)
spectrum_ratio_kernel(const float2* __restrict__ x, const float2* __restrict__ y, const int* __restrict__ frames, long long ld,
                      const float* __restrict__ win, int a0, int a1, float* __restrict__ ratio) {
    extern __shared__ __align__(16) unsigned char spec_smem[];
    double2* sm = reinterpret_cast<double2*>(spec_smem);
    const int t = threadIdx.x, s = blockIdx.x;
    const long long start = (long long)frames[s] * kHop;
    const int nb = spec_bins_of_thread(t);
    float mag[2][9];
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const float2* src = (side == 0 ? x : y) + start;
        for (int n = t; n < kNfft; n += 256) {
            const cplx64 z = spec_window_sample(src[n], win[n]);
            sm[bitrev12(n)] = make_double2(z.x, z.y);
        }
        __syncthreads();
        fft4096_f64(sm, t);                                 // ends with a barrier
#pragma unroll
        for (int i = 0; i < 9; ++i) mag[side][i] = (i < nb) ? spec_mean_mag(reinterpret_cast<const cplx64*>(sm), spec_bin(t, i)) : 0.f;
        __syncthreads();                                    // the next side overwrites the buffer
    }
    float r[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) r[i] = spec_ratio(mag[1][i], mag[0][i]);
    if (a1 >= a0) {
        float* R = reinterpret_cast<float*>(spec_smem);
#pragma unroll
        for (int i = 0; i < 9; ++i) if (i < nb) R[spec_bin(t, i)] = r[i];
        __syncthreads();
        const float g = spec_anchor_gain(R, a0, a1);
        if (g > 0.f) {
#pragma unroll
            for (int i = 0; i < 9; ++i) r[i] = r[i] / g;
        }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) if (i < nb) ratio[(long long)spec_bin(t, i) * ld + s] = r[i];
}

// np.median(ratios, axis=0): one CTA per bin, radix select (4 x 8 bits, most significant first) on the bit patterns of the
// non-negative floats of one column; even counts average the two middle elements in float32 like np.mean.
__global__ void __launch_bounds__(256)
column_median_kernel(const float* __restrict__ ratio, int n, long long ld, float* __restrict__ med) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_rank;
    const unsigned* col = reinterpret_cast<const unsigned*>(ratio + (long long)blockIdx.x * ld);
    float vals[2] = {0.f, 0.f};
    for (int which = 0; which < 2; ++which) {
        unsigned rank = which == 0 ? (unsigned)(n - 1) / 2u : (unsigned)n / 2u;
        unsigned prefix = 0, mask = 0;
        for (int shift = 24; shift >= 0; shift -= 8) {
            hist[threadIdx.x] = 0;
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += 256) {
                const unsigned b = col[i];
                if ((b & mask) == prefix) atomicAdd(&hist[(b >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned cum = 0;
                int d = 0;
                for (; d < 255; ++d) {
                    if (cum + hist[d] > rank) break;
                    cum += hist[d];
                }
                s_prefix = prefix | ((unsigned)d << shift);
                s_rank = rank - cum;
            }
            __syncthreads();
            prefix = s_prefix;
            rank = s_rank;
            mask |= 0xffu << shift;
            __syncthreads();
        }
        vals[which] = __uint_as_float(prefix);
    }
    if (threadIdx.x == 0) med[blockIdx.x] = (n & 1) ? vals[0] : __fmul_rn(__fadd_rn(vals[0], vals[1]), 0.5f);
}

// ------------------------------------------------------------------------------------------------
// N4 calibration front end (src/calibrate_to_baseline_v2.py); per-thread code in calib.cuh.
__global__ void __launch_bounds__(256)
calib_decimate_kernel(const float2* __restrict__ x, long long n_in, const float* __restrict__ h, int len_h, int up, int down,
                      long long n_pre_remove, long long n_out, float* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += stride)
        out[j] = decimate_sample(x, n_in, h, len_h, up, down, j + n_pre_remove);
}

constexpr int kSumBlocks = 1024;
__global__ void __launch_bounds__(256) calib_partial_sum_kernel(const float* __restrict__ v, long long n, double* __restrict__ partial) {
    __shared__ double red[256];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)kSumBlocks * 256) acc += (double)v[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
__global__ void __launch_bounds__(256) calib_subtract_kernel(float* __restrict__ v, long long n, float mean) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[i] = __fsub_rn(v[i], mean);
}

// corr[k] = sum_j a[k + j] * b[j], k in [0, na - nb]: a CTA owns 1024 consecutive lags (4 per thread), b is walked in tiles
// of 2048 staged in shared memory with the matching window of a; each thread slides an 8-value register window over a,
// one 128-bit load per 16 multiply-adds; float32 partial sums are folded into double accumulators once per tile.
constexpr int kXcLags = 1024, kXcTile = 2048;
__global__ void __launch_bounds__(256)
calib_xcorr_kernel(const float* __restrict__ a, long long na, const float* __restrict__ b, int nb, float* __restrict__ corr,
                   long long n_lags) {
    __shared__ __align__(16) float sa[kXcLags + kXcTile + 8];
    __shared__ __align__(16) float sb[kXcTile];
    const int t = threadIdx.x;
    const long long k0 = (long long)blockIdx.x * kXcLags;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j0 = 0; j0 < nb; j0 += kXcTile) {
        __syncthreads();
        for (int i = t; i < kXcLags + kXcTile + 8; i += 256) {
            const long long idx = k0 + j0 + i;
            sa[i] = idx < na ? a[idx] : 0.f;
        }
        for (int i = t; i < kXcTile; i += 256) sb[i] = (j0 + i < nb) ? b[j0 + i] : 0.f;
        __syncthreads();
        float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
        float4 lo = *reinterpret_cast<const float4*>(&sa[4 * t]);
#pragma unroll 4
        for (int j = 0; j < kXcTile; j += 4) {
            const float4 hi = *reinterpret_cast<const float4*>(&sa[4 * t + j + 4]);
            const float4 bb = *reinterpret_cast<const float4*>(&sb[j]);
            f0 = fmaf(lo.x, bb.x, f0); f1 = fmaf(lo.y, bb.x, f1); f2 = fmaf(lo.z, bb.x, f2); f3 = fmaf(lo.w, bb.x, f3);
            f0 = fmaf(lo.y, bb.y, f0); f1 = fmaf(lo.z, bb.y, f1); f2 = fmaf(lo.w, bb.y, f2); f3 = fmaf(hi.x, bb.y, f3);
            f0 = fmaf(lo.z, bb.z, f0); f1 = fmaf(lo.w, bb.z, f1); f2 = fmaf(hi.x, bb.z, f2); f3 = fmaf(hi.y, bb.z, f3);
            f0 = fmaf(lo.w, bb.w, f0); f1 = fmaf(hi.x, bb.w, f1); f2 = fmaf(hi.y, bb.w, f2); f3 = fmaf(hi.z, bb.w, f3);
            lo = hi;
        }
        acc[0] += (double)f0; acc[1] += (double)f1; acc[2] += (double)f2; acc[3] += (double)f3;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const long long k = k0 + 4 * t + r;
        if (k < n_lags) corr[k] = (float)acc[r];
    }
}

// stft_band_tilt: energies of two bin ranges of |rfft(power_mono(frame) * hann)|^2, one CTA per frame (fp64 transform, rounded
// to complex64 like NumPy's result); the band sums are accumulated in double in a fixed order and rounded to float32.
__global__ void __launch_bounds__(256)
calib_band_kernel(const float2* __restrict__ x, const float* __restrict__ win, int lo0, int lo1, int hi0, int hi1,
                  float* __restrict__ e_lo, float* __restrict__ e_hi) {
    extern __shared__ __align__(16) unsigned char band_smem[];
    double2* sm = reinterpret_cast<double2*>(band_smem);
    const int t = threadIdx.x;
    const float2* src = x + (long long)blockIdx.x * kHop;
    for (int n = t; n < kNfft; n += 256) sm[bitrev12(n)] = make_double2((double)__fmul_rn(power_mono(src[n]), win[n]), 0.0);
    __syncthreads();
    fft4096_f64(sm, t);
    const int nb = spec_bins_of_thread(t);
    float pw[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) pw[i] = (i < nb) ? power_bin(reinterpret_cast<const cplx64*>(sm), spec_bin(t, i)) : 0.f;
    __syncthreads();
    float* P = reinterpret_cast<float*>(band_smem);
#pragma unroll
    for (int i = 0; i < 9; ++i) if (i < nb) P[spec_bin(t, i)] = pw[i];
    __syncthreads();
    const int warp = t >> 5, lane = t & 31;
    if (warp < 2) {
        const int k0 = warp == 0 ? lo0 : hi0, k1 = warp == 0 ? lo1 : hi1;
        double s = 0.0;
        for (int k = k0 + lane; k < k1; k += 32) s += (double)P[k];
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) (warp == 0 ? e_lo : e_hi)[blockIdx.x] = (float)s;
    }
}

__global__ void __launch_bounds__(128)
calib_gate_grid_kernel(const float* __restrict__ level, const long long* __restrict__ start, const unsigned char* __restrict__ want,
                       int n, const float* __restrict__ on, const float* __restrict__ off, const long long* __restrict__ delay,
                       int n_combos, int* __restrict__ mismatches, int* __restrict__ switches, unsigned char* __restrict__ states) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_combos) return;
    int mis, sw;
    gate_grid_combo(level, start, want, n, on[c], off[c], delay[c], &mis, &sw, states ? states + (size_t)c * n : nullptr);
    mismatches[c] = mis;
    switches[c] = sw;
}


// ------------------------------------------------------------------------------------------------
// General FFT sizes (EXPERIMENTAL; csrc/generic.cuh).  Coverage path, not the hot path.
template <typename T>
__global__ void __launch_bounds__(128)
gen_meansq_kernel(const float2* __restrict__ x, long long total, long long first_start, int n_fft, int hop, int n_frames, float sc,
                  int mono_file, T* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_frames) out[k] = gen_frame_meansq<T>(x, total, first_start + (long long)k * hop, n_fft, sc, mono_file != 0);
}

__device__ void fft_pow2_f64(double2* sm, int n, int log2n, int t) {      // in: bit-reversed order, out: natural order, forward
    for (int s = 1; s <= log2n; ++s) {
        const int half = 1 << (s - 1);
        for (int b = t; b < (n >> 1); b += 256) {
            const int pos = b & (half - 1);
            const int i = ((b >> (s - 1)) << s) + pos, j = i + half;
            double sn, cs;
            sincospi(-(double)pos / (double)half, &sn, &cs);
            const double2 u = sm[i], v0 = sm[j];
            const double2 v = make_double2(v0.x * cs - v0.y * sn, v0.x * sn + v0.y * cs);
            sm[i] = make_double2(u.x + v.x, u.y + v.y);
            sm[j] = make_double2(u.x - v.x, u.y - v.y);
        }
        __syncthreads();
    }
}

struct GenFrameParams {
    const float2* x;
    long long total, first_start;
    int n_fft, log2n, hop, n_frames;
    const float* win;
    const float* gains;      // [n_rows][n_fft/2 + 1]
    const uint16_t* rows;    // [n_frames]
    float in_scale;
    int flavour;
    void* frames;            // float2 (flavours 0, 1) or double2 (2) [n_frames][n_fft]
};

__global__ void __launch_bounds__(256) gen_frame_kernel(const GenFrameParams prm) {
    extern __shared__ __align__(16) unsigned char gen_smem[];
    double2* sm = reinterpret_cast<double2*>(gen_smem);
    const int t = threadIdx.x, N = prm.n_fft, sh = 32 - prm.log2n;
    const long long pos0 = prm.first_start + (long long)blockIdx.x * prm.hop;
    for (int n = t; n < N; n += 256) {
        double re, im;
        gen_input(gen_sample(prm.x, prm.total, pos0 + n), prm.in_scale, prm.win[n], prm.flavour, &re, &im);
        sm[__brev((unsigned)n) >> sh] = make_double2(re, im);
    }
    __syncthreads();
    fft_pow2_f64(sm, N, prm.log2n, t);
    const float* g = prm.gains + (size_t)prm.rows[blockIdx.x] * (N / 2 + 1);
    for (int k = t; k < N; k += 256) {                   // real symmetric gain, conjugate for the inverse transform
        const double gg = (double)g[k <= N / 2 ? k : N - k];
        const double2 v = sm[k];
        sm[k] = make_double2(v.x * gg, -v.y * gg);
    }
    __syncthreads();
    for (int k = t; k < N; k += 256) {
        const int r = (int)(__brev((unsigned)k) >> sh);
        if (k < r) { const double2 a = sm[k]; sm[k] = sm[r]; sm[r] = a; }
    }
    __syncthreads();
    fft_pow2_f64(sm, N, prm.log2n, t);                   // conj(FFT(conj(Y))) = N * IFFT(Y)
    const double inv_n = 1.0 / (double)N;
    for (int n = t; n < N; n += 256) {
        const double2 v = sm[n];
        const double yr = v.x * inv_n, yi = -v.y * inv_n;
        const float w = prm.win[n];
        const size_t o = (size_t)blockIdx.x * N + n;
        if (prm.flavour == kGenAdaptiveF64) {
            reinterpret_cast<double2*>(prm.frames)[o] = make_double2(yr * (double)w, yi * (double)w);
        } else {
            float ox, oy;
            gen_output_f32(yr, yi, w, prm.flavour, &ox, &oy);
            reinterpret_cast<float2*>(prm.frames)[o] = make_float2(ox, oy);
        }
    }
}

template <typename A, typename F, typename O>
__global__ void __launch_bounds__(256)
gen_ola_kernel(const F* __restrict__ frames, const float* __restrict__ win, long long total, long long first_start, int n_fft, int hop,
               int n_frames, int clamp_norm, float post, O* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += stride) {
        A ox, oy;
        gen_ola_sample<A, F>(frames, win, s, first_start, n_fft, hop, n_frames, clamp_norm != 0, &ox, &oy);
        O o;
        o.x = ox * (A)post;                               // output gain / restored pre-attenuation (1 = identity)
        o.y = oy * (A)post;
        out[s] = o;
    }
}

__device__ __forceinline__ void gen_atomic_max(float* p, float v) { atomicMax(reinterpret_cast<int*>(p), __float_as_int(v)); }
__device__ __forceinline__ void gen_atomic_max(double* p, double v) {
    atomicMax(reinterpret_cast<unsigned long long*>(p), (unsigned long long)__double_as_longlong(v));
}
// per-chunk peak (non-negative values: the bit patterns order like the numbers) and the limiter of write_clamped
template <typename V, typename T>
__global__ void __launch_bounds__(256) gen_peak_kernel(const V* __restrict__ y, const long long* __restrict__ bounds, T* __restrict__ peaks) {
    const long long s0 = bounds[2 * blockIdx.y], s1 = bounds[2 * blockIdx.y + 1];
    T m = 0;
    for (long long i = s0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < s1; i += (long long)gridDim.x * blockDim.x) {
        const V v = y[i];
        m = fmax(m, fmax(fabs(v.x), fabs(v.y)));
    }
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) gen_atomic_max(peaks + blockIdx.y, m);
}
template <typename V, typename T>
__global__ void __launch_bounds__(256) gen_limit_kernel(V* __restrict__ y, const long long* __restrict__ bounds, const T* __restrict__ peaks, T limit) {
    const T peak = peaks[blockIdx.y];
    if (!(peak > limit)) return;
    const T scale = limit / peak;
    const long long s0 = bounds[2 * blockIdx.y], s1 = bounds[2 * blockIdx.y + 1];
    for (long long i = s0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < s1; i += (long long)gridDim.x * blockDim.x) {
        V v = y[i];
        v.x = v.x * scale;
        v.y = v.y * scale;
        y[i] = v;
    }
}
__global__ void __launch_bounds__(256) gen_to_float_kernel(const double2* __restrict__ y, float2* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = make_float2((float)y[i].x, (float)y[i].y);
}

// ================================================================================================
// host side
template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    bool owned = true;
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
    }
    void view(void* base, size_t count) { release(); p = reinterpret_cast<T*>(base); n = count; owned = false; }   // slice of an arena
    void release() { if (p && owned) cudaFree(p); p = nullptr; n = 0; owned = true; }
    ~DevBuf() { release(); }
};

// One device allocation per plan, carved into its arrays (256-byte aligned).  cudaMalloc / cudaFree are the expensive part of
// a plan's life (two dozen of each cost tens of milliseconds and cudaFree synchronises the device), so the engine also keeps
// the arena of the last destroyed plan and hands it to the next one that fits.
struct Arena {
    unsigned char* base = nullptr;
    size_t size = 0, used = 0;
    void* take(size_t bytes) {
        const size_t off = (used + 255) & ~size_t(255);
        used = off + bytes;
        return base ? base + off : nullptr;
    }
};

}  // namespace

struct tmt_engine {
    int device = 0;
    int n_sms = 0;
    DevBuf<float> win;        // [4096]
    DevBuf<float> swin;       // [2][4096]  synthesis window x interior normalisation (eps | clamp)
    std::mutex stage_mu;                    // pinned staging buffer of tmt_plan_read_many
    unsigned char* stage = nullptr;
    size_t stage_size = 0;
    std::mutex arena_mu;                    // plans of one engine may be created / destroyed from several host threads
    unsigned char* spare_arena = nullptr;   // arena of the last destroyed plan, reused by the next plan that fits
    size_t spare_arena_size = 0;
    int gate_nseg = 0;        // > 0: force this many gate-scan segments per track (TMT_GATE_NSEG, tests)
    DevBuf<float2> tw_bases;  // [256][4]
    DevBuf<float2> tw_a;      // [256][16]
    DevBuf<float> gperm;      // [n_rows][4096]
    DevBuf<float> gnat;       // [n_rows][2049] natural order (fp64 edge frames)
    int n_rows = 0;
    bool have_win = false;
    // fork / join of the fp64 edge frames beside the STFT kernel (see launch_edges_beside): one side stream per engine, events
    // recycled through a small pool (a plan takes a pair at its first fork and returns it when it is destroyed)
    cudaStream_t side = nullptr;
    std::mutex ev_mu;
    std::vector<cudaEvent_t> ev_pool;
};

struct HostTrack {
    tmt_track_desc d;
    int n_frames, frame_base, hs_base, chunk_base, n_chunks;
    int hb_lo, hb_hi, f_lo, f_hi;
    int edge_lo = 0, edge_hi = 0;
    long long first_start;
    long long oc_lo = 0, oc_hi = 0;                              // output positions [oc_lo, oc_hi)
    std::vector<std::pair<long long, long long>> chunk_ranges;   // clipped sample ranges
};

struct tmt_plan {
    ~tmt_plan() { if (arena.base) cudaFree(arena.base); }      // only on error paths; tmt_plan_destroy recycles the arena first
    tmt_engine* e = nullptr;
    int framing = 0;
    int n_tracks = 0, total_frames = 0, total_chunks = 0, n_units = 0;
    int max_hb = 0, max_frames = 0;
    long long max_chunk_len = 0, max_in_len = 0;
    long long total_blocks = 0;  // output hop blocks of all work units
    std::vector<HostTrack> ht;
    std::vector<TrackDev> tracks_h;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    Arena arena;
    DevBuf<TrackDev> tracks;
    DevBuf<UnitDev> units;
    DevBuf<ChunkDev> chunks;
    DevBuf<int> chunk_done;     // fused limiter: finished work units per chunk
    DevBuf<int> unit_counter;   // stft_kernel work queue head
    DevBuf<unsigned long long> dbg;   // [2] dev counters of the fused limiter (see StftParams)
    int n_unfusable = 0;        // chunks the fused limiter must leave to limiter_kernel
    DevBuf<EdgeDev> edges;
    int n_edges = 0;
    DevBuf<float> edge_in_scale, edge_out_scale;
    DevBuf<double> hsum;        // float or double view, [total_frames + n_tracks]
    DevBuf<double> msq;         // float or double view, [total_frames]
    DevBuf<double> gate_f64;    // [total_frames]
    DevBuf<uint8_t> state;
    DevBuf<uint16_t> rows;
    DevBuf<int> c2;
    DevBuf<uint8_t> segmap;     // gate scan scratch: [tracks][segments][states]
    DevBuf<int> segclamp;       // [tracks][segments][3]
    int seg_cap = 0;            // segments available in the scratch (all tracks together)
    DevBuf<float> chunk_peaks, in_peaks, in_scale;
    DevBuf<double> von, voff;
    DevBuf<double> bis_in;      // threshold search: [3][tracks] T_low, T_high, start value
    DevBuf<int> bis_active;     // [tracks]
    DevBuf<double> bis_T;       // [tracks] result
    DevBuf<int> bis_iters;      // [tracks]
    DevBuf<double> bis_trace_T; // [tracks][kBisectMaxIter]
    DevBuf<int> bis_trace_c2;   // [tracks][kBisectMaxIter]
    int64_t launches = 0;
};

namespace {

int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// frames of one file under each framing (see oracle/tomatis_oracle.py frame_layout_streaming and
// src/process_tomatis_adaptive.py:298-300)
int count_frames(int framing, long long total) {
    if (framing == TMT_FRAMING_STREAMING) {
        const long long pad = kNfft / 2;
        long long r = (total - kNfft) % kHop;          // C++ % truncates toward zero: emulate Python's
        if (r < 0) r += kHop;
        const long long pad_end = (kHop - r) % kHop;
        const long long length = pad + total + pad_end;
        if (length < kNfft) return 0;
        return (int)((length - kNfft) / kHop + 1);
    }
    if (framing == TMT_FRAMING_EQ_PAD) return (int)(total / kHop + 1);          // n_fft/2 zeros on both sides, src/layer2_apply_eq.py:112-125,205-208
    if (framing == TMT_FRAMING_EQ_NOPAD) return total >= kNfft ? (int)((total - kNfft) / kHop + 1) : 0;
    return (int)(total / kHop);
}

// limiter chunks in block units [blo, bhi) (block b = positions [first_start + b*hop, +hop))
void chunk_blocks(int framing, int n_frames, std::vector<std::pair<int, int>>& out) {
    out.clear();
    if (n_frames <= 0) return;
    const int n_blocks = n_frames + 1;
    if (framing != TMT_FRAMING_STREAMING) { out.push_back({0, n_blocks}); return; }
    // replay of the flush rule src/process_tomatis.py:419-426: positions relative to out_base = -pad
    long long out_base = 0, next_start = 0;       // both shifted by +pad
    int flushed = 0;
    for (int j = 0; j < n_frames; ++j) {
        next_start += kHop;
        const long long safe = (next_start - out_base) - kNfft;
        if (safe >= kFlushSafe) {
            const int nb = (int)(safe / kHop);     // safe is a multiple of hop
            out.push_back({flushed, flushed + nb});
            flushed += nb;
            out_base += safe;
        }
    }
    if (flushed < n_blocks) out.push_back({flushed, n_blocks});
}

void fill_tracks_host(tmt_plan* p) {
    p->tracks_h.resize(p->n_tracks);
    for (int i = 0; i < p->n_tracks; ++i) {
        const HostTrack& h = p->ht[i];
        TrackDev& t = p->tracks_h[i];
        t.in = reinterpret_cast<const float2*>(h.d.pcm_in);
        t.out = reinterpret_cast<float2*>(h.d.pcm_out);
        t.total = h.d.total;
        t.in_origin = h.d.in_origin;
        t.in_lo = std::max<long long>(0, h.d.in_origin);
        t.in_hi = std::min<long long>(h.d.total, h.d.in_origin + h.d.in_len);
        t.out_origin = h.d.out_origin;
        t.out_lo = std::max<long long>(h.oc_lo, h.d.out_origin);
        t.out_hi = std::min<long long>(h.oc_hi, h.d.out_origin + h.d.out_len);
        t.first_start = h.first_start;
        t.n_frames = h.n_frames;
        t.frame_base = h.frame_base;
        t.hs_base = h.hs_base;
        t.hb_lo = h.hb_lo; t.hb_hi = h.hb_hi; t.f_lo = h.f_lo; t.f_hi = h.f_hi;
        t.chunk_base = h.chunk_base; t.n_chunks = h.n_chunks;
        t.edge_lo = h.edge_lo; t.edge_hi = h.edge_hi;
    }
}

int build_tracks_dev(tmt_plan* p) {
    fill_tracks_host(p);
    if (p->n_tracks)
        CUDA_TRY(cudaMemcpy(p->tracks.p, p->tracks_h.data(), sizeof(TrackDev) * p->n_tracks, cudaMemcpyHostToDevice));
    return TMT_OK;
}

template <typename T, int AUTO>
int launch_gate(tmt_plan* p, const T* vals, int param, int S, int X, int alpha_init, int count_only, cudaStream_t st) {
    if (S > kMaxGateStates) {            // more states than the scan's byte maps hold: exact serial walk (see gate_serial_kernel)
        gate_serial_kernel<T, AUTO><<<p->n_tracks, 32, 0, st>>>(p->tracks.p, vals, p->von.p, p->voff.p, param, X, alpha_init, count_only,
                                                               p->state.p, p->rows.p, p->c2.p);
        p->launches++;
        CUDA_TRY(cudaGetLastError());
        return TMT_OK;
    }
    // long tracks: cut into segments of >= 8 frames per thread so the scan spreads over the whole GPU
    int nseg = 1;
    if (p->max_frames > 16384) nseg = std::min(2 * p->e->n_sms, std::max(1, p->max_frames / 2048));
    if (p->e->gate_nseg > 0) nseg = p->e->gate_nseg;
    nseg = std::max(1, std::min(nseg, p->seg_cap / std::max(p->n_tracks, 1)));
    int NT = (nseg > 1) ? 256 : 1024;
    auto need = [&](int nt) {
        return (size_t)((S * nt + S * (nt / 32) + nt + nt / 32 + 15) & ~15) + sizeof(int) * (size_t)(4 * nt + 5 * (nt / 32) + 2) +
               (size_t)nseg * std::max(S, 12);
    };
    while (NT > 32 && need(NT) > 96 * 1024) NT >>= 1;
    const size_t smem = need(NT);
    if (smem > 200 * 1024) return fail(TMT_ERR_UNSUPPORTED, "gate automaton with %d states needs %zu B of shared memory", S, smem);
    GateSeg seg{nseg, p->segmap.p, p->segclamp.p};
    const dim3 grid(nseg, p->n_tracks);
    const int attr = (int)std::max<size_t>(smem, 48 * 1024);
#define TMT_GATE_LAUNCH(PASS)                                                                                          \
    do {                                                                                                               \
        CUDA_TRY(cudaFuncSetAttribute(gate_kernel<T, AUTO, PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, attr)); \
        gate_kernel<T, AUTO, PASS><<<grid, NT, smem, st>>>(p->tracks.p, vals, p->von.p, p->voff.p, param, S, X, alpha_init, \
                                                           count_only, p->state.p, p->rows.p, p->c2.p, seg);          \
        p->launches++;                                                                                                 \
        CUDA_TRY(cudaGetLastError());                                                                                  \
    } while (0)
    if (nseg == 1) {
        TMT_GATE_LAUNCH(GATE_FUSED);
    } else {
        if ((size_t)nseg * S > p->segmap.n / std::max(p->n_tracks, 1))
            return fail(TMT_ERR_UNSUPPORTED, "gate scan scratch too small for %d segments x %d states", nseg, S);
        CUDA_TRY(cudaMemsetAsync(p->c2.p, 0, sizeof(int) * p->n_tracks, st));
        TMT_GATE_LAUNCH(GATE_MAP);
        TMT_GATE_LAUNCH(GATE_STATES);
        if (!count_only) TMT_GATE_LAUNCH(GATE_ROWS);
    }
#undef TMT_GATE_LAUNCH
    return TMT_OK;
}

struct ArrInfo { void* ptr; size_t elem; size_t count; };
int arr_info(tmt_plan* p, int which, ArrInfo* a) {
    switch (which) {
        case TMT_ARR_MEANSQ_F32: *a = {p->msq.p, 4, (size_t)p->total_frames}; return TMT_OK;
        case TMT_ARR_MEANSQ_F64: *a = {p->msq.p, 8, (size_t)p->total_frames}; return TMT_OK;
        case TMT_ARR_GATE_F64: *a = {p->gate_f64.p, 8, (size_t)p->total_frames}; return TMT_OK;
        case TMT_ARR_STATE: *a = {p->state.p, 1, (size_t)p->total_frames}; return TMT_OK;
        case TMT_ARR_ROW: *a = {p->rows.p, 2, (size_t)p->total_frames}; return TMT_OK;
        case TMT_ARR_C2_COUNT: *a = {p->c2.p, 4, (size_t)p->n_tracks}; return TMT_OK;
        case TMT_ARR_CHUNK_PEAK: *a = {p->chunk_peaks.p, 4, (size_t)p->total_chunks}; return TMT_OK;
        case TMT_ARR_INPUT_PEAK: *a = {p->in_peaks.p, 4, (size_t)p->n_tracks}; return TMT_OK;
        case TMT_ARR_HOPSUM_F32: *a = {p->hsum.p, 4, (size_t)(p->total_frames + p->n_tracks)}; return TMT_OK;
        case TMT_ARR_HOPSUM_F64: *a = {p->hsum.p, 8, (size_t)(p->total_frames + p->n_tracks)}; return TMT_OK;
        case TMT_ARR_BISECT_T: *a = {p->bis_T.p, 8, (size_t)p->n_tracks}; return TMT_OK;
        case TMT_ARR_BISECT_ITERS: *a = {p->bis_iters.p, 4, (size_t)p->n_tracks}; return TMT_OK;
        case TMT_ARR_BISECT_TRACE_T: *a = {p->bis_trace_T.p, 8, (size_t)p->n_tracks * kBisectMaxIter}; return TMT_OK;
        case TMT_ARR_BISECT_TRACE_C2: *a = {p->bis_trace_c2.p, 4, (size_t)p->n_tracks * kBisectMaxIter}; return TMT_OK;
    }
    return fail(TMT_ERR_INVALID, "unknown plan array %d", which);
}

}  // namespace

// ================================================================================================
extern "C" {

int tmt_version(void) { return 100; }

const char* tmt_error_string(int code) {
    switch (code) {
        case TMT_OK: return "ok";
        case TMT_ERR_INVALID: return "invalid argument";
        case TMT_ERR_CUDA: return "CUDA error";
        case TMT_ERR_UNSUPPORTED: return "unsupported configuration";
        case TMT_ERR_NOMEM: return "out of memory";
    }
    return "unknown error";
}

const char* tmt_last_error(void) { return g_err; }

int tmt_engine_create(tmt_engine** out, int device, int n_fft, int hop) {
    if (!out) return fail(TMT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n_fft != kNfft || hop != kHop)
        return fail(TMT_ERR_UNSUPPORTED, "only n_fft=%d hop=%d is implemented on the GPU path (got %d/%d)", kNfft, kHop, n_fft, hop);
    CUDA_TRY(cudaSetDevice(device));
    tmt_engine* e = new (std::nothrow) tmt_engine();
    if (!e) return fail(TMT_ERR_NOMEM, "host allocation failed");
    e->device = device;
    cudaDeviceProp prop;
    cudaError_t ce = cudaGetDeviceProperties(&prop, device);
    if (ce != cudaSuccess) { delete e; return fail(TMT_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(ce)); }
    e->n_sms = prop.multiProcessorCount;
    auto twb = build_tw_bases();
    auto twa = build_tw_stage_a();
    if (e->tw_bases.alloc(twb.size()) != cudaSuccess || e->tw_a.alloc(twa.size()) != cudaSuccess || e->win.alloc(kNfft) != cudaSuccess ||
        e->swin.alloc(2 * kNfft) != cudaSuccess) {
        delete e;
        return fail(TMT_ERR_NOMEM, "device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    cudaMemcpy(e->tw_bases.p, twb.data(), sizeof(float2) * twb.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(e->tw_a.p, twa.data(), sizeof(float2) * twa.size(), cudaMemcpyHostToDevice);
    tw64_init_kernel<<<8, 256>>>();
    ce = cudaDeviceSynchronize();           // one-time: later work may run on non-blocking streams that do not order after this launch
    if (ce != cudaSuccess) { delete e; return fail(TMT_ERR_CUDA, "twiddle table initialisation: %s", cudaGetErrorString(ce)); }
    ce = cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStftSmem);
    if (ce != cudaSuccess) { delete e; return fail(TMT_ERR_CUDA, "cudaFuncSetAttribute(stft_kernel): %s", cudaGetErrorString(ce)); }
    if (const char* sv = getenv("TMT_GATE_NSEG")) e->gate_nseg = atoi(sv);

    ce = cudaFuncSetAttribute(edge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kEdgeKernelSmemBytes);
    if (ce != cudaSuccess) { delete e; return fail(TMT_ERR_CUDA, "cudaFuncSetAttribute(edge_kernel): %s", cudaGetErrorString(ce)); }
    *out = e;
    return TMT_OK;
}

int tmt_engine_destroy(tmt_engine* e) {
    if (!e) return TMT_OK;
    cudaSetDevice(e->device);
    if (e->spare_arena) cudaFree(e->spare_arena);
    if (e->stage) cudaFreeHost(e->stage);
    for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
    if (e->side) cudaStreamDestroy(e->side);
    delete e;
    return TMT_OK;
}

int tmt_engine_set_window(tmt_engine* e, const float* win, int n) {
    if (!e || !win || n != kNfft) return fail(TMT_ERR_INVALID, "window must have %d taps", kNfft);
    CUDA_TRY(cudaSetDevice(e->device));
    // float32 arithmetic exactly as the reference: win2 = (win*win).astype(float32); w_buf = win2[n+hop] + win2[n]
    // interior blocks are covered by two frames: y = (sum of windowed frames) / (w2[n+hop] + w2[n] (+eps | clamped));
    // the reciprocal is folded into the synthesis window the kernel multiplies with
    std::vector<float> sw(2 * kNfft);
    for (int i = 0; i < kHop; ++i) {
        const volatile float lo = win[i] * win[i];
        const volatile float hi = win[i + kHop] * win[i + kHop];
        const volatile float nrm = hi + lo;
        const volatile float d0 = nrm + 1e-12f;
        const double r0 = 1.0 / (double)d0;                                  // src/process_tomatis.py:422
        const double r1 = 1.0 / (double)std::max((float)nrm, 1e-8f);         // src/process_tomatis_adaptive.py:330
        sw[i] = (float)((double)win[i] * r0);
        sw[i + kHop] = (float)((double)win[i + kHop] * r0);
        sw[kNfft + i] = (float)((double)win[i] * r1);
        sw[kNfft + i + kHop] = (float)((double)win[i + kHop] * r1);
    }
    CUDA_TRY(cudaMemcpy(e->win.p, win, sizeof(float) * kNfft, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(e->swin.p, sw.data(), sizeof(float) * 2 * kNfft, cudaMemcpyHostToDevice));
    e->have_win = true;
    return TMT_OK;
}

int tmt_engine_set_gain_rows(tmt_engine* e, const float* rows, int n_rows, int n_bins) {
    if (!e || !rows || n_rows <= 0 || n_bins != kNfft / 2 + 1) return fail(TMT_ERR_INVALID, "gain rows must be [n_rows>0][%d]", kNfft / 2 + 1);
    if (n_rows > 65535) return fail(TMT_ERR_UNSUPPORTED, "at most 65535 gain rows");
    CUDA_TRY(cudaSetDevice(e->device));
    std::vector<float> perm((size_t)n_rows * kPassLen);
    for (int r = 0; r < n_rows; ++r) permute_gain_row(rows + (size_t)r * n_bins, perm.data() + (size_t)r * kPassLen);
    if ((size_t)n_rows * kPassLen > e->gperm.n) {
        if (e->gperm.alloc((size_t)n_rows * kPassLen) != cudaSuccess) return fail(TMT_ERR_NOMEM, "gain table allocation failed");
    }
    CUDA_TRY(cudaMemcpy(e->gperm.p, perm.data(), sizeof(float) * perm.size(), cudaMemcpyHostToDevice));
    if ((size_t)n_rows * n_bins > e->gnat.n) {
        if (e->gnat.alloc((size_t)n_rows * n_bins) != cudaSuccess) return fail(TMT_ERR_NOMEM, "gain table allocation failed");
    }
    CUDA_TRY(cudaMemcpy(e->gnat.p, rows, sizeof(float) * (size_t)n_rows * n_bins, cudaMemcpyHostToDevice));
    e->n_rows = n_rows;
    return TMT_OK;
}

int tmt_plan_create(tmt_engine* e, tmt_plan** out, int framing, int n_tracks, const tmt_track_desc* tracks, int unit_blocks) {
    if (!e || !out || n_tracks < 0 || (n_tracks > 0 && !tracks)) return fail(TMT_ERR_INVALID, "bad arguments");
    if (framing < TMT_FRAMING_STREAMING || framing > TMT_FRAMING_EQ_NOPAD) return fail(TMT_ERR_INVALID, "unknown framing %d", framing);
    const bool eq_framing = (framing == TMT_FRAMING_EQ_PAD || framing == TMT_FRAMING_EQ_NOPAD);
    if (n_tracks > 65535) return fail(TMT_ERR_UNSUPPORTED, "at most 65535 tracks per plan");
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(e->device));
    tmt_plan* p = new (std::nothrow) tmt_plan();
    if (!p) return fail(TMT_ERR_NOMEM, "host allocation failed");
    p->e = e;
    p->framing = framing;
    p->n_tracks = n_tracks;
    bool jitter = true;
    if (unit_blocks <= 0) {
        // Default unit length.  Small jobs first: when the whole plan fits ONE wave of the persistent grid (2 CTAs per SM), the
        // shortest units that still give every CTA at most one of them finish earliest -- the time of the launch is the longest
        // unit plus its warm-up frame -- and uneven lengths would only lengthen it (a 10-minute file: 293 units of 48 blocks
        // instead of 1 758 of 8 with a frame of warm-up each; a 60 s file: 264 units of 5 instead of 165 of 8).
        // Otherwise: a unit costs one redundant warm-up frame (1/u of its work), and the dynamic queue leaves about one unit of
        // idle time per CTA at the end (u * CTAs / blocks of the work); the sum is smallest at u = sqrt(blocks / CTAs).  59 (half
        // a limiter chunk) is the cap: 128 tracks x 5 min gives 53.
        long long blocks = 0;
        std::vector<int> lens;
        std::vector<std::pair<int, int>> cbs;
        for (int i = 0; i < n_tracks; ++i) {
            const int nfr = count_frames(framing, tracks[i].total);
            const long long nb = nfr > 0 ? nfr + 1 : 0;
            const long long lo = std::max<long long>(0, tracks[i].block_lo), hi = tracks[i].block_hi < 0 ? nb : std::min<long long>(nb, tracks[i].block_hi);
            blocks += std::max<long long>(0, hi - lo);
            chunk_blocks(framing, nfr, cbs);
            for (const auto& c : cbs) {
                const long long a = std::max<long long>(c.first, lo), b = std::min<long long>(c.second, hi);
                if (b > a) lens.push_back((int)(b - a));
            }
        }
        // an SM that hosts an fp64 edge CTA (running beside the STFT kernel) has room for one STFT CTA only
        const long long slots = 2LL * e->n_sms - std::min<long long>(2LL * n_tracks, e->n_sms / 2);
        int one_wave = 0;
        constexpr int kBpp = 2048 / kHop;            // hop blocks per pass of the kernel: unit lengths scale with it
        if (blocks > 0 && blocks <= 59LL * kBpp * slots && (long long)lens.size() <= slots) {
            for (int u = (int)std::max<long long>(1, (blocks + slots - 1) / slots); u <= 59 * kBpp && !one_wave; ++u) {
                long long n = 0;
                for (int len : lens) n += (len + u - 1) / u;
                if (n <= slots) one_wave = u;
            }
        }
        if (one_wave) {
            unit_blocks = one_wave;
            jitter = false;
        } else {
            unit_blocks = kBpp * (int)std::lround(std::sqrt((double)blocks / kBpp / (2.0 * e->n_sms)));
            unit_blocks = std::max(8 * kBpp, std::min(59 * kBpp, unit_blocks));
        }
    }
    std::vector<UnitDev> units;
    std::vector<ChunkDev> chunks;
    std::vector<std::pair<int, int>> cb;
    std::vector<EdgeDev> edges;
    long long frames = 0;
    for (int i = 0; i < n_tracks; ++i) {
        HostTrack h;
        h.d = tracks[i];
        if (h.d.total < 0 || h.d.in_len < 0 || h.d.out_len < 0) { delete p; return fail(TMT_ERR_INVALID, "track %d: negative length", i); }
        if (h.d.total > 0 && (!h.d.pcm_in || (!h.d.pcm_out && h.d.out_len > 0))) { delete p; return fail(TMT_ERR_INVALID, "track %d: NULL audio buffer", i); }
        h.first_start = (framing == TMT_FRAMING_STREAMING || framing == TMT_FRAMING_EQ_PAD) ? -(long long)(kNfft / 2) : 0;
        h.n_frames = count_frames(framing, h.d.total);
        // positions that belong to the output: the file itself, or (static EQ) everything the frames cover
        h.oc_lo = eq_framing ? h.first_start : 0;
        h.oc_hi = eq_framing ? h.first_start + (h.n_frames > 0 ? (long long)(h.n_frames + 1) * kHop : 0) : h.d.total;
        h.frame_base = (int)frames;
        h.hs_base = (int)frames + i;
        frames += h.n_frames;
        if (frames > 0x7fff0000LL) { delete p; return fail(TMT_ERR_UNSUPPORTED, "too many frames in one plan"); }
        const int n_blocks = h.n_frames > 0 ? h.n_frames + 1 : 0;
        int blo = (int)std::max<long long>(0, h.d.block_lo);
        int bhi = (h.d.block_hi < 0) ? n_blocks : (int)std::min<long long>(n_blocks, h.d.block_hi);
        if (bhi < blo) bhi = blo;
        // hop-block sums / frame mean squares this plan computes: frames touching its blocks
        h.f_lo = std::max(0, blo - 1);
        h.f_hi = std::min(h.n_frames, bhi);
        if (h.d.block_hi < 0 && h.d.block_lo <= 0) { h.f_lo = 0; h.f_hi = h.n_frames; }
        h.hb_lo = h.f_lo;
        h.hb_hi = (h.f_hi > h.f_lo) ? h.f_hi + 1 : h.f_lo;
        chunk_blocks(framing, h.n_frames, cb);
        h.chunk_base = (int)chunks.size();
        h.n_chunks = (int)cb.size();
        for (size_t c = 0; c < cb.size(); ++c) {
            long long s0 = h.first_start + (long long)cb[c].first * kHop, s1 = h.first_start + (long long)cb[c].second * kHop;
            s0 = std::max<long long>(h.oc_lo, s0);
            s1 = std::min<long long>(h.oc_hi, s1);
            if (s1 < s0) s1 = s0;
            h.chunk_ranges.push_back({s0, s1});
            chunks.push_back(ChunkDev{i, s0, s1, 0, 0});
            p->max_chunk_len = std::max(p->max_chunk_len, s1 - s0);
            // work units: slices of this chunk restricted to [blo,bhi), skipping blocks with no file samples
            int u0 = std::max(cb[c].first, blo), u1 = std::min(cb[c].second, bhi);
            while (u0 < u1 && h.first_start + (long long)(u0 + 1) * kHop <= h.oc_lo) ++u0;
            while (u1 > u0 && h.first_start + (long long)(u1 - 1) * kHop >= h.oc_hi) --u1;
            if (u1 <= u0) continue;
            const int n_sub = ceil_div(u1 - u0, unit_blocks);
            // interior cut points are jittered by up to +-30 % of a unit (deterministic hash of the chunk index): same
            // number of units and of redundant warm-up frames, but uneven lengths, so the persistent CTAs de-synchronise
            std::vector<int> cut(n_sub + 1);
            for (int s = 0; s <= n_sub; ++s) cut[s] = u0 + (int)((long long)(u1 - u0) * s / n_sub);
            for (int s = 1; s < n_sub && jitter; ++s) {
                const uint32_t hsh = (uint32_t)(chunks.size() * 2654435761u + (uint32_t)s * 40503u) * 2246822519u;
                const int span = std::min(cut[s] - cut[s - 1], cut[s + 1] - cut[s]);
                cut[s] += (int)(((int)((hsh >> 12) & 1023) - 512) * (long long)(span * 3 / 10) / 512);
                cut[s] = std::max(cut[s - 1] + 1, std::min(cut[s], u1 - (n_sub - s)));
            }
            for (int s = 0; s < n_sub; ++s)
                if (cut[s + 1] > cut[s]) { units.push_back(UnitDev{i, cut[s], cut[s + 1], h.chunk_base + (int)c}); chunks.back().n_units++; }
            // the in-kernel limiter needs the whole chunk (every block that holds file samples) in this plan, and
            // per-chunk limiting at all (streaming framing); a whole-file chunk is far too large for one CTA
            chunks.back().fusable = (framing == TMT_FRAMING_STREAMING && s1 > s0 && cb[c].first >= blo && cb[c].second <= bhi) ? 1 : 0;
        }
        // single-frame (ill-conditioned) edge blocks, recomputed in fp64 by edge_kernel
        auto chunk_of = [&](int blk) { for (size_t c = 0; c < cb.size(); ++c) if (blk >= cb[c].first && blk < cb[c].second) return h.chunk_base + (int)c; return h.chunk_base; };
        if (h.n_frames > 0) {
            if ((framing == TMT_FRAMING_WHOLEFILE || eq_framing) && blo <= 0 && bhi > 0) {
                h.edge_lo = 1;
                edges.push_back(EdgeDev{i, 0, 0, chunk_of(0)});
            }
            const int tb = h.n_frames;     // tail block: second half of the last frame only
            if (tb >= blo && tb < bhi && h.first_start + (long long)tb * kHop < h.oc_hi) {
                h.edge_hi = 1;
                edges.push_back(EdgeDev{i, h.n_frames - 1, 1, chunk_of(tb)});
            }
        }
        p->max_hb = std::max(p->max_hb, h.hb_hi - h.hb_lo);
        p->max_frames = std::max(p->max_frames, h.n_frames);
        p->max_in_len = std::max<long long>(p->max_in_len, h.d.in_len);
        p->ht.push_back(std::move(h));
    }
    p->total_frames = (int)frames;
    p->seg_cap = kGateSegCap;
    p->total_chunks = (int)chunks.size();
    // chunks this plan writes samples of (work units or an fp64 edge block) without producing them completely: their peaks must be
    // combined across shards before limiting.  Chunks the plan does not touch at all (other shards' chunks) do not count.
    {
        std::vector<char> touched(chunks.size(), 0);
        for (size_t c = 0; c < chunks.size(); ++c) touched[c] = chunks[c].n_units > 0;
        for (const EdgeDev& ed : edges) if (ed.chunk >= 0 && (size_t)ed.chunk < chunks.size()) touched[ed.chunk] = 1;
        for (size_t c = 0; c < chunks.size(); ++c)
            p->n_unfusable += (chunks[c].fusable || chunks[c].s1 <= chunks[c].s0 || !touched[c]) ? 0 : 1;
    }
    p->n_units = (int)units.size();
    for (const UnitDev& u : units) p->total_blocks += u.b1 - u.b0;
    p->n_edges = (int)edges.size();
    const size_t nf = (size_t)frames, nt = (size_t)n_tracks;
    // one arena for every per-plan device array: first pass sizes it, second pass hands out the slices
    for (int pass = 0; pass < 2; ++pass) {
        Arena& a = p->arena;
        a.used = 0;
#define TMT_SLICE(buf, count) (buf).view(a.take(sizeof(*(buf).p) * (count)), (count))
        TMT_SLICE(p->tracks, std::max<size_t>(nt, 1));
        TMT_SLICE(p->units, std::max<size_t>(units.size(), 1));
        TMT_SLICE(p->chunks, std::max<size_t>(chunks.size(), 1));
        TMT_SLICE(p->chunk_done, chunks.size() + 1);
        TMT_SLICE(p->unit_counter, 1);
        TMT_SLICE(p->dbg, 2);
        TMT_SLICE(p->edges, std::max<size_t>(edges.size(), 1));
        TMT_SLICE(p->edge_in_scale, nt + 1);
        TMT_SLICE(p->edge_out_scale, nt + 1);
        TMT_SLICE(p->hsum, nf + nt + 1);
        TMT_SLICE(p->msq, nf + 1);
        TMT_SLICE(p->gate_f64, nf + 1);
        TMT_SLICE(p->state, nf + 1);
        TMT_SLICE(p->rows, nf + 1);
        TMT_SLICE(p->c2, nt + 1);
        TMT_SLICE(p->segmap, (size_t)kGateSegCap * 256);
        TMT_SLICE(p->segclamp, (size_t)kGateSegCap * 3);
        TMT_SLICE(p->chunk_peaks, chunks.size() + 1);
        TMT_SLICE(p->in_peaks, nt + 1);
        TMT_SLICE(p->in_scale, nt + 1);
        TMT_SLICE(p->von, nt + 1);
        TMT_SLICE(p->voff, nt + 1);
        TMT_SLICE(p->bis_in, 3 * nt + 1);
        TMT_SLICE(p->bis_active, nt + 1);
        TMT_SLICE(p->bis_T, nt + 1);
        TMT_SLICE(p->bis_iters, nt + 1);
        TMT_SLICE(p->bis_trace_T, nt * kBisectMaxIter + 1);
        TMT_SLICE(p->bis_trace_c2, nt * kBisectMaxIter + 1);
#undef TMT_SLICE
        if (pass == 0) {
            const size_t need = a.used + 256;
            {
                std::lock_guard<std::mutex> lock(e->arena_mu);
                if (e->spare_arena && e->spare_arena_size >= need) {
                    a.base = e->spare_arena; a.size = e->spare_arena_size;
                    e->spare_arena = nullptr; e->spare_arena_size = 0;
                }
            }
            if (a.base) {
            } else if (cudaMalloc(reinterpret_cast<void**>(&a.base), need) == cudaSuccess) {
                a.size = need;
            } else {
                delete p;
                return fail(TMT_ERR_NOMEM, "device allocation of %zu bytes failed: %s", need, cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    // the slices hsum | msq | gate_f64 | state | rows are adjacent in the arena: one clear instead of four
    {
        char* z0 = reinterpret_cast<char*>(p->hsum.p);
        char* z1 = reinterpret_cast<char*>(p->rows.p + nf + 1);
        cudaMemset(z0, 0, (size_t)(z1 - z0));
    }
    // descriptors: tracks | units | chunks | chunk_done | unit_counter | edges are adjacent as well -> one packed upload
    fill_tracks_host(p);
    {
        char* d0 = reinterpret_cast<char*>(p->tracks.p);
        char* d1 = reinterpret_cast<char*>(p->edges.p + std::max<size_t>(edges.size(), 1));
        std::vector<char> blob((size_t)(d1 - d0), 0);
        auto put = [&](const void* dev, const void* src, size_t bytes) { if (bytes) memcpy(blob.data() + (reinterpret_cast<const char*>(dev) - d0), src, bytes); };
        put(p->tracks.p, p->tracks_h.data(), sizeof(TrackDev) * p->tracks_h.size());
        put(p->units.p, units.data(), sizeof(UnitDev) * units.size());
        put(p->chunks.p, chunks.data(), sizeof(ChunkDev) * chunks.size());
        put(p->edges.p, edges.data(), sizeof(EdgeDev) * edges.size());
        const cudaError_t ce = cudaMemcpy(d0, blob.data(), blob.size(), cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) { delete p; return fail(TMT_ERR_CUDA, "descriptor upload: %s", cudaGetErrorString(ce)); }
    }
    *out = p;
    return TMT_OK;
}

int tmt_plan_destroy(tmt_plan* p) {
    if (!p) return TMT_OK;
    tmt_engine* e = p->e;
    cudaSetDevice(e->device);
    if (p->ev_fork) {
        std::lock_guard<std::mutex> lock(e->ev_mu);
        e->ev_pool.push_back(p->ev_fork);
        e->ev_pool.push_back(p->ev_join);
        p->ev_fork = p->ev_join = nullptr;
    }
    if (p->arena.base) {
        // work queued on the plan's arrays may still be running: the next owner only touches the arena through stream-ordered
        // calls on the same device after a device-wide synchronisation here (cudaFree would have implied the same)
        cudaDeviceSynchronize();
        unsigned char* to_free = p->arena.base;
        {
            std::lock_guard<std::mutex> lock(e->arena_mu);
            if (p->arena.size > e->spare_arena_size && p->arena.size <= (size_t(1) << 30)) {
                to_free = e->spare_arena;
                e->spare_arena = p->arena.base; e->spare_arena_size = p->arena.size;
            }
        }
        if (to_free) cudaFree(to_free);
        p->arena.base = nullptr;
    }
    delete p;
    return TMT_OK;
}

int tmt_plan_set_buffers(tmt_plan* p, int track, const void* pcm_in, void* pcm_out) {
    if (!p || track < 0 || track >= p->n_tracks) return fail(TMT_ERR_INVALID, "bad track index");
    p->ht[track].d.pcm_in = pcm_in;
    p->ht[track].d.pcm_out = pcm_out;
    p->tracks_h[track].in = reinterpret_cast<const float2*>(pcm_in);
    p->tracks_h[track].out = reinterpret_cast<float2*>(pcm_out);
    CUDA_TRY(cudaSetDevice(p->e->device));
    CUDA_TRY(cudaMemcpy(p->tracks.p + track, &p->tracks_h[track], sizeof(TrackDev), cudaMemcpyHostToDevice));
    return TMT_OK;
}

int tmt_plan_set_level_ranges(tmt_plan* p, int track, int hb_lo, int hb_hi, int f_lo, int f_hi) {
    if (!p || track < 0 || track >= p->n_tracks) return fail(TMT_ERR_INVALID, "bad track index");
    HostTrack& h = p->ht[track];
    const int nb = h.n_frames > 0 ? h.n_frames + 1 : 0;
    if (hb_lo < 0 || hb_hi > nb || hb_hi < hb_lo || f_lo < 0 || f_hi > h.n_frames || f_hi < f_lo)
        return fail(TMT_ERR_INVALID, "level ranges outside the track (%d hop blocks, %d frames)", nb, h.n_frames);
    h.hb_lo = hb_lo; h.hb_hi = hb_hi; h.f_lo = f_lo; h.f_hi = f_hi;
    TrackDev& t = p->tracks_h[track];
    t.hb_lo = hb_lo; t.hb_hi = hb_hi; t.f_lo = f_lo; t.f_hi = f_hi;
    p->max_hb = 0;
    for (const HostTrack& x : p->ht) p->max_hb = std::max(p->max_hb, x.hb_hi - x.hb_lo);
    CUDA_TRY(cudaSetDevice(p->e->device));
    CUDA_TRY(cudaMemcpy(p->tracks.p + track, &t, sizeof(TrackDev), cudaMemcpyHostToDevice));
    return TMT_OK;
}

int tmt_plan_total_frames(const tmt_plan* p) { return p ? p->total_frames : -1; }
int tmt_plan_total_chunks(const tmt_plan* p) { return p ? p->total_chunks : -1; }
int tmt_plan_total_units(const tmt_plan* p) { return p ? p->n_units : -1; }
int tmt_plan_unfusable_chunks(const tmt_plan* p) { return p ? p->n_unfusable : -1; }
int tmt_plan_debug_counters(tmt_plan* p, uint64_t* out2, int reset) {
    if (!p || !out2) return fail(TMT_ERR_INVALID, "bad arguments");
    CUDA_TRY(cudaSetDevice(p->e->device));
    CUDA_TRY(cudaMemcpy(out2, p->dbg.p, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (reset) CUDA_TRY(cudaMemset(p->dbg.p, 0, 2 * sizeof(uint64_t)));
    return TMT_OK;
}
int tmt_plan_track_frames(const tmt_plan* p, int t) { return (p && t >= 0 && t < p->n_tracks) ? p->ht[t].n_frames : -1; }
int tmt_plan_track_frame_base(const tmt_plan* p, int t) { return (p && t >= 0 && t < p->n_tracks) ? p->ht[t].frame_base : -1; }
int tmt_plan_track_chunks(const tmt_plan* p, int t) { return (p && t >= 0 && t < p->n_tracks) ? p->ht[t].n_chunks : -1; }
int tmt_plan_track_chunk_base(const tmt_plan* p, int t) { return (p && t >= 0 && t < p->n_tracks) ? p->ht[t].chunk_base : -1; }
int tmt_plan_chunk_range(const tmt_plan* p, int t, int c, int64_t* s0, int64_t* s1) {
    if (!p || t < 0 || t >= p->n_tracks || c < 0 || c >= p->ht[t].n_chunks || !s0 || !s1) return fail(TMT_ERR_INVALID, "bad chunk index");
    *s0 = p->ht[t].chunk_ranges[c].first;
    *s1 = p->ht[t].chunk_ranges[c].second;
    return TMT_OK;
}
int tmt_plan_geometry(const tmt_plan* p, int32_t* n_frames, int32_t* frame_base, int32_t* n_chunks, int32_t* chunk_base) {
    if (!p) return fail(TMT_ERR_INVALID, "plan is NULL");
    for (int t = 0; t < p->n_tracks; ++t) {
        if (n_frames) n_frames[t] = p->ht[t].n_frames;
        if (frame_base) frame_base[t] = p->ht[t].frame_base;
        if (n_chunks) n_chunks[t] = p->ht[t].n_chunks;
        if (chunk_base) chunk_base[t] = p->ht[t].chunk_base;
    }
    return TMT_OK;
}
int tmt_plan_chunk_ranges(const tmt_plan* p, int64_t* ranges) {
    if (!p || !ranges) return fail(TMT_ERR_INVALID, "bad arguments");
    for (int t = 0; t < p->n_tracks; ++t)
        for (int c = 0; c < p->ht[t].n_chunks; ++c) {
            ranges[2 * ((size_t)p->ht[t].chunk_base + c)] = p->ht[t].chunk_ranges[c].first;
            ranges[2 * ((size_t)p->ht[t].chunk_base + c) + 1] = p->ht[t].chunk_ranges[c].second;
        }
    return TMT_OK;
}
int64_t tmt_plan_launch_count(const tmt_plan* p) { return p ? p->launches : -1; }

int tmt_plan_read(tmt_plan* p, int which, int64_t offset, int64_t count, void* ptr, int is_device, void* stream) {
    if (!p || !ptr) return fail(TMT_ERR_INVALID, "bad arguments");
    ArrInfo a;
    int rc = arr_info(p, which, &a);
    if (rc) return rc;
    if (offset < 0 || count < 0 || (size_t)(offset + count) > a.count) return fail(TMT_ERR_INVALID, "range [%lld,+%lld) outside array %d of %zu", (long long)offset, (long long)count, which, a.count);
    if (count == 0) return TMT_OK;
    CUDA_TRY(cudaSetDevice(p->e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaMemcpyAsync(ptr, static_cast<char*>(a.ptr) + offset * a.elem, count * a.elem,
                             is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (!is_device) CUDA_TRY(cudaStreamSynchronize(st));
    return TMT_OK;
}

int tmt_plan_read_many(tmt_plan* p, int n, const int32_t* which, void* const* dst, void* stream) {
    if (!p || n < 0 || (n > 0 && (!which || !dst))) return fail(TMT_ERR_INVALID, "bad arguments");
    if (n == 0) return TMT_OK;
    tmt_engine* e = p->e;
    std::vector<ArrInfo> info((size_t)n);
    std::vector<size_t> off((size_t)n);
    size_t need = 0;
    for (int i = 0; i < n; ++i) {
        int rc = arr_info(p, which[i], &info[i]);
        if (rc) return rc;
        if (!dst[i]) return fail(TMT_ERR_INVALID, "dst[%d] is NULL", i);
        off[i] = need;
        need += (info[i].elem * info[i].count + 255) & ~size_t(255);
    }
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    std::lock_guard<std::mutex> lock(e->stage_mu);      // one pinned staging buffer per engine
    if (need > e->stage_size) {
        if (e->stage) cudaFreeHost(e->stage);
        e->stage = nullptr; e->stage_size = 0;
        const size_t want = std::max<size_t>(need, size_t(1) << 20);
        if (cudaMallocHost(reinterpret_cast<void**>(&e->stage), want) != cudaSuccess)
            return fail(TMT_ERR_NOMEM, "pinned staging buffer of %zu bytes: %s", want, cudaGetErrorString(cudaGetLastError()));
        e->stage_size = want;
    }
    for (int i = 0; i < n; ++i)
        if (info[i].count) CUDA_TRY(cudaMemcpyAsync(e->stage + off[i], info[i].ptr, info[i].elem * info[i].count, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));                // the only host wait, whatever the number of arrays
    for (int i = 0; i < n; ++i)
        if (info[i].count) memcpy(dst[i], e->stage + off[i], info[i].elem * info[i].count);
    return TMT_OK;
}

int tmt_plan_write(tmt_plan* p, int which, int64_t offset, int64_t count, const void* ptr, int is_device, void* stream) {
    if (!p || !ptr) return fail(TMT_ERR_INVALID, "bad arguments");
    ArrInfo a;
    int rc = arr_info(p, which, &a);
    if (rc) return rc;
    if (offset < 0 || count < 0 || (size_t)(offset + count) > a.count) return fail(TMT_ERR_INVALID, "range outside array %d", which);
    if (count == 0) return TMT_OK;
    CUDA_TRY(cudaSetDevice(p->e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(a.ptr) + offset * a.elem, ptr, count * a.elem,
                             is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    if (!is_device) CUDA_TRY(cudaStreamSynchronize(st));
    return TMT_OK;
}

int tmt_plan_input_peaks(tmt_plan* p, void* stream) {
    if (!p) return fail(TMT_ERR_INVALID, "plan is NULL");
    if (p->n_tracks == 0) return TMT_OK;
    CUDA_TRY(cudaSetDevice(p->e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaMemsetAsync(p->in_peaks.p, 0, sizeof(float) * p->n_tracks, st));
    const int gx = (int)std::max<long long>(1, std::min<long long>(8LL * p->e->n_sms, (p->max_in_len + 256 * 8 - 1) / (256 * 8)));
    input_peak_kernel<<<dim3(gx, p->n_tracks), 256, 0, st>>>(p->tracks.p, p->in_peaks.p);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

int tmt_plan_levels(tmt_plan* p, int flags, const float* in_scale, void* stream) {
    const int use_f64 = flags & TMT_LEVELS_F64;
    const int mono = (flags & TMT_LEVELS_MONO) ? 1 : (flags & TMT_LEVELS_LEFT) ? 2 : (flags & TMT_LEVELS_RIGHT) ? 3
                     : (flags & TMT_LEVELS_POWER_EPS) ? 4 : 0;
    if (!p) return fail(TMT_ERR_INVALID, "plan is NULL");
    if (((flags & TMT_LEVELS_MONO) != 0) + ((flags & TMT_LEVELS_LEFT) != 0) + ((flags & TMT_LEVELS_RIGHT) != 0) +
            ((flags & TMT_LEVELS_POWER_EPS) != 0) > 1)
        return fail(TMT_ERR_INVALID, "TMT_LEVELS_MONO, _LEFT, _RIGHT and _POWER_EPS exclude each other");
    if (p->n_tracks == 0 || (p->max_hb == 0 && !(flags & TMT_LEVELS_MEANSQ_ONLY))) return TMT_OK;
    CUDA_TRY(cudaSetDevice(p->e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const float* sc = nullptr;
    if (in_scale) {
        CUDA_TRY(cudaMemcpyAsync(p->in_scale.p, in_scale, sizeof(float) * p->n_tracks, cudaMemcpyHostToDevice, st));
        sc = p->in_scale.p;
    }
    const bool do_sums = !(flags & TMT_LEVELS_MEANSQ_ONLY), do_msq = !(flags & TMT_LEVELS_HOPSUM_ONLY);
    const dim3 g1(ceil_div(std::max(p->max_hb, 1), kLevelWarps), p->n_tracks);
    const dim3 g2(ceil_div(std::max(p->max_frames, 1), 256), p->n_tracks);
    if (use_f64) {
        if (do_sums) levels_kernel<double><<<g1, kLevelWarps * 32, 0, st>>>(p->tracks.p, sc, p->hsum.p, mono);
        if (do_msq) meansq_kernel<double><<<g2, 256, 0, st>>>(p->tracks.p, p->hsum.p, p->msq.p);
    } else {
        if (do_sums) levels_kernel<float><<<g1, kLevelWarps * 32, 0, st>>>(p->tracks.p, sc, reinterpret_cast<float*>(p->hsum.p), mono);
        if (do_msq) meansq_kernel<float><<<g2, 256, 0, st>>>(p->tracks.p, reinterpret_cast<const float*>(p->hsum.p), reinterpret_cast<float*>(p->msq.p));
    }
    p->launches += (do_sums ? 1 : 0) + (do_msq ? 1 : 0);
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

int tmt_plan_levels_multichannel(tmt_plan* p, int flags, const float* in_scale, const float* x, int channels, void* stream) {
    if (!p || !x) return fail(TMT_ERR_INVALID, "bad arguments");
    if (channels < 1 || channels > 128) return fail(TMT_ERR_INVALID, "channels must lie in [1, 128], got %d", channels);
    if (flags & ~(TMT_LEVELS_F64 | TMT_LEVELS_HOPSUM_ONLY)) return fail(TMT_ERR_INVALID, "flags: TMT_LEVELS_F64 and TMT_LEVELS_HOPSUM_ONLY only");
    if (p->n_tracks == 0 || p->max_hb == 0) return TMT_OK;
    const HostTrack& h0 = p->ht[0];
    for (const HostTrack& h : p->ht)
        if (h.n_frames != h0.n_frames || h.hb_lo != h0.hb_lo || h.hb_hi != h0.hb_hi || h.d.total != h0.d.total || h.d.in_origin != h0.d.in_origin ||
            h.d.in_len != h0.d.in_len)
            return fail(TMT_ERR_INVALID, "the channel pairs of one file must be tracks of identical geometry");
    CUDA_TRY(cudaSetDevice(p->e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const float* sc = nullptr;
    if (in_scale) {
        CUDA_TRY(cudaMemcpyAsync(p->in_scale.p, in_scale, sizeof(float) * p->n_tracks, cudaMemcpyHostToDevice, st));
        sc = p->in_scale.p;
    }
    const dim3 g1(ceil_div(std::max(p->max_hb, 1), kLevelWarps));
    const dim3 g2(ceil_div(std::max(p->max_frames, 1), 256), p->n_tracks);
    const bool do_msq = !(flags & TMT_LEVELS_HOPSUM_ONLY);
    if (flags & TMT_LEVELS_F64) {
        levels_multi_kernel<double><<<g1, kLevelWarps * 32, 0, st>>>(p->tracks.p, p->n_tracks, sc, x, channels, p->hsum.p);
        if (do_msq) meansq_kernel<double><<<g2, 256, 0, st>>>(p->tracks.p, p->hsum.p, p->msq.p);
    } else {
        levels_multi_kernel<float><<<g1, kLevelWarps * 32, 0, st>>>(p->tracks.p, p->n_tracks, sc, x, channels, reinterpret_cast<float*>(p->hsum.p));
        if (do_msq) meansq_kernel<float><<<g2, 256, 0, st>>>(p->tracks.p, reinterpret_cast<const float*>(p->hsum.p), reinterpret_cast<float*>(p->msq.p));
    }
    p->launches += 1 + (do_msq ? 1 : 0);
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

int tmt_channels_split(const float* in, int64_t total, int channels, float* pairs, void* stream) {
    if (!in || !pairs || total < 0 || channels < 1) return fail(TMT_ERR_INVALID, "bad arguments");
    if (total == 0) return TMT_OK;
    const long long n = total * ((channels + 1) / 2);
    channels_split_kernel<<<(int)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        in, total, channels, reinterpret_cast<float2*>(pairs));
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

int tmt_channels_merge(const float* pairs, int64_t total, int channels, float* out, void* stream) {
    if (!out || !pairs || total < 0 || channels < 1) return fail(TMT_ERR_INVALID, "bad arguments");
    if (total == 0) return TMT_OK;
    const long long n = total * ((channels + 1) / 2);
    channels_merge_kernel<<<(int)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2*>(pairs), total, channels, out);
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

// ---- peer-memory exchange of the sharded long file (see peer_publish_kernel) -------------------------------------------
size_t tmt_peer_bytes(int n_hop_blocks) { return kPeerOffHsum + 2 * peer_nb_pad(std::max(n_hop_blocks, 0)) * sizeof(float); }

int tmt_peer_alloc(int device, size_t bytes, void** ptr, unsigned char* handle64) {
    if (!ptr || !handle64 || bytes < kPeerOffHsum) return fail(TMT_ERR_INVALID, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CUDA_TRY(cudaSetDevice(device));
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(TMT_ERR_CUDA, "exchange buffer: %s", cudaGetErrorString(e)); }
    memcpy(handle64, &h, 64);
    *ptr = p;
    return TMT_OK;
}
int tmt_peer_open(int device, const unsigned char* handle64, void** ptr) {
    if (!ptr || !handle64) return fail(TMT_ERR_INVALID, "bad arguments");
    CUDA_TRY(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TMT_OK;
}
int tmt_peer_close(int device, void* ptr) {
    CUDA_TRY(cudaSetDevice(device));
    if (ptr) CUDA_TRY(cudaIpcCloseMemHandle(ptr));
    return TMT_OK;
}
int tmt_peer_free(int device, void* ptr) {
    CUDA_TRY(cudaSetDevice(device));
    if (ptr) CUDA_TRY(cudaFree(ptr));
    return TMT_OK;
}
int tmt_peer_status(int device, const void* base, int32_t* status) {
    if (!base || !status) return fail(TMT_ERR_INVALID, "bad arguments");
    CUDA_TRY(cudaSetDevice(device));
    unsigned v = 0;
    CUDA_TRY(cudaMemcpy(&v, static_cast<const unsigned char*>(base) + 8, 4, cudaMemcpyDeviceToHost));
    *status = (int32_t)v;
    return TMT_OK;
}

int tmt_plan_peer_publish(tmt_plan* p, int rank, int world, void* const* bases, const void* first_hop, const void* last_hop, void* stream) {
    if (!p || !bases || p->n_tracks != 1) return fail(TMT_ERR_INVALID, "peer exchange works on a one-track (one shard) plan");
    if (world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world) return fail(TMT_ERR_INVALID, "bad rank / world (max %d ranks)", kPeerMaxWorld);
    if ((rank > 0 && !first_hop) || (rank + 1 < world && !last_hop)) return fail(TMT_ERR_INVALID, "edge hops missing");
    CUDA_TRY(cudaSetDevice(p->e->device));
    const HostTrack& h = p->ht[0];
    PeerPublishParams prm;
    for (int r = 0; r < kPeerMaxWorld; ++r) prm.bases[r] = r < world ? static_cast<unsigned char*>(bases[r]) : nullptr;
    prm.rank = rank; prm.world = world;
    prm.hsum = reinterpret_cast<const float*>(p->hsum.p) + h.hs_base;
    prm.hb_lo = h.hb_lo; prm.hb_hi = h.hb_hi; prm.nb = h.n_frames > 0 ? h.n_frames + 1 : 0;
    prm.first_hop = static_cast<const float2*>(first_hop); prm.last_hop = static_cast<const float2*>(last_hop);
    const int grid = std::max(1, std::min(32, ceil_div(std::max(h.hb_hi - h.hb_lo, kHop), 256)));
    peer_publish_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(prm);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

int tmt_plan_peer_wait(tmt_plan* p, int world, void* base, void* window, int64_t window_len, int left, int right, double timeout_s, void* stream) {
    if (!p || !base || p->n_tracks != 1 || !window) return fail(TMT_ERR_INVALID, "bad arguments");
    if (world < 1 || world > kPeerMaxWorld || left < 0 || right < 0 || left > kHop || right > kHop || left + right > window_len)
        return fail(TMT_ERR_INVALID, "bad world / halo lengths");
    CUDA_TRY(cudaSetDevice(p->e->device));
    const HostTrack& h = p->ht[0];
    const int nb = h.n_frames > 0 ? h.n_frames + 1 : 0;
    const int grid = std::max(1, std::min(64, ceil_div(std::max(nb, kHop), 1024)));
    peer_wait_unpack_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        static_cast<unsigned char*>(base), world, reinterpret_cast<float*>(p->hsum.p) + h.hs_base, nb, static_cast<float2*>(window), window_len,
        left, right, (unsigned long long)(std::max(timeout_s, 0.001) * 1e9));
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

int tmt_plan_gate(tmt_plan* p, int automaton, int gate_input, const double* on, const double* off, int param,
                  int xfade_frames, int alpha_init_to_target, int count_only, void* stream) {
    if (!p || ((on == nullptr) != (off == nullptr))) return fail(TMT_ERR_INVALID, "bad arguments");
    if (param < 0 || xfade_frames < 0 || xfade_frames > 65534) return fail(TMT_ERR_INVALID, "bad gate parameters");
    if (p->n_tracks == 0) return TMT_OK;
    int S;
    if (automaton == TMT_GATE_UPDELAY) { if (param < 1) return fail(TMT_ERR_INVALID, "run_frames must be >= 1"); S = param + 1; }
    else if (automaton == TMT_GATE_MINHOLD) S = 2 * (param + 1);
    else return fail(TMT_ERR_INVALID, "unknown automaton %d", automaton);
    CUDA_TRY(cudaSetDevice(p->e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (on) {               // NULL / NULL: keep the thresholds already on the device (left there by tmt_plan_bisect)
        CUDA_TRY(cudaMemcpyAsync(p->von.p, on, sizeof(double) * p->n_tracks, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(p->voff.p, off, sizeof(double) * p->n_tracks, cudaMemcpyHostToDevice, st));
    }
    const int X = xfade_frames, ai = alpha_init_to_target ? 1 : 0, co = count_only ? 1 : 0;
    if (gate_input == TMT_ARR_MEANSQ_F32) {
        const float* v = reinterpret_cast<const float*>(p->msq.p);
        return automaton == TMT_GATE_UPDELAY ? launch_gate<float, TMT_GATE_UPDELAY>(p, v, param, S, X, ai, co, st)
                                             : launch_gate<float, TMT_GATE_MINHOLD>(p, v, param, S, X, ai, co, st);
    }
    const double* v = (gate_input == TMT_ARR_MEANSQ_F64) ? p->msq.p : (gate_input == TMT_ARR_GATE_F64) ? p->gate_f64.p : nullptr;
    if (!v) return fail(TMT_ERR_INVALID, "gate_input must be MEANSQ_F32, MEANSQ_F64 or GATE_F64");
    return automaton == TMT_GATE_UPDELAY ? launch_gate<double, TMT_GATE_UPDELAY>(p, v, param, S, X, ai, co, st)
                                         : launch_gate<double, TMT_GATE_MINHOLD>(p, v, param, S, X, ai, co, st);
}

static int plan_bisect_impl(tmt_plan* p, const double* t_low, const double* t_high, const double* start, const int32_t* active, double hyst_db,
                            double target_c2, int hold_frames, int max_iter, int emit, int xfade_frames, int alpha_init, void* stream) {
    if (!p || !t_low || !t_high || !start || !active) return fail(TMT_ERR_INVALID, "bad arguments");
    if (hold_frames < 0 || max_iter < 0 || max_iter > kBisectMaxIter) return fail(TMT_ERR_INVALID, "bad search parameters (max_iter <= %d)", kBisectMaxIter);
    if (p->n_tracks == 0) return TMT_OK;
    const int S = 2 * (hold_frames + 1);
    if (S > kMaxGateStates) return fail(TMT_ERR_UNSUPPORTED, "gate automaton needs %d states (max %d): hold too long for the GPU scan", S, kMaxGateStates);
    if (p->max_frames > 16384 || p->e->gate_nseg > 1)
        return fail(TMT_ERR_UNSUPPORTED, "in-kernel threshold search covers tracks of up to 16384 frames (single-segment scan)");
    CUDA_TRY(cudaSetDevice(p->e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t nt = (size_t)p->n_tracks;
    CUDA_TRY(cudaMemcpyAsync(p->bis_in.p, t_low, sizeof(double) * nt, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(p->bis_in.p + nt, t_high, sizeof(double) * nt, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(p->bis_in.p + 2 * nt, start, sizeof(double) * nt, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(p->bis_active.p, active, sizeof(int) * nt, cudaMemcpyHostToDevice, st));
    int NT = 1024;
    auto need = [&](int n) { return (size_t)((S * n + S * (n / 32) + n + n / 32 + 15) & ~15) + sizeof(int) * (size_t)(n / 32 + 1); };
    while (NT > 32 && need(NT) > 96 * 1024) NT >>= 1;
    const size_t smem = need(NT);
    if (smem > 200 * 1024) return fail(TMT_ERR_UNSUPPORTED, "gate automaton with %d states needs %zu B of shared memory", S, smem);
    CUDA_TRY(cudaFuncSetAttribute(bisect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
    BisectParams prm;
    prm.tracks = p->tracks.p; prm.levels = p->gate_f64.p;
    prm.t_low = p->bis_in.p; prm.t_high = p->bis_in.p + nt; prm.best0 = p->bis_in.p + 2 * nt; prm.active = p->bis_active.p;
    prm.half_hyst = hyst_db / 2; prm.target = target_c2; prm.hold = hold_frames; prm.S = S; prm.max_iter = max_iter;
    prm.von = p->von.p; prm.voff = p->voff.p; prm.best_T = p->bis_T.p; prm.n_iter = p->bis_iters.p;
    prm.trace_T = p->bis_trace_T.p; prm.trace_c2 = p->bis_trace_c2.p;
    prm.emit = 0; prm.X = xfade_frames; prm.alpha_init = alpha_init ? 1 : 0;
    prm.state = p->state.p; prm.rows = p->rows.p; prm.c2_count = p->c2.p;
    const bool fast = S <= 16 && p->max_frames <= 1024 * kBisectFastFrames && !getenv("TMT_BISECT_GENERIC");
    if (fast) {
        prm.emit = (emit && !getenv("TMT_BISECT_NO_EMIT")) ? 1 : 0;      // the fast kernel also writes the final states and rows
        bisect_fast_kernel<<<p->n_tracks, 1024, 0, st>>>(prm);
    } else {
        bisect_kernel<<<p->n_tracks, NT, smem, st>>>(prm);
    }
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    if (emit && !prm.emit)                                                // generic search: the gate scan follows as its own launch
        return tmt_plan_gate(p, TMT_GATE_MINHOLD, TMT_ARR_GATE_F64, nullptr, nullptr, hold_frames, xfade_frames, alpha_init, 0, stream);
    return TMT_OK;
}

int tmt_plan_bisect(tmt_plan* p, const double* t_low, const double* t_high, const double* start, const int32_t* active, double hyst_db,
                    double target_c2, int hold_frames, int max_iter, void* stream) {
    return plan_bisect_impl(p, t_low, t_high, start, active, hyst_db, target_c2, hold_frames, max_iter, 0, 0, 0, stream);
}

int tmt_plan_bisect_gate(tmt_plan* p, const double* t_low, const double* t_high, const double* start, const int32_t* active, double hyst_db,
                         double target_c2, int hold_frames, int max_iter, int xfade_frames, int alpha_init_to_target, void* stream) {
    if (xfade_frames < 0 || xfade_frames > 65534) return fail(TMT_ERR_INVALID, "bad gate parameters");
    return plan_bisect_impl(p, t_low, t_high, start, active, hyst_db, target_c2, hold_frames, max_iter, 1, xfade_frames, alpha_init_to_target,
                            stream);
}

static int launch_stft(tmt_plan* p, float post_gain, float limit, cudaStream_t st) {
    tmt_engine* e = p->e;
    StftParams prm;
    prm.tracks = p->tracks.p;
    prm.units = p->units.p;
    prm.n_units = p->n_units;
    prm.rows = p->rows.p;
    prm.gperm = e->gperm.p;
    prm.win = e->win.p;
    prm.swin = e->swin.p + ((p->framing == TMT_FRAMING_WHOLEFILE) ? kNfft : 0);
    prm.tw_bases = e->tw_bases.p;
    prm.tw_a = e->tw_a.p;
    prm.chunk_peaks = p->chunk_peaks.p;
    prm.post_gain = post_gain;
    prm.chunks = p->chunks.p;
    prm.chunk_done = p->chunk_done.p;
    prm.unit_counter = p->unit_counter.p;
    prm.limit = limit;
    prm.dbg = p->dbg.p;
    CUDA_TRY(cudaMemsetAsync(p->unit_counter.p, 0, sizeof(int), st));
    const int grid = std::min(p->n_units, 2 * e->n_sms);            // persistent: two CTAs per SM
    stft_kernel<<<grid, kThreads, kStftSmem, st>>>(prm);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

int tmt_plan_clear_peaks(tmt_plan* p, void* stream) {
    if (!p) return fail(TMT_ERR_INVALID, "plan is NULL");
    CUDA_TRY(cudaSetDevice(p->e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (p->total_chunks) {
        CUDA_TRY(cudaMemsetAsync(p->chunk_peaks.p, 0, sizeof(float) * p->total_chunks, st));
        CUDA_TRY(cudaMemsetAsync(p->chunk_done.p, 0, sizeof(int) * p->total_chunks, st));
    }
    return TMT_OK;
}

int tmt_plan_stft(tmt_plan* p, float post_gain, int skip_edges, void* stream) {
    if (!p) return fail(TMT_ERR_INVALID, "plan is NULL");
    tmt_engine* e = p->e;
    if (!e->have_win || e->n_rows == 0) return fail(TMT_ERR_INVALID, "engine window / gain rows not set");
    int rc = tmt_plan_clear_peaks(p, stream);
    if (rc) return rc;
    if (p->n_units == 0) return TMT_OK;
    rc = launch_stft(p, post_gain, 0.f, reinterpret_cast<cudaStream_t>(stream));
    if (rc) return rc;
    // the single-frame edge blocks are never produced by stft_kernel; without an explicit
    // tmt_plan_edge_frames call they are filled in here with default scales
    if (!skip_edges) return tmt_plan_edge_frames(p, post_gain, nullptr, nullptr, 0, stream);
    return TMT_OK;
}

static bool stft_fused_limiter(const tmt_plan* p) {
    bool fuse = p->total_blocks >= 48LL * (2048 / kHop) * p->e->n_sms;
    if (const char* fv = getenv("TMT_LIMITER_FUSED")) fuse = atoi(fv) != 0;        // dev switch: A/B of the two limiter placements
    return fuse;
}

int tmt_plan_stft_limited(tmt_plan* p, float post_gain, float limit, void* stream) {
    if (!p) return fail(TMT_ERR_INVALID, "plan is NULL");
    tmt_engine* e = p->e;
    if (!e->have_win || e->n_rows == 0) return fail(TMT_ERR_INVALID, "engine window / gain rows not set");
    if (!(limit > 0.f)) return fail(TMT_ERR_INVALID, "limit must be positive");
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // Small jobs (a single short file): the kernel lasts a few frame times, every chunk finishes at its very end, and the in-kernel
    // rescale -- one CTA per chunk, latency-bound -- would be a serial tail.  The separate limiter pass spreads the same bytes over
    // all SMs and costs one launch.  Large jobs hide the rescale under the other CTAs' butterflies and keep it fused.
    const bool fuse = stft_fused_limiter(p);
    if (p->n_units) {
        int rc = launch_stft(p, post_gain, fuse ? limit : 0.f, st);
        if (rc) return rc;
    }
    if ((fuse && p->n_unfusable == 0) || p->max_chunk_len == 0) return TMT_OK;
    const long long per_cta = 256LL * 16;
    int gx = (int)std::min<long long>((p->max_chunk_len + per_cta - 1) / per_cta, 8LL * e->n_sms);
    limiter_kernel<<<dim3(std::max(gx, 1), p->total_chunks), 256, 0, st>>>(p->tracks.p, p->chunks.p, p->chunk_peaks.p, limit, fuse ? 1 : 0);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

// The fp64 edge frames (<= 2 CTAs per track, latency-bound: 45 - 60 us whatever the job) write blocks the STFT kernel skips
// and only share the chunk-peak atomics with it, so the two can run side by side.  beside = true: the edge kernel goes to the
// engine's side stream, forked from `st` here; the caller launches the STFT kernel on `st` and then calls join_edges.  A single
// 60 s file spends more time in edge_kernel than in stft_kernel; this takes it off the critical path.
static int launch_edges(tmt_plan* p, float post_gain, const float* in_scale, const float* out_scale, int pipeline_f64, cudaStream_t st,
                        bool beside) {
    tmt_engine* e = p->e;
    EdgeParams prm;
    prm.tracks = p->tracks.p;
    prm.edges = p->edges.p;
    prm.rows = p->rows.p;
    prm.gnat = e->gnat.p;
    prm.win = e->win.p;
    prm.in_scale = nullptr;
    prm.out_scale = nullptr;
    if (in_scale) {
        CUDA_TRY(cudaMemcpyAsync(p->edge_in_scale.p, in_scale, sizeof(float) * p->n_tracks, cudaMemcpyHostToDevice, st));
        prm.in_scale = p->edge_in_scale.p;
    }
    if (out_scale) {
        CUDA_TRY(cudaMemcpyAsync(p->edge_out_scale.p, out_scale, sizeof(float) * p->n_tracks, cudaMemcpyHostToDevice, st));
        prm.out_scale = p->edge_out_scale.p;
    }
    prm.chunk_peaks = p->chunk_peaks.p;
    prm.norm_clamp = (p->framing == TMT_FRAMING_WHOLEFILE) ? 1 : 0;
    prm.pipeline_f64 = pipeline_f64 ? 1 : 0;
    prm.post_gain = post_gain;
    cudaStream_t where = st;
    if (beside) {
        if (!p->ev_fork) {
            std::lock_guard<std::mutex> lock(e->ev_mu);
            if (!e->side) CUDA_TRY(cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking));
            for (cudaEvent_t* ev : {&p->ev_fork, &p->ev_join}) {
                if (!e->ev_pool.empty()) { *ev = e->ev_pool.back(); e->ev_pool.pop_back(); }
                else CUDA_TRY(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
            }
        }
        CUDA_TRY(cudaEventRecord(p->ev_fork, st));
        CUDA_TRY(cudaStreamWaitEvent(e->side, p->ev_fork, 0));
        where = e->side;
    }
    edge_kernel<<<p->n_edges, 256, kEdgeKernelSmemBytes, where>>>(prm);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    if (beside) CUDA_TRY(cudaEventRecord(p->ev_join, e->side));
    return TMT_OK;
}
static int join_edges(tmt_plan* p, cudaStream_t st) {
    CUDA_TRY(cudaStreamWaitEvent(st, p->ev_join, 0));
    return TMT_OK;
}

int tmt_plan_edge_frames(tmt_plan* p, float post_gain, const float* in_scale, const float* out_scale, int pipeline_f64,
                         void* stream) {
    if (!p) return fail(TMT_ERR_INVALID, "plan is NULL");
    tmt_engine* e = p->e;
    if (!e->have_win || e->n_rows == 0) return fail(TMT_ERR_INVALID, "engine window / gain rows not set");
    if (p->n_edges == 0) return TMT_OK;
    CUDA_TRY(cudaSetDevice(e->device));
    return launch_edges(p, post_gain, in_scale, out_scale, pipeline_f64, reinterpret_cast<cudaStream_t>(stream), false);
}

int tmt_plan_stft_with_edges(tmt_plan* p, float post_gain, const float* in_scale, const float* out_scale, int pipeline_f64, void* stream) {
    if (!p) return fail(TMT_ERR_INVALID, "plan is NULL");
    tmt_engine* e = p->e;
    if (!e->have_win || e->n_rows == 0) return fail(TMT_ERR_INVALID, "engine window / gain rows not set");
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = tmt_plan_clear_peaks(p, stream);
    if (rc) return rc;
    const bool beside = p->n_edges > 0 && p->n_units > 0 && !getenv("TMT_EDGES_SERIAL");
    if (p->n_edges) { rc = launch_edges(p, post_gain, in_scale, out_scale, pipeline_f64, st, beside); if (rc) return rc; }
    if (p->n_units) { rc = launch_stft(p, post_gain, 0.f, st); if (rc) return rc; }
    if (beside) return join_edges(p, st);
    return TMT_OK;
}

int tmt_plan_limiter(tmt_plan* p, float limit, void* stream) {
    if (!p) return fail(TMT_ERR_INVALID, "plan is NULL");
    if (p->total_chunks == 0 || p->max_chunk_len == 0) return TMT_OK;
    CUDA_TRY(cudaSetDevice(p->e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long per_cta = 256LL * 16;
    int gx = (int)std::min<long long>((p->max_chunk_len + per_cta - 1) / per_cta, 8LL * p->e->n_sms);
    gx = std::max(gx, 1);
    limiter_kernel<<<dim3(gx, p->total_chunks), 256, 0, st>>>(p->tracks.p, p->chunks.p, p->chunk_peaks.p, limit, 0);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

static int run_streaming_tail(tmt_plan* p, double m_on, double m_off, int run_frames, int xfade_frames, float post_gain, float limit,
                              void* stream);

int tmt_plan_run_streaming(tmt_plan* p, double m_on, double m_off, int run_frames, int xfade_frames, float post_gain,
                           float limit, void* stream) {
    if (!p) return fail(TMT_ERR_INVALID, "plan is NULL");
    int rc = tmt_plan_levels(p, 0, nullptr, stream);
    if (rc) return rc;
    return run_streaming_tail(p, m_on, m_off, run_frames, xfade_frames, post_gain, limit, stream);
}

int tmt_plan_pcm_levels(tmt_plan* p, const void* pcm, int64_t track_stride_bytes, int format, void* stream) {
    if (!p || !pcm) return fail(TMT_ERR_INVALID, "bad arguments");
    if (format != TMT_PCM_S16 && format != TMT_PCM_S24) return fail(TMT_ERR_INVALID, "unknown PCM format %d", format);
    if (p->n_tracks == 0) return TMT_OK;
    for (const HostTrack& h : p->ht)
        if (h.d.in_origin != 0 || h.d.in_len != h.d.total || h.hb_lo != 0 || (h.n_frames > 0 && h.hb_hi != h.n_frames + 1))
            return fail(TMT_ERR_UNSUPPORTED, "tmt_plan_pcm_levels needs whole-track plans (input window = the file, all hop blocks)");
    CUDA_TRY(cudaSetDevice(p->e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const dim3 grid(ceil_div(p->max_frames + 1, kLevelWarps), p->n_tracks);
    if (format == TMT_PCM_S16)
        pcm_levels_kernel<0><<<grid, kLevelWarps * 32, 0, st>>>(p->tracks.p, reinterpret_cast<const unsigned char*>(pcm), track_stride_bytes,
                                                              reinterpret_cast<float*>(p->hsum.p));
    else
        pcm_levels_kernel<1><<<grid, kLevelWarps * 32, 0, st>>>(p->tracks.p, reinterpret_cast<const unsigned char*>(pcm), track_stride_bytes,
                                                              reinterpret_cast<float*>(p->hsum.p));
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

int tmt_plan_run_streaming_pcm(tmt_plan* p, const void* pcm, int64_t track_stride_bytes, int format, double m_on, double m_off,
                               int run_frames, int xfade_frames, float post_gain, float limit, void* stream) {
    int rc = tmt_plan_pcm_levels(p, pcm, track_stride_bytes, format, stream);      // conversion + hop-block sums in one pass
    if (rc) return rc;
    rc = tmt_plan_levels(p, TMT_LEVELS_MEANSQ_ONLY, nullptr, stream);
    if (rc) return rc;
    return run_streaming_tail(p, m_on, m_off, run_frames, xfade_frames, post_gain, limit, stream);
}

static int run_streaming_tail(tmt_plan* p, double m_on, double m_off, int run_frames, int xfade_frames, float post_gain, float limit,
                              void* stream) {
    int rc;
    std::vector<double> on((size_t)std::max(p->n_tracks, 1), m_on), off((size_t)std::max(p->n_tracks, 1), m_off);
    rc = tmt_plan_gate(p, TMT_GATE_UPDELAY, TMT_ARR_MEANSQ_F32, on.data(), off.data(), run_frames, xfade_frames, 0, 0, stream);
    if (rc) return rc;
    if (!stft_fused_limiter(p)) {
        // small job: separate limiter pass anyway, so the edge frames can run beside the STFT kernel
        if (!(limit > 0.f)) return fail(TMT_ERR_INVALID, "limit must be positive");
        rc = tmt_plan_stft_with_edges(p, post_gain, nullptr, nullptr, 0, stream);
        if (rc) return rc;
        return tmt_plan_limiter(p, limit, stream);
    }
    rc = tmt_plan_clear_peaks(p, stream);
    if (rc) return rc;
    rc = tmt_plan_edge_frames(p, post_gain, nullptr, nullptr, 0, stream);      // edge blocks first: their samples and peaks
    if (rc) return rc;                                                         // must be in place when a chunk is finished
    return tmt_plan_stft_limited(p, post_gain, limit, stream);
}

int tmt_pcm_to_float(const void* pcm, int format, int64_t n_values, float* out, void* stream) {
    if (n_values < 0 || (n_values > 0 && (!pcm || !out))) return fail(TMT_ERR_INVALID, "bad arguments");
    if (n_values == 0) return TMT_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool aligned = ((reinterpret_cast<uintptr_t>(pcm) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    const int grid = 148 * 8;
    if (format == TMT_PCM_S16) {
        const long long n8 = aligned ? n_values / 8 : 0;
        s16_to_float_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const int4*>(pcm), reinterpret_cast<float4*>(out), n8,
                                                  reinterpret_cast<const short*>(pcm), out, n_values);
    } else if (format == TMT_PCM_S24) {
        const long long n4 = aligned ? n_values / 4 : 0;
        s24_to_float_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned*>(pcm), reinterpret_cast<float4*>(out), n4,
                                                  reinterpret_cast<const unsigned char*>(pcm), out, n_values);
    } else {
        return fail(TMT_ERR_INVALID, "unknown PCM format %d", format);
    }
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

int tmt_requantise_scale(float* y, int64_t n_values, float scale, void* stream) {
    if (n_values < 0 || (n_values > 0 && !y)) return fail(TMT_ERR_INVALID, "bad arguments");
    if (n_values == 0) return TMT_OK;
    requantise_scale_kernel<<<148 * 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(y, n_values, scale);
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

int tmt_float_to_pcm(const float* in, int format, int64_t n_values, void* pcm, void* stream) {
    if (n_values < 0 || (n_values > 0 && (!pcm || !in))) return fail(TMT_ERR_INVALID, "bad arguments");
    if (n_values == 0) return TMT_OK;
    if (format != TMT_PCM_S24) return fail(TMT_ERR_UNSUPPORTED, "only PCM_24 output is implemented (the reference writes PCM_24)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool aligned = ((reinterpret_cast<uintptr_t>(pcm) | reinterpret_cast<uintptr_t>(in)) & 15u) == 0;
    const long long n4 = aligned ? n_values / 4 : 0;
    float_to_s24_kernel<<<148 * 8, 256, 0, st>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<unsigned*>(pcm), n4, in,
                                                 reinterpret_cast<unsigned char*>(pcm), n_values);
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

int tmt_cond_spectrum(tmt_engine* e, const void* x, const void* y, int64_t total, const int32_t* frames, int n_frames,
                      int anchor_bin_lo, int anchor_bin_hi, float* median_out, void* stream) {
    if (!e || !frames || !median_out || n_frames < 0 || total < 0) return fail(TMT_ERR_INVALID, "bad arguments");
    if (!e->have_win) return fail(TMT_ERR_INVALID, "engine window not set");
    if (n_frames == 0) return TMT_OK;
    if (!x || !y) return fail(TMT_ERR_INVALID, "NULL audio buffer");
    if (anchor_bin_hi >= anchor_bin_lo && (anchor_bin_lo < 0 || anchor_bin_hi >= kBins)) return fail(TMT_ERR_INVALID, "anchor band outside [0, %d]", kBins - 1);
    for (int i = 0; i < n_frames; ++i)
        if (frames[i] < 0 || (long long)frames[i] * kHop + kNfft > total)
            return fail(TMT_ERR_INVALID, "frame %d (index %d) does not lie inside the file", i, frames[i]);
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaFuncSetAttribute(spectrum_ratio_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSpecSmemBytes));
    DevBuf<int> d_frames;
    DevBuf<float> d_ratio, d_med;
    CUDA_TRY(d_frames.alloc((size_t)n_frames));
    CUDA_TRY(d_ratio.alloc((size_t)kBins * (size_t)n_frames));
    CUDA_TRY(d_med.alloc(kBins));
    CUDA_TRY(cudaMemcpyAsync(d_frames.p, frames, sizeof(int) * (size_t)n_frames, cudaMemcpyHostToDevice, st));
    spectrum_ratio_kernel<<<n_frames, 256, kSpecSmemBytes, st>>>(reinterpret_cast<const float2*>(x), reinterpret_cast<const float2*>(y),
                                                                 d_frames.p, (long long)n_frames, e->win.p,
                                                                 anchor_bin_lo, anchor_bin_hi, d_ratio.p);
    CUDA_TRY(cudaGetLastError());
    column_median_kernel<<<kBins, 256, 0, st>>>(d_ratio.p, n_frames, (long long)n_frames, d_med.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(median_out, d_med.p, sizeof(float) * kBins, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return TMT_OK;
}

int tmt_calib_envelope_decimate(tmt_engine* e, const void* x, int64_t n_in, const float* h, int len_h, int up, int down,
                                int64_t n_pre_remove, int64_t n_out, float* out, void* stream) {
    if (!e || !h || len_h <= 0 || up < 1 || down < 1 || n_in < 0 || n_out < 0 || n_pre_remove < 0) return fail(TMT_ERR_INVALID, "bad arguments");
    if (n_out == 0) return TMT_OK;
    if (!x || !out) return fail(TMT_ERR_INVALID, "NULL buffer");
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DevBuf<float> d_h;
    DevBuf<double> d_part;
    CUDA_TRY(d_h.alloc((size_t)len_h));
    CUDA_TRY(d_part.alloc(kSumBlocks));
    CUDA_TRY(cudaMemcpyAsync(d_h.p, h, sizeof(float) * (size_t)len_h, cudaMemcpyHostToDevice, st));
    const int grid = (int)std::min<long long>((n_out + 255) / 256, 148LL * 16);
    calib_decimate_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float2*>(x), n_in, d_h.p, len_h, up, down, n_pre_remove, n_out, out);
    CUDA_TRY(cudaGetLastError());
    // x - np.mean(x): block partial sums in double, folded on the host in a fixed order
    calib_partial_sum_kernel<<<kSumBlocks, 256, 0, st>>>(out, n_out, d_part.p);
    CUDA_TRY(cudaGetLastError());
    std::vector<double> part(kSumBlocks);
    CUDA_TRY(cudaMemcpyAsync(part.data(), d_part.p, sizeof(double) * kSumBlocks, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    double sum = 0.0;
    for (double v : part) sum += v;
    calib_subtract_kernel<<<grid, 256, 0, st>>>(out, n_out, (float)(sum / (double)n_out));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(st));
    return TMT_OK;
}

int tmt_calib_xcorr_valid(tmt_engine* e, const float* a, int64_t na, const float* b, int64_t nb, float* corr, void* stream) {
    if (!e || na < 0 || nb <= 0 || nb > na || nb > 0x7fffffffLL) return fail(TMT_ERR_INVALID, "need 0 < nb <= na");
    if (!a || !b || !corr) return fail(TMT_ERR_INVALID, "NULL buffer");
    CUDA_TRY(cudaSetDevice(e->device));
    const long long n_lags = na - nb + 1;
    calib_xcorr_kernel<<<(unsigned)((n_lags + kXcLags - 1) / kXcLags), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, na, b, (int)nb, corr, n_lags);
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

int tmt_calib_band_energies(tmt_engine* e, const void* x, int64_t total, int n_frames, int lo0, int lo1, int hi0, int hi1,
                            float* e_lo, float* e_hi, void* stream) {
    if (!e || n_frames < 0 || total < 0) return fail(TMT_ERR_INVALID, "bad arguments");
    if (!e->have_win) return fail(TMT_ERR_INVALID, "engine window not set");
    if (n_frames == 0) return TMT_OK;
    if (!x || !e_lo || !e_hi) return fail(TMT_ERR_INVALID, "NULL buffer");
    if ((long long)(n_frames - 1) * kHop + kNfft > total) return fail(TMT_ERR_INVALID, "%d frames do not fit into %lld sample-frames", n_frames, (long long)total);
    if (lo0 < 0 || lo1 > kBins || hi0 < 0 || hi1 > kBins) return fail(TMT_ERR_INVALID, "band outside [0, %d]", kBins);
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaFuncSetAttribute(calib_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSpecSmemBytes));
    DevBuf<float> d_lo, d_hi;
    CUDA_TRY(d_lo.alloc((size_t)n_frames));
    CUDA_TRY(d_hi.alloc((size_t)n_frames));
    calib_band_kernel<<<n_frames, 256, kSpecSmemBytes, st>>>(reinterpret_cast<const float2*>(x), e->win.p, lo0, lo1, hi0, hi1, d_lo.p, d_hi.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(e_lo, d_lo.p, sizeof(float) * (size_t)n_frames, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(e_hi, d_hi.p, sizeof(float) * (size_t)n_frames, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return TMT_OK;
}

int tmt_calib_gate_grid(tmt_engine* e, const float* level, const int64_t* start, const uint8_t* want, int n, const float* on,
                        const float* off, const int64_t* delay, int n_combos, int32_t* mismatches, int32_t* switches, uint8_t* states,
                        void* stream) {
    if (!e || n < 0 || n_combos < 0) return fail(TMT_ERR_INVALID, "bad arguments");
    if (n_combos == 0) return TMT_OK;
    if ((n > 0 && (!level || !start || !want)) || !on || !off || !delay || !mismatches || !switches) return fail(TMT_ERR_INVALID, "NULL buffer");
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DevBuf<float> d_level, d_on, d_off;
    DevBuf<long long> d_start, d_delay;
    DevBuf<unsigned char> d_want, d_states;
    DevBuf<int> d_mis, d_sw;
    if (states && n > 0) CUDA_TRY(d_states.alloc((size_t)n * (size_t)n_combos));
    CUDA_TRY(d_level.alloc((size_t)std::max(n, 1))); CUDA_TRY(d_start.alloc((size_t)std::max(n, 1))); CUDA_TRY(d_want.alloc((size_t)std::max(n, 1)));
    CUDA_TRY(d_on.alloc((size_t)n_combos)); CUDA_TRY(d_off.alloc((size_t)n_combos)); CUDA_TRY(d_delay.alloc((size_t)n_combos));
    CUDA_TRY(d_mis.alloc((size_t)n_combos)); CUDA_TRY(d_sw.alloc((size_t)n_combos));
    if (n > 0) {
        CUDA_TRY(cudaMemcpyAsync(d_level.p, level, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_start.p, start, sizeof(long long) * (size_t)n, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_want.p, want, (size_t)n, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemcpyAsync(d_on.p, on, sizeof(float) * (size_t)n_combos, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_off.p, off, sizeof(float) * (size_t)n_combos, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_delay.p, delay, sizeof(long long) * (size_t)n_combos, cudaMemcpyHostToDevice, st));
    calib_gate_grid_kernel<<<(n_combos + 127) / 128, 128, 0, st>>>(d_level.p, d_start.p, d_want.p, n, d_on.p, d_off.p, d_delay.p, n_combos, d_mis.p, d_sw.p,
                                                                   (states && n > 0) ? d_states.p : nullptr);
    CUDA_TRY(cudaGetLastError());
    if (states && n > 0) CUDA_TRY(cudaMemcpyAsync(states, d_states.p, (size_t)n * (size_t)n_combos, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(mismatches, d_mis.p, sizeof(int) * (size_t)n_combos, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(switches, d_sw.p, sizeof(int) * (size_t)n_combos, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return TMT_OK;
}

// ---- general FFT sizes (EXPERIMENTAL) ------------------------------------------------------------
static int gen_check(tmt_engine* e, int n_fft, int hop, int n_frames, int* log2n) {
    if (!e) return fail(TMT_ERR_INVALID, "engine is NULL");
    int l = 0;
    while ((1 << l) < n_fft) ++l;
    if ((1 << l) != n_fft || n_fft < 128 || n_fft > 8192) return fail(TMT_ERR_UNSUPPORTED, "n_fft must be a power of two in [128, 8192], got %d", n_fft);
    if (hop < 1 || hop > n_fft) return fail(TMT_ERR_UNSUPPORTED, "hop must lie in [1, n_fft], got %d", hop);
    if (n_frames < 0) return fail(TMT_ERR_INVALID, "negative frame count");
    *log2n = l;
    return TMT_OK;
}

int tmt_generic_meansq(tmt_engine* e, const void* x, int64_t total, int64_t first_start, int n_fft, int hop, int n_frames, int use_f64,
                       float in_scale, int mono_file, void* meansq_out, void* stream) {
    int l2;
    int rc = gen_check(e, n_fft, hop, n_frames, &l2);
    if (rc) return rc;
    if (n_frames == 0) return TMT_OK;
    if (!x || !meansq_out || total < 0) return fail(TMT_ERR_INVALID, "bad buffers");
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = (n_frames + 127) / 128;
    if (use_f64)
        gen_meansq_kernel<double><<<grid, 128, 0, st>>>(reinterpret_cast<const float2*>(x), total, first_start, n_fft, hop, n_frames, in_scale, mono_file, reinterpret_cast<double*>(meansq_out));
    else
        gen_meansq_kernel<float><<<grid, 128, 0, st>>>(reinterpret_cast<const float2*>(x), total, first_start, n_fft, hop, n_frames, in_scale, mono_file, reinterpret_cast<float*>(meansq_out));
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

int tmt_generic_frames(tmt_engine* e, const void* x, int64_t total, int64_t first_start, int n_fft, int hop, int n_frames, const float* win,
                       const float* gains, const uint16_t* rows, float in_scale, int flavour, void* frames_out, void* stream) {
    GenFrameParams prm;
    int rc = gen_check(e, n_fft, hop, n_frames, &prm.log2n);
    if (rc) return rc;
    if (n_frames == 0) return TMT_OK;
    if (!x || !win || !gains || !rows || !frames_out || total < 0 || flavour < kGenStreaming || flavour > kGenAdaptiveF64) return fail(TMT_ERR_INVALID, "bad arguments");
    CUDA_TRY(cudaSetDevice(e->device));
    prm.x = reinterpret_cast<const float2*>(x); prm.total = total; prm.first_start = first_start; prm.n_fft = n_fft; prm.hop = hop;
    prm.n_frames = n_frames; prm.win = win; prm.gains = gains; prm.rows = rows; prm.in_scale = in_scale; prm.flavour = flavour; prm.frames = frames_out;
    const int smem = n_fft * (int)sizeof(double2);
    CUDA_TRY(cudaFuncSetAttribute(gen_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    gen_frame_kernel<<<n_frames, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(prm);
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

int tmt_generic_overlap_add(tmt_engine* e, const void* frames, int flavour, int64_t total, int64_t first_start, int n_fft, int hop, int n_frames,
                            const float* win, float post_scale, void* y_out, void* stream) {
    int l2;
    int rc = gen_check(e, n_fft, hop, n_frames, &l2);
    if (rc) return rc;
    if (total <= 0) return TMT_OK;
    if ((n_frames > 0 && !frames) || !win || !y_out || flavour < kGenStreaming || flavour > kGenAdaptiveF64) return fail(TMT_ERR_INVALID, "bad arguments");
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = (int)std::min<long long>((total + 255) / 256, 16LL * e->n_sms);
    if (flavour == kGenAdaptiveF64)
        gen_ola_kernel<double, double2, double2><<<grid, 256, 0, st>>>(reinterpret_cast<const double2*>(frames), win, total, first_start, n_fft, hop, n_frames, 1, post_scale, reinterpret_cast<double2*>(y_out));
    else
        gen_ola_kernel<float, float2, float2><<<grid, 256, 0, st>>>(reinterpret_cast<const float2*>(frames), win, total, first_start, n_fft, hop, n_frames, flavour == kGenAdaptiveF32, post_scale, reinterpret_cast<float2*>(y_out));
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

int tmt_generic_limit(tmt_engine* e, void* y, int use_f64, const int64_t* bounds, int n_chunks, double limit, void* peaks_out, void* stream) {
    if (!e || n_chunks < 0) return fail(TMT_ERR_INVALID, "bad arguments");
    if (n_chunks == 0) return TMT_OK;
    if (!y || !bounds || !peaks_out) return fail(TMT_ERR_INVALID, "NULL buffer");
    CUDA_TRY(cudaSetDevice(e->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const dim3 grid(4 * e->n_sms, n_chunks);
    const long long* b = reinterpret_cast<const long long*>(bounds);
    CUDA_TRY(cudaMemsetAsync(peaks_out, 0, (use_f64 ? sizeof(double) : sizeof(float)) * (size_t)n_chunks, st));
    if (use_f64) {
        gen_peak_kernel<double2, double><<<grid, 256, 0, st>>>(reinterpret_cast<const double2*>(y), b, reinterpret_cast<double*>(peaks_out));
        gen_limit_kernel<double2, double><<<grid, 256, 0, st>>>(reinterpret_cast<double2*>(y), b, reinterpret_cast<const double*>(peaks_out), limit);
    } else {
        gen_peak_kernel<float2, float><<<grid, 256, 0, st>>>(reinterpret_cast<const float2*>(y), b, reinterpret_cast<float*>(peaks_out));
        gen_limit_kernel<float2, float><<<grid, 256, 0, st>>>(reinterpret_cast<float2*>(y), b, reinterpret_cast<const float*>(peaks_out), (float)limit);
    }
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

int tmt_generic_to_float(tmt_engine* e, const void* y_f64, int64_t n, void* out_f32, void* stream) {
    if (!e || n < 0) return fail(TMT_ERR_INVALID, "bad arguments");
    if (n == 0) return TMT_OK;
    if (!y_f64 || !out_f32) return fail(TMT_ERR_INVALID, "NULL buffer");
    CUDA_TRY(cudaSetDevice(e->device));
    const int grid = (int)std::min<long long>((n + 255) / 256, 16LL * e->n_sms);
    gen_to_float_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const double2*>(y_f64), reinterpret_cast<float2*>(out_f32), n);
    CUDA_TRY(cudaGetLastError());
    return TMT_OK;
}

}  // extern "C"
