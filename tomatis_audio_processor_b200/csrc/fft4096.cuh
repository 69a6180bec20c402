// 4096-point complex FFT building blocks for one 256-thread CTA (sm_100a), fp32.
//
// The reference does, per STFT frame and per channel, rfft -> real gain -> irfft
// (/root/reference/src/process_tomatis.py:394-398).  Here the stereo pair is packed as one
// complex signal z = L + iR (the gain is real and symmetric, so IFFT(g*FFT(z)) = yL + i*yR),
// and N = 4096 = 16*16*16 is done as three register-resident radix-16 stages.  Index split
//     n = 256*n1 + 16*n2 + n3          k = k1 + 16*k2 + 256*k3
// forward:  A (threads (n2,n3), DFT over n1, twiddle W4096^((16*n2+n3)*k1))  -> exchange E1 (shared memory)
//           B (threads (k1,n3), DFT over n2, twiddle W256^(n3*k2)) -> exchange E2 (TENSOR MEMORY, see below)
//           C (threads (k1,k2), DFT over n3)  -> X[k1+16*k2+256*k3] in registers
// inverse:  exactly the mirror (C' B' A') with conjugate twiddles applied on stage inputs, so
// the spectrum never leaves registers between forward and inverse and the time-domain result
// lands in the same thread/register layout the input was loaded in (n = 256*j + t).
//
// E2 swaps the 4-bit register index k2 with the 4-bit thread index n3 inside groups of 16 threads.  All 16 threads of such
// a group sit in one warp (lane = 2*n3 + (k1 & 1)), so the exchange can use the warp's own 32 lanes of tensor memory instead
// of shared memory: a tcgen05.st in the .32x32b shape (lane i, column c <-> thread i, register c) followed by a tcgen05.ld in
// the .16x256b shape (lane r, column c <-> thread 4*(r%8) + (c%8)/2, register 4*(c/8) + 2*(r/8) + c%2) moves two thread bits
// into the register index and two register bits into the thread index.  Two such round trips make the 4-bit exchange; the
// radix-16 butterfly of stage C is split into its two radix-4 layers and one round trip sits on either side of the first
// layer (n3 = 4a + b: bits a arrive with the first trip, bits b with the second), so every round trip is adjacent to
// arithmetic and costs no register moves.  The inverse path runs the two trips backwards (st .16x256b, ld .32x32b).
// Measured on B200 (tools/mb_tmem_xpose.cu): 41 SM-cycles per warp-level exchange on the tensor-memory pipe against 64
// wavefront-cycles on the shared-memory pipe that co-limits this kernel; same latency (~200 cycles).
//
// Everything here is __host__ __device__ so tests/ can run the same code on the CPU
// (csrc/host_emul.cu) -- there is no GPU in the build container.
#pragma once
#include <cuda_runtime.h>

#define TMT_HD __host__ __device__ __forceinline__

namespace tmt {

// The library is compiled once per supported frame size: TMT_NFFT = 4096 (default; hop 2048) or 2048 (hop 1024, "pair mode":
// two consecutive 2048-point frames ride one 4096-wide pass of the same machinery as the even and odd samples of the packed
// vector, see the pair-mode section below).  Everything that is not the FFT derives its geometry from kNfft / kHop.
#ifndef TMT_NFFT
#define TMT_NFFT 4096
#endif
static_assert(TMT_NFFT == 4096 || TMT_NFFT == 2048, "fused path: n_fft 4096 (hop 2048) or 2048 (hop 1024)");
constexpr int kNfft = TMT_NFFT;
constexpr int kHop = TMT_NFFT / 2;
constexpr bool kPair = (TMT_NFFT == 2048);
constexpr int kPassLen = 4096;      // complex points of one pass of the FFT machinery (one 4096-frame or two 2048-frames)
constexpr int kThreads = 256;       // one pass per CTA, 16 points per thread
#ifdef TMT_E1_PADDED_ROWS             // the earlier layout: 16 padded rows, 64-bit accesses on both sides
constexpr int kE1Row = 264;         // E1 row stride in float2 (256 + 8): the B side reads rows k1 and k1 + 1 from the two lane
                                    // parities of a warp, the 64-byte skew keeps every half warp on 32 distinct banks
constexpr int kE1Float2 = 16 * kE1Row;       // the E1 exchange buffer (33 792 B)
#else
// Rows interleaved in pairs: element (k1, col) lives at (k1 >> 1) * 512 + 2 * col + (k1 & 1).  The A side (thread t = col, rows
// j = 0..15) then moves rows 2jp and 2jp + 1 with ONE 128-bit access (a warp covers 512 contiguous bytes), the B side (a warp
// owns exactly one row pair, lane = 2 * n3 + (k1 & 1)) reads 32 consecutive float2 per access: no padding, no bank conflicts,
// 16 shared-memory instructions fewer per frame.
constexpr int kE1Row = 256;
constexpr int kE1Float2 = 16 * kE1Row;       // the E1 exchange buffer (32 768 B)
#endif

// Complex arithmetic on float2.  On the device every operation is a packed FP32x2 instruction
// (FADD2 / FMUL2 / FFMA2, new on sm_100): re/im live in one 64-bit register pair, and the swap,
// per-half sign and scalar broadcast these formulas need are free operand modifiers in SASS
// (R.F32x2.LO_HI.NP, R.F32), so a complex add is ONE instruction and a complex multiply TWO.
// The host versions (CPU emulation in tests) are the plain scalar formulas.
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
TMT_HD float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
TMT_HD float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
TMT_HD float2 cscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
// a + (-i)*b = (a.x + b.y, a.y - b.x)
TMT_HD float2 cadd_mi(float2 a, float2 b) { return __ffma2_rn(make_float2(b.y, b.x), make_float2(1.f, -1.f), a); }
// a + (+i)*b = (a.x - b.y, a.y + b.x)
TMT_HD float2 cadd_pi(float2 a, float2 b) { return __ffma2_rn(make_float2(b.y, b.x), make_float2(-1.f, 1.f), a); }
TMT_HD float2 cmul(float2 a, float2 w) {
    return __ffma2_rn(make_float2(a.y, a.x), make_float2(-w.y, w.y), __fmul2_rn(a, make_float2(w.x, w.x)));
}
// a * conj(w)
TMT_HD float2 cmulc(float2 a, float2 w) {
    return __ffma2_rn(make_float2(a.y, a.x), make_float2(w.y, -w.y), __fmul2_rn(a, make_float2(w.x, w.x)));
}
// acc + a * w  /  acc + a * conj(w): two FFMA2, no separate add
TMT_HD float2 cfma(float2 acc, float2 a, float2 w) {
    return __ffma2_rn(make_float2(a.y, a.x), make_float2(-w.y, w.y), __ffma2_rn(a, make_float2(w.x, w.x), acc));
}
TMT_HD float2 cfmac(float2 acc, float2 a, float2 w) {
    return __ffma2_rn(make_float2(a.y, a.x), make_float2(w.y, -w.y), __ffma2_rn(a, make_float2(w.x, w.x), acc));
}
// 2*p - t  (the other output of a twiddled radix-2 butterfly: p - w*b = 2p - (p + w*b))
TMT_HD float2 twice_minus(float2 p, float2 t) { return __ffma2_rn(p, make_float2(2.f, 2.f), make_float2(-t.x, -t.y)); }
// a * s + p  /  a * s - p  with a real scale s
TMT_HD float2 cfms(float2 a, float s, float2 p) { return __ffma2_rn(a, make_float2(s, s), p); }
TMT_HD float2 cfmsn(float2 a, float s, float2 p) { return __ffma2_rn(a, make_float2(s, s), make_float2(-p.x, -p.y)); }
#else
TMT_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
TMT_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
TMT_HD float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
TMT_HD float2 cadd_mi(float2 a, float2 b) { return make_float2(a.x + b.y, a.y - b.x); }
TMT_HD float2 cadd_pi(float2 a, float2 b) { return make_float2(a.x - b.y, a.y + b.x); }
TMT_HD float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
TMT_HD float2 cmulc(float2 a, float2 w) { return make_float2(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y); }
TMT_HD float2 cfma(float2 acc, float2 a, float2 w) { return make_float2(acc.x + a.x * w.x - a.y * w.y, acc.y + a.x * w.y + a.y * w.x); }
TMT_HD float2 cfmac(float2 acc, float2 a, float2 w) { return make_float2(acc.x + a.x * w.x + a.y * w.y, acc.y + a.y * w.x - a.x * w.y); }
TMT_HD float2 twice_minus(float2 p, float2 t) { return make_float2(2.f * p.x - t.x, 2.f * p.y - t.y); }
TMT_HD float2 cfms(float2 a, float s, float2 p) { return make_float2(a.x * s + p.x, a.y * s + p.y); }
TMT_HD float2 cfmsn(float2 a, float s, float2 p) { return make_float2(a.x * s - p.x, a.y * s - p.y); }
#endif

// 4-point DFT in place: (x0..x3) -> (X0..X3), forward W4 = -i, inverse W4 = +i
template <bool INV>
TMT_HD void radix4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = csub(x1, x3);
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    if (INV) {
        x1 = cadd_pi(t1, t3);
        x3 = cadd_mi(t1, t3);
    } else {
        x1 = cadd_mi(t1, t3);
        x3 = cadd_pi(t1, t3);
    }
}

// 4-point DFT of (s0*x0, s1*x1, s2*x2, s3*x3) with REAL scales (window taps, spectral gains) fused into the first butterfly level:
// a*s0 +- c*s2 costs one multiply and two fused multiply-adds instead of two multiplies, an add and a subtract (10 packed
// instructions instead of 12; one rounding fewer per sum).
template <bool INV>
TMT_HD void radix4_scaled(float2& x0, float2& x1, float2& x2, float2& x3, float s0, float s1, float s2, float s3) {
    const float2 p2 = cscale(x2, s2), p3 = cscale(x3, s3);
    const float2 t0 = cfms(x0, s0, p2), t1 = cfmsn(x0, s0, p2), t2 = cfms(x1, s1, p3), t3 = cfmsn(x1, s1, p3);
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    if (INV) {
        x1 = cadd_pi(t1, t3);
        x3 = cadd_mi(t1, t3);
    } else {
        x1 = cadd_mi(t1, t3);
        x3 = cadd_pi(t1, t3);
    }
}

// 4-point DFT of (x0, w1*x1, w2*x2, w3*x3) (CONJ: conj(w)), twiddles fused into the first butterfly level:
// p + w*b costs two FFMA2 and p - w*b = 2p - (p + w*b) one more, instead of multiply (2) + add + subtract.
// TW0: x0 is also multiplied by w0.  12 packed instructions (14 with TW0) instead of 14 (16).
template <bool INV, bool CONJ, bool TW0>
TMT_HD void radix4_tw(float2& x0, float2& x1, float2& x2, float2& x3, float2 w0, float2 w1, float2 w2, float2 w3) {
    const float2 p0 = TW0 ? (CONJ ? cmulc(x0, w0) : cmul(x0, w0)) : x0;
    const float2 t0 = CONJ ? cfmac(p0, x2, w2) : cfma(p0, x2, w2);
    const float2 t1 = twice_minus(p0, t0);
    const float2 p1 = CONJ ? cmulc(x1, w1) : cmul(x1, w1);
    const float2 t2 = CONJ ? cfmac(p1, x3, w3) : cfma(p1, x3, w3);
    const float2 t3 = twice_minus(p1, t2);
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    if (INV) {
        x1 = cadd_pi(t1, t3);
        x3 = cadd_mi(t1, t3);
    } else {
        x1 = cadd_mi(t1, t3);
        x3 = cadd_pi(t1, t3);
    }
}

// W16^M as a value (forward e^{-2*pi*i*M/16}; the inverse transform uses its conjugate)
template <int M>
TMT_HD float2 w16() {
    constexpr float kC1 = 0.92387953251128673848f, kS1 = 0.38268343236508978178f, kH = 0.70710678118654752440f;
    static_assert(M == 1 || M == 2 || M == 3 || M == 6 || M == 9, "unsupported W16 power");
    return (M == 1) ? make_float2(kC1, -kS1) : (M == 2) ? make_float2(kH, -kH) : (M == 3) ? make_float2(kS1, -kC1)
         : (M == 6) ? make_float2(-kH, -kH) : make_float2(-kC1, kS1);
}

// second radix-4 layer of the 16-point DFT: over b for each c, twiddles W16^(b*c) fused in, then the 4x4 transpose of the
// register names (free after unrolling); V[c + 4d] ends up in v[c + 4d]
template <bool INV>
TMT_HD void dft16_layer2(float2 (&v)[16]) {
    radix4<INV>(v[0], v[1], v[2], v[3]);                                                           // c = 0: no twiddles
    radix4_tw<INV, INV, false>(v[4], v[5], v[6], v[7], make_float2(1.f, 0.f), w16<1>(), w16<2>(), w16<3>());       // c = 1
    {                                                                                              // c = 2: W^0, W^2, W^4 = -+i, W^6
        const float2 t0 = INV ? cadd_pi(v[8], v[10]) : cadd_mi(v[8], v[10]);
        const float2 t1 = INV ? cadd_mi(v[8], v[10]) : cadd_pi(v[8], v[10]);
        const float2 p1 = INV ? cmulc(v[9], w16<2>()) : cmul(v[9], w16<2>());
        const float2 t2 = INV ? cfmac(p1, v[11], w16<6>()) : cfma(p1, v[11], w16<6>());
        const float2 t3 = twice_minus(p1, t2);
        v[8] = cadd(t0, t2);
        v[10] = csub(t0, t2);
        v[9] = INV ? cadd_pi(t1, t3) : cadd_mi(t1, t3);
        v[11] = INV ? cadd_mi(t1, t3) : cadd_pi(t1, t3);
    }
    radix4_tw<INV, INV, false>(v[12], v[13], v[14], v[15], make_float2(1.f, 0.f), w16<3>(), w16<6>(), w16<9>());   // c = 3
    float2 t;
    t = v[1]; v[1] = v[4]; v[4] = t;
    t = v[2]; v[2] = v[8]; v[8] = t;
    t = v[3]; v[3] = v[12]; v[12] = t;
    t = v[6]; v[6] = v[9]; v[9] = t;
    t = v[7]; v[7] = v[13]; v[13] = t;
    t = v[11]; v[11] = v[14]; v[14] = t;
}

// 16-point DFT, natural order in and out, fully in registers:
//   n = 4a + b, k = c + 4d:  V[c+4d] = sum_b W4^(bd) * W16^(bc) * sum_a W4^(ac) v[4a+b]
template <bool INV>
TMT_HD void dft16(float2 (&v)[16]) {
    // first layer: radix-4 over a for each b; u_b[c] is left in v[4c + b]
    radix4<INV>(v[0], v[4], v[8], v[12]);
    radix4<INV>(v[1], v[5], v[9], v[13]);
    radix4<INV>(v[2], v[6], v[10], v[14]);
    radix4<INV>(v[3], v[7], v[11], v[15]);
    dft16_layer2<INV>(v);
}

// 16-point DFT of (s[n] * v[n]) with real scales s (the analysis window fused into stage A's first layer)
template <bool INV>
TMT_HD void dft16_scaled(float2 (&v)[16], const float (&s)[16]) {
    radix4_scaled<INV>(v[0], v[4], v[8], v[12], s[0], s[4], s[8], s[12]);
    radix4_scaled<INV>(v[1], v[5], v[9], v[13], s[1], s[5], s[9], s[13]);
    radix4_scaled<INV>(v[2], v[6], v[10], v[14], s[2], s[6], s[10], s[14]);
    radix4_scaled<INV>(v[3], v[7], v[11], v[15], s[3], s[7], s[11], s[15]);
    dft16_layer2<INV>(v);
}

// Inverse 16-point DFT of (v[k] * conj(p[k])), k = 0..15, p[0] = 1: the inter-stage twiddles of the inverse transform are
// fused into the first radix-4 layer (8 packed instructions fewer than multiply-then-transform).
TMT_HD void dft16_inv_tw(float2 (&v)[16], const float2 (&p)[16]) {
    radix4_tw<true, true, false>(v[0], v[4], v[8], v[12], p[0], p[4], p[8], p[12]);
    radix4_tw<true, true, true>(v[1], v[5], v[9], v[13], p[1], p[5], p[9], p[13]);
    radix4_tw<true, true, true>(v[2], v[6], v[10], v[14], p[2], p[6], p[10], p[14]);
    radix4_tw<true, true, true>(v[3], v[7], v[11], v[15], p[3], p[7], p[11], p[15]);
    dft16_layer2<true>(v);
}

// ---- thread maps ------------------------------------------------------------------------------
// A side (and the time domain): thread t holds n = 256*j + t, i.e. (n2, n3) = (t >> 4, t & 15).
// B side: warp w = t >> 5 owns k1 in {2w, 2w + 1}; lane = 2*n3 + (k1 & 1).
// C side (after the tensor-memory exchange): k1 unchanged per warp, lane = 16*(k1 & 1) + 8*q1 + 4*q0 + 2*q3 + q2 for
// k2 = (q3 q2 q1 q0), i.e. k1 = t >> 4 and k2 = 4*(t & 3) + ((t >> 2) & 3).
TMT_HD int b_k1(int t) { return 2 * (t >> 5) + (t & 1); }
TMT_HD int b_n3(int t) { return (t & 31) >> 1; }
TMT_HD int c_k1(int t) { return t >> 4; }
TMT_HD int c_k2(int t) { return 4 * (t & 3) + ((t >> 2) & 3); }
// bin held in register j of thread t after stage C (and expected by stage C')
TMT_HD int bin_of(int t, int j) { return c_k1(t) + 16 * c_k2(t) + 256 * j; }

// ---- shared-memory exchange E1 (float2 units):  idx = k1*264 + n2*16 + n3 ----------------------
#ifdef TMT_E1_PADDED_ROWS
TMT_HD int e1_a(int t, int j) { return j * kE1Row + t; }
TMT_HD int e1_b(int t, int j) { return b_k1(t) * kE1Row + j * 16 + b_n3(t); }
#else
TMT_HD int e1_a(int t, int j) { return (j >> 1) * 512 + 2 * t + (j & 1); }
TMT_HD int e1_b(int t, int j) { return (b_k1(t) >> 1) * 512 + 2 * (j * 16 + b_n3(t)) + (b_k1(t) & 1); }     // = (t >> 5) * 512 + 32 * j + (t & 31)
#endif

// Twiddles.  Both twiddle stages have the form v[k] *= b^k with a PER-THREAD base:
//   stage A (thread t = 16*n2 + n3, output k1):  W256^(n2*k1) * W4096^(n3*k1) = (W4096^t)^k1
//   stage B (thread (k1,n3),        output k2):  W256^(n3*k2)                 = (W256^n3)^k2
// ncu showed the shared-memory data pipe (not FP32) to be the limiter of this kernel, and twiddle-table
// reads were a quarter of its wavefronts, so the powers are recomputed every frame from b and b^4
// (exact, host-computed in double, 8 registers per thread): 14 extra complex multiplies per stage,
// chain depth <= 3 roundings, no table traffic at all.
struct TwBase { float2 b1, b4; };

template <bool CONJ>
TMT_HD void tw_pow(float2 (&v)[16], const TwBase w) {
    const float2 b1 = w.b1, b4 = w.b4;
    const float2 b2 = cmul(b1, b1), b3 = cmul(b2, b1);
    const float2 b8 = cmul(b4, b4), b12 = cmul(b8, b4);
#define TMT_TW(k, p) v[k] = CONJ ? cmulc(v[k], p) : cmul(v[k], p)
    TMT_TW(1, b1); TMT_TW(2, b2); TMT_TW(3, b3); TMT_TW(4, b4);
    TMT_TW(5, cmul(b4, b1)); TMT_TW(6, cmul(b4, b2)); TMT_TW(7, cmul(b4, b3)); TMT_TW(8, b8);
    TMT_TW(9, cmul(b8, b1)); TMT_TW(10, cmul(b8, b2)); TMT_TW(11, cmul(b8, b3)); TMT_TW(12, b12);
    TMT_TW(13, cmul(b12, b1)); TMT_TW(14, cmul(b12, b2)); TMT_TW(15, cmul(b12, b3));
#undef TMT_TW
}

// p[k] = b^k, k = 0..15, from the per-thread bases (same products as tw_pow)
TMT_HD void tw_table(float2 (&p)[16], const TwBase w) {
    const float2 b1 = w.b1, b4 = w.b4;
    const float2 b2 = cmul(b1, b1), b3 = cmul(b2, b1);
    const float2 b8 = cmul(b4, b4), b12 = cmul(b8, b4);
    p[0] = make_float2(1.f, 0.f); p[1] = b1; p[2] = b2; p[3] = b3;
    p[4] = b4; p[5] = cmul(b4, b1); p[6] = cmul(b4, b2); p[7] = cmul(b4, b3);
    p[8] = b8; p[9] = cmul(b8, b1); p[10] = cmul(b8, b2); p[11] = cmul(b8, b3);
    p[12] = b12; p[13] = cmul(b12, b1); p[14] = cmul(b12, b2); p[15] = cmul(b12, b3);
}

// ---- exchange pieces (E1) -----------------------------------------------------------------------
// The device accesses are spelled as PTX: written as plain C++ assignments, ptxas copied most of the values into one fixed register
// pair in front of their STS.64 (36 MOVs per frame in the two store groups; -34 instructions, -1.5 % kernel time,
// profiles/r02/ab_asm_sts.txt).
#if defined(__CUDA_ARCH__) && !defined(TMT_E1_PADDED_ROWS)
TMT_HD void st_e1a(const float2 (&v)[16], int t, float2* buf) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(buf + 2 * t);
#pragma unroll
    for (int jp = 0; jp < 8; ++jp)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a + jp * 4096), "f"(v[2 * jp].x), "f"(v[2 * jp].y), "f"(v[2 * jp + 1].x),
                     "f"(v[2 * jp + 1].y) : "memory");
}
TMT_HD void ld_e1a(float2 (&v)[16], int t, const float2* buf) {
    const float4* p = reinterpret_cast<const float4*>(buf) + t;
#pragma unroll
    for (int jp = 0; jp < 8; ++jp) {
        const float4 x = p[jp * 256];
        v[2 * jp] = make_float2(x.x, x.y);
        v[2 * jp + 1] = make_float2(x.z, x.w);
    }
}
TMT_HD void ld_e1b(float2 (&v)[16], int t, const float2* buf) {
    const float2* p = buf + e1_b(t, 0);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = p[j * 32];
}
TMT_HD void st_e1b(const float2 (&v)[16], int t, float2* buf) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(buf + e1_b(t, 0));
#pragma unroll
    for (int j = 0; j < 16; ++j) asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a + j * 256), "f"(v[j].x), "f"(v[j].y) : "memory");
}
#else
TMT_HD void st_e1a(const float2 (&v)[16], int t, float2* buf) {
#if defined(__CUDA_ARCH__)
    const unsigned a = (unsigned)__cvta_generic_to_shared(buf + t);
#pragma unroll
    for (int j = 0; j < 16; ++j) asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a + j * kE1Row * 8), "f"(v[j].x), "f"(v[j].y) : "memory");
#else
#pragma unroll
    for (int j = 0; j < 16; ++j) buf[e1_a(t, j)] = v[j];
#endif
}
TMT_HD void ld_e1b(float2 (&v)[16], int t, const float2* buf) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = buf[e1_b(t, j)];
}
TMT_HD void st_e1b(const float2 (&v)[16], int t, float2* buf) {
#if defined(__CUDA_ARCH__)
    const unsigned a = (unsigned)__cvta_generic_to_shared(buf + e1_b(t, 0));
#pragma unroll
    for (int j = 0; j < 16; ++j) asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a + j * 128), "f"(v[j].x), "f"(v[j].y) : "memory");
#else
#pragma unroll
    for (int j = 0; j < 16; ++j) buf[e1_b(t, j)] = v[j];
#endif
}
TMT_HD void ld_e1a(float2 (&v)[16], int t, const float2* buf) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = buf[e1_a(t, j)];
}
#endif

// ---- exchange E2 through tensor memory: register naming around the four round trips -------------
// A round trip works on the raw 32-register image of a thread.  "Column order" is what .32x32b stores / loads
// (register c <-> column c: float2 index c/2, component c%2); "row order" is what .16x256b loads / stores
// (register 16*I + 4*g + 2*h + e <-> lane 16*I + 8*h + T/4, column 8*g + 2*(T%4) + e of thread T).
TMT_HD int row_reg(int hi2, int g) { return 16 * (hi2 >> 1) + 4 * g + 2 * (hi2 & 1); }   // float2 (2-bit arrival index hi2, column group g)

// forward trip 1, send: B outputs v[k2] in column order
TMT_HD void x_fwd1_pack(const float2 (&v)[16], float (&r)[32]) {
#pragma unroll
    for (int q = 0; q < 16; ++q) { r[2 * q] = v[q].x; r[2 * q + 1] = v[q].y; }
}
// forward, between the trips: r (row order) holds x[a][g] (a = n3 >> 2 just arrived, g = k2 >> 2 still here);
// first radix-4 layer of stage C over a, result x[c][g] packed in column order at float2 column 4*c + g
template <bool INV>
TMT_HD void x_layer_a(const float (&r)[32], float (&s)[32]) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float2 x[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) x[a] = make_float2(r[row_reg(a, g)], r[row_reg(a, g) + 1]);
        radix4<INV>(x[0], x[1], x[2], x[3]);
#pragma unroll
        for (int c = 0; c < 4; ++c) { s[2 * (4 * c + g)] = x[c].x; s[2 * (4 * c + g) + 1] = x[c].y; }
    }
}
// forward trip 2, receive: s (row order) holds y[b][c] (b = n3 & 3 just arrived, c = column group); second layer of stage C
// (twiddles W16^(b*c) fused) -> v[k3] natural
TMT_HD void x_fwd2_finish(const float (&s)[32], float2 (&v)[16]) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int b = 0; b < 4; ++b) v[4 * c + b] = make_float2(s[row_reg(b, c)], s[row_reg(b, c) + 1]);
    dft16_layer2<false>(v);
}
// conj(W16^M) * v for the inverse transform's inner twiddles (M = b*c)
template <int M>
TMT_HD float2 mul_w16c(float2 v) {
    if (M == 0) return v;
    if (M == 4) return make_float2(-v.y, v.x);                  // conj(-i) = +i
    if (M == 1) return cmulc(v, w16<1>());
    if (M == 2) return cmulc(v, w16<2>());
    if (M == 3) return cmulc(v, w16<3>());
    if (M == 6) return cmulc(v, w16<6>());
    return cmulc(v, w16<9>());
}
// inverse trip 1 (undoes forward trip 2), send: first layer of stage C' over d (k3 = c + 4d) gives b = n3 & 3, inner
// twiddles conj(W16^(b*c)) applied while both indices are still in registers; packed in row order with b as the index that leaves
// g: optional real gains per input (the tilt gain row), fused into the first layer's butterflies
TMT_HD void x_inv1_pack(float2 (&v)[16], float (&r)[32], const float* g = nullptr) {
    if (g) {
        radix4_scaled<true>(v[0], v[4], v[8], v[12], g[0], g[4], g[8], g[12]);
        radix4_scaled<true>(v[1], v[5], v[9], v[13], g[1], g[5], g[9], g[13]);
        radix4_scaled<true>(v[2], v[6], v[10], v[14], g[2], g[6], g[10], g[14]);
        radix4_scaled<true>(v[3], v[7], v[11], v[15], g[3], g[7], g[11], g[15]);
    } else {
        radix4<true>(v[0], v[4], v[8], v[12]);
        radix4<true>(v[1], v[5], v[9], v[13]);
        radix4<true>(v[2], v[6], v[10], v[14]);
        radix4<true>(v[3], v[7], v[11], v[15]);
    }                                                            // now v[4b + c]
#define TMT_XW(b, c) { const float2 y = mul_w16c<(b) * (c)>(v[4 * (b) + (c)]); r[row_reg(b, c)] = y.x; r[row_reg(b, c) + 1] = y.y; }
    TMT_XW(0, 0) TMT_XW(0, 1) TMT_XW(0, 2) TMT_XW(0, 3)
    TMT_XW(1, 0) TMT_XW(1, 1) TMT_XW(1, 2) TMT_XW(1, 3)
    TMT_XW(2, 0) TMT_XW(2, 1) TMT_XW(2, 2) TMT_XW(2, 3)
    TMT_XW(3, 0) TMT_XW(3, 1) TMT_XW(3, 2) TMT_XW(3, 3)
#undef TMT_XW
}
// inverse, between the trips: r (column order) holds x[c][g] at float2 column 4*c + g (g = k2 >> 2 is back); second layer
// of stage C' over c gives a = n3 >> 2, packed in row order with a as the index that leaves
TMT_HD void x_layer_c_inv(const float (&r)[32], float (&s)[32]) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float2 x[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = make_float2(r[2 * (4 * c + g)], r[2 * (4 * c + g) + 1]);
        radix4<true>(x[0], x[1], x[2], x[3]);
#pragma unroll
        for (int a = 0; a < 4; ++a) { s[row_reg(a, g)] = x[a].x; s[row_reg(a, g) + 1] = x[a].y; }
    }
}
// inverse trip 2, receive: column order = v[k2] natural
TMT_HD void x_inv2_unpack(const float (&s)[32], float2 (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = make_float2(s[2 * q], s[2 * q + 1]);
}

// ---- pair mode: two 2048-point transforms in one 4096-wide pass ---------------------------------------------------------
// z[2m + p] = frame_p[m] (p = 0, 1; m = 0..2047).  A decimation-in-time 4096-point transform of z that stops before its last
// radix-2 butterfly IS the two 2048-point transforms of the even and the odd samples.  With n = 256*n1 + 16*n2 + n3 the parity
// p is the lowest bit of n3, which no stage but C combines: stages A and B and both exchanges keep their layouts and only see
// other twiddles -- m = 128*n1 + 8*n2 + n3' (n3 = 2*n3' + p), k = k1 + 16*k2 + 256*k3' (k3' = 0..7):
//     stage A: DFT16 over n1, twiddle W2048^((8*n2 + n3')*k1) = (W4096^(t & ~1))^k1         (thread t = 16*n2 + n3)
//     stage B: DFT16 over n2, twiddle W128^(n3'*k2)           = (W256^(n3 & ~1))^k2
//     stage C: DFT8 over n3' = 2a + b1 for each p: the first radix-4 layer over a is the 4096 one (n3 = 4a + b, b = 2*b1 + p);
//              the second layer shrinks to a radix-2 over b1 with twiddle W8^(b1*c): X_p[c + 4d'] = y[p][c] + (-1)^d' W8^c y[2+p][c]
// After stage C register j = 8p + k3' of thread (k1, k2) holds bin k1 + 16*k2 + 256*k3' of frame p.
template <int C>
TMT_HD float2 w8() {                 // W8^C = e^{-2*pi*i*C/8}, C = 1, 3 (0 and 2 are handled without a multiply)
    constexpr float kH = 0.70710678118654752440f;
    static_assert(C == 1 || C == 3, "unsupported W8 power");
    return (C == 1) ? make_float2(kH, -kH) : make_float2(-kH, -kH);
}
// forward, second layer of stage C: s (row order) holds y[b][c] (b = n3 & 3 just arrived, c = column group)
TMT_HD void x_fwd2_finish_pair(const float (&s)[32], float2 (&v)[16]) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
#define TMT_Y(b, c) make_float2(s[row_reg(b, c)], s[row_reg(b, c) + 1])
        { const float2 lo = TMT_Y(p, 0), hi = TMT_Y(2 + p, 0); v[8 * p + 0] = cadd(lo, hi); v[8 * p + 4] = csub(lo, hi); }
        { const float2 lo = TMT_Y(p, 1), t = cfma(lo, TMT_Y(2 + p, 1), w8<1>()); v[8 * p + 1] = t; v[8 * p + 5] = twice_minus(lo, t); }
        { const float2 lo = TMT_Y(p, 2), hi = TMT_Y(2 + p, 2); v[8 * p + 2] = cadd_mi(lo, hi); v[8 * p + 6] = cadd_pi(lo, hi); }
        { const float2 lo = TMT_Y(p, 3), t = cfma(lo, TMT_Y(2 + p, 3), w8<3>()); v[8 * p + 3] = t; v[8 * p + 7] = twice_minus(lo, t); }
#undef TMT_Y
    }
}
// inverse, first layer of stage C' (undoes x_fwd2_finish_pair): radix-2 over d' for each (c, p) with the real gains fused in,
// inner twiddle conj(W8^(b1*c)), packed in row order with b = 2*b1 + p as the index that leaves
TMT_HD void x_inv1_pack_pair(const float2 (&v)[16], float (&r)[32], const float* g = nullptr) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float2 a = v[8 * p + c], b = v[8 * p + c + 4];
            float2 y0, y1;
            if (g) {
                const float2 p2 = cscale(b, g[8 * p + c + 4]);
                y0 = cfms(a, g[8 * p + c], p2);
                y1 = cfmsn(a, g[8 * p + c], p2);
            } else {
                y0 = cadd(a, b);
                y1 = csub(a, b);
            }
            if (c == 1) y1 = cmulc(y1, w8<1>());
            if (c == 2) y1 = make_float2(-y1.y, y1.x);              // conj(-i) = +i
            if (c == 3) y1 = cmulc(y1, w8<3>());
            r[row_reg(p, c)] = y0.x; r[row_reg(p, c) + 1] = y0.y;
            r[row_reg(2 + p, c)] = y1.x; r[row_reg(2 + p, c) + 1] = y1.y;
        }
    }
}
// bin of frame (j >> 3) held in register j of thread t after stage C in pair mode
TMT_HD int bin_of_pair(int t, int j) { return c_k1(t) + 16 * c_k2(t) + 256 * (j & 7); }

// The data movement of the round trips, for the CPU emulation (csrc/host_emul.cu): src/dst index a warp's 32 x 32 register image.
// forward trip (st .32x32b, ld .16x256b): thread T register 16*I + 4*g + 2*h + e  <-  thread (16*I + 8*h + T/4) register 8*g + 2*(T%4) + e
inline void emul_trip_fwd(const float* src, float* dst) {
    for (int T = 0; T < 32; ++T)
        for (int I = 0; I < 2; ++I) for (int g = 0; g < 4; ++g) for (int h = 0; h < 2; ++h) for (int e = 0; e < 2; ++e)
            dst[T * 32 + 16 * I + 4 * g + 2 * h + e] = src[(16 * I + 8 * h + T / 4) * 32 + 8 * g + 2 * (T % 4) + e];
}
// inverse trip (st .16x256b, ld .32x32b): the inverse permutation
inline void emul_trip_inv(const float* src, float* dst) {
    for (int T = 0; T < 32; ++T)
        for (int I = 0; I < 2; ++I) for (int g = 0; g < 4; ++g) for (int h = 0; h < 2; ++h) for (int e = 0; e < 2; ++e)
            dst[(16 * I + 8 * h + T / 4) * 32 + 8 * g + 2 * (T % 4) + e] = src[T * 32 + 16 * I + 4 * g + 2 * h + e];
}

}  // namespace tmt
