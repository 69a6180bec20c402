// 4096-point complex FFT building blocks for one 256-thread CTA (sm_100a), fp32.
//
// The reference does, per STFT frame and per channel, rfft -> real gain -> irfft
// (/root/reference/src/process_tomatis.py:394-398).  Here the stereo pair is packed as one
// complex signal z = L + iR (the gain is real and symmetric, so IFFT(g*FFT(z)) = yL + i*yR),
// and N = 4096 = 16*16*16 is done as three register-resident radix-16 stages with two
// shared-memory exchanges per direction.  Index split
//     n = 256*n1 + 16*n2 + n3          k = k1 + 16*k2 + 256*k3
// forward:  A (threads (n2,n3), DFT over n1, twiddle W4096^((16*n2+n3)*k1))  -> smem E1
//           B (threads (k1,n3), DFT over n2, twiddle W256^(n3*k2)) -> smem E2
//           C (threads (k1,k2), DFT over n3)  -> X[k1+16*k2+256*k3] in registers
// inverse:  exactly the mirror (C' B' A') with conjugate twiddles applied on stage inputs, so
// the spectrum never leaves registers between forward and inverse and the time-domain result
// lands in the same thread/register layout the input was loaded in (n = 256*j + t).
//
// Everything here is __host__ __device__ so tests/ can run the same code on the CPU
// (csrc/host_emul.cu) -- there is no GPU in the build container.
#pragma once
#include <cuda_runtime.h>

#define TMT_HD __host__ __device__ __forceinline__

namespace tmt {

constexpr int kNfft = 4096;
constexpr int kHop = 2048;
constexpr int kThreads = 256;       // one frame per CTA pass, 16 points per thread
constexpr int kRowPad = 18;         // E2 row stride in float2 (16 + 2): rows stay 16-byte aligned, so the C side moves two
                                    // points per 128-bit access (8 instead of 16 instructions); conflict-free both ways
constexpr int kExchFloat2 = 256 * kRowPad;   // the E2 exchange buffer (36 864 B)

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

// Inverse 16-point DFT of (v[k] * conj(p[k])), k = 0..15, p[0] = 1: the inter-stage twiddles of the inverse transform are
// fused into the first radix-4 layer (8 packed instructions fewer than multiply-then-transform).
TMT_HD void dft16_inv_tw(float2 (&v)[16], const float2 (&p)[16]) {
    radix4_tw<true, true, false>(v[0], v[4], v[8], v[12], p[0], p[4], p[8], p[12]);
    radix4_tw<true, true, true>(v[1], v[5], v[9], v[13], p[1], p[5], p[9], p[13]);
    radix4_tw<true, true, true>(v[2], v[6], v[10], v[14], p[2], p[6], p[10], p[14]);
    radix4_tw<true, true, true>(v[3], v[7], v[11], v[15], p[3], p[7], p[11], p[15]);
    dft16_layer2<true>(v);
}

// ---- shared-memory exchange layouts (float2 units) -------------------------------------
// E1 (linear):  idx = k1*256 + n2*16 + n3        A side: j*256 + t       B side: (t>>4)*256 + j*16 + (t&15)
// E2 (padded):  idx = (k1*16 + k2)*18 + n3       B side: ((t>>4)*16 + j)*18 + (t&15)    C side: t*18 + j
TMT_HD int e1_a(int t, int j) { return j * 256 + t; }
TMT_HD int e1_b(int t, int j) { return (t >> 4) * 256 + j * 16 + (t & 15); }
TMT_HD int e2_b(int t, int j) { return ((t >> 4) * 16 + j) * kRowPad + (t & 15); }
TMT_HD int e2_c(int t, int j) { return t * kRowPad + j; }

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

// ---- exchange pieces -----------------------------------------------------------------------
TMT_HD void st_e1a(const float2 (&v)[16], int t, float2* buf) {
#pragma unroll
    for (int j = 0; j < 16; ++j) buf[e1_a(t, j)] = v[j];
}
TMT_HD void ld_e1b(float2 (&v)[16], int t, const float2* buf) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = buf[e1_b(t, j)];
}
TMT_HD void st_e2b(const float2 (&v)[16], int t, float2* buf) {
#pragma unroll
    for (int j = 0; j < 16; ++j) buf[e2_b(t, j)] = v[j];
}
TMT_HD void ld_e2c(float2 (&v)[16], int t, const float2* buf) {
    const float4* row = reinterpret_cast<const float4*>(buf + e2_c(t, 0));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 x = row[j];
        v[2 * j] = make_float2(x.x, x.y);
        v[2 * j + 1] = make_float2(x.z, x.w);
    }
}
TMT_HD void st_e2c(const float2 (&v)[16], int t, float2* buf) {
    float4* row = reinterpret_cast<float4*>(buf + e2_c(t, 0));
#pragma unroll
    for (int j = 0; j < 8; ++j) row[j] = make_float4(v[2 * j].x, v[2 * j].y, v[2 * j + 1].x, v[2 * j + 1].y);
}
TMT_HD void ld_e2b(float2 (&v)[16], int t, const float2* buf) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = buf[e2_b(t, j)];
}
TMT_HD void st_e1b(const float2 (&v)[16], int t, float2* buf) {
#pragma unroll
    for (int j = 0; j < 16; ++j) buf[e1_b(t, j)] = v[j];
}
TMT_HD void ld_e1a(float2 (&v)[16], int t, const float2* buf) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = buf[e1_a(t, j)];
}

// ---- composite forward stages (in: v[j] = windowed z[256*j + t]) ---------------------------
TMT_HD void fwd_a(float2 (&v)[16], int t, const TwBase wa, float2* bufP) {
    dft16<false>(v);
    tw_pow<false>(v, wa);
    st_e1a(v, t, bufP);
}
TMT_HD void fwd_b(float2 (&v)[16], int t, const TwBase wb, const float2* bufP, float2* bufQ) {
    ld_e1b(v, t, bufP);
    dft16<false>(v);
    tw_pow<false>(v, wb);
    st_e2b(v, t, bufQ);
}
// out: v[j] = Z[(t>>4) + 16*(t&15) + 256*j]
TMT_HD void fwd_c(float2 (&v)[16], int t, const float2* bufQ) {
    ld_e2c(v, t, bufQ);
    dft16<false>(v);
}
// ---- composite inverse stages (unnormalised; the 1/4096 is folded into the gain table) ------
TMT_HD void inv_c(float2 (&v)[16], int t, float2* bufP) {
    dft16<true>(v);
    st_e2c(v, t, bufP);
}
TMT_HD void inv_b(float2 (&v)[16], int t, const TwBase wb, const float2* bufP, float2* bufQ) {
    ld_e2b(v, t, bufP);
    float2 p[16];
    tw_table(p, wb);
    dft16_inv_tw(v, p);
    st_e1b(v, t, bufQ);
}
// out: v[j] = 4096 * y[256*j + t]
TMT_HD void inv_a(float2 (&v)[16], int t, const TwBase wa, const float2* bufQ) {
    ld_e1a(v, t, bufQ);
    float2 p[16];
    tw_table(p, wa);
    dft16_inv_tw(v, p);
}

// bin held in register j of thread t after fwd_c (and expected by inv_c)
TMT_HD int bin_of(int t, int j) { return (t >> 4) + 16 * (t & 15) + 256 * j; }

}  // namespace tmt
