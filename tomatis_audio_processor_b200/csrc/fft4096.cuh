// 4096-point complex FFT building blocks for one 256-thread CTA (sm_100a), fp32.
//
// The reference does, per STFT frame and per channel, rfft -> real gain -> irfft
// (/root/reference/src/process_tomatis.py:394-398).  Here the stereo pair is packed as one
// complex signal z = L + iR (the gain is real and symmetric, so IFFT(g*FFT(z)) = yL + i*yR),
// and N = 4096 = 16*16*16 is done as three register-resident radix-16 stages with two
// shared-memory exchanges per direction.  Index split
//     n = 256*n1 + 16*n2 + n3          k = k1 + 16*k2 + 256*k3
// forward:  A (threads (n2,n3), DFT over n1, twiddle W256^(n2*k1))  -> smem E1
//           B (threads (k1,n3), DFT over n2, twiddle W4096^(n3*(k1+16*k2))) -> smem E2
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
constexpr int kRowPad = 17;         // E2 row stride in float2 (16 + 1): conflict-free both ways
constexpr int kExchFloat2 = 256 * kRowPad;   // one exchange buffer (34 816 B)

TMT_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
TMT_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
TMT_HD float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
// a * conj(w)
TMT_HD float2 cmulc(float2 a, float2 w) { return make_float2(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y); }

// multiply by W16^M (forward, e^{-2*pi*i*M/16}) or its conjugate (INV)
template <int M, bool INV>
TMT_HD float2 mul_w16(float2 a) {
    constexpr float kC1 = 0.92387953251128673848f;   // cos(pi/8)
    constexpr float kS1 = 0.38268343236508978178f;   // sin(pi/8)
    constexpr float kH = 0.70710678118654752440f;    // sqrt(1/2)
    if constexpr (M == 0) {
        return a;
    } else if constexpr (M == 4) {                   // -i (fwd) / +i (inv)
        return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
    } else if constexpr (M == 2) {                   // (1 -/+ i)/sqrt2
        return INV ? make_float2((a.x - a.y) * kH, (a.x + a.y) * kH)
                   : make_float2((a.x + a.y) * kH, (a.y - a.x) * kH);
    } else if constexpr (M == 6) {                   // (-1 -/+ i)/sqrt2
        return INV ? make_float2((-a.x - a.y) * kH, (a.x - a.y) * kH)
                   : make_float2((a.y - a.x) * kH, (-a.x - a.y) * kH);
    } else {
        // general: W16^M = (c, -s) forward, (c, +s) inverse
        constexpr float c = (M == 1) ? kC1 : (M == 3) ? kS1 : (M == 9) ? -kC1 : 0.f;
        constexpr float s = (M == 1) ? kS1 : (M == 3) ? kC1 : (M == 9) ? -kS1 : 0.f;
        static_assert(M == 1 || M == 3 || M == 9, "unsupported W16 power");
        return INV ? make_float2(a.x * c - a.y * s, a.y * c + a.x * s)
                   : make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
    }
}

// 4-point DFT in place: (x0..x3) -> (X0..X3), forward W4 = -i, inverse W4 = +i
template <bool INV>
TMT_HD void radix4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = csub(x1, x3);
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    if (INV) {
        x1 = make_float2(t1.x - t3.y, t1.y + t3.x);
        x3 = make_float2(t1.x + t3.y, t1.y - t3.x);
    } else {
        x1 = make_float2(t1.x + t3.y, t1.y - t3.x);
        x3 = make_float2(t1.x - t3.y, t1.y + t3.x);
    }
}

// 16-point DFT, natural order in and out, fully in registers:
//   n = 4a + b, k = c + 4d:  V[c+4d] = sum_b W4^(bd) * W16^(bc) * sum_a W4^(ac) v[4a+b]
template <bool INV>
TMT_HD void dft16(float2 (&v)[16]) {
    // step 1: radix-4 over a for each b; u_b[c] is left in v[4c + b]
    radix4<INV>(v[0], v[4], v[8], v[12]);
    radix4<INV>(v[1], v[5], v[9], v[13]);
    radix4<INV>(v[2], v[6], v[10], v[14]);
    radix4<INV>(v[3], v[7], v[11], v[15]);
    // step 2: twiddle u_b[c] *= W16^(b*c)   (v index 4c + b)
    v[5] = mul_w16<1, INV>(v[5]);
    v[6] = mul_w16<2, INV>(v[6]);
    v[7] = mul_w16<3, INV>(v[7]);
    v[9] = mul_w16<2, INV>(v[9]);
    v[10] = mul_w16<4, INV>(v[10]);
    v[11] = mul_w16<6, INV>(v[11]);
    v[13] = mul_w16<3, INV>(v[13]);
    v[14] = mul_w16<6, INV>(v[14]);
    v[15] = mul_w16<9, INV>(v[15]);
    // step 3: radix-4 over b for each c; V[c + 4d] is left in v[4c + d]
    radix4<INV>(v[0], v[1], v[2], v[3]);
    radix4<INV>(v[4], v[5], v[6], v[7]);
    radix4<INV>(v[8], v[9], v[10], v[11]);
    radix4<INV>(v[12], v[13], v[14], v[15]);
    // step 4: 4x4 transpose of the register names (free after unrolling)
    float2 t;
    t = v[1]; v[1] = v[4]; v[4] = t;
    t = v[2]; v[2] = v[8]; v[8] = t;
    t = v[3]; v[3] = v[12]; v[12] = t;
    t = v[6]; v[6] = v[9]; v[9] = t;
    t = v[7]; v[7] = v[13]; v[13] = t;
    t = v[11]; v[11] = v[14]; v[14] = t;
}

// ---- shared-memory exchange layouts (float2 units) -------------------------------------
// E1 (linear):  idx = k1*256 + n2*16 + n3        A side: j*256 + t       B side: (t>>4)*256 + j*16 + (t&15)
// E2 (padded):  idx = (k1*16 + k2)*17 + n3       B side: ((t>>4)*16 + j)*17 + (t&15)    C side: t*17 + j
TMT_HD int e1_a(int t, int j) { return j * 256 + t; }
TMT_HD int e1_b(int t, int j) { return (t >> 4) * 256 + j * 16 + (t & 15); }
TMT_HD int e2_b(int t, int j) { return ((t >> 4) * 16 + j) * kRowPad + (t & 15); }
TMT_HD int e2_c(int t, int j) { return t * kRowPad + j; }

// twiddle tables (built on the host in double precision, see tomatis_b200.cu):
//   twA[k1*16 + n2]  = W256^(n2*k1)                       (256 entries; warp-broadcast reads)
//   twB[k2*256 + t]  = W4096^((t&15) * ((t>>4) + 16*k2))  (4096 entries; coalesced reads)

// ---- forward stages ---------------------------------------------------------------------
// in : v[j] = windowed z[256*j + t]
TMT_HD void fwd_a(float2 (&v)[16], int t, const float2* twA, float2* bufP) {
    dft16<false>(v);
    const int n2 = t >> 4;
#pragma unroll
    for (int j = 1; j < 16; ++j) v[j] = cmul(v[j], twA[j * 16 + n2]);
#pragma unroll
    for (int j = 0; j < 16; ++j) bufP[e1_a(t, j)] = v[j];
}
TMT_HD void fwd_b(float2 (&v)[16], int t, const float2* twB, const float2* bufP, float2* bufQ) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = bufP[e1_b(t, j)];
    dft16<false>(v);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = cmul(v[j], twB[j * 256 + t]);
#pragma unroll
    for (int j = 0; j < 16; ++j) bufQ[e2_b(t, j)] = v[j];
}
// out: v[j] = Z[(t>>4) + 16*(t&15) + 256*j]
TMT_HD void fwd_c(float2 (&v)[16], int t, const float2* bufQ) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = bufQ[e2_c(t, j)];
    dft16<false>(v);
}
// ---- inverse stages (unnormalised; the 1/4096 is folded into the gain table) --------------
TMT_HD void inv_c(float2 (&v)[16], int t, float2* bufP) {
    dft16<true>(v);
#pragma unroll
    for (int j = 0; j < 16; ++j) bufP[e2_c(t, j)] = v[j];
}
TMT_HD void inv_b(float2 (&v)[16], int t, const float2* twB, const float2* bufP, float2* bufQ) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = cmulc(bufP[e2_b(t, j)], twB[j * 256 + t]);
    dft16<true>(v);
#pragma unroll
    for (int j = 0; j < 16; ++j) bufQ[e1_b(t, j)] = v[j];
}
// out: v[j] = 4096 * y[256*j + t]
TMT_HD void inv_a(float2 (&v)[16], int t, const float2* twA, const float2* bufQ) {
    const int n2 = t >> 4;
    v[0] = bufQ[e1_a(t, 0)];
#pragma unroll
    for (int j = 1; j < 16; ++j) v[j] = cmulc(bufQ[e1_a(t, j)], twA[j * 16 + n2]);
    dft16<true>(v);
}

// bin held in register j of thread t after fwd_c (and expected by inv_c)
TMT_HD int bin_of(int t, int j) { return (t >> 4) + 16 * (t & 15) + 256 * j; }

}  // namespace tmt
