"""Numerical feasibility of doing the radix-16 stages of the 4096-point FFT on tensor cores (tcgen05.mma, fp32 accumulate)
with split operands -- CPU study in NumPy, no GPU needed.  Question: which operand format / number of partial products keeps
the STFT -> gain -> ISTFT round trip inside the 1e-5-of-full-scale parity bar (BASELINE.json), given that the reference
itself computes in float32 (pocketfft, ~1e-7)?

Model.  N = 4096 = 16^3, three Cooley-Tukey stages.  Each stage is Y = F16 @ X over the stage's 16-point axis (complex F16
as the real 32 x 32 block matrix [[Fr, -Fi], [Fi, Fr]]) followed by the inter-stage twiddles in float32 on CUDA cores.
Operands are rounded to the tensor-core input format (tf32: 10-bit mantissa, fp16: 10-bit mantissa + 5-bit exponent with a
per-frame power-of-two scale, bf16: 7-bit), split hi + lo, and the kept partial products are accumulated in float32.
Prints the max abs error of the round trip x -> IFFT(g * FFT(w * x)) * w against float64, for a full-scale noise frame.

    python tools/experiments/tc_fft_feasibility.py
"""
import numpy as np


def round_mantissa(x, bits):
    """Round float32 values to `bits` explicit mantissa bits (round to nearest even), exponent range of float32."""
    x = np.asarray(x, np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    drop = 23 - bits
    u = (u + ((1 << (drop - 1)) - 1) + ((u >> drop) & 1)) >> drop << drop
    return u.astype(np.uint32).view(np.float32)


def to_fmt(x, fmt):
    if fmt == "tf32":
        return round_mantissa(x, 10)
    if fmt == "bf16":
        return round_mantissa(x, 7)
    if fmt == "fp16":
        with np.errstate(over="ignore"):
            return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)
    raise ValueError(fmt)


def split(x, fmt, terms):
    parts, r = [], np.asarray(x, np.float32)
    for _ in range(terms):
        p = to_fmt(r, fmt)
        parts.append(p)
        r = (r - p).astype(np.float32)
    return parts


def dft16_block(inv):
    k = np.arange(16)
    f = np.exp((2j if inv else -2j) * np.pi * np.outer(k, k) / 16)
    return np.block([[f.real, -f.imag], [f.imag, f.real]])           # float64 [32, 32]


def tc_stage(z, inv, fmt, x_terms, f_terms, products):
    """z complex64 [16, M] -> F16 @ z on the 'tensor core': fp32 accumulation of the kept hi/lo partial products."""
    x = np.concatenate([z.real, z.imag]).astype(np.float32)          # [32, M]
    scale = np.float32(1.0)
    if fmt == "fp16":                                                # per-frame power-of-two scale into fp16's range
        scale = np.float32(2.0 ** -np.ceil(np.log2(max(np.abs(x).max(), 1e-30) / 1024.0)))
    xs = split(x * scale, fmt, x_terms)
    fs = split(dft16_block(inv).astype(np.float32), fmt, f_terms)
    acc = np.zeros_like(x, dtype=np.float32)
    for fi, xi in products:
        if fi < f_terms and xi < x_terms:
            acc = (acc + (fs[fi].astype(np.float64) @ xs[xi].astype(np.float64)).astype(np.float32)).astype(np.float32)
    acc = acc / scale
    return (acc[:16] + 1j * acc[16:]).astype(np.complex64)


def fft4096(z, inv, stage):
    """Three radix-16 stages, twiddles in float32 (exact tables), n = 256 n1 + 16 n2 + n3 -> k = k1 + 16 k2 + 256 k3."""
    sgn = 1 if inv else -1
    a = z.reshape(16, 16, 16)                                                        # [n1, n2, n3]
    a = stage(a.reshape(16, 256), inv).reshape(16, 16, 16)                           # [k1, n2, n3]
    k1, n2, n3 = np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij")
    a = (a * np.exp(sgn * 2j * np.pi * k1 * (16 * n2 + n3) / 4096).astype(np.complex64)).astype(np.complex64)
    a = np.moveaxis(a, 1, 0)                                                         # [n2, k1, n3]
    a = stage(a.reshape(16, 256), inv).reshape(16, 16, 16)                           # [k2, k1, n3]
    k2, k1b, n3b = np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij")
    a = (a * np.exp(sgn * 2j * np.pi * k2 * n3b / 256).astype(np.complex64)).astype(np.complex64)
    a = np.moveaxis(a, 2, 0)                                                         # [n3, k2, k1]
    a = stage(a.reshape(16, 256), inv).reshape(16, 16, 16)                           # [k3, k2, k1]
    return a.reshape(4096)                                                           # k = 256 k3 + 16 k2 + k1


def exact_stage(z, inv):
    k = np.arange(16)
    return (np.exp((2j if inv else -2j) * np.pi * np.outer(k, k) / 16) @ z.astype(np.complex128)).astype(np.complex64)


def main():
    rng = np.random.default_rng(0)
    n, sr = 4096, 48000
    win = np.hanning(n)
    f = np.fft.fftfreq(n, 1 / sr)
    x = np.log2(np.maximum(np.abs(f), 1.0) / 1000.0)
    g = 10 ** (np.where(x < 0, np.minimum(12 * -x, 15), -np.minimum(12 * x, 15)) / 20)   # C1 tilt, +-15 dB platforms
    z = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n))                             # full-scale L + iR
    ref = np.fft.ifft(np.fft.fft(z * win) * g) * win
    zin = (z * win).astype(np.complex64)
    assert np.allclose(fft4096(zin, False, exact_stage), np.fft.fft(zin.astype(np.complex128)), atol=2e-2)
    P3 = [(0, 0), (0, 1), (1, 0)]
    cases = [("float32 butterflies (what stft_kernel does now)", exact_stage),
             ("tf32 x1 (plain TF32 GEMM)", lambda a, inv: tc_stage(a, inv, "tf32", 1, 1, [(0, 0)])),
             ("tf32 x3 (Fh*xh + Fh*xl + Fl*xh)", lambda a, inv: tc_stage(a, inv, "tf32", 2, 2, P3)),
             ("tf32 x4 (+ Fl*xl)", lambda a, inv: tc_stage(a, inv, "tf32", 2, 2, P3 + [(1, 1)])),
             ("fp16 x3, per-frame scale", lambda a, inv: tc_stage(a, inv, "fp16", 2, 2, P3)),
             ("bf16 x3", lambda a, inv: tc_stage(a, inv, "bf16", 2, 2, P3)),
             ("bf16 x6 (three-way split, products down to 2^-24)",
              lambda a, inv: tc_stage(a, inv, "bf16", 3, 3, [(0, 0), (0, 1), (1, 0), (0, 2), (1, 1), (2, 0)]))]
    print(f"{'variant':58s} {'max |err| of the round trip':>28s}   (bar: 1e-5 of full scale)")
    for name, st in cases:
        spec = fft4096(zin, False, st)
        y = fft4096((spec * g.astype(np.float32)).astype(np.complex64), True, st) / 4096
        err = np.abs(y * win - ref).max()
        print(f"{name:58s} {err:28.3e}   {'ok' if err < 1e-5 else 'FAILS'}")


if __name__ == "__main__":
    main()
