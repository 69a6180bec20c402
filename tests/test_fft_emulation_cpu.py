"""CPU: the FFT building blocks of the CUDA kernel (csrc/fft4096.cuh is __host__ __device__) emulated thread by thread
(csrc/host_emul.cu, a host-only build of the same stage functions) against numpy's FFT -- index maps, twiddles, exchange
layouts and the permuted gain row can be checked in the GPU-less build container."""
import ctypes as C

import numpy as np
import pytest

from tomatis_audio_processor_b200 import build


@pytest.fixture(scope="module")
def emul():
    lib = C.CDLL(build.build_emulation())
    lib.tmt_emul_forward.argtypes = [C.c_void_p, C.c_void_p]
    lib.tmt_emul_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def test_forward_fft_matches_numpy(emul):
    rng = np.random.default_rng(0)
    z = (rng.standard_normal(4096) + 1j * rng.standard_normal(4096)).astype(np.complex64)
    out = np.empty(4096, np.complex64)
    assert emul.tmt_emul_forward(z.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) == 0
    ref = np.fft.fft(z.astype(np.complex128))
    err = np.abs(out - ref).max() / np.abs(ref).max()
    assert err < 2e-6, err


def test_filter_operator_matches_numpy(emul):
    """out = IFFT(G * FFT(L + iR)) equals the per-channel rfft * g * irfft of the reference (process_tomatis.py:394-398)."""
    rng = np.random.default_rng(1)
    l, r = rng.standard_normal(4096).astype(np.float32), rng.standard_normal(4096).astype(np.float32)
    g = np.exp(rng.uniform(-1.5, 1.5, 2049)).astype(np.float32)
    z = (l + 1j * r).astype(np.complex64)
    out = np.empty(4096, np.complex64)
    assert emul.tmt_emul_filter(z.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) == 0
    yl = np.fft.irfft(np.fft.rfft(l.astype(np.float64)) * g, 4096)
    yr = np.fft.irfft(np.fft.rfft(r.astype(np.float64)) * g, 4096)
    scale = max(np.abs(yl).max(), np.abs(yr).max())
    assert np.abs(out.real - yl).max() / scale < 3e-6 and np.abs(out.imag - yr).max() / scale < 3e-6


# ---- pair mode (n_fft = 2048): two frames as the even / odd samples of one 4096-wide pass ------------------------------------
@pytest.fixture(scope="module")
def emul_pair(emul):
    emul.tmt_emul_forward_pair.argtypes = [C.c_void_p, C.c_void_p]
    emul.tmt_emul_filter_pair.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return emul


def test_pair_forward_is_two_2048_point_ffts(emul_pair):
    rng = np.random.default_rng(2)
    f = (rng.standard_normal((2, 2048)) + 1j * rng.standard_normal((2, 2048))).astype(np.complex64)
    z = np.empty(4096, np.complex64)
    z[0::2], z[1::2] = f[0], f[1]
    out = np.empty(4096, np.complex64)
    assert emul_pair.tmt_emul_forward_pair(z.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) == 0
    ref = np.fft.fft(f.astype(np.complex128), axis=1).reshape(-1)
    err = np.abs(out - ref).max() / np.abs(ref).max()
    assert err < 2e-6, err


def test_pair_filter_applies_each_frames_own_gain_row(emul_pair):
    """Each frame of the pair: IFFT(G_p * FFT(L_p + iR_p)) = the per-channel rfft * g_p * irfft of the reference with n_fft = 2048."""
    rng = np.random.default_rng(3)
    lr = rng.standard_normal((2, 2, 2048)).astype(np.float32)                 # [frame][channel][sample]
    g = np.exp(rng.uniform(-1.5, 1.5, (2, 1025))).astype(np.float32)
    z = np.empty(4096, np.complex64)
    z[0::2], z[1::2] = lr[0, 0] + 1j * lr[0, 1], lr[1, 0] + 1j * lr[1, 1]
    out = np.empty(4096, np.complex64)
    assert emul_pair.tmt_emul_filter_pair(z.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) == 0
    for p in range(2):
        y = out[p::2]
        for ch, got in ((0, y.real), (1, y.imag)):
            want = np.fft.irfft(np.fft.rfft(lr[p, ch].astype(np.float64)) * g[p], 2048)
            assert np.abs(got - want).max() / np.abs(want).max() < 3e-6
