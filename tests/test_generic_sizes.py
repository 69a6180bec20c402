"""General FFT sizes (generic.py + csrc/generic.cuh).  CPU: frame geometry and limiter chunks
against the oracle, and the whole path for n_fft / hop other than 4096 / 2048 with the kernels' per-thread code run on the CPU
(csrc/host_emul.cu) against the oracle, which tests/test_oracle_vs_reference.py pins to the executed reference for these
sizes -- mean squares bit-exact, states / rows / chunk lengths exact, PCM within 1e-5 of the float64-FFT evaluation everywhere
and of the float32-FFT reference wherever that is well-conditioned.  GPU: the same through the C ABI (engine.run with
n_fft / hop keywords, the way the command lines call it)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import tomatis_oracle as orc
from tomatis_audio_processor_b200 import _lib as L, build, generic, synth, tables as tb

SIZES = [(2048, 1024), (1024, 512), (4096, 1024), (8192, 4096), (2048, 512), (512, 384), (256, 256)]
PCM_TOL = 1e-5


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def emul():
    lib = C.CDLL(build.build_emulation())
    ll, i, f, p = C.c_longlong, C.c_int, C.c_float, C.c_void_p
    lib.tmt_emul_gen_meansq.argtypes = [p, ll, ll, i, i, i, i, f, i, p]
    lib.tmt_emul_gen_frames.argtypes = [p, ll, ll, i, i, i, p, p, p, f, i, p]
    lib.tmt_emul_gen_ola.argtypes = [p, i, ll, ll, i, i, i, p, f, p]
    return lib


class EmulatedGenericKernels:
    """generic.CudaKernels' interface on the CPU: csrc/generic.cuh's per-thread functions through host_emul.cu; the gate
    automata (device code of the main path, not part of generic.cuh) are restated here in Python.  Test infrastructure."""

    def __init__(self, lib):
        self.lib, self.launches = lib, 0

    def upload(self, x):
        return np.ascontiguousarray(x, np.float32)

    def upload_tables(self, win, gains):
        return np.ascontiguousarray(win, np.float32), np.ascontiguousarray(gains, np.float32)

    def input_peak(self, xd):
        return np.float32(np.max(np.abs(xd)))

    def meansq(self, xd, total, first, n_fft, hop, n_frames, use_f64, sc, mono):
        out = np.zeros(n_frames, np.float64 if use_f64 else np.float32)
        assert self.lib.tmt_emul_gen_meansq(_p(xd), total, first, n_fft, hop, n_frames, int(use_f64), float(sc), int(mono), _p(out)) == 0
        return out

    def gate(self, automaton, values, on, off, param, xfade_frames, init_to_target=False, count_only=False):
        n = len(values)
        states = np.zeros(n, np.uint8)
        state, run, since = 1, 0, param
        for i, v in enumerate(values):
            hi, lo = v >= on, v <= off
            if automaton == L.GATE_UPDELAY:                      # param-th consecutive "hi" frame switches up
                if state == 1:
                    run = run + 1 if hi else 0
                    if run >= param:
                        state, run = 2, 0
                elif lo:
                    state, run = 1, 0
            else:                                                # min-hold (src/process_tomatis_adaptive.py:87-121)
                since += 1
                if since >= param:
                    if state == 1 and hi:
                        state, since = 2, 0
                    elif state == 2 and lo:
                        state, since = 1, 0
            states[i] = state
        if count_only:
            return int((states == 2).sum())
        X = max(int(xfade_frames), 1)
        rows, k = np.zeros(n, np.uint16), 0
        for i, s in enumerate(states):
            target = 0 if s == 1 else X
            if (init_to_target and i == 0) or xfade_frames <= 0 or abs(target - k) <= 1:
                k = target
            else:
                k += 1 if target > k else -1
            rows[i] = k
        return states, rows

    def frames(self, xd, total, first, n_fft, hop, n_frames, win, gains, rows, sc, flavour):
        fr = np.zeros((max(1, n_frames), n_fft, 2), np.float64 if flavour == generic.ADAPTIVE_F64 else np.float32)
        rows = np.ascontiguousarray(rows, np.uint16)
        assert self.lib.tmt_emul_gen_frames(_p(xd), total, first, n_fft, hop, n_frames, _p(win), _p(gains), _p(rows), float(sc), flavour, _p(fr)) == 0
        return fr

    def overlap_add(self, fr, flavour, total, first, n_fft, hop, n_frames, win, post):
        y = np.zeros((total, 2), np.float64 if flavour == generic.ADAPTIVE_F64 else np.float32)
        assert self.lib.tmt_emul_gen_ola(_p(fr), flavour, total, first, n_fft, hop, n_frames, _p(win), float(post), _p(y)) == 0
        return y

    def limit(self, y, use_f64, bounds):
        T = np.float64 if use_f64 else np.float32
        peaks = np.zeros(len(bounds), T)
        for c, (a, b) in enumerate(bounds):
            peaks[c] = np.abs(y[a:b]).max() if b > a else 0
            if peaks[c] > T(tb.PEAK_LIMIT):
                y[a:b] *= T(tb.PEAK_LIMIT) / peaks[c]
        return peaks

    def to_host(self, y, use_f64):
        return y.astype(np.float32)


def _q(x):
    return synth.pcm16_to_float(synth.quantise_pcm16(x))


@pytest.mark.parametrize("n_fft,hop", SIZES + [(4096, 2048), (4096, 1536)])
def test_geometry_matches_oracle(n_fft, hop):
    for total in (0, 1, hop - 1, hop, n_fft // 2, n_fft - 1, n_fft, n_fft + 1, 3 * n_fft + 7, 250000, 500001):
        if total < 0:
            continue
        pad, pad_end, starts = orc.frame_layout_streaming(total, n_fft, hop)
        first, nf = generic.streaming_layout(total, n_fft, hop)
        assert nf == len(starts) and (nf == 0 or first == starts[0])
        want = []
        for a, b in orc.flush_schedule(nf, n_fft, hop):
            s, e = max(0, a), min(total, b)
            if e > s:
                want.append((s, e))
        assert generic.flush_sample_ranges(nf, total, n_fft, hop) == want
        # adaptive: the frames compute_frame_levels keeps
        x = np.zeros((total, 2), np.float32)
        levels = orc.compute_frame_levels(x, 48000, n_fft, hop)[0] if total else np.zeros(0)
        first_a, nf_a = generic.adaptive_layout(total, n_fft, hop)
        assert nf_a == len(levels)
        assert first_a == -(-(n_fft // 2) // hop) * hop - n_fft // 2 and 0 <= first_a < hop


def _compare(mode, o, o64, r):
    assert r["meansq"].dtype == np.asarray(o["meansq"]).dtype and np.array_equal(r["meansq"], np.asarray(o["meansq"]))
    assert np.array_equal(r["levels"], np.asarray(o["levels"], np.float64))
    assert np.array_equal(r["states"], o["states"])
    assert np.allclose(r["rows"] / max(r["xfade_frames"], 1), o["alphas"], atol=1e-9)
    assert r["chunk_lengths"] == o["chunk_lengths"]
    y, ref, ref64 = r["out"].astype(np.float64), o["out"].astype(np.float64), o64["out"].astype(np.float64)
    assert y.shape == ref.shape
    assert np.abs(y - ref64).max(initial=0.0) <= PCM_TOL                              # everywhere vs the float64-FFT evaluation
    self_noise = np.abs(ref - ref64).max(axis=1, initial=0.0)
    d = np.abs(y - ref).max(axis=1, initial=0.0)
    assert np.all(d <= PCM_TOL + self_noise)                                          # vs the float32-FFT reference, pointwise
    if mode == "adaptive":
        assert r["optimal_T"] == o["optimal_T"] and r["trace"] == o["trace"] and r["pipeline_dtype"] == o["pipeline_dtype"]


@pytest.mark.parametrize("n_fft,hop", SIZES)
def test_emulated_general_path_matches_oracle(emul, n_fft, hop):
    k = EmulatedGenericKernels(emul)
    sr = 48000
    secs = 0.6 if n_fft <= 1024 else 1.3
    xs = _q(synth.recipe_gated_pink(secs, sr, 90, env_hz=4.0, hi_dbfs=-18.0))
    for mode, kw in (("standard", dict(gate_ui=50, up_delay_ms=30.0, output_gain_db=-2.0)), ("xfade", dict(gate_ui=62, xfade_ms=60.0, up_delay_ms=20.0))):
        r = generic.run_streaming(mode, [xs], sr, kernels=k, n_fft=n_fft, hop=hop, **kw)[0]
        _compare(mode, orc.run(mode, xs, sr, n_fft=n_fft, hop=hop, **kw), orc.run(mode, xs, sr, n_fft=n_fft, hop=hop, fft_dtype="float64", **kw), r)
    for peak in (0.5, 0.08):                                                          # float32 and float64 pipelines
        xa = _q(synth.recipe_swept_pink(secs, sr, 91, period_s=0.3, peak=peak))
        kw = dict(min_hold_ms=40.0, xfade_ms=80.0, n_fft=n_fft, hop=hop)
        r = generic.run_adaptive([xa], sr, kernels=k, **kw)[0]
        o = orc.run("adaptive", xa, sr, **kw)
        assert r["pipeline_dtype"] == ("float32" if peak == 0.5 else "float64")
        _compare("adaptive", o, orc.run("adaptive", xa, sr, fft_dtype="float64", **kw), r)


def test_general_path_is_on_by_default_and_checks_sizes(monkeypatch):
    monkeypatch.delenv("TMT_GENERIC_FFT", raising=False)
    assert generic.enabled()
    monkeypatch.setenv("TMT_GENERIC_FFT", "0")
    assert not generic.enabled()
    for bad in ((100, 50), (16384, 8192), (3000, 1500), (2048, 0), (2048, 4096)):
        with pytest.raises(NotImplementedError):
            generic.check_sizes(*bad)
    generic.check_sizes(2048, 1024)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("n_fft,hop", SIZES)
def test_general_path_gpu_matches_oracle(n_fft, hop):
    from tomatis_audio_processor_b200 import engine
    sr = 48000
    xs = _q(synth.recipe_gated_pink(1.5, sr, 90, env_hz=4.0, hi_dbfs=-18.0))
    for mode, kw in (("standard", dict(gate_ui=50, up_delay_ms=30.0, output_gain_db=-2.0)), ("xfade", dict(gate_ui=62, xfade_ms=60.0, up_delay_ms=20.0))):
        r = engine.run(mode, [xs], sr, n_fft=n_fft, hop=hop, **kw)[0]
        _compare(mode, orc.run(mode, xs, sr, n_fft=n_fft, hop=hop, **kw), orc.run(mode, xs, sr, n_fft=n_fft, hop=hop, fft_dtype="float64", **kw), r)
    for peak in (0.5, 0.08):
        xa = _q(synth.recipe_swept_pink(1.5, sr, 91, period_s=0.3, peak=peak))
        kw = dict(min_hold_ms=40.0, xfade_ms=80.0, n_fft=n_fft, hop=hop)
        r = engine.run("adaptive", [xa], sr, **kw)[0]
        _compare("adaptive", orc.run("adaptive", xa, sr, **kw), orc.run("adaptive", xa, sr, fft_dtype="float64", **kw), r)


def test_emulated_general_path_random_cases(emul):
    """Seeded sweep: sizes down to 128, hops that do not divide n_fft, three sample rates, files shorter than a frame."""
    k = EmulatedGenericKernels(emul)
    rng = np.random.default_rng(7)
    for case in range(18):
        n_fft = int(rng.choice([128, 256, 512, 1024]))
        hop = int(rng.choice([n_fft, n_fft // 2, n_fft // 4, max(1, n_fft // 3), (3 * n_fft) // 4]))
        sr = int(rng.choice([44100, 48000, 96000]))
        total = int(rng.choice([1, n_fft // 2, n_fft - 1, n_fft + 1, int(rng.integers(2 * n_fft, 12000))]))
        env = 10.0 ** (rng.uniform(-3.0, -0.5, size=(total // 512 + 1)).repeat(512)[:total, None])
        x = _q((env * rng.standard_normal((total, 2))).astype(np.float32).clip(-1, 1))
        mode = ("standard", "xfade", "adaptive")[case % 3]
        if mode == "adaptive":
            kw = dict(min_hold_ms=float(rng.choice([0, 10, 40])), xfade_ms=float(rng.choice([0, 30, 80])), n_fft=n_fft, hop=hop)
            r = generic.run_adaptive([x], sr, kernels=k, **kw)[0]
        else:
            kw = dict(gate_ui=float(rng.uniform(40, 60)), up_delay_ms=float(rng.choice([0, 5, 30])), n_fft=n_fft, hop=hop)
            if mode == "xfade":
                kw["xfade_ms"] = float(rng.choice([0, 20, 60]))
            r = generic.run_streaming(mode, [x], sr, kernels=k, **kw)[0]
        o, o64 = orc.run(mode, x, sr, **kw), orc.run(mode, x, sr, fft_dtype="float64", **kw)
        if mode == "adaptive" and len(o["states"]) == 0:                     # no frame: the threshold is NaN on both sides
            assert np.isnan(r["optimal_T"]) and np.isnan(o["optimal_T"]) and np.array_equal(r["out"], o["out"].astype(np.float32))
            continue
        _compare(mode, o, o64, r)
