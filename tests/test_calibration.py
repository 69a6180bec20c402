"""Calibration front end (src/calibrate_to_baseline_v2.py): oracle vs the executed reference's fixtures
(CPU), the kernels' per-thread code (csrc/calib.cuh, __host__ __device__) run on the CPU vs scipy / the oracle, the product's
host logic driven by those emulated kernels vs the fixtures (CPU), and the CUDA path vs fixtures and oracle (GPU).

Tolerances.  Frame levels, reference states, every mismatch / switch count of the grid and therefore the saved JSON: exact.
Decimated envelopes: scipy filters in float32, the kernel accumulates in double -> 2e-6 of the envelope's peak; the
correlation peak index (and the delay) exact.  Band tilt: the kernel transforms in fp64 and sums the band in double, NumPy in
float32 -> 1e-4 dB.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from helpers import cal_golden_names, cal_kwargs, load_cal_golden
from oracle import calibrate_oracle as co
from tomatis_audio_processor_b200 import build, calibrate_to_baseline_v2 as prod, tables as tb

TILT_TOL_DB = 1e-4
ENV_TOL = 2e-6


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def emul():
    lib = C.CDLL(build.build_emulation())
    lib.tmt_emul_decimate.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_void_p]
    lib.tmt_emul_power_levels.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p]
    lib.tmt_emul_band_energies.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.tmt_emul_gate_grid.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3
    return lib


class EmulatedKernels:
    """engine.calib_* with the same signatures, running csrc/calib.cuh's per-thread functions on the CPU (test infrastructure:
    lets the product's host logic be checked without a GPU).  The cross-correlation, whose kernel has no per-thread
    function to share, is NumPy's."""

    def __init__(self, lib):
        self.lib = lib

    def to_device(self, arrays, device=0):
        return [np.ascontiguousarray(a, dtype=np.float32) for a in arrays]

    def calib_envelope(self, x, lo, hi, up, down, device=0):
        x = np.ascontiguousarray(x[lo:hi], dtype=np.float32)
        plan = tb.resample_poly_plan(hi - lo, up, down)
        h = plan["h"] if plan["h"] is not None else np.ones(1, np.float32)
        out = np.empty(plan["n_out"], np.float32)
        assert self.lib.tmt_emul_decimate(_p(x), len(x), _p(h), len(h), plan["up"], plan["down"], plan["n_pre_remove"],
                                          plan["n_out"], _p(out)) == 0
        out = out - np.float32(out.astype(np.float64).sum() / max(1, out.size))
        return _HostTensor(out)

    def calib_xcorr(self, a, b):
        if a.arr.size == 0 or b.arr.size == 0:
            raise ValueError("cross-correlation of an empty envelope")
        if b.arr.size > a.arr.size:                      # scipy swaps the operands of a "valid" convolution
            return self.calib_xcorr(b, a)[::-1].copy()
        return np.correlate(a.arr.astype(np.float64), b.arr.astype(np.float64), mode="valid").astype(np.float32)

    def calib_frame_levels(self, x, device=0):
        x = np.ascontiguousarray(x, dtype=np.float32)
        mono = np.empty(len(x), np.float32)
        assert self.lib.tmt_emul_power_levels(_p(x), len(x), _p(mono)) == 0
        n = 1 + (len(x) - 4096) // 2048 if len(x) >= 4096 else 0
        msq = np.array([np.mean(mono[i * 2048:i * 2048 + 4096] * mono[i * 2048:i * 2048 + 4096]) for i in range(n)], np.float32)
        return tb.levels_from_meansq(msq).astype(np.float32)

    def calib_band_energies(self, x, n_frames, lo_bins, hi_bins, device=0):
        x = np.ascontiguousarray(x, dtype=np.float32)
        win = np.hanning(4096).astype(np.float32)
        e_lo, e_hi = np.zeros(n_frames, np.float32), np.zeros(n_frames, np.float32)
        a, b = C.c_float(), C.c_float()
        for i in range(n_frames):
            fr = np.ascontiguousarray(x[i * 2048:i * 2048 + 4096])
            assert self.lib.tmt_emul_band_energies(_p(fr), _p(win), lo_bins[0], lo_bins[1], hi_bins[0], hi_bins[1],
                                                   C.byref(a), C.byref(b)) == 0
            e_lo[i], e_hi[i] = a.value, b.value
        return e_lo, e_hi

    def calib_gate_grid(self, level, start, want, on, off, delay, want_states=False, device=0):
        level = np.ascontiguousarray(level, np.float32); start = np.ascontiguousarray(start, np.int64)
        want = np.ascontiguousarray(want, np.uint8); on = np.ascontiguousarray(on, np.float32)
        off = np.ascontiguousarray(off, np.float32); delay = np.ascontiguousarray(delay, np.int64)
        mis, sw = np.zeros(on.size, np.int32), np.zeros(on.size, np.int32)
        st = np.zeros((on.size, level.size), np.uint8) if want_states else None
        assert self.lib.tmt_emul_gate_grid(_p(level), _p(start), _p(want), level.size, _p(on), _p(off), _p(delay), on.size,
                                           _p(mis), _p(sw), _p(st) if want_states else None) == 0
        return (mis, sw, st) if want_states else (mis, sw)


class _HostTensor:
    def __init__(self, arr):
        self.arr = arr

    def numel(self):
        return self.arr.size


@pytest.fixture
def emulated_engine(emul, monkeypatch):
    from tomatis_audio_processor_b200 import engine
    k = EmulatedKernels(emul)
    for name in ("to_device", "calib_envelope", "calib_xcorr", "calib_frame_levels", "calib_band_energies", "calib_gate_grid"):
        monkeypatch.setattr(engine, name, getattr(k, name))
    return k


# ------------------------------------------------------------------------------------------------ oracle
@pytest.mark.parametrize("name", cal_golden_names())
def test_calibration_oracle_matches_reference_fixture(name):
    g = load_cal_golden(name)
    o = co.calibrate(g["orig"], g["base"], g["sr"], **cal_kwargs(g["words"]))
    assert o["json"] == g["json"]
    assert abs(g["json"]["delay_samples_orig_minus_base"] - g["true_delay"]) <= g["sr"] // 2000   # one decimated sample
    d = co.find_delay(g["orig"], g["base"], sr=g["sr"])
    import scipy
    exact = np.__version__ == g["numpy"] and scipy.__version__ == g["scipy"]
    for got, ref in ((d["mo_ds"], g["mo_ds"]), (d["mb_ds"], g["mb_ds"]), (o["orig_level"], g["orig_level"]),
                     (o["base_level"], g["base_level"]), (o["tilts"], g["tilts"])):
        assert np.array_equal(got, ref) if exact else np.allclose(got, ref, rtol=1e-5, atol=1e-6)
    assert d["k"] == g["k"]


# ------------------------------------------------------------------------------------------------ host tables / logic
@pytest.mark.parametrize("n_in,up,down", [(48000, 2000, 48000), (44100, 2000, 44100), (1000, 1, 24), (999, 3, 7), (25, 1, 24),
                                          (5000, 7, 3), (4800, 1, 1), (1, 1, 24)])
def test_emulated_decimation_matches_scipy(emul, n_in, up, down):
    """tables.resample_poly_plan + decimate_sample against scipy.signal.resample_poly on the float32 envelope."""
    from scipy.signal import resample_poly
    rng = np.random.default_rng(n_in + up)
    x = (0.3 * rng.standard_normal((n_in, 2))).astype(np.float32)
    env = co.power_mono(x)
    assert env.dtype == np.float32
    ref = resample_poly(env, up, down)
    plan = tb.resample_poly_plan(n_in, up, down)
    assert plan["n_out"] == len(ref)
    h = plan["h"] if plan["h"] is not None else np.ones(1, np.float32)
    out = np.empty(plan["n_out"], np.float32)
    assert emul.tmt_emul_decimate(_p(x), n_in, _p(h), len(h), plan["up"], plan["down"], plan["n_pre_remove"], plan["n_out"], _p(out)) == 0
    assert np.abs(out - ref).max() <= ENV_TOL * max(1.0, np.abs(ref).max())
    mono = np.empty(n_in, np.float32)
    emul.tmt_emul_power_levels(_p(x), n_in, _p(mono))
    assert np.array_equal(mono, env)


def test_median_filter_matches_scipy():
    from scipy.signal import medfilt
    rng = np.random.default_rng(4)
    for n in (1, 2, 3, 7, 100):
        x = rng.standard_normal(n).astype(np.float32)
        for k in (3, 5, 7):
            assert np.array_equal(tb.medfilt_zero_padded(x, k), medfilt(x, kernel_size=k))


def test_host_bookkeeping_matches_oracle():
    rng = np.random.default_rng(8)
    for n in (1, 2, 5, 40, 300):
        st = (1 + (np.cumsum(rng.random(n) < 0.3) % 2)).astype(np.int32)
        for min_run in (1, 2, 3, 4):
            assert np.array_equal(prod.debounce_state(st, min_run), co.debounce_state(st, min_run))
        x = np.concatenate([rng.normal(-8, 2, n), rng.normal(9, 3, n // 2)]).astype(np.float32)
        a, b = prod.kmeans2_1d(x), co.kmeans2_1d(x)
        assert np.array_equal(a[0], b[0]) and a[1:] == b[1:]
        mask = rng.random(x.size) < 0.8
        if mask.any():
            for k in (4, 5, 1):
                sa, ta = prod.baseline_states(x, mask, k)
                sb, tb_ = co.baseline_states(x, mask, k)
                assert np.array_equal(sa, sb) and np.array_equal(ta, tb_)
    # one cluster only: every frame keeps C1 whichever label it got
    x = np.full(30, -3.0, np.float32)
    assert np.array_equal(prod.baseline_states(x, np.ones(30, bool))[0], co.baseline_states(x, np.ones(30, bool))[0])
    assert prod.band_bins(48000, 4096, (200, 1000)) == (18, 86) and prod.band_bins(48000, 4096, (30000, 40000)) == (0, 0)


def test_emulated_gate_grid_matches_oracle(emul):
    """gate_grid_combo against simulate_state on irregular frame positions, thresholds in float32."""
    rng = np.random.default_rng(12)
    k = EmulatedKernels(emul)
    for n in (0, 1, 7, 400):
        level = rng.uniform(-60, -20, n).astype(np.float32)
        start = np.cumsum(rng.integers(1, 4, n) * 2048).astype(np.int64)
        want = rng.integers(1, 3, n).astype(np.uint8)
        Ts = rng.uniform(-50, -30, 12).astype(np.float32).astype(np.float64)
        hyst = rng.choice([0.0, 1.0, 3.0, 6.0], 12)
        up_ms = rng.choice([0.0, 50.0, 130.0, 250.0], 12)
        if n:
            Ts[0], hyst[0] = float(level[n // 2]), 0.0                       # a level exactly on both thresholds
        on, off = (Ts + hyst / 2).astype(np.float32), (Ts - hyst / 2).astype(np.float32)
        delay = np.array([int(round(48000 * v / 1000.0)) for v in up_ms], np.int64)
        mis, sw, st = k.calib_gate_grid(level, start, want, on, off, delay, want_states=True)
        for c in range(12):
            ref = co.simulate_state(level, start, 48000, float(Ts[c]), float(hyst[c]), float(up_ms[c]))
            assert np.array_equal(st[c], ref)
            assert mis[c] == int((ref != want).sum()) and sw[c] == int((ref[1:] != ref[:-1]).sum())


@pytest.mark.parametrize("name", cal_golden_names())
def test_emulated_band_energies_match_reference_tilts(emul, name):
    g = load_cal_golden(name)
    k = EmulatedKernels(emul)
    d = g["json"]["delay_samples_orig_minus_base"]
    xb = g["base"][max(0, -d):]
    n = min(40, len(g["tilts"]))
    e_lo, e_hi = k.calib_band_energies(xb, n, prod.band_bins(g["sr"], 4096, g["tilt_lo"]), prod.band_bins(g["sr"], 4096, g["tilt_hi"]))
    assert np.abs(prod.tilt_from_energies(e_lo, e_hi) - g["tilts"][:n]).max() < TILT_TOL_DB


# ------------------------------------------------------------------------------------------------ product host logic
@pytest.mark.parametrize("name", cal_golden_names())
def test_front_end_on_emulated_kernels_matches_reference_fixture(emulated_engine, name):
    """calibrate() with its kernels emulated on the CPU: the JSON the reference saved, exactly."""
    g = load_cal_golden(name)
    r = prod.calibrate(g["orig"], g["base"], g["sr"], **cal_kwargs(g["words"]))
    assert r["json"] == g["json"]
    assert np.array_equal(r["orig_level"], g["orig_level"]) and np.array_equal(r["base_level"], g["base_level"])
    assert np.abs(r["tilts"] - g["tilts"]).max() < TILT_TOL_DB
    d = prod.find_delay(g["orig"], g["base"], sr=g["sr"])
    assert d["k"] == g["k"]
    assert np.abs(d["mo_ds"].arr - g["mo_ds"]).max() <= ENV_TOL * np.abs(g["mo_ds"]).max()
    assert np.abs(d["corr"] - g["corr"]).max() <= 1e-4 * np.abs(g["corr"]).max()


def test_front_end_grid_table_matches_oracle(emulated_engine):
    """Every (gain, delay, hysteresis, threshold) combination's mismatch and switch counts, in the reference's loop order."""
    g = load_cal_golden("cal_48k_default")
    o = co.calibrate(g["orig"], g["base"], g["sr"], **cal_kwargs(g["words"]))
    args = (o["orig_level"], o["base_level"], o["base_state"], o["starts"], o["music_mask"], g["sr"], (0, 3), (0, 100.0))
    kw = dict(gain_search_pm_db=1.0, gain_step_db=0.5, T_pm_db=4.0, T_step_db=0.5, want_table=True)
    best_o, g0_o, table_o = co.grid_search(*args, **kw)
    best_p, g0_p, table_p = prod.grid_search(*args, **kw)
    assert best_p == best_o and g0_p == g0_o and len(table_p) > 100
    assert [tuple(t) for t in table_p] == [tuple(t) for t in table_o]
    st = prod.simulate_state(o["orig_level"], o["starts"], g["sr"], -40.0, 3.0, 100.0)
    assert st.dtype == np.int32 and np.array_equal(st, co.simulate_state(o["orig_level"], o["starts"], g["sr"], -40.0, 3.0, 100.0))


def test_front_end_errors(emulated_engine):
    g = load_cal_golden("cal_44k1_bands")
    with pytest.raises(ValueError):
        prod.calibrate(g["orig"][:60000], g["base"][:3000], g["sr"])                  # overlap <= n_fft
    with pytest.raises(NotImplementedError):
        prod.calibrate(g["orig"], g["base"], g["sr"], n_fft=2048, hop=1024)
    with pytest.raises(RuntimeError):
        prod.calibrate(g["orig"][:90000], g["base"][:40000], g["sr"])                 # under 10 frames per state -> no optimum
    with pytest.raises(IndexError):
        prod.calibrate(g["orig"], g["base"], g["sr"], music_dbfs=0.0)                 # no music frame: np.percentile of nothing,
    with pytest.raises(IndexError):                                                   # as in the reference (:33)
        co.calibrate(g["orig"], g["base"], g["sr"], music_dbfs=0.0)
    with pytest.raises(AssertionError):
        prod.calibrate(g["orig"][:, :1], g["base"], g["sr"])


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", cal_golden_names())
def test_calibration_gpu_matches_reference_fixture(name):
    g = load_cal_golden(name)
    r = prod.calibrate(g["orig"], g["base"], g["sr"], **cal_kwargs(g["words"]))
    assert r["json"] == g["json"]
    assert np.array_equal(r["orig_level"], g["orig_level"]) and np.array_equal(r["base_level"], g["base_level"])
    tilt_err = float(np.abs(r["tilts"] - g["tilts"]).max())
    d = prod.find_delay(g["orig"], g["base"], sr=g["sr"])
    mo, mb = d["mo_ds"].cpu().numpy(), d["mb_ds"].cpu().numpy()
    env_err = max(float(np.abs(mo - g["mo_ds"]).max() / np.abs(g["mo_ds"]).max()), float(np.abs(mb - g["mb_ds"]).max() / np.abs(g["mb_ds"]).max()))
    corr_err = float(np.abs(d["corr"] - g["corr"]).max() / np.abs(g["corr"]).max())
    print(f"{name}: JSON exact, levels exact, tilt max |dB| error {tilt_err:.2e}, envelope {env_err:.2e}, correlation {corr_err:.2e} (relative to peak)")
    assert d["k"] == g["k"] and tilt_err < TILT_TOL_DB and env_err <= ENV_TOL and corr_err <= 1e-4


@pytest.mark.gpu
def test_calibration_gpu_kernels_against_numpy_and_emulation(emul):
    import torch
    from tomatis_audio_processor_b200 import engine
    rng = np.random.default_rng(21)
    k = EmulatedKernels(emul)
    # cross-correlation: tile / lag-block edges, one lag, long b
    for na, nb in ((5000, 5000), (5000, 4999), (3000, 1), (70000, 2048), (70000, 2049), (9000, 4100), (1025, 3)):
        a, b = rng.standard_normal(na).astype(np.float32), rng.standard_normal(nb).astype(np.float32)
        got = engine.calib_xcorr(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
        ref = np.correlate(a.astype(np.float64), b.astype(np.float64), mode="valid")
        assert got.shape == ref.shape and np.abs(got - ref).max() <= 2e-5 * np.sqrt(nb) * 4
    from scipy.signal import fftconvolve
    a, b = rng.standard_normal(700).astype(np.float32), rng.standard_normal(1900).astype(np.float32)
    got = engine.calib_xcorr(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())     # second operand longer: scipy's swap
    ref = fftconvolve(a.astype(np.float64), b[::-1].astype(np.float64), mode="valid")
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-3
    with pytest.raises(ValueError):
        engine.calib_xcorr(torch.zeros(10).cuda(), torch.zeros(0).cuda())
    # envelope + decimation against the emulated per-thread code (same arithmetic) and scipy
    from scipy.signal import resample_poly
    for n_in, up, down in ((48000, 2000, 48000), (44100, 2000, 44100), (999, 3, 7), (25, 1, 24), (4800, 1, 1)):
        x = (0.3 * rng.standard_normal((n_in + 100, 2))).astype(np.float32)
        got = engine.calib_envelope(x, 50, 50 + n_in, up, down).cpu().numpy()
        ref = resample_poly(co.power_mono(x[50:50 + n_in]), up, down).astype(np.float32)
        ref = ref - np.mean(ref)
        assert got.shape == ref.shape and np.abs(got - ref).max() <= ENV_TOL * max(1.0, np.abs(ref).max())
    with pytest.raises(ValueError):
        engine.calib_envelope(x, 10, len(x) + 1, 1, 24)
    # gate grid: bit-identical to the emulated per-thread code, states included
    n = 3000
    level = rng.uniform(-60, -20, n).astype(np.float32)
    start = np.cumsum(rng.integers(1, 4, n) * 2048).astype(np.int64)
    want = rng.integers(1, 3, n).astype(np.uint8)
    on = rng.uniform(-45, -30, 700).astype(np.float32)
    off = (on - rng.choice([0.0, 1.0, 3.0], 700)).astype(np.float32)
    delay = rng.choice([0, 2400, 4800, 12000], 700).astype(np.int64)
    got, ref = engine.calib_gate_grid(level, start, want, on, off, delay, want_states=True), k.calib_gate_grid(level, start, want, on, off, delay, want_states=True)
    assert all(np.array_equal(a, b) for a, b in zip(got, ref))
    mis, sw = engine.calib_gate_grid(level[:0], start[:0], want[:0], on, off, delay)
    assert not mis.any() and not sw.any()
    # frame levels / band energies on ragged lengths: levels exact against NumPy, energies against the emulation
    for total in (4096, 4097, 6143, 6144, 30000):
        x = (0.2 * rng.standard_normal((total, 2))).astype(np.float32)
        lv = engine.calib_frame_levels(x)
        assert np.array_equal(lv, k.calib_frame_levels(x))
        ge, ke = engine.calib_band_energies(x, len(lv), (18, 86), (171, 683)), k.calib_band_energies(x, len(lv), (18, 86), (171, 683))
        assert np.allclose(ge[0], ke[0], rtol=2e-6) and np.allclose(ge[1], ke[1], rtol=2e-6)
    assert engine.calib_frame_levels(x[:4095]).size == 0
    with pytest.raises(RuntimeError):
        engine.calib_band_energies(x[:5000], 2, (18, 86), (171, 683))                    # second frame past the end


@pytest.mark.gpu
def test_calibration_gpu_cli_files(tmp_path):
    """The command line on WAV files: the JSON file the reference saves, key for key."""
    from tomatis_audio_processor_b200 import audio_io
    g = load_cal_golden("cal_48k_default")
    po, pb, pj = (str(tmp_path / n) for n in ("orig.wav", "base.wav", "cal.json"))
    audio_io.write(po, g["orig"], g["sr"], subtype="PCM_16")
    audio_io.write(pb, g["base"], g["sr"], subtype="PCM_16")
    assert prod.main(["--orig", po, "--base", pb, "--out_json", pj] + [str(w) for w in g["words"]]) == 0
    with open(pj, encoding="utf-8") as f:
        saved = json.load(f)
    assert list(saved)[:2] == ["orig", "base"] and saved["orig"] == po
    assert {k: v for k, v in saved.items() if k not in ("orig", "base")} == g["json"]
    assert list(saved)[2:] == list(g["json"])
    assert prod.find_delay_by_corr(po, pb, sr=g["sr"]) == g["json"]["delay_samples_orig_minus_base"]
    assert os.path.getsize(pj) > 0


def _pair(sr, secs, base_secs, seed, lead, up_ms=100.0):
    """orig = x[lead:], base = the oracle's standard-mode output of the first base_secs of x: the baseline starts `lead`
    samples EARLIER than the original (negative delay)."""
    from oracle import tomatis_oracle as orc
    from tomatis_audio_processor_b200 import synth
    x = synth.pcm16_to_float(synth.quantise_pcm16(synth.recipe_level_steps(secs, sr, seed, min_s=0.25, max_s=0.8)))
    y = orc.run("standard", x[:int(base_secs * sr)], sr, gate_ui=50, up_delay_ms=up_ms)["out"].astype(np.float32)
    return np.ascontiguousarray(x[lead:]), synth.pcm16_to_float(synth.quantise_pcm16(0.9 * y))


@pytest.mark.parametrize("sr,lead,kw", [(48000, 30007, dict(hyst_list=(2, 3), delay_list_ms=(50, 100))),
                                        (44100, 12345, dict(hyst_list=(3,), delay_list_ms=(100,), max_minutes=0.08, tilt_medfilt=7))])
def test_front_end_negative_delay_matches_oracle(emulated_engine, sr, lead, kw):
    """The baseline leads the original (delay < 0: base_start = -delay, orig_start = 0, :168-169), short overlap cap.  The
    baseline has to be longer than the 25 s chunk for that: its middle chunk must lie inside the original."""
    orig, base = _pair(sr, 31.0, 29.0, 77, lead)
    o = co.calibrate(orig, base, sr, **kw)
    r = prod.calibrate(orig, base, sr, **kw)
    assert o["delay"] < 0 and abs(o["delay"] + lead) <= sr // 2000 + 1
    assert r["json"] == o["json"] and r["delay"] == o["delay"]
    assert np.array_equal(r["orig_level"], o["orig_level"]) and np.array_equal(r["base_level"], o["base_level"])
    assert np.array_equal(r["base_state"], o["base_state"]) and np.array_equal(r["music_mask"], o["music_mask"])
    assert np.abs(r["tilts"] - o["tilts"]).max() < TILT_TOL_DB


def test_front_end_baseline_chunk_longer_than_original(emulated_engine):
    """scipy's "valid" convolution swaps its operands when the second is the longer one, so the reference does not fail on
    an original shorter than the baseline chunk (:77-78): same (meaningless) delay, same JSON."""
    orig, base = _pair(48000, 8.0, 8.0, 78, 30007)
    assert len(base) > len(orig)
    kw = dict(hyst_list=(3,), delay_list_ms=(100,))
    o, r = co.calibrate(orig, base, 48000, **kw), prod.calibrate(orig, base, 48000, **kw)
    assert r["delay"] == o["delay"] and r["json"] == o["json"]
