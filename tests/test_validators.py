"""Validator kernels (src/validate_layer1.py:110-163,244-389, src/verify_tomatis_15db_v2.py:254-369):
oracle vs the executed reference's fixtures (CPU), the spectrum kernel's host/device code emulated on the CPU vs the oracle,
CUDA path vs fixtures and oracle (GPU).

Tolerances.  Gate re-simulation: levels and states exact.  Conditional spectrum: the kernel transforms in fp64 and rounds
once to float32 like NumPy's rfft of float32 data, so ~60 % of the frame ratios are bit-identical and the rest 1 ulp off;
the medians in dB then differ by a few 1e-6 dB -> bar 1e-4 dB on every bin (the reference's own pass criteria are RMSE
thresholds of ~1 dB).  Frame counts exact."""
import ctypes as C

import numpy as np
import pytest

from helpers import load_val_golden, val_golden_names
from oracle import validate_oracle as vo
from tomatis_audio_processor_b200 import build, validators as prod

DB_TOL = 1e-4


def _v2_kwargs(g):
    return {k: (tuple(v) if k == "anchor_band" else v) for k, v in g["kwargs"].items() if k in ("level_percentile", "anchor_band")}


def _gate_kwargs(g):
    return {k: g["kwargs"][k] for k in ("threshold_dbfs", "hyst_db", "up_delay_ms")}


@pytest.mark.parametrize("name", val_golden_names())
def test_validator_oracle_matches_reference_fixture(name):
    g = load_val_golden(name)
    st, lv = vo.simulate_gate(g["x"], g["sr"], 4096, 2048, **_gate_kwargs(g))
    assert st == g["states"] and np.array_equal(np.array(lv), g["levels"])
    f, c1, c2, n1, n2, _ = vo.compute_conditional_spectrum(g["x"], g["y"], g["sr"], st, 4096, 2048)
    assert (n1, n2) == (g["n_c1"], g["n_c2"]) and min(n1, n2) > 10
    f, a1, a2, m1, m2, _ = vo.compute_conditional_spectrum_v2(g["x"], g["y"], g["sr"], st, np.array(lv), 4096, 2048, **_v2_kwargs(g))
    assert (m1, m2) == (g["v2_n_c1"], g["v2_n_c2"])
    exact = np.__version__ == g["numpy"]
    for got, ref in ((c1, g["c1_db"]), (c2, g["c2_db"]), (a1, g["v2_c1_db"]), (a2, g["v2_c2_db"])):
        assert np.array_equal(got, ref) if exact else np.abs(got - ref).max() < 1e-4


def test_stable_frames_match_oracle():
    rng = np.random.default_rng(3)
    for n in (0, 3, 4, 5, 6, 40, 500):
        st = ["C1" if v else "C2" for v in (np.cumsum(rng.random(n) < 0.15) % 2 == 0)]
        for margin in (1, 2, 3):
            assert prod.find_stable_frames(st, margin) == tuple(vo.find_stable_frames(st, margin))
        codes = np.array([1 if s == "C1" else 2 for s in st], dtype=np.uint8)
        assert prod.find_stable_frames(codes, 2) == tuple(vo.find_stable_frames(st, 2))


@pytest.fixture(scope="module")
def emul():
    lib = C.CDLL(build.build_emulation())
    lib.tmt_emul_spectrum_ratio.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int, C.c_void_p]
    return lib


@pytest.mark.parametrize("name", val_golden_names())
def test_emulated_spectrum_kernel_matches_oracle(emul, name):
    """spectrum_ratio_kernel's per-thread code (csrc/spectrum.cuh + fft4096.cuh, __host__ __device__) run thread by thread
    on the CPU: frame ratios, anchor normalisation and the resulting median curves against the oracle."""
    g = load_val_golden(name)
    x, y, sr = g["x"], g["y"], g["sr"]
    win = np.hanning(4096).astype(np.float32)
    freqs = np.fft.rfftfreq(4096, 1 / sr)
    band = _v2_kwargs(g).get("anchor_band", (900, 1100))
    am = np.nonzero((freqs >= band[0]) & (freqs <= band[1]))[0]

    def ratio(i, a0=0, a1=-1):
        fx, fy = np.ascontiguousarray(x[i * 2048:i * 2048 + 4096]), np.ascontiguousarray(y[i * 2048:i * 2048 + 4096])
        out = np.empty(2049, np.float32)
        assert emul.tmt_emul_spectrum_ratio(fx.ctypes.data_as(C.c_void_p), fy.ctypes.data_as(C.c_void_p),
                                            win.ctypes.data_as(C.c_void_p), a0, a1, out.ctypes.data_as(C.c_void_p)) == 0
        return out

    _, c1, c2, _, _, used = vo.compute_conditional_spectrum(x, y, sr, g["states"], 4096, 2048)
    _, a1, a2, _, _, used2 = vo.compute_conditional_spectrum_v2(x, y, sr, g["states"], g["levels"], 4096, 2048, **_v2_kwargs(g))
    for sel, ref in ((used[0], c1), (used[1], c2)):
        db = 20 * np.log10(np.median(np.array([ratio(i) for i in sel]), axis=0) + 1e-12)
        assert np.abs(db - ref).max() < DB_TOL
    for sel, ref in ((used2[0], a1), (used2[1], a2)):
        db = 20 * np.log10(np.median(np.array([ratio(i, int(am[0]), int(am[-1])) for i in sel]), axis=0) + 1e-12)
        assert np.abs(db - ref).max() < DB_TOL


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", val_golden_names())
def test_validators_gpu_match_reference_fixture(name):
    g = load_val_golden(name)
    x, y, sr = g["x"], g["y"], g["sr"]
    st, lv = prod.simulate_gate(x, sr, 4096, 2048, **_gate_kwargs(g))
    assert st == g["states"] and np.array_equal(np.array(lv), g["levels"])
    f, c1, c2, n1, n2 = prod.compute_conditional_spectrum(x, y, sr, st, 4096, 2048)
    assert (n1, n2) == (g["n_c1"], g["n_c2"]) and np.array_equal(f, np.fft.rfftfreq(4096, 1 / sr))
    f, a1, a2, m1, m2 = prod.compute_conditional_spectrum_v2(x, y, sr, st, np.array(lv), 4096, 2048, **_v2_kwargs(g))
    assert (m1, m2) == (g["v2_n_c1"], g["v2_n_c2"])
    errs = [float(np.abs(a - b).max()) for a, b in ((c1, g["c1_db"]), (c2, g["c2_db"]), (a1, g["v2_c1_db"]), (a2, g["v2_c2_db"]))]
    print(f"{name}: max |dB| error vs the reference c1 {errs[0]:.2e} c2 {errs[1]:.2e} anchored c1 {errs[2]:.2e} c2 {errs[3]:.2e}")
    assert max(errs) < DB_TOL


@pytest.mark.gpu
def test_validators_gpu_edge_cases():
    """Mono input, odd and even frame counts in the median, nothing selected, level threshold excluding everything,
    a frame list that reaches the end of the file, an output file that is too short."""
    from tomatis_audio_processor_b200 import engine, synth
    sr = 48000
    x = synth.recipe_gated_pink(3.0, sr, 61, env_hz=1.0, hi_dbfs=-26.0)
    rng = np.random.default_rng(9)
    y = (x * 0.5 + 0.01 * rng.standard_normal(x.shape)).astype(np.float32)
    # mono: levels / states exact, spectrum within tolerance
    st, lv = prod.simulate_gate(x[:, 0], sr, 4096, 2048, -40.0, 3.0, 50.0)
    st_o, lv_o = vo.simulate_gate(x[:, 0], sr, 4096, 2048, -40.0, 3.0, 50.0)
    assert st == st_o and lv == lv_o
    got = prod.compute_conditional_spectrum(x[:, 0], y[:, 0], sr, st, 4096, 2048)
    ref = vo.compute_conditional_spectrum(x[:, 0], y[:, 0], sr, st, 4096, 2048)
    assert got[3:] == ref[3:5] and np.abs(got[1] - ref[1]).max() < DB_TOL and np.abs(got[2] - ref[2]).max() < DB_TOL
    # the median itself, odd / even / single frame counts, against NumPy on the same device ratios
    for frames in ([5], [3, 9], [2, 4, 6], list(range(1, 41)), list(range(0, 69))):
        med = engine.cond_spectrum_median(x, y, frames)
        cols = np.stack([engine.cond_spectrum_median(x, y, [i]) for i in frames])       # a one-frame median is that frame's ratio
        assert np.array_equal(med, np.median(cols, axis=0))
    # nothing selected -> zeros (src/validate_layer1.py:380-381)
    f, c1, c2, n1, n2 = prod.compute_conditional_spectrum(x, y, sr, st, 4096, 2048, level_threshold=0.0)
    assert (n1, n2) == (0, 0) and not c1.any() and not c2.any() and c1.shape == (2049,)
    # empty anchor band -> no normalisation
    a = prod.compute_conditional_spectrum_v2(x, y, sr, st, np.array(lv), 4096, 2048, anchor_band=(1, 2))
    b = vo.compute_conditional_spectrum_v2(x, y, sr, st, np.array(lv), 4096, 2048, anchor_band=(1, 2))
    assert a[3:] == b[3:5] and np.abs(a[1] - b[1]).max() < DB_TOL and np.abs(a[2] - b[2]).max() < DB_TOL
    with pytest.raises(ValueError):
        prod.compute_conditional_spectrum(x, y[:50000], sr, st, 4096, 2048)
    with pytest.raises(RuntimeError):
        engine.cond_spectrum_median(x, y, [len(x) // 2048])                              # frame past the end of the file
    with pytest.raises(NotImplementedError):
        prod.simulate_gate(x, sr, 1024, 512, -40.0, 3.0, 50.0)                           # neither 4096 / 2048 nor 2048 / 1024
