"""GPU: seeded random sweep of parameters, lengths, sample rates and batch compositions against the oracle -- exact gate
data, PCM within the north-star tolerance (edge windows against the float64-FFT evaluation, see test_gpu_parity.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
PCM_TOL = 1e-5
EDGE = 256


def _cmp(r, o, o64):
    """(excess over the reference's own float32-vs-float64 FFT noise, error vs the float64-FFT evaluation).  The reference is
    ill-conditioned at its edge blocks, and in adaptive mode an edge sample can set the GLOBAL limiter scale, which moves every
    sample of its float32 run by ~1e-4 relative; hence the pointwise triangle bound instead of a window."""
    y = r["out"].astype(np.float64)
    d = np.abs(y - o["out"].astype(np.float64)).max(axis=1)
    d64 = np.abs(y - o64["out"].astype(np.float64)).max(axis=1)
    noise = np.abs(o["out"].astype(np.float64) - o64["out"].astype(np.float64)).max(axis=1)
    return float((d - noise).max(initial=0.0)), float(d64.max(initial=0.0))


@pytest.mark.parametrize("seed", range(12))
def test_random_streaming_batches(seed):
    from oracle import tomatis_oracle as orc
    from tomatis_audio_processor_b200 import engine, synth
    rng = np.random.default_rng(1000 + seed)
    sr = int(rng.choice([44100, 48000, 96000]))
    mode = str(rng.choice(["standard", "xfade"]))
    kw = dict(gate_ui=float(rng.uniform(35, 65)), hysteresis_db=float(rng.choice([0.0, 1.0, 3.0, 6.0])),
              up_delay_ms=float(rng.choice([0.0, 20.0, 80.0, 250.0])), fc=float(rng.choice([500.0, 1000.0, 2000.0])),
              slope=float(rng.choice([6.0, 12.0, 18.0])))
    if mode == "xfade":
        kw["xfade_ms"] = float(rng.choice([0.0, 40.0, 200.0, 500.0]))
    else:
        kw["gate_mode"] = str(rng.choice(["linear", "log_percent"]))
        kw["output_gain_db"] = float(rng.choice([0.0, -3.0, 1.5]))
    T = -40.0 + (kw["gate_ui"] - 50) * (1.0 if mode == "xfade" or kw.get("gate_mode") == "linear" else 0.8)
    if mode == "xfade" or kw.get("gate_mode") == "linear":
        T = kw["gate_ui"] - 100.0
    xs = []
    for k in range(int(rng.integers(1, 5))):
        n = int(rng.integers(1, 6 * sr // 2))
        x = synth.recipe_threshold_ramps(max(n / sr, 0.01), sr, int(rng.integers(1 << 30)), t_on=T + 1.5, t_off=T - 1.5,
                                         period_s=float(rng.uniform(0.2, 1.5)))[:n]
        xs.append(x)
    rs = engine.run(mode, xs, sr, **kw)
    for x, r in zip(xs, rs):
        o = orc.run(mode, x, sr, **kw)
        o64 = orc.run(mode, x, sr, fft_dtype="float64", **kw)
        assert r["chunk_lengths"] == o["chunk_lengths"]
        assert np.array_equal(r["meansq"], o["meansq"]) and np.array_equal(r["states"], o["states"]), (mode, sr, kw, len(x))
        assert np.allclose(r["rows"] / max(r["xfade_frames"], 1), o["alphas"], atol=1e-9)
        inner, e64 = _cmp(r, o, o64)
        scale = max(1.0, float(np.abs(o64["out"]).max(initial=0.0)))
        assert inner <= PCM_TOL * scale and e64 <= PCM_TOL * scale, (mode, sr, kw, len(x), inner, e64)


@pytest.mark.parametrize("seed", range(8))
def test_random_adaptive_batches(seed):
    from oracle import tomatis_oracle as orc
    from tomatis_audio_processor_b200 import engine, synth
    rng = np.random.default_rng(2000 + seed)
    sr = int(rng.choice([44100, 48000, 96000]))
    kw = dict(target_c2=float(rng.choice([0.3, 0.5, 0.7])), hyst_db=float(rng.choice([1.0, 3.0, 5.0])),
              min_hold_ms=float(rng.choice([0.0, 100.0, 250.0])), xfade_ms=float(rng.choice([50.0, 200.0, 500.0])),
              headroom_margin=float(rng.choice([0.0, 2.0])))
    xs = []
    for k in range(int(rng.integers(1, 4))):
        n = int(rng.integers(2048, 5 * sr // 2))
        peak = float(rng.choice([0.05, 0.12, 0.3, 0.9]))
        xs.append(synth.recipe_swept_pink(max(n / sr, 0.05), sr, int(rng.integers(1 << 30)), period_s=float(rng.uniform(0.3, 1.2)),
                                          peak=peak)[:n])
    rs = engine.run("adaptive", xs, sr, **kw)
    for x, r in zip(xs, rs):
        o = orc.run("adaptive", x, sr, **kw)
        o64 = orc.run("adaptive", x, sr, fft_dtype="float64", **kw)
        assert r["pipeline_dtype"] == o["pipeline_dtype"]
        assert np.array_equal(r["meansq"], np.asarray(o["meansq"])) and np.array_equal(r["states"], o["states"]), (sr, kw, len(x))
        assert r["optimal_T"] == o["optimal_T"] and r["trace"] == o["trace"]
        inner, e64 = _cmp(r, o, o64)
        scale = max(1.0, float(np.abs(o64["out"]).max(initial=0.0)))
        assert inner <= PCM_TOL * scale and e64 <= PCM_TOL * scale, (sr, kw, len(x), inner, e64)
