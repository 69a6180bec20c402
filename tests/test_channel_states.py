"""Per-channel state analyser (src/analyze_stereo_state.py): oracle vs the executed reference's
fixtures (CPU), front end (CPU), CUDA path vs oracle and fixtures (GPU; CSV text must be identical)."""
import csv

import numpy as np
import pytest

from helpers import chan_golden_names, load_chan_golden
from oracle import analysis_oracle as ao
from tomatis_audio_processor_b200 import analyze_stereo_state as prod, synth


def _as_text(rows):
    return [[str(c) for c in r] for r in rows]


@pytest.mark.parametrize("name", chan_golden_names())
def test_channel_oracle_matches_reference_fixture(name):
    g = load_chan_golden(name)
    o = ao.analyze(g["x"], g["sr"], **g["kwargs"])
    assert o["rows"] == g["csv"]                       # header, times, 2-decimal levels, states: verbatim
    assert f"T={o['left_T']:.2f} dBFS, C2={o['left_c2'] * 100:.1f}%" in g["stdout"]
    assert f"T={o['right_T']:.2f} dBFS, C2={o['right_c2'] * 100:.1f}%" in g["stdout"]
    assert not np.array_equal(o["left_states"], o["right_states"])      # the fixture does tell the channels apart


def test_channel_front_end_mirrors_reference():
    assert prod.CSV_HEADER == load_chan_golden(chan_golden_names()[0])["csv"][0]
    flags = {s for a in prod.build_parser()._actions for s in a.option_strings if s.startswith("--")} - {"--help"}
    assert flags == {"--input", "--output", "--target_c2", "--hyst_db", "--min_hold_ms", "--device"}
    d = {a.dest: a.default for a in prod.build_parser()._actions}
    assert (d["target_c2"], d["hyst_db"], d["min_hold_ms"]) == (0.5, 3.0, 250.0)
    for t in (0.0, 59.999, 60.0, 61.5, 3599.99, 7261.25):
        assert prod.format_time(t) == ao.format_time(t)
    assert prod.format_time(61.5) == "1:01.50"


def test_channel_oracle_silence_and_search_exit():
    o = ao.analyze(np.zeros((30000, 2), np.float32), 48000)
    assert o["left_T"] == -120.0 and o["left_c2"] == 0.0 and len(o["rows"]) == 15
    with pytest.raises(ZeroDivisionError):
        ao.analyze(np.zeros((1000, 2), np.float32), 48000)             # no frame at all (src/analyze_stereo_state.py:126)


# ------------------------------------------------------------------------------------------------ GPU
def _check_result(r, o):
    assert np.array_equal(r["left_levels"], o["left_levels"]) and np.array_equal(r["right_levels"], o["right_levels"])
    assert r["left_T"] == o["left_T"] and r["right_T"] == o["right_T"]
    assert np.array_equal(r["left_states"], o["left_states"]) and np.array_equal(r["right_states"], o["right_states"])
    assert np.array_equal(r["times"], o["times"])
    assert _as_text(prod.csv_rows(r)) == o["rows"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", chan_golden_names())
def test_channel_gpu_matches_reference_fixture(name):
    from tomatis_audio_processor_b200 import engine
    g = load_chan_golden(name)
    r = engine.run_channel_states([g["x"]], g["sr"], **g["kwargs"])[0]
    assert _as_text(prod.csv_rows(r)) == g["csv"]
    _check_result(r, ao.analyze(g["x"], g["sr"], **g["kwargs"]))


@pytest.mark.gpu
def test_channel_gpu_batch_ragged_and_degenerate_tracks():
    """Several tracks in one plan (lock-step searches), ragged lengths, a silent track, a one-sided track, a track with
    a single frame, a frameless track."""
    from tomatis_audio_processor_b200 import engine
    sr = 48000
    rng = np.random.default_rng(5)
    xs = []
    for i, n in enumerate((100000, 2048 * 31, 2048 * 31 + 2047, 4096 + 5)):
        x = synth.recipe_swept_pink(n / sr, sr, 40 + i, period_s=0.4 + 0.1 * i, peak=0.3)[:n]
        x[:, 1] = np.roll(x[:, 1], 3000 + 500 * i) * (0.4 + 0.1 * i)
        xs.append(np.ascontiguousarray(x))
    xs.append(np.zeros((50000, 2), np.float32))                                   # silent: no valid level
    one = synth.recipe_gated_pink(1.0, sr, 47, env_hz=3.0)
    one[:, 1] = 0.0
    xs.append(one)                                                                # right channel silent
    xs.append((rng.standard_normal((2048, 2)) * 0.05).astype(np.float32))         # exactly one frame
    xs.append(np.zeros((1000, 2), np.float32))                                    # no frame
    res = engine.run_channel_states(xs, sr, min_hold_ms=120.0, target_c2=0.4)
    for x, r in zip(xs[:-1], res[:-1]):
        _check_result(r, ao.analyze(x, sr, min_hold_ms=120.0, target_c2=0.4))
    assert len(res[-1]["times"]) == 0 and len(res[-1]["left_states"]) == 0
    # one long track: the segmented gate scan (> 16 384 frames) under the same search
    n = 2048 * 17000 + 123
    x = synth.recipe_swept_pink(60.0, sr, 48, period_s=3.0, peak=0.4)
    x = np.ascontiguousarray(np.tile(x, (n // len(x) + 1, 1))[:n])
    x[:, 1] = np.roll(x[:, 1], 12345) * 0.5
    _check_result(engine.run_channel_states([x], sr)[0], ao.analyze(x, sr))


@pytest.mark.gpu
def test_channel_gpu_file_level(tmp_path):
    from tomatis_audio_processor_b200 import audio_io
    g = load_chan_golden("chan_48k_default")
    wav, out = str(tmp_path / "in.wav"), str(tmp_path / "state.csv")
    audio_io.write(wav, g["x"], g["sr"], subtype="PCM_16")
    assert prod.main(["-i", wav, "-o", out]) == 0
    with open(out, "r", encoding="utf-8", newline="") as f:
        assert list(csv.reader(f)) == g["csv"]
    mono = str(tmp_path / "mono.wav")
    audio_io.write(mono, g["x"][:, :1], g["sr"], subtype="PCM_16")
    assert prod.analyze(mono, out) == 1                                           # src/analyze_stereo_state.py:83-85
    short = str(tmp_path / "short.wav")
    audio_io.write(short, g["x"][:1000], g["sr"], subtype="PCM_16")
    with pytest.raises(ZeroDivisionError):
        prod.analyze(short, out)
