"""Live check of the restatement against the reference executed in place (build container only).

Skipped wherever /root/reference is absent (e.g. the GPU box) -- the committed fixtures in
tests/golden/ carry the same evidence there.
"""
import numpy as np
import pytest

from oracle import ref_harness as rh
from oracle import tomatis_oracle as orc
from tomatis_audio_processor_b200 import synth

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference sources not present")


def _check(mode, x, sr, **kw):
    r = rh.run_reference(mode, x, sr, **kw)
    o = orc.run(mode, x, sr, **kw)
    assert r["chunk_lengths"] == o["chunk_lengths"]
    assert r["out"].dtype == o["out"].dtype
    assert np.array_equal(r["out"], o["out"])
    assert orc.csv_rows(mode, o) == r["csv"]
    return o


def test_standard_default_48k():
    x = synth.recipe_gated_pink(6.0, 48000, 1, env_hz=0.8)
    o = _check("standard", x, 48000, gate_ui=50)
    assert len(o["chunk_lengths"]) == 2 and o["chunk_lengths"][0] == 239616


def test_standard_linear_output_gain_44k1():
    x = synth.recipe_gated_pink(3.1, 44100, 2, env_hz=1.5)
    _check("standard", x, 44100, gate_ui=55, gate_mode="linear", output_gain_db=-2.5, up_delay_ms=80.0)


def test_xfade_ramps_and_hard():
    x = synth.recipe_threshold_ramps(4.0, 48000, 3, t_on=-48.5, t_off=-51.5, period_s=1.3)
    _check("xfade", x, 48000, gate_ui=50, xfade_ms=500.0)
    _check("xfade", x, 48000, gate_ui=50)


def test_adaptive_both_dtype_branches():
    x = synth.recipe_swept_pink(4.0, 48000, 4, period_s=1.1, peak=0.5)
    o = _check("adaptive", x, 48000)
    assert o["pipeline_dtype"] == "float32"
    x = synth.recipe_swept_pink(4.0, 48000, 5, period_s=1.1, peak=0.1)
    o = _check("adaptive", x, 48000)
    assert o["pipeline_dtype"] == "float64"


@pytest.mark.parametrize("n", [1, 100, 2047, 2048, 2049, 4095, 4096, 4097, 6144, 10000])
def test_ragged_short_inputs(n):
    x = synth.recipe_swept_pink(0.25, 48000, 6)[:n]
    _check("standard", x, 48000)
    _check("xfade", x, 48000, xfade_ms=100.0)
    if n >= 2048:                      # the reference divides by len(levels)==0 below one hop
        _check("adaptive", x, 48000)


def test_reference_error_behaviour_pinned():
    x0 = np.zeros((0, 2), np.float32)
    with pytest.raises(ZeroDivisionError):
        rh.run_reference("standard", x0, 48000)
    with pytest.raises(ValueError):
        rh.run_reference("adaptive", x0, 48000)
    with pytest.raises(ZeroDivisionError):
        rh.run_reference("adaptive", np.zeros((100, 2), np.float32) + 0.1, 48000)


def test_adaptive_mono_file():
    """The adaptive script accepts single-channel files (src/process_tomatis_adaptive.py:180-181)."""
    x = synth.recipe_swept_pink(3.0, 48000, 8, period_s=1.1, peak=0.5)[:, :1]
    o = _check("adaptive", x, 48000)
    assert o["out"].shape == (len(x), 1)


@pytest.mark.parametrize("kw", [dict(), dict(pad=False, global_gain_db=-3.0), dict(auto_gain_protect=False, global_gain_db=2.0)])
def test_static_eq_restatement(kw):
    """oracle/layer2_oracle.py against src/layer2_apply_eq.py executed in place, including the gain-protected second file
    the reference makes by re-reading its own PCM_24 output."""
    from oracle import layer2_oracle as l2
    x = synth.recipe_gated_pink(2.0, 48000, 71, env_hz=1.0, hi_dbfs=-14.0)
    fr, db = [20.0, 100.0, 500.0, 1000.0, 4000.0, 12000.0, 20000.0], [6.0, 4.0, 0.0, -2.0, 3.0, 8.0, 10.0]
    r = rh.run_reference_eq(x, 48000, fr, db, **kw)
    g = l2.build_gain_per_bin(48000, 4096, r["eq_freqs"], r["eq_db"])
    assert np.array_equal(g, r["gain_bins"])
    o = l2.apply_eq(x, 48000, g, **kw)
    assert np.array_equal(o["out"], r["out"])
    assert (o["out_gp"] is None) == (r["out_gp"] is None)
    if r["out_gp"] is not None:
        assert np.array_equal(o["out_gp"], r["out_gp"])
