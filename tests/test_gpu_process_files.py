"""GPU: the reference-facing `process(in_path, out_path, ...)` / CLI front ends on real files
(WAV PCM_24 in this image: no libsndfile, so the FLAC request takes the reference's own WAV fallback)."""
import csv
import os

import numpy as np
import pytest

from tomatis_audio_processor_b200 import audio_io, synth

pytestmark = pytest.mark.gpu
Q24 = 1.0 / 8388607.0


def _oracle():
    from oracle import tomatis_oracle as orc
    return orc


def _write_input(tmp_path, x, sr):
    p = str(tmp_path / "in.wav")
    audio_io.write(p, x, sr, subtype="PCM_24")
    xq, _ = audio_io.read(p, dtype="float32")          # what every implementation actually sees
    return p, xq


def _read_csv(path):
    with open(path, newline="", encoding="utf-8") as f:
        return list(csv.reader(f))


def _out_path(tmp_path):
    want = str(tmp_path / "out.flac")
    return want, (want if audio_io.have_soundfile() else want.replace(".flac", ".wav"))


def test_standard_process_files(tmp_path, capsys):
    from tomatis_audio_processor_b200 import process_tomatis as pt
    x = synth.recipe_gated_pink(6.0, 48000, 21, env_hz=0.9, hi_dbfs=-22.0)
    p, xq = _write_input(tmp_path, x, 48000)
    want, got_path = _out_path(tmp_path)
    csv_path = str(tmp_path / "state.csv")
    assert pt.process(p, want, gate_ui=50, state_csv_path=csv_path) is None
    y, sr = audio_io.read(got_path, dtype="float32")
    o = _oracle().run("standard", xq, 48000, gate_ui=50)
    assert sr == 48000 and y.shape == o["out"].shape                     # output length == input length
    d = np.abs(y.astype(np.float64) - o["out"].astype(np.float64)).max(axis=1)
    assert float(d[:-256].max()) <= 1e-5 + Q24
    rows = _read_csv(csv_path)
    assert rows == _oracle().csv_rows("standard", o)                     # levels, states: verbatim
    assert "C2" in capsys.readouterr().out


def test_xfade_cli_files(tmp_path, capsys):
    from tomatis_audio_processor_b200 import process_tomatis_xfade as px
    x = synth.recipe_threshold_ramps(3.0, 48000, 22, t_on=-48.5, t_off=-51.5, period_s=1.0)
    p, xq = _write_input(tmp_path, x, 48000)
    want, got_path = _out_path(tmp_path)
    csv_path = str(tmp_path / "state.csv")
    rc = px.main(["-i", p, "-o", want, "--gate_ui", "50", "--xfade_ms", "200", "--up_delay_ms", "60", "--state_csv", csv_path])
    assert rc == 0
    y, _ = audio_io.read(got_path, dtype="float32")
    o = _oracle().run("xfade", xq, 48000, gate_ui=50, xfade_ms=200.0, up_delay_ms=60.0)
    d = np.abs(y.astype(np.float64) - o["out"].astype(np.float64)).max(axis=1)
    assert float(d[:-256].max()) <= 1e-5 + Q24
    assert _read_csv(csv_path) == _oracle().csv_rows("xfade", o)
    capsys.readouterr()


def test_adaptive_process_files(tmp_path, capsys):
    from tomatis_audio_processor_b200 import process_tomatis_adaptive as pa
    x = synth.recipe_swept_pink(4.0, 44100, 23, period_s=1.1, peak=0.5)      # no sample-rate guard in adaptive
    p = str(tmp_path / "in.wav")
    audio_io.write(p, x, 44100, subtype="PCM_24")
    xq, _ = audio_io.read(p, dtype="float32")
    out = str(tmp_path / "out.wav")
    csv_path = str(tmp_path / "state.csv")
    assert pa.process(p, out, state_csv_path=csv_path) == 0
    y, sr = audio_io.read(out, dtype="float32")
    o = _oracle().run("adaptive", xq, 44100)
    o64 = _oracle().run("adaptive", xq, 44100, fft_dtype="float64")
    assert sr == 44100 and y.shape == o["out"].shape
    d = np.abs(y.astype(np.float64) - o["out"].astype(np.float64)).max(axis=1)
    d64 = np.abs(y.astype(np.float64) - o64["out"].astype(np.float64)).max(axis=1)
    assert float(d[256:-256].max()) <= 1e-5 + Q24 and float(d64.max()) <= 1e-5 + Q24
    assert _read_csv(csv_path) == _oracle().csv_rows("adaptive", o)
    capsys.readouterr()


def test_adaptive_four_channel_file(tmp_path, capsys):
    """More than two channels (the reference loops over channels, src/process_tomatis_adaptive.py:307-313): the pairs share
    input peak, level, gate and limiter scale."""
    from tomatis_audio_processor_b200 import process_tomatis_adaptive as pa
    a = synth.recipe_swept_pink(3.0, 48000, 61, period_s=0.9, peak=0.5)
    b = synth.recipe_swept_pink(3.0, 48000, 62, period_s=0.6, peak=0.35)
    x = np.concatenate([a, b], axis=1)
    p = str(tmp_path / "in.wav")
    audio_io.write(p, x, 48000, subtype="PCM_24")
    xq, _ = audio_io.read(p, dtype="float32")
    assert xq.shape[1] == 4
    out = str(tmp_path / "out.wav")
    csv_path = str(tmp_path / "state.csv")
    assert pa.process(p, out, state_csv_path=csv_path, min_hold_ms=120.0) == 0
    y, sr = audio_io.read(out, dtype="float32")
    o = _oracle().run("adaptive", xq, 48000, min_hold_ms=120.0)
    o64 = _oracle().run("adaptive", xq, 48000, min_hold_ms=120.0, fft_dtype="float64")
    assert sr == 48000 and y.shape == o["out"].shape == (len(x), 4)
    d = np.abs(y.astype(np.float64) - o["out"].astype(np.float64)).max(axis=1)
    d64 = np.abs(y.astype(np.float64) - o64["out"].astype(np.float64)).max(axis=1)
    assert float(d[256:-256].max()) <= 1e-5 + Q24 and float(d64.max()) <= 1e-5 + Q24
    assert _read_csv(csv_path) == _oracle().csv_rows("adaptive", o)
    capsys.readouterr()


def test_44k1_needs_any_sr_extension(tmp_path, capsys):
    from tomatis_audio_processor_b200 import process_tomatis as pt
    x = synth.recipe_gated_pink(2.0, 44100, 24, env_hz=2.0, hi_dbfs=-28.0)
    p, xq = _write_input(tmp_path, x, 44100)
    want, got_path = _out_path(tmp_path)
    assert pt.main(["-i", p, "-o", want]) == 1                                # the reference's ValueError -> exit code 1
    assert not os.path.exists(got_path)
    try:
        assert pt.main(["-i", p, "-o", want, "--any_sr"]) == 0
    finally:
        pt.REFERENCE_GUARDS = True
    y, _ = audio_io.read(got_path, dtype="float32")
    o = _oracle().run("standard", xq, 44100, gate_ui=50)
    d = np.abs(y.astype(np.float64) - o["out"].astype(np.float64)).max(axis=1)
    assert float(d[:-256].max()) <= 1e-5 + Q24
    capsys.readouterr()
