"""CPU tests of the reference-facing host layer: audio file edge, state CSV formatting, CLI flag sets,
the C-ABI export list.  No compute calls (no GPU here); the oracle and the golden fixtures are the checkers."""
import ctypes
import os
import re

import numpy as np
import pytest

from helpers import golden_names, load_golden
from tomatis_audio_processor_b200 import audio_io, report, tables as tb
from tomatis_audio_processor_b200 import process_tomatis, process_tomatis_adaptive, process_tomatis_xfade

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ audio file edge
@pytest.mark.parametrize("subtype,tol", [("PCM_24", 2e-7), ("PCM_16", 5e-5), ("FLOAT", 0.0)])
def test_wav_roundtrip(tmp_path, subtype, tol):
    rng = np.random.default_rng(0)
    y = np.clip(rng.standard_normal((1000, 2)) * 0.3, -0.99, 0.99).astype(np.float32)
    y[0] = [0.999, -0.999]
    p = str(tmp_path / "a.wav")
    audio_io.write(p, y, 44100, subtype=subtype)
    i = audio_io.info(p)
    assert (i.samplerate, i.channels, i.frames, i.subtype) == (44100, 2, 1000, subtype)
    x, sr = audio_io.read(p, dtype="float32")
    assert sr == 44100 and x.shape == y.shape and x.dtype == np.float32
    assert float(np.abs(x - y).max()) <= tol


def test_pcm24_quantiser_matches_libsndfile_rule():
    """libsndfile's clipping conversions (python-soundfile switches clipping on): FLAC rounds x * 2^23 half to even and pins,
    WAV keeps the top three bytes of lrint(x * 2^31), i.e. the floor."""
    L = 1.0 / 8388608.0
    y = np.array([0.0, 1.0, -1.0, 0.5, 1.5, -1.5, 0.999, 2.5 * L, 3.5 * L, -2.5 * L, -0.25 * L, 0.75 * L])
    assert audio_io.quantise_pcm24(y).tolist() == [0, 8388607, -8388608, 4194304, 8388607, -8388608, 8380219, 2, 4, -2, 0, 1]
    assert audio_io.quantise_pcm24(y, "WAV").tolist() == [0, 8388607, -8388608, 4194304, 8388607, -8388608, 8380219, 2, 3, -3, -1, 0]
    y32 = np.float32([0.999, -0.999, 0.1])
    assert audio_io.quantise_pcm24(y32).tolist() == [int(np.rint(float(v) * 8388608.0)) for v in y32]
    assert audio_io.quantise_pcm16(np.array([0.5, -1.0, 1.0, 1.25 / 32768.0, -0.5 / 32768.0])).tolist() == [16384, -32768, 32767, 1, -1]


def test_flac_without_libsndfile_raises_format_unavailable(tmp_path):
    if audio_io.have_soundfile():
        pytest.skip("soundfile is installed: FLAC works")
    with pytest.raises(audio_io.AudioFormatUnavailable):
        audio_io.write(str(tmp_path / "a.flac"), np.zeros((4, 2), np.float32), 48000, subtype="PCM_24", format="FLAC")
    with pytest.raises(audio_io.AudioFormatUnavailable):
        audio_io.read(str(tmp_path / "missing.flac"))


# ------------------------------------------------------------------ state CSV (a14)
def _as_engine_result(mode, o):
    """Shape an oracle result like engine.run_* returns it (the fields report.py consumes)."""
    res = dict(states=np.asarray(o["states"]), levels=np.asarray(o["levels"], dtype=np.float64),
               xfade_frames=o["xfade_frames"], sr=o["sr"])
    if mode == "adaptive":
        res["times"] = o["times"]
    else:
        res["frame_starts"], res["csv_mask"] = o["frame_starts"], o["csv_mask"]
    return res


@pytest.mark.parametrize("name", golden_names())
def test_state_csv_rows_match_reference_files(name):
    from oracle import tomatis_oracle as orc
    g = load_golden(name)
    o = orc.run(g["mode"], g["x"], g["sr"], **g["kwargs"])
    rows = report.state_csv_rows(g["mode"], _as_engine_result(g["mode"], o))
    if np.__version__ == g["numpy"]:
        assert rows == g["csv"]                      # verbatim: the reference's own CSV file
    else:
        assert [r[3] for r in rows] == [r[3] for r in g["csv"]]


def test_gate_statistics():
    st = report.gate_statistics(np.array([1, 1, 2, 2, 2, 1, 2, 2], np.uint8), 48000 * 60, 48000, min_hold_frames=2)
    assert st["frames"] == 8 and st["c2_frames"] == 5 and st["switches"] == 3
    assert st["switches_per_min"] == 3.0 and st["short_run_ratio"] == 0.25


# ------------------------------------------------------------------ CLI flag sets
def _flags(parser):
    out = set()
    for a in parser._actions:
        out.update(s for s in a.option_strings if s.startswith("--"))
    return out - {"--help"}


def test_cli_flags_match_reference():
    std = {"--input", "--output", "--gate_ui", "--gate_mode", "--dynamic_range", "--gate_scale", "--gate_offset",
           "--hyst_db", "--up_delay_ms", "--fc", "--slope", "--c1_low", "--c1_high", "--c2_low", "--c2_high",
           "--n_fft", "--hop", "--state_csv", "--output_gain_db"}
    ext = {"--any_sr", "--device"}
    assert _flags(process_tomatis.build_parser()) == std | ext
    xf = (std - {"--gate_mode", "--dynamic_range", "--output_gain_db"}) | {"--xfade_ms"}
    assert _flags(process_tomatis_xfade.build_parser()) == xf | ext
    ad = {"--input", "--output", "--state_csv", "--fc", "--slope", "--c1_low", "--c1_high", "--c2_low", "--c2_high",
          "--target_c2", "--hyst_db", "--min_hold_ms", "--xfade_ms", "--headroom_margin", "--n_fft", "--hop"}
    assert _flags(process_tomatis_adaptive.build_parser()) == ad | {"--device"}
    assert "--gate_ui" not in _flags(process_tomatis_adaptive.build_parser())
    a = process_tomatis_xfade.build_parser().parse_args(["-i", "a", "-o", "b"])
    assert a.xfade_ms == 0.0 and a.gate_ui == 50
    a = process_tomatis.build_parser().parse_args(["-i", "a", "-o", "b"])
    assert a.gate_mode == "log_percent" and a.up_delay_ms == 250.0 and a.hyst_db == 3.0


def test_process_signatures_match_reference():
    import inspect
    s = inspect.signature(process_tomatis.process)
    assert list(s.parameters) == ["in_path", "out_path", "gate_ui", "gate_mode", "dynamic_range", "gate_scale", "gate_offset",
                                  "hysteresis_db", "fc", "slope", "c1_low", "c1_high", "c2_low", "c2_high", "up_delay_ms",
                                  "n_fft", "hop", "state_csv_path", "output_gain_db"]
    assert s.parameters["gate_offset"].default == -100 and s.parameters["gate_mode"].default == "log_percent"
    s = inspect.signature(process_tomatis_xfade.process)
    assert list(s.parameters) == ["in_path", "out_path", "gate_ui", "gate_scale", "gate_offset", "hysteresis_db", "fc", "slope",
                                  "c1_low", "c1_high", "c2_low", "c2_high", "up_delay_ms", "xfade_ms", "n_fft", "hop",
                                  "state_csv_path"]
    assert s.parameters["xfade_ms"].default == 0.0
    s = inspect.signature(process_tomatis_adaptive.process)
    assert list(s.parameters) == ["in_path", "out_path", "fc", "slope", "c1_low", "c1_high", "c2_low", "c2_high", "target_c2",
                                  "hyst_db", "min_hold_ms", "xfade_ms", "headroom_margin", "n_fft", "hop", "state_csv_path"]
    assert s.parameters["xfade_ms"].default == 500.0 and s.parameters["min_hold_ms"].default == 250.0


def test_guard_and_error_exit_code_without_gpu(tmp_path, capsys):
    """sr != 48 kHz raises ValueError before any device work; main() maps it to exit code 1
    (src/process_tomatis.py:234-237, 519-544)."""
    p = str(tmp_path / "in.wav")
    audio_io.write(p, np.zeros((5000, 2), np.float32), 44100, subtype="PCM_16")
    with pytest.raises(ValueError):
        process_tomatis.process(p, str(tmp_path / "o.flac"))
    assert process_tomatis.main(["-i", p, "-o", str(tmp_path / "o.flac")]) == 1
    assert process_tomatis_xfade.main(["-i", p, "-o", str(tmp_path / "o.flac")]) == 1
    m = str(tmp_path / "mono.wav")
    audio_io.write(m, np.zeros((5000, 1), np.float32), 48000, subtype="PCM_16")
    with pytest.raises(ValueError):
        process_tomatis.process(m, str(tmp_path / "o.flac"))
    capsys.readouterr()


# ------------------------------------------------------------------ C ABI
def test_library_exports_every_declared_symbol():
    from tomatis_audio_processor_b200 import _lib, build
    for n_fft, path in build.LIB_PATHS.items():            # one build per fused frame size (4096 / 2048, 2048 / 1024)
        if not os.path.exists(path):
            build.build_library(n_fft=n_fft)
    with open(os.path.join(ROOT, "include", "tomatis_b200.h")) as f:
        hdr = f.read()
    declared = set(re.findall(r"\b(tmt_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no prototypes found"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for n_fft, path in build.LIB_PATHS.items():
        lib = ctypes.CDLL(path)
        for name in declared:
            getattr(lib, name)                 # AttributeError = missing export
        assert _lib.load(n_fft).tmt_version() >= 100 and _lib.load(n_fft).n_fft == n_fft


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tomatis_audio_processor_b200 import engine
    with pytest.raises(RuntimeError):
        engine.run("standard", [np.zeros((4096, 2), np.float32)], 48000)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tomatis_audio_processor_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn


def test_each_fused_frame_size_has_its_own_library_and_no_fallback(monkeypatch, tmp_path):
    """engine.fused_size names the sizes the fused kernels serve; a missing build of the library is an error, not a fallback."""
    from tomatis_audio_processor_b200 import _lib, build, engine
    assert build.FUSED_SIZES == {4096: 2048, 2048: 1024} and set(build.LIB_PATHS) == set(build.FUSED_SIZES)
    assert engine.fused_size(4096, 2048) and engine.fused_size(2048, 1024)
    assert not engine.fused_size(2048, 512) and not engine.fused_size(1024, 512) and not engine.fused_size(4096, 1024)
    monkeypatch.setenv("TMT_FUSED_2048", "0")                      # comparisons: send 2048 / 1024 to the general-size path
    assert not engine.fused_size(2048, 1024) and engine.fused_size(4096, 2048)
    monkeypatch.delenv("TMT_FUSED_2048")
    with pytest.raises(NotImplementedError):
        engine.Engine(0, 1024, 512)                                # checked before anything touches CUDA
    monkeypatch.setattr(_lib, "_libs", {})
    monkeypatch.setenv("TMT_LIB_2048", str(tmp_path / "missing.so"))
    with pytest.raises(RuntimeError, match="not built"):
        _lib.load(2048)
    with pytest.raises(RuntimeError):
        _lib.load(1024)
