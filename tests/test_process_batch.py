"""File-level batch front end (process_batch.py; the caller of the path: the reference's ForEach loop,
docs/Tomatis处理器使用指南.md:243-249).  CPU: flag forwarding, rank assignment, wave building, per-file error isolation and
the statistics all-reduce over gloo (world size 2) with the engine call replaced by the oracle (test infrastructure).
GPU: the real thing on files -- every output equals the single-file front end's."""
import os
import socket
import sys

import numpy as np
import pytest

from tomatis_audio_processor_b200 import audio_io, process_batch as pb, synth

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _make_files(d, specs):
    """specs: (name, sr, seconds, channels, seed) -> PCM_16 WAV files; returns paths."""
    paths = []
    for name, sr, secs, ch, seed in specs:
        x = synth.recipe_gated_pink(secs, sr, seed, env_hz=2.0, hi_dbfs=-22.0)[:, :ch]
        p = os.path.join(str(d), name + ".wav")
        audio_io.write(p, x, sr, subtype="PCM_16")
        paths.append(p)
    return paths


def _oracle_engine_run(calls):
    """Stand-in for engine.run: the oracle, shaped like the engine's results (fields the front end consumes)."""
    from oracle import tomatis_oracle as orc

    def run(mode, xs, sr, device=0, **kw):
        calls.append((mode, len(xs), sr, device, dict(kw)))
        out = []
        for x in xs:
            o = orc.run(mode, np.asarray(x, np.float32), sr, **{k: v for k, v in kw.items()})
            y = o["out"].astype(np.float32)
            peaks = np.array([np.abs(y).max()], np.float32)
            r = dict(out=y, states=np.asarray(o["states"]), levels=np.asarray(o["levels"], np.float64), sr=sr,
                     xfade_frames=o["xfade_frames"], chunk_lengths=list(o["chunk_lengths"]), chunk_peaks=peaks)
            if mode == "adaptive":
                r["times"] = o["times"]
            else:
                r["frame_starts"], r["csv_mask"] = o["frame_starts"], o["csv_mask"]
            out.append(r)
        return out
    return run


def test_flag_forwarding_and_assignment():
    for mode in ("standard", "xfade", "adaptive"):
        args = pb.MODES[mode].build_parser().parse_args(["-i", "_", "-o", "_", "--hyst_db", "2.5", "--fc", "900"])
        kw = pb.engine_kwargs(mode, args)
        assert kw["fc"] == 900 and kw["n_fft"] == 4096 and kw["hop"] == 2048
        assert kw["hyst_db" if mode == "adaptive" else "hysteresis_db"] == 2.5
        # exactly the keyword arguments the single-file process() of the mode exposes (minus paths and the CSV)
        import inspect
        sig = set(inspect.signature(pb.MODES[mode].process).parameters) - {"in_path", "out_path", "state_csv_path"}
        assert set(kw) == sig
    assert pb.assignment(7, 0, 2) == [0, 2, 4, 6] and pb.assignment(7, 1, 2) == [1, 3, 5] and pb.assignment(0, 0, 4) == []
    assert sorted(sum((pb.assignment(10, r, 4) for r in range(4)), [])) == list(range(10))
    assert pb.output_path("/a/b/song.x.flac", "out", "_t") == os.path.join("out", "song.x_t.flac")


def test_batch_on_oracle_engine(tmp_path, monkeypatch, capsys):
    from tomatis_audio_processor_b200 import engine, process_tomatis
    calls = []
    monkeypatch.setattr(engine, "run", _oracle_engine_run(calls))
    files = _make_files(tmp_path, [("a", 48000, 1.2, 2, 1), ("b", 48000, 0.8, 2, 2), ("c", 44100, 1.0, 2, 3), ("d", 48000, 0.5, 1, 4)])
    files.insert(2, os.path.join(str(tmp_path), "missing.wav"))
    out_dir, csv_dir = str(tmp_path / "out"), str(tmp_path / "csv")
    rc = pb.main(["--mode", "standard", "-i"] + files + ["--out_dir", out_dir, "--state_csv_dir", csv_dir,
                                                        "--wave_sample_frames", "60000", "--gate_ui", "55", "--up_delay_ms", "40"])
    text = capsys.readouterr().out
    assert rc == 1                                                     # three of five files fail and are reported
    assert "[DONE] 2 of 5 file(s)" in text and text.count("[FAILED]") == 3
    assert "expected 48 kHz, got 44100 Hz" in text and "expected stereo, got 1 channel(s)" in text
    # a and b do not fit one 60 000-sample-frame wave together: two engine calls, flags forwarded
    assert [(c[0], c[1], c[2]) for c in calls] == [("standard", 1, 48000), ("standard", 1, 48000)]
    assert calls[0][4]["gate_ui"] == 55 and calls[0][4]["up_delay_ms"] == 40 and calls[0][4]["gate_mode"] == "log_percent"
    # outputs: WAV fallback next to the requested FLAC name (no libsndfile here), same samples as the oracle, CSV per file
    from oracle import tomatis_oracle as orc
    for name in ("a", "b"):
        x, sr = audio_io.read(os.path.join(str(tmp_path), name + ".wav"))
        ext = ".flac" if audio_io.have_soundfile() else ".wav"
        y, _ = audio_io.read(os.path.join(out_dir, name + "_tomatis" + ext))
        ref = orc.run("standard", x, sr, gate_ui=55, up_delay_ms=40)["out"]
        assert y.shape == x.shape and np.abs(y - ref).max() <= 2e-7       # PCM_24 rounding
        assert os.path.getsize(os.path.join(csv_dir, name + "_state.csv")) > 0
    # --any_sr lifts the guard like the single-file command line; adaptive accepts mono and any rate
    calls.clear()
    rc = pb.main(["--mode", "xfade", "-i", files[3], "--out_dir", out_dir, "--any_sr", "--xfade_ms", "100"])
    assert rc == 0 and calls[0][:3] == ("xfade", 1, 44100) and calls[0][4]["xfade_ms"] == 100
    calls.clear()
    rc = pb.main(["--mode", "adaptive", "-i", files[0], files[4], files[3], "--out_dir", out_dir, "--suffix", "_ad", "--min_hold_ms", "90"])
    assert rc == 0 and sorted((c[1], c[2]) for c in calls) == [(1, 44100), (1, 48000), (1, 48000)]      # stereo / mono / 44.1 kHz apart


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp, files):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import io
    from contextlib import redirect_stdout
    from tomatis_audio_processor_b200 import engine, process_batch
    import test_process_batch as me
    calls = []
    engine.run = me._oracle_engine_run(calls)
    buf = io.StringIO()
    with redirect_stdout(buf):
        rc = process_batch.main(["--mode", "standard", "-i"] + files + ["--out_dir", os.path.join(tmp, "out2"), "--up_delay_ms", "40"])
    text = buf.getvalue()
    assert rc == 1                                                     # the 44.1 kHz file fails on rank 0; EVERY rank knows
    assert sum(c[1] for c in calls) == (2 if rank == 0 else 2)         # files 0, 2, (4 fails) on rank 0; 1, 3 on rank 1
    if rank == 0:
        assert "[DONE] 4 of 5 file(s) processed on 2 GPU(s)" in text and text.count("[FAILED]") == 1
    else:
        assert "[DONE]" not in text


def test_batch_two_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp
    files = _make_files(tmp_path, [("r0", 48000, 0.6, 2, 11), ("r1", 48000, 0.7, 2, 12), ("r2", 48000, 0.5, 2, 13),
                                   ("r3", 48000, 0.9, 2, 14), ("r4", 44100, 0.5, 2, 15)])
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), files), nprocs=2, join=True)
    ext = ".flac" if audio_io.have_soundfile() else ".wav"
    assert sorted(os.listdir(tmp_path / "out2")) == [f"r{k}_tomatis{ext}" for k in range(4)]


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_batch_files_equal_single_file_front_end(tmp_path, capsys):
    from tomatis_audio_processor_b200 import process_tomatis, process_tomatis_adaptive
    files = _make_files(tmp_path, [("g0", 48000, 6.0, 2, 21), ("g1", 48000, 2.5, 2, 22), ("g2", 48000, 0.3, 2, 23)])
    ext = ".flac" if audio_io.have_soundfile() else ".wav"
    for mode, mod, flags in (("standard", process_tomatis, ["--gate_ui", "52"]), ("adaptive", process_tomatis_adaptive, ["--min_hold_ms", "120"])):
        out_dir = str(tmp_path / ("batch_" + mode))
        assert pb.main(["--mode", mode, "-i"] + files + ["--out_dir", out_dir] + flags) == 0
        for f in files:
            single = str(tmp_path / ("single_" + mode + "_" + os.path.basename(f)))
            single = single.replace(".wav", ext)
            assert mod.main(["-i", f, "-o", single.replace(ext, ".flac") if mode == "standard" else single] + flags) == 0
            a, _ = audio_io.read(os.path.join(out_dir, os.path.splitext(os.path.basename(f))[0] + "_tomatis" + ext))
            b, _ = audio_io.read(single)
            assert np.array_equal(a, b)
    assert "[DONE] 3 of 3 file(s)" in capsys.readouterr().out
