"""Long-file command line (process_sharded.py): world size 2 over gloo on the CPU with the NumPy stand-in backend -- every
rank reads only its own sample range of the WAV file, rank 0 writes the gathered output and the state CSV; the oracle on the
whole file is the checker.  GPU (needs 2 GPUs): the same through torchrun and NCCL."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from tomatis_audio_processor_b200 import audio_io, synth

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp, mode, kw):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import io
    from contextlib import redirect_stdout
    import torch.distributed as dist
    from shard_standin import NumpyShardBackend
    from tomatis_audio_processor_b200 import audio_io as aio, process_sharded, sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        reads = []
        real = aio.read_range
        aio.read_range = lambda path, a, b, dtype="float32": (reads.append((a, b)), real(path, a, b, dtype))[1]
        make = lambda shard, window, rows, key: NumpyShardBackend(shard, window, rows, key)
        buf = io.StringIO()
        with redirect_stdout(buf):
            r = process_sharded.process_sharded(mode, os.path.join(tmp, "in.wav"), os.path.join(tmp, f"out_{mode}.flac"),
                                                sharded.Comm(None, "cpu"), state_csv_path=os.path.join(tmp, f"state_{mode}.csv"),
                                                make_backend=make, **kw)
        me = r["shard"]
        assert reads == [(me.own_lo, me.own_hi)] and 0 < me.own_hi - me.own_lo < me.total      # only its own samples
        assert ("[OK]" in buf.getvalue()) == (rank == 0)
        with pytest.raises(ValueError):
            process_sharded.process_sharded("standard", os.path.join(tmp, "in441.wav"), os.path.join(tmp, "x.flac"),
                                            sharded.Comm(None, "cpu"), make_backend=make)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,kw", [("standard", dict(gate_ui=50, up_delay_ms=80.0)), ("xfade", dict(gate_ui=60, xfade_ms=120.0, up_delay_ms=40.0)),
                                     ("adaptive", dict(min_hold_ms=100.0, xfade_ms=200.0))])
def test_long_file_two_ranks_gloo(tmp_path, mode, kw):
    import torch.multiprocessing as mp
    from helpers import csv_states  # noqa: F401  (tests/ on sys.path)
    from oracle import tomatis_oracle as orc
    from tomatis_audio_processor_b200 import report
    sr = 48000
    x = (synth.recipe_swept_pink(6.0, sr, 41, period_s=1.1, peak=0.5) if mode == "adaptive"
         else synth.recipe_gated_pink(11.0, sr, 40, env_hz=0.9, hi_dbfs=-22.0))
    audio_io.write(str(tmp_path / "in.wav"), x, sr, subtype="PCM_16")
    audio_io.write(str(tmp_path / "in441.wav"), x[:5000], 44100, subtype="PCM_16")
    x, _ = audio_io.read(str(tmp_path / "in.wav"))
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), mode, kw), nprocs=2, join=True)
    ext = ".flac" if audio_io.have_soundfile() else ".wav"
    y, sr_out = audio_io.read(str(tmp_path / f"out_{mode}{ext}"))
    o = orc.run(mode, x, sr, **kw)
    assert sr_out == sr and y.shape == x.shape
    assert np.abs(y - o["out"]).max() <= (2e-7 if mode == "standard" else 1.2e-6)      # PCM_24 rounding (+ the alpha note of test_sharded_gloo)
    with open(tmp_path / f"state_{mode}.csv", encoding="utf-8") as f:
        rows = [line.rstrip("\n").split(",") for line in f]
    res = dict(states=np.asarray(o["states"]), levels=np.asarray(o["levels"], np.float64), xfade_frames=o["xfade_frames"], sr=sr)
    if mode == "adaptive":
        res["times"] = o["times"]
    else:
        res["frame_starts"], res["csv_mask"] = o["frame_starts"], o["csv_mask"]
    assert rows == report.state_csv_rows(mode, res)


@pytest.mark.gpu
def test_long_file_cli_torchrun_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    from tomatis_audio_processor_b200 import process_tomatis
    sr = 48000
    x = synth.recipe_gated_pink(16.0, sr, 42, env_hz=0.9, hi_dbfs=-22.0)
    src, ext = str(tmp_path / "in.wav"), (".flac" if audio_io.have_soundfile() else ".wav")
    audio_io.write(src, x, sr, subtype="PCM_16")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), "-m", "tomatis_audio_processor_b200.process_sharded", "--mode", "standard",
           "-i", src, "-o", str(tmp_path / "sharded.flac"), "--gate_ui", "50"]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert process_tomatis.main(["-i", src, "-o", str(tmp_path / "single.flac"), "--gate_ui", "50"]) == 0
    a, _ = audio_io.read(str(tmp_path / ("sharded" + ext)))
    b, _ = audio_io.read(str(tmp_path / ("single" + ext)))
    assert np.array_equal(a, b)


@pytest.mark.gpu
def test_one_process_is_the_single_file_front_end(tmp_path):
    from tomatis_audio_processor_b200 import process_sharded, process_tomatis_xfade
    sr = 48000
    x = synth.recipe_threshold_ramps(3.0, sr, 43, t_on=-48.5, t_off=-51.5, period_s=1.0)
    src, ext = str(tmp_path / "in.wav"), (".flac" if audio_io.have_soundfile() else ".wav")
    audio_io.write(src, x, sr, subtype="PCM_16")
    env = {k: os.environ.pop(k, None) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        assert process_sharded.main(["--mode", "xfade", "-i", src, "-o", str(tmp_path / "a.flac"), "--xfade_ms", "200"]) == 0
    finally:
        os.environ.update({k: v for k, v in env.items() if v is not None})
    assert process_tomatis_xfade.main(["-i", src, "-o", str(tmp_path / "b.flac"), "--xfade_ms", "200"]) == 0
    a, _ = audio_io.read(str(tmp_path / ("a" + ext)))
    b, _ = audio_io.read(str(tmp_path / ("b" + ext)))
    assert np.array_equal(a, b)
