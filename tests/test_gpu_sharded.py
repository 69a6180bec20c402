"""GPU: one file split over several ranks (time-chunk sharding) through the CUDA backend.
On the single-GPU box the ranks are threads of one process (tests/thread_comm.py); with >= 2 GPUs the same
driver also runs under torchrun + NCCL (tests/run_sharded_nccl.py)."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run_threads(world, fn):
    from thread_comm import ThreadWorld
    tw = ThreadWorld(world)
    res, err = [None] * world, []

    def body(r):
        try:
            res[r] = fn(tw.comm(r), r)
        except BaseException as e:          # noqa: BLE001 - re-raised in the main thread
            err.append(e)
            tw.barrier.abort()
    th = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    if err:
        raise err[0]
    return res


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("mode,sr,seconds,kw", [
    ("standard", 48000, 11.0, dict(gate_ui=50)),
    ("standard", 44100, 2.5, dict(gate_ui=50, up_delay_ms=80.0, output_gain_db=-1.5)),
    ("xfade", 96000, None, dict(gate_ui=60, xfade_ms=100.0, up_delay_ms=40.0)),
    ("adaptive", 48000, 3.0, dict(min_hold_ms=100.0, xfade_ms=200.0)),
    ("adaptive", 44100, 2.0, dict()),
])
def test_sharded_equals_unsharded(world, mode, sr, seconds, kw):
    import torch
    from tomatis_audio_processor_b200 import engine, sharded, synth
    if mode == "adaptive":
        x = synth.recipe_swept_pink(seconds, sr, 31, period_s=1.1, peak=(0.5 if sr == 48000 else 0.1))
    elif seconds is None:
        x = synth.recipe_threshold_ramps(2048 * 40 / 96000, 96000, 18, t_on=-38.5, t_off=-41.5, period_s=0.4)[:2048 * 40]
    else:
        x = synth.recipe_gated_pink(seconds, sr, 30, env_hz=0.9, hi_dbfs=-22.0)
    total = len(x)
    whole = engine.run(mode, [x], sr, **kw)[0]
    engine.get_engine(0)
    framing = sharded.WHOLEFILE if mode == "adaptive" else sharded.STREAMING
    shards = sharded.plan_shards(total, world, framing)

    def rank_fn(comm, r):
        me = shards[r]
        own = torch.from_numpy(x[me.own_lo:me.own_hi].copy()).cuda()
        if mode == "adaptive":
            out = sharded.run_adaptive_sharded(own, sr, total, comm, gather_to=0, **kw)
        else:
            out = sharded.run_streaming_sharded(mode, own, sr, total, comm, gather_to=0, **kw)
        out["out"] = out["out"].cpu().numpy()
        if out["full"] is not None:
            out["full"] = out["full"].cpu().numpy()
        return out
    res = _run_threads(world, rank_fn)
    for r, o in enumerate(res):
        me = shards[r]
        assert np.array_equal(o["meansq"], whole["meansq"])
        assert np.array_equal(o["states"], whole["states"]) and np.array_equal(o["rows"], whole["rows"])
        assert np.array_equal(o["out"], whole["out"][me.own_lo:me.own_hi]), (mode, world, r)     # same kernels: bit-identical
        if mode == "adaptive":
            assert o["optimal_T"] == whole["optimal_T"] and o["trace"] == whole["trace"]
    assert np.array_equal(res[0]["full"], whole["out"])


@pytest.mark.parametrize("world", [2, 3])
def test_shard_session_repeated_steps(world):
    """The persistent per-rank session bench.py times: halo hand-off overlapped with the hop sums of the owned blocks,
    all-reduce of the hop sums instead of the mean squares.  Two passes with different data, each equal to the unsharded run."""
    import torch
    from tomatis_audio_processor_b200 import engine, sharded, synth
    sr = 48000
    xa = synth.recipe_gated_pink(11.0, sr, 33, env_hz=0.9, hi_dbfs=-22.0)
    xb = synth.recipe_gated_pink(11.0, sr, 34, env_hz=1.3, hi_dbfs=-24.0)
    total = len(xa)
    wa = engine.run("standard", [xa], sr, gate_ui=50)[0]
    wb = engine.run("standard", [xb], sr, gate_ui=50)[0]
    shards = sharded.plan_shards(total, world, sharded.STREAMING)

    def rank_fn(comm, r):
        me = shards[r]
        sess = sharded.StreamingShardSession("standard", torch.from_numpy(xa[me.own_lo:me.own_hi].copy()).cuda(), sr, total, comm, gate_ui=50)
        try:
            sess.step()
            ya = sess.out.cpu().numpy().copy()
            sess.own.copy_(torch.from_numpy(xb[me.own_lo:me.own_hi].copy()).cuda())
            sess.step()
            yb = sess.out.cpu().numpy().copy()
        finally:
            sess.close()
        return ya, yb
    res = _run_threads(world, rank_fn)
    for r, (ya, yb) in enumerate(res):
        me = shards[r]
        assert np.array_equal(ya, wa["out"][me.own_lo:me.own_hi]), r
        assert np.array_equal(yb, wb["out"][me.own_lo:me.own_hi]), r


@pytest.mark.parametrize("n_fft", [4096, 2048])
def test_nccl_two_gpus_torchrun(n_fft):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(HERE, "run_sharded_nccl.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=dict(os.environ, TMT_TEST_NFFT=str(n_fft)))
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "SHARDED-NCCL-OK" in p.stdout


def test_full_size_two_hour_file_sharded_equals_unsharded():
    """BASELINE configs[4] at full size (2 h @ 96 kHz = 691 200 000 sample-frames, pad_end = 0): 4 time shards (threads
    on one GPU) against the unsharded run -- bit-identical output, states and peaks; plus size-independent properties."""
    import torch
    from tomatis_audio_processor_b200 import engine, sharded, synth
    total, sr = 691_200_000, 96000
    x = synth.device_long_file_range(0, total, sr, 5000, "cuda:0")
    whole = engine.run("standard", [x], sr, want_host=False, gate_ui=50)[0]
    y = whole["out"]
    assert y.shape == x.shape and bool(torch.isfinite(y[::97]).all())
    assert float(y.abs().max()) <= 0.999 + 1e-6
    assert len(whole["states"]) == 337_500 and len(whole["chunk_lengths"]) == 2861
    assert whole["chunk_lengths"][0] == 239_616 and sum(whole["chunk_lengths"]) == total
    st = whole["states"]
    assert (st == 1).any() and (st == 2).any()
    world = 4
    shards = sharded.plan_shards(total, world, sharded.STREAMING)
    assert all(s.block_lo % 118 == 0 for s in shards)

    def rank_fn(comm, r):
        me = shards[r]
        out = sharded.run_streaming_sharded("standard", x[me.own_lo:me.own_hi], sr, total, comm, gate_ui=50)
        ok = bool(torch.equal(out["out"], y[me.own_lo:me.own_hi]))
        return ok, np.array_equal(out["states"], st), out["comm_bytes"]
    res = _run_threads(world, rank_fn)
    assert all(r[0] for r in res), [r[0] for r in res]
    assert all(r[1] for r in res)
