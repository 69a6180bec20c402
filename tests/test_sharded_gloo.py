"""CPU, world_size 2, gloo: the multi-rank path for one long file -- shard planning,
halo hand-off, exact level gather, redundant gate scan, chunk-peak all-reduce, final gather -- driven through
the product's sharded.py with the NumPy stand-in backend; the oracle on the whole file is the checker."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import tomatis_oracle as orc
    from shard_standin import NumpyShardBackend
    from tomatis_audio_processor_b200 import sharded, synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        comm = sharded.Comm(None, "cpu")
        mode, sr, seconds, kw = case
        if mode == "adaptive":
            x = synth.recipe_swept_pink(seconds, sr, 31, period_s=1.1, peak=kw.pop("peak", 0.5))
        else:
            x = synth.recipe_gated_pink(seconds, sr, 30, env_hz=0.9, hi_dbfs=-22.0)
        x = x[:kw.pop("trim", len(x))]
        total = len(x)
        framing = sharded.WHOLEFILE if mode == "adaptive" else sharded.STREAMING
        me = sharded.plan_shards(total, world, framing)[rank]
        own = torch.from_numpy(x[me.own_lo:me.own_hi].copy())          # the rank holds ONLY its own samples
        make = lambda shard, window, rows, key: NumpyShardBackend(shard, window, rows, key)
        if mode == "adaptive":
            r = sharded.run_adaptive_sharded(own, sr, total, comm, make_backend=make, gather_to=0, **kw)
        else:
            r = sharded.run_streaming_sharded(mode, own, sr, total, comm, make_backend=make, gather_to=0, **kw)
        o = orc.run(mode, x, sr, **kw)
        # per-frame data is complete and identical on EVERY rank (redundant scan)
        assert np.array_equal(r["meansq"], np.asarray(o["meansq"]))
        assert np.array_equal(r["states"], o["states"])
        xe = max(r["xfade_frames"], 1)
        assert np.allclose(r["rows"] / xe, o["alphas"], atol=1e-9)
        assert r["out"].shape[0] == me.own_hi - me.own_lo
        ref = o["out"].astype(np.float32)
        # standard: bit-exact.  Crossfaded modes: the device counts the crossfade in integer steps k/X while the
        # reference accumulates alpha in float64, so a downward ramp differs by ~1e-16 in alpha and an occasional
        # gain bin by one float32 ulp.
        tol = 0.0 if mode == "standard" else 1e-6
        same = lambda a, b: np.array_equal(a, b) if tol == 0.0 else float(np.abs(a - b).max(initial=0.0)) <= tol
        assert same(r["out"].numpy(), ref[me.own_lo:me.own_hi])                    # own shard
        if mode == "adaptive":
            assert r["optimal_T"] == o["optimal_T"] and r["trace"] == o["trace"] and r["pipeline_dtype"] == o["pipeline_dtype"]
        if rank == 0:
            assert r["full"].shape[0] == total
            assert same(r["full"].numpy(), ref)                                    # gathered file
        else:
            assert r["full"] is None
        assert r["comm_bytes"] > 0
        dist.barrier()
    finally:
        dist.destroy_process_group()


CASES = {
    "standard_chunk_aligned": ("standard", 48000, 11.0, dict(gate_ui=50)),                       # 3 chunks -> split on a chunk boundary
    "standard_chunk_straddles_ranks": ("standard", 48000, 3.0, dict(gate_ui=50, up_delay_ms=80.0)),   # 1 chunk -> peak all-reduce decides
    "xfade_padend0": ("xfade", 96000, 1.0, dict(gate_ui=60, xfade_ms=100.0, up_delay_ms=40.0, trim=2048 * 40)),
    "adaptive_f32": ("adaptive", 48000, 3.0, dict(min_hold_ms=100.0, xfade_ms=200.0)),
    "adaptive_f64": ("adaptive", 44100, 2.0, dict(peak=0.1)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_two_ranks_match_whole_file_oracle(name):
    mode, sr, seconds, kw = CASES[name]
    mp.spawn(_worker, args=(2, _free_port(), (mode, sr, seconds, dict(kw))), nprocs=2, join=True)


def test_shard_plan_properties():
    from tomatis_audio_processor_b200 import sharded, tables as tb
    for total in (691_200_000, 28_800_000, 2_646_000, 100_000, 5000, 2048, 1, 0):
        for world in (1, 2, 3, 8):
            for framing in (sharded.STREAMING, sharded.WHOLEFILE):
                sh = sharded.plan_shards(total, world, framing)
                assert sh[0].own_lo == 0 and sh[-1].own_hi == total
                assert sh[0].frame_lo == 0 and sh[-1].frame_hi == sh[0].n_frames
                for a, b in zip(sh, sh[1:]):
                    assert a.own_hi == b.own_lo and a.block_hi == b.block_lo and a.frame_hi == b.frame_lo
                for s in sh:
                    assert s.in_lo <= s.own_lo <= s.own_hi <= s.in_hi
                    assert s.own_lo - s.in_lo <= tb.HOP and s.in_hi - s.own_hi <= tb.HOP      # one-hop halos
    # 2 h @ 96 kHz over 8 GPUs: every cut is a limiter-chunk boundary (multiples of 118 blocks)
    sh = sharded.plan_shards(691_200_000, 8, sharded.STREAMING)
    assert all(s.block_lo % 118 == 0 for s in sh) and sh[0].n_frames == 337_500
    assert tb.flush_chunk_blocks(337_500)[0] == (0, 118) and len(tb.flush_chunk_blocks(337_500)) == 2861


def test_geometry_helpers_match_oracle():
    from oracle import tomatis_oracle as orc
    from tomatis_audio_processor_b200 import tables as tb
    for total in (0, 1, 100, 2047, 2048, 2049, 4096, 6144, 10000, 262144, 2_646_000):
        _, _, starts = orc.frame_layout_streaming(total, 4096, 2048)
        assert tb.streaming_frame_count(total) == len(starts)
        sched = orc.flush_schedule(len(starts), 4096, 2048)
        blocks = tb.flush_chunk_blocks(len(starts))
        assert [(-2048 + a * 2048, -2048 + b * 2048) for a, b in blocks] == sched
