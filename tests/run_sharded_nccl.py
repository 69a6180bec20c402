"""torchrun entry (>= 2 GPUs): the sharded long-file path over NCCL against the oracle.  Launched by
tests/test_gpu_sharded.py::test_nccl_two_gpus_torchrun or by hand:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_sharded_nccl.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import tomatis_oracle as orc                       # noqa: E402  (checker)
from tomatis_audio_processor_b200 import sharded, synth        # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    comm = sharded.Comm(None, f"cuda:{local}")
    n_fft = int(os.environ.get("TMT_TEST_NFFT", "4096"))       # 2048: the second build of the library (hop 1024)
    sz = dict(n_fft=n_fft, hop=n_fft // 2) if n_fft != 4096 else {}
    cases = [("standard", 48000, synth.recipe_gated_pink(11.0, 48000, 30, env_hz=0.9, hi_dbfs=-22.0), dict(gate_ui=50)),
             ("xfade", 48000, synth.recipe_threshold_ramps(4.0, 48000, 3, t_on=-48.5, t_off=-51.5, period_s=1.3), dict(gate_ui=50, xfade_ms=300.0)),
             ("adaptive", 48000, synth.recipe_swept_pink(4.0, 48000, 4, period_s=1.1, peak=0.5), dict())]
    def say(msg):
        print(f"[rank {rank}] {msg}", flush=True)

    for mode, sr, x, kw in cases:
        say(f"case {mode}")
        total = len(x)
        framing = sharded.WHOLEFILE if mode == "adaptive" else sharded.STREAMING
        kw = dict(kw, **sz)
        me = sharded.plan_shards(total, world, framing, n_fft, n_fft // 2)[rank]
        own = torch.from_numpy(x[me.own_lo:me.own_hi].copy()).cuda()
        if mode == "adaptive":
            r = sharded.run_adaptive_sharded(own, sr, total, comm, device_index=local, gather_to=0, **kw)
        else:
            r = sharded.run_streaming_sharded(mode, own, sr, total, comm, device_index=local, gather_to=0, **kw)
        o = orc.run(mode, x, sr, **kw)
        o64 = orc.run(mode, x, sr, fft_dtype="float64", **kw)
        assert np.array_equal(r["meansq"], np.asarray(o["meansq"])) and np.array_equal(r["states"], o["states"])
        if rank == 0:
            y = r["full"].cpu().numpy().astype(np.float64)
            d = np.abs(y - o["out"]).max(axis=1)
            d64 = np.abs(y - o64["out"]).max()
            assert d[256:-256].max() <= 1e-5 and d64 <= 1e-5, (mode, d[256:-256].max(), d64)
            print(f"{mode}: world {world}, {total} samples, interior err {d[256:-256].max():.2e}, vs fp64-FFT {d64:.2e}, "
                  f"comm {r['comm_bytes']} B")
        dist.barrier()
    # the persistent session bench.py times, in both exchange flavours: peer memory (publish / wait kernels over CUDA IPC) and
    # collectives (all-gather halo hand-off + hop-sum all-reduce).  Three passes on one input, then fresh samples in place and two
    # more passes (a stale parity buffer or halo would show), each flavour against the oracle and against the other, bit for bit.
    x = cases[0][2]
    x2 = synth.recipe_gated_pink(11.0, 48000, 77, env_hz=1.3, hi_dbfs=-20.0)
    total = len(x)
    me = sharded.plan_shards(total, world, sharded.STREAMING, n_fft, n_fft // 2)[rank]
    outs = {}
    for flavour in ("peer", "collectives"):
        sess = sharded.StreamingShardSession("standard", torch.from_numpy(x[me.own_lo:me.own_hi].copy()).cuda(), 48000, total, comm,
                                             device_index=local, gate_ui=50, use_peer=(flavour == "peer"), **sz)
        say(f"{flavour}: session created, unfusable chunks {sess.be.plan.unfusable_chunks}, graph {sess.use_graph}, "
            f"peer exchange {sess.peer is not None}")
        assert (sess.peer is not None) == (flavour == "peer"), "peer-memory exchange not available on this box"
        for src in (x, x2):
            o = orc.run("standard", src, 48000, gate_ui=50, **sz)
            sess.own.copy_(torch.from_numpy(src[me.own_lo:me.own_hi].copy()).cuda())
            for k in range(3 if src is x else 2):
                sess.step()
                torch.cuda.synchronize()
            say(f"{flavour}: passes done (graph replay: {sess._graph is not None})")
            if sess.peer is not None:
                assert sess.peer.status() == 0, sess.peer.status()
            y = sess.out.cpu().numpy()
            d = np.abs(y.astype(np.float64) - o["out"][me.own_lo:me.own_hi]).max(axis=1)
            if rank == world - 1:
                d = d[:-256]
            assert d.max() <= 1e-5, (flavour, rank, d.max())
            outs[(flavour, src is x)] = y.copy()
        sess.close()
        dist.barrier()
    for key in (True, False):
        assert np.array_equal(outs[("peer", key)], outs[("collectives", key)]), "peer-memory and collective exchange differ"
    dist.barrier()
    if rank == 0:
        print("SHARDED-NCCL-OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
