#!/usr/bin/env python3
"""One long file on several B200s -- the time-chunk sharding of SURVEY.md section 8e behind the single-file command line.

    torchrun --nproc-per-node 8 -m tomatis_audio_processor_b200.process_sharded --mode standard -i long.flac -o out.flac --gate_ui 50

`--mode` picks the front end (`process_tomatis.py` / `_xfade.py` / `_adaptive.py`); every other flag is that front end's
own and means the same.  Each rank reads ONLY its own sample range of the input (`audio_io.read_range`), cut at the
reference's limiter-chunk boundaries; the halo hand-off, the exact level gather with the redundant gate scan, the peak
all-reduce and the final gather to rank 0 are `sharded.run_streaming_sharded` / `run_adaptive_sharded` over NCCL (gloo when
there is no GPU, used by the CPU tests with a stand-in backend).  Rank 0 writes the output file (and the state CSV) exactly
like the single-file front end; the samples are bit-identical to the one-GPU result.  With one process it simply calls the
single-file front end.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from . import audio_io, report, tables as tb
from . import process_batch, process_tomatis

MODES = process_batch.MODES


def build_parser():
    ap = argparse.ArgumentParser(description="Tomatis processing of one long file, time-sharded over the GPUs (B200)",
                                 epilog="remaining flags are those of the mode's single-file command line", add_help=False)
    ap.add_argument("--mode", choices=sorted(MODES), default="standard")
    return ap


def process_sharded(mode: str, in_path: str, out_path: str, comm, device_index: int = 0, state_csv_path=None,
                    any_sr: bool = False, make_backend=None, log=print, **params):
    """Run `mode` on `in_path` over comm.world ranks.  Returns the sharded driver's result dict; rank 0 has written the files."""
    import torch
    from . import sharded
    info = audio_io.info(in_path)
    sr, total, ch = info.samplerate, info.frames, info.channels
    if mode != "adaptive" and not any_sr and sr != 48000:
        raise ValueError(f"expected 48 kHz, got {sr} Hz")                       # src/process_tomatis.py:234-235
    if ch != 2:
        if mode != "adaptive":
            raise ValueError(f"expected stereo, got {ch} channel(s)")           # :236-237
        raise NotImplementedError(f"the sharded adaptive path takes stereo files, got {ch} channel(s); use one GPU")
    if total == 0:
        raise ValueError("empty input file")
    n_fft, hop = params.get("n_fft", tb.N_FFT), params.get("hop", tb.HOP)
    from .engine import fused_size
    if not fused_size(n_fft, hop):                         # checked BEFORE any collective: every rank takes the same exit
        raise NotImplementedError(f"the time-sharded path implements n_fft/hop = 4096/2048 and 2048/1024 (got {n_fft}/{hop}); "
                                  f"other sizes run on one GPU (process_tomatis*.py)")
    framing = sharded.WHOLEFILE if mode == "adaptive" else sharded.STREAMING
    me = sharded.plan_shards(total, comm.world, framing, n_fft, hop)[comm.rank]
    # a rank whose read fails must not leave the others waiting in the first collective: agree on success first
    own, err = None, None
    try:
        own = torch.from_numpy(np.ascontiguousarray(audio_io.read_range(in_path, me.own_lo, me.own_hi, dtype="float32"))).to(comm.device)
    except Exception as e:                                  # noqa: BLE001 - re-raised on every rank below
        err = e
    n_bad = int(comm.allreduce(np.array([0.0 if err is None else 1.0], np.float32), "sum")[0])
    if n_bad:
        raise RuntimeError(f"{n_bad} rank(s) could not read their sample range of {in_path}" + (f": {err}" if err is not None else ""))
    log(f"samples [{me.own_lo}, {me.own_hi}) of {total} ({(me.own_hi - me.own_lo) / sr:.1f} s of {total / sr:.1f} s), "
        f"output blocks [{me.block_lo}, {me.block_hi})")
    run = sharded.run_adaptive_sharded if mode == "adaptive" else (lambda *a, **k: sharded.run_streaming_sharded(mode, *a, **k))
    r = run(own, sr, total, comm, make_backend=make_backend, device_index=device_index, gather_to=0, **params)
    if comm.rank == 0:
        y = r["full"].cpu().numpy()
        written, _ = process_tomatis._write_output(out_path, y, sr)
        if state_csv_path:
            report.write_state_csv(state_csv_path, mode, r)
        st = report.gate_statistics(r["states"], total, sr)
        peaks = np.asarray(r["chunk_peaks"]) if mode != "adaptive" else np.asarray([r["output_peak"]], np.float32)
        print(f"[OK] {written}: {total / sr:.2f} s @ {sr} Hz on {comm.world} GPU(s), {st['frames']} frames, "
              f"C2 {st['c2_ratio'] * 100:.1f}%, {int((peaks > np.float32(tb.PEAK_LIMIT)).sum())}/{len(peaks)} limiter chunks scaled, "
              f"{r['comm_bytes']} B sent by this rank", flush=True)
    return r


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    args, rest = build_parser().parse_known_args(argv)
    mod = MODES[args.mode]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return mod.main(rest) or 0
    m = mod.build_parser().parse_args(rest)
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    from . import sharded
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}")) if use_cuda else dist.init_process_group("gloo")
    comm = sharded.Comm(None, f"cuda:{local}" if use_cuda else "cpu")
    log = lambda s: print(f"[rank {rank}/{world}] {s}", flush=True)
    rc = 0
    try:
        process_sharded(args.mode, m.input, m.output, comm, device_index=local, state_csv_path=getattr(m, "state_csv", None),
                        any_sr=getattr(m, "any_sr", False), log=log, **process_batch.engine_kwargs(args.mode, m))
    except Exception as e:                                  # like main() of src/process_tomatis.py:519-544
        print(f"\n[ERR] rank {rank}: {e}", flush=True)
        import traceback
        traceback.print_exc()
        rc = 1
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
