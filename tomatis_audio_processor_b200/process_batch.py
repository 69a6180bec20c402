#!/usr/bin/env python3
"""Many files in one go -- the batch loop of the reference's usage guide (`docs/Tomatis处理器使用指南.md:243-249`: a
PowerShell ForEach over `process_tomatis.py -i file -o file_tomatis.flac`), on one or several B200s.

    python -m tomatis_audio_processor_b200.process_batch --mode standard -i a.flac b.flac ... --out_dir out --gate_ui 50
    torchrun --nproc-per-node 8 -m tomatis_audio_processor_b200.process_batch --mode adaptive -i *.flac --out_dir out

Every flag the single-file front end of the chosen mode takes (`process_tomatis*.py`) is accepted after the batch flags
and means the same.  Partitioning is the one of SURVEY.md section 8e: whole tracks, file k goes to rank k % world, every
rank runs the one-GPU pipeline on its files (tracks of equal sample rate share one plan: the frames of many tracks in each
launch), no collective on the data path; the per-file statistics are summed to every rank once at the end (one all-reduce
of a zero-filled table).  Like the ForEach loop, a file that fails (unreadable, wrong sample rate in standard / xfade mode
without --any_sr, ...) does not stop the others; the exit code is 1 if any file failed.  Outputs are written with the
rule of `src/process_tomatis.py:242-251` in every mode (FLAC PCM_24, else WAV PCM_24 next to it); file names are
`<stem><suffix>.flac` in --out_dir.
"""
from __future__ import annotations

import argparse
import glob
import os
import sys
import time

import numpy as np

from . import audio_io, report, tables as tb
from . import process_tomatis, process_tomatis_adaptive, process_tomatis_xfade

MODES = {"standard": process_tomatis, "xfade": process_tomatis_xfade, "adaptive": process_tomatis_adaptive}
# sample-frames of input one engine call may hold (input + output float32 = 16 B each): 1.5 G -> 24 GB of the 180 GB
WAVE_SAMPLE_FRAMES = 1_500_000_000
STAT_COLS = ("ok", "sample_rate", "samples", "frames", "c2_frames", "chunks", "chunks_limited")


def engine_kwargs(mode: str, args) -> dict:
    """Parsed flags of the mode's own command line -> keyword arguments of engine.run, as its process() passes them."""
    tilt = dict(fc=args.fc, slope=args.slope, c1_low=args.c1_low, c1_high=args.c1_high, c2_low=args.c2_low,
                c2_high=args.c2_high, n_fft=args.n_fft, hop=args.hop)
    if mode == "standard":
        return dict(tilt, gate_ui=args.gate_ui, gate_mode=args.gate_mode, dynamic_range=args.dynamic_range,
                    gate_scale=args.gate_scale, gate_offset=args.gate_offset, hysteresis_db=args.hyst_db,
                    up_delay_ms=args.up_delay_ms, output_gain_db=args.output_gain_db)
    if mode == "xfade":
        return dict(tilt, gate_ui=args.gate_ui, gate_scale=args.gate_scale, gate_offset=args.gate_offset,
                    hysteresis_db=args.hyst_db, up_delay_ms=args.up_delay_ms, xfade_ms=args.xfade_ms)
    return dict(tilt, target_c2=args.target_c2, hyst_db=args.hyst_db, min_hold_ms=args.min_hold_ms, xfade_ms=args.xfade_ms,
                headroom_margin=args.headroom_margin)


def assignment(n_files: int, rank: int, world: int):
    """File indices of one rank: track k -> rank k % world (batch row)."""
    return list(range(rank, n_files, world))


def output_path(in_path: str, out_dir: str, suffix: str) -> str:
    stem = os.path.splitext(os.path.basename(in_path))[0]
    return os.path.join(out_dir, stem + suffix + ".flac")


def _load(mode: str, path: str, any_sr: bool):
    """Read one file and apply the mode's input rules (src/process_tomatis.py:234-237, _xfade.py:106-109; adaptive takes
    any rate and mono, _adaptive.py:180-181)."""
    x, sr = audio_io.read(path, dtype="float32")
    if mode != "adaptive":
        if not any_sr and sr != 48000:
            raise ValueError(f"expected 48 kHz, got {sr} Hz")
        if x.shape[1] != 2:
            raise ValueError(f"expected stereo, got {x.shape[1]} channel(s)")
    elif x.shape[1] > 2:
        raise NotImplementedError(f"adaptive mode on the GPU takes mono or stereo files, got {x.shape[1]} channels")
    elif len(x) == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")   # _adaptive.py:201
    return x, sr


def process_files(mode: str, in_paths, out_paths, csv_paths=None, device: int = 0, any_sr: bool = False,
                  wave_sample_frames: int = WAVE_SAMPLE_FRAMES, log=print, **params):
    """Process `in_paths` -> `out_paths` on one GPU.  Returns one dict per file: ok, error, written, and the STAT_COLS."""
    from . import engine
    n = len(in_paths)
    csv_paths = csv_paths or [None] * n
    results = [dict(ok=False, error=None, written=None, **{c: 0 for c in STAT_COLS[1:]}) for _ in range(n)]
    loaded = {}
    for i, path in enumerate(in_paths):
        try:
            loaded[i] = _load(mode, path, any_sr)
        except Exception as e:                                      # keep going, like the shell loop
            results[i]["error"] = f"{type(e).__name__}: {e}"
            log(f"[ERR] {path}: {results[i]['error']}")
    # waves: files of one sample rate and channel count, in input order, up to the sample-frame budget
    by_sr = {}
    for i, (x, sr) in loaded.items():
        by_sr.setdefault((sr, x.shape[1]), []).append(i)
    for (sr, _), idx in sorted(by_sr.items()):
        wave, used = [], 0
        for i in idx + [None]:
            size = len(loaded[i][0]) if i is not None else 0
            if wave and (i is None or used + size > wave_sample_frames):
                _run_wave(engine, mode, wave, loaded, sr, device, params, results, out_paths, csv_paths, log)
                wave, used = [], 0
            if i is not None:
                wave.append(i)
                used += size
    return results


def _run_wave(engine, mode, wave, loaded, sr, device, params, results, out_paths, csv_paths, log):
    t0 = time.perf_counter()
    try:
        res = engine.run(mode, [loaded[i][0] for i in wave], sr, device=device, **params)
    except Exception as e:
        for i in wave:
            results[i]["error"] = f"{type(e).__name__}: {e}"
        log(f"[ERR] {len(wave)} file(s) at {sr} Hz: {type(e).__name__}: {e}")
        return
    dt = time.perf_counter() - t0
    for i, r in zip(wave, res):
        x = loaded.pop(i)[0]
        try:
            written, _ = process_tomatis._write_output(out_paths[i], r["out"], sr)
            if csv_paths[i]:
                report.write_state_csv(csv_paths[i], mode, r)
            st = report.gate_statistics(r["states"], len(x), sr)
            peaks = np.asarray(r["chunk_peaks"]) if "chunk_peaks" in r else np.asarray([r.get("output_peak", 0.0)], np.float32)
            results[i].update(ok=True, written=written, sample_rate=int(sr), samples=int(len(x)), frames=st["frames"],
                              c2_frames=st["c2_frames"], chunks=len(r["chunk_lengths"]),
                              chunks_limited=int((peaks > np.float32(tb.PEAK_LIMIT)).sum()))
            log(f"[OK] {written}: {len(x) / sr:.2f} s @ {sr} Hz, {st['frames']} frames, C2 {st['c2_ratio'] * 100:.1f}%, "
                f"{results[i]['chunks_limited']}/{results[i]['chunks']} limiter chunks scaled")
        except Exception as e:
            results[i]["error"] = f"{type(e).__name__}: {e}"
            log(f"[ERR] {out_paths[i]}: {results[i]['error']}")
    log(f"[WAVE] {len(wave)} file(s) at {sr} Hz in {dt * 1e3:.1f} ms on device {device}")


def stats_table(results, mine, n_files: int) -> np.ndarray:
    """Zero-filled [n_files, len(STAT_COLS)] int64 table with this rank's rows filled in (summed over ranks afterwards)."""
    t = np.zeros((n_files, len(STAT_COLS)), dtype=np.int64)
    for k, r in zip(mine, results):
        t[k] = [int(r[c]) for c in STAT_COLS]
    return t


def reduce_table(table: np.ndarray, world: int, device: int) -> np.ndarray:
    """Sum of every rank's table on every rank (NCCL when CUDA is there, gloo otherwise); the only collective of the run."""
    if world == 1:
        return table
    import torch
    import torch.distributed as dist
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(device)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{device}")) if use_cuda else dist.init_process_group("gloo")
    t = torch.from_numpy(table.copy())
    if use_cuda:
        t = t.to(f"cuda:{device}")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def build_parser():
    ap = argparse.ArgumentParser(description="Tomatis processing of many files, tracks sharded over the GPUs (B200)",
                                 epilog="remaining flags are those of the mode's single-file command line")
    ap.add_argument("--mode", choices=sorted(MODES), default="standard")
    ap.add_argument("-i", "--input", nargs="+", required=True, help="input files (glob patterns are expanded)")
    ap.add_argument("--out_dir", required=True)
    ap.add_argument("--suffix", default="_tomatis", help="appended to the input's stem")
    ap.add_argument("--state_csv_dir", default=None, help="write <stem>_state.csv per file here")
    ap.add_argument("--wave_sample_frames", type=int, default=WAVE_SAMPLE_FRAMES, help="input sample-frames per engine call")
    return ap


def main(argv=None):
    args, rest = build_parser().parse_known_args(argv)
    mode_args = MODES[args.mode].build_parser().parse_args(["-i", "_", "-o", "_"] + rest)
    files = []
    for pat in args.input:
        hits = sorted(glob.glob(pat))
        files.extend(hits if hits else [pat])
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    device = int(os.environ.get("LOCAL_RANK", getattr(mode_args, "device", 0) or 0))
    if rank == 0:
        os.makedirs(args.out_dir, exist_ok=True)
        if args.state_csv_dir:
            os.makedirs(args.state_csv_dir, exist_ok=True)
    mine = assignment(len(files), rank, world)
    tag = f"[rank {rank}/{world}] " if world > 1 else ""
    log = lambda s: print(tag + s, flush=True)
    log(f"{len(mine)} of {len(files)} file(s), mode {args.mode}, device {device}")
    if world > 1:
        reduce_table(np.zeros((1, 1), np.int64), world, device)      # rendezvous first: rank 0 has made the directories
    stem = lambda p: os.path.splitext(os.path.basename(p))[0]
    results = process_files(
        args.mode, [files[k] for k in mine], [output_path(files[k], args.out_dir, args.suffix) for k in mine],
        [os.path.join(args.state_csv_dir, stem(files[k]) + "_state.csv") if args.state_csv_dir else None for k in mine],
        device=device, any_sr=getattr(mode_args, "any_sr", False), wave_sample_frames=args.wave_sample_frames, log=log,
        **engine_kwargs(args.mode, mode_args))
    table = reduce_table(stats_table(results, mine, len(files)), world, device)
    if rank == 0:
        ok = table[:, 0] == 1
        secs = float(np.sum(table[ok, 2] / np.maximum(table[ok, 1], 1)))
        print(f"[DONE] {int(ok.sum())} of {len(files)} file(s) processed on {world} GPU(s): {secs:.1f} s of audio, "
              f"{int(table[:, 3].sum())} frames, C2 {100.0 * table[:, 4].sum() / max(1, table[:, 3].sum()):.1f}%, "
              f"{int(table[:, 6].sum())}/{int(table[:, 5].sum())} limiter chunks scaled")
        for k in np.flatnonzero(~ok):
            print(f"[FAILED] {files[k]}")
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    return 0 if bool((table[:, 0] == 1).all()) else 1


if __name__ == "__main__":
    sys.exit(main())
