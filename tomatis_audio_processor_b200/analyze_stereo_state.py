#!/usr/bin/env python3
"""Per-channel state analyser -- drop-in for the reference's `src/analyze_stereo_state.py`.

Same `analyze(in_path, out_csv, target_c2, hyst_db, min_hold_ms, n_fft, hop)` signature and return codes (1 for a
single-channel file, :83-85), same CLI flags (:163-171), same CSV: header (:135-143) and number formats (:146-154)
byte for byte.  Levels and every gate simulation run in the CUDA library through engine.run_channel_states; there is
no CPU path.
"""
from __future__ import annotations

import argparse
import csv
import sys

import numpy as np

from . import audio_io

DEVICE = 0

CSV_HEADER = ["Frame", "音频秒数(秒)", "音频时间(分:秒)", "Left_dBFS", "Left_Channel", "Right_dBFS", "Right_Channel"]


def format_time(seconds) -> str:
    """minutes:seconds as the reference prints it (src/analyze_stereo_state.py:22-26)"""
    m = int(seconds // 60)
    s = seconds % 60
    return f"{m}:{s:05.2f}"


def csv_rows(res: dict):
    """One result of engine.run_channel_states -> the rows csv.writer gets in the reference (:135-154)."""
    name = {1: "C1", 2: "C2"}
    rows = [list(CSV_HEADER)]
    for i, t in enumerate(res["times"]):
        t = float(t)
        rows.append([i + 1, f"{t:.3f}", format_time(t), f"{float(res['left_levels'][i]):.2f}",
                     name[int(res["left_states"][i])], f"{float(res['right_levels'][i]):.2f}",
                     name[int(res["right_states"][i])]])
    return rows


def analyze(in_path, out_csv, target_c2=0.5, hyst_db=3.0, min_hold_ms=250.0, n_fft=4096, hop=2048):
    from . import engine

    print(f"reading: {in_path}")
    x, sr = audio_io.read(in_path, dtype="float32")
    if x.shape[1] == 1:
        print("error: the input is mono, a stereo file is needed")
        return 1
    ch = x.shape[1]
    print(f"sample rate: {sr} Hz, channels: {ch}")
    res = engine.run_channel_states([np.ascontiguousarray(x[:, :2])], sr, device=DEVICE, target_c2=target_c2, hyst_db=hyst_db,
                                    min_hold_ms=min_hold_ms, n_fft=n_fft, hop=hop)[0]
    n = len(res["times"])
    print(f"frames: {n}")
    if n == 0:
        raise ZeroDivisionError("division by zero")            # C2 share of no frames, like :126
    print(f"left : T={res['left_T']:.2f} dBFS, C2={res['left_c2'] * 100:.1f}%")
    print(f"right: T={res['right_T']:.2f} dBFS, C2={res['right_c2'] * 100:.1f}%")
    print(f"writing: {out_csv}")
    with open(out_csv, "w", newline="", encoding="utf-8") as f:
        csv.writer(f).writerows(csv_rows(res))
    print("done")
    return 0


def build_parser():
    ap = argparse.ArgumentParser(description="stereo state analyser: per-channel dBFS and C1/C2 state (B200)")
    ap.add_argument("-i", "--input", required=True, help="input audio")
    ap.add_argument("-o", "--output", required=True, help="output CSV")
    ap.add_argument("--target_c2", type=float, default=0.5)
    ap.add_argument("--hyst_db", type=float, default=3.0)
    ap.add_argument("--min_hold_ms", type=float, default=250.0)
    ap.add_argument("--device", type=int, default=0, help="extension: CUDA device index")
    return ap


def main(argv=None):
    global DEVICE
    args = build_parser().parse_args(argv)
    DEVICE = args.device
    return analyze(args.input, args.output, args.target_c2, args.hyst_db, args.min_hold_ms)


if __name__ == "__main__":
    sys.exit(main())
