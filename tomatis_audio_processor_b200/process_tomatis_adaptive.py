#!/usr/bin/env python3
"""Adaptive mode -- drop-in for the reference's `src/process_tomatis_adaptive.py`.

Same `process()` signature, defaults and return value 0 (src/process_tomatis_adaptive.py:157-172,373), same
CLI (no --gate_ui; :376-399), no sample-rate guard and no try/except in main() (errors propagate, as in the
reference).  Whole file in HBM; input peak, per-frame levels, every gate simulation of the threshold
bisection, the final gate + alpha counter, STFT/OLA, restore gain and the global limiter run on the device
(engine.run_adaptive); percentiles and the bisection bookkeeping stay on the host.

Mono files are accepted like in the reference (:180-181; they ride in the L lane of the stereo kernels).  Files with more
than two channels (the reference's `for c in range(ch)`, :307-313) are cut into channel pairs on the device; the pairs share
input peak, frame level, gate and limiter scale, all taken over every channel like the reference does
(engine.run_adaptive_multichannel; default n_fft / hop only).
"""
from __future__ import annotations

import argparse
import sys

import numpy as np

from . import audio_io, report, tables as tb

DEVICE = 0


def process(
    in_path,
    out_path,
    fc=1000.0,
    slope=12.0,
    c1_low=15.0, c1_high=-15.0,
    c2_low=-15.0, c2_high=15.0,
    target_c2=0.5,
    hyst_db=3.0,
    min_hold_ms=250.0,
    xfade_ms=500.0,
    headroom_margin=2.0,
    n_fft=4096,
    hop=2048,
    state_csv_path=None
):
    from . import engine

    print("=" * 60)
    print("Tomatis adaptive processor (B200)")
    print("=" * 60)
    print(f"\nreading: {in_path}")
    x, sr = audio_io.read(in_path, dtype="float32")
    ch = x.shape[1]
    total = len(x)
    print(f"  sample rate: {sr} Hz\n  channels: {ch}\n  duration: {total / sr:.2f} s")
    if total == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")   # np.max(np.abs(x)), :201

    hold, xf = tb.adaptive_frame_counts(sr, min_hold_ms, xfade_ms, hop)
    frame_ms = hop / sr * 1000
    print(f"\ngate: hyst {hyst_db} dB, min_hold {min_hold_ms} ms ({hold} frames), xfade {xfade_ms} ms ({xf} frames), "
          f"frame {frame_ms:.2f} ms")

    res = engine.run_adaptive(
        [x], sr, device=DEVICE, fc=fc, slope=slope, c1_low=c1_low, c1_high=c1_high, c2_low=c2_low, c2_high=c2_high,
        target_c2=target_c2, hyst_db=hyst_db, min_hold_ms=min_hold_ms, xfade_ms=xfade_ms,
        headroom_margin=headroom_margin, n_fft=n_fft, hop=hop)[0]

    with np.errstate(divide="ignore"):
        in_peak_db = 20 * np.log10(np.float32(res["input_peak"]) + tb.EPS)
    print(f"\npre-attenuation: input peak {in_peak_db:.2f} dBFS, max gain +{max(abs(c1_low), abs(c2_high))} dB, "
          f"margin {headroom_margin} dB -> {-res['atten_db']:.2f} dB ({res['pipeline_dtype']} pipeline)")
    levels, states = res["levels"], res["states"]
    n = len(levels)
    valid = int((levels > -70).sum())
    print(f"\nadaptive gate:\n  frames: {n}\n  valid frames: {valid} ({valid / n * 100:.1f}%)")   # ZeroDivisionError below one hop (:222)
    st = report.gate_statistics(states, total, sr, hold)
    y = res["out"]
    out_peak = float(np.max(np.abs(y))) if y.size else 0.0
    if res["atten_db"] > 0:
        print(f"  restored pre-attenuation: +{res['atten_db']:.2f} dB")
    if res["output_peak"] > tb.PEAK_LIMIT:
        print(f"  peak protection: scaled by {20 * np.log10(tb.PEAK_LIMIT / res['output_peak']):.2f} dB")

    audio_io.write(out_path, y, sr, subtype="PCM_24")
    print(f"\noutput saved: {out_path}")
    if state_csv_path:
        report.write_state_csv(state_csv_path, "adaptive", res)
        print(f"state CSV saved: {state_csv_path}")

    print("\nstatistics:")
    print(f"  pre-attenuation: {-res['atten_db']:.2f} dB")
    print(f"  optimal threshold T: {res['optimal_T']:.2f} dBFS")
    print(f"  C2 ratio: {st['c2_ratio'] * 100:.1f}%")
    print(f"  switches: {st['switches']} ({st['switches_per_min']:.1f}/min)")
    print(f"  short-run ratio: {st['short_run_ratio'] * 100:.1f}%")
    print(f"  output peak: {20 * np.log10(out_peak + tb.EPS):.2f} dBFS")
    return 0


def build_parser():
    p = argparse.ArgumentParser(description="Tomatis adaptive processor (B200)")
    p.add_argument("-i", "--input", required=True, help="input audio")
    p.add_argument("-o", "--output", required=True, help="output audio")
    p.add_argument("--state_csv", help="state CSV output path")
    p.add_argument("--fc", type=float, default=1000)
    p.add_argument("--slope", type=float, default=12)
    p.add_argument("--c1_low", type=float, default=15.0)
    p.add_argument("--c1_high", type=float, default=-15.0)
    p.add_argument("--c2_low", type=float, default=-15.0)
    p.add_argument("--c2_high", type=float, default=15.0)
    p.add_argument("--target_c2", type=float, default=0.5, help="target C2 ratio")
    p.add_argument("--hyst_db", type=float, default=3.0, help="hysteresis dB")
    p.add_argument("--min_hold_ms", type=float, default=250.0, help="minimum hold ms")
    p.add_argument("--xfade_ms", type=float, default=500.0, help="crossfade time ms")
    p.add_argument("--headroom_margin", type=float, default=2.0, help="pre-attenuation margin dB")
    p.add_argument("--n_fft", type=int, default=4096)
    p.add_argument("--hop", type=int, default=2048)
    p.add_argument("--device", type=int, default=0, help="extension: CUDA device index")
    return p


def main(argv=None):
    global DEVICE
    a = build_parser().parse_args(argv)
    DEVICE = a.device
    return process(a.input, a.output, fc=a.fc, slope=a.slope, c1_low=a.c1_low, c1_high=a.c1_high, c2_low=a.c2_low,
                   c2_high=a.c2_high, target_c2=a.target_c2, hyst_db=a.hyst_db, min_hold_ms=a.min_hold_ms,
                   xfade_ms=a.xfade_ms, headroom_margin=a.headroom_margin, n_fft=a.n_fft, hop=a.hop,
                   state_csv_path=a.state_csv)


if __name__ == "__main__":
    sys.exit(main())
