#!/usr/bin/env python3
"""Xfade mode -- drop-in for the reference's `src/process_tomatis_xfade.py`.

Same `process()` signature and defaults (src/process_tomatis_xfade.py:55-71: linear gate map only,
`xfade_ms=0.0` = hard switching, no output gain), same CLI flags (:368-391), same 48 kHz / stereo
ValueError (:106-109), FLAC PCM_24 output with WAV fallback (:113-123), state CSV with the alpha column
(:180,293-295).  The slew-limited alpha follower and the dB-domain gain mix (:251-274) run on the device as
an integer crossfade counter that indexes host-built gain rows (tables.gain_rows_xfade).
"""
from __future__ import annotations

import argparse
import sys

import numpy as np

from . import report, tables as tb
from . import process_tomatis as _std

REFERENCE_GUARDS = True
DEVICE = 0


def process(
    in_path,
    out_path,
    gate_ui=50,
    gate_scale=1.0,
    gate_offset=-100,
    hysteresis_db=3.0,
    fc=1000.0,
    slope=12.0,
    c1_low=+15.0, c1_high=-15.0,
    c2_low=-15.0, c2_high=+15.0,
    up_delay_ms=250.0,
    xfade_ms=0.0,
    n_fft=4096,
    hop=2048,
    state_csv_path=None,
):
    """Gate-controlled C1/C2 tilt filter with crossfaded transitions (B200 path)."""
    from . import engine

    print("=" * 70)
    print("Tomatis audio processor with crossfade (B200)")
    print("=" * 70)
    print(f"\ninput : {in_path}\noutput: {out_path}\n")
    T = tb.gate_threshold_linear(gate_ui, gate_scale, gate_offset)
    Ton, Toff = tb.hysteresis_pair(T, hysteresis_db)
    print(f"gate: ui={gate_ui} T={T:.1f} dBFS  up {Ton:.1f}  down {Toff:.1f}  hysteresis {hysteresis_db} dB  "
          f"up-delay {up_delay_ms} ms")
    print(f"crossfade: {xfade_ms} ms" + (" (hard switching)" if xfade_ms <= 0 else ""))
    print(f"tilt: fc={fc} Hz slope={slope} dB/oct  C1 {c1_low:+.1f}/{c1_high:+.1f} dB  C2 {c2_low:+.1f}/{c2_high:+.1f} dB")
    print(f"stft: n_fft={n_fft} hop={hop}\n")

    saved = _std.REFERENCE_GUARDS
    _std.REFERENCE_GUARDS = REFERENCE_GUARDS
    try:
        x, sr = _std._open_check(in_path, "xfade")
    finally:
        _std.REFERENCE_GUARDS = saved
    total = len(x)

    res = engine.run_streaming(
        "xfade", [x], sr, device=DEVICE, gate_ui=gate_ui, gate_scale=gate_scale, gate_offset=gate_offset,
        hysteresis_db=hysteresis_db, fc=fc, slope=slope, c1_low=c1_low, c1_high=c1_high, c2_low=c2_low,
        c2_high=c2_high, up_delay_ms=up_delay_ms, xfade_ms=xfade_ms, n_fft=n_fft, hop=hop)[0]

    written, _ = _std._write_output(out_path, res["out"], sr)
    if state_csv_path:
        report.write_state_csv(state_csv_path, "xfade", res)

    st = report.gate_statistics(res["states"], total, sr)
    n = st["frames"]
    print("\n" + "=" * 70 + "\ndone\n" + "=" * 70)
    print(f"frames: {n}")
    print(f"  C1: {st['c1_frames']} ({st['c1_frames'] / n * 100:.1f}%)")
    print(f"  C2: {st['c2_frames']} ({st['c2_frames'] / n * 100:.1f}%)")
    if xfade_ms > 0:
        print(f"  crossfade: {xfade_ms} ms ({res['xfade_frames']} frames)")
    limited = int((res["chunk_peaks"] > np.float32(tb.PEAK_LIMIT)).sum())
    print(f"limiter chunks: {len(res['chunk_lengths'])} ({limited} scaled to {tb.PEAK_LIMIT})")
    print(f"output: {written}")
    if state_csv_path:
        print(f"state CSV: {state_csv_path}")
    print()


def build_parser():
    ap = argparse.ArgumentParser(description="Tomatis audio processor - gate-controlled C1/C2 tilt filter with crossfade (B200)",
                                 formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("-i", "--input", required=True, help="input FLAC/WAV file")
    ap.add_argument("-o", "--output", required=True, help="output FLAC file")
    ap.add_argument("--gate_ui", type=float, default=50, help="gate UI value (0-100)")
    ap.add_argument("--gate_scale", type=float, default=1.0, help="gate scale")
    ap.add_argument("--gate_offset", type=float, default=-100, help="gate offset")
    ap.add_argument("--hyst_db", type=float, default=3.0, help="hysteresis (dB)")
    ap.add_argument("--up_delay_ms", type=float, default=250.0, help="C1->C2 up-delay (ms)")
    ap.add_argument("--xfade_ms", type=float, default=0.0, help="crossfade time (ms), 0 = hard switching")
    ap.add_argument("--fc", type=float, default=1000.0, help="pivot frequency (Hz)")
    ap.add_argument("--slope", type=float, default=12.0, help="slope (dB/octave)")
    ap.add_argument("--c1_low", type=float, default=15.0, help="C1 low-frequency gain (dB)")
    ap.add_argument("--c1_high", type=float, default=-15.0, help="C1 high-frequency gain (dB)")
    ap.add_argument("--c2_low", type=float, default=-15.0, help="C2 low-frequency gain (dB)")
    ap.add_argument("--c2_high", type=float, default=15.0, help="C2 high-frequency gain (dB)")
    ap.add_argument("--n_fft", type=int, default=4096, help="FFT length")
    ap.add_argument("--hop", type=int, default=2048, help="hop length")
    ap.add_argument("--state_csv", default=None, help="per-frame state CSV path")
    ap.add_argument("--any_sr", action="store_true", help="extension: lift the reference's 48 kHz guard")
    ap.add_argument("--device", type=int, default=0, help="extension: CUDA device index")
    return ap


def main(argv=None):
    global REFERENCE_GUARDS, DEVICE
    args = build_parser().parse_args(argv)
    if args.any_sr:
        REFERENCE_GUARDS = False
    DEVICE = args.device
    try:
        process(args.input, args.output, gate_ui=args.gate_ui, gate_scale=args.gate_scale,
                gate_offset=args.gate_offset, hysteresis_db=args.hyst_db, fc=args.fc, slope=args.slope,
                c1_low=args.c1_low, c1_high=args.c1_high, c2_low=args.c2_low, c2_high=args.c2_high,
                up_delay_ms=args.up_delay_ms, xfade_ms=args.xfade_ms, n_fft=args.n_fft, hop=args.hop,
                state_csv_path=args.state_csv)
    except Exception as e:
        print(f"\n[ERR] {e}")
        import traceback
        traceback.print_exc()
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
