#!/usr/bin/env python3
"""Standard mode -- drop-in for the reference's `src/process_tomatis.py`.

Same `process(in_path, out_path, ...)` signature and defaults (src/process_tomatis.py:160-178), same CLI
flags (`-i/-o/--gate_ui/--gate_mode/...`, :488-515), same error behaviour (ValueError unless 48 kHz stereo,
:234-237; main() prints the error and returns 1, :519-544), same outputs (FLAC PCM_24 with the WAV fallback
of :242-251, optional state CSV of :305,408-409).  Everything between "samples read" and "samples written"
runs in the CUDA library through engine.run_streaming -- there is no CPU path.

Extensions that do not change the reference behaviour: `--any_sr` (CLI) / `process_tomatis.REFERENCE_GUARDS =
False` lift the 48 kHz guard so the 44.1 / 96 kHz BASELINE configs can be processed; `--device`.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np

from . import audio_io, report, tables as tb

REFERENCE_GUARDS = True      # raise on sr != 48000 / ch != 2 exactly like the reference
DEVICE = 0


def _open_check(in_path, mode_name):
    x, sr = audio_io.read(in_path, dtype="float32")
    ch = x.shape[1]
    print(f"[OK] sample rate: {sr} Hz")
    print(f"[OK] channels: {ch}")
    print(f"[OK] length: {len(x)} samples ({len(x) / sr:.2f} s)")
    if REFERENCE_GUARDS and sr != 48000:
        raise ValueError(f"expected 48 kHz, got {sr} Hz")
    if ch != 2:
        raise ValueError(f"expected stereo, got {ch} channel(s)")
    return x, sr


def _write_output(out_path, y, sr):
    """FLAC PCM_24 first, else WAV PCM_24 next to it (src/process_tomatis.py:242-251). Returns the path written."""
    try:
        audio_io.write(out_path, y, sr, subtype="PCM_24", format="FLAC")
        print("[OK] output format: FLAC 24-bit")
        return out_path, True
    except Exception as e:                      # same breadth as the reference's `except Exception`
        print(f"[WARN] FLAC write failed: {e}")
        wav_path = out_path.replace(".flac", ".wav")
        audio_io.write(wav_path, y, sr, subtype="PCM_24", format="WAV")
        print("[OK] output format: WAV 24-bit (convert to FLAC afterwards)")
        return wav_path, False


def process(
    in_path,
    out_path,
    gate_ui=50,
    gate_mode="log_percent",
    dynamic_range=80.0,
    gate_scale=1.0,
    gate_offset=-100,
    hysteresis_db=3.0,
    fc=1000.0,
    slope=12.0,
    c1_low=+15.0, c1_high=-15.0,
    c2_low=-15.0, c2_high=+15.0,
    up_delay_ms=250.0,
    n_fft=4096,
    hop=2048,
    state_csv_path=None,
    output_gain_db=0.0,
):
    """Gate-controlled C1/C2 tilt filter on `in_path` -> `out_path` (B200 path)."""
    from . import engine

    print("=" * 70)
    print("Tomatis audio processor (B200)")
    print("=" * 70)
    print(f"\ninput : {in_path}\noutput: {out_path}\n")
    if gate_mode == "log_percent":
        T = tb.gate_threshold_log_percent(gate_ui, dynamic_range)
        mode_str = f"log_percent (dynamic range {dynamic_range} dB)"
    else:
        T = tb.gate_threshold_linear(gate_ui, gate_scale, gate_offset)
        mode_str = f"linear (scale={gate_scale}, offset={gate_offset})"
    Ton, Toff = tb.hysteresis_pair(T, hysteresis_db)
    print(f"gate: ui={gate_ui} mode={mode_str} T={T:.1f} dBFS  up(C1->C2) {Ton:.1f}  down(C2->C1) {Toff:.1f}  "
          f"hysteresis {hysteresis_db} dB  up-delay {up_delay_ms} ms")
    print(f"tilt: fc={fc} Hz slope={slope} dB/oct  C1 {c1_low:+.1f}/{c1_high:+.1f} dB  C2 {c2_low:+.1f}/{c2_high:+.1f} dB")
    print(f"stft: n_fft={n_fft} hop={hop}\n")

    x, sr = _open_check(in_path, "standard")
    total = len(x)
    pad_end = (hop - ((total - n_fft) % hop)) % hop
    print(f"boundary padding: start {n_fft // 2}, end {pad_end} samples")

    res = engine.run_streaming(
        "standard", [x], sr, device=DEVICE, gate_ui=gate_ui, gate_mode=gate_mode, dynamic_range=dynamic_range,
        gate_scale=gate_scale, gate_offset=gate_offset, hysteresis_db=hysteresis_db, fc=fc, slope=slope,
        c1_low=c1_low, c1_high=c1_high, c2_low=c2_low, c2_high=c2_high, up_delay_ms=up_delay_ms,
        n_fft=n_fft, hop=hop, output_gain_db=output_gain_db)[0]

    written, is_flac = _write_output(out_path, res["out"], sr)
    if state_csv_path:
        report.write_state_csv(state_csv_path, "standard", res)
        print(f"[OK] state CSV: {state_csv_path}")

    st = report.gate_statistics(res["states"], total, sr)
    n = st["frames"]
    print("\n" + "=" * 70 + "\ndone\n" + "=" * 70)
    print(f"frames: {n}")
    print(f"  C1: {st['c1_frames']} ({st['c1_frames'] / n * 100:.1f}%)")      # ZeroDivisionError on empty input,
    print(f"  C2: {st['c2_frames']} ({st['c2_frames'] / n * 100:.1f}%)")      # like src/process_tomatis.py:463
    limited = int((res["chunk_peaks"] > np.float32(tb.PEAK_LIMIT)).sum())
    print(f"limiter chunks: {len(res['chunk_lengths'])} ({limited} scaled to {tb.PEAK_LIMIT})")
    print(f"output: {written}  ({total} samples, same as the input)")
    if not is_flac:
        print(f'convert with: ffmpeg -y -i "{written}" -c:a flac -compression_level 8 "{out_path}"')
    print()


def build_parser():
    ap = argparse.ArgumentParser(description="Tomatis audio processor - gate-controlled C1/C2 tilt filter (B200)",
                                 formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("-i", "--input", required=True, help="input FLAC/WAV file")
    ap.add_argument("-o", "--output", required=True, help="output FLAC file")
    ap.add_argument("--gate_ui", type=float, default=50, help="gate UI value (0-100)")
    ap.add_argument("--gate_mode", choices=["linear", "log_percent"], default="log_percent", help="gate map")
    ap.add_argument("--dynamic_range", type=float, default=80.0, help="dynamic range (dB), log_percent map")
    ap.add_argument("--gate_scale", type=float, default=1.0, help="gate scale (linear map)")
    ap.add_argument("--gate_offset", type=float, default=-100, help="gate offset (linear map)")
    ap.add_argument("--hyst_db", type=float, default=3.0, help="hysteresis (dB)")
    ap.add_argument("--up_delay_ms", type=float, default=250.0, help="C1->C2 up-delay (ms)")
    ap.add_argument("--fc", type=float, default=1000.0, help="pivot frequency (Hz)")
    ap.add_argument("--slope", type=float, default=12.0, help="slope (dB/octave)")
    ap.add_argument("--c1_low", type=float, default=15.0, help="C1 low-frequency gain (dB)")
    ap.add_argument("--c1_high", type=float, default=-15.0, help="C1 high-frequency gain (dB)")
    ap.add_argument("--c2_low", type=float, default=-15.0, help="C2 low-frequency gain (dB)")
    ap.add_argument("--c2_high", type=float, default=15.0, help="C2 high-frequency gain (dB)")
    ap.add_argument("--n_fft", type=int, default=4096, help="FFT length")
    ap.add_argument("--hop", type=int, default=2048, help="hop length")
    ap.add_argument("--state_csv", default=None, help="per-frame state CSV path")
    ap.add_argument("--output_gain_db", type=float, default=0.0, help="output gain (dB)")
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
        process(args.input, args.output, gate_ui=args.gate_ui, gate_mode=args.gate_mode,
                dynamic_range=args.dynamic_range, gate_scale=args.gate_scale, gate_offset=args.gate_offset,
                hysteresis_db=args.hyst_db, fc=args.fc, slope=args.slope, c1_low=args.c1_low, c1_high=args.c1_high,
                c2_low=args.c2_low, c2_high=args.c2_high, up_delay_ms=args.up_delay_ms, n_fft=args.n_fft,
                hop=args.hop, state_csv_path=args.state_csv, output_gain_db=args.output_gain_db)
    except Exception as e:
        print(f"\n[ERR] {e}")
        import traceback
        traceback.print_exc()
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
