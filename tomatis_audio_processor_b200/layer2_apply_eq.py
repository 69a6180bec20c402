#!/usr/bin/env python3
"""Static STFT equaliser -- drop-in for the reference's `src/layer2_apply_eq.py`, the stage
that follows the Tomatis processor in docs/Workflow_v2.md.

Same `apply_eq_stft(in_path, out_path, eq_csv, n_fft, hop, pad, global_gain_db, auto_gain_protect, peak_target)` signature
(src/layer2_apply_eq.py:66-76), same CLI flags (:239-263), same 48 kHz / stereo ValueError (:81-84), same outputs: the
equalised file (FLAC PCM_24, else WAV), and when its peak exceeds `peak_target` a second, gain-protected file
`<out>_gp.flac` made from the first one (:220-234).  The EQ curve handling (CSV columns, log-frequency interpolation onto
the rfft bins) is host code with the reference's NumPy expressions; the filtering runs in the fused STFT/OLA kernel with a
one-row gain table.
"""
from __future__ import annotations

import argparse
import csv
import sys

import numpy as np

from . import audio_io

EPS = 1e-12
DEVICE = 0


def db_to_lin(db):
    return 10.0 ** (db / 20.0)


def load_eq_csv(eq_csv_path):
    """(freqs, dbs) float32, sorted by frequency.  Accepts the reference's column aliases (src/layer2_apply_eq.py:11-46):
    frequency = freq_hz | freq | hz | f; gain = delta_db_smooth | delta_db | db | gain_db | delta | gain."""
    with open(eq_csv_path, "r", encoding="utf-8") as f:
        reader = csv.DictReader(f)
        cols = [c.lower().strip() for c in reader.fieldnames]
        pick = lambda cands: next((c for c in cands if c in cols), None)
        f_col = pick(["freq_hz", "freq", "hz", "f"])
        d_col = pick(["delta_db_smooth", "delta_db", "db", "gain_db", "delta", "gain"])
        if f_col is None or d_col is None:
            raise ValueError(f"unexpected EQ CSV columns: {reader.fieldnames}")
        print(f"[EQ_LOAD] Using columns: freq='{f_col}', gain='{d_col}'")
        rows = [(float(r[f_col]), float(r[d_col])) for r in reader]
    freqs = np.array([r[0] for r in rows], np.float32)
    dbs = np.array([r[1] for r in rows], np.float32)
    idx = np.argsort(freqs)
    return freqs[idx], dbs[idx]


def build_gain_per_bin(sr, n_fft, eq_freqs, eq_db):
    """dB curve -> linear gain per rfft bin, interpolated on a log-frequency axis, clamped to the end values
    (src/layer2_apply_eq.py:48-64)."""
    f_bins = np.fft.rfftfreq(n_fft, 1.0 / sr).astype(np.float32)
    x = np.log10(np.maximum(eq_freqs, 1.0))
    xb = np.log10(np.maximum(f_bins, 1.0))
    yb = np.interp(xb, x, eq_db, left=eq_db[0], right=eq_db[-1]).astype(np.float32)
    return db_to_lin(yb).astype(np.float32)


def _write(path, y, sr):
    try:
        audio_io.write(path, y, sr, subtype="PCM_24", format="FLAC")
        return path, True
    except Exception as e:
        wav = path.replace(".flac", ".wav")
        print(f"[WARN] FLAC write failed, writing WAV: {e}")
        audio_io.write(wav, y, sr, subtype="PCM_24", format="WAV")
        return wav, False


def apply_eq_stft(
    in_path,
    out_path,
    eq_csv,
    n_fft=4096,
    hop=2048,
    pad=True,
    global_gain_db=0.0,
    auto_gain_protect=True,
    peak_target=0.99,
):
    from . import engine

    x, sr = audio_io.read(in_path, dtype="float32")
    if sr != 48000:
        raise ValueError(f"expected 48 kHz, got {sr}")
    if x.shape[1] != 2:
        raise ValueError(f"expected stereo, got {x.shape[1]} channel(s)")
    eq_freqs, eq_db = load_eq_csv(eq_csv)
    gain_bins = build_gain_per_bin(sr, n_fft, eq_freqs, eq_db)
    r = engine.run_eq([x], sr, gain_bins, device=DEVICE, pad=pad, global_gain_db=global_gain_db,
                      auto_gain_protect=auto_gain_protect, peak_target=peak_target, n_fft=n_fft, hop=hop)[0]
    written, is_flac = _write(out_path, r["out"], sr)
    if r["out_gp"] is not None:
        print(f"[GAIN_PROTECT] peak={r['peak_seen']:.4f} > {peak_target}, apply scale={r['scale']:.4f}")
        gp, _ = _write(out_path.replace(".flac", "_gp.flac"), r["out_gp"], sr)
        print(f"[DONE] gain-protected file: {gp}")
    print("[DONE] EQ applied.")
    if not is_flac:
        print(f"[NOTE] output is WAV: {written}; convert with ffmpeg if FLAC is needed.")


def build_parser():
    ap = argparse.ArgumentParser()
    ap.add_argument("-i", "--input", required=True)
    ap.add_argument("-o", "--output", required=True)
    ap.add_argument("--eq_csv", required=True)
    ap.add_argument("--n_fft", type=int, default=4096)
    ap.add_argument("--hop", type=int, default=2048)
    ap.add_argument("--no_pad", action="store_true")
    ap.add_argument("--gain_db", type=float, default=0.0, help="extra overall gain (dB)")
    ap.add_argument("--no_gain_protect", action="store_true")
    ap.add_argument("--device", type=int, default=0, help="extension: CUDA device index")
    return ap


def main(argv=None):
    global DEVICE
    a = build_parser().parse_args(argv)
    DEVICE = a.device
    apply_eq_stft(a.input, a.output, a.eq_csv, n_fft=a.n_fft, hop=a.hop, pad=(not a.no_pad), global_gain_db=a.gain_db,
                  auto_gain_protect=(not a.no_gain_protect))


if __name__ == "__main__":
    sys.exit(main())
