#!/usr/bin/env python3
"""Calibration front end -- drop-in for the reference's `src/calibrate_to_baseline_v2.py`: from an
original file and a recording of the hardware's output ("baseline") derive the gate parameters (`gate_offset`, `hyst_db`,
`up_delay_ms`) the processors take.  Same CLI flags (:130-158), same printed figures, same JSON (:289-308).

On the device (csrc/calib.cuh, tmt_calib_* in include/tomatis_b200.h):
  * envelope + polyphase decimation of both files and their valid cross-correlation (find_delay_by_corr, :44-86);
  * frame levels of both files and the band energies behind the tilt of every baseline frame (:186-196);
  * every gate simulation of the grid search (:237-262), one thread per (delay, hysteresis, threshold) combination.
On the host: the small per-frame bookkeeping in between (median filter, two-means clustering, debouncing, medians, the
pick of the first best score).  There is no CPU path for the device parts.  Supported: n_fft / hop = 4096 / 2048.
"""
from __future__ import annotations

import argparse
import json
import sys

import numpy as np

from . import audio_io
from . import tables as tb

EPS = 1e-12
DEVICE = 0


def _check_fft(n_fft, hop):
    if n_fft != tb.N_FFT or hop != tb.HOP:
        raise NotImplementedError(f"GPU path implements n_fft={tb.N_FFT}, hop={tb.HOP}; got {n_fft}/{hop}")


def _stereo32(x):
    x = np.asarray(x, dtype=np.float32)
    assert x.ndim == 2 and x.shape[1] == 2                       # the reference asserts two channels (:47,165)
    return np.ascontiguousarray(x)


# ------------------------------------------------------------------------------------------------ delay estimate
def find_delay(orig, base, sr=48000, ds_sr=2000, chunk_sec=25) -> dict:
    """find_delay_by_corr (:44-86) on arrays (host float32 [N, 2] or device tensors): the envelope of the middle `chunk_sec`
    seconds of the baseline is slid over the envelope of the whole original, both at `ds_sr`.
    Returns dict(delay, k, corr, s, e, n_base_ds)."""
    from . import engine
    orig_d, base_d = engine.to_device([orig, base], DEVICE)
    n_base = int(base_d.shape[0])
    mid, half = int(0.5 * n_base), int(0.5 * chunk_sec * sr)
    s, e = max(0, mid - half), min(n_base, mid + half)
    mb_ds = engine.calib_envelope(base_d, s, e, ds_sr, sr, DEVICE)
    mo_ds = engine.calib_envelope(orig_d, 0, int(orig_d.shape[0]), ds_sr, sr, DEVICE)
    corr = engine.calib_xcorr(mo_ds, mb_ds)                     # raises ValueError if the chunk is longer than the original
    k = int(np.argmax(corr))
    n_b = int(mb_ds.numel())
    base_center = (s + (e - s) // 2) / sr
    orig_center = (k + n_b // 2) / ds_sr
    return dict(delay=int(round((orig_center - base_center) * sr)), k=k, corr=corr, s=s, e=e, n_base_ds=n_b,
                mo_ds=mo_ds, mb_ds=mb_ds)


def find_delay_by_corr(orig_path, base_path, sr=48000, ds_sr=2000, chunk_sec=25) -> int:
    """Reference signature (:44): delay of the original against the baseline in samples."""
    xo, sr_o = audio_io.read(orig_path, dtype="float32")
    xb, sr_b = audio_io.read(base_path, dtype="float32")
    assert sr_o == sr and sr_b == sr
    return find_delay(_stereo32(xo), _stereo32(xb), sr, ds_sr, chunk_sec)["delay"]


# ------------------------------------------------------------------------------------------------ per-frame features
def band_bins(sr, n_fft, band):
    """[first, last + 1) of the rfft bins with band[0] <= f < band[1] (:24-26)."""
    freqs = np.fft.rfftfreq(n_fft, 1 / sr)
    idx = np.flatnonzero((freqs >= band[0]) & (freqs < band[1]))
    return (int(idx[0]), int(idx[-1]) + 1) if idx.size else (0, 0)


def tilt_from_energies(e_lo, e_hi) -> np.ndarray:
    """10*log10(Ehi / Elo + EPS) with the reference's precisions (:28-30): the epsilon joins the float32 band sum, the
    ratio and the logarithm are float64; stored as float32 (:196)."""
    lo = (np.asarray(e_lo, np.float32) + np.float32(EPS)).astype(np.float64)
    hi = (np.asarray(e_hi, np.float32) + np.float32(EPS)).astype(np.float64)
    return (10 * np.log10(hi / lo + EPS)).astype(np.float32)


def frame_features(xo, xb, sr, n_fft=4096, hop=2048, lo=(200, 1000), hi=(2000, 8000)):
    """Levels of both files and the baseline's band tilt for the frames [i*hop, i*hop + n_fft) (:179-196).
    Returns (frame_starts int64, orig_level, base_level, tilts), the last three float32."""
    from . import engine
    _check_fft(n_fft, hop)
    xo_d, xb_d = engine.to_device([xo, xb], DEVICE)
    avail = int(min(xo_d.shape[0], xb_d.shape[0]))
    n_frames = 1 + (avail - n_fft) // hop if avail >= n_fft else 0
    starts = (np.arange(n_frames) * hop).astype(np.int64)
    orig_level = engine.calib_frame_levels(xo_d[:avail], DEVICE)[:n_frames]
    base_level = engine.calib_frame_levels(xb_d[:avail], DEVICE)[:n_frames]
    e_lo, e_hi = engine.calib_band_energies(xb_d[:avail], n_frames, band_bins(sr, n_fft, lo), band_bins(sr, n_fft, hi), DEVICE)
    return starts, orig_level, base_level, tilt_from_energies(e_lo, e_hi)


# ------------------------------------------------------------------------------------------------ host bookkeeping
def kmeans2_1d(x, iters=25):
    """Two-means on a line started from the 30th / 70th percentiles, ties to the first centre (:32-42).
    Returns (labels int32: 1 = nearer the second centre, m1, m2)."""
    x = np.asarray(x)
    m1, m2 = (float(v) for v in np.percentile(x, [30, 70]))
    for _ in range(iters):
        first = np.abs(x - m1) <= np.abs(x - m2)
        n_first = int(first.sum())
        if n_first:
            m1 = float(np.mean(x[first]))
        if n_first < x.size:
            m2 = float(np.mean(x[~first]))
    return (np.abs(x - m2) < np.abs(x - m1)).astype(np.int32), m1, m2


def debounce_state(state, min_run=3):
    """Runs shorter than min_run are overwritten with the value on their left, scanning left to right so that a rewritten
    run merges with what precedes it (:114-131); a short run at the very start takes the value on its right."""
    s = np.array(state, copy=True)
    n = len(s)
    i = 0
    while i < n:
        j = i + 1
        while j < n and s[j] == s[i]:
            j += 1
        if j - i < min_run:
            if i > 0:
                s[i:j] = s[i - 1]
            elif j < n:
                s[i:j] = s[j]
        i = j
    return s


def baseline_states(tilts, music_mask, tilt_medfilt=5):
    """Reference states read off the baseline (:205-225): median-filtered tilt, two clusters over the music frames, the
    cluster with the higher mean tilt is C2, debounced.  Returns (states int32, smoothed tilts float32)."""
    k = int(tilt_medfilt)
    k += (k % 2 == 0)
    k = max(k, 3)
    ts = tb.medfilt_zero_padded(np.asarray(tilts, np.float32), k).astype(np.float32)
    sel = ts[music_mask]
    lab, _, _ = kmeans2_1d(sel)
    hi_is_1 = True
    if np.any(lab == 1) or np.any(lab == 0):
        mean1 = float(np.mean(sel[lab == 1])) if np.any(lab == 1) else -1e9
        mean0 = float(np.mean(sel[lab == 0])) if np.any(lab == 0) else -1e9
        hi_is_1 = not (mean0 > mean1)
    state = np.ones(len(ts), np.int32)
    state[music_mask] = np.where(lab == (1 if hi_is_1 else 0), 2, 1).astype(np.int32)
    return debounce_state(state, min_run=3), ts


def simulate_state(level_dbfs, frame_starts, sr, T, hyst, up_delay_ms):
    """Reference signature (:88-112): the up-delay gate over frames at arbitrary positions -> int32 states (1 / 2)."""
    from . import engine
    level = np.asarray(level_dbfs, dtype=np.float32)
    on, off = np.float32(T + hyst / 2), np.float32(T - hyst / 2)
    delay = int(round(sr * up_delay_ms / 1000.0))
    _, _, st = engine.calib_gate_grid(level, frame_starts, np.ones(level.size, np.uint8), [on], [off], [delay],
                                      want_states=True, device=DEVICE)
    return st[0].astype(np.int32)


def grid_search(orig_level, base_level, base_state, frame_starts, music_mask, sr, hyst_list, delay_list_ms,
                gain_search_pm_db=3.0, gain_step_db=0.5, T_pm_db=10.0, T_step_db=0.25, want_table=False):
    """The fit (:227-270): for every trial gain, every (up-delay, hysteresis, threshold) combination is simulated on the
    music frames in one launch; the best is the first combination, in the reference's loop order, with the strictly
    smallest mismatch + 1e-5 * switches.  Returns (best dict or None, gain_db0, table of (gain, up_ms, hyst, T,
    mismatches, switches) if asked)."""
    from . import engine
    gain_db0 = float(np.median((base_level - orig_level)[music_mask]))
    gains = np.arange(gain_db0 - gain_search_pm_db, gain_db0 + gain_search_pm_db + 1e-9, gain_step_db).astype(np.float32)
    idx = np.flatnonzero(music_mask)
    fs_fit = np.asarray(frame_starts)[idx]
    s_fit = np.asarray(base_state)[idx]
    want = s_fit.astype(np.uint8)
    n = int(idx.size)
    delays_ms = np.asarray([float(v) for v in delay_list_ms], dtype=np.float64)
    hysts = np.asarray([float(v) for v in hyst_list], dtype=np.float64)
    delay_samples = np.array([int(round(sr * v / 1000.0)) for v in delays_ms], dtype=np.int64)
    best, table = None, []
    for gain_db in gains:
        levels_adj = (orig_level + gain_db)[idx]                                     # float32
        c1, c2 = levels_adj[s_fit == 1], levels_adj[s_fit == 2]
        if len(c1) < 10 or len(c2) < 10:
            continue
        T0 = 0.5 * (float(np.median(c1)) + float(np.median(c2)))
        Ts = np.arange(T0 - T_pm_db, T0 + T_pm_db + 1e-9, T_step_db).astype(np.float32).astype(np.float64)
        if Ts.size == 0 or delays_ms.size == 0 or hysts.size == 0:
            continue
        # combinations in the reference's nesting: delay, then hysteresis, then threshold (fastest)
        d_i, h_i, t_i = np.meshgrid(np.arange(delays_ms.size), np.arange(hysts.size), np.arange(Ts.size), indexing="ij")
        d_i, h_i, t_i = d_i.ravel(), h_i.ravel(), t_i.ravel()
        on = (Ts[t_i] + hysts[h_i] / 2).astype(np.float32)                           # compared in float32 with float32 levels
        off = (Ts[t_i] - hysts[h_i] / 2).astype(np.float32)
        mis, sw = engine.calib_gate_grid(levels_adj, fs_fit, want, on, off, delay_samples[d_i], device=DEVICE)
        score = mis.astype(np.float64) / n + 1e-5 * sw.astype(np.float64)
        if want_table:
            table.extend(zip([float(gain_db)] * score.size, delays_ms[d_i].tolist(), hysts[h_i].tolist(), Ts[t_i].tolist(),
                             mis.tolist(), sw.tolist()))
        c = int(np.argmin(score))                                                    # first of equal scores
        if best is None or score[c] < best["score"]:
            best = dict(score=float(score[c]), mismatch=float(mis[c] / n), switches=int(sw[c]), T=float(Ts[t_i[c]]),
                        hyst=float(hysts[h_i[c]]), up_ms=float(delays_ms[d_i[c]]), gain_db=float(gain_db), T0=float(T0))
    return best, gain_db0, table


# ------------------------------------------------------------------------------------------------ the whole run
def calibrate(orig, base, sr=48000, gate_ui=50.0, gate_scale=1.0, n_fft=4096, hop=2048, max_minutes=6.0,
              hyst_list=(0, 1, 2, 3, 4, 6), delay_list_ms=(0, 50, 100, 150, 200, 250), tilt_lo=(200, 1000),
              tilt_hi=(2000, 8000), tilt_medfilt=5, music_dbfs=-65.0, gain_search_pm_db=3.0, gain_step_db=0.5,
              T_pm_db=10.0, T_step_db=0.25, log=None) -> dict:
    """main() of the reference (:160-308) on arrays.  `json` is the dictionary the CLI saves (without the two paths)."""
    from . import engine
    _check_fft(n_fft, hop)
    say = log if log is not None else (lambda *_: None)
    orig_d, base_d = engine.to_device([_stereo32(orig) if isinstance(orig, np.ndarray) else orig,
                                       _stereo32(base) if isinstance(base, np.ndarray) else base], DEVICE)
    delay = find_delay(orig_d, base_d, sr=sr)["delay"]
    say(f"[ALIGN] estimated delay (orig - base): {delay} samples ({delay / sr * 1000:.2f} ms)")
    base_start, orig_start = max(0, -delay), max(0, delay)
    avail = min(int(base_d.shape[0]) - base_start, int(orig_d.shape[0]) - orig_start, int(max_minutes * 60 * sr))
    if avail <= n_fft:
        raise ValueError("the overlap of the two files is too short to calibrate")
    xb, xo = base_d[base_start:base_start + avail], orig_d[orig_start:orig_start + avail]
    starts, orig_level, base_level, tilts = frame_features(xo, xb, sr, n_fft, hop, tuple(tilt_lo), tuple(tilt_hi))

    music_mask = base_level > music_dbfs
    music_ratio = float(np.mean(music_mask))
    say(f"[MASK] music frames ratio: {music_ratio * 100:.1f}% (threshold {music_dbfs} dBFS)")
    if music_ratio < 0.2:
        say("[WARN] few music frames: consider a lower --music_dbfs (for example -70)")
    base_state, tilts_s = baseline_states(tilts, music_mask, tilt_medfilt)
    best, gain_db0, _ = grid_search(orig_level, base_level, base_state, starts, music_mask, sr, hyst_list, delay_list_ms,
                                    gain_search_pm_db, gain_step_db, T_pm_db, T_step_db)
    say(f"[GAIN] initial gain_db0 (base - orig): {gain_db0:.2f} dB")
    if best is None:
        raise RuntimeError("no usable optimum: relax --music_dbfs or raise --max_minutes")
    # T was fitted on levels shifted by the gain; the processors see the raw file (:272-278)
    T_raw = best["T"] - best["gain_db"]
    out = {
        "delay_samples_orig_minus_base": int(delay),
        "music_dbfs": float(music_dbfs),
        "gain_db_base_minus_orig": float(best["gain_db"]),
        "T_adj_dbfs": float(best["T"]),
        "T_raw_dbfs": float(T_raw),
        "gate_ui": float(gate_ui),
        "gate_scale": float(gate_scale),
        "gate_offset": float(T_raw - gate_scale * gate_ui),
        "hyst_db": float(best["hyst"]),
        "up_delay_ms": float(best["up_ms"]),
        "mismatch": float(best["mismatch"]),
        "switches": int(best["switches"]),
    }
    return dict(json=out, best=best, gain_db0=gain_db0, orig_level=orig_level, base_level=base_level, tilts=tilts,
                tilts_s=tilts_s, base_state=base_state, music_mask=music_mask, starts=starts, delay=delay)


def build_parser():
    ap = argparse.ArgumentParser(description="derive gate parameters from an original / baseline recording pair (B200)")
    ap.add_argument("--orig", required=True)
    ap.add_argument("--base", required=True)
    ap.add_argument("--gate_ui", type=float, default=50.0)
    ap.add_argument("--gate_scale", type=float, default=1.0)
    ap.add_argument("--n_fft", type=int, default=4096)
    ap.add_argument("--hop", type=int, default=2048)
    ap.add_argument("--sr", type=int, default=48000)
    ap.add_argument("--max_minutes", type=float, default=6.0)
    ap.add_argument("--hyst_list", type=float, nargs="+", default=[0, 1, 2, 3, 4, 6])
    ap.add_argument("--delay_list_ms", type=float, nargs="+", default=[0, 50, 100, 150, 200, 250])
    ap.add_argument("--tilt_lo", type=int, nargs=2, default=[200, 1000])
    ap.add_argument("--tilt_hi", type=int, nargs=2, default=[2000, 8000])
    ap.add_argument("--tilt_medfilt", type=int, default=5, help="median filter length for the tilt (odd): 3/5/7")
    ap.add_argument("--music_dbfs", type=float, default=-65.0, help="fit only frames whose baseline level is above this")
    ap.add_argument("--gain_search_pm_db", type=float, default=3.0, help="search range around the initial gain, +-dB")
    ap.add_argument("--gain_step_db", type=float, default=0.5)
    ap.add_argument("--T_pm_db", type=float, default=10.0, help="search range around T0, +-dB")
    ap.add_argument("--T_step_db", type=float, default=0.25)
    ap.add_argument("--out_json", default="calibration_v2.json")
    ap.add_argument("--device", type=int, default=0, help="extension: CUDA device index")
    return ap


def main(argv=None):
    global DEVICE
    args = build_parser().parse_args(argv)
    DEVICE = args.device
    sr = args.sr
    xo, sr_o = audio_io.read(args.orig, dtype="float32")
    xb, sr_b = audio_io.read(args.base, dtype="float32")
    assert sr_o == sr and sr_b == sr
    assert xo.shape[1] == 2 and xb.shape[1] == 2
    r = calibrate(xo, xb, sr=sr, gate_ui=args.gate_ui, gate_scale=args.gate_scale, n_fft=args.n_fft, hop=args.hop,
                  max_minutes=args.max_minutes, hyst_list=args.hyst_list, delay_list_ms=args.delay_list_ms,
                  tilt_lo=args.tilt_lo, tilt_hi=args.tilt_hi, tilt_medfilt=args.tilt_medfilt, music_dbfs=args.music_dbfs,
                  gain_search_pm_db=args.gain_search_pm_db, gain_step_db=args.gain_step_db, T_pm_db=args.T_pm_db,
                  T_step_db=args.T_step_db, log=print)
    best, out = r["best"], r["json"]
    print("\n[BEST]")
    print(best)
    print(f"\n[RECOMMEND] gain_db (diagnostic only): {out['gain_db_base_minus_orig']:+.2f} dB (base - orig)")
    print(f"[RECOMMEND] T_adj (on leveled orig): {out['T_adj_dbfs']:.2f} dBFS")
    print(f"[RECOMMEND] T_raw (for process_tomatis): {out['T_raw_dbfs']:.2f} dBFS")
    print(f"[RECOMMEND] gate_ui={args.gate_ui:.1f}, gate_scale={args.gate_scale:.2f}, gate_offset={out['gate_offset']:.2f}")
    print(f"[RECOMMEND] hyst_db={best['hyst']:.1f}, up_delay_ms={best['up_ms']:.0f}")
    print(f"[RECOMMEND] mismatch={best['mismatch'] * 100:.2f}%, switches={best['switches']} (on music frames)")
    saved = {"orig": args.orig, "base": args.base}
    saved.update(out)
    with open(args.out_json, "w", encoding="utf-8") as f:
        json.dump(saved, f, ensure_ascii=False, indent=2)
    print(f"\n[SAVED] {args.out_json}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
