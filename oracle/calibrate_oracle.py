"""TEST INFRASTRUCTURE ONLY -- NumPy/SciPy restatement of the reference's calibration front end
(`src/calibrate_to_baseline_v2.py`) on in-memory arrays.

Only `tests/` may import this file.  Pinned to the executed reference by
`tests/test_oracle_vs_reference.py::test_calibration_restatement` (live, build container only) and by the frozen
fixtures `tests/golden/cal_*.npz` (`oracle/make_golden.py`).  The reference takes `resample_poly`, `fftconvolve` and
`medfilt` from scipy.signal (installed here: 1.18.1; not pinned by the reference); the oracle calls the same functions.
"""
from __future__ import annotations

import numpy as np
from scipy.signal import fftconvolve, medfilt, resample_poly

EPS = 1e-12


def power_mono(x_lr):
    """src/calibrate_to_baseline_v2.py:8-11"""
    p = 0.5 * (x_lr[:, 0] * x_lr[:, 0] + x_lr[:, 1] * x_lr[:, 1])
    return np.sqrt(p + EPS)


def rms_dbfs_from_mono(mono) -> float:
    """:13-15"""
    r = np.sqrt(np.mean(mono * mono) + EPS)
    return float(20 * np.log10(r + EPS))


def stft_band_tilt(frame_lr, sr, n_fft, lo=(200, 1000), hi=(2000, 8000)) -> float:
    """:17-30 -- 10*log10 of the high-band over the low-band energy of the windowed power-mono frame."""
    win = np.hanning(n_fft).astype(np.float32)
    X = np.fft.rfft(power_mono(frame_lr) * win)
    P = (X.real * X.real + X.imag * X.imag).astype(np.float32)
    freqs = np.fft.rfftfreq(n_fft, 1 / sr)
    e_lo = float(np.sum(P[(freqs >= lo[0]) & (freqs < lo[1])]) + EPS)
    e_hi = float(np.sum(P[(freqs >= hi[0]) & (freqs < hi[1])]) + EPS)
    return float(10 * np.log10(e_hi / e_lo + EPS))


def kmeans2_1d(x, iters=25):
    """:32-42"""
    m1, m2 = np.percentile(x, [30, 70]).astype(float)
    for _ in range(iters):
        near1 = np.abs(x - m1) <= np.abs(x - m2)
        if near1.any():
            m1 = float(np.mean(x[near1]))
        if (~near1).any():
            m2 = float(np.mean(x[~near1]))
    return (np.abs(x - m2) < np.abs(x - m1)).astype(np.int32), m1, m2


def find_delay(orig, base, sr=48000, ds_sr=2000, chunk_sec=25) -> dict:
    """:44-86 on arrays: envelope of the middle `chunk_sec` of the baseline against the envelope of the whole original,
    both decimated to ds_sr, peak of the cross-correlation."""
    n_base = len(base)
    mid, half = int(0.5 * n_base), int(0.5 * chunk_sec * sr)
    s, e = max(0, mid - half), min(n_base, mid + half)
    mb_ds = resample_poly(power_mono(base[s:e]), ds_sr, sr).astype(np.float32)
    mb_ds = mb_ds - np.mean(mb_ds)
    mo = power_mono(orig).astype(np.float32)
    mo_ds = resample_poly(mo, ds_sr, sr).astype(np.float32)
    mo_ds = mo_ds - np.mean(mo_ds)
    corr = fftconvolve(mo_ds, mb_ds[::-1], mode="valid")
    k = int(np.argmax(corr))
    base_center = (s + (e - s) // 2) / sr
    orig_center = (k + len(mb_ds) // 2) / ds_sr
    return dict(delay=int(round((orig_center - base_center) * sr)), k=k, corr=corr, mo_ds=mo_ds, mb_ds=mb_ds, s=s, e=e)


def simulate_state(level_dbfs, frame_starts, sr, T, hyst, up_delay_ms):
    """:88-112 -- the up-delay automaton on an irregular list of frame positions."""
    t_on, t_off = T + hyst / 2, T - hyst / 2
    delay = int(round(sr * up_delay_ms / 1000.0))
    state, pending = 1, None
    out = np.zeros_like(level_dbfs, dtype=np.int32)
    for i, (lv, st) in enumerate(zip(level_dbfs, frame_starts)):
        if state == 1:
            if lv >= t_on:
                if pending is None:
                    pending = st + delay
            else:
                pending = None
            if pending is not None and st >= pending:
                state, pending = 2, None
        elif lv <= t_off:
            state, pending = 1, None
        out[i] = state
    return out


def debounce_state(state, min_run=3):
    """:114-131 -- runs shorter than min_run take the state to their left."""
    s = state.copy()
    n, i = len(s), 0
    while i < n:
        j = i + 1
        while j < n and s[j] == s[i]:
            j += 1
        if j - i < min_run:
            s[i:j] = s[i - 1] if i > 0 else s[j] if j < n else s[i]
        i = j
    return s


def frame_features(xo, xb, sr, n_fft, hop, lo, hi):
    """:179-196"""
    n_frames = 1 + (len(xo) - n_fft) // hop
    starts = (np.arange(n_frames) * hop).astype(np.int64)
    orig_level, base_level, tilts = (np.zeros(n_frames, np.float32) for _ in range(3))
    for i, st in enumerate(starts):
        orig_level[i] = rms_dbfs_from_mono(power_mono(xo[st:st + n_fft, :]))
        base_level[i] = rms_dbfs_from_mono(power_mono(xb[st:st + n_fft, :]))
        tilts[i] = stft_band_tilt(xb[st:st + n_fft, :], sr, n_fft, lo=lo, hi=hi)
    return starts, orig_level, base_level, tilts


def baseline_states(tilts, music_mask, tilt_medfilt=5):
    """:205-225 -- median-filtered tilt, two clusters on the music frames, higher tilt = C2, debounced."""
    k = int(tilt_medfilt)
    if k % 2 == 0:
        k += 1
    k = max(k, 3)
    ts = medfilt(tilts, kernel_size=k).astype(np.float32)
    lab, _, _ = kmeans2_1d(ts[music_mask])
    state = np.ones(len(tilts), np.int32)
    state[music_mask] = np.where(lab == 1, 2, 1).astype(np.int32)
    mean1 = float(np.mean(ts[music_mask][lab == 1])) if np.any(lab == 1) else -1e9
    mean0 = float(np.mean(ts[music_mask][lab == 0])) if np.any(lab == 0) else -1e9
    if mean0 > mean1:
        state[music_mask] = np.where(lab == 0, 2, 1).astype(np.int32)
    return debounce_state(state, min_run=3), ts


def grid_search(orig_level, base_level, base_state, starts, music_mask, sr, hyst_list, delay_list_ms,
                gain_search_pm_db=3.0, gain_step_db=0.5, T_pm_db=10.0, T_step_db=0.25, want_table=False):
    """:227-270 -- first strict minimum of mismatch + 1e-5 * switches over gain x delay x hysteresis x threshold."""
    gain_db0 = float(np.median((base_level - orig_level)[music_mask]))
    gains = np.arange(gain_db0 - gain_search_pm_db, gain_db0 + gain_search_pm_db + 1e-9, gain_step_db).astype(np.float32)
    idx = np.flatnonzero(music_mask)
    fs_fit = starts[idx]
    s_fit = base_state[idx]
    best, table = None, []
    for gain_db in gains:
        levels_adj = (orig_level + gain_db)[idx]
        c1, c2 = levels_adj[s_fit == 1], levels_adj[s_fit == 2]
        if len(c1) < 10 or len(c2) < 10:
            continue
        T0 = 0.5 * (float(np.median(c1)) + float(np.median(c2)))
        Ts = np.arange(T0 - T_pm_db, T0 + T_pm_db + 1e-9, T_step_db).astype(np.float32)
        for up_ms in delay_list_ms:
            for hyst in hyst_list:
                for T in Ts:
                    pred = simulate_state(levels_adj, fs_fit, sr, float(T), float(hyst), float(up_ms))
                    mismatch = float(np.mean(pred != s_fit))
                    switches = int(np.sum(pred[1:] != pred[:-1]))
                    score = mismatch + 1e-5 * switches
                    if want_table:
                        table.append((float(gain_db), float(up_ms), float(hyst), float(T), int((pred != s_fit).sum()), switches))
                    if best is None or score < best["score"]:
                        best = dict(score=score, mismatch=mismatch, switches=switches, T=float(T), hyst=float(hyst),
                                    up_ms=float(up_ms), gain_db=float(gain_db), T0=float(T0))
    return best, gain_db0, table


def calibrate(orig, base, sr=48000, gate_ui=50.0, gate_scale=1.0, n_fft=4096, hop=2048, max_minutes=6.0,
              hyst_list=(0, 1, 2, 3, 4, 6), delay_list_ms=(0, 50, 100, 150, 200, 250), tilt_lo=(200, 1000),
              tilt_hi=(2000, 8000), tilt_medfilt=5, music_dbfs=-65.0, gain_search_pm_db=3.0, gain_step_db=0.5,
              T_pm_db=10.0, T_step_db=0.25) -> dict:
    """main() of src/calibrate_to_baseline_v2.py:130-313 on arrays; `json` = the dictionary it saves (without paths)."""
    delay = find_delay(orig, base, sr=sr)["delay"]
    base_start, orig_start = max(0, -delay), max(0, delay)
    avail = min(len(base) - base_start, len(orig) - orig_start, int(max_minutes * 60 * sr))
    if avail <= n_fft:
        raise ValueError("overlap too short to calibrate")
    xb, xo = base[base_start:base_start + avail], orig[orig_start:orig_start + avail]
    starts, orig_level, base_level, tilts = frame_features(xo, xb, sr, n_fft, hop, tuple(tilt_lo), tuple(tilt_hi))
    music_mask = base_level > music_dbfs
    base_state, tilts_s = baseline_states(tilts, music_mask, tilt_medfilt)
    best, gain_db0, _ = grid_search(orig_level, base_level, base_state, starts, music_mask, sr, hyst_list, delay_list_ms,
                                    gain_search_pm_db, gain_step_db, T_pm_db, T_step_db)
    if best is None:
        raise RuntimeError("no usable optimum")
    T_raw = best["T"] - best["gain_db"]
    out = dict(delay_samples_orig_minus_base=int(delay), music_dbfs=float(music_dbfs),
               gain_db_base_minus_orig=float(best["gain_db"]), T_adj_dbfs=float(best["T"]), T_raw_dbfs=float(T_raw),
               gate_ui=float(gate_ui), gate_scale=float(gate_scale), gate_offset=float(T_raw - gate_scale * gate_ui),
               hyst_db=float(best["hyst"]), up_delay_ms=float(best["up_ms"]), mismatch=float(best["mismatch"]),
               switches=int(best["switches"]))
    return dict(json=out, best=best, gain_db0=gain_db0, orig_level=orig_level, base_level=base_level, tilts=tilts,
                tilts_s=tilts_s, base_state=base_state, music_mask=music_mask, starts=starts, delay=delay)
