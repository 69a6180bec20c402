"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's per-channel state analyser
(`src/analyze_stereo_state.py`) on in-memory arrays.

Only `tests/` may import this file.  Pinned to the executed reference by
`tests/test_oracle_vs_reference.py::test_stereo_state_*` (live, build container only) and by the frozen fixture
`tests/golden/stereo_state_*.npz` (`oracle/make_golden.py`).
"""
from __future__ import annotations

import numpy as np

EPS = 1e-12


def rms_dbfs(x_mono) -> float:
    """src/analyze_stereo_state.py:16-19 -- one channel, float32 all the way, then a Python float."""
    r = np.sqrt(np.mean(x_mono * x_mono) + EPS)
    return float(20.0 * np.log10(r + EPS))


def format_time(seconds) -> str:
    """src/analyze_stereo_state.py:22-26"""
    m = int(seconds // 60)
    s = seconds % 60
    return f"{m}:{s:05.2f}"


def simulate_gate(levels, threshold_dbfs, hyst_db=3.0, min_hold_frames=6):
    """src/analyze_stereo_state.py:29-49 (min-hold automaton); uint8 states 1 = C1, 2 = C2."""
    t_on = threshold_dbfs + hyst_db / 2
    t_off = threshold_dbfs - hyst_db / 2
    state, since = 1, min_hold_frames
    out = np.zeros(len(levels), dtype=np.uint8)
    for i, level in enumerate(levels):
        since += 1
        if since >= min_hold_frames:
            if state == 1 and level >= t_on:
                state, since = 2, 0
            elif state == 2 and level <= t_off:
                state, since = 1, 0
        out[i] = state
    return out


def find_optimal_threshold(levels, target_c2=0.5, hyst_db=3.0, min_hold_frames=6):
    """src/analyze_stereo_state.py:52-76.  Unlike the adaptive mode's search it keeps the LAST midpoint, not the best."""
    valid = levels[levels > -70]
    if len(valid) == 0:
        return np.median(levels)
    t_low = np.percentile(valid, 5)
    t_high = np.percentile(valid, 95)
    best = np.median(valid)
    for _ in range(30):
        t_mid = (t_low + t_high) / 2
        st = simulate_gate(levels, t_mid, hyst_db, min_hold_frames)
        ratio = int((st == 2).sum()) / len(st)
        if abs(ratio - target_c2) < 0.01:
            return t_mid
        if ratio < target_c2:
            t_high = t_mid
        else:
            t_low = t_mid
        best = t_mid
    return best


def analyze(x, sr, target_c2=0.5, hyst_db=3.0, min_hold_ms=250.0, n_fft=4096, hop=2048) -> dict:
    """src/analyze_stereo_state.py:79-160 on a float32 array [N, ch >= 2]; `rows` = the CSV as csv.writer renders it."""
    x = np.asarray(x, dtype=np.float32)
    ch = x.shape[1]
    frame_ms = hop / sr * 1000
    hold = int(np.ceil(min_hold_ms / frame_ms))
    pad = n_fft // 2
    x_pad = np.vstack([np.zeros((pad, ch), dtype=x.dtype), x, np.zeros((pad, ch), dtype=x.dtype)])
    left, right, times = [], [], []
    nxt, total = 0, len(x)
    while nxt + n_fft <= len(x_pad):
        orig = nxt - pad
        if 0 <= orig < total:
            frame = x_pad[nxt:nxt + n_fft, :]
            left.append(rms_dbfs(frame[:, 0]))
            right.append(rms_dbfs(frame[:, 1]))
            times.append(orig / sr)
        nxt += hop
    left, right = np.array(left), np.array(right)
    res = dict(times=np.array(times), left_levels=left, right_levels=right, min_hold_frames=hold)
    for name, lv in (("left", left), ("right", right)):
        T = find_optimal_threshold(lv, target_c2, hyst_db, hold)
        st = simulate_gate(lv, T, hyst_db, hold)
        res[name + "_T"] = float(T)
        res[name + "_states"] = st
        res[name + "_c2"] = int((st == 2).sum()) / len(st)          # ZeroDivisionError on an empty file, like :126
    rows = [["Frame", "音频秒数(秒)", "音频时间(分:秒)", "Left_dBFS", "Left_Channel", "Right_dBFS", "Right_Channel"]]
    nm = {1: "C1", 2: "C2"}
    for i, t in enumerate(times):
        rows.append([str(i + 1), f"{t:.3f}", format_time(t), f"{left[i]:.2f}", nm[int(res['left_states'][i])],
                     f"{right[i]:.2f}", nm[int(res['right_states'][i])]])
    res["rows"] = rows
    return res
