"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's validator kernels:
gate re-simulation and conditional spectrum of `src/validate_layer1.py` and the anchored variant of
`src/verify_tomatis_15db_v2.py`, on in-memory arrays.

Only `tests/` may import this file.  Pinned to the executed reference by
`tests/test_oracle_vs_reference.py::test_validators_*` (live, build container only) and by the frozen fixtures
`tests/golden/val_*.npz` (`oracle/make_golden.py`).
"""
from __future__ import annotations

import numpy as np

EPS = 1e-12


def rms_dbfs(x_mono) -> float:
    """src/validate_layer1.py:36-39"""
    r = np.sqrt(np.mean(x_mono * x_mono) + EPS)
    return float(20.0 * np.log10(r + EPS))


def _padded(a, n_fft):
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    z = np.zeros((n_fft // 2, a.shape[1]), dtype=a.dtype)
    return np.vstack([z, a, z])


def simulate_gate(x, sr, n_fft, hop, threshold_dbfs, hyst_db, up_delay_ms):
    """src/validate_layer1.py:110-163: levels of the frames that start inside the file, up-delay automaton with the
    delay counted in samples from the first frame at or above Ton."""
    t_on, t_off = threshold_dbfs + hyst_db / 2, threshold_dbfs - hyst_db / 2
    delay = int(up_delay_ms * sr / 1000)
    total = len(x)
    xp = _padded(x, n_fft)
    pad = n_fft // 2
    state, pending = 1, None
    states, levels = [], []
    pos = 0
    while pos + n_fft <= len(xp):
        if 0 <= pos - pad < total:
            frame = xp[pos:pos + n_fft, :]
            level = rms_dbfs(np.sqrt(np.mean(frame ** 2, axis=1)))
            if state == 1:
                if level >= t_on:
                    if pending is None:
                        pending = pos + delay
                else:
                    pending = None
                if pending is not None and pos >= pending:
                    state, pending = 2, None
            elif level <= t_off:
                state, pending = 1, None
            states.append("C1" if state == 1 else "C2")
            levels.append(level)
        pos += hop
    return states, levels


def find_stable_frames(states, margin=2):
    """src/validate_layer1.py:244-258 / src/verify_tomatis_15db_v2.py:254-267: frames whose +-margin neighbours agree."""
    c1, c2 = [], []
    for i in range(margin, len(states) - margin):
        w = set(states[i - margin:i + margin + 1])
        if w == {"C1"}:
            c1.append(i)
        elif w == {"C2"}:
            c2.append(i)
    return c1, c2


def _frame_ratio(xp, yp, start, n_fft, win, fft_dtype=None):
    """|Y| / |X| of one frame, magnitudes averaged over the channels, X floored at 1e-10
    (src/validate_layer1.py:350-361).  fft_dtype='float64' evaluates the same expression with a float64 FFT (the
    conditioning check of the tests); None keeps NumPy's float32 transform like the reference."""
    fx, fy = xp[start:start + n_fft, :], yp[start:start + n_fft, :]
    ch = fx.shape[1]
    n_bins = n_fft // 2 + 1
    X = np.zeros(n_bins, dtype=np.float32)
    Y = np.zeros(n_bins, dtype=np.float32)
    for c in range(ch):
        a, b = fx[:, c] * win, fy[:, c] * win
        if fft_dtype is not None:
            a, b = a.astype(fft_dtype), b.astype(fft_dtype)
        X += np.abs(np.fft.rfft(a))
        Y += np.abs(np.fft.rfft(b))
    X /= ch
    Y /= ch
    return Y / np.maximum(X, 1e-10)


def _median_db(ratios, n_bins):
    if not ratios:
        return np.zeros(n_bins)
    return 20 * np.log10(np.median(np.array(ratios), axis=0) + EPS)


def compute_conditional_spectrum(x, y, sr, states, n_fft, hop, level_threshold=-60, fft_dtype=None):
    """src/validate_layer1.py:261-389 (the second, effective pair of loops :338-374 and the medians :376-389)."""
    xp, yp = _padded(x, n_fft), _padded(y, n_fft)
    pad = n_fft // 2
    stable = find_stable_frames(states, margin=2)
    freqs = np.fft.rfftfreq(n_fft, 1 / sr)
    win = np.hanning(n_fft).astype(np.float32)
    out, used = [], []
    for frames in stable:
        ratios, sel = [], []
        for idx in frames:
            o = idx * hop
            if o < 0 or o + n_fft > len(x):
                continue
            fx = xp[o + pad:o + pad + n_fft, :]
            if rms_dbfs(np.sqrt(np.mean(fx ** 2, axis=1))) < level_threshold:
                continue
            ratios.append(_frame_ratio(xp, yp, o + pad, n_fft, win, fft_dtype))
            sel.append(idx)
        out.append(_median_db(ratios, len(freqs)))
        used.append(sel)
    return freqs, out[0], out[1], len(used[0]), len(used[1]), used


def compute_conditional_spectrum_v2(x, y, sr, states, levels, n_fft, hop, level_percentile=10, anchor_band=(900, 1100),
                                    fft_dtype=None):
    """src/verify_tomatis_15db_v2.py:270-369: frames below the level percentile dropped, every frame's ratio divided by
    its mean over the anchor band."""
    xp, yp = _padded(x, n_fft), _padded(y, n_fft)
    pad = n_fft // 2
    thr = np.percentile(levels, level_percentile)
    stable = find_stable_frames(states, margin=2)
    freqs = np.fft.rfftfreq(n_fft, 1 / sr)
    win = np.hanning(n_fft).astype(np.float32)
    anchor = (freqs >= anchor_band[0]) & (freqs <= anchor_band[1])
    out, used = [], []
    for frames in stable:
        ratios, sel = [], []
        for idx in frames:
            if levels[idx] < thr:
                continue
            o = idx * hop
            if o < 0 or o + n_fft > len(x):
                continue
            r = _frame_ratio(xp, yp, o + pad, n_fft, win, fft_dtype)
            g = np.mean(r[anchor])
            if g > 0:
                r = r / g
            ratios.append(r)
            sel.append(idx)
        out.append(_median_db(ratios, len(freqs)))
        used.append(sel)
    return freqs, out[0], out[1], len(used[0]), len(used[1]), used
