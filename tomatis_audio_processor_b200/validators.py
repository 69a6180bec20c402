"""Validator kernels of the reference on the GPU: independent gate re-simulation and the
conditional spectrum, with the reference's function names, arguments and return values
(`src/validate_layer1.py:110-163,244-389`, `src/verify_tomatis_15db_v2.py:254-369`).

Frame levels, the gate automaton, the windowed FFTs of both files and the per-bin medians run in the CUDA library;
frame selection (stable frames, level threshold / percentile) is host bookkeeping on the per-frame arrays.  There is no
CPU path.  Supported: n_fft / hop = 4096 / 2048 (the reference defaults) and 2048 / 1024, one or two channels.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from . import tables as tb

EPS = 1e-12
DEVICE = 0


def _stereo(a):
    """[N] / [N,1] -> the channel twice (mean of two equal magnitudes = the single-channel magnitude); [N,2] as is."""
    a = np.asarray(a, dtype=np.float32)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.shape[1] == 1:
        return np.ascontiguousarray(np.repeat(a, 2, axis=1)), True
    if a.shape[1] != 2:
        raise NotImplementedError(f"GPU validators take mono or stereo files, got {a.shape[1]} channels")
    return np.ascontiguousarray(a), False


def _check_fft(n_fft, hop):
    from .engine import fused_size
    if not fused_size(n_fft, hop):
        raise NotImplementedError(f"GPU path implements n_fft/hop = 4096/2048 and 2048/1024 here; got {n_fft}/{hop}")


def _names(states):
    return ["C1" if int(s) == 1 else "C2" for s in states]


def _codes(states) -> np.ndarray:
    a = np.asarray(states)
    if a.dtype.kind in "US":
        return np.where(a == "C1", 1, 2).astype(np.uint8)
    return a.astype(np.uint8)


def simulate_gate(x, sr, n_fft, hop, threshold_dbfs, hyst_db, up_delay_ms):
    """Gate states recomputed from the input alone (src/validate_layer1.py:110-163) -> (states 'C1'/'C2', levels)."""
    from . import engine
    _check_fft(n_fft, hop)
    xs, mono = _stereo(x)
    if mono:
        xs[:, 1] = 0.0                                  # single channel rides in the L lane (mono level formula)
    t_on, t_off = threshold_dbfs + hyst_db / 2, threshold_dbfs - hyst_db / 2
    msq, levels, plan = engine.frame_levels_wholefile(xs, DEVICE, mono=mono, n_fft=n_fft, hop=hop)
    try:
        plan.gate(L.GATE_UPDELAY, L.ARR_MEANSQ_F32, tb.meansq_threshold_on(t_on, np.float32),
                  tb.meansq_threshold_off(t_off, np.float32), tb.updelay_run_frames(sr, up_delay_ms, hop), 0)
        states = plan.read(L.ARR_STATE)
    finally:
        plan.close()
    return _names(states), [float(v) for v in levels]


def find_stable_frames(states, margin=2):
    """Frames whose +-margin neighbours all share their state (src/validate_layer1.py:244-258) -> (C1 list, C2 list)."""
    st = _codes(states)
    n = len(st)
    if n < 2 * margin + 1:
        return [], []
    same = np.ones(n - 2 * margin, dtype=bool)
    mid = st[margin:n - margin]
    for d in range(-margin, margin + 1):
        same &= st[margin + d:n - margin + d] == mid
    idx = np.arange(margin, n - margin)
    return [int(i) for i in idx[same & (mid == 1)]], [int(i) for i in idx[same & (mid == 2)]]


def _median_db(x, y, frames, n_bins, anchor_bins=None):
    from . import engine
    if len(frames) == 0:
        return np.zeros(n_bins)                         # src/validate_layer1.py:380-381
    n_fft = 2 * (n_bins - 1)                            # n_bins = len(rfftfreq(n_fft)); the fused sizes have hop = n_fft / 2
    med = engine.cond_spectrum_median(x, y, frames, anchor_bins, DEVICE, n_fft=n_fft, hop=n_fft // 2)
    return 20 * np.log10(med + EPS)


def _inside(frames, n_x, n_y, n_fft, hop):
    frames = [i for i in frames if i * hop + n_fft <= n_x]          # `orig_start + n_fft > len(x): continue`
    if any(i * hop + n_fft > n_y for i in frames):
        raise ValueError("the output file is shorter than the input: a selected frame does not fit into it")
    return frames


def compute_conditional_spectrum(x, y, sr, states, n_fft, hop, level_threshold=-60):
    """Delta(f) = 20*log10 median_frames(|Y| / |X|) over the stable C1 and the stable C2 frames whose input level is at
    least level_threshold (src/validate_layer1.py:261-389) -> (freqs, c1_db, c2_db, n_c1, n_c2)."""
    from . import engine
    _check_fft(n_fft, hop)
    xs, mono = _stereo(x)
    ys, _ = _stereo(y)
    lx = xs.copy()
    if mono:
        lx[:, 1] = 0.0
    _, levels, plan = engine.frame_levels_wholefile(lx, DEVICE, mono=mono, n_fft=n_fft, hop=hop)
    plan.close()
    freqs = np.fft.rfftfreq(n_fft, 1 / sr)
    xd, yd = engine.to_device([xs, ys], DEVICE)
    out, used = [], []
    for frames in find_stable_frames(states, margin=2):
        frames = [i for i in _inside(frames, len(xs), len(ys), n_fft, hop) if not levels[i] < level_threshold]
        out.append(_median_db(xd, yd, frames, len(freqs)))
        used.append(len(frames))
    return freqs, out[0], out[1], used[0], used[1]


def compute_conditional_spectrum_v2(x, y, sr, states, levels, n_fft, hop, level_percentile=10, anchor_band=(900, 1100)):
    """The anchored variant (src/verify_tomatis_15db_v2.py:270-369): frames below the level percentile are dropped and
    each frame's ratio is normalised to unit mean gain over anchor_band -> (freqs, c1_db, c2_db, c1_used, c2_used)."""
    _check_fft(n_fft, hop)
    xs, _ = _stereo(x)
    ys, _ = _stereo(y)
    levels = np.asarray(levels)
    thr = np.percentile(levels, level_percentile)
    freqs = np.fft.rfftfreq(n_fft, 1 / sr)
    sel = np.nonzero((freqs >= anchor_band[0]) & (freqs <= anchor_band[1]))[0]
    anchor = (int(sel[0]), int(sel[-1])) if len(sel) else None      # no bin in the band: mean of nothing is never > 0
    from . import engine
    xd, yd = engine.to_device([xs, ys], DEVICE)
    out, used = [], []
    for frames in find_stable_frames(states, margin=2):
        frames = _inside([i for i in frames if not levels[i] < thr], len(xs), len(ys), n_fft, hop)
        out.append(_median_db(xd, yd, frames, len(freqs), anchor))
        used.append(len(frames))
    return freqs, out[0], out[1], used[0], used[1]
