"""TEST INFRASTRUCTURE ONLY -- CPU (NumPy) restatement of the reference's Tomatis path.

This file is the *oracle*: a plain NumPy restatement of what
`/root/reference/src/process_tomatis.py`, `process_tomatis_xfade.py` and
`process_tomatis_adaptive.py` compute between "samples read" and "samples written".
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import it, and only as the checker (or as the thing timed on the host cores) --
never as part of the product path.  The product (`tomatis_audio_processor_b200`) never
imports `oracle/` and raises if its CUDA library is missing.

Parity pinning: the reference ships NO golden vectors or known-answer tests for this path
(SURVEY.md section 4 / 8c).  The restatement is therefore pinned by *executing the reference
itself* in the build container (`oracle/ref_harness.py`, in-memory `soundfile` stand-in) and
(a) comparing live (`tests/test_oracle_vs_reference.py`, skipped where `/root/reference` is
absent) and (b) committing the reference's outputs as fixtures (`tests/golden/*.npz`, made by
`oracle/make_golden.py`) which `tests/test_oracle_golden.py` checks everywhere.  On the pinned
NumPy (2.3.5) the restatement is bit-identical to the reference (output PCM, chunk lengths,
per-frame levels, states, alpha).

All arithmetic deliberately uses the same NumPy calls as the reference (pocketfft via
`np.fft`, pairwise `np.mean`, scalar `np.log10`), so dtype promotion (NEP 50) and rounding
are the reference's.  The control structure is different: whole-array buffers instead of the
reference's streaming closures.

`fft_dtype="float64"` reproduces the NumPy-1.x behaviour of the same source (rfft of a
float32 frame computed in double); it is used to quantify the reference's own fp32 noise
at the ill-conditioned edges, not as the default.
"""
from __future__ import annotations


import numpy as np

EPS = 1e-12           # src/process_tomatis.py:40
PEAK_LIMIT = 0.999    # src/process_tomatis.py:41
FLUSH_SAFE = 48000 * 5  # src/process_tomatis.py:420 -- a sample count, independent of sr


# ------------------------------------------------------------------ primitives (L1)

def rms_dbfs(x_mono: np.ndarray) -> float:
    """src/process_tomatis.py:43-52 (copies: _adaptive.py:32-34, _xfade.py:22-25)."""
    r = np.sqrt(np.mean(x_mono * x_mono) + EPS)
    return float(20.0 * np.log10(r + EPS))


def frame_meansq(frame: np.ndarray):
    """The intermediate `np.mean(mono*mono)` of the level chain, with
    `mono = np.sqrt(np.mean(frame**2, axis=1))` (src/process_tomatis.py:370, :51).
    Exposed because the CUDA path reproduces THIS value bit-exactly and thresholds in its domain."""
    mono = np.sqrt(np.mean(frame ** 2, axis=1))
    return np.mean(mono * mono)


def level_from_meansq(m) -> float:
    """Tail of rms_dbfs from the mean-square scalar (same dtype promotion as the reference)."""
    r = np.sqrt(m + EPS)
    return float(20.0 * np.log10(r + EPS))


def frame_level(frame: np.ndarray) -> float:
    """src/process_tomatis.py:370-371."""
    mono = np.sqrt(np.mean(frame ** 2, axis=1))
    return rms_dbfs(mono)


def gate_ui_to_dbfs(gate_ui, gate_scale=1.0, gate_offset=-100.0):
    """src/process_tomatis.py:54-80."""
    return gate_scale * gate_ui + gate_offset


def gate_ui_to_dbfs_log_percent(gate_ui, dynamic_range=80.0):
    """src/process_tomatis.py:82-103."""
    return -dynamic_range + dynamic_range * gate_ui / 100.0


def db_to_lin_f32(db):
    """src/process_tomatis.py:105-107 (pow in the dtype of `db`, then float32)."""
    return (10.0 ** (db / 20.0)).astype(np.float32)


def db_to_lin_any(db):
    """src/process_tomatis_adaptive.py:37-38 (no cast)."""
    return 10 ** (np.asarray(db) / 20.0)


def build_tilt_gain_db(freqs, fc, slope_db_per_oct, low_gain_db, high_gain_db):
    """src/process_tomatis.py:109-158 (identical copies in _adaptive.py:41-54, _xfade.py:38-52)."""
    f = np.maximum(freqs, 1.0)
    x = np.log2(f / fc).astype(np.float32)
    g = np.zeros_like(x, dtype=np.float32)
    d_low = slope_db_per_oct * np.maximum(0.0, -x)
    g_low = np.sign(low_gain_db) * np.minimum(d_low, abs(low_gain_db))
    g[x < 0] = g_low[x < 0]
    d_hi = slope_db_per_oct * np.maximum(0.0, x)
    g_hi = np.sign(high_gain_db) * np.minimum(d_hi, abs(high_gain_db))
    g[x > 0] = g_hi[x > 0]
    return g


def _rfft(v, fft_dtype):
    if fft_dtype == "float64":
        v = v.astype(np.float64)
    return np.fft.rfft(v)


# ------------------------------------------------------------------ standard / xfade (streaming modes)

def frame_layout_streaming(total: int, n_fft: int, hop: int):
    """Frame starts of the streaming modes: first frame at -pad, tail zero-padded by pad_end
    (src/process_tomatis.py:270-272, 310-312, 364-367, 447-449).  Returns (pad, pad_end, starts)."""
    pad = n_fft // 2
    pad_end = (hop - ((total - n_fft) % hop)) % hop
    length = pad + total + pad_end                 # samples available to the frame loop
    starts = []
    s = -pad
    while (s + pad) + n_fft <= length:
        starts.append(s)
        s += hop
    return pad, pad_end, starts


def flush_schedule(n_frames: int, n_fft: int, hop: int):
    """Chunk boundaries (in absolute sample positions, origin = first input sample) produced by
    the periodic flush rule src/process_tomatis.py:419-426 plus the final flush :451-453.
    Returns a list of (abs_start, abs_end) -- unclipped; write_clamped clips to [0,total)."""
    pad = n_fft // 2
    out_base = -pad
    next_start = -pad
    chunks = []
    for _ in range(n_frames):
        next_start += hop
        safe = (next_start - out_base) - n_fft
        if safe >= FLUSH_SAFE:
            chunks.append((out_base, out_base + safe))
            out_base += safe
    end = (next_start - hop) + n_fft if n_frames > 0 else out_base
    if end > out_base:
        chunks.append((out_base, end))
    return chunks


def _process_streaming(x, sr, *, Ton, Toff, up_delay_ms, fc, slope, c1_low, c1_high, c2_low, c2_high,
                       n_fft, hop, xfade_ms, xfade_mode, output_gain_db, fft_dtype):
    x = np.ascontiguousarray(x, dtype=np.float32)
    total, ch = x.shape
    freqs = np.fft.rfftfreq(n_fft, d=1.0 / sr)
    g1_db = build_tilt_gain_db(freqs, fc, slope, c1_low, c1_high)
    g2_db = build_tilt_gain_db(freqs, fc, slope, c2_low, c2_high)
    g1 = db_to_lin_f32(g1_db)
    g2 = db_to_lin_f32(g2_db)
    win = np.hanning(n_fft).astype(np.float32)
    win2 = (win * win).astype(np.float32)

    pad, pad_end, starts = frame_layout_streaming(total, n_fft, hop)
    up_delay_samples = int(sr * up_delay_ms / 1000.0)

    if xfade_mode:
        frame_duration_ms = hop / sr * 1000.0                                   # _xfade.py:153
        xfade_frames = max(1, int(np.ceil(xfade_ms / frame_duration_ms))) if xfade_ms > 0 else 0
        alpha_step = 1.0 / xfade_frames if xfade_frames > 0 else 1.0
    else:
        xfade_frames, alpha_step = 0, 1.0

    xp = np.zeros((pad + total + pad_end, ch), dtype=np.float32)
    xp[pad:pad + total] = x
    n_frames = len(starts)
    span = (n_frames - 1) * hop + n_fft if n_frames else 0
    out_buf = np.zeros((span, ch), dtype=np.float32)
    w_buf = np.zeros((span,), dtype=np.float32)

    state = 1
    pending_c2_at = None
    current_alpha = 0.0
    levels = np.zeros(n_frames, dtype=np.float64)
    meansq = np.zeros(n_frames, dtype=np.float32)
    states = np.zeros(n_frames, dtype=np.uint8)
    alphas = np.zeros(n_frames, dtype=np.float64)

    for k, next_start in enumerate(starts):
        rel = next_start + pad
        frame = xp[rel:rel + n_fft, :]
        mono = np.sqrt(np.mean(frame ** 2, axis=1))
        m = np.mean(mono * mono)
        level = level_from_meansq(m)
        meansq[k] = m
        levels[k] = level

        # gate state machine, src/process_tomatis.py:373-385 (= _xfade.py:237-249)
        if state == 1:
            if level >= Ton:
                if pending_c2_at is None:
                    pending_c2_at = next_start + up_delay_samples
            else:
                pending_c2_at = None
            if pending_c2_at is not None and next_start >= pending_c2_at:
                state = 2
                pending_c2_at = None
        else:
            if level <= Toff:
                state = 1
                pending_c2_at = None
        states[k] = state

        if xfade_mode:
            # _xfade.py:251-274
            target_alpha = 0.0 if state == 1 else 1.0
            if xfade_frames > 0:
                diff = target_alpha - current_alpha
                if abs(diff) <= alpha_step:
                    current_alpha = target_alpha
                else:
                    current_alpha += alpha_step * np.sign(diff)
            else:
                current_alpha = target_alpha
            if xfade_ms > 0 and 0 < current_alpha < 1:
                mixed_gain_db = (1 - current_alpha) * g1_db + current_alpha * g2_db
                gain = db_to_lin_f32(mixed_gain_db)
            else:
                gain = g1 if current_alpha < 0.5 else g2
            alphas[k] = current_alpha
        else:
            gain = g1 if state == 1 else g2
            alphas[k] = 0.0 if state == 1 else 1.0

        y = np.zeros_like(frame, dtype=np.float32)
        for c in range(ch):
            X = _rfft(frame[:, c] * win, fft_dtype)
            X *= gain
            y[:, c] = np.fft.irfft(X, n=n_fft).astype(np.float32) * win
        out_buf[rel:rel + n_fft, :] += y
        w_buf[rel:rel + n_fft] += win2

    # flushes: normalise, clip to [0,total), optional output gain, per-chunk peak limit
    chunks = []
    for (a, b) in flush_schedule(n_frames, n_fft, hop):
        y_out = out_buf[a + pad:b + pad, :] / (w_buf[a + pad:b + pad, None] + EPS)
        s, e = max(0, a), min(total, b)
        if e <= s:
            continue
        out_chunk = y_out[s - a:e - a]
        if output_gain_db != 0.0:
            out_chunk = out_chunk * (10.0 ** (output_gain_db / 20.0))
        peak = np.max(np.abs(out_chunk))
        if peak > PEAK_LIMIT:
            out_chunk = out_chunk * (PEAK_LIMIT / peak)
        chunks.append(out_chunk)
    out = np.concatenate(chunks, axis=0) if chunks else np.zeros((0, ch), np.float32)
    in_file = np.array([0 <= s < total for s in starts], dtype=bool)
    return dict(out=out, chunk_lengths=[len(c) for c in chunks], levels=levels, meansq=meansq,
                states=states, alphas=alphas, frame_starts=np.array(starts, dtype=np.int64),
                csv_mask=in_file, Ton=Ton, Toff=Toff, up_delay_samples=up_delay_samples,
                xfade_frames=xfade_frames, pad_end=pad_end, sr=sr)


def process_standard(x, sr, gate_ui=50, gate_mode="log_percent", dynamic_range=80.0, gate_scale=1.0,
                     gate_offset=-100, hysteresis_db=3.0, fc=1000.0, slope=12.0, c1_low=+15.0,
                     c1_high=-15.0, c2_low=-15.0, c2_high=+15.0, up_delay_ms=250.0, n_fft=4096,
                     hop=2048, output_gain_db=0.0, fft_dtype=None):
    """Restates process() of src/process_tomatis.py:160-479 on an in-memory [N,2] float32 array."""
    if gate_mode == "log_percent":
        T = gate_ui_to_dbfs_log_percent(gate_ui, dynamic_range)
    else:
        T = gate_ui_to_dbfs(gate_ui, gate_scale, gate_offset)
    Ton = T + hysteresis_db / 2.0
    Toff = T - hysteresis_db / 2.0
    return _process_streaming(x, sr, Ton=Ton, Toff=Toff, up_delay_ms=up_delay_ms, fc=fc, slope=slope,
                              c1_low=c1_low, c1_high=c1_high, c2_low=c2_low, c2_high=c2_high,
                              n_fft=n_fft, hop=hop, xfade_ms=0.0, xfade_mode=False,
                              output_gain_db=output_gain_db, fft_dtype=fft_dtype)


def process_xfade(x, sr, gate_ui=50, gate_scale=1.0, gate_offset=-100, hysteresis_db=3.0, fc=1000.0,
                  slope=12.0, c1_low=+15.0, c1_high=-15.0, c2_low=-15.0, c2_high=+15.0,
                  up_delay_ms=250.0, xfade_ms=0.0, n_fft=4096, hop=2048, fft_dtype=None):
    """Restates process() of src/process_tomatis_xfade.py:55-359."""
    T = gate_ui_to_dbfs(gate_ui, gate_scale, gate_offset)
    Ton = T + hysteresis_db / 2.0
    Toff = T - hysteresis_db / 2.0
    return _process_streaming(x, sr, Ton=Ton, Toff=Toff, up_delay_ms=up_delay_ms, fc=fc, slope=slope,
                              c1_low=c1_low, c1_high=c1_high, c2_low=c2_low, c2_high=c2_high,
                              n_fft=n_fft, hop=hop, xfade_ms=xfade_ms, xfade_mode=True,
                              output_gain_db=0.0, fft_dtype=fft_dtype)


# ------------------------------------------------------------------ adaptive (whole-file mode)

def compute_frame_levels(x, sr, n_fft, hop, silence_threshold=-70):
    """src/process_tomatis_adaptive.py:57-84.  Also returns the mean-square intermediates."""
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    ch = x.shape[1]
    pad_len = n_fft // 2
    x_pad = np.vstack([np.zeros((pad_len, ch), dtype=x.dtype), x, np.zeros((pad_len, ch), dtype=x.dtype)])
    levels, msq = [], []
    next_start = 0
    total_frames = len(x)
    while next_start + n_fft <= len(x_pad):
        orig_start = next_start - pad_len
        if 0 <= orig_start < total_frames:
            frame = x_pad[next_start:next_start + n_fft, :]
            mono = np.sqrt(np.mean(frame ** 2, axis=1))
            m = np.mean(mono * mono)
            msq.append(m)
            levels.append(level_from_meansq(m))
        next_start += hop
    levels = np.array(levels)
    valid_mask = levels > silence_threshold
    frame_sec = hop / sr
    times = [(i + 1) * frame_sec for i in range(len(levels))]
    return levels, valid_mask, times, np.array(msq)


def simulate_gate(levels, threshold_dbfs, hyst_db=3.0, min_hold_frames=6):
    """src/process_tomatis_adaptive.py:87-121; returns uint8 states (1=C1, 2=C2)."""
    Ton = threshold_dbfs + hyst_db / 2
    Toff = threshold_dbfs - hyst_db / 2
    state = 1
    states = np.zeros(len(levels), dtype=np.uint8)
    frames_since_switch = min_hold_frames
    for i, level in enumerate(levels):
        frames_since_switch += 1
        if frames_since_switch >= min_hold_frames:
            if state == 1:
                if level >= Ton:
                    state = 2
                    frames_since_switch = 0
            else:
                if level <= Toff:
                    state = 1
                    frames_since_switch = 0
        states[i] = state
    return states


def find_optimal_threshold(levels, valid_mask, hyst_db=3.0, min_hold_frames=6, target_c2=0.5):
    """src/process_tomatis_adaptive.py:124-154.  Returns (best_T, trace of (T_mid, c2_ratio))."""
    valid_levels = levels[valid_mask]
    trace = []
    if len(valid_levels) == 0:
        return np.median(levels), trace
    T_low = np.percentile(valid_levels, 5)
    T_high = np.percentile(valid_levels, 95)
    best_T = np.median(valid_levels)
    best_diff = 1.0
    for _ in range(30):
        T_mid = (T_low + T_high) / 2
        states = simulate_gate(levels, T_mid, hyst_db, min_hold_frames)
        c2_ratio = int(np.sum(states == 2)) / len(states)
        trace.append((float(T_mid), c2_ratio))
        diff = abs(c2_ratio - target_c2)
        if diff < best_diff:
            best_diff = diff
            best_T = T_mid
        if diff < 0.01:
            break
        if c2_ratio < target_c2:
            T_high = T_mid
        else:
            T_low = T_mid
    return best_T, trace


def alpha_follow(states, xfade_frames):
    """src/process_tomatis_adaptive.py:253-265."""
    target_alpha = np.array([0.0 if s == 1 else 1.0 for s in states])
    alpha = np.zeros_like(target_alpha)
    if len(alpha) == 0:
        return alpha
    alpha[0] = target_alpha[0]
    step = 1.0 / xfade_frames if xfade_frames > 0 else 1.0
    for i in range(1, len(alpha)):
        diff = target_alpha[i] - alpha[i - 1]
        if abs(diff) <= step:
            alpha[i] = target_alpha[i]
        else:
            alpha[i] = alpha[i - 1] + step * np.sign(diff)
    return alpha


def process_adaptive(x, sr, fc=1000.0, slope=12.0, c1_low=15.0, c1_high=-15.0, c2_low=-15.0, c2_high=15.0,
                     target_c2=0.5, hyst_db=3.0, min_hold_ms=250.0, xfade_ms=500.0, headroom_margin=2.0,
                     n_fft=4096, hop=2048, fft_dtype=None):
    """Restates process() of src/process_tomatis_adaptive.py:157-373."""
    x = np.asarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    ch = x.shape[1]
    total_frames = len(x)

    frame_ms = hop / sr * 1000
    min_hold_frames = int(np.ceil(min_hold_ms / frame_ms))
    xfade_frames = int(np.ceil(xfade_ms / frame_ms))

    input_peak = np.max(np.abs(x))
    input_peak_dbfs = 20 * np.log10(input_peak + EPS)
    max_gain = max(abs(c1_low), abs(c2_high))
    atten_db = max(0, input_peak_dbfs + max_gain + headroom_margin)
    atten_lin = db_to_lin_any(-atten_db)
    x_atten = x * atten_lin                      # float32, or float64 when atten_db is int 0

    levels, valid_mask, times, msq = compute_frame_levels(x_atten, sr, n_fft, hop)
    optimal_T, trace = find_optimal_threshold(levels, valid_mask, hyst_db, min_hold_frames, target_c2)
    states = simulate_gate(levels, optimal_T, hyst_db, min_hold_frames)
    alpha = alpha_follow(states, xfade_frames)

    freqs = np.fft.rfftfreq(n_fft, 1 / sr)
    c1_gain_db = build_tilt_gain_db(freqs, fc, slope, c1_low, c1_high)
    c2_gain_db = build_tilt_gain_db(freqs, fc, slope, c2_low, c2_high)
    win = np.hanning(n_fft).astype(np.float32)

    pad_len = n_fft // 2
    x_pad = np.vstack([np.zeros((pad_len, ch), dtype=x_atten.dtype), x_atten,
                       np.zeros((pad_len, ch), dtype=x_atten.dtype)])
    y = np.zeros_like(x_atten)
    norm = np.zeros(total_frames, dtype=np.float32)

    next_start = 0
    frame_idx = 0
    while next_start + n_fft <= len(x_pad):
        orig_start = next_start - pad_len
        if 0 <= orig_start < total_frames and frame_idx < len(states):
            a = alpha[frame_idx]
            mixed_gain_db = (1 - a) * c1_gain_db + a * c2_gain_db
            gain = db_to_lin_any(mixed_gain_db).astype(np.float32)
            frame = x_pad[next_start:next_start + n_fft, :]
            y_frame = np.zeros_like(frame)
            for c in range(ch):
                X = _rfft(frame[:, c] * win, fft_dtype)
                X *= gain
                y_frame[:, c] = np.fft.irfft(X, n_fft) * win
            write_start = max(0, orig_start)
            write_end = min(total_frames, orig_start + n_fft)
            fs = write_start - orig_start
            fe = write_end - orig_start
            y[write_start:write_end] += y_frame[fs:fe]
            norm[write_start:write_end] += win[fs:fe] ** 2
            frame_idx += 1
        next_start += hop

    norm = np.maximum(norm, 1e-8)
    for c in range(ch):
        y[:, c] /= norm
    if atten_db > 0:
        restore_lin = db_to_lin_any(atten_db)
        y *= restore_lin
    output_peak = np.max(np.abs(y)) if y.size else 0.0
    scale = None
    if output_peak > PEAK_LIMIT:
        scale = PEAK_LIMIT / output_peak
        y *= scale
    return dict(out=y, chunk_lengths=[len(y)], levels=levels, meansq=msq, states=states, alphas=alpha,
                times=times, optimal_T=float(optimal_T), trace=trace, atten_db=float(atten_db),
                pipeline_dtype=str(x_atten.dtype), min_hold_frames=min_hold_frames,
                xfade_frames=xfade_frames, output_peak=float(output_peak), limiter_scale=scale, sr=sr)


# ------------------------------------------------------------------ CSV formatting (a14)

def csv_rows(mode: str, res: dict):
    """State-CSV rows as the reference's csv.writer would emit them (lists of str).
    standard: src/process_tomatis.py:305,408-409; xfade: _xfade.py:180,293-295;
    adaptive: _adaptive.py:355-362."""
    rows = []
    if mode == "adaptive":
        rows.append(['frame_idx', 'time_sec', 'level_dbfs', 'state', 'alpha'])
        for i, (t, lvl, st) in enumerate(zip(res["times"], res["levels"], res["states"])):
            a = res["alphas"][i]
            rows.append([str(i + 1), f'{t:.6f}', f'{lvl:.4f}', 'C1' if st == 1 else 'C2', f'{a:.4f}'])
        return rows
    sr = res["sr"]
    if mode == "standard":
        rows.append(["frame_idx", "time_sec", "level_dbfs", "state"])
    else:
        rows.append(["frame_idx", "time_sec", "level_dbfs", "state", "alpha"])
    for k, start in enumerate(res["frame_starts"]):
        if not res["csv_mask"][k]:
            continue
        st = "C1" if res["states"][k] == 1 else "C2"
        lvl = float(res["levels"][k])
        if mode == "standard":
            rows.append([str(k), str(int(start) / sr), str(lvl), st])
        else:
            rows.append([str(k), str(int(start) / sr), f"{lvl:.2f}", st, f"{res['alphas'][k]:.3f}"])
    return rows


PROCESSORS = {"standard": process_standard, "xfade": process_xfade, "adaptive": process_adaptive}


def run(mode: str, x, sr, **params):
    return PROCESSORS[mode](x, sr, **params)
