"""Host-side scalar/table logic of the Tomatis path.

Everything that is O(n_fft) or O(1) stays on the host and is built with the SAME NumPy
expressions the reference uses, so the tables the kernels consume are bit-identical to the
reference's (SURVEY.md section 7.2 "Gain tables on host"):

* tilt curves and dB->linear   -- src/process_tomatis.py:105-158
* gate maps and thresholds     -- src/process_tomatis.py:54-103, 277-285
* threshold -> mean-square     -- inverse of rms_dbfs (src/process_tomatis.py:43-52) by bisection
                                  over float bit patterns, so the GPU can compare in the domain of
                                  the bit-exact mean square and gate decisions match exactly
* crossfade gain rows          -- src/process_tomatis_xfade.py:251-274, _adaptive.py:253-265,302-304
"""
from __future__ import annotations

import math

import numpy as np

EPS = 1e-12
PEAK_LIMIT = 0.999
N_FFT = 4096
HOP = 2048


# ------------------------------------------------------------------ gate maps
def gate_threshold_linear(gate_ui, gate_scale=1.0, gate_offset=-100.0):
    """T = scale * ui + offset  (src/process_tomatis.py:54-80)."""
    return gate_scale * gate_ui + gate_offset


def gate_threshold_log_percent(gate_ui, dynamic_range=80.0):
    """T = -DR + DR * ui / 100  (src/process_tomatis.py:82-103)."""
    return -dynamic_range + dynamic_range * gate_ui / 100.0


def hysteresis_pair(T, hysteresis_db):
    """(Ton, Toff), src/process_tomatis.py:283-284."""
    return T + hysteresis_db / 2.0, T - hysteresis_db / 2.0


def updelay_run_frames(sr, up_delay_ms, hop=HOP) -> int:
    """Number of CONSECUTIVE frames with level >= Ton after which C1 switches to C2.

    The reference arms `pending_c2_at = next_start + up_delay_samples` on the first such frame
    and switches once `next_start >= pending_c2_at` (src/process_tomatis.py:285, 373-381); frames
    advance by hop, so that is the (ceil(D/hop) + 1)-th consecutive frame."""
    d = int(sr * up_delay_ms / 1000.0)
    return max(0, -(-d // hop)) + 1


def adaptive_frame_counts(sr, min_hold_ms, xfade_ms, hop=HOP):
    """(min_hold_frames, xfade_frames), src/process_tomatis_adaptive.py:190-192."""
    frame_ms = hop / sr * 1000
    return int(np.ceil(min_hold_ms / frame_ms)), int(np.ceil(xfade_ms / frame_ms))


def xfade_frame_count(sr, xfade_ms, hop=HOP) -> int:
    """src/process_tomatis_xfade.py:153-154."""
    frame_duration_ms = hop / sr * 1000.0
    return max(1, int(np.ceil(xfade_ms / frame_duration_ms))) if xfade_ms > 0 else 0


# ------------------------------------------------------------------ level chain and its inverse
def level_from_meansq_scalar(m):
    """rms_dbfs tail for a NumPy scalar m (dtype promotion as in the reference, NEP 50)."""
    r = np.sqrt(m + EPS)
    return float(20.0 * np.log10(r + EPS))


def levels_from_meansq(m: np.ndarray) -> np.ndarray:
    """Vectorised rms_dbfs tail; returns float64 values equal to the reference's Python floats.
    float32 input keeps the reference's float32 arithmetic, float64 input its float64 arithmetic."""
    m = np.asarray(m)
    with np.errstate(divide="ignore"):
        lv = 20.0 * np.log10(np.sqrt(m + EPS) + EPS)
    return lv.astype(np.float64)


def _bits_to_float(bits: int, dtype):
    if dtype == np.float32:
        return np.array([bits], dtype=np.uint32).view(np.float32)[0]
    return np.array([bits], dtype=np.uint64).view(np.float64)[0]


def meansq_threshold_on(t_on_db: float, dtype=np.float32) -> float:
    """Smallest non-negative m of `dtype` whose level is >= t_on_db (inf if none).
    The chain m -> level is monotone non-decreasing, so `level >= Ton` <=> `m >= m_on`."""
    hi = 0x7F800000 if dtype == np.float32 else 0x7FF0000000000000      # +inf
    lo = 0
    if level_from_meansq_scalar(_bits_to_float(lo, dtype)) >= t_on_db:
        return 0.0
    with np.errstate(over="ignore", invalid="ignore"):
        if not level_from_meansq_scalar(_bits_to_float(hi, dtype)) >= t_on_db:
            return math.inf
        while hi - lo > 1:                      # invariant: level(lo) < T <= level(hi)
            mid = (lo + hi) // 2
            if level_from_meansq_scalar(_bits_to_float(mid, dtype)) >= t_on_db:
                hi = mid
            else:
                lo = mid
    return float(_bits_to_float(hi, dtype))


def meansq_threshold_off(t_off_db: float, dtype=np.float32) -> float:
    """Largest non-negative m of `dtype` whose level is <= t_off_db (-1.0 if none: m is never negative)."""
    hi = 0x7F800000 if dtype == np.float32 else 0x7FF0000000000000
    lo = 0
    if not level_from_meansq_scalar(_bits_to_float(lo, dtype)) <= t_off_db:
        return -1.0
    with np.errstate(over="ignore", invalid="ignore"):
        if level_from_meansq_scalar(_bits_to_float(hi, dtype)) <= t_off_db:
            return math.inf
        while hi - lo > 1:                      # invariant: level(lo) <= T < level(hi)
            mid = (lo + hi) // 2
            if level_from_meansq_scalar(_bits_to_float(mid, dtype)) <= t_off_db:
                lo = mid
            else:
                hi = mid
    return float(_bits_to_float(lo, dtype))


# ------------------------------------------------------------------ tilt gains
def tilt_gain_db(freqs, fc, slope_db_per_oct, low_gain_db, high_gain_db):
    """Tilt curve pivoting at fc: slope dB/oct towards each side, clamped at the platform gains
    (src/process_tomatis.py:109-158).  log2 distance is rounded to float32 as in the reference."""
    x = np.log2(np.maximum(freqs, 1.0) / fc).astype(np.float32)
    g = np.zeros_like(x, dtype=np.float32)
    below, above = x < 0, x > 0
    lo = np.sign(low_gain_db) * np.minimum(slope_db_per_oct * np.maximum(0.0, -x), abs(low_gain_db))
    hi = np.sign(high_gain_db) * np.minimum(slope_db_per_oct * np.maximum(0.0, x), abs(high_gain_db))
    g[below] = lo[below]
    g[above] = hi[above]
    return g


def db_to_lin_f32(db):
    """standard/xfade flavour: pow in the dtype of `db`, cast to float32 (src/process_tomatis.py:105-107)."""
    return (10.0 ** (db / 20.0)).astype(np.float32)


def db_to_lin_keep(db):
    """adaptive flavour: no cast (src/process_tomatis_adaptive.py:37-38)."""
    return 10 ** (np.asarray(db) / 20.0)


def hann_window(n_fft=N_FFT):
    """np.hanning(n_fft).astype(float32): symmetric Hann, end points exactly 0 (src/process_tomatis.py:266)."""
    return np.hanning(n_fft).astype(np.float32)


def tilt_curves_db(sr, n_fft, fc, slope, c1_low, c1_high, c2_low, c2_high):
    freqs = np.fft.rfftfreq(n_fft, d=1.0 / sr)
    return (tilt_gain_db(freqs, fc, slope, c1_low, c1_high), tilt_gain_db(freqs, fc, slope, c2_low, c2_high))


def alpha_ramp(xfade_frames: int) -> np.ndarray:
    """alpha value of crossfade counter k = 0..X as the reference's follower produces it on an
    upward ramp: repeated `alpha += step` in float64, snapped to the target at the ends."""
    xe = max(int(xfade_frames), 1)
    step = 1.0 / xfade_frames if xfade_frames > 0 else 1.0
    a = np.zeros(xe + 1, dtype=np.float64)
    cur = np.float64(0.0)
    for k in range(1, xe):
        cur = cur + step * np.sign(1.0 - cur)
        a[k] = cur
    a[xe] = 1.0
    return a


def gain_rows_standard(g1_db, g2_db):
    """rows: 0 = C1, 1 = C2 (src/process_tomatis.py:259-260, 392)."""
    return np.stack([db_to_lin_f32(g1_db), db_to_lin_f32(g2_db)]).astype(np.float32)


def gain_rows_xfade(g1_db, g2_db, xfade_frames):
    """rows k = 0..X.  alpha in {0,1} takes the precomputed float32 curves, 0 < alpha < 1 the dB-domain
    mix with a float64 alpha (src/process_tomatis_xfade.py:270-274)."""
    a = alpha_ramp(xfade_frames)
    rows = [db_to_lin_f32(g1_db)]
    for k in range(1, len(a) - 1):
        ak = np.float64(a[k])
        rows.append(db_to_lin_f32((1 - ak) * g1_db + ak * g2_db))
    rows.append(db_to_lin_f32(g2_db))
    return np.stack(rows).astype(np.float32)


def gain_rows_adaptive(c1_db, c2_db, xfade_frames):
    """rows k = 0..X, always the dB-domain mix in float64 then float32 (src/process_tomatis_adaptive.py:302-304)."""
    a = alpha_ramp(xfade_frames)
    rows = []
    for k in range(len(a)):
        ak = np.float64(a[k])
        rows.append(db_to_lin_keep((1 - ak) * c1_db + ak * c2_db).astype(np.float32))
    return np.stack(rows).astype(np.float32)


# ------------------------------------------------------------------ adaptive host scalars
def adaptive_attenuation(input_peak, c1_low, c2_high, headroom_margin):
    """(atten_db, atten_lin, use_f64) of src/process_tomatis_adaptive.py:201-215.

    `input_peak` must be the np.float32 max|x|.  `atten_db = max(0, ...)` is the Python int 0 when no
    attenuation is needed; then `atten_lin` is a float64 1.0 and `x * atten_lin` promotes the reference's
    whole pipeline (levels, FFT, OLA) to float64 (use_f64 = True).  Otherwise both are np.float32."""
    input_peak = np.float32(input_peak)
    with np.errstate(divide="ignore"):
        input_peak_dbfs = 20 * np.log10(input_peak + EPS)
    max_gain = max(abs(c1_low), abs(c2_high))
    atten_db = max(0, input_peak_dbfs + max_gain + headroom_margin)
    atten_lin = db_to_lin_keep(-atten_db)
    use_f64 = not isinstance(atten_db, np.floating)
    return atten_db, atten_lin, use_f64


def alpha_follow_exact(states: np.ndarray, xfade_frames: int, start_at_target: bool) -> np.ndarray:
    """The reference's float64 alpha follower (for the state CSV only; the device uses the integer
    counter).  start_at_target: adaptive (alpha[0] = target[0], _adaptive.py:257) vs xfade (starts 0.0)."""
    n = len(states)
    alpha = np.zeros(n, dtype=np.float64)
    step = 1.0 / xfade_frames if xfade_frames > 0 else 1.0
    cur = 0.0
    for i in range(n):
        target = 0.0 if states[i] == 1 else 1.0
        if (start_at_target and i == 0) or xfade_frames <= 0:
            cur = target
        else:
            diff = target - cur
            if abs(diff) <= step:
                cur = target
            else:
                cur = cur + step * np.sign(diff)
        alpha[i] = cur
    return alpha


# ------------------------------------------------------------------ frame / block / chunk geometry
FLUSH_SAFE = 48000 * 5      # src/process_tomatis.py:420 -- a sample count, whatever the sample rate


def streaming_frame_count(total: int, n_fft=N_FFT, hop=HOP) -> int:
    """Frames of standard/xfade: first frame at -n_fft/2, tail zero-padded by pad_end
    (src/process_tomatis.py:270-272, 310-312, 447-449)."""
    pad = n_fft // 2
    pad_end = (hop - ((total - n_fft) % hop)) % hop
    length = pad + total + pad_end
    return 0 if length < n_fft else (length - n_fft) // hop + 1


def wholefile_frame_count(total: int, hop=HOP) -> int:
    """Frames of adaptive: starts k*hop with 0 <= k*hop < total that fit the padded signal
    (src/process_tomatis_adaptive.py:298-300)."""
    return total // hop


def flush_chunk_blocks(n_frames: int, n_fft=N_FFT, hop=HOP):
    """Limiter chunks of the streaming modes as output-block ranges [(b0, b1), ...] (block b = positions
    [-n_fft/2 + b*hop, +hop)): the flush rule src/process_tomatis.py:419-426 + final flush :451-453.

    The rule flushes `safe = (next_start - out_base) - n_fft` samples whenever safe >= 240 000 after a frame.  With
    n_fft a multiple of hop that is periodic: the first flush happens after ceil((240000 + n_fft) / hop) frames, every
    later one ceil(240000 / hop) frames after the previous (120 and 118 frames at the defaults: chunks of 118 blocks)."""
    if n_frames <= 0:
        return []
    if n_fft % hop:
        return _flush_chunk_blocks_replay(n_frames, n_fft, hop)
    first = -(-(FLUSH_SAFE + n_fft) // hop)               # frames before the first flush
    nb1 = first - n_fft // hop                            # blocks it writes
    per = -(-FLUSH_SAFE // hop)                           # frames (= blocks) between later flushes
    out, flushed = [], 0
    if n_frames >= first:
        out.append((0, nb1))
        flushed = nb1
        k = (n_frames - first) // per
        out.extend((nb1 + i * per, nb1 + (i + 1) * per) for i in range(k))
        flushed += k * per
    if flushed < n_frames + 1:
        out.append((flushed, n_frames + 1))
    return out


def _flush_chunk_blocks_replay(n_frames: int, n_fft=N_FFT, hop=HOP):
    """Frame-by-frame replay of the same rule (any n_fft / hop); also the cross-check of the closed form in the tests."""
    if n_frames <= 0:
        return []
    out, flushed, out_base, next_start = [], 0, 0, 0
    for _ in range(n_frames):
        next_start += hop
        safe = (next_start - out_base) - n_fft
        if safe >= FLUSH_SAFE:
            nb = safe // hop
            out.append((flushed, flushed + nb))
            flushed += nb
            out_base += safe
    if flushed < n_frames + 1:
        out.append((flushed, n_frames + 1))
    return out


# ------------------------------------------------------------------------------------------------ calibration (N4)
def resample_poly_plan(n_in: int, up: int, down: int):
    """What scipy.signal.resample_poly(x, up, down) does to a float32 signal of n_in samples, as tables: the Kaiser(5)
    windowed-sinc low-pass of firwin (2*10*max(up, down) + 1 taps, cutoff 1/max(up, down), unit DC gain) cast to float32,
    times up, zero-padded so the kept outputs are centred; y[j] = upfirdn(h, x, up, down)[j + n_pre_remove], j < n_out.
    Returns dict(up, down, h float32, n_pre_remove, n_out) with up/down reduced; h is None when up == down (identity).
    The reference reaches it through find_delay_by_corr (src/calibrate_to_baseline_v2.py:60,73)."""
    import math
    g = math.gcd(int(up), int(down))
    up, down = int(up) // g, int(down) // g
    if up == down == 1:
        return dict(up=1, down=1, h=None, n_pre_remove=0, n_out=int(n_in))
    n_out = n_in * up
    n_out = n_out // down + bool(n_out % down)
    max_rate = max(up, down)
    f_c = 1.0 / max_rate
    half_len = 10 * max_rate
    numtaps = 2 * half_len + 1
    m = np.arange(0, numtaps, dtype=np.float64) - 0.5 * (numtaps - 1)
    h = f_c * np.sinc(f_c * m) * np.kaiser(numtaps, 5.0)
    h /= np.sum(h)
    h = h.astype(np.float32)
    h *= up
    n_pre_pad = down - half_len % down
    n_post_pad = 0
    n_pre_remove = (half_len + n_pre_pad) // down

    def out_len(len_h):                      # scipy.signal._upfirdn._output_len
        return (((n_in - 1) * up + len_h) - 1) // down + 1

    while out_len(numtaps + n_pre_pad + n_post_pad) < n_out + n_pre_remove:
        n_post_pad += 1
    h = np.concatenate([np.zeros(n_pre_pad, np.float32), h, np.zeros(n_post_pad, np.float32)])
    return dict(up=up, down=down, h=h, n_pre_remove=int(n_pre_remove), n_out=int(n_out))


def medfilt_zero_padded(x: np.ndarray, kernel_size: int) -> np.ndarray:
    """scipy.signal.medfilt on a 1-D array: sliding median of an odd window over the zero-extended signal
    (src/calibrate_to_baseline_v2.py:209)."""
    k = int(kernel_size)
    assert k % 2 == 1
    x = np.asarray(x)
    if x.size == 0:
        return x.copy()
    pad = np.zeros(k // 2, dtype=x.dtype)
    w = np.lib.stride_tricks.sliding_window_view(np.concatenate([pad, x, pad]), k)
    return np.sort(w, axis=1)[:, k // 2]
