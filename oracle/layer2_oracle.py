"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's static-EQ processor
`/root/reference/src/layer2_apply_eq.py` on in-memory arrays.

Pinned like the Tomatis oracle: the reference has no golden vectors for it, so `oracle/ref_harness.py` executes
the reference itself (in-memory soundfile stand-in, including the PCM_24 write/read round trip its gain-protect pass
performs on its own output file) and `oracle/make_golden.py` commits those outputs as `tests/golden/eq_*.npz`;
`tests/test_oracle_golden.py` checks this restatement against them (bit-identical on NumPy 2.3.5).
Only tests/, smoke() and bench.py's CPU legs may import this file.
"""
from __future__ import annotations

import numpy as np

EPS = 1e-12          # src/layer2_apply_eq.py:6


def db_to_lin(db):
    """src/layer2_apply_eq.py:8-9."""
    return 10.0 ** (db / 20.0)


def build_gain_per_bin(sr, n_fft, eq_freqs, eq_db):
    """src/layer2_apply_eq.py:48-64: dB curve interpolated on a log-frequency axis onto the rfft bins."""
    f_bins = np.fft.rfftfreq(n_fft, 1.0 / sr).astype(np.float32)
    f_safe = np.maximum(f_bins, 1.0)
    x = np.log10(np.maximum(eq_freqs, 1.0))
    xb = np.log10(f_safe)
    yb = np.interp(xb, x, eq_db, left=eq_db[0], right=eq_db[-1]).astype(np.float32)
    return db_to_lin(yb).astype(np.float32)


def pcm24_roundtrip(y):
    """What soundfile returns (dtype float32) after writing `y` as FLAC PCM_24 (the container of the gain-protect pass,
    src/layer2_apply_eq.py:226): libsndfile's clipping conversion lrint(y * 2^23) pinned to [-2^23, 2^23 - 1]
    (src/flac.c f2flac24_clip_array; python-soundfile switches clipping on), read back as value / 2^23."""
    q = np.clip(np.rint(np.asarray(y, dtype=np.float64) * 8388608.0), -8388608, 8388607)
    return (q / 8388608.0).astype(np.float32)


def apply_eq(x, sr, gain_bins, n_fft=4096, hop=2048, pad=True, global_gain_db=0.0, auto_gain_protect=True,
             peak_target=0.99, fft_dtype=None):
    """Restates apply_eq_stft (src/layer2_apply_eq.py:66-237) between "samples read" and "samples written".
    Returns dict(out = float array handed to the first file, peak_seen, scale, out_gp = array handed to the
    gain-protected second file or None)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    N, ch = x.shape
    win = np.hanning(n_fft).astype(np.float32)
    win2 = (win * win).astype(np.float32)
    pad_len = n_fft // 2 if pad else 0
    g_global = db_to_lin(global_gain_db)
    xs = (x * g_global).astype(np.float32)                                  # :146
    buf = np.vstack([np.zeros((pad_len, ch), np.float32), xs, np.zeros((pad_len, ch), np.float32)]) if pad_len else xs
    n_frames = (len(buf) - n_fft) // hop + 1 if len(buf) >= n_fft else 0
    span = (n_frames - 1) * hop + n_fft if n_frames else 0
    out_buf = np.zeros((span, ch), np.float32)
    w_buf = np.zeros((span,), np.float32)
    for k in range(n_frames):
        frame = buf[k * hop:k * hop + n_fft, :]
        y = np.zeros_like(frame, np.float32)
        for c in range(ch):
            fr = frame[:, c] * win
            X = np.fft.rfft(fr.astype(np.float64) if fft_dtype == "float64" else fr)
            X *= gain_bins
            y[:, c] = np.fft.irfft(X, n=n_fft).astype(np.float32) * win
        out_buf[k * hop:k * hop + n_fft, :] += y
        w_buf[k * hop:k * hop + n_fft] += win2
    seg = out_buf / (w_buf[:, None] + EPS)
    peak_seen = float(np.max(np.abs(seg))) if seg.size else 0.0
    scale, out_gp = None, None
    if auto_gain_protect and peak_seen > peak_target:
        scale = peak_target / max(peak_seen, EPS)
        out_gp = (pcm24_roundtrip(seg) * scale).astype(np.float32)          # second pass re-reads the PCM_24 file, :226-231
    return dict(out=seg, peak_seen=peak_seen, scale=scale, out_gp=out_gp, n_frames=n_frames, pad_len=pad_len)
