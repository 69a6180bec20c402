"""Seeded synthetic stereo test signals (SURVEY.md section 8d recipes).

No audio fixtures ship with the reference (`/root/reference/.gitignore:20-22`), so every
parity test, golden vector and benchmark input is synthesised.  All generators return
interleaved stereo float32 arrays of shape [N, 2] in [-1, 1].

Host (NumPy) generators are deterministic in `seed`; `device_batch()` is the torch-on-GPU
variant used only by bench.py to fill HBM quickly (same recipe, different RNG stream).
"""
from __future__ import annotations

import numpy as np


def pink_noise(n: int, channels: int, rng: np.random.Generator) -> np.ndarray:
    """White Gaussian noise shaped by 1/sqrt(f) in the rfft domain, per channel independent,
    each channel normalised to unit RMS.  Returns float64 [n, channels]."""
    out = np.empty((n, channels), dtype=np.float64)
    f = np.fft.rfftfreq(n)
    shape = np.ones_like(f)
    shape[1:] = 1.0 / np.sqrt(f[1:] / f[1])
    shape[0] = 0.0
    for c in range(channels):
        w = rng.standard_normal(n)
        X = np.fft.rfft(w) * shape
        y = np.fft.irfft(X, n)
        y /= np.sqrt(np.mean(y * y)) + 1e-30
        out[:, c] = y
    return out


def _db(x):
    return 10.0 ** (np.asarray(x, dtype=np.float64) / 20.0)


def tone_bursts(n: int, sr: int, level_dbfs: float = -12.0, every_s: float = 2.0,
                dur_s: float = 0.2, freqs=(1000.0, 3000.0)) -> np.ndarray:
    """200 ms tone bursts alternating between `freqs`, raised-cosine edges (5 ms). float64 [n]."""
    t = np.arange(n, dtype=np.float64) / sr
    y = np.zeros(n, dtype=np.float64)
    amp = _db(level_dbfs) * np.sqrt(2.0)      # sine RMS = level
    k = 0
    start = 0.5 * every_s
    edge = 0.005
    while start < n / sr:
        a = int(start * sr)
        b = min(n, int((start + dur_s) * sr))
        if b > a:
            tt = t[a:b] - t[a]
            env = np.minimum(1.0, np.minimum(tt, tt[-1] - tt) / edge)
            env = 0.5 - 0.5 * np.cos(np.pi * env)
            y[a:b] = amp * env * np.sin(2 * np.pi * freqs[k % len(freqs)] * tt)
        k += 1
        start += every_s
    return y


def recipe_gated_pink(seconds: float, sr: int, seed: int, lo_dbfs: float = -55.0,
                      hi_dbfs: float = -25.0, env_hz: float = 0.2, bursts: bool = True,
                      peak: float | None = None) -> np.ndarray:
    """C1/C4 recipe: pink noise with an on/off RMS envelope between lo and hi dBFS
    (square wave at env_hz, 50 ms smoothed edges) plus tone bursts at -12 dBFS every 2 s."""
    n = int(round(seconds * sr))
    rng = np.random.default_rng(seed)
    noise = pink_noise(n, 2, rng)
    t = np.arange(n, dtype=np.float64) / sr
    phase = (t * env_hz + rng.uniform()) % 1.0
    edge = 0.05 * env_hz
    up = np.clip(phase / edge, 0, 1) * np.clip((0.5 - phase) / edge, 0, 1)
    up = np.clip(up, 0, 1)
    lev_db = lo_dbfs + (hi_dbfs - lo_dbfs) * up
    x = noise * _db(lev_db)[:, None]
    if bursts:
        x += tone_bursts(n, sr)[:, None]
    if peak is not None:
        x *= peak / (np.max(np.abs(x)) + 1e-30)
    np.clip(x, -1.0, 1.0, out=x)
    return x.astype(np.float32)


def recipe_swept_pink(seconds: float, sr: int, seed: int, centre_dbfs: float = -32.0,
                      sweep_db: float = 12.0, period_s: float = 7.0, peak: float | None = 0.5,
                      bursts: bool = False) -> np.ndarray:
    """C2/C5 recipe: pink noise with a slow +-sweep_db sinusoidal level sweep so the level
    distribution is continuous (adaptive's [p5, p95] threshold search needs that)."""
    n = int(round(seconds * sr))
    rng = np.random.default_rng(seed)
    noise = pink_noise(n, 2, rng)
    t = np.arange(n, dtype=np.float64) / sr
    lev_db = centre_dbfs + sweep_db * np.sin(2 * np.pi * t / period_s + rng.uniform(0, 2 * np.pi))
    x = noise * _db(lev_db)[:, None]
    if bursts:
        x += tone_bursts(n, sr)[:, None]
    if peak is not None:
        x *= peak / (np.max(np.abs(x)) + 1e-30)
    np.clip(x, -1.0, 1.0, out=x)
    return x.astype(np.float32)


def recipe_threshold_ramps(seconds: float, sr: int, seed: int, t_on: float, t_off: float,
                           margin_db: float = 6.0, period_s: float = 10.0) -> np.ndarray:
    """C3 recipe: pink noise whose RMS ramps linearly in dB across [t_off-margin, t_on+margin]
    and back every period_s -- every ramp straddles the gate thresholds."""
    n = int(round(seconds * sr))
    rng = np.random.default_rng(seed)
    noise = pink_noise(n, 2, rng)
    t = np.arange(n, dtype=np.float64) / sr
    tri = 2.0 * np.abs(((t / period_s) % 1.0) - 0.5)          # 1 -> 0 -> 1
    lo, hi = t_off - margin_db, t_on + margin_db
    lev_db = lo + (hi - lo) * (1.0 - tri)
    x = noise * _db(lev_db)[:, None]
    np.clip(x, -1.0, 1.0, out=x)
    return x.astype(np.float32)


def recipe_level_steps(seconds: float, sr: int, seed: int, lo_dbfs: float = -52.0, hi_dbfs: float = -28.0,
                       min_s: float = 0.4, max_s: float = 1.8, edge_s: float = 0.02) -> np.ndarray:
    """Calibration recipe: pink noise that alternates between a quiet and a loud plateau (each jittered by +-4 dB) held
    for random durations in [min_s, max_s] -- an aperiodic envelope, so the envelope cross-correlation of
    src/calibrate_to_baseline_v2.py:44-86 has one clear peak, and a bimodal level distribution for its gate fit."""
    n = int(round(seconds * sr))
    rng = np.random.default_rng(seed)
    noise = pink_noise(n, 2, rng)
    lev_db = np.empty(n, dtype=np.float64)
    pos, loud = 0, False
    while pos < n:
        ln = int(rng.uniform(min_s, max_s) * sr)
        lev_db[pos:pos + ln] = (hi_dbfs if loud else lo_dbfs) + rng.uniform(-4.0, 4.0)
        pos += ln
        loud = not loud
    k = max(1, int(edge_s * sr))
    lev_db = np.convolve(np.pad(lev_db, (k // 2, k - 1 - k // 2), mode="edge"), np.ones(k) / k, mode="valid")
    x = noise * _db(lev_db)[:, None]
    np.clip(x, -1.0, 1.0, out=x)
    return x.astype(np.float32)


def quantise_pcm16(x: np.ndarray) -> np.ndarray:
    """Round to the int16 grid (values stay float32 and exactly representable) so golden
    fixtures can store inputs as int16."""
    q = np.clip(np.rint(x.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)
    return q


def pcm16_to_float(q: np.ndarray) -> np.ndarray:
    return (q.astype(np.float32) / np.float32(32768.0)).astype(np.float32)


def device_batch(n_tracks: int, n_samples: int, sr: int, seed0: int, device, out=None):
    """bench.py only: fill a [n_tracks, n_samples, 2] float32 CUDA tensor with the C1/C4 recipe
    (pink noise + 0.2 Hz on/off envelope -55/-25 dBFS + tone bursts), one seed per track.
    Uses torch.fft for the 1/sqrt(f) shaping -- this is input synthesis, not the timed path."""
    import torch

    if out is None:
        out = torch.empty((n_tracks, n_samples, 2), dtype=torch.float32, device=device)
    t = torch.arange(n_samples, device=device, dtype=torch.float64) / sr
    f = torch.fft.rfftfreq(n_samples, device=device, dtype=torch.float32)
    shape = torch.ones_like(f)
    shape[1:] = torch.rsqrt(f[1:] / f[1])
    shape[0] = 0.0
    bursts = torch.from_numpy(tone_bursts(n_samples, sr).astype(np.float32)).to(device)
    g = torch.Generator(device=device)
    for i in range(n_tracks):
        g.manual_seed(seed0 + i)
        w = torch.randn((2, n_samples), generator=g, device=device, dtype=torch.float32)
        y = torch.fft.irfft(torch.fft.rfft(w, dim=1) * shape, n=n_samples, dim=1)
        y = y / (y.pow(2).mean(dim=1, keepdim=True).sqrt() + 1e-30)
        ph = float(torch.rand((), generator=g, device=device))
        phase = ((t * 0.2 + ph) % 1.0).to(torch.float32)
        edge = 0.05 * 0.2
        up = (phase / edge).clamp(0, 1) * ((0.5 - phase) / edge).clamp(0, 1)
        lev = torch.pow(10.0, (-55.0 + 30.0 * up.clamp(0, 1)) / 20.0)
        x = y * lev + bursts
        out[i].copy_(x.t().clamp_(-1.0, 1.0))
        del w, y, x
    return out


def device_long_file_range(lo: int, hi: int, sr: int, seed0: int, device, segment_seconds: float = 300.0):
    """bench.py only: samples [lo, hi) of one long synthetic file made of consecutive 5-minute segments of the
    C1/C4 recipe (segment k uses seed seed0 + k, so every rank synthesises exactly its own range of the same file)."""
    import torch

    seg = int(round(segment_seconds * sr))
    out = torch.empty((hi - lo, 2), dtype=torch.float32, device=device)
    k = lo // seg
    while k * seg < hi:
        a, b = max(lo, k * seg), min(hi, (k + 1) * seg)
        x = device_batch(1, seg, sr, seed0 + k, device)[0]
        out[a - lo:b - lo] = x[a - k * seg:b - k * seg]
        del x
        k += 1
    return out
