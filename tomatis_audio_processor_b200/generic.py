"""General FFT sizes: every `--n_fft` / `--hop` other than the fused kernels' 4096 / 2048 (TMT_GENERIC_FFT=0 turns it off).

The reference exposes `--n_fft` / `--hop` (`src/process_tomatis.py:509-510`, `_adaptive.py:396-397`, `_xfade.py:388-389`);
the fused kernels implement the defaults 4096 / 2048 and everything else raises NotImplementedError.  This module is the
plain path for the other sizes (power-of-two n_fft in [128, 8192], any hop in [1, n_fft]) on the `tmt_generic_*` entry
points of the C ABI (csrc/generic.cuh): per-frame levels in NumPy's pairwise order, the gate automata of the main path (run
on an audio-less plan over the frames), one CTA per frame for window -> FFT -> gain -> IFFT -> window in double precision
with the reference's float32 roundings around it, gather overlap-add in the reference's frame order, limiter.  Coverage, not
speed.  The per-thread arithmetic is checked on the CPU against the oracle (tests/test_generic_sizes.py, through
csrc/host_emul.cu) and the CUDA path on the B200 (same file, -m gpu; first hardware run: profiles/r02/generic_first_run.md).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence

import numpy as np

from . import _lib as L
from . import tables as tb

STREAMING, ADAPTIVE_F32, ADAPTIVE_F64 = 0, 1, 2          # kGenStreaming / kGenAdaptiveF32 / kGenAdaptiveF64


def enabled() -> bool:
    return os.environ.get("TMT_GENERIC_FFT", "1") != "0"


def check_sizes(n_fft: int, hop: int):
    if n_fft < 128 or n_fft > 8192 or n_fft & (n_fft - 1):
        raise NotImplementedError(f"general path: n_fft must be a power of two in [128, 8192], got {n_fft}")
    if not 1 <= hop <= n_fft:
        raise NotImplementedError(f"general path: hop must lie in [1, n_fft], got {hop}")


# ------------------------------------------------------------------------------------------------ frame geometry
def streaming_layout(total: int, n_fft: int, hop: int):
    """(first_start, n_frames) of standard / xfade: frames from -n_fft/2 over the file padded by pad_end
    (src/process_tomatis.py:270-272,310-312,447-449)."""
    return -(n_fft // 2), tb.streaming_frame_count(total, n_fft, hop)


def adaptive_layout(total: int, n_fft: int, hop: int):
    """(first_start, n_frames) of adaptive: frame starts k*hop - n_fft/2 of the signal padded by n_fft/2 on both sides, kept
    while 0 <= start < total and the frame fits the padded signal (src/process_tomatis_adaptive.py:62-82,296-325)."""
    pad = n_fft // 2
    first = -(-pad // hop) * hop - pad                      # first k*hop - pad that is >= 0
    last = min(total - 1, total + pad - n_fft)              # start < total and start + pad + n_fft <= total + 2 * pad
    return first, (0 if last < first else (last - first) // hop + 1)


def flush_sample_ranges(n_frames: int, total: int, n_fft: int, hop: int):
    """Limiter chunks of the streaming modes as file sample ranges [(s0, s1)], clipped to the file, empty ones dropped: the
    rule of src/process_tomatis.py:419-426 flushes everything no later frame touches once that is >= 240 000 samples, i.e.
    the first time after ceil((240000 + n_fft) / hop) frames and then every ceil(240000 / hop) frames; the final flush
    (:451-453) takes the rest."""
    if n_frames <= 0:
        return []
    pad = n_fft // 2
    first_flush = -(-(tb.FLUSH_SAFE + n_fft) // hop)
    period = -(-tb.FLUSH_SAFE // hop)
    cuts = [-pad]
    f = first_flush
    while f <= n_frames:
        cuts.append(-pad + f * hop - n_fft)
        f += period
    cuts.append(-pad + (n_frames - 1) * hop + n_fft)
    out = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        s, e = max(0, a), min(total, b)
        if b > a and e > s:
            out.append((s, e))
    return out


# ------------------------------------------------------------------------------------------------ kernels (CUDA)
class CudaKernels:
    """The device side: torch tensors as buffers, tmt_generic_* for the arithmetic, an audio-less plan for the gate."""

    def __init__(self, device: int = 0):
        from . import engine
        self.engine_mod = engine
        self.torch = engine._torch()
        self.eng = engine.get_engine(device)
        self.device = device
        self.launches = 0

    def _s(self):
        return self.engine_mod._stream_ptr(self.torch)

    def upload(self, x: np.ndarray):
        return self.torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(f"cuda:{self.device}")

    def upload_tables(self, win, gains):
        t = self.torch
        return (t.from_numpy(np.ascontiguousarray(win, np.float32)).to(f"cuda:{self.device}"),
                t.from_numpy(np.ascontiguousarray(gains, np.float32)).to(f"cuda:{self.device}"))

    def input_peak(self, xd) -> np.float32:
        n = int(xd.shape[0])
        plan = self.engine_mod.Plan(self.eng, L.FRAMING_WHOLEFILE, [L.TrackDesc(xd.data_ptr(), None, n, 0, n, 0, 0, 0, -1)])
        try:
            plan.input_peaks()
            return np.float32(plan.read(L.ARR_INPUT_PEAK)[0])
        finally:
            plan.close()

    def meansq(self, xd, total, first, n_fft, hop, n_frames, use_f64, sc, mono):
        out = self.torch.empty(max(1, n_frames), dtype=self.torch.float64 if use_f64 else self.torch.float32, device=xd.device)
        L.check(self.eng.lib.tmt_generic_meansq(self.eng.h, C.c_void_p(xd.data_ptr()), total, first, n_fft, hop, n_frames, int(use_f64),
                                                float(sc), int(mono), C.c_void_p(out.data_ptr()), self._s()), "tmt_generic_meansq")
        self.launches += 1
        return out[:n_frames].cpu().numpy()

    def _gate_plan(self, n_frames):
        """A plan without audio whose per-frame arrays have n_frames entries (whole-file framing counts total // 2048 frames)."""
        self._dummy = self.torch.zeros(8, dtype=self.torch.float32, device=f"cuda:{self.device}")
        fake = n_frames * tb.HOP
        return self.engine_mod.Plan(self.eng, L.FRAMING_WHOLEFILE, [L.TrackDesc(self._dummy.data_ptr(), None, fake, 0, fake, 0, 0, 0, -1)])

    def gate(self, automaton, values, on, off, param, xfade_frames, init_to_target=False, count_only=False):
        """values: float32 mean squares (up-delay automaton) or float64 levels (min-hold).  Returns the C2 count if count_only,
        else (states uint8, rows uint16)."""
        n = int(len(values))
        if getattr(self, "_plan", None) is None or self._plan_n != n or self._plan_vals is not values:
            self.close_gate()
            self._plan, self._plan_n, self._plan_vals = self._gate_plan(n), n, values
            which = L.ARR_MEANSQ_F32 if values.dtype == np.float32 else L.ARR_GATE_F64
            self._plan.write(which, values)
            self._plan_which = which
        self._plan.gate(automaton, self._plan_which, on, off, param, xfade_frames, alpha_init_to_target=init_to_target, count_only=count_only)
        self.launches += 1
        if count_only:
            return int(self._plan.read(L.ARR_C2_COUNT)[0])
        return self._plan.read(L.ARR_STATE), self._plan.read(L.ARR_ROW)

    def close_gate(self):
        if getattr(self, "_plan", None) is not None:
            self._plan.close()
            self._plan = None

    def frames(self, xd, total, first, n_fft, hop, n_frames, win_d, gains_d, rows, sc, flavour):
        t = self.torch
        rows_d = t.from_numpy(np.ascontiguousarray(rows, np.uint16).view(np.int16)).to(xd.device)
        shape = (max(1, n_frames), n_fft, 2)
        fr = t.empty(shape, dtype=t.float64 if flavour == ADAPTIVE_F64 else t.float32, device=xd.device)
        L.check(self.eng.lib.tmt_generic_frames(self.eng.h, C.c_void_p(xd.data_ptr()), total, first, n_fft, hop, n_frames,
                                                C.c_void_p(win_d.data_ptr()), C.c_void_p(gains_d.data_ptr()),
                                                C.c_void_p(rows_d.data_ptr()), float(sc), flavour, C.c_void_p(fr.data_ptr()),
                                                self._s()), "tmt_generic_frames")
        self.launches += 1
        return fr

    def overlap_add(self, fr, flavour, total, first, n_fft, hop, n_frames, win_d, post):
        t = self.torch
        y = t.empty((max(1, total), 2), dtype=t.float64 if flavour == ADAPTIVE_F64 else t.float32, device=fr.device)
        L.check(self.eng.lib.tmt_generic_overlap_add(self.eng.h, C.c_void_p(fr.data_ptr()), flavour, total, first, n_fft, hop, n_frames,
                                                     C.c_void_p(win_d.data_ptr()), float(post), C.c_void_p(y.data_ptr()), self._s()),
                "tmt_generic_overlap_add")
        self.launches += 1
        return y[:total]

    def limit(self, y, use_f64, bounds) -> np.ndarray:
        t = self.torch
        b = np.ascontiguousarray(bounds, dtype=np.int64).reshape(-1, 2)
        if len(b) == 0:
            return np.zeros(0, np.float64 if use_f64 else np.float32)
        bd = t.from_numpy(b).to(y.device)
        peaks = t.zeros(len(b), dtype=t.float64 if use_f64 else t.float32, device=y.device)
        L.check(self.eng.lib.tmt_generic_limit(self.eng.h, C.c_void_p(y.data_ptr()), int(use_f64), C.c_void_p(bd.data_ptr()), len(b),
                                               float(tb.PEAK_LIMIT), C.c_void_p(peaks.data_ptr()), self._s()), "tmt_generic_limit")
        self.launches += 2
        return peaks.cpu().numpy()

    def to_host(self, y, use_f64) -> np.ndarray:
        if not use_f64:
            return y.cpu().numpy()
        t = self.torch
        out = t.empty((max(1, y.shape[0]), 2), dtype=t.float32, device=y.device)
        L.check(self.eng.lib.tmt_generic_to_float(self.eng.h, C.c_void_p(y.data_ptr()), int(y.shape[0]), C.c_void_p(out.data_ptr()),
                                                  self._s()), "tmt_generic_to_float")
        self.launches += 1
        return out[:y.shape[0]].cpu().numpy()


# ------------------------------------------------------------------------------------------------ the three modes
def run_streaming(mode: str, xs: Sequence[np.ndarray], sr: int, kernels=None, device: int = 0, **params) -> List[dict]:
    """standard / xfade for any supported n_fft / hop; one result dict per track with the fields of engine.run_streaming."""
    from .engine import streaming_params
    n_fft, hop = int(params.get("n_fft", tb.N_FFT)), int(params.get("hop", tb.HOP))
    check_sizes(n_fft, hop)
    k = kernels or CudaKernels(device)
    sp = streaming_params(mode, sr, **params)
    win_d, gains_d = k.upload_tables(tb.hann_window(n_fft), sp.rows)
    res = []
    try:
        for x in xs:
            x = np.ascontiguousarray(x, dtype=np.float32)
            if x.ndim != 2 or x.shape[1] != 2:
                raise ValueError(f"expected interleaved stereo [N,2], got {x.shape}")
            total = len(x)
            first, nf = streaming_layout(total, n_fft, hop)
            xd = k.upload(x)
            msq = k.meansq(xd, total, first, n_fft, hop, nf, False, 1.0, False)
            if nf:
                states, rows = k.gate(L.GATE_UPDELAY, msq, sp.m_on, sp.m_off, sp.run_frames, sp.xfade_frames)
            else:
                states, rows = np.zeros(0, np.uint8), np.zeros(0, np.uint16)
            fr = k.frames(xd, total, first, n_fft, hop, nf, win_d, gains_d, rows, 1.0, STREAMING)
            y = k.overlap_add(fr, STREAMING, total, first, n_fft, hop, nf, win_d, sp.post_gain)
            ranges = flush_sample_ranges(nf, total, n_fft, hop)
            peaks = k.limit(y, False, ranges)
            starts = first + hop * np.arange(nf, dtype=np.int64)
            written = ranges[-1][1] if ranges else 0            # the reference writes its chunks only: with a hop that does not
            res.append(dict(out=k.to_host(y, False)[:written],  # divide n_fft the frames can end short of the file
                            chunk_lengths=[b - a for a, b in ranges], chunk_ranges=ranges,
                            chunk_peaks=peaks, meansq=msq, levels=tb.levels_from_meansq(msq), states=np.asarray(states),
                            rows=np.asarray(rows), frame_starts=starts, csv_mask=(starts >= 0) & (starts < total),
                            xfade_frames=sp.xfade_frames, sr=sr, Ton=sp.Ton, Toff=sp.Toff, launches=k.launches))
    finally:
        if hasattr(k, "close_gate"):
            k.close_gate()
    return res


def run_adaptive(xs: Sequence[np.ndarray], sr: int, kernels=None, device: int = 0, fc=1000.0, slope=12.0, c1_low=15.0, c1_high=-15.0,
                 c2_low=-15.0, c2_high=15.0, target_c2=0.5, hyst_db=3.0, min_hold_ms=250.0, xfade_ms=500.0, headroom_margin=2.0,
                 n_fft=tb.N_FFT, hop=tb.HOP) -> List[dict]:
    """adaptive mode for any supported n_fft / hop (src/process_tomatis_adaptive.py:157-373); fields of engine.run_adaptive."""
    n_fft, hop = int(n_fft), int(hop)
    check_sizes(n_fft, hop)
    k = kernels or CudaKernels(device)
    hold, xf = tb.adaptive_frame_counts(sr, min_hold_ms, xfade_ms, hop)
    c1_db, c2_db = tb.tilt_curves_db(sr, n_fft, fc, slope, c1_low, c1_high, c2_low, c2_high)
    win_d, gains_d = k.upload_tables(tb.hann_window(n_fft), tb.gain_rows_adaptive(c1_db, c2_db, xf))
    res = []
    try:
        for x in xs:
            x = np.asarray(x, dtype=np.float32)
            mono = x.ndim == 1 or x.shape[1] == 1
            if mono:
                x = np.stack([x.reshape(-1), np.zeros(x.size, np.float32)], axis=1)
            if x.shape[1] != 2:
                raise NotImplementedError(f"adaptive mode on the GPU takes mono or stereo files, got {x.shape[1]} channels")
            total = len(x)
            if total == 0:
                raise ValueError("zero-size array to reduction operation maximum which has no identity")      # :201
            first, nf = adaptive_layout(total, n_fft, hop)
            xd = k.upload(x)
            in_peak = k.input_peak(xd)
            atten_db, atten_lin, use_f64 = tb.adaptive_attenuation(in_peak, c1_low, c2_high, headroom_margin)
            scale = np.float32(atten_lin)
            msq = k.meansq(xd, total, first, n_fft, hop, nf, use_f64, scale, mono)
            levels = tb.levels_from_meansq(msq)
            # threshold search (:124-154), every gate simulation on the device
            trace = []
            valid = levels > -70
            if valid.any():
                vl = levels[valid]
                T_low, T_high, best_T, best_diff = np.percentile(vl, 5), np.percentile(vl, 95), np.median(vl), 1.0
                for _ in range(30):
                    T_mid = (T_low + T_high) / 2
                    ratio = k.gate(L.GATE_MINHOLD, levels, T_mid + hyst_db / 2, T_mid - hyst_db / 2, hold, xf, True, True) / nf
                    trace.append((float(T_mid), ratio))
                    diff = abs(ratio - target_c2)
                    if diff < best_diff:
                        best_diff, best_T = diff, T_mid
                    if diff < 0.01:
                        break
                    if ratio < target_c2:
                        T_high = T_mid
                    else:
                        T_low = T_mid
            else:
                best_T = np.median(levels)                          # nan for a file shorter than one frame start, like the reference
            if nf:
                states, rows = k.gate(L.GATE_MINHOLD, levels, best_T + hyst_db / 2, best_T - hyst_db / 2, hold, xf, True, False)
            else:
                states, rows = np.zeros(0, np.uint8), np.zeros(0, np.uint16)
            flavour = ADAPTIVE_F64 if use_f64 else ADAPTIVE_F32
            fr = k.frames(xd, total, first, n_fft, hop, nf, win_d, gains_d, rows, 1.0 if use_f64 else scale, flavour)
            restore = 1.0 if use_f64 else float(np.float32(tb.db_to_lin_keep(atten_db)))
            y = k.overlap_add(fr, flavour, total, first, n_fft, hop, nf, win_d, restore)
            peaks = k.limit(y, use_f64, [(0, total)])
            out = k.to_host(y, use_f64)
            res.append(dict(out=(out[:, :1] if mono else out), chunk_lengths=[total], meansq=msq, levels=levels,
                            states=np.asarray(states), rows=np.asarray(rows), times=(np.arange(1, nf + 1) * (hop / sr)),
                            optimal_T=float(best_T), trace=trace, atten_db=float(atten_db),
                            pipeline_dtype="float64" if use_f64 else "float32", min_hold_frames=hold, xfade_frames=xf,
                            output_peak=float(peaks[0]) if len(peaks) else 0.0, sr=sr, input_peak=float(in_peak),
                            launches=k.launches))
    finally:
        if hasattr(k, "close_gate"):
            k.close_gate()
    return res
