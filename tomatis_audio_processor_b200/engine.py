"""Device pipeline of the Tomatis path: thin Python orchestration over the C ABI.

PyTorch is used for device buffers and streams only; every arithmetic step on audio runs in the
hand-written kernels of csrc/tomatis_b200.cu.  The three `run_*` functions mirror the inside of the
reference's `process()` functions (src/process_tomatis.py:160, _xfade.py:55, _adaptive.py:157) on
in-memory arrays and accept a *batch* of tracks (one plan, one launch sequence for all of them).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L
from . import tables as tb


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("tomatis_audio_processor_b200 needs a CUDA device (no CPU fallback)")
    return torch


def _stream_ptr(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    """One per device: window, twiddles and the gain table (tmt_engine)."""

    def __init__(self, device: int = 0, n_fft: int = tb.N_FFT, hop: int = tb.HOP):
        if not fused_size(n_fft, hop):
            raise NotImplementedError(
                f"the fused GPU path implements n_fft/hop = 4096/2048 (the reference defaults) and 2048/1024; got {n_fft}/{hop}")
        self.lib = L.load(n_fft)                         # one build of the library per fused frame size
        torch = _torch()
        self.device = int(device)
        torch.cuda.set_device(self.device)
        torch.cuda.current_stream()                      # make sure the primary context exists
        h = C.c_void_p()
        L.check(self.lib.tmt_engine_create(C.byref(h), self.device, n_fft, hop), "tmt_engine_create")
        self.h = h
        self.n_fft, self.hop = n_fft, hop
        win = np.ascontiguousarray(tb.hann_window(n_fft))
        L.check(self.lib.tmt_engine_set_window(self.h, win.ctypes.data_as(C.c_void_p), n_fft), "set_window")
        self._rows_key = None

    def set_gain_rows(self, rows: np.ndarray, key=None):
        if key is not None and key == self._rows_key:
            return
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        L.check(self.lib.tmt_engine_set_gain_rows(self.h, rows.ctypes.data_as(C.c_void_p), rows.shape[0], rows.shape[1]),
                "set_gain_rows")
        self._rows_key = key

    def close(self):
        if getattr(self, "h", None):
            self.lib.tmt_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_ARR_DTYPES = {L.ARR_MEANSQ_F32: np.float32, L.ARR_MEANSQ_F64: np.float64, L.ARR_GATE_F64: np.float64,
               L.ARR_STATE: np.uint8, L.ARR_ROW: np.uint16, L.ARR_C2_COUNT: np.int32,
               L.ARR_CHUNK_PEAK: np.float32, L.ARR_INPUT_PEAK: np.float32,
               L.ARR_HOPSUM_F32: np.float32, L.ARR_HOPSUM_F64: np.float64, L.ARR_BISECT_T: np.float64,
               L.ARR_BISECT_ITERS: np.int32, L.ARR_BISECT_TRACE_T: np.float64, L.ARR_BISECT_TRACE_C2: np.int32}


class Plan:
    """Geometry + scratch for one batch of tracks / file shards (tmt_plan)."""

    def __init__(self, engine: Engine, framing: int, descs: Sequence[L.TrackDesc], unit_blocks: int = 0):
        self.engine, self.lib = engine, engine.lib
        self.n_tracks = len(descs)
        arr = (L.TrackDesc * max(1, self.n_tracks))(*descs)
        h = C.c_void_p()
        L.check(self.lib.tmt_plan_create(engine.h, C.byref(h), framing, self.n_tracks, arr, unit_blocks), "tmt_plan_create")
        self.h = h
        self.framing = framing
        self.total_frames = self.lib.tmt_plan_total_frames(h)
        self.total_chunks = self.lib.tmt_plan_total_chunks(h)
        self.total_units = self.lib.tmt_plan_total_units(h)
        self.unfusable_chunks = self.lib.tmt_plan_unfusable_chunks(h)
        geo = np.zeros((4, max(1, self.n_tracks)), dtype=np.int32)            # one call instead of four per track
        L.check(self.lib.tmt_plan_geometry(h, *(geo[i].ctypes.data_as(C.c_void_p) for i in range(4))), "tmt_plan_geometry")
        self.track_frames, self.frame_base, self.track_chunks, self.chunk_base = (geo[i, :self.n_tracks].tolist() for i in range(4))
        self._chunk_ranges = None

    # -- geometry
    def chunk_ranges(self, track: int):
        """[(s0, s1)] of the track's limiter chunks, file sample ranges."""
        if self._chunk_ranges is None:
            r = np.zeros((max(1, self.total_chunks), 2), dtype=np.int64)
            L.check(self.lib.tmt_plan_chunk_ranges(self.h, r.ctypes.data_as(C.c_void_p)), "tmt_plan_chunk_ranges")
            self._chunk_ranges = r.tolist()
        cb = self.chunk_base[track]
        return [tuple(v) for v in self._chunk_ranges[cb:cb + self.track_chunks[track]]]

    # -- arrays
    def _count(self, which):
        if which in (L.ARR_C2_COUNT, L.ARR_INPUT_PEAK, L.ARR_BISECT_T, L.ARR_BISECT_ITERS):
            return self.n_tracks
        if which in (L.ARR_BISECT_TRACE_T, L.ARR_BISECT_TRACE_C2):
            return self.n_tracks * L.BISECT_MAX_ITER
        if which == L.ARR_CHUNK_PEAK:
            return self.total_chunks
        if which in (L.ARR_HOPSUM_F32, L.ARR_HOPSUM_F64):
            return self.total_frames + self.n_tracks
        return self.total_frames

    def read(self, which: int, offset: int = 0, count: Optional[int] = None) -> np.ndarray:
        torch = _torch()
        if count is None:
            count = self._count(which) - offset
        out = np.empty(count, dtype=_ARR_DTYPES[which])
        if count:
            L.check(self.lib.tmt_plan_read(self.h, which, offset, count, out.ctypes.data_as(C.c_void_p), 0,
                                           _stream_ptr(torch)), "tmt_plan_read")
        return out

    def read_many(self, *which) -> list:
        """Whole arrays, one host wait for all of them (tmt_plan_read_many)."""
        torch = _torch()
        outs = [np.empty(max(1, self._count(w)), dtype=_ARR_DTYPES[w]) for w in which]
        ids = (C.c_int32 * len(which))(*which)
        ptrs = (C.c_void_p * len(which))(*(o.ctypes.data for o in outs))
        L.check(self.lib.tmt_plan_read_many(self.h, len(which), ids, ptrs, _stream_ptr(torch)), "tmt_plan_read_many")
        return [o[:self._count(w)] for o, w in zip(outs, which)]

    def write(self, which: int, data: np.ndarray, offset: int = 0):
        torch = _torch()
        data = np.ascontiguousarray(data, dtype=_ARR_DTYPES[which])
        if data.size:
            L.check(self.lib.tmt_plan_write(self.h, which, offset, data.size, data.ctypes.data_as(C.c_void_p), 0,
                                            _stream_ptr(torch)), "tmt_plan_write")

    def write_device(self, which: int, dev_ptr: int, count: int, offset: int = 0):
        torch = _torch()
        L.check(self.lib.tmt_plan_write(self.h, which, offset, count, C.c_void_p(dev_ptr), 1, _stream_ptr(torch)))

    def read_device(self, which: int, dev_ptr: int, count: int, offset: int = 0):
        torch = _torch()
        L.check(self.lib.tmt_plan_read(self.h, which, offset, count, C.c_void_p(dev_ptr), 1, _stream_ptr(torch)))

    # -- kernels
    def input_peaks(self):
        L.check(self.lib.tmt_plan_input_peaks(self.h, _stream_ptr(_torch())), "tmt_plan_input_peaks")

    def set_level_ranges(self, track: int, hb_lo: int, hb_hi: int, f_lo: int, f_hi: int):
        L.check(self.lib.tmt_plan_set_level_ranges(self.h, track, hb_lo, hb_hi, f_lo, f_hi), "tmt_plan_set_level_ranges")

    def levels(self, use_f64: bool = False, in_scale: Optional[np.ndarray] = None, mono: bool = False, part: str = "all",
               channel: Optional[str] = None):
        """part: "all" | "hopsums" (hop-block sums only) | "meansq" (mean squares from the sums already in the plan);
        channel: None (the mode's stereo / mono formula) | "left" | "right" (np.mean(x*x) of that channel alone) |
        "power_eps" (the calibration front end's power-average mono, epsilon inside the root)"""
        ptr = None
        if in_scale is not None:
            in_scale = np.ascontiguousarray(in_scale, dtype=np.float32)
            assert in_scale.size == self.n_tracks
            ptr = in_scale.ctypes.data_as(C.c_void_p)
        flags = (int(bool(use_f64)) | (L.LEVELS_MONO if mono else 0)
                 | {"all": 0, "hopsums": L.LEVELS_HOPSUM_ONLY, "meansq": L.LEVELS_MEANSQ_ONLY}[part]
                 | {None: 0, "left": L.LEVELS_LEFT, "right": L.LEVELS_RIGHT, "power_eps": L.LEVELS_POWER_EPS}[channel])
        L.check(self.lib.tmt_plan_levels(self.h, flags, ptr, _stream_ptr(_torch())),
                "tmt_plan_levels")

    def levels_multichannel(self, x_dev, channels: int, use_f64: bool = False, in_scale: Optional[np.ndarray] = None):
        """The level all channel pairs of one file share (tmt_plan_levels_multichannel): x_dev = the interleaved file [N, channels]."""
        ptr = None
        if in_scale is not None:
            in_scale = np.ascontiguousarray(in_scale, dtype=np.float32)
            assert in_scale.size == self.n_tracks
            ptr = in_scale.ctypes.data_as(C.c_void_p)
        L.check(self.lib.tmt_plan_levels_multichannel(self.h, int(bool(use_f64)), ptr, C.c_void_p(x_dev.data_ptr()), int(channels),
                                                      _stream_ptr(_torch())), "tmt_plan_levels_multichannel")

    def gate(self, automaton: int, gate_input: int, on, off, param: int, xfade_frames: int,
             alpha_init_to_target: bool = False, count_only: bool = False):
        """on / off None: keep the thresholds already on the device (left there by bisect())."""
        if on is None and off is None:
            pon = poff = None
        else:
            on = np.ascontiguousarray(np.broadcast_to(np.asarray(on, dtype=np.float64), (max(1, self.n_tracks),)))
            off = np.ascontiguousarray(np.broadcast_to(np.asarray(off, dtype=np.float64), (max(1, self.n_tracks),)))
            pon, poff = on.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p)
        L.check(self.lib.tmt_plan_gate(self.h, automaton, gate_input, pon, poff, int(param), int(xfade_frames),
                                       int(alpha_init_to_target), int(count_only), _stream_ptr(_torch())), "tmt_plan_gate")

    def can_bisect(self) -> bool:
        """The in-kernel threshold search covers tracks the gate scan handles in one segment."""
        return max(self.track_frames, default=0) <= 16384

    def bisect(self, t_low, t_high, start, active, hyst_db: float, target_c2: float, hold: int, max_iter: int = 30, final_gate=None):
        """find_optimal_threshold for every track in one launch (tmt_plan_bisect); results in ARR_BISECT_*.
        final_gate = (xfade_frames, alpha_init_to_target): also run the final gate with the thresholds found
        (tmt_plan_bisect_gate; states, rows and C2 counts are final afterwards)."""
        n = max(1, self.n_tracks)
        arrs = [np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,))) for a in (t_low, t_high, start)]
        act = np.ascontiguousarray(np.broadcast_to(np.asarray(active, dtype=np.int32), (n,)))
        if final_gate is not None:
            rc = self.lib.tmt_plan_bisect_gate(self.h, *(a.ctypes.data_as(C.c_void_p) for a in arrs), act.ctypes.data_as(C.c_void_p),
                                               float(hyst_db), float(target_c2), int(hold), int(max_iter), int(final_gate[0]),
                                               int(bool(final_gate[1])), _stream_ptr(_torch()))
        else:
            rc = self.lib.tmt_plan_bisect(self.h, *(a.ctypes.data_as(C.c_void_p) for a in arrs), act.ctypes.data_as(C.c_void_p),
                                          float(hyst_db), float(target_c2), int(hold), int(max_iter), _stream_ptr(_torch()))
        if rc == L.ERR_UNSUPPORTED:          # long tracks / forced multi-segment scan: the caller drives the search from the host
            return False
        L.check(rc, "tmt_plan_bisect")
        return True

    def stft(self, post_gain: float = 1.0, skip_edges: bool = True):
        L.check(self.lib.tmt_plan_stft(self.h, float(post_gain), int(skip_edges), _stream_ptr(_torch())), "tmt_plan_stft")

    def clear_peaks(self):
        L.check(self.lib.tmt_plan_clear_peaks(self.h, _stream_ptr(_torch())), "tmt_plan_clear_peaks")

    def stft_limited(self, post_gain: float = 1.0, limit: float = tb.PEAK_LIMIT):
        """Fused STFT/OLA + per-chunk limiter; call clear_peaks() and edge_frames() first."""
        L.check(self.lib.tmt_plan_stft_limited(self.h, float(post_gain), float(np.float32(limit)), _stream_ptr(_torch())),
                "tmt_plan_stft_limited")

    def edge_frames(self, post_gain: float = 1.0, in_scale=None, out_scale=None, pipeline_f64: bool = False):
        pi = po = None
        if in_scale is not None:
            in_scale = np.ascontiguousarray(in_scale, dtype=np.float32)
            pi = in_scale.ctypes.data_as(C.c_void_p)
        if out_scale is not None:
            out_scale = np.ascontiguousarray(out_scale, dtype=np.float32)
            po = out_scale.ctypes.data_as(C.c_void_p)
        L.check(self.lib.tmt_plan_edge_frames(self.h, float(post_gain), pi, po, int(pipeline_f64), _stream_ptr(_torch())),
                "tmt_plan_edge_frames")

    def stft_with_edges(self, post_gain: float = 1.0, in_scale=None, out_scale=None, pipeline_f64: bool = False):
        """stft() + edge_frames() with the fp64 edge kernel running beside the STFT kernel (tmt_plan_stft_with_edges)."""
        pi = po = None
        if in_scale is not None:
            in_scale = np.ascontiguousarray(in_scale, dtype=np.float32)
            pi = in_scale.ctypes.data_as(C.c_void_p)
        if out_scale is not None:
            out_scale = np.ascontiguousarray(out_scale, dtype=np.float32)
            po = out_scale.ctypes.data_as(C.c_void_p)
        L.check(self.lib.tmt_plan_stft_with_edges(self.h, float(post_gain), pi, po, int(pipeline_f64), _stream_ptr(_torch())),
                "tmt_plan_stft_with_edges")

    def limiter(self, limit: float = tb.PEAK_LIMIT):
        L.check(self.lib.tmt_plan_limiter(self.h, float(np.float32(limit)), _stream_ptr(_torch())), "tmt_plan_limiter")

    def run_streaming(self, m_on, m_off, run_frames, xfade_frames, post_gain=1.0, limit=tb.PEAK_LIMIT):
        L.check(self.lib.tmt_plan_run_streaming(self.h, float(m_on), float(m_off), int(run_frames), int(xfade_frames),
                                                float(post_gain), float(np.float32(limit)), _stream_ptr(_torch())),
                "tmt_plan_run_streaming")

    def pcm_levels(self, pcm_dev, fmt: int):
        """Integer PCM of all tracks (device tensor [T, N, 2] int16 or [T, N, 6] uint8) -> the plan's float input buffers and
        the hop-block sums, one pass (tmt_plan_pcm_levels)."""
        stride = pcm_dev.stride(0) * pcm_dev.element_size() if pcm_dev.dim() == 3 else pcm_dev.numel() * pcm_dev.element_size()
        L.check(self.lib.tmt_plan_pcm_levels(self.h, C.c_void_p(pcm_dev.data_ptr()), int(stride), int(fmt), _stream_ptr(_torch())),
                "tmt_plan_pcm_levels")

    def run_streaming_pcm(self, pcm_dev, fmt: int, m_on, m_off, run_frames, xfade_frames, post_gain=1.0, limit=tb.PEAK_LIMIT):
        """run_streaming with integer PCM input: conversion and hop-block sums in one pass (tmt_plan_run_streaming_pcm)."""
        stride = pcm_dev.stride(0) * pcm_dev.element_size() if pcm_dev.dim() == 3 else pcm_dev.numel() * pcm_dev.element_size()
        L.check(self.lib.tmt_plan_run_streaming_pcm(self.h, C.c_void_p(pcm_dev.data_ptr()), int(stride), int(fmt), float(m_on), float(m_off),
                                                    int(run_frames), int(xfade_frames), float(post_gain), float(np.float32(limit)),
                                                    _stream_ptr(_torch())), "tmt_plan_run_streaming_pcm")

    def launch_count(self) -> int:
        return int(self.lib.tmt_plan_launch_count(self.h))

    def set_buffers(self, track: int, in_ptr: int, out_ptr: int):
        L.check(self.lib.tmt_plan_set_buffers(self.h, track, C.c_void_p(in_ptr), C.c_void_p(out_ptr)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.tmt_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pcm_to_float(pcm_dev, fmt: int, out_dev):
    """Device tensors: integer PCM (int16 values, or uint8 bytes of packed PCM_24) -> float32, on the current stream."""
    torch = _torch()
    n = out_dev.numel()
    L.check(L.load().tmt_pcm_to_float(C.c_void_p(pcm_dev.data_ptr()), int(fmt), n, C.c_void_p(out_dev.data_ptr()),
                                      _stream_ptr(torch)), "tmt_pcm_to_float")


def float_to_pcm24(x_dev, out_bytes_dev):
    """Device tensors: float32 -> packed PCM_24 bytes (3 per value), on the current stream."""
    torch = _torch()
    L.check(L.load().tmt_float_to_pcm(C.c_void_p(x_dev.data_ptr()), L.PCM_S24, x_dev.numel(),
                                      C.c_void_p(out_bytes_dev.data_ptr()), _stream_ptr(torch)), "tmt_float_to_pcm")


def whole_track_desc(x_dev, y_dev, total: Optional[int] = None) -> L.TrackDesc:
    n = int(x_dev.shape[0]) if total is None else int(total)
    return L.TrackDesc(x_dev.data_ptr(), y_dev.data_ptr(), n, 0, n, 0, n, 0, -1)


_engines = {}


def fused_size(n_fft, hop) -> bool:
    """Frame sizes the fused kernels serve: the reference's defaults and its documented faster setting 2048 / 1024
    (TMT_FUSED_2048=0 sends the latter to the general-size path, for comparisons)."""
    import os
    from .build import FUSED_SIZES
    if FUSED_SIZES.get(n_fft) != hop:
        return False
    return n_fft == tb.N_FFT or os.environ.get("TMT_FUSED_2048", "1") != "0"


def get_engine(device: int = 0, n_fft: int = tb.N_FFT, hop: int = tb.HOP) -> Engine:
    key = device if n_fft == tb.N_FFT else (device, n_fft)
    if key not in _engines:
        _engines[key] = Engine(device, n_fft, hop)
    return _engines[key]


# ------------------------------------------------------------------------------------------------
@dataclass
class StreamingParams:
    """Scalars of standard / xfade mode after the reference's host-side preamble
    (src/process_tomatis.py:256-285, _xfade.py:128-155)."""
    sr: int
    Ton: float
    Toff: float
    run_frames: int
    xfade_frames: int
    rows: np.ndarray
    rows_key: tuple
    post_gain: float
    m_on: float = field(default=0.0)
    m_off: float = field(default=0.0)


_params_cache = {}


def streaming_params(mode: str, sr: int, **kw) -> StreamingParams:
    """Host-side preamble of standard / xfade, cached per parameter set: the tilt tables cost more host time than a whole
    single-track call costs on the GPU (0.2 - 0.4 ms of NumPy against 0.25 ms of kernels for a 60 s file)."""
    try:
        key = (mode, sr, tuple(sorted(kw.items())))
        hit = _params_cache.get(key)
    except TypeError:                     # unhashable argument (array-valued): no caching
        key, hit = None, None
    if hit is None:
        hit = _streaming_params(mode, sr, **kw)
        if key is not None:
            if len(_params_cache) > 64:
                _params_cache.clear()
            _params_cache[key] = hit
    return hit


def _streaming_params(mode: str, sr: int, *, gate_ui=50, gate_mode="log_percent", dynamic_range=80.0, gate_scale=1.0,
                      gate_offset=-100, hysteresis_db=3.0, fc=1000.0, slope=12.0, c1_low=+15.0, c1_high=-15.0,
                      c2_low=-15.0, c2_high=+15.0, up_delay_ms=250.0, xfade_ms=0.0, n_fft=tb.N_FFT, hop=tb.HOP,
                      output_gain_db=0.0) -> StreamingParams:
    if mode == "standard" and gate_mode == "log_percent":
        T = tb.gate_threshold_log_percent(gate_ui, dynamic_range)
    else:
        T = tb.gate_threshold_linear(gate_ui, gate_scale, gate_offset)
    Ton, Toff = tb.hysteresis_pair(T, hysteresis_db)
    g1_db, g2_db = tb.tilt_curves_db(sr, n_fft, fc, slope, c1_low, c1_high, c2_low, c2_high)
    if mode == "xfade":
        xf = tb.xfade_frame_count(sr, xfade_ms, hop)
        rows = tb.gain_rows_xfade(g1_db, g2_db, xf) if xf > 0 else tb.gain_rows_standard(g1_db, g2_db)
    else:
        xf = 0
        rows = tb.gain_rows_standard(g1_db, g2_db)
    key = (mode, sr, fc, slope, c1_low, c1_high, c2_low, c2_high, xf, n_fft)
    post_gain = 1.0 if output_gain_db == 0.0 else float(np.float32(10.0 ** (output_gain_db / 20.0)))
    return StreamingParams(sr=sr, Ton=Ton, Toff=Toff, run_frames=tb.updelay_run_frames(sr, up_delay_ms, hop),
                           xfade_frames=xf, rows=rows, rows_key=key, post_gain=post_gain,
                           m_on=tb.meansq_threshold_on(Ton), m_off=tb.meansq_threshold_off(Toff))


def _host_array(x) -> np.ndarray:
    return x if isinstance(x, np.ndarray) else x.cpu().numpy()


def _to_device(torch, xs, device):
    out = []
    for x in xs:
        if isinstance(x, np.ndarray):
            x = np.ascontiguousarray(x, dtype=np.float32)
            if x.ndim != 2 or x.shape[1] != 2:
                raise ValueError(f"expected interleaved stereo [N,2], got {x.shape}")
            t = torch.from_numpy(x).to(f"cuda:{device}", non_blocking=False)
        else:
            t = x
            if t.dtype != torch.float32 or t.dim() != 2 or t.shape[1] != 2 or not t.is_contiguous():
                raise ValueError("device tracks must be contiguous float32 [N,2]")
        out.append(t)
    return out


def run_streaming(mode: str, xs: Sequence, sr: int, device: int = 0, want_host: bool = True, outs=None,
                  unit_blocks: int = 0, **params) -> List[dict]:
    """standard / xfade on a batch of tracks.  Returns one dict per track: out (numpy [N,2] float32 if
    want_host else the device tensor), chunk_lengths, meansq, levels, states, rows, frame geometry."""
    torch = _torch()
    n_fft, hop = params.get("n_fft", tb.N_FFT), params.get("hop", tb.HOP)
    eng = get_engine(device, n_fft, hop) if fused_size(n_fft, hop) else get_engine(device)
    if n_fft != eng.n_fft or hop != eng.hop:
        from . import generic
        if generic.enabled():                                         # general-size path (csrc/generic.cuh)
            return generic.run_streaming(mode, [_host_array(x) for x in xs], sr, device=device, **params)
        raise NotImplementedError(f"GPU path implements n_fft={eng.n_fft}, hop={eng.hop}; got {n_fft}/{hop}")
    sp = streaming_params(mode, sr, **params)
    eng.set_gain_rows(sp.rows, key=sp.rows_key)
    xd = _to_device(torch, xs, device)
    yd = outs if outs is not None else [torch.empty_like(x) for x in xd]
    plan = Plan(eng, L.FRAMING_STREAMING, [whole_track_desc(x, y) for x, y in zip(xd, yd)], unit_blocks)
    try:
        plan.run_streaming(sp.m_on, sp.m_off, sp.run_frames, sp.xfade_frames, sp.post_gain)
        msq, states, rows, peaks = plan.read_many(L.ARR_MEANSQ_F32, L.ARR_STATE, L.ARR_ROW, L.ARR_CHUNK_PEAK)
        levels_all = tb.levels_from_meansq(msq)                        # one vectorised pass for the whole batch
        launches = plan.launch_count()
        res = []
        for t in range(plan.n_tracks):
            fb, nf = plan.frame_base[t], plan.track_frames[t]
            total = int(xd[t].shape[0])
            ranges = plan.chunk_ranges(t)
            cb = plan.chunk_base[t]
            starts = -(n_fft // 2) + hop * np.arange(nf, dtype=np.int64)
            res.append(dict(
                out=(yd[t].cpu().numpy() if want_host else yd[t]),
                chunk_lengths=[int(b - a) for (a, b) in ranges if b > a],
                chunk_ranges=ranges, chunk_peaks=peaks[cb:cb + len(ranges)].copy(),
                meansq=msq[fb:fb + nf].copy(), levels=levels_all[fb:fb + nf].copy(), states=states[fb:fb + nf].copy(),
                rows=rows[fb:fb + nf].copy(), frame_starts=starts, csv_mask=(starts >= 0) & (starts < total),
                xfade_frames=sp.xfade_frames, sr=sr, Ton=sp.Ton, Toff=sp.Toff, launches=launches))
        return res
    finally:
        plan.close()


def _percentile_sorted(s: np.ndarray, q: float):
    """np.percentile(s, q) (method "linear") of an already sorted float64 array, bit for bit: NumPy's virtual index
    (n - 1) * (q / 100), its two neighbours and its two-sided lerp (a + (b - a) * g, or b - (b - a) * (1 - g) once g >= 0.5)."""
    n = len(s)
    vi = (n - 1) * np.true_divide(q, 100)
    lo = int(np.floor(vi))
    if vi >= n - 1:
        lo = hi = n - 1
    else:
        hi = lo + 1
    g = vi - lo
    a, b = s[lo], s[hi]
    d = b - a
    return b - d * (1 - g) if g >= 0.5 else a + d * g


def _percentile_thresholds(levels, valid):
    """(p5, p95, median) of the valid levels (src/process_tomatis_adaptive.py:131-135).  One sort and three reads instead of
    np.percentile + np.median (0.6 ms of host time per 10-minute track against 0.5 ms of kernels for the whole file); the
    values are NumPy's own, bit for bit (tests/test_tables_cpu.py)."""
    vl = levels[valid]
    n = len(vl)
    if n == 0:
        return None
    s = np.sort(vl)
    med = s[n // 2] if n % 2 else (s[n // 2 - 1] + s[n // 2]) / 2.0
    return _percentile_sorted(s, 5), _percentile_sorted(s, 95), med


def run_adaptive(xs: Sequence, sr: int, device: int = 0, want_host: bool = True, outs=None, unit_blocks: int = 0,
                 fc=1000.0, slope=12.0, c1_low=15.0, c1_high=-15.0, c2_low=-15.0, c2_high=15.0, target_c2=0.5,
                 hyst_db=3.0, min_hold_ms=250.0, xfade_ms=500.0, headroom_margin=2.0, n_fft=tb.N_FFT, hop=tb.HOP,
                 _linked=None) -> List[dict]:
    """adaptive mode on a batch of tracks (src/process_tomatis_adaptive.py:157-373).

    Host: scalars, percentiles and the bisection bookkeeping.  Device: input peaks, levels, every gate
    simulation of the bisection (count-only scans), the final gate + crossfade counter, STFT, limiter.

    A host array with more than two channels is one file whose channel pairs share the gate (run_adaptive_multichannel);
    `_linked` = (interleaved device tensor, channels) is that function's hook: the tracks are the channel pairs of one file,
    so input peak, level and output peak are taken over all of them."""
    torch = _torch()
    eng = get_engine(device, n_fft, hop) if fused_size(n_fft, hop) else get_engine(device)
    if _linked is None and any(isinstance(x, np.ndarray) and x.ndim == 2 and x.shape[1] > 2 for x in xs):
        if not fused_size(n_fft, hop):
            raise NotImplementedError(f"files with more than two channels need n_fft/hop = 4096/2048 or 2048/1024; got {n_fft}/{hop}")
        kw = dict(device=device, want_host=want_host, unit_blocks=unit_blocks, fc=fc, slope=slope, c1_low=c1_low, c1_high=c1_high,
                  c2_low=c2_low, c2_high=c2_high, target_c2=target_c2, hyst_db=hyst_db, min_hold_ms=min_hold_ms, xfade_ms=xfade_ms,
                  headroom_margin=headroom_margin, n_fft=n_fft, hop=hop)
        return [run_adaptive_multichannel(x, sr, **kw) if (isinstance(x, np.ndarray) and x.ndim == 2 and x.shape[1] > 2)
                else run_adaptive([x], sr, **kw)[0] for x in xs]
    if n_fft != eng.n_fft or hop != eng.hop:
        from . import generic
        if generic.enabled():                                         # general-size path (csrc/generic.cuh)
            return generic.run_adaptive([_host_array(x) for x in xs], sr, device=device, fc=fc, slope=slope, c1_low=c1_low,
                                        c1_high=c1_high, c2_low=c2_low, c2_high=c2_high, target_c2=target_c2, hyst_db=hyst_db,
                                        min_hold_ms=min_hold_ms, xfade_ms=xfade_ms, headroom_margin=headroom_margin,
                                        n_fft=n_fft, hop=hop)
        raise NotImplementedError(f"GPU path implements n_fft={eng.n_fft}, hop={eng.hop}; got {n_fft}/{hop}")
    hold, xf = tb.adaptive_frame_counts(sr, min_hold_ms, xfade_ms, hop)
    rows_key = ("adaptive", sr, fc, slope, c1_low, c1_high, c2_low, c2_high, xf, n_fft)
    if eng._rows_key != rows_key:                 # the table build is host work worth skipping (see streaming_params)
        c1_db, c2_db = tb.tilt_curves_db(sr, n_fft, fc, slope, c1_low, c1_high, c2_low, c2_high)
        eng.set_gain_rows(tb.gain_rows_adaptive(c1_db, c2_db, xf), key=rows_key)
    # single-channel files (accepted by the reference, _adaptive.py:180-181) ride in the L lane with R = 0
    mono_in = [isinstance(x, np.ndarray) and (x.ndim == 1 or x.shape[1] == 1) for x in xs]
    if any(mono_in) and not all(mono_in):
        raise ValueError("mono and stereo tracks cannot share one adaptive batch")
    mono = bool(mono_in) and all(mono_in)
    if mono:
        xs = [np.stack([np.asarray(x, np.float32).reshape(-1), np.zeros(len(x), np.float32)], axis=1) for x in xs]
    xd = _to_device(torch, xs, device)
    yd = outs if outs is not None else [torch.empty_like(x) for x in xd]
    results: List[Optional[dict]] = [None] * len(xd)

    # pass 0: input peaks decide pre-attenuation and the float32/float64 branch per track
    plan0 = Plan(eng, L.FRAMING_WHOLEFILE, [whole_track_desc(x, y) for x, y in zip(xd, yd)], unit_blocks)
    try:
        plan0.input_peaks()
        in_peaks = plan0.read(L.ARR_INPUT_PEAK)
        if _linked is not None:                       # np.max(np.abs(x)) runs over every channel of the file (:201)
            in_peaks = np.full_like(in_peaks, in_peaks.max() if len(in_peaks) else 0.0)
    except Exception:
        plan0.close()
        raise
    branch = [tb.adaptive_attenuation(pk, c1_low, c2_high, headroom_margin) for pk in in_peaks]
    single_branch = len({b[2] for b in branch}) <= 1          # the usual case: the first plan serves the whole call
    if not single_branch:
        plan0.close()

    for use_f64 in (False, True):
        idx = [i for i, b in enumerate(branch) if b[2] == use_f64]
        if not idx:
            continue
        plan = plan0 if single_branch else Plan(eng, L.FRAMING_WHOLEFILE, [whole_track_desc(xd[i], yd[i]) for i in idx], unit_blocks)
        try:
            scale = np.array([np.float32(branch[i][1]) for i in idx], dtype=np.float32)
            if _linked is not None:
                plan.levels_multichannel(_linked[0], _linked[1], use_f64=use_f64, in_scale=scale)
            else:
                plan.levels(use_f64=use_f64, in_scale=scale, mono=mono)
            msq = plan.read(L.ARR_MEANSQ_F64 if use_f64 else L.ARR_MEANSQ_F32)
            levels_all = tb.levels_from_meansq(msq)
            plan.write(L.ARR_GATE_F64, levels_all)
            nt = len(idx)
            lv = [levels_all[plan.frame_base[t]:plan.frame_base[t] + plan.track_frames[t]] for t in range(nt)]
            # threshold bisection, all tracks in lock step (src/process_tomatis_adaptive.py:124-154)
            T_low = np.zeros(nt); T_high = np.zeros(nt); best_T = np.zeros(nt); best_diff = np.ones(nt)
            active = np.ones(nt, dtype=bool)
            for t in range(nt):
                if plan.track_frames[t] == 0:
                    active[t] = False
                    continue
                valid = lv[t] > -70
                pt = _percentile_thresholds(lv[t], valid)
                if pt is None:
                    best_T[t] = np.median(lv[t]); active[t] = False
                else:
                    T_low[t], T_high[t], best_T[t] = pt
            traces = [[] for _ in range(nt)]
            # the whole search in one launch when the tracks fit the single-segment scan; thresholds stay on the device for the
            # final gate, results come back at the end
            in_kernel = plan.can_bisect() and plan.bisect(T_low, T_high, best_T, active, hyst_db, target_c2, hold, 30,
                                                          final_gate=(xf, True))
            if not in_kernel:
                for _ in range(30):
                    if not active.any():
                        break
                    T_mid = (T_low + T_high) / 2
                    plan.gate(L.GATE_MINHOLD, L.ARR_GATE_F64, T_mid + hyst_db / 2, T_mid - hyst_db / 2, hold, xf,
                              alpha_init_to_target=True, count_only=True)
                    c2 = plan.read(L.ARR_C2_COUNT)
                    for t in range(nt):
                        if not active[t]:
                            continue
                        ratio = int(c2[t]) / plan.track_frames[t]
                        traces[t].append((float(T_mid[t]), ratio))
                        diff = abs(ratio - target_c2)
                        if diff < best_diff[t]:
                            best_diff[t] = diff; best_T[t] = T_mid[t]
                        if diff < 0.01:
                            active[t] = False
                            continue
                        if ratio < target_c2:
                            T_high[t] = T_mid[t]
                        else:
                            T_low[t] = T_mid[t]
                plan.gate(L.GATE_MINHOLD, L.ARR_GATE_F64, best_T + hyst_db / 2, best_T - hyst_db / 2, hold, xf,
                          alpha_init_to_target=True, count_only=False)
            if use_f64:
                plan.stft_with_edges(1.0, None, None, pipeline_f64=True)
            else:      # restore_lin = db_to_lin(atten_db), float32 (src/process_tomatis_adaptive.py:335-337)
                restore = np.array([np.float32(tb.db_to_lin_keep(branch[i][0])) for i in idx], dtype=np.float32)
                plan.stft_with_edges(1.0, scale, restore, pipeline_f64=False)
            if _linked is not None:                   # output_peak = np.max(np.abs(y)) over every channel (:340): one scale for all pairs
                pk = plan.read(L.ARR_CHUNK_PEAK)
                plan.write(L.ARR_CHUNK_PEAK, np.full_like(pk, pk.max() if len(pk) else 0.0))
            plan.limiter()
            if in_kernel:
                states, rows, peaks, best_T, iters, tr_T, tr_c = plan.read_many(
                    L.ARR_STATE, L.ARR_ROW, L.ARR_CHUNK_PEAK, L.ARR_BISECT_T, L.ARR_BISECT_ITERS, L.ARR_BISECT_TRACE_T, L.ARR_BISECT_TRACE_C2)
                tr_T, tr_c = tr_T.reshape(nt, L.BISECT_MAX_ITER), tr_c.reshape(nt, L.BISECT_MAX_ITER)
                traces = [[(float(tr_T[t, k]), int(tr_c[t, k]) / plan.track_frames[t]) for k in range(int(iters[t]))] for t in range(nt)]
            else:
                states, rows, peaks = plan.read_many(L.ARR_STATE, L.ARR_ROW, L.ARR_CHUNK_PEAK)
            for t, i in enumerate(idx):
                fb, nf = plan.frame_base[t], plan.track_frames[t]
                results[i] = dict(
                    out=((yd[i].cpu().numpy()[:, :1] if mono else yd[i].cpu().numpy()) if want_host else yd[i]),
                    chunk_lengths=[int(xd[i].shape[0])],
                    meansq=msq[fb:fb + nf].copy(), levels=lv[t].copy(), states=states[fb:fb + nf].copy(),
                    rows=rows[fb:fb + nf].copy(), times=(np.arange(1, nf + 1) * (hop / sr)),
                    optimal_T=float(best_T[t]), trace=traces[t], atten_db=float(branch[i][0]),
                    pipeline_dtype="float64" if use_f64 else "float32", min_hold_frames=hold, xfade_frames=xf,
                    output_peak=float(peaks[plan.chunk_base[t]]) if plan.track_chunks[t] else 0.0, sr=sr,
                    input_peak=float(in_peaks[i]), launches=plan.launch_count())
        finally:
            plan.close()
    return results


def run_adaptive_multichannel(x: np.ndarray, sr: int, device: int = 0, want_host: bool = True, **params) -> dict:
    """adaptive mode on one file with more than two channels (the reference's `for c in range(ch)`,
    src/process_tomatis_adaptive.py:307-313): the file is cut into channel pairs on the device, every pair is a track of one
    plan of the stereo path, and what the reference computes over all channels -- input peak (:201), frame level (:74), output
    peak (:340) -- is shared by the pairs, so that there is one gate, one pre-attenuation and one limiter scale."""
    torch = _torch()
    x = np.ascontiguousarray(x, dtype=np.float32)
    total, ch = x.shape
    if ch > 128:
        raise NotImplementedError(f"up to 128 channels, got {ch}")
    npairs = (ch + 1) // 2
    lib = L.load()
    xd = torch.from_numpy(x).to(f"cuda:{device}")
    pairs = torch.empty((npairs, max(1, total), 2), dtype=torch.float32, device=xd.device)
    L.check(lib.tmt_channels_split(C.c_void_p(xd.data_ptr()), total, ch, C.c_void_p(pairs.data_ptr()), _stream_ptr(torch)), "tmt_channels_split")
    outs = torch.empty_like(pairs)
    res = run_adaptive([pairs[p, :total] for p in range(npairs)], sr, device=device, want_host=False,
                       outs=[outs[p, :total] for p in range(npairs)], _linked=(xd, ch), **params)
    yd = torch.empty_like(xd)
    L.check(lib.tmt_channels_merge(C.c_void_p(outs.data_ptr()), total, ch, C.c_void_p(yd.data_ptr()), _stream_ptr(torch)), "tmt_channels_merge")
    r = dict(res[0])
    r["out"] = yd.cpu().numpy() if want_host else yd
    r["output_peak"] = max(float(q["output_peak"]) for q in res)
    r["launches"] = res[-1]["launches"] + 2
    r["channels"] = ch
    return r


# ------------------------------------------------------------------------------ per-channel state analysis (N3)
def run_channel_states(xs: Sequence, sr: int, device: int = 0, target_c2=0.5, hyst_db=3.0, min_hold_ms=250.0,
                       n_fft=tb.N_FFT, hop=tb.HOP) -> List[dict]:
    """Left and right channel analysed on their own: level per frame, threshold search, min-hold gate
    (src/analyze_stereo_state.py:79-128).  Device: both level passes and every gate simulation of the two searches, all
    tracks in lock step; host: percentiles and the search bookkeeping.  No audio is written (analysis-only plan)."""
    torch = _torch()
    if not fused_size(n_fft, hop):
        raise NotImplementedError(f"GPU path implements n_fft/hop = 4096/2048 and 2048/1024 here; got {n_fft}/{hop}")
    eng = get_engine(device, n_fft, hop)
    frame_ms = hop / sr * 1000                                        # :89-90
    hold = int(np.ceil(min_hold_ms / frame_ms))
    xd = _to_device(torch, xs, device)
    descs = [L.TrackDesc(x.data_ptr(), None, int(x.shape[0]), 0, int(x.shape[0]), 0, 0, 0, -1) for x in xd]
    plan = Plan(eng, L.FRAMING_WHOLEFILE, descs)
    try:
        nt = plan.n_tracks
        res = [dict(min_hold_frames=hold, sr=sr, times=np.arange(plan.track_frames[t]) * hop / sr)
               for t in range(nt)]                                   # frame i covers [i*hop, i*hop + n_fft): orig_start / sr (:114)
        for name in ("left", "right"):
            plan.levels(channel=name)
            levels_all = tb.levels_from_meansq(plan.read(L.ARR_MEANSQ_F32))      # float32 chain of rms_dbfs (:16-19)
            plan.write(L.ARR_GATE_F64, levels_all)
            lv = [levels_all[plan.frame_base[t]:plan.frame_base[t] + plan.track_frames[t]] for t in range(nt)]
            # find_optimal_threshold (:52-76): returns the midpoint that hits the target, else the last one tried
            T_low = np.zeros(nt); T_high = np.zeros(nt); T = np.zeros(nt)
            active = np.ones(nt, dtype=bool)
            for t in range(nt):
                pt = _percentile_thresholds(lv[t], lv[t] > -70)
                if pt is None:
                    with np.errstate(invalid="ignore"):
                        T[t] = np.median(lv[t]) if len(lv[t]) else np.nan
                    active[t] = False
                else:
                    T_low[t], T_high[t], T[t] = pt
            for _ in range(30):
                if not active.any():
                    break
                T_mid = np.where(active, (T_low + T_high) / 2, T)
                T_mid = np.nan_to_num(T_mid)                          # frameless tracks carry NaN (np.median of nothing)
                plan.gate(L.GATE_MINHOLD, L.ARR_GATE_F64, T_mid + hyst_db / 2, T_mid - hyst_db / 2, hold, 0, count_only=True)
                c2 = plan.read(L.ARR_C2_COUNT)
                for t in np.nonzero(active)[0]:
                    ratio = int(c2[t]) / plan.track_frames[t]
                    T[t] = T_mid[t]
                    if abs(ratio - target_c2) < 0.01:
                        active[t] = False
                    elif ratio < target_c2:
                        T_high[t] = T_mid[t]
                    else:
                        T_low[t] = T_mid[t]
            Tg = np.nan_to_num(T)
            plan.gate(L.GATE_MINHOLD, L.ARR_GATE_F64, Tg + hyst_db / 2, Tg - hyst_db / 2, hold, 0, count_only=False)
            states = plan.read(L.ARR_STATE)
            for t in range(nt):
                fb, nf = plan.frame_base[t], plan.track_frames[t]
                st = states[fb:fb + nf].copy()
                res[t].update({name + "_levels": lv[t].copy(), name + "_T": float(T[t]), name + "_states": st,
                               name + "_c2": (int((st == 2).sum()) / nf if nf else float("nan"))})
        return res
    finally:
        plan.close()


# ------------------------------------------------------------------------------------------- validators (N3)
def frame_levels_wholefile(x, device: int = 0, mono: bool = False, n_fft=tb.N_FFT, hop=tb.HOP):
    """float32 mean squares + levels of the frames [i*hop, i*hop + n_fft), i < len(x) // hop (zero padded past the end):
    the frames validate_layer1.simulate_gate keeps (src/validate_layer1.py:132-141).  Returns (meansq f32, levels f64,
    plan) -- the analysis-only plan stays open for gate runs; the caller closes it."""
    torch = _torch()
    eng = get_engine(device, n_fft, hop)
    xd = _to_device(torch, [x], device)[0]
    n = int(xd.shape[0])
    plan = Plan(eng, L.FRAMING_WHOLEFILE, [L.TrackDesc(xd.data_ptr(), None, n, 0, n, 0, 0, 0, -1)])
    try:
        plan.levels(mono=mono)
        msq = plan.read(L.ARR_MEANSQ_F32)
    except Exception:
        plan.close()
        raise
    plan._keep = xd                                            # the plan reads this buffer
    return msq, tb.levels_from_meansq(msq), plan


def to_device(arrays, device: int = 0):
    """float32 [N, 2] host arrays -> device tensors (device tensors pass through)."""
    return _to_device(_torch(), arrays, device)


def cond_spectrum_median(x, y, frames, anchor_bins=None, device: int = 0, n_fft=tb.N_FFT, hop=tb.HOP) -> np.ndarray:
    """Per-bin median over `frames` of |Y| / |X| (tmt_cond_spectrum); x, y float32 [N, 2] (host or device)."""
    torch = _torch()
    eng = get_engine(device, n_fft, hop)
    xd, yd = _to_device(torch, [x, y], device)
    frames = np.ascontiguousarray(frames, dtype=np.int32)
    total = int(min(xd.shape[0], yd.shape[0]))
    a0, a1 = (0, -1) if anchor_bins is None else (int(anchor_bins[0]), int(anchor_bins[1]))
    out = np.zeros(eng.n_fft // 2 + 1, dtype=np.float32)
    L.check(eng.lib.tmt_cond_spectrum(eng.h, C.c_void_p(xd.data_ptr()), C.c_void_p(yd.data_ptr()), total,
                                      frames.ctypes.data_as(C.c_void_p), int(frames.size), a0, a1,
                                      out.ctypes.data_as(C.c_void_p), _stream_ptr(torch)), "tmt_cond_spectrum")
    return out


# ------------------------------------------------------------------------------------------- calibration (N4)
def calib_envelope(x, lo: int, hi: int, up: int, down: int, device: int = 0):
    """Envelope of x[lo:hi] for the delay estimate (src/calibrate_to_baseline_v2.py:57-61,71-74): power_mono, resample_poly
    (up, down), mean removed.  x: float32 [N, 2] host array or device tensor.  Returns a float32 device tensor."""
    torch = _torch()
    eng = get_engine(device)
    xd = _to_device(torch, [x], device)[0]
    lo, hi = int(lo), int(hi)
    if not 0 <= lo <= hi <= int(xd.shape[0]):
        raise ValueError(f"range [{lo}, {hi}) outside the {int(xd.shape[0])} sample-frames of the file")
    plan = tb.resample_poly_plan(hi - lo, up, down)
    out = torch.empty(plan["n_out"], dtype=torch.float32, device=xd.device)
    h = plan["h"] if plan["h"] is not None else np.ones(1, np.float32)          # up == down: resample_poly returns x itself
    h = np.ascontiguousarray(h, dtype=np.float32)
    L.check(eng.lib.tmt_calib_envelope_decimate(eng.h, C.c_void_p(xd.data_ptr() + 8 * lo), hi - lo,
                                                h.ctypes.data_as(C.c_void_p), int(h.size), plan["up"], plan["down"],
                                                plan["n_pre_remove"], plan["n_out"], C.c_void_p(out.data_ptr()),
                                                _stream_ptr(torch)), "tmt_calib_envelope_decimate")
    return out


def calib_xcorr(a_dev, b_dev) -> np.ndarray:
    """fftconvolve(a, b[::-1], "valid") (:77-78) -> host float32: corr[k] = sum_j a[k + j] * b[j], k = 0 .. len(a) - len(b).
    Like scipy, a second operand longer than the first swaps the roles: the result is then sum_j a[j] * b[len(b) - len(a) - k + j]."""
    torch = _torch()
    eng = get_engine(a_dev.device.index or 0)
    na, nb = int(a_dev.numel()), int(b_dev.numel())
    if na == 0 or nb == 0:
        raise ValueError("cross-correlation of an empty envelope")
    if nb > na:
        return calib_xcorr(b_dev, a_dev)[::-1].copy()
    corr = torch.empty(na - nb + 1, dtype=torch.float32, device=a_dev.device)
    L.check(eng.lib.tmt_calib_xcorr_valid(eng.h, C.c_void_p(a_dev.data_ptr()), na, C.c_void_p(b_dev.data_ptr()), nb,
                                          C.c_void_p(corr.data_ptr()), _stream_ptr(torch)), "tmt_calib_xcorr_valid")
    return corr.cpu().numpy()


def calib_frame_levels(x, device: int = 0) -> np.ndarray:
    """rms_dbfs_from_mono(power_mono(frame)) (:8-15,193-194) of the frames [i*hop, i*hop + n_fft) that lie inside x -> float32."""
    torch = _torch()
    eng = get_engine(device)
    xd = _to_device(torch, [x], device)[0]
    n = int(xd.shape[0])
    plan = Plan(eng, L.FRAMING_EQ_NOPAD, [L.TrackDesc(xd.data_ptr(), None, n, 0, n, 0, 0, 0, -1)])
    try:
        plan.levels(channel="power_eps")
        return tb.levels_from_meansq(plan.read(L.ARR_MEANSQ_F32)).astype(np.float32)
    finally:
        plan.close()


def calib_band_energies(x, n_frames: int, lo_bins, hi_bins, device: int = 0):
    """Band energies of stft_band_tilt (:17-30) for the first n_frames frames of x: sums of |rfft(power_mono * hann)|^2 over
    the bin ranges [lo_bins[0], lo_bins[1]) and [hi_bins[0], hi_bins[1]) -> (e_lo, e_hi) float32."""
    torch = _torch()
    eng = get_engine(device)
    xd = _to_device(torch, [x], device)[0]
    e_lo, e_hi = np.zeros(n_frames, np.float32), np.zeros(n_frames, np.float32)
    L.check(eng.lib.tmt_calib_band_energies(eng.h, C.c_void_p(xd.data_ptr()), int(xd.shape[0]), int(n_frames),
                                            int(lo_bins[0]), int(lo_bins[1]), int(hi_bins[0]), int(hi_bins[1]),
                                            e_lo.ctypes.data_as(C.c_void_p), e_hi.ctypes.data_as(C.c_void_p),
                                            _stream_ptr(torch)), "tmt_calib_band_energies")
    return e_lo, e_hi


def calib_gate_grid(level, start, want, on, off, delay, want_states: bool = False, device: int = 0):
    """simulate_state (:88-112) for len(on) parameter sets over the same frames -> (mismatches, switches[, states])."""
    torch = _torch()
    eng = get_engine(device)
    level = np.ascontiguousarray(level, dtype=np.float32)
    start = np.ascontiguousarray(start, dtype=np.int64)
    want = np.ascontiguousarray(want, dtype=np.uint8)
    on, off = np.ascontiguousarray(on, dtype=np.float32), np.ascontiguousarray(off, dtype=np.float32)
    delay = np.ascontiguousarray(delay, dtype=np.int64)
    n, nc = int(level.size), int(on.size)
    assert start.size == n and want.size == n and off.size == nc and delay.size == nc
    mis, sw = np.zeros(nc, np.int32), np.zeros(nc, np.int32)
    states = np.zeros((nc, n), np.uint8) if want_states else None
    L.check(eng.lib.tmt_calib_gate_grid(eng.h, level.ctypes.data_as(C.c_void_p), start.ctypes.data_as(C.c_void_p),
                                        want.ctypes.data_as(C.c_void_p), n, on.ctypes.data_as(C.c_void_p),
                                        off.ctypes.data_as(C.c_void_p), delay.ctypes.data_as(C.c_void_p), nc,
                                        mis.ctypes.data_as(C.c_void_p), sw.ctypes.data_as(C.c_void_p),
                                        states.ctypes.data_as(C.c_void_p) if want_states else None, _stream_ptr(torch)),
            "tmt_calib_gate_grid")
    return (mis, sw, states) if want_states else (mis, sw)


def run(mode: str, xs: Sequence, sr: int, **kw) -> List[dict]:
    if mode == "adaptive":
        return run_adaptive(xs, sr, **kw)
    return run_streaming(mode, xs, sr, **kw)


# ------------------------------------------------------------------------------------------------ static EQ (N1)
def run_eq(xs: Sequence, sr: int, gain_bins: np.ndarray, device: int = 0, pad: bool = True, global_gain_db: float = 0.0,
           auto_gain_protect: bool = True, peak_target: float = 0.99, want_host: bool = True, unit_blocks: int = 0,
           n_fft=tb.N_FFT, hop=tb.HOP) -> List[dict]:
    """Static EQ through the same fused STFT/OLA kernel with a one-row gain table (src/layer2_apply_eq.py:66-237).

    Per track: out = what the reference writes to its output file (every position its frames cover: length
    (n_frames+1)*hop, shifted by n_fft/2 when pad), peak_seen, scale and out_gp = the gain-protected second file
    (PCM_24 round trip of `out` times peak_target/peak) when the peak exceeds peak_target."""
    torch = _torch()
    if not fused_size(n_fft, hop):
        raise NotImplementedError(f"GPU path implements n_fft/hop = 4096/2048 and 2048/1024 here; got {n_fft}/{hop}")
    eng = get_engine(device, n_fft, hop)
    gain_bins = np.ascontiguousarray(gain_bins, dtype=np.float32).reshape(1, -1)
    eng.set_gain_rows(gain_bins, key=None)
    g_global = 10.0 ** (global_gain_db / 20.0)                      # db_to_lin, src/layer2_apply_eq.py:8-9,115
    xd = _to_device(torch, xs, device)
    framing = L.FRAMING_EQ_PAD if pad else L.FRAMING_EQ_NOPAD
    first = -(n_fft // 2) if pad else 0
    descs, yd = [], []
    for x in xd:
        n = int(x.shape[0])
        nf = (n // hop + 1) if pad else ((n - n_fft) // hop + 1 if n >= n_fft else 0)
        out_len = (nf + 1) * hop if nf > 0 else 0
        y = torch.empty((out_len, 2), dtype=torch.float32, device=x.device)
        yd.append(y)
        descs.append(L.TrackDesc(x.data_ptr(), y.data_ptr(), n, 0, n, first, out_len, 0, -1))
    plan = Plan(eng, framing, descs, unit_blocks)
    try:
        nt = plan.n_tracks
        plan.stft(float(np.float32(g_global)), skip_edges=True)      # rows are all zero: the single EQ row
        plan.edge_frames(1.0, in_scale=np.full(nt, np.float32(g_global), np.float32) if global_gain_db != 0.0 else None)
        peaks = plan.read(L.ARR_CHUNK_PEAK)
        res = []
        for t in range(nt):
            peak_seen = float(peaks[plan.chunk_base[t]]) if plan.track_chunks[t] else 0.0
            scale, y_gp = None, None
            if auto_gain_protect and peak_seen > peak_target:
                scale = peak_target / max(peak_seen, tb.EPS)
                y_gp = yd[t].clone()
                L.check(eng.lib.tmt_requantise_scale(C.c_void_p(y_gp.data_ptr()), y_gp.numel(), float(np.float32(scale)),
                                                     _stream_ptr(torch)), "tmt_requantise_scale")
            res.append(dict(out=(yd[t].cpu().numpy() if want_host else yd[t]), peak_seen=peak_seen, scale=scale,
                            out_gp=(None if y_gp is None else (y_gp.cpu().numpy() if want_host else y_gp)),
                            n_frames=plan.track_frames[t], launches=plan.launch_count()))
        return res
    finally:
        plan.close()
        eng._rows_key = None
