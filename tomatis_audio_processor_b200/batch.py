"""Batch / host-buffer front end: many equal-length tracks through the streaming modes.

`DeviceBatch` keeps a plan over device-resident tracks (the resident-in-HBM benchmark arm);
`HostBatchPipeline` is the host-buffer entry point: tracks live in pinned host memory, are staged to
the GPU in waves, processed, and copied back, with H2D / kernels / D2H of consecutive waves overlapped
on three streams.  Both mirror what a user of the reference does with a directory of files
(`docs/Tomatis处理器使用指南.md:243-249` loops process() over files); sharding by whole track across
ranks needs no collective.
"""
from __future__ import annotations

from . import _lib as L
from .engine import Plan, get_engine, streaming_params, whole_track_desc, _torch


class DeviceBatch:
    """Plan over `x` [T, N, 2] float32 CUDA tensor -> `y` (same shape), standard or xfade mode."""

    def __init__(self, x, y, sr: int, mode: str = "standard", device: int = 0, unit_blocks: int = 0, **params):
        self.torch = _torch()
        self.eng = get_engine(device, params.get("n_fft", 4096), params.get("hop", 2048))      # a fused frame size (engine.fused_size)
        self.sp = streaming_params(mode, sr, **params)
        self.eng.set_gain_rows(self.sp.rows, key=self.sp.rows_key)
        self.x, self.y, self.sr = x, y, sr
        self.plan = Plan(self.eng, L.FRAMING_STREAMING, [whole_track_desc(x[i], y[i]) for i in range(x.shape[0])], unit_blocks)

    # the individual launches of one step (bench.py brackets `stft` with its own events)
    def levels(self):
        self.plan.levels(False, None)

    def gate(self):
        sp = self.sp
        self.plan.gate(L.GATE_UPDELAY, L.ARR_MEANSQ_F32, sp.m_on, sp.m_off, sp.run_frames, sp.xfade_frames)

    def edges(self):
        self.plan.clear_peaks()
        self.plan.edge_frames(self.sp.post_gain)

    def stft(self):
        """Fused STFT/OLA + per-chunk limiter (edge blocks must already be in place: call edges() first)."""
        self.plan.stft_limited(self.sp.post_gain)

    def step(self):
        self.eng.set_gain_rows(self.sp.rows, key=self.sp.rows_key)     # no-op unless another caller replaced the engine's table
        self.levels(); self.gate(); self.edges(); self.stft()

    def close(self):
        self.plan.close()


class HostBatchPipeline:
    """process(host_in, host_out): pinned host tracks -> pinned host tracks.

    Waves of `wave_tracks` tracks rotate through `n_slots` device staging slots; copy-in, compute and
    copy-out run on separate streams, ordered by events, so PCIe transfers overlap the kernels.  Small waves keep the
    pipeline's fill and drain short: 2 x 5-minute tracks per wave measured 14 % faster end to end than 8 (PCIe gives
    46-48 GB/s each way on this box when both directions run; the pipeline reaches ~45).

    in_format / out_format select what crosses PCIe:
      "f32"  float32 [T, N, 2]                       8 B per sample-frame each way (what the reference holds in memory)
      "s16"  int16   [T, N, 2]  (input only)         4 B/sf; converted on the device as soundfile would (value / 32768)
      "s24"  uint8   [T, N, 6]  packed PCM_24        6 B/sf; input value / 8388608, output lrint(y * 2^23) clipped (FLAC rule) --
                                                     the reference's output files are PCM_24 (src/process_tomatis.py:243)
    """

    def __init__(self, n_samples: int, sr: int, mode: str = "standard", device: int = 0, wave_tracks: int = 2,
                 n_slots: int = 3, unit_blocks: int = 0, in_format: str = "f32", out_format: str = "f32", **params):
        torch = self.torch = _torch()
        from .engine import pcm_to_float, float_to_pcm24
        self._to_float, self._to_pcm24 = pcm_to_float, float_to_pcm24
        if in_format not in ("f32", "s16", "s24") or out_format not in ("f32", "s24"):
            raise ValueError("in_format: f32|s16|s24, out_format: f32|s24")
        self.in_format, self.out_format = in_format, out_format
        import os
        self.fused_pcm = os.environ.get("TMT_PCM_FUSED", "1") != "0"       # 0: separate conversion + levels pass (comparisons)
        self.eng = get_engine(device, params.get("n_fft", 4096), params.get("hop", 2048))      # a fused frame size (engine.fused_size)
        self.sp = streaming_params(mode, sr, **params)
        self.eng.set_gain_rows(self.sp.rows, key=self.sp.rows_key)
        dev = f"cuda:{device}"
        self.W, self.N = wave_tracks, n_samples
        self.s_in, self.s_c, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
        self.slots = []
        for _ in range(n_slots):
            x = torch.empty((wave_tracks, n_samples, 2), dtype=torch.float32, device=dev)
            y = torch.empty_like(x)
            raw_in = raw_out = None
            if in_format == "s16":
                raw_in = torch.empty((wave_tracks, n_samples, 2), dtype=torch.int16, device=dev)
            elif in_format == "s24":
                raw_in = torch.empty((wave_tracks, n_samples, 6), dtype=torch.uint8, device=dev)
            if out_format == "s24":
                raw_out = torch.empty((wave_tracks, n_samples, 6), dtype=torch.uint8, device=dev)
            plan = Plan(self.eng, L.FRAMING_STREAMING, [whole_track_desc(x[i], y[i]) for i in range(wave_tracks)], unit_blocks)
            self.slots.append(dict(x=x, y=y, raw_in=raw_in, raw_out=raw_out, plan=plan, ev_in=torch.cuda.Event(),
                                   ev_c=torch.cuda.Event(), ev_out=torch.cuda.Event(), used=False))
        self.launches = 0

    def bytes_per_sample_frame(self):
        return {"f32": 8, "s16": 4, "s24": 6}[self.in_format], {"f32": 8, "s24": 6}[self.out_format]

    def process(self, host_in, host_out):
        torch, sp = self.torch, self.sp
        self.eng.set_gain_rows(sp.rows, key=sp.rows_key)               # no-op unless another caller replaced the engine's table
        T = host_in.shape[0]
        if T % self.W:
            raise ValueError(f"track count {T} must be a multiple of the wave size {self.W}")
        cur = torch.cuda.current_stream()
        for s in (self.s_in, self.s_c, self.s_out):
            s.wait_stream(cur)
        for w in range(T // self.W):
            sl = self.slots[w % len(self.slots)]
            lo, hi = w * self.W, (w + 1) * self.W
            with torch.cuda.stream(self.s_in):
                if sl["used"]:
                    self.s_in.wait_event(sl["ev_c"])          # previous compute on this slot has consumed its input
                (sl["x"] if sl["raw_in"] is None else sl["raw_in"]).copy_(host_in[lo:hi], non_blocking=True)
                sl["ev_in"].record(self.s_in)
            with torch.cuda.stream(self.s_c):
                self.s_c.wait_event(sl["ev_in"])
                if sl["used"]:
                    self.s_c.wait_event(sl["ev_out"])         # previous copy-out of this slot has finished
                before = sl["plan"].launch_count()
                if sl["raw_in"] is not None and self.fused_pcm:
                    # integer input: conversion and hop-block sums in ONE pass over the samples, no levels pass over the floats
                    sl["plan"].run_streaming_pcm(sl["raw_in"], L.PCM_S16 if self.in_format == "s16" else L.PCM_S24,
                                                 sp.m_on, sp.m_off, sp.run_frames, sp.xfade_frames, sp.post_gain)
                else:
                    if sl["raw_in"] is not None:
                        self._to_float(sl["raw_in"], L.PCM_S16 if self.in_format == "s16" else L.PCM_S24, sl["x"])
                        self.launches += 1
                    sl["plan"].run_streaming(sp.m_on, sp.m_off, sp.run_frames, sp.xfade_frames, sp.post_gain)
                if sl["raw_out"] is not None:
                    self._to_pcm24(sl["y"], sl["raw_out"])
                    self.launches += 1
                self.launches += sl["plan"].launch_count() - before
                sl["ev_c"].record(self.s_c)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(sl["ev_c"])
                host_out[lo:hi].copy_(sl["y"] if sl["raw_out"] is None else sl["raw_out"], non_blocking=True)
                sl["ev_out"].record(self.s_out)
            sl["used"] = True
        for s in (self.s_in, self.s_c, self.s_out):
            cur.wait_stream(s)

    def close(self):
        for sl in self.slots:
            sl["plan"].close()
