"""One long file from / to host buffers in bounded device memory (standard / xfade).

The reference streams: its frame loop holds a few seconds of audio at a time, whatever the file length
(src/process_tomatis.py:359-453, flush every 240 000 samples :419-426).  `engine.run_streaming` instead keeps the whole file in HBM
and `sharded.StreamingShardSession` each rank's whole share.  `HostFileStreamer` is the streaming form of the same kernels for a file
that lives in (pinned) host memory: the file is cut into time slabs on limiter-chunk boundaries (`sharded.plan_shards`, the cut that
keeps every per-chunk limiter inside one slab), the slabs rotate through a few device slots, and the copy-in of slab k+1, the kernels
of slab k and the copy-out of slab k-1 run on three streams.  The gate is a chain along the file: a slab adds the hop-block sums of
the blocks it owns (plus the one block after them, which its right halo covers and its last frame needs) to a per-file array on the
device, and scans the gate over everything up to its own end -- a slab never needs anything from a LATER slab, so the pipeline
never stalls.  Device memory: n_slots x (slab + two hops in, slab out), independent of the file length.

The output is bit-identical to the whole-file call (the same kernels on the same sample positions; the slabs are the shards of the
time-sharded path, whose output equals the unsharded one bit for bit, tests/test_gpu_sharded.py)."""
from __future__ import annotations

import numpy as np

from . import _lib as L
from . import tables as tb
from .engine import _torch, streaming_params
from .sharded import STREAMING, CudaShardBackend, plan_shards


def plan_slabs(total: int, sr: int, slab_seconds: float = 300.0, n_fft: int = tb.N_FFT, hop: int = tb.HOP):
    """Host-side geometry of the streamed pass: [(shard, hb_lo, hb_hi, f_hi)] in file order.  shard = the slab's block / sample
    ranges (`sharded.plan_shards`, cut on limiter-chunk boundaries); [hb_lo, hb_hi) = the hop blocks the slab sums itself -- its
    own and, when its right halo covers it, the one after them, which its last frame spans; f_hi = frames [0, f_hi) get their
    mean square (and a valid gate state) when the slab runs: everything up to the slab's own last frame."""
    n_frames = tb.streaming_frame_count(int(total), n_fft, hop)
    n_chunks = max(1, len(tb.flush_chunk_blocks(n_frames, n_fft, hop)))
    k = int(np.clip(round(total / max(1.0, slab_seconds * sr)), 1, n_chunks))
    nb = n_frames + 1 if n_frames > 0 else 0
    out = []
    for s in plan_shards(int(total), k, STREAMING, n_fft, hop):
        if s.own_hi <= s.own_lo:
            continue
        hb_lo = min(s.block_lo, nb)
        hb_hi = min(s.block_hi + (1 if s.in_hi >= s.first_start + (s.block_hi + 1) * hop else 0), nb)
        out.append((s, hb_lo, hb_hi, min(s.block_hi, n_frames)))
    return out


class HostFileStreamer:
    def __init__(self, mode: str, total: int, sr: int, device: int = 0, slab_seconds: float = 300.0, n_slots: int = 3,
                 unit_blocks: int = 0, **params):
        if int(total) <= 0:
            raise ValueError("empty input file")
        if mode not in ("standard", "xfade"):
            raise ValueError("HostFileStreamer serves the streaming modes (standard, xfade); adaptive needs the whole file's peak first")
        torch = self.torch = _torch()
        from .engine import fused_size
        n_fft, hop = params.get("n_fft", tb.N_FFT), params.get("hop", tb.HOP)
        if not fused_size(n_fft, hop):
            raise NotImplementedError(f"HostFileStreamer implements n_fft/hop = 4096/2048 and 2048/1024; got {n_fft}/{hop}")
        self.total, self.sr = int(total), sr
        self.sp = streaming_params(mode, sr, **params)
        dev = torch.device(f"cuda:{device}")
        slabs = plan_slabs(self.total, sr, slab_seconds, n_fft, hop)
        self.shards = [sl[0] for sl in slabs]
        n_frames = self.shards[0].n_frames if self.shards else 0
        self.n_frames = n_frames
        max_in = max(s.in_hi - s.in_lo for s in self.shards)
        max_own = max(s.own_hi - s.own_lo for s in self.shards)
        n_slots = min(n_slots, len(self.shards))
        self.slots = [dict(win=torch.empty((max_in, 2), dtype=torch.float32, device=dev),
                           out=torch.empty((max_own, 2), dtype=torch.float32, device=dev),
                           ev_in=torch.cuda.Event(), ev_c=torch.cuda.Event(), ev_out=torch.cuda.Event(), used=False)
                      for _ in range(n_slots)]
        self.hsum = torch.zeros(n_frames + 2, dtype=torch.float32, device=dev)      # hop-block sums of the file, filled slab by slab
        self.s_in, self.s_c, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
        self.slabs = []
        for i, (s, hb_lo, hb_hi, f_hi) in enumerate(slabs):
            sl = self.slots[i % n_slots]
            be = CudaShardBackend(s, sl["win"][:s.in_hi - s.in_lo], device, self.sp.rows, self.sp.rows_key, unit_blocks,
                                  out=sl["out"][:s.own_hi - s.own_lo])
            if be.plan.unfusable_chunks:
                raise RuntimeError("slab cut through a limiter chunk")
            be.plan.set_level_ranges(0, hb_lo, hb_hi, 0, f_hi)
            self.slabs.append(dict(shard=s, be=be, slot=sl, hb=(hb_lo, hb_hi), thresholds=False))
        self.launches = 0

    def process(self, host_in, host_out):
        """host_in / host_out: float32 [total, 2] tensors (pinned for asynchronous copies).  Returns when everything is queued;
        the caller's stream waits for the three pipeline streams."""
        torch, sp = self.torch, self.sp
        assert host_in.shape[0] == self.total and host_out.shape[0] == self.total
        cur = torch.cuda.current_stream()
        for st in (self.s_in, self.s_c, self.s_out):
            st.wait_stream(cur)
        for sb in self.slabs:
            s, be, sl = sb["shard"], sb["be"], sb["slot"]
            with torch.cuda.stream(self.s_in):
                if sl["used"]:
                    self.s_in.wait_event(sl["ev_c"])                       # the previous slab in this slot has read its input
                sl["win"][:s.in_hi - s.in_lo].copy_(host_in[s.in_lo:s.in_hi], non_blocking=True)
                sl["ev_in"].record(self.s_in)
            with torch.cuda.stream(self.s_c):
                self.s_c.wait_event(sl["ev_in"])
                if sl["used"]:
                    self.s_c.wait_event(sl["ev_out"])                      # ... and its output has left the slot
                before = be.plan.launch_count()
                lo, hi = sb["hb"]
                be.plan.levels(part="hopsums")
                if hi > lo:
                    be.plan.read_device(L.ARR_HOPSUM_F32, self.hsum[lo:].data_ptr(), hi - lo, offset=lo)
                if hi > 0:
                    be.plan.write_device(L.ARR_HOPSUM_F32, self.hsum.data_ptr(), hi)      # everything up to this slab's end
                be.plan.levels(part="meansq")
                if not sb["thresholds"]:
                    be.gate(L.GATE_UPDELAY, L.ARR_MEANSQ_F32, sp.m_on, sp.m_off, sp.run_frames, sp.xfade_frames)
                    sb["thresholds"] = True
                else:
                    be.gate(L.GATE_UPDELAY, L.ARR_MEANSQ_F32, None, None, sp.run_frames, sp.xfade_frames)
                be.plan.clear_peaks()
                be.edge_frames(sp.post_gain)
                be.plan.stft_limited(sp.post_gain)
                self.launches += be.plan.launch_count() - before
                sl["ev_c"].record(self.s_c)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(sl["ev_c"])
                host_out[s.own_lo:s.own_hi].copy_(sl["out"][:s.own_hi - s.own_lo], non_blocking=True)
                sl["ev_out"].record(self.s_out)
            sl["used"] = True
        for st in (self.s_in, self.s_c, self.s_out):
            cur.wait_stream(st)

    def states_rows(self):
        """Gate states and gain-row indices of every frame of the file (the last slab scanned them all)."""
        return self.slabs[-1]["be"].states_rows()

    def device_bytes(self) -> int:
        return sum(sl["win"].numel() * 4 + sl["out"].numel() * 4 for sl in self.slots)

    def close(self):
        for sb in self.slabs:
            sb["be"].close()
