"""One long file on several GPUs: time-chunk sharding with one-hop halos.

Rank r owns a contiguous run of output hop-blocks -- aligned to the reference's limiter-chunk boundaries
(src/process_tomatis.py:419-426) whenever there are at least as many chunks as ranks -- and holds only its
own samples.  Cross-rank traffic, all over torch.distributed (NCCL on GPUs, gloo in the CPU tests):

  1. halo hand-off: the hop before and the hop after the owned range come from the neighbouring ranks
     (point-to-point send/recv; a frame overlaps its neighbours by 50 %);
  2. gate state: every rank reduces its frames' mean squares (bit-exact values, each frame owned by exactly
     one rank, so an all-reduce(SUM) over zero-filled arrays is an exact gather), then every rank runs the
     identical gate scan over the whole file -- the "carried gate state" is recomputed instead of passed
     along, which keeps the ranks independent and deterministic (337 500 frames = 1.35 MB for 2 h @ 96 kHz);
  3. limiter: all-reduce(MAX) of the per-chunk peaks (adaptive: also of the input peak);
  4. optional final gather of the output shards to rank 0.

The numeric work is done by a *backend* with the interface of engine.Plan (CudaShardBackend below); the CPU
tests drive the same code with a NumPy stand-in built from the oracle.

`run_streaming_sharded` / `run_adaptive_sharded` are the one-shot drivers (they also return the per-frame data of the
whole file on every rank).  `StreamingShardSession` keeps plan, window and output on the device for repeated passes and
trims the hand-offs: halos as one all-gather of each rank's first and last hop, issued first and awaited only before the
STFT; levels as an all-reduce of hop-block sums, which need no halo at all.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import tables as tb

STREAMING, WHOLEFILE = 0, 1


@dataclass
class Shard:
    rank: int
    world: int
    total: int            # file length in sample-frames
    framing: int
    n_frames: int
    first_start: int      # position of frame 0
    block_lo: int         # owned output blocks [block_lo, block_hi)
    block_hi: int
    own_lo: int           # owned sample range [own_lo, own_hi) (what the rank holds and produces)
    own_hi: int
    in_lo: int            # samples the rank's frames read: [in_lo, in_hi) (own range + halos)
    in_hi: int
    frame_lo: int         # frames whose level this rank is the owner of: [frame_lo, frame_hi)
    frame_hi: int
    n_fft: int = tb.N_FFT # frame size of the plan (the fused kernels serve 4096 / 2048 and 2048 / 1024)
    hop: int = tb.HOP


def plan_shards(total: int, world: int, framing: int, n_fft=tb.N_FFT, hop=tb.HOP) -> List[Shard]:
    """Split the output blocks of one file into `world` contiguous runs."""
    if framing == STREAMING:
        n_frames = tb.streaming_frame_count(total, n_fft, hop)
        first = -(n_fft // 2)
        chunks = tb.flush_chunk_blocks(n_frames, n_fft, hop)
    else:
        n_frames = tb.wholefile_frame_count(total, hop)
        first = 0
        chunks = [(0, n_frames + 1)] if n_frames > 0 else []
    n_blocks = n_frames + 1 if n_frames > 0 else 0
    if len(chunks) >= world:            # whole limiter chunks per rank: chunk peaks stay rank-local
        cuts = [chunks[(len(chunks) * r) // world][0] for r in range(world)] + [n_blocks]
    else:                               # fewer chunks than ranks: split by blocks (peaks are all-reduced anyway)
        cuts = [(n_blocks * r) // world for r in range(world)] + [n_blocks]
    shards = []
    for r in range(world):
        blo, bhi = cuts[r], cuts[r + 1]
        pos = lambda b: min(total, max(0, first + b * hop))
        own_lo = pos(blo) if r > 0 else 0
        own_hi = pos(bhi) if r < world - 1 else total
        f_lo, f_hi = max(0, blo - 1), min(n_frames, bhi)
        if f_hi > f_lo:
            in_lo, in_hi = max(0, first + f_lo * hop), min(total, first + (f_hi - 1) * hop + n_fft)
        else:
            in_lo = in_hi = own_lo
        in_lo, in_hi = min(in_lo, own_lo), max(in_hi, own_hi)
        shards.append(Shard(r, world, total, framing, n_frames, first, blo, bhi, own_lo, own_hi, in_lo, in_hi,
                            min(blo, n_frames), min(bhi, n_frames), n_fft, hop))
    return shards


# ------------------------------------------------------------------------------------------------ collectives
class Comm:
    """Thin wrapper over a torch.distributed process group (tensors live on `device`)."""

    def __init__(self, group=None, device="cpu"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group, self.device = torch, dist, group, torch.device(device)
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.bytes_sent = 0

    def _t(self, a: np.ndarray):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

    def allreduce(self, a, op: str):
        """In-place all-reduce of a tensor already on the communicator's device (no host round trip), or of a NumPy
        array (copied to the device and back)."""
        is_np = isinstance(a, np.ndarray)
        t = self._t(a) if is_np else a
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM if op == "sum" else self.dist.ReduceOp.MAX, group=self.group)
        self.bytes_sent += t.numel() * t.element_size()
        return t.cpu().numpy() if is_np else t

    def exchange_halos(self, own, shard: Shard, shards: List[Shard], window=None):
        """own: tensor [own_hi-own_lo, 2] on self.device.  Returns the rank's input window [in_hi-in_lo, 2]:
        own samples plus the halos fetched from whichever ranks own them (normally the two neighbours).
        window: existing window buffer of which `own` is already the middle slice (only the halos are moved)."""
        torch, dist = self.torch, self.dist
        if window is None:
            win = torch.zeros((shard.in_hi - shard.in_lo, 2), dtype=own.dtype, device=own.device)
            win[shard.own_lo - shard.in_lo: shard.own_hi - shard.in_lo] = own
        else:
            win = window
        ops, recvs = [], []
        for other in shards:
            if other.rank == shard.rank:
                continue
            # what `other` needs from me
            lo, hi = max(other.in_lo, shard.own_lo), min(other.in_hi, shard.own_hi)
            if hi > lo:
                buf = own[lo - shard.own_lo: hi - shard.own_lo].contiguous()
                ops.append(dist.P2POp(dist.isend, buf, other.rank, group=self.group))
                self.bytes_sent += buf.numel() * buf.element_size()
            # what I need from `other`
            lo, hi = max(shard.in_lo, other.own_lo), min(shard.in_hi, other.own_hi)
            if hi > lo:
                buf = torch.empty((hi - lo, 2), dtype=own.dtype, device=own.device)
                ops.append(dist.P2POp(dist.irecv, buf, other.rank, group=self.group))
                recvs.append((lo, hi, buf))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for lo, hi, buf in recvs:
            win[lo - shard.in_lo: hi - shard.in_lo] = buf
        return win

    def exchange_halos_begin(self, own, shard: Shard, shards: List[Shard]):
        """Start the halo hand-off without waiting; finish it with exchange_halos_end.  Lets the rank sum its own hop blocks
        and reduce the levels while the halos travel.

        Usual case (every halo is at most one hop and comes from the adjacent rank): ONE all-gather of each rank's first
        and last hop (32 KB per rank) instead of a group of point-to-point transfers -- the group's host-side launch cost
        (~0.25 ms) was a fifth of the 8-GPU step.  Anything else falls back to point-to-point."""
        torch, dist = self.torch, self.dist
        hop = shard.hop
        if len(shards) == 1:
            return ("none", None, None)
        simple = all(s.own_hi - s.own_lo >= hop and s.own_lo - s.in_lo <= hop and s.in_hi - s.own_hi <= hop for s in shards)
        if simple:
            edge = torch.empty((2, hop, 2), dtype=own.dtype, device=own.device)
            edge[0] = own[:hop]
            edge[1] = own[-hop:]
            allg = torch.empty((len(shards), 2, hop, 2), dtype=own.dtype, device=own.device)
            work = dist.all_gather_into_tensor(allg, edge, group=self.group, async_op=True)
            self.bytes_sent += edge.numel() * edge.element_size()
            return ("allgather", work, allg)
        ops, recvs = [], []
        for other in shards:
            if other.rank == shard.rank:
                continue
            lo, hi = max(other.in_lo, shard.own_lo), min(other.in_hi, shard.own_hi)
            if hi > lo:
                buf = own[lo - shard.own_lo: hi - shard.own_lo].contiguous()
                ops.append(dist.P2POp(dist.isend, buf, other.rank, group=self.group))
                self.bytes_sent += buf.numel() * buf.element_size()
            lo, hi = max(shard.in_lo, other.own_lo), min(shard.in_hi, other.own_hi)
            if hi > lo:
                buf = torch.empty((hi - lo, 2), dtype=own.dtype, device=own.device)
                ops.append(dist.P2POp(dist.irecv, buf, other.rank, group=self.group))
                recvs.append((lo, hi, buf))
        return ("p2p", (dist.batch_isend_irecv(ops) if ops else []), recvs)

    def exchange_halos_end(self, pending, shard: Shard, window):
        kind, work, data = pending
        if kind == "none":
            return
        if kind == "allgather":
            work.wait()
            left, right = shard.own_lo - shard.in_lo, shard.in_hi - shard.own_hi
            if left > 0:                                     # tail of the previous rank's last hop
                window[:left] = data[shard.rank - 1, 1, shard.hop - left:]
            if right > 0:                                    # head of the next rank's first hop
                window[window.shape[0] - right:] = data[shard.rank + 1, 0, :right]
            return
        for req in work:
            req.wait()
        for lo, hi, buf in data:
            window[lo - shard.in_lo: hi - shard.in_lo] = buf

    def gather_output(self, own_out, shards: List[Shard], dst: int = 0):
        """Concatenate the ranks' output shards on rank `dst` (None elsewhere)."""
        torch, dist = self.torch, self.dist
        if self.rank == dst:
            total = shards[0].total
            full = torch.empty((total, 2), dtype=own_out.dtype, device=own_out.device)
            me = shards[dst]
            full[me.own_lo:me.own_hi] = own_out
            ops, bufs = [], []
            for s in shards:
                if s.rank != dst and s.own_hi > s.own_lo:
                    buf = full[s.own_lo:s.own_hi]            # contiguous row slice: receive in place
                    ops.append(dist.P2POp(dist.irecv, buf, s.rank, group=self.group))
            for req in (dist.batch_isend_irecv(ops) if ops else []):
                req.wait()
            return full
        if own_out.shape[0] > 0:
            for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, own_out.contiguous(), dst, group=self.group)]):
                req.wait()
            self.bytes_sent += own_out.numel() * own_out.element_size()
        return None


# ------------------------------------------------------------------------------------------------ peer-memory exchange
class PeerExchange:
    """The per-pass exchange of the persistent session through peer memory instead of collectives (include/tomatis_b200.h,
    "peer memory"): every rank owns an exchange buffer, maps the other ranks' buffers through CUDA IPC (handles travel once, over
    the process group), and a pass is two small kernels -- publish (own hop sums into every rank's buffer, edge hops into the
    neighbours', flags raised) and wait + unpack -- instead of an all-gather, an all-reduce and the copies around them.
    Construction is collective; `ok` is the same on every rank (False: some rank could not map a peer -> the session keeps
    using the collectives)."""

    def __init__(self, comm: "Comm", shard: Shard, device_index: int, timeout_s: float = 5.0):
        import ctypes as C
        from . import _lib as L
        self.L, self.C, self.lib = L, C, L.load(shard.n_fft)
        self.comm, self.shard, self.device, self.timeout_s = comm, shard, device_index, timeout_s
        self.nb = shard.n_frames + 1 if shard.n_frames > 0 else 0
        self.own_ptr, self.peers, self.ok = None, {}, False
        torch, dist = comm.torch, comm.dist
        handle = (C.c_ubyte * 64)()
        ptr = C.c_void_p()
        good = 1
        try:
            L.check(self.lib.tmt_peer_alloc(device_index, self.lib.tmt_peer_bytes(self.nb), C.byref(ptr), handle), "tmt_peer_alloc")
            self.own_ptr = ptr.value
        except Exception:
            good = 0
        mine = torch.tensor(list(bytes(handle)) + [good], dtype=torch.uint8, device=comm.device)
        allh = torch.empty((comm.world, 65), dtype=torch.uint8, device=comm.device)
        dist.all_gather_into_tensor(allh, mine, group=comm.group)
        allh = allh.cpu().numpy()
        good = int(allh[:, 64].min())
        if good:
            for r in range(comm.world):
                if r == comm.rank:
                    continue
                q = C.c_void_p()
                h = (C.c_ubyte * 64)(*allh[r, :64].tolist())
                if self.lib.tmt_peer_open(device_index, h, C.byref(q)) != 0:
                    good = 0
                    break
                self.peers[r] = q.value
        flag = torch.tensor([good], dtype=torch.int32, device=comm.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=comm.group)
        self.ok = bool(int(flag.item()))
        if not self.ok:
            self.close()
            return
        self.bases = (C.c_void_p * comm.world)(*[(self.own_ptr if r == comm.rank else self.peers[r]) for r in range(comm.world)])

    def publish(self, plan, own):
        C, hop = self.C, self.shard.hop
        first = own.data_ptr() if self.shard.rank > 0 else None
        last = own[own.shape[0] - hop:].data_ptr() if self.shard.rank + 1 < self.shard.world else None
        from .engine import _stream_ptr
        self.L.check(self.lib.tmt_plan_peer_publish(plan.h, self.shard.rank, self.shard.world, self.bases, C.c_void_p(first), C.c_void_p(last),
                                                    _stream_ptr(self.comm.torch)), "tmt_plan_peer_publish")

    def wait_unpack(self, plan, window):
        C, s = self.C, self.shard
        from .engine import _stream_ptr
        self.L.check(self.lib.tmt_plan_peer_wait(plan.h, s.world, C.c_void_p(self.own_ptr), C.c_void_p(window.data_ptr()), int(window.shape[0]),
                                                 s.own_lo - s.in_lo, s.in_hi - s.own_hi, float(self.timeout_s), _stream_ptr(self.comm.torch)),
                     "tmt_plan_peer_wait")

    def status(self) -> int:
        """0, or 1 + the rank whose flag never arrived within the timeout (the pass that saw it produced garbage)."""
        v = self.C.c_int32(0)
        self.L.check(self.lib.tmt_peer_status(self.device, self.C.c_void_p(self.own_ptr), self.C.byref(v)), "tmt_peer_status")
        return int(v.value)

    def close(self):
        """Collective (every rank calls it, also on the failure path of the constructor): unmap the peers, wait for everybody,
        free the own buffer."""
        if getattr(self, "_closed", False):
            return
        self._closed = True
        for q in self.peers.values():
            self.lib.tmt_peer_close(self.device, self.C.c_void_p(q))
        self.peers = {}
        if self.comm.world > 1:
            self.comm.dist.barrier(group=self.comm.group)          # nobody frees a buffer a peer still has mapped and may write to
        if self.own_ptr:
            self.lib.tmt_peer_free(self.device, self.C.c_void_p(self.own_ptr))
            self.own_ptr = None


# ------------------------------------------------------------------------------------------------ CUDA backend
class CudaShardBackend:
    """engine.Plan over one shard: input window + owned output buffer on this rank's GPU."""

    def __init__(self, shard: Shard, window, device_index: int, gain_rows: np.ndarray, rows_key, unit_blocks: int = 0, out=None):
        import torch
        from . import _lib as L
        from .engine import Plan, get_engine
        self.L, self.shard = L, shard
        self.eng = get_engine(device_index, shard.n_fft, shard.hop)
        self.eng.set_gain_rows(gain_rows, key=rows_key)
        self.window = window
        # out: a caller-owned buffer for the shard's output (streamed.HostFileStreamer rotates a few device slots)
        self.out = out if out is not None else torch.empty((shard.own_hi - shard.own_lo, 2), dtype=torch.float32, device=window.device)
        assert self.out.shape[0] == shard.own_hi - shard.own_lo
        desc = L.TrackDesc(window.data_ptr(), self.out.data_ptr(), shard.total, shard.in_lo, shard.in_hi - shard.in_lo,
                           shard.own_lo, shard.own_hi - shard.own_lo, shard.block_lo, shard.block_hi)
        self.plan = Plan(self.eng, L.FRAMING_STREAMING if shard.framing == STREAMING else L.FRAMING_WHOLEFILE, [desc], unit_blocks)
        assert self.plan.track_frames[0] == shard.n_frames, (self.plan.track_frames, shard.n_frames)

    # -- levels
    def input_peak(self) -> np.float32:
        self.plan.input_peaks()
        return self.plan.read(self.L.ARR_INPUT_PEAK)[0]

    def local_meansq(self, use_f64=False, in_scale=None):
        """Mean squares of the frames touching this shard, as a device tensor [n_frames] (zeros elsewhere)."""
        import torch
        self.plan.levels(use_f64=use_f64, in_scale=None if in_scale is None else np.array([in_scale], np.float32))
        t = torch.empty(self.shard.n_frames, dtype=torch.float64 if use_f64 else torch.float32, device=self.window.device)
        if t.numel():
            self.plan.read_device(self.L.ARR_MEANSQ_F64 if use_f64 else self.L.ARR_MEANSQ_F32, t.data_ptr(), t.numel())
        return t

    def own_hop_sums(self):
        """Pairwise sums of the hop blocks this rank owns (needs no halo), device tensor [n_frames + 1], zeros elsewhere."""
        import torch
        s = self.shard
        nb = s.n_frames + 1 if s.n_frames > 0 else 0
        if not getattr(self, "_ranges_set", False):
            self.plan.set_level_ranges(0, min(s.block_lo, nb), min(s.block_hi, nb), 0, s.n_frames)
            self._ranges_set = True
        self.plan.levels(part="hopsums")
        t = torch.zeros(nb, dtype=torch.float32, device=self.window.device)
        lo, hi = min(s.block_lo, nb), min(s.block_hi, nb)
        if hi > lo:
            self.plan.read_device(self.L.ARR_HOPSUM_F32, t[lo:].data_ptr(), hi - lo, offset=lo)
        return t

    def set_hop_sums(self, h):
        """All hop-block sums of the file -> frame mean squares of every frame."""
        if h.numel():
            self.plan.write_device(self.L.ARR_HOPSUM_F32, h.data_ptr(), h.numel())
        self.plan.levels(part="meansq")

    def set_meansq(self, m):
        import torch
        if m.numel():
            self.plan.write_device(self.L.ARR_MEANSQ_F64 if m.dtype == torch.float64 else self.L.ARR_MEANSQ_F32,
                                   m.data_ptr(), m.numel())

    def set_gate_input(self, lv: np.ndarray):
        self.plan.write(self.L.ARR_GATE_F64, lv)

    # -- gate
    def gate(self, automaton, gate_input, on, off, param, xfade_frames, alpha_init_to_target=False, count_only=False):
        self.plan.gate(automaton, gate_input, on, off, param, xfade_frames, alpha_init_to_target, count_only)

    def c2_count(self) -> int:
        return int(self.plan.read(self.L.ARR_C2_COUNT)[0])

    def states_rows(self):
        return self.plan.read(self.L.ARR_STATE), self.plan.read(self.L.ARR_ROW)

    # -- audio
    def stft(self, post_gain=1.0):
        self.plan.stft(post_gain, skip_edges=True)

    def edge_frames(self, post_gain=1.0, in_scale=None, out_scale=None, pipeline_f64=False):
        self.plan.edge_frames(post_gain, None if in_scale is None else np.array([in_scale], np.float32),
                              None if out_scale is None else np.array([out_scale], np.float32), pipeline_f64)

    def chunk_peaks(self):
        import torch
        t = torch.empty(self.plan.total_chunks, dtype=torch.float32, device=self.window.device)
        if t.numel():
            self.plan.read_device(self.L.ARR_CHUNK_PEAK, t.data_ptr(), t.numel())
        return t

    def set_chunk_peaks(self, p):
        if p.numel():
            self.plan.write_device(self.L.ARR_CHUNK_PEAK, p.data_ptr(), p.numel())

    def limiter(self):
        self.plan.limiter()

    def launches(self) -> int:
        return self.plan.launch_count()

    def close(self):
        self.plan.close()


def _cuda_backend_factory(device_index, unit_blocks=0):
    def make(shard, window, gain_rows, rows_key):
        return CudaShardBackend(shard, window, device_index, gain_rows, rows_key, unit_blocks)
    return make


def _owned(shard: Shard, arr):
    """Zero everything but the frames this rank owns (so that a SUM all-reduce assembles the array exactly).
    Works on NumPy arrays and torch tensors alike."""
    out = np.zeros_like(arr) if isinstance(arr, np.ndarray) else arr.new_zeros(arr.shape)
    out[shard.frame_lo:shard.frame_hi] = arr[shard.frame_lo:shard.frame_hi]
    return out


def _host(a) -> np.ndarray:
    return a if isinstance(a, np.ndarray) else a.cpu().numpy()


# ------------------------------------------------------------------------------------------------ drivers
def run_streaming_sharded(mode: str, own, sr: int, total: int, comm: Comm, make_backend=None, device_index: int = 0,
                          gather_to: Optional[int] = None, unit_blocks: int = 0, **params) -> dict:
    """standard / xfade on one file of `total` samples spread over comm.world ranks.

    own: this rank's samples (tensor [own_hi-own_lo, 2], float32, on the rank's device) for the shard
    plan_shards(total, world, STREAMING)[rank].  Returns dict(out = this rank's output shard (same shape as own),
    full = gathered output on rank gather_to, states/meansq/levels/rows for the whole file, chunk_peaks, shard)."""
    from . import _lib as L
    from .engine import streaming_params
    shards = plan_shards(total, comm.world, STREAMING, params.get("n_fft", tb.N_FFT), params.get("hop", tb.HOP))
    me = shards[comm.rank]
    assert own.shape[0] == me.own_hi - me.own_lo, (own.shape, me)
    sp = streaming_params(mode, sr, **params)
    window = comm.exchange_halos(own, me, shards)                                  # 1. halo hand-off
    be = (make_backend or _cuda_backend_factory(device_index, unit_blocks))(me, window, sp.rows, sp.rows_key)
    try:
        msq = comm.allreduce(_owned(me, be.local_meansq()), "sum")                 # 2. levels -> every rank (on the device)
        be.set_meansq(msq)
        be.gate(L.GATE_UPDELAY, L.ARR_MEANSQ_F32, sp.m_on, sp.m_off, sp.run_frames, sp.xfade_frames)
        be.stft(sp.post_gain)
        be.edge_frames(sp.post_gain)
        peaks = comm.allreduce(be.chunk_peaks(), "max")                            # 3. limiter chunks that straddle ranks
        be.set_chunk_peaks(peaks)
        be.limiter()
        states, rows = be.states_rows()
        full = comm.gather_output(be.out, shards, gather_to) if gather_to is not None else None   # 4.
        msq, peaks = _host(msq), _host(peaks)
        n_fft, hop = params.get("n_fft", tb.N_FFT), params.get("hop", tb.HOP)
        starts = -(n_fft // 2) + hop * np.arange(me.n_frames, dtype=np.int64)
        return dict(out=be.out, full=full, shard=me, shards=shards, meansq=msq, levels=tb.levels_from_meansq(msq),
                    states=states, rows=rows, chunk_peaks=peaks, frame_starts=starts,
                    csv_mask=(starts >= 0) & (starts < total), xfade_frames=sp.xfade_frames, sr=sr,
                    launches=be.launches(), comm_bytes=comm.bytes_sent)
    finally:
        be.close()


def run_adaptive_sharded(own, sr: int, total: int, comm: Comm, make_backend=None, device_index: int = 0,
                         gather_to: Optional[int] = None, unit_blocks: int = 0, fc=1000.0, slope=12.0, c1_low=15.0,
                         c1_high=-15.0, c2_low=-15.0, c2_high=15.0, target_c2=0.5, hyst_db=3.0, min_hold_ms=250.0,
                         xfade_ms=500.0, headroom_margin=2.0, n_fft=tb.N_FFT, hop=tb.HOP) -> dict:
    """adaptive mode on one file spread over comm.world ranks (src/process_tomatis_adaptive.py:157-373)."""
    from . import _lib as L
    shards = plan_shards(total, comm.world, WHOLEFILE, n_fft, hop)
    me = shards[comm.rank]
    assert own.shape[0] == me.own_hi - me.own_lo, (own.shape, me)
    hold, xf = tb.adaptive_frame_counts(sr, min_hold_ms, xfade_ms, hop)
    c1_db, c2_db = tb.tilt_curves_db(sr, n_fft, fc, slope, c1_low, c1_high, c2_low, c2_high)
    rows_tab = tb.gain_rows_adaptive(c1_db, c2_db, xf)
    window = comm.exchange_halos(own, me, shards)
    be = (make_backend or _cuda_backend_factory(device_index, unit_blocks))(
        me, window, rows_tab, ("adaptive", sr, fc, slope, c1_low, c1_high, c2_low, c2_high, xf, n_fft))
    try:
        in_peak = np.float32(comm.allreduce(np.array([be.input_peak()], np.float32), "max")[0])
        atten_db, atten_lin, use_f64 = tb.adaptive_attenuation(in_peak, c1_low, c2_high, headroom_margin)
        scale = np.float32(atten_lin)
        msq = _host(comm.allreduce(_owned(me, be.local_meansq(use_f64, scale)), "sum"))
        levels = tb.levels_from_meansq(msq)
        be.set_gate_input(levels)
        # threshold bisection, identical on every rank (src/process_tomatis_adaptive.py:124-154)
        n = me.n_frames
        valid = levels > -70
        trace = []
        if valid.any():
            vl = levels[valid]
            T_low, T_high, best_T, best_diff = np.percentile(vl, 5), np.percentile(vl, 95), np.median(vl), 1.0
            for _ in range(30):
                T_mid = (T_low + T_high) / 2
                be.gate(L.GATE_MINHOLD, L.ARR_GATE_F64, T_mid + hyst_db / 2, T_mid - hyst_db / 2, hold, xf, True, True)
                ratio = be.c2_count() / n
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
            best_T = np.median(levels)
        be.gate(L.GATE_MINHOLD, L.ARR_GATE_F64, best_T + hyst_db / 2, best_T - hyst_db / 2, hold, xf, True, False)
        be.stft(1.0)
        if use_f64:
            be.edge_frames(1.0, None, None, True)
        else:
            be.edge_frames(1.0, scale, np.float32(tb.db_to_lin_keep(atten_db)), False)
        peaks = comm.allreduce(be.chunk_peaks(), "max")           # one chunk: the global output peak
        be.set_chunk_peaks(peaks)
        be.limiter()
        states, rows = be.states_rows()
        full = comm.gather_output(be.out, shards, gather_to) if gather_to is not None else None
        peaks = _host(peaks)
        return dict(out=be.out, full=full, shard=me, shards=shards, meansq=msq, levels=levels, states=states, rows=rows,
                    optimal_T=float(best_T), trace=trace, atten_db=float(atten_db), input_peak=float(in_peak),
                    pipeline_dtype="float64" if use_f64 else "float32", output_peak=float(peaks[0]) if len(peaks) else 0.0,
                    min_hold_frames=hold, xfade_frames=xf, times=(np.arange(1, n + 1) * (hop / sr)), sr=sr,
                    launches=be.launches(), comm_bytes=comm.bytes_sent)
    finally:
        be.close()


class StreamingShardSession:
    """Persistent per-rank state for repeated passes over one sharded file (what bench.py times): the plan, the input
    window and the output shard are created once; `step()` is one pass of the whole path -- halo hand-off, levels,
    level all-reduce, gate scan, STFT/OLA, edge frames, peak all-reduce, limiter -- with no host synchronisation."""

    def __init__(self, mode: str, own, sr: int, total: int, comm: Comm, device_index: int = 0, unit_blocks: int = 0,
                 use_graph: bool = True, use_peer: Optional[bool] = None, **params):
        import os
        from . import _lib as L
        from .engine import streaming_params
        self.L, self.comm, self.own = L, comm, own
        # graph replay needs a capturable communicator: torch.distributed / NCCL (the in-process stand-ins of the tests are not)
        self.use_graph, self._graph, self._thresholds_set = (use_graph and isinstance(comm, Comm)), None, False
        self.shards = plan_shards(total, comm.world, STREAMING, params.get("n_fft", tb.N_FFT), params.get("hop", tb.HOP))
        self.me = self.shards[comm.rank]
        assert own.shape[0] == self.me.own_hi - self.me.own_lo
        self.sp = streaming_params(mode, sr, **params)
        window = comm.exchange_halos(own, self.me, self.shards)
        self.be = CudaShardBackend(self.me, window, device_index, self.sp.rows, self.sp.rows_key, unit_blocks)
        # from here on the rank's samples live inside the window buffer: refresh them through this view
        self.own = window[self.me.own_lo - self.me.in_lo: self.me.own_hi - self.me.in_lo]
        # per-pass exchange through peer memory (PeerExchange) when there is more than one rank, the communicator is a real
        # process group and every halo is at most one hop from the adjacent rank; TMT_PEER_EXCHANGE=0 keeps the collectives
        self.peer = None
        if use_peer is None:
            use_peer = os.environ.get("TMT_PEER_EXCHANGE", "1") != "0"
        hop = self.me.hop
        simple = all(s.own_hi - s.own_lo >= hop and s.own_lo - s.in_lo <= hop and s.in_hi - s.own_hi <= hop for s in self.shards)
        if use_peer and comm.world > 1 and isinstance(comm, Comm) and comm.device.type == "cuda" and simple:
            px = PeerExchange(comm, self.me, device_index)
            if px.ok:
                self.peer = px

    def _step_eager(self, stft_events=None, marks=None):
        comm, be, me, sp, L = self.comm, self.be, self.me, self.sp, self.L

        def mark(label):
            if marks is not None:
                import torch
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((label, e))
        mark("start")
        if self.peer is not None:
            # hop sums of the blocks this rank owns -> published straight into every rank's exchange buffer together with the
            # edge hops the neighbours need; then one kernel that waits for everybody's flag and unpacks sums and halos
            s = me
            nb = s.n_frames + 1 if s.n_frames > 0 else 0
            if not getattr(be, "_ranges_set", False):
                be.plan.set_level_ranges(0, min(s.block_lo, nb), min(s.block_hi, nb), 0, s.n_frames)
                be._ranges_set = True
            be.plan.levels(part="hopsums")
            mark("levels")
            self.peer.publish(be.plan, self.own)
            mark("peer_publish")
            self.peer.wait_unpack(be.plan, be.window)
            mark("peer_wait")
            be.plan.levels(part="meansq")
            if not self._thresholds_set:
                be.gate(L.GATE_UPDELAY, L.ARR_MEANSQ_F32, sp.m_on, sp.m_off, sp.run_frames, sp.xfade_frames)
                self._thresholds_set = True
            else:
                be.gate(L.GATE_UPDELAY, L.ARR_MEANSQ_F32, None, None, sp.run_frames, sp.xfade_frames)
            mark("gate")
            return self._step_audio(stft_events, mark)
        pending = comm.exchange_halos_begin(self.own, me, self.shards)    # 1. halo hand-off starts (fresh data every pass) ...
        mark("halo_issue")
        hsum = be.own_hop_sums()                                        # ... while the rank sums the hop blocks it owns
        mark("levels")
        if comm.world > 1:
            hsum = comm.allreduce(hsum, "sum")                           # 2. every hop block has one owner: exact gather
        be.set_hop_sums(hsum)
        mark("allreduce_levels")
        if not self._thresholds_set:                                     # thresholds go to the device once, not every pass
            be.gate(L.GATE_UPDELAY, L.ARR_MEANSQ_F32, sp.m_on, sp.m_off, sp.run_frames, sp.xfade_frames)
            self._thresholds_set = True
        else:
            be.gate(L.GATE_UPDELAY, L.ARR_MEANSQ_F32, None, None, sp.run_frames, sp.xfade_frames)
        mark("gate")
        comm.exchange_halos_end(pending, me, be.window)                  # the STFT is the first consumer of the halos
        mark("halo_wait")
        return self._step_audio(stft_events, mark)

    def _step_audio(self, stft_events, mark):
        comm, be, sp = self.comm, self.be, self.sp
        if be.plan.unfusable_chunks == 0:
            # every limiter chunk lies inside this rank's range (shards are cut on chunk boundaries): the per-chunk limiter runs
            # inside the STFT kernel, no peak exchange, no separate pass (src/process_tomatis.py:331-357)
            be.plan.clear_peaks()
            be.edge_frames(sp.post_gain)
            mark("edge")
            if stft_events:
                stft_events[0].record()
            be.plan.stft_limited(sp.post_gain)
            if stft_events:
                stft_events[1].record()
            mark("stft")
            return
        if stft_events:
            stft_events[0].record()
        be.stft(sp.post_gain)
        if stft_events:
            stft_events[1].record()
        mark("stft")
        be.edge_frames(sp.post_gain)
        mark("edge")
        peaks = comm.allreduce(be.chunk_peaks(), "max")                # 3.
        be.set_chunk_peaks(peaks)
        mark("allreduce_peaks")
        be.limiter()
        mark("limiter")

    def step(self, stft_events=None, marks=None):
        """One pass of the whole path over the rank's shard.  marks: optional list that receives (label, torch.cuda.Event)
        after each phase (bench breakdown; runs eagerly).  Otherwise the launch sequence -- kernels and collectives -- is
        captured into a CUDA graph on the first plain call and replayed afterwards: at 8 GPUs a pass lasts about a millisecond
        and the host-side launch of three small collectives was a quarter of it."""
        if marks is not None or stft_events is not None or not self.use_graph:
            return self._step_eager(stft_events, marks)
        import torch
        if self._graph is None:
            self._step_eager()                                           # warm-up outside capture (lazy initialisations)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g):
                    self._step_eager()
                self._graph = g
            except Exception as exc:                                     # capture not possible here: stay eager, say so once
                import warnings
                warnings.warn(f"StreamingShardSession: CUDA graph capture failed ({exc}); running eagerly")
                self.use_graph = False
                torch.cuda.synchronize()
                return self._step_eager()
        self._graph.replay()

    @property
    def out(self):
        return self.be.out

    def close(self):
        if self._graph is not None:          # a live graph keeps the captured NCCL kernels' communicator busy: destroying the
            import torch                     # process group with it still alive hangs
            torch.cuda.synchronize()
            self._graph = None
        if self.peer is not None:
            import torch
            torch.cuda.synchronize()
            self.peer.close()
            self.peer = None
        self.be.close()
