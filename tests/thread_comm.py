"""TEST INFRASTRUCTURE: in-process emulation of sharded.Comm -- N "ranks" are N Python threads of one process
sharing one GPU (the GPU box of the parity run has a single device; kernels of different ranks never wait on
one another, only the host threads do).  Same interface as sharded.Comm."""
import threading

import numpy as np


class ThreadWorld:
    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world
        self.lock = threading.Lock()

    def comm(self, rank):
        return ThreadComm(self, rank)


class ThreadComm:
    def __init__(self, w, rank):
        self.w, self.rank, self.world, self.bytes_sent = w, rank, w.world, 0

    def _all(self, value):
        self.w.slots[self.rank] = value
        self.w.barrier.wait()
        vals = list(self.w.slots)
        self.w.barrier.wait()
        return vals

    def allreduce(self, a, op):
        if isinstance(a, np.ndarray):
            vals = self._all(np.array(a, copy=True))
            self.bytes_sent += a.nbytes
            out = vals[0].copy()
            for v in vals[1:]:
                out = out + v if op == "sum" else np.maximum(out, v)
            return out.astype(a.dtype)
        import torch
        torch.cuda.synchronize()
        vals = self._all(a.clone())
        self.bytes_sent += a.numel() * a.element_size()
        out = vals[0].clone()
        for v in vals[1:]:
            out = out + v if op == "sum" else torch.maximum(out, v)
        return out

    def exchange_halos(self, own, shard, shards):
        import torch
        owns = self._all(own)
        win = torch.zeros((shard.in_hi - shard.in_lo, 2), dtype=own.dtype, device=own.device)
        for other in shards:
            lo, hi = max(shard.in_lo, other.own_lo), min(shard.in_hi, other.own_hi)
            if hi > lo:
                win[lo - shard.in_lo:hi - shard.in_lo] = owns[other.rank][lo - other.own_lo:hi - other.own_lo]
                if other.rank != shard.rank:
                    self.bytes_sent += (hi - lo) * 8
        torch.cuda.synchronize()
        self.w.barrier.wait()
        return win

    def exchange_halos_begin(self, own, shard, shards):
        import torch
        torch.cuda.synchronize()
        owns = self._all(own)
        recvs = []
        for other in shards:
            if other.rank == shard.rank:
                continue
            lo, hi = max(shard.in_lo, other.own_lo), min(shard.in_hi, other.own_hi)
            if hi > lo:
                recvs.append((lo, hi, owns[other.rank][lo - other.own_lo:hi - other.own_lo].clone()))
                self.bytes_sent += (hi - lo) * 8
        torch.cuda.synchronize()
        self.w.barrier.wait()
        return "p2p", [], recvs

    def exchange_halos_end(self, pending, shard, window):
        for lo, hi, buf in pending[2]:
            window[lo - shard.in_lo:hi - shard.in_lo] = buf

    def gather_output(self, own_out, shards, dst=0):
        import torch
        outs = self._all(own_out)
        if self.rank != dst:
            return None
        return torch.cat([outs[s.rank] for s in shards], dim=0)
