"""TEST INFRASTRUCTURE: NumPy stand-in for sharded.CudaShardBackend, built from the oracle's primitives.

It lets the multi-rank driver (tomatis_audio_processor_b200/sharded.py: shard planning, halo hand-off,
level all-reduce, redundant gate scan, chunk-peak all-reduce, final gather) run on CPU under gloo.  It sees
ONLY the rank's input window and raises if a frame needs a sample the halo exchange did not deliver."""
import numpy as np
import torch

from oracle import tomatis_oracle as orc
from tomatis_audio_processor_b200 import tables as tb
from tomatis_audio_processor_b200.sharded import STREAMING

N_FFT, HOP = 4096, 2048
GATE_UPDELAY, GATE_MINHOLD = 0, 1


class NumpyShardBackend:
    def __init__(self, shard, window, gain_rows, rows_key):
        self.s = shard
        self.win_in = window.numpy()
        assert self.win_in.shape[0] == shard.in_hi - shard.in_lo
        self.rows_tab = np.asarray(gain_rows, dtype=np.float32)
        self.out = torch.zeros((shard.own_hi - shard.own_lo, 2), dtype=torch.float32)
        self.f_lo, self.f_hi = max(0, shard.block_lo - 1), min(shard.n_frames, shard.block_hi)
        self.msq = self.gate_in = None
        self.states = np.zeros(shard.n_frames, np.uint8)
        self.rows = np.zeros(shard.n_frames, np.uint16)
        self.c2 = 0
        self.n_launch = 0
        if shard.framing == STREAMING:
            self.chunks = [(shard.first_start + a * HOP, shard.first_start + b * HOP)
                           for a, b in tb.flush_chunk_blocks(shard.n_frames)]
        else:
            self.chunks = [(0, shard.total)] if shard.n_frames else []
        self.peaks = np.zeros(len(self.chunks), np.float32)

    # ---- input access restricted to the delivered window
    def _frame(self, f, dtype=np.float32):
        s = self.s
        p0 = s.first_start + f * HOP
        fr = np.zeros((N_FFT, 2), dtype=dtype)
        lo, hi = max(p0, 0), min(p0 + N_FFT, s.total)
        if hi > lo:
            assert lo >= s.in_lo and hi <= s.in_hi, f"frame {f} needs [{lo},{hi}) outside the window [{s.in_lo},{s.in_hi})"
            fr[lo - p0:hi - p0] = self.win_in[lo - s.in_lo:hi - s.in_lo]
        return fr

    # ---- levels
    def input_peak(self):
        return np.float32(np.max(np.abs(self.win_in))) if self.win_in.size else np.float32(0)

    def local_meansq(self, use_f64=False, in_scale=None):
        out = np.zeros(self.s.n_frames, np.float64 if use_f64 else np.float32)
        for f in range(self.f_lo, self.f_hi):
            fr = self._frame(f)
            if in_scale is not None:
                fr = fr * (np.float64(in_scale) if use_f64 else np.float32(in_scale))   # x * atten_lin (_adaptive.py:215)
            elif use_f64:
                fr = fr.astype(np.float64)
            out[f] = orc.frame_meansq(fr)
        self.n_launch += 2
        return out

    def set_meansq(self, m):
        self.msq = np.asarray(m)

    def set_gate_input(self, lv):
        self.gate_in = np.asarray(lv, np.float64)

    # ---- gate
    def gate(self, automaton, gate_input, on, off, param, xfade_frames, alpha_init_to_target=False, count_only=False):
        v = self.gate_in if gate_input == 2 else self.msq
        v = v.astype(np.float32 if v.dtype == np.float32 else np.float64)
        on = v.dtype.type(on)
        off = v.dtype.type(off)
        n = len(v)
        st = np.ones(n, np.uint8)
        if automaton == GATE_UPDELAY:
            state, run = 1, 0
            for i in range(n):
                if state == 1:
                    run = run + 1 if v[i] >= on else 0
                    if run >= param:
                        state, run = 2, 0
                elif v[i] <= off:
                    state, run = 1, 0
                st[i] = state
        else:
            state, since = 1, param
            for i in range(n):
                since += 1
                if since >= param:
                    if state == 1 and v[i] >= on:
                        state, since = 2, 0
                    elif state == 2 and v[i] <= off:
                        state, since = 1, 0
                st[i] = state
        self.c2 = int((st == 2).sum())
        self.n_launch += 1
        if count_only:
            return
        xe = max(xfade_frames, 1)
        k, rows = 0, np.zeros(n, np.uint16)
        for i in range(n):
            t2 = int(st[i] == 2)
            k = t2 * xe if (alpha_init_to_target and i == 0) else min(max(k + (1 if t2 else -1), 0), xe)
            rows[i] = k
        self.states, self.rows = st, rows

    def c2_count(self):
        return self.c2

    def states_rows(self):
        return self.states, self.rows

    # ---- audio: deferred to edge_frames(), which knows the adaptive scales
    def stft(self, post_gain=1.0):
        self.post_gain = post_gain

    def edge_frames(self, post_gain=1.0, in_scale=None, out_scale=None, pipeline_f64=False):
        s = self.s
        streaming = s.framing == STREAMING
        dt = np.float64 if pipeline_f64 else np.float32
        win = np.hanning(N_FFT).astype(np.float32)
        span0 = s.first_start + self.f_lo * HOP
        n_local = (self.f_hi - self.f_lo - 1) * HOP + N_FFT if self.f_hi > self.f_lo else 0
        acc = np.zeros((n_local, 2), dtype=dt)
        nrm = np.zeros(n_local, np.float32)
        for f in range(self.f_lo, self.f_hi):
            fr = self._frame(f)
            if in_scale is not None:
                fr = fr * np.float32(in_scale)
            elif pipeline_f64:
                fr = fr.astype(np.float64)
            gain = self.rows_tab[self.rows[f]]
            y = np.zeros_like(fr)
            for c in range(2):
                X = np.fft.rfft(fr[:, c] * win)
                X *= gain
                y[:, c] = (np.fft.irfft(X, n=N_FFT).astype(np.float32) * win) if streaming else (np.fft.irfft(X, N_FFT) * win)
            o = (f - self.f_lo) * HOP
            if streaming:
                acc[o:o + N_FFT] += y
                nrm[o:o + N_FFT] += (win * win).astype(np.float32)
            else:                                  # clipped writes (src/process_tomatis_adaptive.py:316-323)
                p0 = s.first_start + f * HOP
                e = min(s.total, p0 + N_FFT) - p0
                acc[o:o + e] += y[:e]
                nrm[o:o + e] += win[:e] ** 2
        lo, hi = s.own_lo, s.own_hi
        a, b = lo - span0, hi - span0
        if hi > lo:
            if streaming:
                y = acc[a:b] / (nrm[a:b, None] + orc.EPS)
                if post_gain != 1.0:
                    y = y * post_gain
            else:
                y = acc[a:b] / np.maximum(nrm[a:b], 1e-8)[:, None].astype(np.float32)
                if out_scale is not None:
                    y = y * np.float32(out_scale)
            self.y = y
        else:
            self.y = np.zeros((0, 2), dt)
        for c, (c0, c1) in enumerate(self.chunks):
            u0, u1 = max(c0, lo), min(c1, hi)
            self.peaks[c] = np.max(np.abs(self.y[u0 - lo:u1 - lo])) if u1 > u0 else 0.0
        self.n_launch += 2

    def chunk_peaks(self):
        return self.peaks.copy()

    def set_chunk_peaks(self, p):
        self.peaks = np.asarray(p, np.float32)

    def limiter(self):
        lo, hi = self.s.own_lo, self.s.own_hi
        y = self.y
        for c, (c0, c1) in enumerate(self.chunks):
            u0, u1 = max(c0, lo), min(c1, hi)
            peak = self.peaks[c]
            if u1 > u0 and peak > orc.PEAK_LIMIT:
                y[u0 - lo:u1 - lo] = y[u0 - lo:u1 - lo] * (orc.PEAK_LIMIT / (peak if y.dtype == np.float32 else np.float64(peak)))
        self.out = torch.from_numpy(np.ascontiguousarray(y.astype(np.float32)))
        self.n_launch += 1

    def launches(self):
        return self.n_launch

    def close(self):
        pass
