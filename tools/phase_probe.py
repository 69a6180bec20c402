"""Dev probe (B200 box): where the time of two latency-bound calls goes.

  1. engine.run_adaptive on one 10-minute 48 kHz track (BASELINE configs[1]): wall time per Plan method, host time between them.
  2. generic.run_streaming (--n_fft 2048 --hop 1024 and others) on a 5-minute 48 kHz track: wall time per CudaKernels method.

Every wrapped call is followed by a device synchronise, so the figures are per-phase costs, not the pipelined total (printed too).
"""
import os
import sys
import time
from collections import OrderedDict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tomatis_audio_processor_b200 import engine, generic, synth


def wrap(cls, names, acc):
    saved = {}
    for n in names:
        f = getattr(cls, n)
        saved[n] = f

        def g(self, *a, __f=f, __n=n, **k):
            torch.cuda.synchronize()
            t = time.perf_counter()
            r = __f(self, *a, **k)
            torch.cuda.synchronize()
            acc[__n] = acc.get(__n, 0.0) + time.perf_counter() - t
            return r
        setattr(cls, n, g)
    return saved


def unwrap(cls, saved):
    for n, f in saved.items():
        setattr(cls, n, f)


def adaptive():
    n, sr = 28_800_000, 48000
    x = synth.device_batch(1, n, sr, 3000, "cuda:0")
    x.mul_(0.5 / float(x.abs().max()))
    xs, outs = [x[0]], [torch.empty_like(x[0])]
    for _ in range(3):
        engine.run_adaptive(xs, sr, want_host=False, outs=outs)
    torch.cuda.synchronize()
    t = time.perf_counter()
    reps = 10
    for _ in range(reps):
        engine.run_adaptive(xs, sr, want_host=False, outs=outs)
    torch.cuda.synchronize()
    whole = (time.perf_counter() - t) / reps
    acc = OrderedDict()
    names = ["__init__", "input_peaks", "read", "levels", "write", "bisect", "gate", "stft", "edge_frames", "limiter", "read_many", "close"]
    saved = wrap(engine.Plan, names, acc)
    t = time.perf_counter()
    for _ in range(reps):
        engine.run_adaptive(xs, sr, want_host=False, outs=outs)
    torch.cuda.synchronize()
    serial = (time.perf_counter() - t) / reps
    unwrap(engine.Plan, saved)
    parts = ", ".join(f"{k} {v / reps * 1e3:.3f}" for k, v in acc.items())
    if os.environ.get("TMT_PROFILE"):
        import cProfile, pstats
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(50):
            engine.run_adaptive(xs, sr, want_host=False, outs=outs)
        torch.cuda.synchronize()
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(28)
    print(f"adaptive 600 s @ 48 kHz: whole call {whole * 1e3:.3f} ms; with a synchronise after every phase {serial * 1e3:.3f} ms: {parts}; "
          f"host outside the plan methods {(serial - sum(acc.values()) / reps) * 1e3:.3f}")


def general(n_fft, hop):
    sr = 48000
    x = synth.recipe_gated_pink(300.0, sr, 5, env_hz=0.5)
    for _ in range(2):
        generic.run_streaming("standard", [x], sr, gate_ui=50, n_fft=n_fft, hop=hop)
    torch.cuda.synchronize()
    t = time.perf_counter()
    reps = 3
    for _ in range(reps):
        r = generic.run_streaming("standard", [x], sr, gate_ui=50, n_fft=n_fft, hop=hop)
    torch.cuda.synchronize()
    whole = (time.perf_counter() - t) / reps
    acc = OrderedDict()
    saved = wrap(generic.CudaKernels, ["upload", "upload_tables", "meansq", "gate", "frames", "overlap_add", "limit", "to_host"], acc)
    for _ in range(reps):
        generic.run_streaming("standard", [x], sr, gate_ui=50, n_fft=n_fft, hop=hop)
    unwrap(generic.CudaKernels, saved)
    parts = ", ".join(f"{k} {v / reps * 1e3:.2f}" for k, v in acc.items())
    print(f"general path standard 300 s @ 48 kHz n_fft={n_fft} hop={hop}: whole call {whole * 1e3:.2f} ms ({len(r[0]['states'])} frames): {parts}")


if __name__ == "__main__":
    adaptive()
    if "--general" in sys.argv:
        for nf, hp in ((2048, 1024), (1024, 512), (4096, 1024)):
            general(nf, hp)
