"""GPU box only: device-resident throughput of every front end on the BASELINE config shapes that are not the bench line
(the bench times configs[3], standard mode).  Wall time of the whole engine call (host bookkeeping, plan creation and the
small read-backs included), inputs and outputs resident in HBM, best of 5 after 2 warm-up calls.  One JSON line per case."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tomatis_audio_processor_b200 import engine, synth


def best_of(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best


def batch(T, n, sr, seed, peak=None):
    x = synth.device_batch(T, n, sr, seed, "cuda:0")
    if peak is not None:
        x.mul_(peak / float(x.abs().max()))
    xs = [x[i] for i in range(T)]
    return x, xs, [torch.empty_like(v) for v in xs]


def report(name, T, n, sr, dt, extra=None):
    print(json.dumps(dict(case=name, tracks=T, seconds_per_track=round(n / sr, 2), sample_rate=sr, ms_per_call=round(dt * 1e3, 3),
                          audio_s_per_s=round(T * n / sr / dt, 1), algorithmic_GBps=round(16.0 * T * n / dt / 1e9, 1), **(extra or {}))), flush=True)


def main():
    # configs[0]: standard, one 60 s 44.1 kHz track (latency-bound: one track does not fill the GPU)
    for T in (1, 64):
        x, xs, outs = batch(T, 2_646_000, 44100, 100)
        dt = best_of(lambda: engine.run_streaming("standard", xs, 44100, want_host=False, outs=outs, gate_ui=50))
        report("configs[0] standard 60 s @ 44.1 kHz", T, 2_646_000, 44100, dt)
        del x, xs, outs
    # configs[1]: adaptive, 10 min 48 kHz (threshold bisection = up to 30 dependent gate scans with a host round trip each)
    for T in (1, 16):
        x, xs, outs = batch(T, 28_800_000, 48000, 200, peak=0.5)
        r = [None]
        dt = best_of(lambda: r.__setitem__(0, engine.run_adaptive(xs, 48000, want_host=False, outs=outs)))
        report("configs[1] adaptive 10 min @ 48 kHz", T, 28_800_000, 48000, dt,
               dict(bisection_iterations=len(r[0][0]["trace"]), launches=r[0][0]["launches"]))
        del x, xs, outs
    # configs[2]: xfade with 500 ms crossfades, 120 s 48 kHz, linear gate map
    for T in (1, 64):
        x, xs, outs = batch(T, 5_760_000, 48000, 300)
        dt = best_of(lambda: engine.run_streaming("xfade", xs, 48000, want_host=False, outs=outs, gate_ui=60, xfade_ms=500.0))
        report("configs[2] xfade 120 s @ 48 kHz, xfade_ms 500", T, 5_760_000, 48000, dt)
        del x, xs, outs
    # N1 static EQ, 32 tracks x 5 min 44.1 kHz, gain-protect pass active
    T, n, sr = 32, 13_230_000, 44100
    x, xs, _ = batch(T, n, sr, 400)
    f = np.fft.rfftfreq(4096, 1 / sr)
    gain = (10 ** (np.interp(np.log10(np.maximum(f, 1.0)), np.log10([20, 1000, 20000]), [6.0, 0.0, 9.0]) / 20)).astype(np.float32)
    dt = best_of(lambda: engine.run_eq(xs, sr, gain, want_host=False), reps=3, warm=1)
    report("N1 static EQ 5 min @ 44.1 kHz (layer2_apply_eq, padded, gain protect)", T, n, sr, dt)


if __name__ == "__main__":
    main()
