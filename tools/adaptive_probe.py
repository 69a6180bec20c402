"""Dev probe: wall time of engine.run_adaptive (host round trips of the threshold bisection included)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tomatis_audio_processor_b200 import engine, synth

for name, T, n, sr in (("1 x 10 min @ 48 kHz", 1, 28_800_000, 48000), ("32 x 5 min @ 44.1 kHz", 32, 13_230_000, 44100)):
    x = synth.device_batch(T, n, sr, 3000, "cuda:0")
    x.mul_(0.5 / float(x.abs().max()))
    xs = [x[i] for i in range(T)]
    outs = [torch.empty_like(v) for v in xs]
    for _ in range(2):
        r = engine.run_adaptive(xs, sr, want_host=False, outs=outs)
    torch.cuda.synchronize()
    t = time.perf_counter()
    reps = 5
    for _ in range(reps):
        r = engine.run_adaptive(xs, sr, want_host=False, outs=outs)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / reps
    print(f"{name}: {dt * 1e3:.2f} ms per call, {T * n / sr / dt:.3e} audio-s/s, bisection iterations {[len(q['trace']) for q in r][:4]}, launches {r[0]['launches']}")
    del x, xs, outs

if os.environ.get("TMT_PROFILE"):
    import cProfile, pstats
    T, n, sr = 32, 13_230_000, 44100
    x = synth.device_batch(T, n, sr, 3000, "cuda:0")
    x.mul_(0.5 / float(x.abs().max()))
    xs = [x[i] for i in range(T)]
    outs = [torch.empty_like(v) for v in xs]
    engine.run_adaptive(xs, sr, want_host=False, outs=outs)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        engine.run_adaptive(xs, sr, want_host=False, outs=outs)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
