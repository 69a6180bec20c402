"""Dev probe (B200): the device-resident batch step at the two fused frame sizes, and the general-size path on one track.
    python tools/fused2048_probe.py [tracks]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tomatis_audio_processor_b200 import batch, generic, synth     # noqa: E402


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    sr, secs = 44100, 300.0
    base = [synth.recipe_gated_pink(secs, sr, 1000 + i, env_hz=0.2, hi_dbfs=-25.0) for i in range(min(T, 4))]
    x = torch.stack([torch.from_numpy(base[i % len(base)]) for i in range(T)]).cuda()
    y = torch.empty_like(x)
    for n_fft, hop in ((4096, 2048), (2048, 1024)):
        db = batch.DeviceBatch(x, y, sr, "standard", gate_ui=50, n_fft=n_fft, hop=hop)
        step = timed(db.step)
        db.levels(); db.gate()
        tail = timed(lambda: (db.edges(), db.stft()))              # edges() also resets the per-chunk counters of the fused limiter
        edge = timed(db.edges)
        stft = tail - edge
        lv = timed(db.levels)
        sf = T * x.shape[1]
        print(f"fused n_fft={n_fft} hop={hop}: {T} tracks x {secs:.0f} s: step {step:.3f} ms ({T * secs / step * 1e3:.0f} audio-s/s), "
              f"stft with fused limiter {stft:.3f} ms = {16.0 * sf / stft / 1e6 / 6547.2:.4f} of the HBM roofline, levels {lv:.3f} ms, "
              f"units {db.plan.total_units}, peak {float(y.abs().max()):.6f}")
        db.close()
    os.environ["TMT_FUSED_2048"] = "0"
    x1 = base[0]
    import time
    generic.run_streaming("standard", [x1], sr, gate_ui=50, n_fft=2048, hop=1024)
    t0 = time.perf_counter()
    generic.run_streaming("standard", [x1], sr, gate_ui=50, n_fft=2048, hop=1024)
    print(f"general-size path, 1 track, whole call {1e3 * (time.perf_counter() - t0):.1f} ms")


if __name__ == "__main__":
    main()
