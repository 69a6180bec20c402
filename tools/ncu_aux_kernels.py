"""GPU box only, meant to run under ncu: one adaptive-mode call on 16 tracks x 5 min @ 44.1 kHz, so that the HBM-bound helper
kernels (input_peak_kernel, levels_kernel, limiter_kernel) each launch once on 1.7 GB of audio.

    ncu --set full --clock-control none -k regex:'input_peak_kernel|levels_kernel|limiter_kernel' -c 3 -o out python tools/ncu_aux_kernels.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tomatis_audio_processor_b200 import engine, synth

T, n, sr = 16, 13_230_000, 44100
x = synth.device_batch(T, n, sr, 3000, "cuda:0")
x.mul_(0.5 / float(x.abs().max()))
xs = [x[i] for i in range(T)]
outs = [torch.empty_like(v) for v in xs]
r = engine.run_adaptive(xs, sr, want_host=False, outs=outs)
torch.cuda.synchronize()
print("launches", r[0]["launches"], "limited", [bool(q.get("limited", q.get("scale", 1.0) != 1.0)) for q in r][:4])
