"""GPU box only (dev): a few engine calls of one BASELINE single-file shape, for a launch list under ncu.
usage: one_call.py standard|xfade|adaptive"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tomatis_audio_processor_b200 import engine, synth

mode = sys.argv[1]
n, sr, kw = {"standard": (2_646_000, 44100, dict(gate_ui=50)), "xfade": (5_760_000, 48000, dict(gate_ui=60, xfade_ms=500.0)),
             "adaptive": (28_800_000, 48000, {})}[mode]
x = synth.device_batch(1, n, sr, 100, "cuda:0")[0]
if mode == "adaptive":
    x.mul_(0.5 / float(x.abs().max()))
y = torch.empty_like(x)
for _ in range(3):
    engine.run(mode, [x], sr, want_host=False, outs=[y], **kw)
torch.cuda.synchronize()
print("ok")
