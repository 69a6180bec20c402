"""GPU box only (dev): cProfile of one adaptive engine call on a 10-minute track (BASELINE configs[1] shape)."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tomatis_audio_processor_b200 import engine, synth

x = synth.device_batch(1, 28_800_000, 48000, 200, "cuda:0")[0]
x.mul_(0.5 / float(x.abs().max()))
y = torch.empty_like(x)
for _ in range(3):
    engine.run_adaptive([x], 48000, want_host=False, outs=[y])
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    engine.run_adaptive([x], 48000, want_host=False, outs=[y])
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
