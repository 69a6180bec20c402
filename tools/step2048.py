"""Profiling target: a few device-resident batch steps at --n_fft 2048 --hop 1024 (python tools/step2048.py [tracks] [steps])."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tomatis_audio_processor_b200 import batch, synth     # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sr = 44100
x = synth.device_batch(T, int(300.0 * sr), sr, 1000, "cuda:0")
y = torch.empty_like(x)
db = batch.DeviceBatch(x, y, sr, "standard", gate_ui=50, n_fft=2048, hop=1024)
for _ in range(steps):
    db.step()
torch.cuda.synchronize()
print("peak", float(y.abs().max()))
