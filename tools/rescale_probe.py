"""GPU box only (dev): time the CTAs spend in the fused limiter's chunk rescale during one bench-like step."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tomatis_audio_processor_b200 import synth
from tomatis_audio_processor_b200.batch import DeviceBatch

T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x = synth.device_batch(T, 13_230_000, 44100, 1000, "cuda:0")
y = torch.empty_like(x)
db = DeviceBatch(x, y, 44100, "standard", gate_ui=50)
for _ in range(3):
    db.step()
torch.cuda.synchronize()
out = (C.c_uint64 * 2)()
db.plan.lib.tmt_plan_debug_counters(db.plan.h, out, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
db.levels(); db.gate(); db.edges()
e0.record(); db.stft(); e1.record()
torch.cuda.synchronize()
db.plan.lib.tmt_plan_debug_counters(db.plan.h, out, 1)
ms = e0.elapsed_time(e1)
print(f"{T} tracks: stft {ms:.3f} ms, {out[1]} rescales, {out[0] / max(1, out[1]) / 1e3:.1f} us per rescale, "
      f"{out[0] / 1e6:.2f} CTA-ms in rescales = {out[0] / 1e6 / (296 * ms) * 100:.1f} % of the CTA time")
