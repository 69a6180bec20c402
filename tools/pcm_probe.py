"""Dev probe (B200): GPU time of one wave of the integer-PCM pipeline, conversion fused with the hop-block sums against the separate
conversion + levels passes (device-resident int16 input, 32 tracks x 5 min @ 44.1 kHz).  python tools/pcm_probe.py [tracks]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tomatis_audio_processor_b200 import _lib as L, synth                                         # noqa: E402
from tomatis_audio_processor_b200.engine import Plan, get_engine, pcm_to_float, streaming_params, whole_track_desc   # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sr, n = 44100, int(300.0 * 44100)
x = synth.device_batch(T, n, sr, 1000, "cuda:0")
raw = (x * 32767.0).round().clamp(-32768, 32767).to(torch.int16)
y = torch.empty_like(x)
eng = get_engine(0)
sp = streaming_params("standard", sr, gate_ui=50)
eng.set_gain_rows(sp.rows, key=sp.rows_key)
plan = Plan(eng, L.FRAMING_STREAMING, [whole_track_desc(x[i], y[i]) for i in range(T)])


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def separate():
    pcm_to_float(raw, L.PCM_S16, x)
    plan.run_streaming(sp.m_on, sp.m_off, sp.run_frames, sp.xfade_frames, sp.post_gain)


def fused():
    plan.run_streaming_pcm(raw, L.PCM_S16, sp.m_on, sp.m_off, sp.run_frames, sp.xfade_frames, sp.post_gain)


t_sep, t_fus = timed(separate), timed(fused)
t_conv = timed(lambda: pcm_to_float(raw, L.PCM_S16, x))
t_lev = timed(lambda: plan.levels(part="hopsums"))
t_pl = timed(lambda: plan.pcm_levels(raw, L.PCM_S16))
sf = T * n
print(f"{T} tracks x 300 s, int16 in: separate conversion + levels + rest {t_sep:.3f} ms | fused {t_fus:.3f} ms ({100 * (1 - t_fus / t_sep):.1f} % less)")
print(f"  conversion alone {t_conv:.3f} ms ({12.0 * sf / t_conv / 1e6:.0f} GB/s), hop sums alone {t_lev:.3f} ms ({8.0 * sf / t_lev / 1e6:.0f} GB/s), "
      f"fused pass {t_pl:.3f} ms ({12.0 * sf / t_pl / 1e6:.0f} GB/s of 4 B read + 8 B written per sample-frame)")
from tomatis_audio_processor_b200.engine import float_to_pcm24          # noqa: E402
out24 = torch.empty((T, n, 6), dtype=torch.uint8, device="cuda")
t_o = timed(lambda: float_to_pcm24(y, out24))
print(f"  float -> PCM_24 of the output: {t_o:.3f} ms ({14.0 * sf / t_o / 1e6:.0f} GB/s of 8 B read + 6 B written per sample-frame)")
