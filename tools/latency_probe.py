"""GPU box only: where the wall time of a single-track engine call goes (BASELINE configs[0] / configs[1] / configs[2] shapes).
Host phases are timed with perf_counter and a device synchronize after each, so the figures are upper bounds of what each phase
costs in the un-instrumented call (printed beside it)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tomatis_audio_processor_b200 import engine, synth, _lib as L, tables as tb


def tick(label, acc, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    acc[label] = acc.get(label, 0.0) + (t1 - t0)
    return time.perf_counter()


def probe_streaming(mode, n, sr, reps=20, **kw):
    x = synth.device_batch(1, n, sr, 100, "cuda:0")[0]
    y = torch.empty_like(x)
    for _ in range(3):
        engine.run_streaming(mode, [x], sr, want_host=False, outs=[y], **kw)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        engine.run_streaming(mode, [x], sr, want_host=False, outs=[y], **kw)
    torch.cuda.synchronize()
    whole = (time.perf_counter() - t) / reps
    acc = {}
    eng = engine.get_engine(0)
    for _ in range(reps):
        t0 = time.perf_counter()
        sp = engine.streaming_params(mode, sr, **kw)
        eng.set_gain_rows(sp.rows, key=sp.rows_key)
        t0 = tick("params+rows", acc, t0)
        plan = engine.Plan(eng, L.FRAMING_STREAMING, [engine.whole_track_desc(x, y)], 0)
        t0 = tick("plan_create", acc, t0)
        plan.run_streaming(sp.m_on, sp.m_off, sp.run_frames, sp.xfade_frames, sp.post_gain)
        t0 = tick("launches+kernels", acc, t0)
        msq = plan.read(L.ARR_MEANSQ_F32); st = plan.read(L.ARR_STATE); rows = plan.read(L.ARR_ROW); pk = plan.read(L.ARR_CHUNK_PEAK)
        t0 = tick("readbacks", acc, t0)
        tb.levels_from_meansq(msq); plan.chunk_ranges(0)
        t0 = tick("host_post", acc, t0)
        plan.close()
        t0 = tick("plan_close", acc, t0)
    print(f"{mode} {n / sr:.0f} s @ {sr}: whole call {whole * 1e3:.3f} ms | " + ", ".join(f"{k} {v / reps * 1e3:.3f}" for k, v in acc.items()), flush=True)


def probe_adaptive(n, sr, reps=10):
    x = synth.device_batch(1, n, sr, 200, "cuda:0")[0]
    x.mul_(0.5 / float(x.abs().max()))
    y = torch.empty_like(x)
    for _ in range(3):
        r = engine.run_adaptive([x], sr, want_host=False, outs=[y])
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        r = engine.run_adaptive([x], sr, want_host=False, outs=[y])
    torch.cuda.synchronize()
    print(f"adaptive {n / sr:.0f} s @ {sr}: whole call {(time.perf_counter() - t) / reps * 1e3:.3f} ms, {len(r[0]['trace'])} search steps, "
          f"{r[0]['launches']} launches", flush=True)


if __name__ == "__main__":
    probe_streaming("standard", 2_646_000, 44100, gate_ui=50)
    probe_streaming("xfade", 5_760_000, 48000, gate_ui=60, xfade_ms=500.0)
    probe_adaptive(28_800_000, 48000)
