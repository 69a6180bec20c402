"""Per-kernel timing probe (CUDA events on the launching stream).  Dev tool, not the bench."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tomatis_audio_processor_b200 import _lib as L, engine as E, synth, tables as tb


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), ts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=16)
    ap.add_argument("--seconds", type=float, default=300.0)
    ap.add_argument("--sr", type=int, default=44100)
    ap.add_argument("--unit_blocks", type=int, default=0)
    a = ap.parse_args()
    n = int(a.seconds * a.sr)
    t0 = time.time()
    x = synth.device_batch(a.tracks, n, a.sr, 1000, "cuda:0")
    y = torch.empty_like(x)
    torch.cuda.synchronize()
    print(f"synth {time.time()-t0:.1f}s  tracks={a.tracks} n={n}")
    eng = E.get_engine(0)
    sp = E.streaming_params("standard", a.sr, gate_ui=50)
    eng.set_gain_rows(sp.rows, key=sp.rows_key)
    plan = E.Plan(eng, L.FRAMING_STREAMING, [E.whole_track_desc(x[i], y[i]) for i in range(a.tracks)], a.unit_blocks)
    sf = a.tracks * n
    print(f"frames={plan.total_frames} chunks={plan.total_chunks} units={plan.total_units}")
    for name, fn, bytes_per_sf in [
        ("levels", lambda: plan.levels(False, None), 8),
        ("gate", lambda: plan.gate(L.GATE_UPDELAY, L.ARR_MEANSQ_F32, sp.m_on, sp.m_off, sp.run_frames, 0), 0),
        ("stft", lambda: plan.stft(1.0), 16),
        ("limiter", lambda: plan.limiter(), 16),
        ("stft+lim", lambda: (plan.clear_peaks(), plan.edge_frames(1.0), plan.stft_limited(1.0)), 16),
        ("stft+lim none", lambda: (plan.clear_peaks(), plan.edge_frames(1.0), plan.stft_limited(1.0, 1e9)), 16),
        ("stft+lim all", lambda: (plan.clear_peaks(), plan.edge_frames(1.0), plan.stft_limited(1.0, 1e-9)), 16),
        ("step", lambda: plan.run_streaming(sp.m_on, sp.m_off, sp.run_frames, 0), 16),
    ]:
        ms, ts = ev_time(fn)
        gbs = bytes_per_sf * sf / (ms * 1e-3) / 1e9
        print(f"{name:8s} {ms:9.3f} ms  {sf/(ms*1e-3)/1e9:8.3f} Gsf/s  {gbs:8.1f} GB/s(alg)  audio-s/s={sf/a.sr/(ms*1e-3):.3e}  all={['%.3f'%t for t in ts]}")
    st = plan.read(L.ARR_STATE)
    pk = plan.read(L.ARR_CHUNK_PEAK)
    print("C2 frac", float((st == 2).mean()), "chunks over limit", float((pk > 0.999).mean()), "max peak", float(pk.max()))


if __name__ == "__main__":
    main()
