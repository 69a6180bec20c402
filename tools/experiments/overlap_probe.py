"""Dev probe: does levels_slim_kernel run beside stft_kernel?  Times STFT alone, slim levels alone, both on two streams."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tomatis_audio_processor_b200 import _lib as L, engine as E, synth

n, sr, T = 13_230_000, 44100, 32
x = synth.device_batch(2 * T, n, sr, 1000, "cuda:0")
y = torch.empty_like(x)
eng = E.get_engine(0)
sp = E.streaming_params("standard", sr, gate_ui=50)
eng.set_gain_rows(sp.rows, key=sp.rows_key)
A = E.Plan(eng, L.FRAMING_STREAMING, [E.whole_track_desc(x[i], y[i]) for i in range(T)])
B = E.Plan(eng, L.FRAMING_STREAMING, [E.whole_track_desc(x[i], y[i]) for i in range(T, 2 * T)])
for p in (A, B):
    p.levels(); p.gate(L.GATE_UPDELAY, L.ARR_MEANSQ_F32, sp.m_on, sp.m_off, sp.run_frames, 0); p.clear_peaks(); p.edge_frames(1.0)
side = torch.cuda.Stream()
torch.cuda.synchronize()

def timed(fn, reps=4):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]

def both():
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    A.clear_peaks(); A.stft_limited(1.0)
    with torch.cuda.stream(side):
        B.levels(slim=True)
    main.wait_stream(side)

print("stft(A) alone      %.3f ms" % timed(lambda: (A.clear_peaks(), A.stft_limited(1.0))))
print("levels(B) full     %.3f ms" % timed(lambda: B.levels()))
print("levels(B) slim     %.3f ms" % timed(lambda: B.levels(slim=True)))
print("stft(A) || slim(B) %.3f ms" % timed(both))
