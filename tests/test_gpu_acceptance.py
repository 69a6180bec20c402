"""GPU: secondary end-to-end acceptance in the reference's own terms: the criteria of its validator
`src/validate_layer1.py` restated compactly -- (A) same length, finite; (B) gate re-simulation from the input agrees with the
state CSV; (D) conditional spectrum |Y|/|X| of stable C1 / C2 frames follows the theoretical tilt curves, RMSE < 1.5 dB in
100-800 Hz, 800-1200 Hz and 2000-8000 Hz (validate_layer1.py:261-389, 568-589).  Much looser than the parity tests; it
checks that the output is the *intended* filter, independently of the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N_FFT, HOP = 4096, 2048


def _stable(states, margin=2):
    st = np.asarray(states)
    ok = np.ones(len(st), bool)
    for d in range(1, margin + 1):
        ok[d:] &= st[d:] == st[:-d]
        ok[:-d] &= st[:-d] == st[d:]
    ok[:margin] = ok[-margin:] = False
    return ok


def test_conditional_spectrum_follows_the_tilt_curves():
    from tomatis_audio_processor_b200 import engine, synth, tables as tb
    sr = 48000
    x = synth.recipe_gated_pink(40.0, sr, 90, lo_dbfs=-62.0, hi_dbfs=-36.0, env_hz=0.25, bursts=False)   # quiet: limiter idle
    r = engine.run("standard", [x], sr, gate_ui=50)[0]
    y = r["out"]
    assert y.shape == x.shape and np.isfinite(y).all() and float(np.abs(y).max()) < 0.98          # validator check A
    assert float(r["chunk_peaks"].max()) <= 0.999                                                 # limiter did not touch it
    # B: independent re-simulation of the gate from the levels of the INPUT (hysteresis + up-delay, frame grid of the CSV)
    starts, mask = r["frame_starts"], r["csv_mask"]
    state, run, sim = 1, 0, []
    need = tb.updelay_run_frames(sr, 250.0)
    for k in range(len(starts)):
        s0 = int(starts[k])
        fr = np.zeros((N_FFT, 2), np.float32)
        a, b = max(s0, 0), min(s0 + N_FFT, len(x))
        fr[a - s0:b - s0] = x[a:b]
        lvl = 20 * np.log10(np.sqrt(np.mean(np.mean(fr.astype(np.float64) ** 2, axis=1)) + 1e-12) + 1e-12)
        if state == 1:
            run = run + 1 if lvl >= r["Ton"] else 0
            if run >= need:
                state, run = 2, 0
        elif lvl <= r["Toff"]:
            state, run = 1, 0
        sim.append(state)
    mismatch = float(np.mean(np.array(sim)[mask] != r["states"][mask]))
    assert mismatch < 0.01, mismatch
    # D: conditional spectrum of stable frames against the theoretical C1 / C2 curves
    freqs = np.fft.rfftfreq(N_FFT, 1 / sr)
    g1_db, g2_db = tb.tilt_curves_db(sr, N_FFT, 1000.0, 12.0, 15.0, -15.0, -15.0, 15.0)
    win = np.hanning(N_FFT)
    stable = _stable(r["states"])
    ratios = {1: [], 2: []}
    for k in np.nonzero(stable & mask)[0]:
        s0 = int(starts[k])
        if s0 < 0 or s0 + N_FFT > len(x) or r["levels"][k] < -60:
            continue
        X = sum(np.abs(np.fft.rfft(x[s0:s0 + N_FFT, c] * win)) for c in range(2)) / 2
        Y = sum(np.abs(np.fft.rfft(y[s0:s0 + N_FFT, c] * win)) for c in range(2)) / 2
        ratios[int(r["states"][k])].append(Y / np.maximum(X, 1e-10))
    assert len(ratios[1]) > 20 and len(ratios[2]) > 20
    for st, theory in ((1, g1_db), (2, g2_db)):
        meas = 20 * np.log10(np.median(np.array(ratios[st]), axis=0) + 1e-12)
        for lo, hi in ((100, 800), (800, 1200), (2000, 8000)):
            m = (freqs >= lo) & (freqs <= hi)
            rmse = float(np.sqrt(np.mean((meas[m] - theory[m]) ** 2)))
            assert rmse < 1.5, (st, lo, hi, rmse)
