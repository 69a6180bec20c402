"""GPU parity against the ORACLE at the BASELINE.json sizes (run on the B200 box: pytest -m gpu).

The short fixtures of test_gpu_parity.py never reach the parts of the path that only exist at length: the limiter's
chunk schedule over dozens of chunks (src/process_tomatis.py:331-357,419-426), the fused last-finisher rescale with
296 CTAs racing over thousands of chunks, the dynamic work queue with jittered unit lengths, the segmented gate scan
(> 16 384 frames), the adaptive bisection on 14 062 frames.  Every case below compares with the NumPy restatement of
the reference on the same samples:

  * mean squares bit-exact, gate states / crossfade counters / chunk lengths / optimal_T / bisection trace exact;
  * PCM: max-abs error <= 1e-5 of full scale against the float64-FFT evaluation of the reference source everywhere,
    and pointwise <= 1e-5 + the reference's own float32-vs-float64 self-noise against its float32-FFT output (the
    rule of test_gpu_parity.py::test_golden_fixture).
"""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PCM_TOL = 1e-5


def _engine():
    from tomatis_audio_processor_b200 import engine
    return engine


def _oracle():
    from oracle import tomatis_oracle as orc
    return orc


def _check_pcm(tag, got, o32, o64):
    """got / o32 / o64: [N, 2] arrays (GPU, float32-FFT oracle, float64-FFT oracle).  Returns the two error figures."""
    got = np.asarray(got, dtype=np.float64)
    d64 = np.abs(got - o64.astype(np.float64)).max(axis=1)
    d32 = np.abs(got - o32.astype(np.float64)).max(axis=1)
    self_noise = np.abs(o32.astype(np.float64) - o64.astype(np.float64)).max(axis=1)
    e64, e32 = float(d64.max()), float(d32.max())
    print(f"{tag}: {len(got)} sample-frames, max-abs error vs float64-FFT oracle {e64:.3e}, vs float32-FFT oracle {e32:.3e} "
          f"(oracle self-noise {float(self_noise.max()):.1e})")
    assert e64 <= PCM_TOL, (tag, e64, int(d64.argmax()))
    assert np.all(d32 <= PCM_TOL + self_noise), (tag, int((d32 - self_noise).argmax()))
    return e64, e32


def _check_streaming(tag, r, x, sr, mode, kw):
    orc = _oracle()
    o = orc.run(mode, x, sr, **kw)
    o64 = orc.run(mode, x, sr, fft_dtype="float64", **kw)
    assert np.array_equal(r["meansq"], np.asarray(o["meansq"])), tag
    assert np.array_equal(r["states"], o["states"]), tag
    assert r["chunk_lengths"] == o["chunk_lengths"], tag
    xe = max(r["xfade_frames"], 1)
    assert np.allclose(r["rows"] / xe, o["alphas"], atol=1e-9), tag
    out = r["out"] if isinstance(r["out"], np.ndarray) else r["out"].cpu().numpy()
    _check_pcm(tag, out, o["out"], o64["out"])
    return o


def test_config0_standard_60s_44k1():
    """BASELINE configs[0]: process_tomatis standard, --gate_ui 50, 60 s @ 44.1 kHz (SURVEY.md 8d recipe C1, seed 1)."""
    from tomatis_audio_processor_b200 import synth
    x = synth.recipe_gated_pink(60.0, 44100, 1)
    r = _engine().run("standard", [x], 44100, gate_ui=50)[0]
    o = _check_streaming("configs[0]", r, x, 44100, "standard", dict(gate_ui=50))
    assert len(o["states"]) == 1292 and len(o["chunk_lengths"]) == 11
    assert (o["states"] == 2).any() and (o["states"] == 1).any()


def test_config3_tracks_inside_a_64_track_batch():
    """BASELINE configs[3] shape: 5 min @ 44.1 kHz tracks processed inside one 64-track plan (fused limiter, dynamic
    work queue and jittered units at scale: 3 520 chunks, ~7 000 work units over 296 CTAs); three tracks against the oracle."""
    import torch
    from tomatis_audio_processor_b200 import synth
    n, sr, T = 13_230_000, 44100, 64
    xb = synth.device_batch(T, n, sr, 1000, "cuda:0")
    rs = _engine().run("standard", [xb[i] for i in range(T)], sr, gate_ui=50, want_host=False)
    over = 0
    for i in (0, 31, 63):
        x = xb[i].cpu().numpy()
        r = dict(rs[i])
        r["out"] = rs[i]["out"].cpu().numpy()
        o = _check_streaming(f"configs[3] track {i} of {T}", r, x, sr, "standard", dict(gate_ui=50))
        assert len(o["chunk_lengths"]) == 55 and o["chunk_lengths"][0] == 239_616
        over += int((r["chunk_peaks"] > 0.999).sum())
    assert over > 0          # the limiter must have engaged somewhere, or this test does not cover the rescale
    del rs, xb
    torch.cuda.empty_cache()


@pytest.mark.parametrize("peak,target", [(0.5, 0.5), (0.1, 0.5), (0.5, 0.35)])
def test_config1_adaptive_10min_48k(peak, target):
    """BASELINE configs[1]: process_tomatis_adaptive on 10 min @ 48 kHz (recipe C2, seed 2); peak 0.5 -> float32 branch with
    pre-attenuation, peak 0.1 -> float64 branch (src/process_tomatis_adaptive.py:201-215)."""
    from tomatis_audio_processor_b200 import synth
    orc = _oracle()
    x = synth.recipe_swept_pink(600.0, 48000, 2, peak=peak)
    kw = dict(target_c2=target)
    r = _engine().run("adaptive", [x], 48000, **kw)[0]
    o = orc.run("adaptive", x, 48000, **kw)
    o64 = orc.run("adaptive", x, 48000, fft_dtype="float64", **kw)
    tag = f"configs[1] peak {peak} target {target}"
    assert r["pipeline_dtype"] == o["pipeline_dtype"] == ("float32" if peak == 0.5 else "float64"), tag
    assert len(o["states"]) == 14062
    assert np.array_equal(r["meansq"], np.asarray(o["meansq"])), tag
    assert r["optimal_T"] == o["optimal_T"] and r["trace"] == o["trace"], tag
    assert np.array_equal(r["states"], o["states"]), tag
    assert np.allclose(r["rows"] / max(r["xfade_frames"], 1), o["alphas"], atol=1e-9), tag
    assert r["chunk_lengths"] == o["chunk_lengths"]
    _check_pcm(tag, r["out"], o["out"], o64["out"])


@pytest.mark.parametrize("gate_ui", [50, 60])
def test_config2_xfade_ramps_120s(gate_ui):
    """BASELINE configs[2]: process_tomatis_xfade, --xfade_ms 500, level ramps straddling the thresholds (recipe C3, seed 3,
    linear gate map: --gate_ui 50 -> T = -50 dBFS; the same file at --gate_ui 60)."""
    from tomatis_audio_processor_b200 import synth
    x = synth.recipe_threshold_ramps(120.0, 48000, 3, t_on=-48.5, t_off=-51.5)
    kw = dict(gate_ui=gate_ui, xfade_ms=500.0)
    r = _engine().run("xfade", [x], 48000, **kw)[0]
    o = _check_streaming(f"configs[2] gate_ui {gate_ui}", r, x, 48000, "xfade", kw)
    assert len(o["states"]) == 2813 and len(o["chunk_lengths"]) == 24
    if gate_ui == 50:
        a = np.asarray(o["alphas"])
        assert ((a > 0) & (a < 1)).any()          # crossfades in progress: the dB-domain mix rows were exercised


def _run_threads(world, fn):
    from thread_comm import ThreadWorld
    tw = ThreadWorld(world)
    res, err = [None] * world, []

    def body(r):
        try:
            res[r] = fn(tw.comm(r), r)
        except BaseException as e:          # noqa: BLE001 - re-raised in the main thread
            err.append(e)
            tw.barrier.abort()
    th = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    if err:
        raise err[0]
    return res


def test_config4_slice_10min_96k_sharded_world4():
    """BASELINE configs[4] shape, 10-minute slice (57 600 000 sample-frames @ 96 kHz, pad_end = 0): four time shards with
    halos and the recomputed gate state (sharded.run_streaming_sharded) against the whole-file oracle."""
    import torch
    from tomatis_audio_processor_b200 import sharded, synth
    total, sr, world = 57_600_000, 96000, 4
    xd = synth.device_long_file_range(0, total, sr, 5000, "cuda:0", segment_seconds=120.0)
    x = xd.cpu().numpy()
    orc = _oracle()
    o = orc.run("standard", x, sr, gate_ui=50)
    o64 = orc.run("standard", x, sr, gate_ui=50, fft_dtype="float64")
    _engine().get_engine(0)
    shards = sharded.plan_shards(total, world, sharded.STREAMING)

    def rank_fn(comm, r):
        me = shards[r]
        out = sharded.run_streaming_sharded("standard", xd[me.own_lo:me.own_hi], sr, total, comm, gate_ui=50)
        return out["out"].cpu().numpy(), out["states"], out["meansq"]
    res = _run_threads(world, rank_fn)
    y = np.concatenate([r[0] for r in res], axis=0)
    assert y.shape == x.shape
    for r in res:
        assert np.array_equal(r[1], o["states"]) and np.array_equal(r[2], np.asarray(o["meansq"]))
    assert len(o["states"]) == 28125
    _check_pcm("configs[4] 10-minute slice, 4 shards", y, o["out"], o64["out"])
