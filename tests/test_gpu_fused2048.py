"""`--n_fft 2048 --hop 1024` (the reference documentation's faster setting, docs/Tomatis技术说明.md:253-258) on the FUSED kernels:
the second build of the library (csrc/libtomatis_b200_n2048.so, TMT_NFFT=2048), whose STFT kernel carries two consecutive
2048-point frames through one 4096-wide pass ("pair mode", csrc/fft4096.cuh).  Everything is compared with the oracle run at the
same sizes (which tests/test_oracle_vs_reference.py pins to the executed reference for non-default n_fft / hop): mean squares
bit-exact, states / rows / chunk lengths / threshold search exact, PCM within 1e-5 of full scale."""
import os

import numpy as np
import pytest

from oracle import tomatis_oracle as orc
from tomatis_audio_processor_b200 import build, synth
from test_generic_sizes import _compare

pytestmark = pytest.mark.gpu
SZ = dict(n_fft=2048, hop=1024)


def _q(x):
    return synth.pcm16_to_float(synth.quantise_pcm16(x))


def _engine():
    from tomatis_audio_processor_b200 import engine
    return engine


def test_second_library_is_loaded_for_2048():
    eng = _engine()
    e = eng.get_engine(0, 2048, 1024)
    assert e.n_fft == 2048 and e.hop == 1024 and e.lib.n_fft == 2048
    assert eng.get_engine(0).lib.n_fft == 4096 and eng.get_engine(0) is not e
    assert build.LIB_PATHS[2048].endswith("libtomatis_b200_n2048.so")


@pytest.mark.parametrize("sr", [48000, 44100])
def test_three_modes_match_oracle(sr):
    eng = _engine()
    xs = _q(synth.recipe_gated_pink(6.5, sr, 190, env_hz=1.3, hi_dbfs=-20.0))          # two limiter chunks at hop 1024 (flush after 240 000 samples)
    for mode, kw in (("standard", dict(gate_ui=50, up_delay_ms=120.0, output_gain_db=-1.5)),
                     ("xfade", dict(gate_ui=62, xfade_ms=150.0, up_delay_ms=60.0)), ("xfade", dict(gate_ui=55))):
        r = eng.run(mode, [xs], sr, **SZ, **kw)[0]
        _compare(mode, orc.run(mode, xs, sr, **SZ, **kw), orc.run(mode, xs, sr, fft_dtype="float64", **SZ, **kw), r)
    for peak in (0.5, 0.08):                                                            # float32 and float64 pipelines
        xa = _q(synth.recipe_swept_pink(5.0, sr, 191, period_s=0.8, peak=peak))
        kw = dict(min_hold_ms=100.0, xfade_ms=200.0, **SZ)
        r = eng.run("adaptive", [xa], sr, **kw)[0]
        _compare("adaptive", orc.run("adaptive", xa, sr, **kw), orc.run("adaptive", xa, sr, fft_dtype="float64", **kw), r)


def test_ragged_lengths_and_short_files():
    """Every parity of frame count and unit length; files shorter than a frame, a hop, one sample; pad_end = 0."""
    eng = _engine()
    sr = 48000
    rng = np.random.default_rng(5)
    totals = [1, 700, 1024, 1500, 2047, 2048, 2049, 3072, 4096 + 17, 1024 * 9, 1024 * 10 + 1, 1024 * 40, 33333, 100001]
    for k, total in enumerate(totals):
        env = 10.0 ** (rng.uniform(-3.0, -0.6, size=(total // 700 + 1)).repeat(700)[:total, None])
        x = _q((env * rng.standard_normal((total, 2))).astype(np.float32).clip(-1, 1))
        mode = ("standard", "xfade", "adaptive")[k % 3]
        if mode == "adaptive":
            kw = dict(min_hold_ms=20.0, xfade_ms=60.0, **SZ)
        else:
            kw = dict(gate_ui=48.0 + k, up_delay_ms=float((0, 15, 40)[k % 3]), **SZ)
            if mode == "xfade":
                kw["xfade_ms"] = 45.0
        r = eng.run(mode, [x], sr, **kw)[0]
        o, o64 = orc.run(mode, x, sr, **kw), orc.run(mode, x, sr, fft_dtype="float64", **kw)
        if mode == "adaptive" and len(o["states"]) == 0:
            assert np.array_equal(r["out"], o["out"].astype(np.float32))
            continue
        _compare(mode, o, o64, r)


def test_batch_of_uneven_tracks_with_forced_unit_lengths():
    """One plan, several tracks, work units of 1, 2, 3, 7 and the default number of blocks: unit boundaries fall on both pass
    parities, so the carry hand-over between the lanes of a pair, the half pass of warm-up and the odd last pass all run."""
    eng = _engine()
    sr = 48000
    xs = [_q(synth.recipe_gated_pink(s, sr, 300 + i, env_hz=2.0, hi_dbfs=-19.0)) for i, s in enumerate((1.1, 2.37, 0.41, 5.3))]
    kw = dict(gate_ui=50, up_delay_ms=50.0, **SZ)
    want = [(orc.run("standard", x, sr, **kw), orc.run("standard", x, sr, fft_dtype="float64", **kw)) for x in xs]
    outs = {}
    for ub in (0, 1, 2, 3, 7):
        rs = eng.run_streaming("standard", xs, sr, unit_blocks=ub, **kw)
        for r, (o, o64) in zip(rs, want):
            _compare("standard", o, o64, r)
        outs[ub] = [r["out"] for r in rs]
    for ub in (1, 2, 3, 7):                                        # the unit split changes nothing but the limiter's reduction order
        for a, b in zip(outs[0], outs[ub]):
            assert np.abs(a - b).max() <= 2e-7


def test_five_minute_track_inside_a_batch_fused_limiter_at_scale():
    """16 tracks x 5 min @ 44.1 kHz at 2048 / 1024: steady loop, dynamic queue and the fused per-chunk limiter at scale; track 3
    against the oracle."""
    import torch
    eng = _engine()
    sr, n = 44100, 16
    xs = [synth.recipe_gated_pink(300.0, sr, 1000 + i, env_hz=0.2, hi_dbfs=-25.0) for i in range(n)]
    kw = dict(gate_ui=50, **SZ)
    xd = [torch.from_numpy(x).cuda() for x in xs]
    rs = eng.run_streaming("standard", xd, sr, want_host=False, **kw)
    r = dict(rs[3])
    r["out"] = r["out"].cpu().numpy()
    o, o64 = orc.run("standard", xs[3], sr, **kw), orc.run("standard", xs[3], sr, fft_dtype="float64", **kw)
    _compare("standard", o, o64, r)
    assert len(r["chunk_lengths"]) == len(o["chunk_lengths"]) >= 50


def test_fused_2048_equals_the_general_size_path(monkeypatch):
    """The fused pair-mode kernels and the plain general-size kernels (csrc/generic.cuh, TMT_FUSED_2048=0) on the same input."""
    eng = _engine()
    sr = 48000
    x = _q(synth.recipe_gated_pink(3.0, sr, 77, env_hz=2.0, hi_dbfs=-21.0))
    kw = dict(gate_ui=50, up_delay_ms=80.0, **SZ)
    a = eng.run("standard", [x], sr, **kw)[0]
    monkeypatch.setenv("TMT_FUSED_2048", "0")
    b = eng.run("standard", [x], sr, **kw)[0]
    assert np.array_equal(a["meansq"], b["meansq"]) and np.array_equal(a["states"], b["states"])
    assert a["chunk_lengths"] == b["chunk_lengths"]
    assert np.abs(a["out"].astype(np.float64) - b["out"]).max() <= 2e-6


def test_adaptive_with_more_than_two_channels():
    """Channel pairs as linked tracks of one plan (engine.run_adaptive_multichannel) at 2048 / 1024: shared level, gate, scale."""
    eng = _engine()
    sr = 48000
    a = synth.recipe_swept_pink(3.0, sr, 401, period_s=0.7, peak=0.5)
    b = synth.recipe_swept_pink(3.0, sr, 402, period_s=0.5, peak=0.3)
    x = _q(np.concatenate([a, b[:, :1]], axis=1))                                       # 3 channels
    kw = dict(min_hold_ms=80.0, xfade_ms=150.0, **SZ)
    r = eng.run("adaptive", [x], sr, **kw)[0]
    o, o64 = orc.run("adaptive", x, sr, **kw), orc.run("adaptive", x, sr, fft_dtype="float64", **kw)
    assert r["out"].shape == x.shape
    _compare("adaptive", o, o64, r)


def test_static_eq_at_2048():
    from oracle import layer2_oracle as l2
    eng = _engine()
    sr = 48000
    x = _q(synth.recipe_gated_pink(1.7, sr, 71, env_hz=2.0, hi_dbfs=-14.0))
    gain = l2.build_gain_per_bin(sr, 2048, np.array([20.0, 100.0, 500.0, 1000.0, 4000.0, 12000.0, 20000.0]),
                                 np.array([6.0, 4.0, 0.0, -2.0, 3.0, 8.0, 10.0]))
    for kw in (dict(), dict(pad=False, global_gain_db=-3.0, auto_gain_protect=False)):
        r = eng.run_eq([x], sr, gain, **SZ, **kw)[0]
        o, o64 = l2.apply_eq(x, sr, gain, **SZ, **kw), l2.apply_eq(x, sr, gain, fft_dtype="float64", **SZ, **kw)
        assert r["out"].shape == o["out"].shape
        d = np.abs(r["out"].astype(np.float64) - o["out"]).max(axis=1)
        d64 = np.abs(r["out"].astype(np.float64) - o64["out"]).max(axis=1) / np.maximum(1.0, np.abs(o64["out"]).max(axis=1))
        assert float(d[2048:-2048].max()) <= 1e-5 and float(d64.max()) <= 1e-5
        assert abs(r["peak_seen"] - o64["peak_seen"]) <= 2e-5 * o64["peak_seen"]
        assert (r["out_gp"] is None) == (o["out_gp"] is None)
        if o["out_gp"] is not None:
            assert np.array_equal(r["out_gp"], (l2.pcm24_roundtrip(r["out"]) * np.float32(r["scale"])).astype(np.float32))


def test_channel_state_analyser_at_2048():
    from oracle import analysis_oracle as ao
    from test_channel_states import _check_result
    eng = _engine()
    sr = 44100
    x = synth.recipe_swept_pink(4.0, sr, 31, period_s=0.9, peak=0.5)
    x[:, 1] = np.roll(x[:, 1], 5000) * 0.6
    x = _q(x)[:1024 * 150 + 333]
    kw = dict(min_hold_ms=100.0, target_c2=0.4, **SZ)
    _check_result(eng.run_channel_states([x], sr, **kw)[0], ao.analyze(x, sr, **kw))


def test_batch_front_ends_at_2048():
    """batch.DeviceBatch and batch.HostBatchPipeline with n_fft / hop keywords use the 2048 build; samples equal engine.run's."""
    import torch
    from tomatis_audio_processor_b200.batch import DeviceBatch, HostBatchPipeline
    eng = _engine()
    n, sr, T = 200000, 48000, 4
    xs = np.stack([_q(synth.recipe_gated_pink(n / sr, sr, 60 + i, env_hz=1.1, hi_dbfs=-22.0))[:n] for i in range(T)])
    want = [r["out"] for r in eng.run("standard", list(xs), sr, gate_ui=50, **SZ)]
    h_in = torch.from_numpy(xs).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    p = HostBatchPipeline(n, sr, "standard", wave_tracks=2, gate_ui=50, **SZ)
    assert p.eng.n_fft == 2048
    p.process(h_in, h_out)
    torch.cuda.synchronize()
    p.close()
    x = h_in.cuda()
    y = torch.empty_like(x)
    db = DeviceBatch(x, y, sr, "standard", gate_ui=50, **SZ)
    db.step()
    torch.cuda.synchronize()
    db.close()
    for i in range(T):
        assert np.abs(h_out[i].numpy() - want[i]).max() <= 2e-7 and np.abs(y[i].cpu().numpy() - want[i]).max() <= 2e-7


def test_validators_at_2048():
    """validate_layer1 / verify_tomatis_15db_v2 kernels (gate re-simulation, conditional spectrum, anchored variant) on a file
    processed at 2048 / 1024, against the oracle restatement of the reference's functions at the same sizes."""
    from oracle import validate_oracle as vo
    from tomatis_audio_processor_b200 import validators as prod
    eng = _engine()
    sr = 48000
    x = _q(synth.recipe_gated_pink(8.0, sr, 88, env_hz=0.6, hi_dbfs=-22.0))
    y = eng.run("standard", [x], sr, gate_ui=50, up_delay_ms=100.0, **SZ)[0]["out"]
    st_o, lv_o = vo.simulate_gate(x, sr, 2048, 1024, -40.0, 3.0, 100.0)
    st, lv = prod.simulate_gate(x, sr, 2048, 1024, -40.0, 3.0, 100.0)
    assert st == st_o and np.array_equal(np.array(lv), np.array(lv_o))
    f, c1, c2, n1, n2 = prod.compute_conditional_spectrum(x, y, sr, st, 2048, 1024)
    fo, o1, o2, m1, m2, _ = vo.compute_conditional_spectrum(x, y, sr, st_o, 2048, 1024)
    assert (n1, n2) == (m1, m2) and n1 > 0 and n2 > 0 and np.array_equal(f, fo) and len(f) == 1025
    assert max(float(np.abs(c1 - o1).max()), float(np.abs(c2 - o2).max())) < 2e-4
    f, a1, a2, k1, k2 = prod.compute_conditional_spectrum_v2(x, y, sr, st, np.array(lv), 2048, 1024)
    fo, b1, b2, j1, j2, _ = vo.compute_conditional_spectrum_v2(x, y, sr, st_o, np.array(lv_o), 2048, 1024)
    assert (k1, k2) == (j1, j2)
    assert max(float(np.abs(a1 - b1).max()), float(np.abs(a2 - b2).max())) < 2e-4


@pytest.mark.parametrize("mode,kw", [("standard", dict(gate_ui=50, up_delay_ms=80.0)), ("xfade", dict(gate_ui=60, xfade_ms=100.0)),
                                     ("adaptive", dict(min_hold_ms=100.0, xfade_ms=200.0))])
def test_time_sharded_equals_unsharded_at_2048(mode, kw):
    """One file over three ranks (threads of this process on one GPU, tests/thread_comm.py) at 2048 / 1024: halos of one 1024-sample
    hop, hop-sum exchange, redundant gate scan -- bit-identical to the whole-file call."""
    import torch
    from test_gpu_sharded import _run_threads
    from tomatis_audio_processor_b200 import sharded
    eng = _engine()
    sr, world = 48000, 3
    x = (synth.recipe_swept_pink(4.0, sr, 31, period_s=1.1, peak=0.5) if mode == "adaptive"
         else synth.recipe_gated_pink(16.0, sr, 30, env_hz=0.9, hi_dbfs=-22.0))
    total = len(x)
    whole = eng.run(mode, [x], sr, **kw, **SZ)[0]
    framing = sharded.WHOLEFILE if mode == "adaptive" else sharded.STREAMING
    shards = sharded.plan_shards(total, world, framing, 2048, 1024)
    assert all(s.hop == 1024 and s.n_fft == 2048 for s in shards)

    def rank_fn(comm, r):
        me = shards[r]
        own = torch.from_numpy(x[me.own_lo:me.own_hi].copy()).cuda()
        if mode == "adaptive":
            out = sharded.run_adaptive_sharded(own, sr, total, comm, gather_to=0, **kw, **SZ)
        else:
            out = sharded.run_streaming_sharded(mode, own, sr, total, comm, gather_to=0, **kw, **SZ)
        out["out"] = out["out"].cpu().numpy()
        if out["full"] is not None:
            out["full"] = out["full"].cpu().numpy()
        return out
    res = _run_threads(world, rank_fn)
    for r, o in enumerate(res):
        me = shards[r]
        assert np.array_equal(o["meansq"], whole["meansq"]) and np.array_equal(o["states"], whole["states"])
        assert np.array_equal(o["out"], whole["out"][me.own_lo:me.own_hi]), (mode, r)
    assert np.array_equal(res[0]["full"], whole["out"])


def test_streamed_file_at_2048():
    import torch
    from tomatis_audio_processor_b200.streamed import HostFileStreamer
    eng = _engine()
    sr = 48000
    x = synth.recipe_gated_pink(31.0, sr, 501, env_hz=0.7, hi_dbfs=-21.0)
    st = HostFileStreamer("standard", len(x), sr, slab_seconds=6.0, n_slots=2, gate_ui=50, **SZ)
    assert len(st.slabs) >= 4
    h_out = torch.empty((len(x), 2), dtype=torch.float32).pin_memory()
    st.process(torch.from_numpy(x).pin_memory(), h_out)
    torch.cuda.synchronize()
    r = eng.run("standard", [x], sr, gate_ui=50, **SZ)[0]
    assert np.array_equal(h_out.numpy(), r["out"]) and np.array_equal(st.states_rows()[0], r["states"])
    st.close()


def test_integer_pcm_pipeline_at_2048():
    """HostBatchPipeline with int16 in / PCM_24 out at 2048 / 1024: the fused conversion + hop-sum pass of the second build against
    the separate passes, byte for byte, and against the float pipeline quantised on the host."""
    import torch
    from tomatis_audio_processor_b200 import audio_io
    from tomatis_audio_processor_b200.batch import HostBatchPipeline
    n, sr, T = 150001, 48000, 2
    xs = np.stack([_q(synth.recipe_gated_pink(n / sr + 0.01, sr, 90 + i, env_hz=1.1, hi_dbfs=-22.0))[:n] for i in range(T)])
    s_in = torch.from_numpy(np.stack([synth.quantise_pcm16(x) for x in xs])).pin_memory()
    outs = []
    for fused in ("1", "0"):
        os.environ["TMT_PCM_FUSED"] = fused
        try:
            p = HostBatchPipeline(n, sr, "standard", wave_tracks=1, in_format="s16", out_format="s24", gate_ui=50, **SZ)
        finally:
            os.environ.pop("TMT_PCM_FUSED")
        s_out = torch.empty((T, n, 6), dtype=torch.uint8).pin_memory()
        p.process(s_in, s_out)
        torch.cuda.synchronize()
        p.close()
        outs.append(s_out.numpy().copy())
    assert np.array_equal(outs[0], outs[1])
    want = _engine().run("standard", list(xs), sr, gate_ui=50, **SZ)
    for i in range(T):
        q = audio_io.quantise_pcm24(want[i]["out"]).reshape(-1)
        got = outs[0][i].reshape(-1, 3).astype(np.int32)
        got = got[:, 0] | (got[:, 1] << 8) | (got[:, 2] << 16)
        got = np.where(got & 0x800000, got - 0x1000000, got)
        assert np.abs(got - q).max() <= 2                          # the batch plan's unit split can move a limited chunk by an LSB or two
