"""CPU: host-side scalar/table logic against the oracle's (= the reference's) expressions."""
import numpy as np
import pytest

from oracle import tomatis_oracle as orc
from tomatis_audio_processor_b200 import tables as tb


@pytest.mark.parametrize("T", [-38.5, -41.5, -48.5, -51.5, -70.0, -20.25, -100.0, -119.9, 5.0])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_meansq_thresholds_are_exact_inverses_of_the_level_chain(T, dtype):
    """level(m) >= Ton  <=>  m >= m_on   and   level(m) <= Toff  <=>  m <= m_off, for every representable m: the GPU compares
    bit-exact mean squares against these, which is what makes its gate decisions identical to the reference's."""
    m_on, m_off = tb.meansq_threshold_on(T, dtype), tb.meansq_threshold_off(T, dtype)
    rng = np.random.default_rng(int(abs(T) * 100))
    probes = []
    for c in (m_on, m_off):
        if np.isfinite(c) and c > 0:
            x = dtype(c)
            probes += [x, np.nextafter(x, dtype(0)), np.nextafter(x, dtype(np.inf))]
    probes += list((10.0 ** rng.uniform(-14, 1, 200)).astype(dtype)) + [dtype(0)]
    for m in probes:
        lvl = orc.level_from_meansq(dtype(m))
        assert (lvl >= T) == (float(m) >= m_on), (m, lvl, T, m_on)
        assert (lvl <= T) == (float(m) <= m_off), (m, lvl, T, m_off)


def test_tables_match_oracle_expressions():
    for sr in (44100, 48000, 96000):
        freqs = np.fft.rfftfreq(4096, d=1.0 / sr)
        for lo, hi in ((15.0, -15.0), (-15.0, 15.0), (6.0, -3.0)):
            assert np.array_equal(tb.tilt_gain_db(freqs, 1000.0, 12.0, lo, hi), orc.build_tilt_gain_db(freqs, 1000.0, 12.0, lo, hi))
    g1, g2 = tb.tilt_curves_db(48000, 4096, 1000.0, 12.0, 15.0, -15.0, -15.0, 15.0)
    assert g1[0] == 15.0 and g1[2048] == -15.0 and abs(g1[85] - 0.0678) < 1e-3 and abs(g1[86] + 0.1347) < 1e-3
    assert np.array_equal(tb.gain_rows_standard(g1, g2)[0], orc.db_to_lin_f32(g1))
    assert tb.hann_window()[0] == 0.0 and abs(float(tb.hann_window()[1]) - 5.8856e-7) < 1e-10
    # derived known answers of SURVEY.md 8c
    assert tb.gate_threshold_log_percent(50, 80.0) == -40.0 and tb.gate_threshold_linear(50, 1.0, -100) == -50.0
    assert tb.hysteresis_pair(-40.0, 3.0) == (-38.5, -41.5)
    assert tb.updelay_run_frames(48000, 250.0) == 7 and tb.updelay_run_frames(48000, 0.0) == 1
    assert tb.adaptive_frame_counts(48000, 250.0, 500.0) == (6, 12) and tb.xfade_frame_count(48000, 0.0) == 0
    assert np.allclose(tb.alpha_ramp(12)[1:4], [1 / 12, 2 / 12, 3 / 12])


def test_adaptive_attenuation_dtype_branch():
    """peak <= 10^(-17/20): atten_db is the Python int 0 -> float64 pipeline; above: float32."""
    a_db, a_lin, f64 = tb.adaptive_attenuation(np.float32(0.05), 15.0, 15.0, 2.0)
    assert f64 and a_db == 0 and isinstance(a_lin, np.float64) and a_lin == 1.0
    a_db, a_lin, f64 = tb.adaptive_attenuation(np.float32(0.5), 15.0, 15.0, 2.0)
    assert (not f64) and isinstance(a_db, np.floating) and a_db > 0
    x = np.float32(0.5)
    assert a_db == max(0, 20 * np.log10(x + 1e-12) + 15.0 + 2.0)


def test_flush_schedule_closed_form_equals_the_frame_by_frame_replay():
    for n in list(range(0, 700)) + [1292, 6460, 14063, 337500, 337501]:
        assert tb.flush_chunk_blocks(n) == tb._flush_chunk_blocks_replay(n), n
    for nf, hp in ((2048, 1024), (8192, 4096), (4096, 1024), (4096, 4096)):
        for n in (0, 1, 100, 500, 1000, 5000):
            assert tb.flush_chunk_blocks(n, nf, hp) == tb._flush_chunk_blocks_replay(n, nf, hp), (nf, hp, n)


def test_random_geometry_and_resampler_tables():
    """Seeded sweeps of the host tables against their checkers: limiter chunk ranges of the general-size path and both frame
    layouts against the oracle's frame-by-frame versions, the polyphase resampler plan against scipy's own output length and
    filter, the up-delay run length against a direct replay of the reference's arming rule."""
    from scipy.signal import firwin, resample_poly
    from oracle import tomatis_oracle as orc
    from tomatis_audio_processor_b200 import generic
    rng = np.random.default_rng(99)
    for _ in range(60):
        n_fft = int(2 ** rng.integers(7, 14))
        hop = int(rng.integers(1, n_fft + 1))
        total = int(rng.integers(0, 900000))
        _, _, starts = orc.frame_layout_streaming(total, n_fft, hop)
        first, nf = generic.streaming_layout(total, n_fft, hop)
        assert nf == len(starts)
        want = [(max(0, a), min(total, b)) for a, b in orc.flush_schedule(nf, n_fft, hop) if min(total, b) > max(0, a)]
        assert generic.flush_sample_ranges(nf, total, n_fft, hop) == want
    for _ in range(40):
        n_in, up, down = int(rng.integers(1, 5000)), int(rng.integers(1, 12)), int(rng.integers(1, 60))
        plan = tb.resample_poly_plan(n_in, up, down)
        x = rng.standard_normal(n_in).astype(np.float32)
        assert plan["n_out"] == len(resample_poly(x, up, down))
        if plan["h"] is not None:
            m = max(plan["up"], plan["down"])
            h = firwin(2 * 10 * m + 1, 1.0 / m, window=("kaiser", 5.0)).astype(np.float32) * plan["up"]
            core = plan["h"][np.flatnonzero(plan["h"])[0]:np.flatnonzero(plan["h"])[-1] + 1]
            ref = h[np.flatnonzero(h)[0]:np.flatnonzero(h)[-1] + 1]
            assert core.shape == ref.shape and np.allclose(core, ref, rtol=2e-6, atol=1e-9)
    for _ in range(200):
        sr, ms, hop = int(rng.choice([44100, 48000, 96000])), float(rng.uniform(0, 400)), int(rng.choice([256, 512, 1024, 2048, 3000]))
        d = int(sr * ms / 1000.0)
        # replay: every frame is loud; frame j (0-based) starts at j*hop, armed at frame 0 with pending = d
        j = 0
        while j * hop < d:
            j += 1
        assert tb.updelay_run_frames(sr, ms, hop) == j + 1


def test_fast_percentiles_equal_numpy_bit_for_bit():
    """engine._percentile_thresholds (one sort, NumPy's lerp restated) against np.percentile / np.median on level-like data of
    many sizes, float32-valued and float64-valued, with ties."""
    from tomatis_audio_processor_b200.engine import _percentile_thresholds
    rng = np.random.default_rng(11)
    for n in list(range(1, 40)) + [100, 101, 1291, 14062, 14063]:
        for kind in range(3):
            v = rng.standard_normal(n) * 12.0 - 40.0
            if kind == 1:
                v = v.astype(np.float32).astype(np.float64)          # float32-valued levels (the float32 branch)
            elif kind == 2:
                v = np.round(v)                                       # many ties
            got = _percentile_thresholds(v, np.ones(n, bool))
            want = (np.percentile(v, 5), np.percentile(v, 95), np.median(v))
            assert all(float(a) == float(b) for a, b in zip(got, want)), (n, kind, got, want)
    assert _percentile_thresholds(np.zeros(3), np.zeros(3, bool)) is None
