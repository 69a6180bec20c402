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
