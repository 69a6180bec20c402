"""Static EQ processor (src/layer2_apply_eq.py): oracle vs the executed reference's fixtures (CPU),
product host helpers (CPU), CUDA path vs oracle (GPU)."""
import numpy as np
import pytest

from helpers import eq_golden_names, load_eq_golden
from oracle import layer2_oracle as l2
from tomatis_audio_processor_b200 import layer2_apply_eq as prod

PCM_TOL = 1e-5


@pytest.mark.parametrize("name", eq_golden_names())
def test_eq_oracle_matches_reference_fixture(name):
    g = load_eq_golden(name)
    gain = l2.build_gain_per_bin(g["sr"], 4096, g["eq_freqs"], g["eq_db"])
    assert np.array_equal(gain, g["gain_bins"])
    o = l2.apply_eq(g["x"], g["sr"], gain, **g["kwargs"])
    assert o["out"].shape == g["out"].shape
    err = float(np.abs(o["out"].astype(np.float64) - g["out"]).max())
    assert err <= 2e-7 * max(1.0, float(np.abs(g["out"]).max()))
    if np.__version__ == g["numpy"]:
        assert np.array_equal(o["out"], g["out"])
        assert (o["out_gp"] is None) == (g["out_gp"] is None)
        if g["out_gp"] is not None:
            assert np.array_equal(o["out_gp"], g["out_gp"])


def test_eq_host_helpers_match_oracle(tmp_path):
    g = load_eq_golden(eq_golden_names()[0])
    p = tmp_path / "eq.csv"
    with open(p, "w") as f:
        f.write("freq,delta_db_smooth,delta_db\n")           # alias + preferred column, unsorted rows
        for a, b in sorted(zip(g["eq_freqs"], g["eq_db"]), key=lambda r: -r[0]):
            f.write(f"{float(a)!r},{float(b)!r},99\n")
    fr, db = prod.load_eq_csv(str(p))
    assert np.array_equal(fr, g["eq_freqs"]) and np.array_equal(db, g["eq_db"])
    assert np.array_equal(prod.build_gain_per_bin(g["sr"], 4096, fr, db), g["gain_bins"])
    flags = {s for a in prod.build_parser()._actions for s in a.option_strings if s.startswith("--")} - {"--help"}
    assert flags == {"--input", "--output", "--eq_csv", "--n_fft", "--hop", "--no_pad", "--gain_db", "--no_gain_protect", "--device"}


def _cmp(got, ref, ref64, interior=4096):
    """interior: <= 1e-5 vs the reference output; everywhere: <= 1e-5 (relative to max(1,|y|)) vs the float64-FFT
    evaluation (the first and last hop are single-frame blocks divided by w^2 -> 0)."""
    d = np.abs(got.astype(np.float64) - ref).max(axis=1)
    d64 = np.abs(got.astype(np.float64) - ref64).max(axis=1) / np.maximum(1.0, np.abs(ref64).max(axis=1))
    return float(d[interior:-interior].max()), float(d64.max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", eq_golden_names())
def test_eq_gpu_matches_reference_fixture(name):
    from tomatis_audio_processor_b200 import engine
    g = load_eq_golden(name)
    r = engine.run_eq([g["x"]], g["sr"], g["gain_bins"], **g["kwargs"])[0]
    o64 = l2.apply_eq(g["x"], g["sr"], g["gain_bins"], fft_dtype="float64", **g["kwargs"])
    assert r["out"].shape == g["out"].shape
    inner, rel64 = _cmp(r["out"], g["out"], o64["out"])
    print(f"{name}: interior {inner:.2e} vs reference, everywhere {rel64:.2e} (relative) vs fp64-FFT; peak {r['peak_seen']:.4f}")
    assert inner <= PCM_TOL and rel64 <= PCM_TOL
    assert abs(r["peak_seen"] - o64["peak_seen"]) <= 2e-5 * o64["peak_seen"]
    assert (r["out_gp"] is None) == (g["out_gp"] is None)
    if g["out_gp"] is not None:
        # the second file = PCM_24 round trip of the first, times peak_target / peak
        want = (l2.pcm24_roundtrip(r["out"]) * np.float32(r["scale"])).astype(np.float32)
        assert np.array_equal(r["out_gp"], want)
        assert float(np.abs(r["out_gp"].astype(np.float64) - g["out_gp"]).max()) <= PCM_TOL


@pytest.mark.gpu
def test_eq_gpu_file_frontend_and_batch(tmp_path, capsys):
    from tomatis_audio_processor_b200 import audio_io, engine, synth
    x = synth.recipe_gated_pink(2.0, 48000, 72, env_hz=1.0, hi_dbfs=-30.0, bursts=False)
    p = str(tmp_path / "in.wav")
    audio_io.write(p, x, 48000, subtype="PCM_24")
    xq, _ = audio_io.read(p, dtype="float32")
    csvp = str(tmp_path / "eq.csv")
    with open(csvp, "w") as f:
        f.write("freq_hz,delta_db\n20,-6\n200,-3\n1000,0\n5000,-2\n20000,-8\n")
    want = str(tmp_path / "out.flac")
    prod.apply_eq_stft(p, want, csvp, global_gain_db=-2.0)
    got_path = want if audio_io.have_soundfile() else want.replace(".flac", ".wav")
    y, sr = audio_io.read(got_path, dtype="float32")
    fr, db = prod.load_eq_csv(csvp)
    gain = l2.build_gain_per_bin(48000, 4096, fr, db)
    o = l2.apply_eq(xq, 48000, gain, global_gain_db=-2.0)
    assert sr == 48000 and y.shape == o["out"].shape
    d = np.abs(y.astype(np.float64) - np.clip(o["out"], -1, 1)).max(axis=1)
    assert float(d[4096:-4096].max()) <= PCM_TOL + 1.2e-7
    # several tracks of different length in one plan
    xs = [xq, xq[:50000], xq[:4096]]
    rs = engine.run_eq(xs, 48000, gain, pad=False, auto_gain_protect=False)
    for xi, ri in zip(xs, rs):
        oi = l2.apply_eq(xi, 48000, gain, pad=False, auto_gain_protect=False)
        o64 = l2.apply_eq(xi, 48000, gain, pad=False, auto_gain_protect=False, fft_dtype="float64")
        assert ri["out"].shape == oi["out"].shape
        if len(oi["out"]) > 8192:
            inner, rel64 = _cmp(ri["out"], oi["out"], o64["out"])
            assert inner <= PCM_TOL and rel64 <= PCM_TOL
    capsys.readouterr()
