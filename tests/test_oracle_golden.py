"""The NumPy restatement (oracle/) against the committed outputs of the reference itself.

tests/golden/*.npz were produced by oracle/make_golden.py, which RUNS the reference's
process() functions (in-memory soundfile stand-in).  On the pinned NumPy (2.3.5) the
restatement is bit-identical; on another NumPy build the PCM tolerance below applies
(float32 pocketfft rounding) while chunk lengths and gate states must still be exact.
"""
import numpy as np
import pytest

from oracle import tomatis_oracle as orc
from helpers import golden_names, load_golden, csv_states

PCM_TOL = 2e-7          # only matters off the pinned NumPy build; 0.0 observed on 2.3.5


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_fixture(name):
    g = load_golden(name)
    res = orc.run(g["mode"], g["x"], g["sr"], **g["kwargs"])
    assert res["chunk_lengths"] == g["chunk_lengths"]
    assert res["out"].shape == g["out"].shape
    assert str(res["out"].dtype) == g["out_dtype"]
    err = float(np.max(np.abs(res["out"].astype(np.float64) - g["out"].astype(np.float64))))
    assert err <= PCM_TOL, err
    rows = orc.csv_rows(g["mode"], res)
    assert csv_states(rows) == csv_states(g["csv"])           # gate decisions: exact
    if np.__version__ == g["numpy"]:
        assert err == 0.0
        assert rows == g["csv"]                                # levels/alpha text: exact


def test_golden_cover_all_modes_and_edges():
    gs = [load_golden(n) for n in golden_names()]
    assert {g["mode"] for g in gs} == {"standard", "xfade", "adaptive"}
    assert any(len(g["chunk_lengths"]) > 1 for g in gs)                 # limiter chunk boundary
    assert any(g["out_dtype"] == "float64" for g in gs)                 # adaptive fp64 branch
    assert any(g["sr"] != 48000 and g["guard_skipped"] for g in gs)     # 44.1 / 96 kHz
    assert any(len(g["x"]) % 2048 == 0 for g in gs)                     # pad_end == 0 tail
    for g in gs:
        st = csv_states(g["csv"])
        assert "C1" in st and "C2" in st                                # every case switches
