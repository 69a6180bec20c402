"""streamed.HostFileStreamer: one file from / to pinned host buffers in bounded device memory (time slabs on limiter-chunk
boundaries through a few device slots) against the whole-file call -- bit-identical samples, identical gate states."""
import numpy as np
import pytest

from oracle import tomatis_oracle as orc
from tomatis_audio_processor_b200 import synth

pytestmark = pytest.mark.gpu


def _pinned(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).pin_memory()


@pytest.mark.parametrize("mode,kw", [("standard", dict(gate_ui=50, up_delay_ms=120.0)), ("xfade", dict(gate_ui=60, xfade_ms=300.0))])
def test_streamed_file_equals_whole_file_call(mode, kw):
    import torch
    from tomatis_audio_processor_b200 import engine
    from tomatis_audio_processor_b200.streamed import HostFileStreamer
    sr = 48000
    xs = [synth.recipe_gated_pink(41.3, sr, 500 + i, env_hz=0.7, hi_dbfs=-21.0) for i in range(2)]
    total = len(xs[0])
    st = HostFileStreamer(mode, total, sr, slab_seconds=8.0, n_slots=2, **kw)
    assert len(st.slabs) >= 4 and len(st.slots) == 2
    assert st.device_bytes() < 0.7 * 2 * total * 8                   # bounded by the slots, not by the file
    h_out = torch.empty((total, 2), dtype=torch.float32).pin_memory()
    for x in xs:                                                      # the second file reuses every slot and every plan
        st.process(_pinned(x), h_out)
        torch.cuda.synchronize()
        r = engine.run(mode, [x], sr, **kw)[0]
        assert np.array_equal(h_out.numpy(), r["out"])
        states, rows = st.states_rows()
        assert np.array_equal(states, r["states"]) and np.array_equal(rows, r["rows"])
    o = orc.run(mode, xs[1], sr, **kw)
    assert np.array_equal(states, o["states"])
    assert np.abs(h_out.numpy().astype(np.float64) - orc.run(mode, xs[1], sr, fft_dtype="float64", **kw)["out"]).max() <= 1e-5
    st.close()


def test_streamed_file_single_slab_and_ragged_tail():
    import torch
    from tomatis_audio_processor_b200 import engine
    from tomatis_audio_processor_b200.streamed import HostFileStreamer
    sr = 44100
    for secs, slab in ((3.0, 300.0), (17.77, 5.0)):
        x = synth.recipe_gated_pink(secs, sr, 77, env_hz=1.5, hi_dbfs=-20.0)[:int(secs * sr) - 13]
        st = HostFileStreamer("standard", len(x), sr, slab_seconds=slab, gate_ui=50)
        h_out = torch.empty((len(x), 2), dtype=torch.float32).pin_memory()
        st.process(_pinned(x), h_out)
        torch.cuda.synchronize()
        assert np.array_equal(h_out.numpy(), engine.run("standard", [x], sr, gate_ui=50)[0]["out"])
        st.close()
