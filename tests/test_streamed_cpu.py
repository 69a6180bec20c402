"""CPU: the host-side geometry of streamed.HostFileStreamer (plan_slabs).  The streamed pass chains the gate through a per-file
array of hop-block sums that is filled slab by slab; this test replays that chain with NumPy on the slabs' own windows only and
checks that, when a slab runs, every frame up to its own last one has exactly the oracle's mean square -- i.e. that a slab never
needs a sample or a sum from a later slab."""
import numpy as np
import pytest

from oracle import tomatis_oracle as orc
from tomatis_audio_processor_b200 import synth, tables as tb
from tomatis_audio_processor_b200.streamed import plan_slabs


@pytest.mark.parametrize("secs,slab,sr", [(23.0, 5.0, 48000), (11.3, 5.0, 44100), (4.0, 300.0, 48000), (31.0, 6.0, 96000)])
def test_slab_chain_reproduces_the_whole_file_mean_squares(secs, slab, sr):
    x = synth.recipe_gated_pink(secs, sr, 9, env_hz=1.0, hi_dbfs=-20.0)[:int(secs * sr) - 7]
    total = len(x)
    want = np.asarray(orc.run("standard", x, sr, gate_ui=50)["meansq"])
    slabs = plan_slabs(total, sr, slab)
    n_frames = slabs[0][0].n_frames
    assert n_frames == len(want) and len(slabs) == max(1, min(round(total / (slab * sr)), len(tb.flush_chunk_blocks(n_frames))))
    # the slabs partition the file, in order
    assert slabs[0][0].own_lo == 0 and slabs[-1][0].own_hi == total
    assert all(a[0].own_hi == b[0].own_lo for a, b in zip(slabs, slabs[1:]))
    m2 = (0.5 * (x[:, 0] * x[:, 0] + x[:, 1] * x[:, 1])).astype(np.float32)        # checked against the oracle below, not assumed
    m2 = np.sqrt(m2) * np.sqrt(m2)
    G = np.zeros(n_frames + 1, np.float32)
    have = np.zeros(n_frames + 1, bool)
    for s, hb_lo, hb_hi, f_hi in slabs:
        assert hb_lo == s.block_lo and hb_hi in (s.block_hi, s.block_hi + 1) and hb_hi <= n_frames + 1
        for q in range(hb_lo, hb_hi):                                             # sums from the slab's own window only
            p0 = s.first_start + q * tb.HOP
            lo, hi = max(p0, 0), min(p0 + tb.HOP, total)
            assert lo >= s.in_lo and hi <= s.in_hi, "hop block outside the slab's window"
            blk = np.zeros(tb.HOP, np.float32)
            if hi > lo:
                blk[lo - p0:hi - p0] = m2[lo:hi]
            G[q] = np.sum(blk, dtype=np.float32)
            have[q] = True
        assert f_hi >= min(s.block_hi, n_frames) and have[:f_hi + 1].all() if f_hi else True
        msq = ((G[:f_hi] + G[1:f_hi + 1]) / np.float32(tb.N_FFT)).astype(np.float32)
        # frames of this slab: block_lo - 1 .. block_hi - 1; everything before was already right when the earlier slabs ran
        f0 = max(0, s.block_lo - 1)
        assert np.allclose(msq[f0:f_hi], want[f0:f_hi], rtol=3e-6, atol=0)         # float32 summation order differs from the kernel's exact one
    assert have.all()


def test_streamer_rejects_what_it_cannot_stream():
    from tomatis_audio_processor_b200.streamed import HostFileStreamer
    with pytest.raises(ValueError):
        HostFileStreamer("standard", 0, 48000)
    with pytest.raises(ValueError):
        HostFileStreamer("adaptive", 48000, 48000)
