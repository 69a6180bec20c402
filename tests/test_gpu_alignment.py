"""The streaming helper kernels use 128-bit accesses with the odd leading / trailing sample-frame peeled (input_peak_kernel,
limiter_kernel): tracks that start on an 8-byte but not 16-byte boundary, with odd and even lengths, must give exactly what
the same samples give from a freshly allocated (256-byte aligned) buffer."""
import numpy as np
import pytest

from tomatis_audio_processor_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["adaptive", "standard"])
def test_unaligned_device_views_bit_identical(mode):
    import torch
    from tomatis_audio_processor_b200 import engine
    sr = 48000
    x = synth.recipe_swept_pink(6.0, sr, 77, period_s=0.9, peak=0.7)                 # loud: every limiter path engages
    base = torch.from_numpy(np.concatenate([np.zeros((1, 2), np.float32), x])).cuda()
    run = ((lambda xs, outs: engine.run_adaptive(xs, sr, want_host=False, outs=outs)) if mode == "adaptive" else
           (lambda xs, outs: engine.run_streaming("standard", xs, sr, want_host=False, outs=outs, gate_ui=50)))
    for n in (len(x), len(x) - 1, len(x) - 2, 4097):
        xa = base[1:1 + n]                                                           # data_ptr % 16 == 8
        assert xa.data_ptr() % 16 == 8 and xa.is_contiguous()
        xb = xa.clone()
        assert xb.data_ptr() % 16 == 0
        ya = torch.full((n + 1, 2), 7.0, device="cuda")[1:]                          # output view misaligned the same way
        yb = torch.empty_like(xb)
        ra, rb = run([xa], [ya])[0], run([xb], [yb])[0]
        assert torch.equal(ya, yb)
        assert float(ya.abs().max()) <= 0.999 + 1e-6
        if mode == "adaptive":
            assert ra["input_peak"] == rb["input_peak"] == float(xa.abs().max())
            assert ra["output_peak"] == rb["output_peak"]
            if n > 4097:
                assert ra["output_peak"] > 0.999                                     # the limiter did run
        else:
            assert np.array_equal(ra["chunk_peaks"], rb["chunk_peaks"])
