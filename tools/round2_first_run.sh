#!/usr/bin/env bash
# First GPU call of the next round (run under gpurun from the repo root): what this round could not run on hardware.
#   gpurun --timeout 600 -- 'bash tools/round2_first_run.sh'
# 1. the experimental general-FFT-size path (generic.py / csrc/generic.cuh) against the oracle -- if green, drop the
#    TMT_GENERIC_FFT gate in engine.py / generic.enabled() and the skipif in tests/test_generic_sizes.py;
# 2. the calibration cross-correlation's operand-swap branch and the 2-GPU command lines (need gpurun --gpus 2);
# 3. a timing of the general path so that DESIGN.md can state its cost next to the fused kernel's.
set -u
out=gpurun_out/round2_first
mkdir -p "$out"
TMT_GENERIC_FFT=1 python -m pytest tests/test_generic_sizes.py -m gpu -q -x > "$out/generic_gpu.log" 2>&1; echo "generic rc=$?"
python -m pytest tests/test_calibration.py -m gpu -q > "$out/calibration_gpu.log" 2>&1; echo "calibration rc=$?"
python -m pytest tests/test_process_sharded.py tests/test_gpu_sharded.py -m gpu -q > "$out/two_gpu.log" 2>&1; echo "sharded rc=$? (2-GPU tests skip on one GPU)"
TMT_GENERIC_FFT=1 python - > "$out/generic_timing.txt" 2>&1 <<'PY'
import time, numpy as np, torch
from tomatis_audio_processor_b200 import engine, synth
x = synth.recipe_gated_pink(300.0, 48000, 1)
for n_fft, hop in ((4096, 2048), (2048, 1024), (1024, 512), (4096, 1024)):
    for _ in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = engine.run("standard", [x], 48000, n_fft=n_fft, hop=hop, gate_ui=50)[0]
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"standard 5 min @ 48 kHz n_fft={n_fft} hop={hop}: {dt*1e3:.1f} ms per call incl. host copies, {len(r['states'])} frames")
PY
echo "timing rc=$?"; tail -3 "$out"/*.log "$out/generic_timing.txt"
