#!/usr/bin/env bash
# Dev loop on the B200 box: golden parity (fast subset) then the device-resident bench line (value, ms, roofline frac, kernel ms).
set -u
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or meansq or ragged or batch_of" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-longfile 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value %.0f audio-s/s  step %.3f ms  stft %.3f ms  frac %.4f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
