#!/usr/bin/env bash
# Dev tool (B200 box): stft_kernel time of the bench workload against the work-unit length (hop blocks per unit).
cd "$(dirname "$0")/.."
for ub in 0 24 30 40 59 118; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-longfile --unit-blocks $ub 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('unit_blocks %-4s step %.3f ms  stft %.3f ms  frac %.4f' % ('$ub', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
done
