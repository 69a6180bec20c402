#!/usr/bin/env bash
# Dev tool.  Build:  tools/ab.sh build NAME "-DTMT_AB_X=0 ..."   (here, nvcc only)   -> tomatis_audio_processor_b200/csrc/ab/NAME.so
#            Run:    tools/ab.sh run [tracks]                     (on the B200 box)   -> one line per built variant, same box, same inputs
set -u
cd "$(dirname "$0")/.."
AB=tomatis_audio_processor_b200/csrc/ab
case "$1" in
  build)
    mkdir -p $AB
    ( cd tomatis_audio_processor_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v $3 \
        -o ab/$2.so tomatis_b200.cu 2>&1 | grep -A2 "stft_kernel" | grep -E "spill|registers" | tr '\n' ' ' ); echo " <- $2 ($3)";;
  run)
    T=${2:-128}
    for so in $AB/*.so; do
      TMT_LIB=$PWD/$so timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-longfile --tracks-per-gpu $T 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('%-28s step %.3f ms  stft %.3f ms  frac %.4f  peak_out %.6f' % ('$(basename $so .so)', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['checks']['output_peak']))"
    done;;
esac
