#!/usr/bin/env bash
# Dev probe: host<->device bandwidth with 1 and with N processes at once (one per GPU), to show what the box's host side gives
# the end-to-end pipeline when every GPU copies at the same time.  usage: tools/pcie_probe_all.sh N
N=${1:-8}
cd "$(dirname "$0")/.."
echo "== one process (GPU 0) =="
CUDA_VISIBLE_DEVICES=0 python tools/pcie_probe.py | tail -1
echo "== $N processes at once, one per GPU =="
for i in $(seq 0 $((N-1))); do
  ( CUDA_VISIBLE_DEVICES=$i python tools/pcie_probe.py | tail -1 | sed "s/^/gpu $i: /" ) &
done
wait
