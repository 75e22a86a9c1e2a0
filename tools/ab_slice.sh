#!/bin/bash
# A/B of k_slice build variants on the GPU box (tuning only): same cached synthetic streams for every library
mkdir -p gpurun_out
for v in "" rtl_433_b200/csrc/variants/lib_slice8.so rtl_433_b200/csrc/variants/lib_slice10.so; do
  for g in "" "--gates"; do
    echo "== lib=${v:-default} $g"
    R433B_LIB=$v python tools/quick_perf.py --streams 4096 --distinct 64 --iters 3 $g 2>&1 | grep "iter 2"
  done
done
