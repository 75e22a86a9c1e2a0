#!/bin/bash
# A/B of k_slice formulations and build variants on the GPU box (tuning only): same cached synthetic streams for every run
mkdir -p gpurun_out
for v in "" rtl_433_b200/csrc/variants/lib_slice8.so rtl_433_b200/csrc/variants/lib_slice10.so; do
  for v1 in 0 1; do
    for g in "" "--gates"; do
      echo "== lib=${v:-default} SLICE_V1=$v1 $g"
      R433B_SLICE_V1=$v1 R433B_LIB=${v:+$PWD/$v} python tools/quick_perf.py --streams 4096 --distinct 64 --iters 3 $g 2>&1 | grep "iter 2\|rror" | head -3
    done
  done
done
echo "== fsk default"
python tools/quick_perf.py --fsk --streams 1024 --distinct 32 --iters 3 2>&1 | grep "iter 2\|rror"
