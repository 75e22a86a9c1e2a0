#!/bin/bash
# A/B of k_slice2 build variants / staging sizes on the GPU box (tuning only): same cached synthetic streams for every run
mkdir -p gpurun_out
for v in "" rtl_433_b200/csrc/variants/lib_slice7.so rtl_433_b200/csrc/variants/lib_slice9.so; do
    for g in "" "--gates"; do
      echo "== lib=${v:-default} $g"
      R433B_LIB=${v:+$PWD/$v} python tools/quick_perf.py --streams 4096 --distinct 64 --iters 3 $g 2>&1 | grep "iter 2\|rror" | head -3
    done
done
for w in 256 512 2048; do
    for g in "" "--gates"; do
      echo "== stage_words=$w $g"
      R433B_STAGE_WORDS=$w python tools/quick_perf.py --streams 4096 --distinct 64 --iters 3 $g 2>&1 | grep "iter 2\|rror" | head -3
    done
done
