#!/bin/bash
# final check of a tree: GPU tests (writes gpurun_out/dispatch_real_decoders.json), smoke, both bench arms
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-120
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_f_reference.json 2> gpurun_out/bench_f_reference.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; tail -3 gpurun_out/bench_f.err
cat gpurun_out/dispatch_real_decoders.json
