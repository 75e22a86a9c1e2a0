#!/bin/bash
# evidence for profiles/: the contract bench (both arms), the launch list of the same command under ncu, one --set full
# capture of every kernel of a step
mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_f_reference.json 2> gpurun_out/bench_f_reference.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; tail -3 gpurun_out/bench_f.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_f.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --parity-streams 0 --no-secondary > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_front|k_detect|k_slice2' -c 3 -o gpurun_out/all_r02f -f \
    python tools/quick_perf.py --streams 4096 --distinct 256 --iters 1 --gates > gpurun_out/ncu_all.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_slice2' -c 1 -o gpurun_out/slice_ungated_r02f -f \
    python tools/quick_perf.py --streams 4096 --distinct 256 --iters 1 > gpurun_out/ncu_slice_ungated.log 2>&1
ls -la gpurun_out
