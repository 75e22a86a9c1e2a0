python -c "import torch" 2>/dev/null
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.json | cut -c1-1500
timeout 600 python bench.py --workload fsk_cs16_1024k --no-cpu-baseline > gpurun_out/bench_fsk.json 2> gpurun_out/bench_fsk.err; tail -1 gpurun_out/bench_fsk.json | cut -c1-600
