python -c "import torch" 2>/dev/null
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_detect -c 1 -o gpurun_out/detect_r01z python tools/quick_perf.py --streams 4096 --distinct 32 --iters 1 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_slice -c 1 -o gpurun_out/slice_r01z python tools/quick_perf.py --streams 4096 --distinct 32 --iters 1 2>&1 | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01z.csv python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1; tail -2 gpurun_out/launches_r01z.csv | cut -c1-200
