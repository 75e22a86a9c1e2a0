python -c "import torch" 2>/dev/null
timeout 600 compute-sanitizer --tool memcheck python tools/quick_perf.py --streams 96 --distinct 8 --iters 1 2>&1 | tail -4
timeout 600 compute-sanitizer --tool memcheck python tools/quick_perf.py --streams 48 --distinct 8 --fsk --iters 1 2>&1 | tail -3
timeout 900 compute-sanitizer --tool racecheck python tools/quick_perf.py --streams 16 --distinct 8 --log2n 18 --iters 1 2>&1 | tail -3
