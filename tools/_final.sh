python -c "import torch" 2>/dev/null
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/quick_perf.py --streams 1024 --distinct 32 --fsk --iters 2 2>&1 | tail -2 | head -1 | cut -c1-120
timeout 300 python tools/quick_perf.py --streams 4096 --distinct 32 --iters 2 2>&1 | tail -2 | head -1 | cut -c1-120
