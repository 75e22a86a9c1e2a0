#!/bin/bash
# one GPU session: tests, the contract bench, pipeline-depth A/B on the FSK workload
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/gputests_e.log; cat gpurun_out/gputests_e.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; tail -3 gpurun_out/bench_e.err
for g in 4 8 16; do
  echo "== fsk pipeline $g"
  python bench.py --workload fsk_cs16_1024k --pipeline $g --steps 2 --warmup 1 --no-cpu-baseline --parity-streams 0 --no-gates 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['e2e']['value'], d['e2e']['breakdown_ms'])"
done
