#!/bin/bash
# multi-GPU check: usage gpu_session_n.sh N [extra bench args]
N=$1; shift
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 "$@" > gpurun_out/bench_n$N$TAG.json 2> gpurun_out/bench_n$N.err
[ -n "$TAG" ] && cp gpurun_out/bench_n$N$TAG.json gpurun_out/bench_n$N.json
tail -5 gpurun_out/bench_n$N.err
python - <<PY
import json
l=[json.loads(x) for x in open("gpurun_out/bench_n$N.json") if x.startswith("{")][-1]
print(l["value"], l["n_gpus"], l["kernel_ms"], l["e2e"]["value"], l["e2e_ungated"]["value"], l["parity_checked"]["ok"], l["config"]["host_binding"], l["config"]["streams_total"])
PY
