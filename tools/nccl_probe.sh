#!/bin/bash
# Diagnostic: where the sharded step's time goes (kernel sum vs all-reduce vs host enqueue), N=$1 GPUs.
N=${1:-2}
show='
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("value",d["value"],"ms",d["ms_per_step"],"host_enqueue",d["host_enqueue_ms_per_step"],"e2e",(d.get("e2e") or {}).get("value")); print({k:round(v["ms"],4) for k,v in d["kernels"].items()})'
echo "== N=1 scale 0.25"; timeout 200 python bench.py --steps 20 --warmup 3 --scale 0.25 --lm-iters 0 --no-cpu-baseline 2>/dev/null | python -c "$show"
echo "== N=$N scale 0.25"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --scale 0.25 --lm-iters 0 2>gpurun_out/probe_err.log | python -c "$show"
