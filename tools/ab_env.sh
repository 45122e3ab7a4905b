#!/bin/bash
# A/B of one diagnostic environment knob on C4 (per-kernel CUDA-event times): bash tools/ab_env.sh GLBA_PREFETCH 0 1
V=$1; shift
show='
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("value %.4e step ms %.4f lm it/s %.1f" % (d["value"], d["ms_per_step"], d["lm"]["lm_iters_per_s"]), {k:round(v["ms"],4) for k,v in d["kernels"].items()})'
for rep in 1 2; do for x in "$@"; do echo -n "$V=$x  "; env $V=$x timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$show"; done; done
