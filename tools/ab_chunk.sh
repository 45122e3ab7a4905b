show='
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("value",d["value"],"ms",d["ms_per_step"]); print({k:round(v["ms"],4) for k,v in d["kernels"].items()})'
for c in 4096 2048 1024 512; do echo "== chunk $c"; GLBA_CHUNK=$c timeout 200 python bench.py --steps 20 --warmup 3 --lm-iters 0 --no-cpu-baseline 2>/dev/null | python -c "$show"; done
