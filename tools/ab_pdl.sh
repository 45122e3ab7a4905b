#!/bin/bash
# A/B of the programmatic-dependent-launch policy on one box: 0 = plain launches, 1 = small grids only (default), 2 = every launch
for rep in 1 2; do for m in 0 1 2; do
  GLBA_PDL=$m timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/ab_pdl_${m}_${rep}.json 2>/dev/null
done; done
python - <<'PY'
import json
for rep in (1, 2):
    for m in (0, 1, 2):
        d = json.load(open(f'gpurun_out/ab_pdl_{m}_{rep}.json'))
        print(f"PDL={m} rep{rep} step {d['ms_per_step']:.4f} ms  lm {d['lm']['lm_iters_per_s']:.1f} it/s  t_solve {d['lm']['device_ms']['t_solve_ms']:.1f}  c2 {d['window']['solve_ms']:.3f} ms  c3 {d['window']['c3_sliding_windows']['solve_s']:.4f} s  e2e {d['e2e']['ms_per_step']:.3f} ms  pose {d['window']['pose_only_500pts_ms']:.3f}")
PY
