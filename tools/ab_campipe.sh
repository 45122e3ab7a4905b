#!/bin/bash
# A/B of the fused camera-major kernel (glba_campipe.cuh) on C4, one GPU: step time and LM rate.
#   GLBA_CAMPIPE=0  separate k_linearize_cm / k_schur_cm;  GLBA_CP_OCC=2|3  register budget of k_cam_pipe
cd "$(dirname "$0")/.."
for cfg in "1 2" "1 3" "0 2"; do
  set -- $cfg
  GLBA_CAMPIPE=$1 GLBA_CP_OCC=$2 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('GLBA_CAMPIPE=$1 GLBA_CP_OCC=$2  step_ms %.4f  G_obs_per_s %.2f  lm_it_per_s %.1f  cost %.10e' % (d['ms_per_step'], d['value']/1e9, d['lm']['lm_iters_per_s'], d['cost_at_initial_point']))"
done
