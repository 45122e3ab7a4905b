#!/bin/bash
# A/B of the two PCG kernel forms on C4 (one GPU): time per PCG iteration, LM rate, per-phase cycles of CTA 0.
#   GLBA_CG_REG=1 (default)  one row per warp, vectors in registers, a CTA's blocks in shared memory, rows balanced by entry count
#   GLBA_CG_REG=0            general kernel (several rows per warp, everything in global memory / L2)
cd "$(dirname "$0")/.."
for reg in 1 0; do
  GLBA_CG_REG=$reg python tools/time_explicit.py C4 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)['explicit']
print('GLBA_CG_REG=$reg  pcg_iteration_us %.2f  lm_it_per_s %.1f  t_solve_ms %.3f' % (1e3*d['kernels']['bsr_spmv_ms'], d['lm_iters_per_s'], d['device_ms']['t_solve_ms']))"
  GLBA_CG_PROF=1 GLBA_CG_REG=$reg python tools/time_explicit.py C4 2>&1 >/dev/null | grep "k_cg_bsr" | tail -1
done
