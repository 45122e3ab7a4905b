#!/bin/bash
# ncu --set full of the small-window (dense path) kernels on C2, exported as CSV on the box.  usage: bash tools/ncu_small.sh <tag>
tag=${1:-r02}
CMD="python tools/time_c2.py"
$CMD > gpurun_out/${tag}_small_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_small_plain.log; exit 1; }
for k in k_dense_schur k_dense_solve k_dense_reduce k_cam_lin_fin; do
  ncu --set full --import-source on --clock-control none -k "regex:$k" -s 6 -c 1 -f -o /tmp/${tag}_$k $CMD > gpurun_out/${tag}_ncu_$k.log 2>&1
  ncu -i /tmp/${tag}_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_${k}_raw.csv 2>/dev/null
  ncu -i /tmp/${tag}_$k.ncu-rep --page source --csv > gpurun_out/${tag}_${k}_source.csv 2>/dev/null
done
ls -la gpurun_out/${tag}_k_*
