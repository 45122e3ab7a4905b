#!/bin/bash
# Round-2 profiling pass on C4 (one GPU), everything exported as CSV on the box:
#   1. launch list (gpu__time_duration.sum, --clock-control none) of a short bench run
#   2. ncu --set full --import-source on of one launch of each hot kernel
# usage (under gpurun): bash tools/ncu_round2.sh <tag>
tag=${1:-r02}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --lm-iters 2"
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > /dev/null 2>&1
for k in "k_lin_pipe" "k_pt_pipe<.int.0>" "k_pt_pipe<.int.1>" "k_linearize_cm" "k_schur_cm" "k_spmv_cm" "k_schur_pairs" "k_cg_bsr" "k_cam_pipe"; do
  n=$(echo $k | tr -d '<>.' )
  skip=2
  [ "$k" = "k_cg_bsr" ] && skip=8      # past the 1- and 33-iteration launches of glba_time_kernels: the first PCG of the LM solve (40 iterations)
  ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$k" -s $skip -c 1 -f -o /tmp/${tag}_$n $CMD > gpurun_out/${tag}_ncu_$n.log 2>&1
  ncu -i /tmp/${tag}_$n.ncu-rep --page raw --csv > gpurun_out/${tag}_${n}_raw.csv 2>/dev/null
  ncu -i /tmp/${tag}_$n.ncu-rep --page source --csv > gpurun_out/${tag}_${n}_source.csv 2>/dev/null
done
ls -la gpurun_out/${tag}_*
