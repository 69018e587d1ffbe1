#!/bin/bash
# ncu evidence for the current build: launch list of a bench run + full captures of the hot kernels
mkdir -p gpurun_out
W=${1:-c2}
CMD="python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$W.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_kernel|select_kernel|rerank_kernel|output_kernel|baseprep|tau_merge|qprep" -s 24 -c 8 -f -o gpurun_out/prof_hot_$W $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
tail -2 gpurun_out/ncu_full.log | cut -c1-200
