#!/bin/bash
# ncu evidence for the current build: launch list of a bench run + full captures of the two hot kernels
mkdir -p gpurun_out
CMD="python bench.py --workload c2 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_kernel|finish_kernel|baseprep" -s 12 -c 4 -o gpurun_out/prof_hot $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
tail -3 gpurun_out/ncu_full.log | cut -c1-300
python bench.py --workload c2 --steps 20 --warmup 3 > gpurun_out/bench_c2_full.log 2>&1; tail -1 gpurun_out/bench_c2_full.log | cut -c1-1500
