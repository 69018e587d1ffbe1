#!/bin/bash
# tests (separate processes), bench, then ncu launch list + full capture of the tensor-core kernels
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name: $*" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; }
PT="python -m pytest -q -m gpu --timeout 600 -p no:cacheprovider"
run t_tile $PT tests/test_gpu_parity.py -k "tc_tile or probe" -s
TAILN=40 run t_parity $PT tests/test_gpu_parity.py -k "not tc_tile and not probe" -s
run t_drivers $PT tests/test_gpu_drivers.py
run smoke python __graft_entry__.py smoke
run bench_c1 python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu-baseline
run bench_c2 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline
if [ "$1" == "ncu" ]; then
  CMD="python bench.py --workload c2 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
  $CMD > gpurun_out/ncu_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "ncu list exit $?"
  $CMD > gpurun_out/ncu_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:tc_kernel -s 6 -c 2 -o gpurun_out/prof_tc $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
  tail -5 gpurun_out/ncu_full.log
fi
