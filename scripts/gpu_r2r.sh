#!/bin/bash
# 1 GPU: full verification of the build (all GPU tests, smoke, the driver's bench line), c5shard stage timings,
# then ncu: launch list of a c5shard bench run and full captures of the kernels this session changed
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name @$SECONDS: $*" ; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "exit $?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-400}; tail -n 2 gpurun_out/$name.err | cut -c1-300; }
PT="python -m pytest -q -m gpu --timeout 600 -p no:cacheprovider"
TAILN=4 run t_gpu $PT tests
run smoke python __graft_entry__.py smoke
CUT=1200 run bench_c2 python bench.py --gpus 1 --steps 20 --warmup 5
CMD5="python bench.py --workload c5shard --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-config3 --no-gpu-baseline"
CUT=3000 run bench_c5 python bench.py --workload c5shard --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-config3 --no-gpu-baseline
echo "=== ncu @$SECONDS"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c5shard_final.csv $CMD5 > gpurun_out/ncu_list_c5.log 2>&1; echo "ncu list exit $?"
ncu --set full --clock-control none --import-source on -k regex:"baseprep|qmax" -s 8 -c 3 -f -o gpurun_out/r02_prof_c5prep $CMD5 > gpurun_out/ncu_full_c5.log 2>&1; echo "ncu c5 prep exit $?"
ncu -i gpurun_out/r02_prof_c5prep.ncu-rep --page raw --csv > gpurun_out/r02_ncu_raw_c5prep.csv 2>/dev/null
MS="python scripts/bench_maxsim.py 1000 32 12500 180 128 100 2"
$MS > gpurun_out/ms_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"ms_exact|ms_eq|ms_final" -s 4 -c 3 -f -o gpurun_out/r02_prof_msx $MS > gpurun_out/ncu_full_msx.log 2>&1; echo "ncu msx exit $?"
ncu -i gpurun_out/r02_prof_msx.ncu-rep --page raw --csv > gpurun_out/r02_ncu_raw_msx.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep.tmp; ls -la gpurun_out/*.ncu-rep 2>/dev/null; echo "done @$SECONDS"
