#!/bin/bash
# First GPU pass: every group in its own process so one faulting kernel cannot poison the rest.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name: $*" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log; }
PT="python -m pytest -q -m gpu --timeout 600 -p no:cacheprovider"
run t_basic $PT tests/test_gpu_parity.py -k "synth or exact_scan or merge or error_behaviour"
run t_tile $PT tests/test_gpu_parity.py -k "tc_tile" -s
run t_tc $PT tests/test_gpu_parity.py -k "tc_knn or gaussian or duplicates or uncertifiable" -s
run t_rest $PT tests/test_gpu_parity.py -k "running or host_entry or large_shape" -s
run t_drivers $PT tests/test_gpu_drivers.py
run smoke python __graft_entry__.py smoke
run bench_tiny python bench.py --workload tiny --steps 3 --warmup 3 --no-cpu-baseline
run bench_c1 python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu-baseline
run bench_c2 python bench.py --workload c2 --steps 5 --warmup 3
