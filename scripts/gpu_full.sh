#!/bin/bash
# full GPU pass for a finished change: tests, smoke, bench lines of every shape, then the ncu evidence
# (launch list + full capture of the hot kernels) for config 2
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name: $*" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-600; }
PT="python -m pytest -q -m gpu --timeout 600 -p no:cacheprovider"
run t_gpu $PT tests
run smoke python __graft_entry__.py smoke
run bench_c2 python bench.py --workload c2 --steps 20 --warmup 3
for w in c1 c3shard c5shard c4slab; do run bench_$w python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e; done
bash scripts/gpu_ncu.sh c2
