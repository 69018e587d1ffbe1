#!/bin/bash
# round 2, first pass: box probe, the whole GPU suite (new config-3 / config-4 parity tests included), bench line
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name: $*" ; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-800}; }
bash scripts/gpu_probe.sh | tail -40
PT="python -m pytest -q -m gpu --timeout 900 -p no:cacheprovider"
TAILN=15 run t_new $PT tests/test_gpu_configs.py -x -s
run t_gpu $PT tests --deselect tests/test_gpu_configs.py
run smoke python __graft_entry__.py smoke
CUT=6000 run bench_c2 python bench.py --steps 10 --warmup 3
