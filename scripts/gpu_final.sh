#!/bin/bash
# final check of a build on one GPU, as the driver does it: all GPU tests, smoke(), the bench line and the reference arm
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name @$SECONDS: $*" ; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "exit $?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-400}; tail -n 2 gpurun_out/$name.err | cut -c1-300; }
PT="python -m pytest -q -m gpu --timeout 600 -p no:cacheprovider"
TAILN=4 run t_gpu $PT tests
run smoke python __graft_entry__.py smoke
CUT=600 run bench_c2 python bench.py --gpus 1 --steps 20 --warmup 5
CUT=300 run bench_ref python bench.py --impl reference --gpus 1 --steps 20 --warmup 5
echo "done @$SECONDS"
