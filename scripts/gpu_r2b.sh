#!/bin/bash
# round 2, second pass (2 GPUs): real NCCL chain + rank-aware entry points, bench at N=2 (c2 + config 3), c4 streaming
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name: $*" ; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "exit $?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-800}; tail -n 5 gpurun_out/$name.err | cut -c1-400; }
bash scripts/gpu_probe.sh > /dev/null; head -12 gpurun_out/probe.txt
PT="python -m pytest -q -m gpu --timeout 900 -p no:cacheprovider"
TAILN=12 run t_world2 $PT tests/test_gpu_configs.py -x -s -k "world2"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
CUT=7000 run bench_2gpu $TR bench.py --gpus 2 --steps 10 --warmup 3
CUT=4000 run bench_c4_2gpu $TR bench.py --gpus 2 --workload c4 --c4-rows-per-gpu 4000000 --steps 2
CUT=4000 run bench_c4_1gpu python bench.py --workload c4 --steps 2
CUT=1500 run bench_ref_2gpu $TR bench.py --impl reference --gpus 2 --steps 1 --warmup 1
