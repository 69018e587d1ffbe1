#!/bin/bash
# 2 GPUs: the default bench line with the config-4 leg forced on (small shard), c4 workload, world-2 tests
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name: $*" ; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "exit $?" >> gpurun_out/$name.log; tail -n ${TAILN:-2} gpurun_out/$name.log | cut -c1-${CUT:-300}; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$\|NCCL version" gpurun_out/$name.err | tail -n 5 | cut -c1-400; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run bench_2gpu_full $TR bench.py --gpus 2 --steps 20 --warmup 5 --config4-min-gpus 2 --c4-rows-per-gpu 3000000
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/bench_2gpu_full.log") if l.startswith("{")][-1])
print("ms/step", j["ms_per_step"], "clocks", j["clocks"])
print("exchange", j["exchange"], j["exchange_ms_per_step"], j["local_knn_ms"])
print("c3", j["config3"]["ms_per_step"], j["config3"]["frac_of_n_x_burst"])
c4 = j["config4"]
print("c4", c4["config"]["workload"]); print(c4["e2e"]); print(c4["resident"])
PY
PT="python -m pytest -q -m gpu --timeout 600 -p no:cacheprovider"
TAILN=4 run t_world2 $PT tests/test_gpu_configs.py -x -k "world2"
