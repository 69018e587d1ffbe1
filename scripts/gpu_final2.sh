#!/bin/bash
# 2 GPUs: the driver's N=2 command on the final build (c2 weak + config 3 strong + e2e)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo "bench exit $? @$SECONDS"
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/bench_2gpu.log") if l.startswith("{")][-1])
print("ms/step", j["ms_per_step"], "exchange", j["exchange"], j["exchange_ms_per_step"], j["local_knn_ms"])
print("e2e", j["e2e"]["ms_per_step"], j["e2e"]["matches_resident_result"])
c3 = j["config3"]; print("c3", c3["ms_per_step"], c3["frac_of_n_x_burst"], c3["sampled_queries_equal_fp64_scan"])
PY
