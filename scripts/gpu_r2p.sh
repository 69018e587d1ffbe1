#!/bin/bash
# 2 GPUs: the tests that need two devices (NCCL chain, peer exchange, rank-aware entry points, nwk_mg_*),
# then the driver's own N=2 command line (c2 weak + config3 strong + e2e)
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu --timeout 600 -p no:cacheprovider"
timeout 600 $PT tests/test_gpu_configs.py -k "world2 or single_process" 2>&1 | tail -4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
S=$SECONDS
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo "bench exit $? in $((SECONDS-S)) s"
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/bench_2gpu.log") if l.startswith("{")][-1])
print("ms/step", j["ms_per_step"], "value", j["value"], "clocks", j["clocks"])
print("exchange", j["exchange"], j["exchange_ms_per_step"], j["local_knn_ms"])
print("e2e", {k: j["e2e"][k] for k in ("value", "ms_per_step", "step_ms", "h2d_gb_per_s_per_gpu")})
c3 = j["config3"]
print("c3", c3["ms_per_step"], c3["frac_of_n_x_burst"], c3["exchange_ms"], c3["sampled_queries_equal_fp64_scan"])
PY
S=$SECONDS
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_ref.log 2> gpurun_out/bench_2gpu_ref.err; echo "reference exit $? in $((SECONDS-S)) s"; cut -c1-400 gpurun_out/bench_2gpu_ref.log
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$\|NCCL version" gpurun_out/bench_2gpu.err | tail -5 | cut -c1-300
