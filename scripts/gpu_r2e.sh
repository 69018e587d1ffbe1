#!/bin/bash
# 8 GPUs: the driver's scaling line (c2 weak + config 3 strong), config 4 streamed, config 5 shapes, exchange comparison
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name: $*" ; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "exit $?" >> gpurun_out/$name.log; tail -n ${TAILN:-2} gpurun_out/$name.log | cut -c1-${CUT:-300}; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$\|NCCL version" gpurun_out/$name.err | tail -n 5 | cut -c1-400; }
bash scripts/gpu_probe.sh > /dev/null; head -14 gpurun_out/probe.txt; cp gpurun_out/probe.txt gpurun_out/probe_8gpu.txt
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run bench_${N}gpu $TR bench.py --gpus $N --steps 10 --warmup 3
python scripts/brief.py gpurun_out/bench_${N}gpu.log
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/bench_${N}gpu.log") if l.startswith("{")][-1])
print("exchange", j["exchange"], j["exchange_ms_per_step"], "local", j["local_knn_ms"])
print("e2e", {k: j["e2e"][k] for k in ("ms_per_step", "step_ms", "h2d_gb_per_s_per_gpu", "matches_resident_result")})
print("c3", {k: j["config3"][k] for k in ("ms_per_step", "exchange_ms", "exchange", "frac_of_n_x_burst", "aggregate_tflops", "sampled_queries_equal_fp64_scan", "local_ms_max_over_ranks", "local_ms_min_over_ranks", "fallback_queries_per_step")})
PY
CUT=900 run h2d_probe_${N}gpu $TR scripts/h2d_probe.py
WC=$(python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/h2d_probe_${N}gpu.log") if l.startswith("{")][-1])
    print("--c4-host-wc" if j["write_combined"]["gb_per_s_per_gpu_min"] > 1.1 * j["pinned"]["gb_per_s_per_gpu_min"] else "")
except Exception:
    print("")
PY
)
echo "c4 host memory flag: [$WC]"
TMO=900 run bench_c4_${N}gpu $TR bench.py --gpus $N --workload c4 --steps 2 $WC
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/bench_c4_${N}gpu.log") if l.startswith("{")][-1])
print("c4", j["config"]["workload"], j["config"]["note"])
print("c4 e2e", j["e2e"])
print("c4 resident", j["resident"], "fill_s", j["fill_seconds"])
PY
NWK_EXCHANGE=allgather run bench_${N}gpu_ag $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-config3
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/bench_${N}gpu_ag.log") if l.startswith("{")][-1])
print("allgather: ms/step", j["ms_per_step"], "exchange", j["exchange"], j["exchange_ms_per_step"], "local", j["local_knn_ms"])
PY
run bench_c5_${N}gpu $TR bench.py --gpus $N --workload c5shard --steps 10 --warmup 3 --no-e2e --no-config3
python scripts/brief.py gpurun_out/bench_c5_${N}gpu.log
CUT=1200 run maxsim_${N}gpu $TR scripts/bench_maxsim.py 1000 32 12500 180 128 100 5
