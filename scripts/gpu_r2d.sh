#!/bin/bash
# 2 GPUs: peer-memory exchange vs packed all-gather (tests + bench lines)
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name: $*" ; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "exit $?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-800}; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/$name.err | tail -n 5 | cut -c1-400; }
PT="python -m pytest -q -m gpu --timeout 600 -p no:cacheprovider"
TAILN=14 run t_world2 $PT tests/test_gpu_configs.py -x -s -k "world2"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
B="bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e"
run bench_2gpu_peer $TR $B
python scripts/brief.py gpurun_out/bench_2gpu_peer.log
python - <<'PY'
import json
for n in ("bench_2gpu_peer",):
    j = json.loads([l for l in open(f"gpurun_out/{n}.log") if l.startswith("{")][-1])
    print(n, "exchange", j["exchange"], j["exchange_ms_per_step"], "local", j["local_knn_ms"], "c3", {k: j["config3"][k] for k in ("ms_per_step", "exchange_ms", "exchange", "frac_of_n_x_burst", "sampled_queries_equal_fp64_scan", "local_ms_max_over_ranks", "local_ms_min_over_ranks")})
PY
NWK_EXCHANGE=allgather run bench_2gpu_ag $TR $B --no-config3
python - <<'PY'
import json
j = json.loads([l for l in open("gpurun_out/bench_2gpu_ag.log") if l.startswith("{")][-1])
print("allgather: ms/step", j["ms_per_step"], "exchange", j["exchange"], j["exchange_ms_per_step"], "local", j["local_knn_ms"])
PY
