#!/bin/bash
# 2 GPUs: peer exchange v2 (slice merge + broadcast) tests and timing, H2D probe
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name: $*" ; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>gpurun_out/$name.err; echo "exit $?" >> gpurun_out/$name.log; tail -n ${TAILN:-3} gpurun_out/$name.log | cut -c1-${CUT:-600}; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$\|NCCL version" gpurun_out/$name.err | tail -n 5 | cut -c1-400; }
PT="python -m pytest -q -m gpu --timeout 600 -p no:cacheprovider"
TAILN=6 run t_world2 $PT tests/test_gpu_configs.py -x -k "world2"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run h2d_probe $TR scripts/h2d_probe.py
B="bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-config3"
for w in c2 c5shard; do
run bench_2gpu_peer_$w $TR $B --workload $w
NWK_EXCHANGE=allgather run bench_2gpu_ag_$w $TR $B --workload $w
python - <<PY
import json
for n in ("peer", "ag"):
    j = json.loads([l for l in open(f"gpurun_out/bench_2gpu_{n}_$w.log") if l.startswith("{")][-1])
    print("$w", n, "ms/step", round(j["ms_per_step"], 3), j["exchange"], "exchange_ms", j["exchange_ms_per_step"], "local", j["local_knn_ms"])
PY
done
