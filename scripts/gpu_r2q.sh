#!/bin/bash
# 1 GPU: stage timings after the short-row fix of baseprep (c5shard, c1) and a re-measurement of the finish stage
# with the F2F rerank (c4slab, c2) on one box
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu --timeout 600 -p no:cacheprovider -x"
timeout 600 $PT tests/test_gpu_parity.py -k "parity or edge or short" 2>&1 | tail -2
for w in ${WORKLOADS:-c5shard c5l2 c1 c2 c4slab}; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-config3 --no-gpu-baseline 2>gpurun_out/q_bench_$w.err | tee gpurun_out/q_bench_$w.log | python -c "
import json,sys
j=json.loads(sys.stdin.readlines()[-1]); s=j['roofline']['stage_ms_mean']
print('$w', 'step', round(j['ms_per_step'],3), {k: round(v,3) for k,v in s.items()})"
done
