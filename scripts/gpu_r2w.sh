#!/bin/bash
# 1 GPU, product build: the existing CTA-pair filter (query tile re-fetched per base tile) on the short-row shape,
# against the default single-CTA kernel with the resident query tile (NWK_PAIR is read by the product library)
mkdir -p gpurun_out
run() { envs=$1; w=$2;
  env ${envs//,/ } timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-config3 --no-gpu-baseline 2>gpurun_out/w.err | python -c "
import json,sys
j=json.loads(sys.stdin.readlines()[-1]); s=j['roofline']['stage_ms_mean']
print('$envs $w', 'filter', round(s['filter'],3), 'step', round(j['ms_per_step'],3), 'clk', j['clocks']['sm_mhz'], 'fb', j['exactness']['fallback_queries_per_step'])"; }
run NWK_PAIR=0 c5shard
run NWK_PAIR=1 c5shard
run NWK_PAIR=0,NWK_ARES=0 c5shard
