#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu --timeout 300 -p no:cacheprovider -x"
timeout 600 $PT tests/test_gpu_parity.py tests/test_gpu_drivers.py > gpurun_out/q_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/q_tests.log
for v in 1 0; do
  for w in c2 c3shard c5shard c1; do
  export NWK_PAIR=$v
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/q_bench_$w.err | python -c "
import json,sys
j=json.loads(sys.stdin.readlines()[-1]); s=j['roofline']['stage_ms_mean']
print('pair=$v', '$w', 'step', round(j['ms_per_step'],3), {k: round(v,3) for k,v in s.items()}, j['exactness']['mean_candidates_per_query'], j['exactness']['fallback_queries_per_step'], j['clocks']['sm_mhz'])"
  done
done
unset NWK_PAIR
