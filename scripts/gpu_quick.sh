#!/bin/bash
# quick check of a kernel change: parity tests, then the stage timings of the bench shapes
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu --timeout 600 -p no:cacheprovider -x"
timeout 900 $PT tests/test_gpu_parity.py tests/test_gpu_drivers.py > gpurun_out/q_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/q_tests.log
for w in ${WORKLOADS:-c2 c1 c3shard c5shard}; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/q_bench_$w.err | tee gpurun_out/q_bench_$w.log | python -c "
import json,sys
j=json.loads(sys.stdin.readlines()[-1]); s=j['roofline']['stage_ms_mean']
print('$w', 'step', round(j['ms_per_step'],3), {k: round(v,3) for k,v in s.items()}, j['exactness'])"
done
