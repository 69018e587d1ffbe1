#!/bin/bash
# 1 GPU: the four-warpgroup direct-key epilogue with hits recorded from the fast path (NWK_EPI_INLINE=1, default)
# against the slow loop over TMEM (=0): parity + fuzz, then the short-row shapes on one box
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu --timeout 600 -p no:cacheprovider -x"
timeout 700 $PT tests/test_gpu_parity.py tests/test_gpu_fuzz.py -k "not maxsim" 2>&1 | tail -3
for inl in 1 0 1 0; do
for w in ${WORKLOADS:-c5shard c5l2}; do
  NWK_EPI_INLINE=$inl timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-config3 --no-gpu-baseline 2>gpurun_out/s_bench_$w.err | tee gpurun_out/s_bench_${w}_$inl.log | python -c "
import json,sys
j=json.loads(sys.stdin.readlines()[-1]); s=j['roofline']['stage_ms_mean']
print('inline=$inl $w', 'step', round(j['ms_per_step'],3), {k: round(v,3) for k,v in s.items()}, 'cand', round(j['exactness']['mean_candidates_per_query'],1), 'fb', j['exactness']['fallback_queries_per_step'], 'clk', j['clocks']['sm_mhz'])"
done
done
