#!/bin/bash
# 1 GPU, EXPERIMENT build (libnwk_exp.so, -DNWK_EXPERIMENTS: the product library does not read these switches;
#   python -c "from neighborhoodwatch_b200 import _build; _build.build(defines=('NWK_EXPERIMENTS',), out_path='neighborhoodwatch_b200/libnwk_exp.so')"
# before calling gpurun): where the d = 128 filter pass spends its time.  NWK_DEBUG_EPI: 1 = no key passes (fast path only),
# 3 = the accumulators are read from TMEM and dropped, 2 = the epilogue does not touch them; NWK_DEBUG_PROF = the MMA issuer's
# cycles per tile.  (The run recorded in profiles/r02_summary.md also had NWK_EPI_INLINE -- hits recorded from the fast path's
# registers, the build of commit "Four-warpgroup epilogue: hits recorded from the fast path" -- which lost 12 % and was removed.)
mkdir -p gpurun_out
export NWK_LIB_PATH=$PWD/neighborhoodwatch_b200/libnwk_exp.so
run() { envs=$1; w=$2;
  env ${envs//,/ } timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-config3 --no-gpu-baseline 2>gpurun_out/u.err | python -c "
import json,sys
j=json.loads(sys.stdin.readlines()[-1]); s=j['roofline']['stage_ms_mean']
print('$envs $w', 'filter', round(s['filter'],3), 'step', round(j['ms_per_step'],3), 'clk', j['clocks']['sm_mhz'], 'cand', round(j['exactness']['mean_candidates_per_query'],1))"
  grep "nwk prof" gpurun_out/u.err | tail -1; }
run NWK_DEBUG_EPI=0 c5shard
run NWK_DEBUG_EPI=0,NWK_DEBUG_PROF=1 c5shard
run NWK_DEBUG_EPI=1,NWK_DEBUG_PROF=1 c5shard
run NWK_DEBUG_EPI=3,NWK_DEBUG_PROF=1 c5shard
run NWK_DEBUG_EPI=2,NWK_DEBUG_PROF=1 c5shard
