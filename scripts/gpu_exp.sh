#!/bin/bash
# Bench lines under experiment switches, on the experiment build of the library
# (python -c "from neighborhoodwatch_b200 import _build; _build.build(defines=('NWK_EXPERIMENTS',), out_path='neighborhoodwatch_b200/libnwk_exp.so')"
#  before calling gpurun).  EXTRA="NWK_PAIR=1:c5shard NWK_DEBUG_EPI=2,NWK_DEBUG_PROF=1:c5shard" STEPS=5
#   NWK_PAIR / NWK_DIRECT / NWK_ARES      choose the filter variant (also read by the product library)
#   NWK_DEBUG_EPI=1|2                      no key passes | the epilogue does not read the accumulators (results are wrong)
#   NWK_DEBUG_NA=1                         a single resident query buffer
#   NWK_DEBUG_PROF=1                       the MMA issuer's wait / issue cycles per tile on stderr
export NWK_LIB_PATH=$PWD/neighborhoodwatch_b200/libnwk_exp.so
for spec in $EXTRA; do
  envs=${spec%%:*}; w=${spec##*:}
  env ${envs//,/ } timeout 600 python bench.py --workload $w --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/exp.err | python -c "
import json,sys
j=json.loads(sys.stdin.readlines()[-1]); s=j['roofline']['stage_ms_mean']
print('$spec', 'step', round(j['ms_per_step'],3), {k: round(v,3) for k,v in s.items()}, j['clocks']['sm_mhz'], j['clocks']['power_w_max'])"
  grep "nwk prof" gpurun_out/exp.err | tail -1
done
