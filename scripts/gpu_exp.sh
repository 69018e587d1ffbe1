#!/bin/bash
# bench lines under experiment switches: EXTRA="NWK_PAIR=1:c5shard NWK_DEBUG_EPI=2,NWK_PAIR=0:c5shard"
for spec in $EXTRA; do
  envs=${spec%%:*}; w=${spec##*:}
  env ${envs//,/ } timeout 600 python bench.py --workload $w --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/exp.err | python -c "
import json,sys
j=json.loads(sys.stdin.readlines()[-1]); s=j['roofline']['stage_ms_mean']
print('$spec', 'step', round(j['ms_per_step'],3), {k: round(v,3) for k,v in s.items()}, j['clocks']['sm_mhz'], j['clocks']['power_w_max'])"
  grep "nwk prof" gpurun_out/exp.err | tail -1
done
