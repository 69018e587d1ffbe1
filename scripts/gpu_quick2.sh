#!/bin/bash
# quick pass + experiment switches: EXTRA="NWK_PAIR=1:c5shard NWK_DIRECT=0:c5shard" runs bench lines with env overrides
bash scripts/gpu_quick.sh
for spec in $EXTRA; do
  envs=${spec%%:*}; w=${spec##*:}
  env ${envs//,/ } timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.readlines()[-1]); s=j['roofline']['stage_ms_mean']
print('$spec', 'step', round(j['ms_per_step'],3), {k: round(v,3) for k,v in s.items()})"
done
