#!/bin/bash
# compare experiment builds of the filter epilogue (NWK_LIB_PATH) on the stage timings of three shapes
mkdir -p gpurun_out
for lib in neighborhoodwatch_b200/exp_*.so; do
  for w in c5shard c1 c2; do
    NWK_LIB_PATH=$PWD/$lib python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.readlines()[-1]); s=j['roofline']['stage_ms_mean']
print('$lib'.split('exp_')[1], '$w', 'step', round(j['ms_per_step'],3), 'sample', round(s['sample'],3), 'filter', round(s['filter'],3))"
  done
done
