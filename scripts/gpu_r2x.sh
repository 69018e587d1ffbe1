#!/bin/bash
# 1 GPU: document tokens per lane of the MaxSim re-scoring kernel (NWK_MSX_PW): conversions amortised against CTAs per SM
mkdir -p gpurun_out
for pw in 16 8 4 16 8; do
  NWK_MSX_PW=$pw timeout 200 python scripts/bench_maxsim.py 1000 32 12500 180 128 100 10 2>gpurun_out/x.err | python -c "
import json,sys
j=json.loads(sys.stdin.readlines()[-1]); s=j['stats']
print('pw=$pw', 'ms', round(j['ms_per_step'],2), 'finish', round(s['ms_finish'],2), 'filter', round(s['ms_filter'],2))"
done
