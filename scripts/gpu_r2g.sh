#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu --timeout 900 -p no:cacheprovider"
timeout 1500 $PT tests 2>&1 | tail -8
B="python bench.py --no-e2e --no-cpu-baseline --no-config3 --no-gpu-baseline --steps 20 --warmup 3"
for W in c1 c2; do timeout 600 $B --workload $W > gpurun_out/z_$W.log 2> gpurun_out/z_$W.err; python scripts/brief.py gpurun_out/z_$W.log; done
