#!/bin/bash
# second chance + threshold ranks: before / after lines, then the full suite
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-config3 --no-gpu-baseline --steps 20 --warmup 3"
r() { name=$1; shift; env "$@" timeout 600 $B --workload $W > gpurun_out/$name.log 2> gpurun_out/$name.err; python scripts/brief.py gpurun_out/$name.log; tail -2 gpurun_out/$name.err; }
for W in c2 c3shard c5shard c4slab c1 c5l2; do
r y_${W}_old NWK_SAMPLE_M=32 NWK_SAMPLE_STRIDE=32
r y_${W}_new NWK_X=0
done
W=c5shard
r y_c5_r8 NWK_SAMPLE_M=8
r y_c5_r16 NWK_SAMPLE_M=16
PT="python -m pytest -q -m gpu --timeout 900 -p no:cacheprovider"
timeout 1500 $PT tests -x 2>&1 | tail -5
