#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-config3 --no-gpu-baseline --steps 20 --warmup 3"
r() { name=$1; shift; env "$@" timeout 600 $B --workload $W > gpurun_out/$name.log 2> gpurun_out/$name.err; python scripts/brief.py gpurun_out/$name.log; tail -2 gpurun_out/$name.err; }
for W in c2 c3shard; do
r g_${W}_nogate NWK_PREP_GATE=0
r g_${W}_gate NWK_PREP_GATE=1
r g_${W}_gate10 NWK_PREP_GATE=1 NWK_PART0_PCT=10
r g_${W}_gate33 NWK_PREP_GATE=1 NWK_PART0_PCT=33
r g_${W}_gate_p2 NWK_PREP_GATE=1 NWK_PARTS=2 NWK_PART0_PCT=15
r g_${W}_nogate2 NWK_PREP_GATE=0
done
