#!/bin/bash
# overlap experiments: parts / sample density, then the full suite
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-config3 --no-gpu-baseline --steps 20 --warmup 3"
r() { name=$1; shift; env "$@" timeout 600 $B --workload $W > gpurun_out/$name.log 2> gpurun_out/$name.err; python scripts/brief.py gpurun_out/$name.log; tail -2 gpurun_out/$name.err; }
W=c2
r x_c2_base NWK_PARTS=1 NWK_SAMPLE_M=32
r x_c2_m16 NWK_PARTS=1
r x_c2_p2 NWK_PARTS=2
r x_c2_p3 NWK_PARTS=3
r x_c2_p4 NWK_PARTS=4
r x_c2_base2 NWK_PARTS=1 NWK_SAMPLE_M=32
W=c3shard
r x_c3_base NWK_PARTS=1 NWK_SAMPLE_M=32
r x_c3_def NWK_X=0
W=c4slab
r x_c4_base NWK_PARTS=1 NWK_SAMPLE_M=32
r x_c4_def NWK_X=0
PT="python -m pytest -q -m gpu --timeout 900 -p no:cacheprovider"
timeout 1200 $PT tests -x 2>&1 | tail -5
