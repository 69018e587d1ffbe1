#!/bin/bash
bash scripts/gpu_ncu_r2.sh c2
SKIP=55 COUNT=11 bash scripts/gpu_ncu_r2.sh c5shard
cuobjdump -sass neighborhoodwatch_b200/libnwk.so | grep -E "UTCHMMA|LDTM|UTMALDG|UTCBAR|UBLKCP|Function :" > gpurun_out/r02_sass_grep.txt; wc -l gpurun_out/r02_sass_grep.txt
