#!/bin/bash
# --set full capture of the MaxSim kernels (one GPU's share of config 5)
CMD="python scripts/bench_maxsim.py 1000 32 12500 180 128 100 1"
$CMD > gpurun_out/ncu_plain_ms.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"maxsim_kernel|ms_exact_kernel" -s 6 -c 2 -f -o gpurun_out/prof_maxsim $CMD > gpurun_out/ncu_full_ms.log 2>&1
echo "ncu full exit $?"
