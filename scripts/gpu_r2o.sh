#!/bin/bash
# 1 GPU: the rebalanced MaxSim re-scoring kernel -- its tests and the MaxSim bench at both shapes
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu --timeout 600 -p no:cacheprovider -x"
timeout 600 $PT tests/test_gpu_parity.py tests/test_gpu_fuzz.py -k "maxsim" 2>&1 | tail -4
timeout 300 python scripts/bench_maxsim.py 1000 32 12500 180 128 100 10 > gpurun_out/maxsim_shard.json 2> gpurun_out/maxsim_shard.err; cat gpurun_out/maxsim_shard.json; tail -2 gpurun_out/maxsim_shard.err
timeout 300 python scripts/bench_maxsim.py 1000 32 100000 180 128 100 5 > gpurun_out/maxsim_full.json 2> gpurun_out/maxsim_full.err; cat gpurun_out/maxsim_full.json; tail -2 gpurun_out/maxsim_full.err
