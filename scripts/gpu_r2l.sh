#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu --timeout 600 -p no:cacheprovider"
timeout 600 $PT tests/test_gpu_configs.py -x -k "single_process or world2" 2>&1 | tail -5
timeout 900 python scripts/bench_loader.py > gpurun_out/bench_loader.log 2> gpurun_out/bench_loader.err; cat gpurun_out/bench_loader.log; tail -3 gpurun_out/bench_loader.err
