#!/bin/bash
# 1 GPU: the rewritten transpose and MaxSim re-scoring kernels -- their tests, the MaxSim bench at both shapes,
# the transpose kernel's rate
mkdir -p gpurun_out
PT="python -m pytest -q -m gpu --timeout 600 -p no:cacheprovider -x"
timeout 600 $PT tests/test_gpu_drivers.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -k "rows_from_columns or maxsim or compute_knn" 2>&1 | tail -4
timeout 300 python scripts/bench_maxsim.py 1000 32 12500 180 128 100 10 > gpurun_out/maxsim_shard.json 2> gpurun_out/maxsim_shard.err; cat gpurun_out/maxsim_shard.json; tail -2 gpurun_out/maxsim_shard.err
timeout 300 python scripts/bench_maxsim.py 1000 32 100000 180 128 100 5 > gpurun_out/maxsim_full.json 2> gpurun_out/maxsim_full.err; cat gpurun_out/maxsim_full.json; tail -2 gpurun_out/maxsim_full.err
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/transpose.json
import json, torch
from neighborhoodwatch_b200 import knn
dev = torch.device("cuda", 0)
out = []
for rows, d in ((1_000_000, 768), (300_000, 1536), (1_000_000, 1024), (2_250_000, 128)):
    cm = torch.randn((d, rows), device=dev)
    y = torch.empty((rows, d), device=dev)
    knn.rows_from_columns(cm, out=y)
    assert torch.equal(y, cm.t())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        knn.rows_from_columns(cm, out=y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    e0.record()
    for _ in range(10):
        y.copy_(cm.view(rows, d))          # plain device copy of the same bytes: the HBM read + write ceiling
    e1.record(); torch.cuda.synchronize()
    ms_copy = e0.elapsed_time(e1) / 10
    out.append({"rows": rows, "d": d, "transpose_ms": ms, "gb_per_s_rw": 2 * rows * d * 4 / ms / 1e6,
                "device_copy_ms": ms_copy, "copy_gb_per_s_rw": 2 * rows * d * 4 / ms_copy / 1e6})
    del cm, y
print(json.dumps(out))
PY
