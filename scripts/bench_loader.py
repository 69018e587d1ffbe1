"""Parquet -> device loader throughput: the reference's path is
pq.read_table -> cudf -> dlpack -> cupy.copy(order='C') (cu_knn.py:112-148,206-211).  Here: read only the
embedding columns, pack them into pinned column-major staging with a thread pool, one H2D copy, our tiled
transpose kernel.  Prints one JSON line per shape: seconds and GB/s of every stage."""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pyarrow as pa
import pyarrow.dataset as ds
import pyarrow.parquet as pq
import torch

from neighborhoodwatch_b200 import cu_knn, cu_knn_ds, knn, synth

dev = torch.device("cuda", 0)
shapes = [(1_000_000, 768), (300_000, 1536)] if len(sys.argv) < 3 else [(int(sys.argv[1]), int(sys.argv[2]))]
for rows, d in shapes:
    tmp = tempfile.mkdtemp(prefix="nwk_loader_")
    path = os.path.join(tmp, "base.parquet")
    names = [f"embedding_{i}" for i in range(d)]
    schema = pa.schema([pa.field("text", pa.string())] + [pa.field(n, pa.float32()) for n in names])
    chunk = 100_000                                      # the reference's writer flushes 100 000-row groups
    with pq.ParquetWriter(path, schema, compression="snappy") as w:
        for lo in range(0, rows, chunk):
            n = min(chunk, rows - lo)
            cm = synth.device_vectors(n, d, synth.BASE_SEED, row0=lo, device=dev).t().contiguous().cpu().numpy()
            w.write_table(pa.Table.from_arrays([pa.array(["x"] * n)] + [pa.array(cm[j]) for j in range(d)], schema=schema))
    gb = rows * d * 4 / 1e9
    out = {"rows": rows, "d": d, "gb_fp32": gb, "file_gb": os.path.getsize(path) / 1e9, "host_cores": len(os.sched_getaffinity(0))}
    t0 = time.perf_counter()
    table = cu_knn.prep_table(tmp, "model", "base.parquet", rows, d)
    t1 = time.perf_counter()
    x = cu_knn.table_to_device(table, dev)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    x2 = cu_knn.table_to_device(table, dev)              # second call: pinned allocation warm
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    out["read_table_s"] = t1 - t0
    out["read_table_gb_per_s"] = gb / (t1 - t0)
    out["table_to_device_s"] = [t2 - t1, t3 - t2]
    out["table_to_device_gb_per_s"] = gb / (t3 - t2)
    # the device part alone: transpose kernel
    cmaj = x2.t().contiguous()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    y = torch.empty_like(x2)
    knn.rows_from_columns(cmaj, out=y)
    e0.record()
    for _ in range(5):
        knn.rows_from_columns(cmaj, out=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out["transpose_kernel_ms"] = ms
    out["transpose_kernel_gb_per_s_rw"] = 2 * gb / (ms * 1e-3)
    assert torch.equal(y, x2)
    del x, x2, y, cmaj, table
    # row-group-selective read of one rank's quarter
    rows_obj = cu_knn.ParquetRows(path, names, rows)
    t0 = time.perf_counter()
    part = rows_obj.slice(rows // 4, rows // 4)
    out["parquet_rows_quarter_s"] = time.perf_counter() - t0
    assert part.num_rows == rows // 4
    del part
    # streaming decode into pinned slabs (cu_knn_ds reader thread's work), no GPU compute
    dset = ds.dataset(path, format="parquet")
    slab = 250_000
    stage = torch.empty((d, slab + 100), dtype=torch.float32, pin_memory=True)
    t0 = time.perf_counter()
    seen = 0
    for off, n, batches in cu_knn_ds._iter_slabs(dset, names, rows, slab, min_rows=100):
        cu_knn_ds._fill_pinned(stage.numpy(), batches, names)
        seen += n
    dt = time.perf_counter() - t0
    assert seen == rows
    out["stream_decode_s"] = dt
    out["stream_decode_gb_per_s"] = gb / dt
    print(json.dumps(out))
    os.remove(path)
    os.rmdir(tmp)
