"""Partial-result merge: same entry point and files as neighborhoodwatch/merge.py, on the GPU.

``merge_indices_and_distances(data_dir, k=100)`` (merge.py:39-114) reads
``partial/indices{i}.parquet`` / ``partial/distances{i}.parquet`` for ``i = 0..P-1`` and
writes ``partial/final_indices.parquet`` / ``partial/final_distances.parquet`` with the
schema of partial 0.  The reference folds the partials pairwise on one CPU thread with
``np.argsort`` over ``[Q, 2k]`` (merge.py:73-101); here all P lists of a query are merged in
one pass by ``nwk_merge_topk`` (csrc/nwk_scan.cu), ordered by (distance, index).

Divergences from the reference, all in code the reference itself leaves unreachable or
broken: the ``RowNum`` column is really dropped (merge.py:33-34
discards the result of ``remove_column``); row batches use ``start * batch_size`` (only
matters beyond 10M queries).
"""
from __future__ import annotations

import glob
import math
import re

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import torch

from . import knn as _knn
from .nw_utils import get_partial_distances_filename, get_partial_indices_filename

_GROUP = 16           # partial files merged per kernel call (bounds device memory)
_ROW_BATCH = 10_000_000   # merge.py:46


def get_file_count(data_dir):
    """merge.py:15-27: number of files named ``indices<digits>.parquet`` in ``data_dir``."""
    pattern = re.compile(re.escape(data_dir) + r"/indices(\d+)\.parquet")
    return sum(1 for filename in sorted(glob.glob(f"{data_dir}/indices*.parquet")) if pattern.match(filename))


def read_ifvec_parquet_with_proper_schema(filename):
    """merge.py:30-36, with the ``RowNum`` column actually removed."""
    table = pq.read_table(filename)
    rownum_index = table.schema.get_field_index('RowNum')
    if rownum_index != -1:
        table = table.remove_column(rownum_index)
    return table


def _table_to_numpy(table: pa.Table, dtype) -> np.ndarray:
    out = np.empty((table.num_rows, table.num_columns), dtype=dtype)
    for j in range(table.num_columns):
        out[:, j] = table.column(j).to_numpy()
    return out


def _numpy_to_table(arr: np.ndarray, schema: pa.Schema) -> pa.Table:
    cols = [pa.array(np.ascontiguousarray(arr[:, j]), type=schema.field(j).type) for j in range(arr.shape[1])]
    return pa.Table.from_arrays(cols, schema=schema)


def merge_indices_and_distances(data_dir, k=100):
    file_count = get_file_count(f"{data_dir}/partial")
    if file_count <= 0:
        return                                                    # merge.py:42
    if not torch.cuda.is_available():
        raise RuntimeError("neighborhoodwatch_b200.merge needs a CUDA device; there is no CPU path")
    device = torch.device("cuda", torch.cuda.current_device())

    indices_table = read_ifvec_parquet_with_proper_schema(get_partial_indices_filename(data_dir, 0))
    distances_table = read_ifvec_parquet_with_proper_schema(get_partial_distances_filename(data_dir, 0))
    nq = len(indices_table)
    k_in = indices_table.num_columns
    k_out = min(k, k_in * file_count)
    if file_count == 1:
        idx_schema, dist_schema = indices_table.schema, distances_table.schema
    else:
        # merge.py:101-102 keeps the first k columns of the concatenation
        idx_schema = pa.schema([indices_table.schema.field(j) for j in range(min(k_out, k_in))])
        dist_schema = pa.schema([distances_table.schema.field(j) for j in range(min(k_out, k_in))])
        if k_out > k_in:   # k larger than one partial's width: name the extra columns in sequence
            extra = range(k_in, k_out)
            idx_schema = pa.schema(list(idx_schema) + [pa.field(str(j), indices_table.schema.field(0).type) for j in extra])
            dist_schema = pa.schema(list(dist_schema) + [pa.field(str(j), distances_table.schema.field(0).type) for j in extra])

    batch_size = min(_ROW_BATCH, nq)
    batch_count = math.ceil(nq / batch_size) if nq else 0
    final_indices_writer = pq.ParquetWriter(get_partial_indices_filename(data_dir, -1), idx_schema)
    final_distances_writer = pq.ParquetWriter(get_partial_distances_filename(data_dir, -1), dist_schema)
    try:
        for start in range(batch_count):
            lo = start * batch_size
            n = min(batch_size, nq - lo)
            if file_count == 1:
                final_indices_writer.write_table(indices_table.slice(lo, n))
                final_distances_writer.write_table(distances_table.slice(lo, n))
                continue
            run_d = run_i = None
            for g0 in range(0, file_count, _GROUP):
                ds, ids = [], []
                if run_d is not None:
                    ds.append(run_d)
                    ids.append(run_i)
                width = None
                for i in range(g0, min(g0 + _GROUP, file_count)):
                    it = read_ifvec_parquet_with_proper_schema(get_partial_indices_filename(data_dir, i)).slice(lo, n)
                    dt = read_ifvec_parquet_with_proper_schema(get_partial_distances_filename(data_dir, i)).slice(lo, n)
                    ids.append(torch.from_numpy(_table_to_numpy(it, np.int64)).to(device))
                    ds.append(torch.from_numpy(_table_to_numpy(dt, np.float32)).to(device))
                width = max(t.shape[1] for t in ds)
                # partial lists may differ in width (running result vs files): pad with +inf
                ds = [_pad(t, width, float("inf")) for t in ds]
                ids = [_pad(t, width, -1) for t in ids]
                keep = min(k_out, width * len(ds))
                run_d, run_i = _knn.merge_topk(torch.stack(ds), torch.stack(ids), keep)
            out_d = run_d[:, :k_out].cpu().numpy()
            out_i = run_i[:, :k_out].cpu().numpy()
            # merge.py:104 -- every merged row must be ascending
            assert np.all(np.diff(out_d, axis=1) >= 0)
            final_indices_writer.write_table(_numpy_to_table(out_i, idx_schema))
            final_distances_writer.write_table(_numpy_to_table(out_d, dist_schema))
    finally:
        final_indices_writer.close()
        final_distances_writer.close()


def _pad(t: torch.Tensor, width: int, value) -> torch.Tensor:
    if t.shape[1] == width:
        return t
    pad = torch.full((t.shape[0], width - t.shape[1]), value, dtype=t.dtype, device=t.device)
    return torch.cat([t, pad], dim=1)
