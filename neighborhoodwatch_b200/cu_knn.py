"""Table-API kNN driver: same entry points as neighborhoodwatch/cu_knn.py, served by libnwk.

``compute_knn`` (cu_knn.py:151-180) and ``process_batches`` (cu_knn.py:193-287) keep the
reference's signatures, batch / split arithmetic, assertions and on-disk contract
(``partial/distances{i}.parquet`` + ``partial/indices{i}.parquet``: Q rows, k columns named
"0".."k-1", float32 ascending distances, int64 global row ids), so ``nw`` / ``ck`` and
``merge_indices_and_distances`` work unchanged on top of it.

What is different underneath: parquet columns go host -> device once per table (no
arrow->cudf->dlpack->cupy round trips, cu_knn.py:206-211,231-235), ``knn()`` is the fused
tcgen05 distance + top-k of csrc/nwk_tc.cu, the ``+ batch_offset`` (cu_knn.py:273) is folded
into the kernel, and the per-split results are concatenated on the device and leave it once.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import torch

from . import knn as _knn
from .nw_utils import (ensure_partial_dir, get_full_filename, get_partial_distances_filename,
                       get_partial_indices_filename, output_dimension_validity_check)

_FLOAT_TYPES = (pa.float32(), pa.float64())


# --------------------------------------------------------------------------------------
# parquet <-> device
# --------------------------------------------------------------------------------------
def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("neighborhoodwatch_b200 needs a CUDA device (sm_100a); there is no CPU path")
    # the reference hard-codes GPU 0 (cu_knn.py:64-65); one process per GPU here, so the
    # process's current device plays that role
    return torch.device("cuda", torch.cuda.current_device())


def float_column_names(table: pa.Table) -> List[str]:
    """Columns the reference keeps with ``select_dtypes(['float32', 'float64'])`` (cu_knn.py:208,232)."""
    return [f.name for f in table.schema if f.type in _FLOAT_TYPES]


def table_to_device(table: pa.Table, device: Optional[torch.device] = None) -> torch.Tensor:
    """Float columns of an arrow table -> row-major fp32 ``[rows, d]`` device tensor.

    The parquet layout is one column per dimension (generate_dataset.py:233-235), i.e.
    column-major.  Columns are packed contiguously into pinned host memory, copied in one
    transfer and transposed on the device (the reference does the same transposition with
    ``cupy.copy(order='C')``, cu_knn.py:211).
    """
    device = device or _device()
    names = float_column_names(table)
    rows = table.num_rows
    if not names:
        raise ValueError("table has no float32/float64 columns")
    staging = torch.empty((len(names), rows), dtype=torch.float32, pin_memory=True)
    host = staging.numpy()
    for j, name in enumerate(names):
        col = table.column(name)
        off = 0
        for chunk in col.chunks:
            n = len(chunk)
            if n:
                host[j, off:off + n] = chunk.to_numpy(zero_copy_only=False)
            off += n
    dev_cm = staging.to(device, non_blocking=True)
    out = dev_cm.t().contiguous()
    torch.cuda.current_stream(device).synchronize()
    return out


def stream_cudf_to_parquet(filename: str, df, chunk_size: int) -> None:
    """cu_knn.py:26-51: one parquet file, ``chunk_size``-row row groups.

    ``df`` is a ``[rows, k]`` numpy array or torch tensor (the reference passes a cuDF frame);
    columns are named "0".."k-1" as at cu_knn.py:278-279.
    """
    if isinstance(df, torch.Tensor):
        df = df.detach().cpu().numpy()
    arr = np.ascontiguousarray(df)
    names = [str(i) for i in range(arr.shape[1])]
    writer = None
    num_chunks = max(1, (len(arr) + chunk_size - 1) // chunk_size)
    try:
        for i in range(num_chunks):
            part = arr[i * chunk_size:(i + 1) * chunk_size]
            table = pa.Table.from_arrays([pa.array(np.ascontiguousarray(part[:, j])) for j in range(arr.shape[1])],
                                         names=names)
            if writer is None:
                writer = pq.ParquetWriter(filename, table.schema)
            writer.write_table(table)
    finally:
        if writer is not None:
            writer.close()


# --------------------------------------------------------------------------------------
# reference helpers kept for API compatibility
# --------------------------------------------------------------------------------------
def load_table(data_dir, filename, start, end):
    """cu_knn.py:112-114 (``slice(start, end)`` is arrow's (offset, length))."""
    return pq.read_table(get_full_filename(data_dir, filename)).slice(start, end)


def drop_columns(table, keep_columns):
    """cu_knn.py:117-123: keep only the named columns, in table order."""
    keep = set(keep_columns)
    return table.select([name for name in table.schema.names if name in keep])


def get_embedding_count(table):
    """cu_knn.py:126-129."""
    return sum(1 for name in table.schema.names if name.startswith("embedding_"))


def prep_table(data_dir, model_name, filename, count, dimension):
    """cu_knn.py:132-148: first ``count`` rows, dimension check, only ``embedding_i`` columns."""
    path = get_full_filename(data_dir, filename)
    schema = pq.read_schema(path)
    total_dimensions = sum("embedding_" in c for c in schema.names)
    assert output_dimension_validity_check(model_name, dimension, total_dimensions), \
        f"Expected {dimension} values, got {total_dimensions} for model {model_name} [filename: {filename}]"
    wanted = {f"embedding_{i}" for i in range(total_dimensions)}
    columns = [name for name in schema.names if name in wanted]
    # reading only the embedding columns gives the table the reference has after drop_columns
    return pq.read_table(path, columns=columns).slice(0, count)


def _total_gpu_memory() -> int:
    return int(torch.cuda.get_device_properties(torch.cuda.current_device()).total_memory)


def tune_memory(table, batch_size, max_memory_threshold, rmm=None):
    """cu_knn.py:54-109 without the trial allocations.

    The reference measures ``cudf.DataFrame.memory_usage()`` of trial batches; for a table of
    fixed-width columns that is exactly ``rows * bytes_per_row``, so the same search (factor
    jump with the hard-coded 0.2, then x1.2 growth until the threshold, then x0.8) is run on
    arithmetic.  Never raises (cu_knn.py:104-107).  ``rmm`` is accepted and ignored.
    """
    try:
        num_rows = len(table)
        bytes_per_row = sum(f.type.bit_width // 8 for f in table.schema if pa.types.is_primitive(f.type))
        bytes_per_row = max(bytes_per_row, 1)
        total_gpu_memory = _total_gpu_memory()
        memory_used = min(num_rows, batch_size) * bytes_per_row
        factor = math.ceil((total_gpu_memory * .2) / max(memory_used, 1))
        batch_size *= factor
        while True:
            if num_rows < batch_size:
                batch_size = num_rows
                break
            memory_used = int(batch_size) * bytes_per_row
            if memory_used > max_memory_threshold * total_gpu_memory:
                batch_size = int(0.8 * batch_size)
                break
            batch_size *= 1.2
        return int(batch_size)
    except Exception as e:  # noqa: BLE001 -- mirrors the reference's catch-all
        print(f"-- exception {e}, max batch size {int(0.8 * batch_size)}")
        return int(0.8 * batch_size)


def cleanup(*args):
    """cu_knn.py:183-190 -- nothing to do: no RMM pool, tensors free with their references."""
    return None


# --------------------------------------------------------------------------------------
# the hot loop
# --------------------------------------------------------------------------------------
def compute_knn(data_dir,
                model_name,
                dimensions,
                query_filename,
                query_count,
                base_filename,
                base_count,
                mem_tune=False,
                k=100,
                initial_batch_size=100000,
                max_memory_threshold=0.1,
                split=True):
    """cu_knn.py:151-180."""
    batch_size = initial_batch_size
    query_table = prep_table(data_dir, model_name, query_filename, query_count, dimensions)
    base_table = prep_table(data_dir, model_name, base_filename, base_count, dimensions)

    if mem_tune:
        batch_size = tune_memory(base_table, batch_size, max_memory_threshold)

    batch_count = math.ceil(len(base_table) / batch_size)
    assert (len(base_table) % batch_size == 0) or k <= (len(base_table) % batch_size), \
        f"Cannot generate k of {k} with only {len(base_table)} rows and batch_size of {batch_size}."

    process_batches(data_dir, base_table, query_table, batch_count, batch_size, k, split)


def process_batches(data_dir,
                    base_table,
                    query_table,
                    batch_count,
                    batch_size,
                    k,
                    split,
                    engine='raft'):
    """cu_knn.py:193-287: one partial (distances, indices) file pair per base batch."""
    metric = _knn.ENGINE_METRIC[engine]
    device = _device()
    ensure_partial_dir(data_dir)
    nq = len(query_table)
    queries = table_to_device(query_table, device)      # resident for the whole run

    for start in range(0, batch_count):
        batch_offset = start * batch_size
        batch_length = batch_size if start != batch_count - 1 else len(base_table) - batch_offset
        dataset = table_to_device(base_table.slice(batch_offset, batch_length), device)

        # cu_knn.py:216-223 -- split_factor = 1
        splits = 1 * batch_count
        rows_per_split = nq // splits
        if not split:
            splits = 1

        distances = torch.empty((nq, k), dtype=torch.float32, device=device)
        indices = torch.empty((nq, k), dtype=torch.int64, device=device)
        # The reference runs one knn() per query split and concatenates (cu_knn.py:225-276); the
        # rows of a kNN result do not depend on which other queries share the call, so the splits
        # are served by ONE call per base batch (the base batch is prepared once).  The split
        # arithmetic is kept for its assertions.
        filled = 0
        for i in range(splits):
            offset = i * rows_per_split
            length = rows_per_split if i != splits - 1 else nq - offset
            assert (k <= len(dataset))                                  # cu_knn.py:237
            filled += max(length, 0)
        if nq > 0:
            _knn.knn(dataset, queries, k, metric=metric, id_offset=batch_offset, out=(distances, indices))

        assert (filled == nq)                                           # cu_knn.py:281-282
        stream_cudf_to_parquet(get_partial_distances_filename(data_dir, start), distances, 100000)
        stream_cudf_to_parquet(get_partial_indices_filename(data_dir, start), indices, 100000)
        del dataset, distances, indices
