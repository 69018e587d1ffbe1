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

Multi-GPU (one process per GPU, ``torch.distributed`` initialised by the launcher): the same entry
points are rank-aware.  The reference slices the base temporally -- batch after batch on GPU 0
(cu_knn.py:201-205) -- and here every batch is additionally split spatially: rank r takes rows
``shard_bounds(batch_length, W, r)`` of the batch, computes its exact local top-k with global ids
(``+ batch_offset`` of cu_knn.py:273 plus the shard start), one packed all-gather exchanges the
``[Q, k]`` lists and the merge kernel reduces them (dist.sharded_knn); rank 0 writes the batch's
partial files, which are the files a single-GPU run writes.  In that mode ``compute_knn`` reads only
the row groups a rank needs (``ParquetRows``) instead of the whole table on every rank.
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


def _world():
    """(rank, world size) of the launcher's process group; (0, 1) for a plain single-GPU run."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class ParquetRows:
    """The first ``count`` rows of some columns of a parquet file, read on demand.

    ``len()`` and ``slice(offset, length) -> pa.Table`` are the two operations process_batches needs
    of ``base_table`` (cu_knn.py:201-206); only the row groups that overlap a slice are read, so a rank
    of a multi-GPU run never holds more than its own share of a batch in host memory (the reference
    reads the whole file, cu_knn.py:114: 70 GB of RSS for 4M x 1536 in docs/nvtop.png)."""

    def __init__(self, path: str, columns: List[str], count: int):
        self._pf = pq.ParquetFile(path)
        md = self._pf.metadata
        self._columns = list(columns)
        self._starts = [0]
        for g in range(md.num_row_groups):
            self._starts.append(self._starts[-1] + md.row_group(g).num_rows)
        self._n = int(min(count, md.num_rows))
        full = self._pf.schema_arrow
        self.schema = pa.schema([full.field(c) for c in self._columns])

    def __len__(self) -> int:
        return self._n

    @property
    def num_rows(self) -> int:
        return self._n

    def slice(self, offset: int = 0, length: Optional[int] = None) -> pa.Table:
        offset = max(0, min(int(offset), self._n))
        end = self._n if length is None else min(self._n, offset + int(length))
        if end <= offset:
            return self.schema.empty_table()
        import bisect
        g0 = bisect.bisect_right(self._starts, offset) - 1
        g1 = bisect.bisect_left(self._starts, end)
        table = self._pf.read_row_groups(list(range(g0, g1)), columns=self._columns, use_threads=True)
        return table.slice(offset - self._starts[g0], end - offset)


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

    def pack(j):
        # arrow -> numpy is zero-copy for float32 chunks without nulls; the copy into the pinned row
        # runs with the GIL released, so the columns are packed by a small thread pool
        off = 0
        for chunk in table.column(names[j]).chunks:
            n = len(chunk)
            if n:
                host[j, off:off + n] = chunk.to_numpy(zero_copy_only=False)
            off += n

    if rows * len(names) >= (1 << 22):
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(16, len(names))) as pool:
            list(pool.map(pack, range(len(names))))
    else:
        for j in range(len(names)):
            pack(j)
    dev_cm = staging.to(device, non_blocking=True)
    out = _knn.rows_from_columns(dev_cm)
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


def plan_batch_size(num_rows: int, d: int, nq: int, k: int, max_memory_threshold: float, world: int = 1,
                    metric: str = "sqeuclidean", free_bytes: Optional[int] = None, total_bytes: Optional[int] = None) -> int:
    """Base rows per batch, planned instead of searched (replaces the trial loop of cu_knn.py:54-109).

    Two limits, both per GPU (a batch is row-sharded over ``world`` ranks):
      * the reference's own criterion: the batch's fp32 rows take at most ``max_memory_threshold`` of the
        device memory (cu_knn.py:89: ``memory_used > max_memory_threshold * total_gpu_memory`` ends the search);
      * what the engine really needs beside them: ``nwk_knn_workspace_bytes`` (fp16 operand copy, candidate logs,
        fallback scratch) must fit in 90 % of the memory that is free now.
    The answer is the largest row count meeting both, at least k, at most ``num_rows``; fewer, larger batches
    are always cheaper for this engine (every batch pays a sample pass and a finish stage per query).
    """
    from . import _native
    lib = _native.load()
    if free_bytes is None or total_bytes is None:
        free_now, total_now = torch.cuda.mem_get_info()
        free_bytes = free_now if free_bytes is None else free_bytes
        total_bytes = total_now if total_bytes is None else total_bytes
    metric_id = _native.METRIC_IDS[metric]
    row_bytes = 4 * d

    def need(rows_per_gpu: int) -> int:
        return rows_per_gpu * row_bytes + lib.nwk_knn_workspace_bytes(rows_per_gpu, max(nq, 1), d, k, metric_id, 0)

    cap_rows = int(max_memory_threshold * total_bytes // row_bytes)          # the reference's criterion
    hi = max(k, min(math.ceil(num_rows / world), cap_rows))
    budget = int(0.9 * free_bytes)
    if need(hi) > budget:
        lo = k
        while lo < hi:                                                      # largest rows with need(rows) <= budget
            mid = (lo + hi + 1) // 2
            if need(mid) <= budget:
                lo = mid
            else:
                hi = mid - 1
        hi = lo
    return int(max(k, min(num_rows, hi * world)))


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
    rank, world = _world()
    if world > 1:
        # every rank needs only its shard of each batch: read row groups on demand
        path = get_full_filename(data_dir, base_filename)
        schema = pq.read_schema(path)
        total_dimensions = sum("embedding_" in c for c in schema.names)
        assert output_dimension_validity_check(model_name, dimensions, total_dimensions), \
            f"Expected {dimensions} values, got {total_dimensions} for model {model_name} [filename: {base_filename}]"
        wanted = {f"embedding_{i}" for i in range(total_dimensions)}
        base_table = ParquetRows(path, [n for n in schema.names if n in wanted], base_count)
    else:
        base_table = prep_table(data_dir, model_name, base_filename, base_count, dimensions)

    if mem_tune:
        # the reference searches for a batch size by trial allocation (cu_knn.py:54-109, replayed on
        # arithmetic by tune_memory); the engine's footprint is known in closed form, so the batch is planned
        batch_size = plan_batch_size(len(base_table), dimensions, len(query_table), k, max_memory_threshold,
                                     world=world)

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
    """cu_knn.py:193-287: one partial (distances, indices) file pair per base batch.

    ``base_table``: a ``pa.Table`` or anything with ``len()`` and ``slice(offset, length) -> pa.Table``
    (``ParquetRows``).  With an initialised process group every batch is row-sharded over the ranks."""
    from . import dist as _dist
    metric = _knn.ENGINE_METRIC[engine]
    device = _device()
    rank, world = _world()
    if rank == 0:
        ensure_partial_dir(data_dir)
    nq = len(query_table)
    queries = table_to_device(query_table, device)      # resident for the whole run

    for start in range(0, batch_count):
        batch_offset = start * batch_size
        batch_length = batch_size if start != batch_count - 1 else len(base_table) - batch_offset
        lo, hi = _dist.shard_bounds(batch_length, world, rank)
        dataset = (table_to_device(base_table.slice(batch_offset + lo, hi - lo), device) if hi > lo
                   else torch.empty((0, queries.shape[1]), dtype=torch.float32, device=device))

        # cu_knn.py:216-223 -- split_factor = 1
        splits = 1 * batch_count
        rows_per_split = nq // splits
        if not split:
            splits = 1

        # The reference runs one knn() per query split and concatenates (cu_knn.py:225-276); the
        # rows of a kNN result do not depend on which other queries share the call, so the splits
        # are served by ONE call per base batch (the base batch is prepared once).  The split
        # arithmetic is kept for its assertions.
        filled = 0
        for i in range(splits):
            offset = i * rows_per_split
            length = rows_per_split if i != splits - 1 else nq - offset
            assert (k <= batch_length)                                  # cu_knn.py:237
            filled += max(length, 0)
        if nq > 0:
            # local top-k with global ids (cu_knn.py:273), exchange, merge: identical on every rank
            distances, indices = _dist.sharded_knn(dataset, queries, k, metric, batch_offset + lo)
        else:
            distances = torch.empty((0, k), dtype=torch.float32, device=device)
            indices = torch.empty((0, k), dtype=torch.int64, device=device)

        assert (filled == nq)                                           # cu_knn.py:281-282
        if rank == 0:
            stream_cudf_to_parquet(get_partial_distances_filename(data_dir, start), distances, 100000)
            stream_cudf_to_parquet(get_partial_indices_filename(data_dir, start), indices, 100000)
        del dataset, distances, indices
    if world > 1:
        import torch.distributed as dist
        dist.barrier()                                  # the partial files exist when any rank returns
