"""Dataset-API (streaming) kNN driver: the entry point of neighborhoodwatch/cu_knn_ds.py.

``compute_knn_ds`` keeps the reference signature (cu_knn_ds.py:136-146) and its output
contract -- it writes ``partial/final_distances.parquet`` / ``partial/final_indices.parquet``
directly (cu_knn_ds.py:232-233), so ``merge_indices_and_distances`` afterwards finds no
numbered partials and returns (merge.py:40-42).

It implements the *intended* semantics of that path: one global
top-k over all base batches.  The reference overwrites the final file with every base batch,
re-uploads all queries per batch (cu_knn_ds.py:206-211) and mis-offsets short record batches
(cu_knn_ds.py:192); here the queries stay resident, base record batches are staged into two
pinned host slabs and copied on a side stream while the previous slab is searched, every
slab's exact top-k is folded into a running device-side state (``nwk_topk_state``), offsets
are the true global row numbers, and the final files are written once.

Multi-GPU (one process per GPU, ``torch.distributed`` initialised by the launcher): rank r streams
only rows ``shard_bounds(base_count, W, r)`` of the base dataset (row groups before its range are
skipped without being read), keeps its own running top-k, and ONE packed all-gather + merge kernel
at the end produce the global result; rank 0 writes the final files.
"""
from __future__ import annotations

import threading
from typing import Iterator, List, Tuple

import numpy as np
import pyarrow as pa
import pyarrow.dataset as ds
import torch

import math

from . import knn as _knn
from .cu_knn import (_device, _world, cleanup, plan_batch_size, stream_cudf_to_parquet,  # noqa: F401
                     tune_memory as _tune_table)
from .nw_utils import (ensure_partial_dir, get_full_filename, get_partial_distances_filename,
                       get_partial_indices_filename)


def get_gpu_memory():
    """cu_knn_ds.py:30-36 (GPU of this process instead of the hard-coded index 0)."""
    return int(torch.cuda.get_device_properties(torch.cuda.current_device()).total_memory)


def load_dataset(data_dir, filename):
    """cu_knn_ds.py:108-110."""
    return ds.dataset(get_full_filename(data_dir, filename), format="parquet")


def dataset_count(dataset):
    """cu_knn_ds.py:118-119."""
    return dataset.count_rows()


def get_dataset_columns(dataset, n):
    """cu_knn_ds.py:122-133, without the UnboundLocalError when neither text column exists.

    The reference keeps a text column only to drop it again with ``select_dtypes``
    (cu_knn_ds.py:195,208); only the embedding columns are returned here.
    """
    return [f"embedding_{i}" for i in range(n)]


def tune_memory(dataset, batch_size, max_memory_threshold, rmm=None, column_names=None):
    """cu_knn_ds.py:57-105 on arithmetic (see cu_knn.tune_memory); never raises."""
    try:
        import math
        num_rows = dataset_count(dataset)
        names = column_names or [n for n in dataset.schema.names if n.startswith("embedding_")]
        bytes_per_row = max(sum(dataset.schema.field(n).type.bit_width // 8 for n in names
                                if pa.types.is_primitive(dataset.schema.field(n).type)), 1)
        total = get_gpu_memory()
        memory_used = min(num_rows, batch_size) * bytes_per_row
        factor = math.ceil((total * max_memory_threshold) / max(memory_used, 1))
        batch_size *= factor
        while True:
            if num_rows < batch_size:
                batch_size = num_rows
                break
            if int(batch_size) * bytes_per_row > max_memory_threshold * total:
                batch_size = int(0.8 * batch_size)
                break
            batch_size *= 1.2
        return int(batch_size)
    except Exception as e:  # noqa: BLE001
        print(f"-- exception {e}, max {int(0.8 * batch_size)}")
        return int(0.8 * batch_size)


def _fragment_rows(piece) -> int:
    """Rows of a (row-group) fragment from the parquet metadata; ``count_rows()`` of a fragment made by
    ``split_by_row_group`` answers for the whole file."""
    try:
        piece.ensure_complete_metadata()
        return int(sum(rg.num_rows for rg in piece.row_groups))
    except Exception:  # noqa: BLE001 -- not a parquet fragment
        return int(piece.count_rows())


def _iter_record_batches(dataset, columns: List[str], start: int, stop: int, batch_rows: int) -> Iterator[pa.RecordBatch]:
    """Record batches covering rows ``[start, stop)`` of the dataset, in order.  Row groups that lie
    wholly before ``start`` are skipped on their metadata (never decoded): a rank of a multi-GPU run
    pays for its own shard only.  A row group is decoded in one call (arrow decodes its columns in
    parallel: one column per dimension gives it hundreds to work on) and the next two row groups are
    decoded ahead on helper threads while the caller packs the current one."""
    from concurrent.futures import ThreadPoolExecutor
    todo = []                                        # (piece, first row of the piece)
    pos = 0
    for frag in dataset.get_fragments():
        try:
            pieces = frag.split_by_row_group()
        except Exception:  # noqa: BLE001 -- not a parquet fragment
            pieces = [frag]
        for piece in pieces:
            if pos >= stop:
                break
            n = _fragment_rows(piece)
            if pos + n > start:
                todo.append((piece, pos))
            pos += n
        if pos >= stop:
            break
    if not todo:
        return
    with ThreadPoolExecutor(max_workers=2) as ahead:
        pending = [ahead.submit(piece.to_table, columns=columns) for piece, _ in todo[:2]]
        for i, (piece, p0) in enumerate(todo):
            table = pending.pop(0).result()
            if i + 2 < len(todo):
                pending.append(ahead.submit(todo[i + 2][0].to_table, columns=columns))
            lo, hi = max(p0, start), min(p0 + table.num_rows, stop)
            if hi <= lo:
                continue
            for batch in table.slice(lo - p0, hi - lo).to_batches(max_chunksize=batch_rows):
                if len(batch):
                    yield batch


def _iter_slabs(dataset, columns: List[str], limit: int, slab_rows: int, start: int = 0,
                min_rows: int = 1) -> Iterator[Tuple[int, int, List[pa.RecordBatch]]]:
    """Yield (global_row_offset, rows, record batches) covering rows ``[start, limit)`` of the dataset
    in slabs of exactly ``slab_rows`` rows; a tail shorter than ``min_rows`` is folded into the slab
    before it (so the last slab holds fewer than ``slab_rows + min_rows`` rows and, unless the whole
    range is shorter, at least ``min_rows``: every knn() call needs k rows, cu_knn_ds.py:204)."""
    total = max(0, limit - start)
    nslab = max(1, -(-total // slab_rows))
    if nslab > 1 and total - (nslab - 1) * slab_rows < min_rows:
        nslab -= 1
    pending: List[pa.RecordBatch] = []
    have = 0
    off = start
    idx = 0
    for batch in _iter_record_batches(dataset, columns, start, limit, min(slab_rows, 131072)):
        while len(batch):
            want = slab_rows if idx < nslab - 1 else total - (nslab - 1) * slab_rows
            take = min(len(batch), want - have)
            pending.append(batch.slice(0, take))
            have += take
            batch = batch.slice(take)
            if have == want:
                yield off, have, pending
                off += have
                idx += 1
                pending, have = [], 0
    if have:
        yield off, have, pending


_pack_pool = None


def _pool():
    """A small thread pool for packing decoded columns into pinned staging (numpy's contiguous copies
    release the GIL; arrow -> numpy is zero-copy for float32 columns without nulls)."""
    global _pack_pool
    if _pack_pool is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 4
        _pack_pool = ThreadPoolExecutor(max_workers=max(2, min(16, cores // 2)))
    return _pack_pool


def _fill_pinned(staging: np.ndarray, batches: List[pa.RecordBatch], columns: List[str]) -> None:
    """Column-per-dimension record batches -> column-major ``[d, rows]`` pinned staging."""
    jobs = []
    off = 0
    for b in batches:
        n = len(b)
        if n:
            idx = [b.schema.get_field_index(name) for name in columns]
            for j0 in range(0, len(columns), 64):
                jobs.append((b, idx, j0, min(j0 + 64, len(columns)), off, n))
        off += n

    def pack(job):
        b, idx, j0, j1, o, n = job
        for j in range(j0, j1):
            staging[j, o:o + n] = b.column(idx[j]).to_numpy(zero_copy_only=False)

    total = sum(job[5] * (job[3] - job[2]) for job in jobs)
    if total >= (1 << 22) and len(jobs) > 1:
        list(_pool().map(pack, jobs))
    else:
        for job in jobs:
            pack(job)


def compute_knn_ds(data_dir,
                   dimensions,
                   query_filename,
                   query_count,
                   base_filename,
                   base_count,
                   mem_tune=False,
                   k=100,
                   initial_batch_size=200000,
                   max_memory_threshold=0.2,
                   split=True):
    """cu_knn_ds.py:136-178."""
    device = _device()
    batch_size = initial_batch_size
    query_dataset = load_dataset(data_dir, query_filename)
    query_column_names = get_dataset_columns(query_dataset, dimensions)
    base_dataset = load_dataset(data_dir, base_filename)
    base_column_names = get_dataset_columns(base_dataset, dimensions)
    for name in base_column_names:
        assert name in base_dataset.schema.names, f"base file lacks column {name}"
    for name in query_column_names:
        assert name in query_dataset.schema.names, f"query file lacks column {name}"
    base_count = min(base_count, dataset_count(base_dataset))
    query_count = min(query_count, dataset_count(query_dataset))

    if mem_tune:
        # planned, not searched (the reference's trial loop is cu_knn_ds.py:57-105): a slab and its
        # double-buffer twin beside the engine's workspace, within max_memory_threshold of the device
        rank, world = _world()
        batch_size = plan_batch_size(math.ceil(base_count / world), dimensions, query_count, k,
                                     max_memory_threshold / 2)
    batch_size = int(min(batch_size, base_count))

    # cu_knn_ds.py:168
    assert (base_count % batch_size == 0) or k <= (base_count % batch_size), \
        f"Cannot generate k of {k} with only {base_count} rows and batch_size of {batch_size}."

    process_dataset_batches(data_dir, base_dataset, base_column_names, query_dataset, query_column_names,
                            batch_size, query_count, k, split, base_count=base_count, device=device)


def process_dataset_batches(data_dir,
                            base_dataset,
                            base_column_names,
                            query_dataset,
                            query_column_names,
                            batch_size,
                            query_count,
                            k,
                            split,
                            base_count=None,
                            device=None):
    """cu_knn_ds.py:181-239 with a running global top-k, a double-buffered loader and rank sharding."""
    from . import dist as _dist
    device = device or _device()
    rank, world = _world()
    if rank == 0:
        ensure_partial_dir(data_dir)
    if base_count is None:
        base_count = dataset_count(base_dataset)
    d = len(base_column_names)
    shard_lo, shard_hi = _dist.shard_bounds(base_count, world, rank)
    shard_rows = shard_hi - shard_lo

    # queries: resident for the whole run
    q_table = query_dataset.head(query_count, columns=query_column_names)
    q_stage = torch.empty((d, len(q_table)), dtype=torch.float32, pin_memory=True)
    _fill_pinned(q_stage.numpy(), q_table.to_batches(), query_column_names)
    queries = _knn.rows_from_columns(q_stage.to(device))
    nq = queries.shape[0]
    assert nq == query_count                                        # cu_knn_ds.py:229-230

    ring_rows = int(min(batch_size + k, max(shard_rows, 1)))        # a folded tail adds fewer than k rows
    copy_stream = torch.cuda.Stream(device=device)
    staging = [torch.empty((d, ring_rows), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    staged_evt = [torch.cuda.Event() for _ in range(2)]
    dev_cm = [torch.empty((d, ring_rows), dtype=torch.float32, device=device) for _ in range(2)]
    dev_rm = torch.empty((ring_rows, d), dtype=torch.float32, device=device)
    state = _knn.RunningTopK(nq, k, device)

    slabs = _iter_slabs(base_dataset, base_column_names, shard_hi, batch_size, start=shard_lo, min_rows=k)
    # a reader thread decodes parquet into the pinned slab while the GPU works
    lock = threading.Semaphore(2)
    ready: list = []
    done = threading.Event()
    cond = threading.Condition()
    error: list = []

    def reader():
        # decodes parquet into the pinned slab AND enqueues its host->device copy, so the copy
        # of slab i+1 runs on the side stream while the main thread is inside knn(slab i)
        try:
            torch.cuda.set_device(device)
            for i, (offset, rows, batches) in enumerate(slabs):
                lock.acquire()                      # slot b is free: knn(slab i-2) has returned
                b = i & 1
                _fill_pinned(staging[b].numpy(), batches, base_column_names)
                with torch.cuda.stream(copy_stream):
                    dev_cm[b][:, :rows].copy_(staging[b][:, :rows], non_blocking=True)
                    staged_evt[b].record(copy_stream)
                with cond:
                    ready.append((i, offset, rows))
                    cond.notify()
        except Exception as e:  # noqa: BLE001
            error.append(e)
        finally:
            done.set()
            with cond:
                cond.notify()

    th = threading.Thread(target=reader, daemon=True)
    th.start()
    main_stream = torch.cuda.current_stream(device)
    while True:
        with cond:
            while not ready and not done.is_set():
                cond.wait(0.05)
            if not ready:
                break
            i, offset, rows = ready.pop(0)
        b = i & 1
        main_stream.wait_event(staged_evt[b])
        dataset = _knn.rows_from_columns(dev_cm[b], rows, out=dev_rm)      # [d, rows] -> [rows, d], our kernel
        if split:
            assert (k <= rows) or rows == shard_rows                # cu_knn_ds.py:204 (a whole shard may be short)
        if rows >= k:
            # knn of this slab folded into the running result; the current k-th distances bound the search
            state.search(dataset, queries, metric="sqeuclidean", id_offset=offset)
        else:
            dd, ii = _knn.knn(dataset, queries, rows, metric="sqeuclidean", id_offset=offset)
            pad_d = torch.full((nq, k), float("inf"), dtype=torch.float32, device=device)
            pad_i = torch.full((nq, k), -1, dtype=torch.int64, device=device)
            pad_d[:, :rows] = dd
            pad_i[:, :rows] = ii
            state.update(pad_d, pad_i)
        lock.release()
    th.join()
    if error:
        raise error[0]
    distances, indices = state.result()
    torch.cuda.current_stream(device).synchronize()
    state.close()
    # every rank holds the exact top-k of its shard with global ids: one exchange, one merge
    distances, indices = _dist.merge_across_ranks(distances, indices, k)
    assert (len(distances) == query_count)                          # cu_knn_ds.py:229
    assert (len(indices) == query_count)                            # cu_knn_ds.py:230
    if rank == 0:
        stream_cudf_to_parquet(get_partial_distances_filename(data_dir, -1), distances, 100000)
        stream_cudf_to_parquet(get_partial_indices_filename(data_dir, -1), indices, 100000)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
