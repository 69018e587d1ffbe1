"""Multi-GPU exact kNN: base rows sharded over ranks, one process per GPU.

The reference has no multi-device code at all ("Parallelism strategies:
none"); its only sharding is temporal -- base batches processed one after another on GPU 0
(cu_knn.py:201-205) and recombined by the CPU merge (merge.py:56-101).  Here that row
sharding is spatial: rank r owns base rows ``[lo_r, hi_r)``, computes its exact local top-k
with ``id_offset = lo_r`` (the ``+ batch_offset`` of cu_knn.py:273), the ``[Q, k]`` lists are
exchanged with ONE all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests), and every
rank reduces the W lists with the k-way merge (merge.py:73-104 semantics, (distance, index)
order).  The exchange is 12*Q*k bytes per rank (12 MB at Q=10^4, k=100).
"""
from __future__ import annotations

import inspect
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range of ``rank``: the first ``total_rows % world_size`` ranks get one extra row."""
    base, extra = divmod(total_rows, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def gather_partials(local_d: torch.Tensor, local_i: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather ``[Q, k]`` partials into ``[W, Q, k]`` as two collectives (kept for callers that hold
    separate tensors; ``sharded_knn`` uses the packed single-collective exchange below)."""
    world = dist.get_world_size(group)
    local_d = local_d.contiguous()
    local_i = local_i.contiguous()
    nq, k = local_d.shape
    # concatenated along dim 0 (the one output layout every backend accepts), viewed as [W, Q, k]
    all_d = torch.empty((world * nq, k), dtype=local_d.dtype, device=local_d.device)
    all_i = torch.empty((world * nq, k), dtype=local_i.dtype, device=local_i.device)
    dist.all_gather_into_tensor(all_d, local_d, group=group)
    dist.all_gather_into_tensor(all_i, local_i, group=group)
    return all_d.view(world, nq, k), all_i.view(world, nq, k)


# ---- packed exchange: ONE collective moves distances and ids together -------------------------
# chunk of a rank = [Q*k f32 | pad to 16 B | Q*k i64] (knn.packed_chunk_bytes); the local top-k is
# written straight into the send chunk by the kernel (no packing pass), the gathered buffer is merged
# in place by nwk_merge_topk_strided (no unpacking pass).
_xbuf: dict = {}
last_exchange: Optional[str] = None      # which exchange served the last sharded call (diagnostics / tests)


def _chunk_bytes(nq: int, k: int) -> int:
    return (nq * k * 4 + 15) // 16 * 16 + nq * k * 8


def _chunk_views(buf: torch.Tensor, nq: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    off = (nq * k * 4 + 15) // 16 * 16
    return (buf[: nq * k * 4].view(torch.float32).view(nq, k), buf[off: off + nq * k * 8].view(torch.int64).view(nq, k))


def exchange_buffers(nq: int, k: int, world: int, device: torch.device) -> Tuple[torch.Tensor, torch.Tensor]:
    """(send chunk, gather buffer of ``world`` chunks), cached per shape and device."""
    key = (nq, k, world, device.type, device.index)
    bufs = _xbuf.get(key)
    if bufs is None:
        _xbuf.clear()                                  # one shape at a time: keeps HBM for the base
        chunk = _chunk_bytes(nq, k)
        bufs = (torch.empty(chunk, dtype=torch.uint8, device=device),
                torch.empty(world * chunk, dtype=torch.uint8, device=device))
        _xbuf[key] = bufs
    return bufs


def sharded_knn(local_base: torch.Tensor, queries: torch.Tensor, k: int, metric: str, row_offset: int,
                group=None, local_knn: Optional[Callable] = None, merge: Optional[Callable] = None
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exact global top-k when every rank holds ``local_base`` = its shard of the base rows.

    ``local_knn(base, queries, k, metric, offset, out=None)`` / ``merge(all_d, all_i, k)`` default to
    the libnwk kernels; the gloo CPU tests inject array stand-ins so the sharding / offset / packing /
    gather logic is exercised without a GPU.  Shards smaller than k contribute all their rows, padded
    with +inf / -1.
    """
    nq = queries.shape[0]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    native = local_knn is None
    if native:
        from . import knn as _knn

        def local_knn(b, q, kk, m, off, out=None):
            return _knn.knn(b, q, kk, metric=m, id_offset=off, out=out)
    kk = min(k, local_base.shape[0])
    if world == 1:
        if kk == k:
            return local_knn(local_base, queries, k, metric, row_offset)
        sbuf = torch.empty(_chunk_bytes(nq, k), dtype=torch.uint8, device=queries.device)
    else:
        sbuf, gbuf = exchange_buffers(nq, k, world, queries.device)
    d_loc, i_loc = _chunk_views(sbuf, nq, k)
    if kk == k:
        # the kernel writes its [Q, k] lists straight into the send chunk
        if "out" in inspect.signature(local_knn).parameters:
            out = local_knn(local_base, queries, k, metric, row_offset, out=(d_loc, i_loc))
        else:                                          # injected stand-in without an ``out`` parameter
            out = local_knn(local_base, queries, k, metric, row_offset)
        if out[0].data_ptr() != d_loc.data_ptr():
            d_loc.copy_(out[0])
            i_loc.copy_(out[1])
    else:
        d_loc.fill_(float("inf"))
        i_loc.fill_(-1)
        if kk > 0:
            dd, ii = local_knn(local_base, queries, kk, metric, row_offset)
            d_loc[:, :kk] = dd
            i_loc[:, :kk] = ii
    if world == 1:
        return d_loc, i_loc
    return _exchange_and_merge(sbuf, gbuf, nq, k, world, group, merge)


def _exchange_and_merge(sbuf, gbuf, nq, k, world, group, merge):
    global last_exchange
    dist.all_gather_into_tensor(gbuf, sbuf, group=group)
    last_exchange = "packed-all-gather"
    if merge is None:
        from . import knn as _knn
        return _knn.merge_topk_packed(gbuf, world, nq, k, k)
    chunk = _chunk_bytes(nq, k)
    views = [_chunk_views(gbuf[p * chunk:(p + 1) * chunk], nq, k) for p in range(world)]
    return merge(torch.stack([v[0] for v in views]), torch.stack([v[1] for v in views]), k)


def merge_across_ranks(local_d: torch.Tensor, local_i: torch.Tensor, k: int, group=None,
                       merge: Optional[Callable] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Global top-k from every rank's finished ``[Q, k]`` list (ids already global; lists shorter than k
    padded with +inf / -1 by the caller): the reducer the streaming driver runs once at the end."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_d, local_i
    nq = local_d.shape[0]
    assert local_d.shape == (nq, k) and local_i.shape == (nq, k)
    sbuf, gbuf = exchange_buffers(nq, k, world, local_d.device)
    d_loc, i_loc = _chunk_views(sbuf, nq, k)
    d_loc.copy_(local_d)
    i_loc.copy_(local_i)
    return _exchange_and_merge(sbuf, gbuf, nq, k, world, group, merge)


def sharded_maxsim(query_tokens: torch.Tensor, local_doc_tokens: torch.Tensor, k: int, doc_offset: int,
                   doc_len: Optional[torch.Tensor] = None, group=None, local_maxsim: Optional[Callable] = None,
                   merge: Optional[Callable] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Document-level MaxSim when every rank holds a shard of the documents: local top-k with global
    document ids (``id_offset = doc_offset``), one all-gather, merge by (score descending, id) --
    the kNN reducer applied to the negated scores.  Shards smaller than k are padded (-inf / -1)."""
    if local_maxsim is None or merge is None:
        from . import knn as _knn
        local_maxsim = local_maxsim or (lambda q, dd, kk, dl, off: _knn.maxsim_topk(q, dd, kk, doc_len=dl, id_offset=off))
        merge = merge or _knn.merge_topk
    nq = query_tokens.shape[0]
    kk = min(k, local_doc_tokens.shape[0])
    if kk > 0:
        s_loc, i_loc = local_maxsim(query_tokens, local_doc_tokens, kk, doc_len, doc_offset)
    else:
        s_loc = torch.empty((nq, 0), dtype=torch.float32, device=query_tokens.device)
        i_loc = torch.empty((nq, 0), dtype=torch.int64, device=query_tokens.device)
    if kk < k:
        s_loc = torch.cat([s_loc, torch.full((nq, k - kk), float("-inf"), dtype=s_loc.dtype, device=s_loc.device)], dim=1)
        i_loc = torch.cat([i_loc, torch.full((nq, k - kk), -1, dtype=i_loc.dtype, device=i_loc.device)], dim=1)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return s_loc, i_loc
    all_d, all_i = gather_partials(-s_loc, i_loc, group)          # distances = negated scores
    m_d, m_i = merge(all_d, all_i, k)
    return -m_d, m_i
