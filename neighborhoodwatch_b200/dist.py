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
    """All-gather ``[Q, k]`` partials into ``[W, Q, k]`` (the only collective on the path)."""
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


def sharded_knn(local_base: torch.Tensor, queries: torch.Tensor, k: int, metric: str, row_offset: int,
                group=None, local_knn: Optional[Callable] = None, merge: Optional[Callable] = None
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exact global top-k when every rank holds ``local_base`` = its shard of the base rows.

    ``local_knn`` / ``merge`` default to the libnwk kernels; the gloo CPU tests inject array
    stand-ins so the sharding / offset / gather logic is exercised without a GPU.
    Shards smaller than k contribute all their rows, padded with +inf / -1.
    """
    if local_knn is None or merge is None:
        from . import knn as _knn
        local_knn = local_knn or (lambda b, q, kk, m, off: _knn.knn(b, q, kk, metric=m, id_offset=off))
        merge = merge or _knn.merge_topk
    kk = min(k, local_base.shape[0])
    if kk > 0:
        d_loc, i_loc = local_knn(local_base, queries, kk, metric, row_offset)
    else:
        d_loc = torch.empty((queries.shape[0], 0), dtype=torch.float32, device=queries.device)
        i_loc = torch.empty((queries.shape[0], 0), dtype=torch.int64, device=queries.device)
    if kk < k:
        pad_d = torch.full((queries.shape[0], k - kk), float("inf"), dtype=d_loc.dtype, device=d_loc.device)
        pad_i = torch.full((queries.shape[0], k - kk), -1, dtype=i_loc.dtype, device=i_loc.device)
        d_loc = torch.cat([d_loc, pad_d], dim=1)
        i_loc = torch.cat([i_loc, pad_i], dim=1)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return d_loc, i_loc
    all_d, all_i = gather_partials(d_loc, i_loc, group)
    return merge(all_d, all_i, k)


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
