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


class _DevArray:
    """A device buffer owned by libnwk, exposed to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerExchange:
    """Gather + merge as one kernel over peer memory (``nwk_xchg_*``, csrc/nwk_xchg.cu).

    Every rank's ``[Q, k]`` result buffer lives in memory the library exports with CUDA IPC; the merge
    kernel of each rank reads the W lists of its queries straight from the owners' HBM over NVLink.
    The handles are swapped once, at construction, with one torch.distributed all-gather; the step
    path has no collective call.  ``slot()`` returns tensors over the buffer the NEXT local top-k
    must be written to; ``merge()`` publishes it, waits (on the device) for the peers and merges."""

    def __init__(self, nq: int, k: int, device: torch.device, group=None):
        from . import _native
        import ctypes
        self._lib = _native.load()
        self._native = _native
        self.nq, self.k, self.device = nq, k, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._h = ctypes.c_void_p()
        self._views: dict = {}
        with torch.cuda.device(device):
            _native.check(self._lib.nwk_xchg_create(self.rank, self.world, nq, k, ctypes.byref(self._h)))
            hb = int(self._lib.nwk_xchg_handle_bytes())
            mine = (ctypes.c_ubyte * hb)()
            _native.check(self._lib.nwk_xchg_local_handle(self._h, mine))
            send = torch.tensor(list(mine), dtype=torch.uint8, device=device)
            recv = torch.empty(self.world * hb, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(recv, send, group=group)
            allh = bytes(recv.cpu().numpy().tobytes())
            rc = self._lib.nwk_xchg_open_peers(self._h, allh)
            # every rank must agree on whether the peer path is usable
            ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                msg = self._lib.nwk_last_error().decode("utf-8", "replace") if rc != 0 else "a peer could not map the buffers"
                self.close()
                raise RuntimeError(f"peer exchange unavailable: {msg}")

    def slot(self) -> Tuple[torch.Tensor, torch.Tensor]:
        import ctypes
        pd, pi = ctypes.c_void_p(), ctypes.c_void_p()
        self._native.check(self._lib.nwk_xchg_next_slot(self._h, ctypes.byref(pd), ctypes.byref(pi)))
        views = self._views.get(pd.value)
        if views is None:                              # two slots: two pairs of views, made once
            views = (torch.as_tensor(_DevArray(pd.value, (self.nq, self.k), "<f4"), device=self.device),
                     torch.as_tensor(_DevArray(pi.value, (self.nq, self.k), "<i8"), device=self.device))
            self._views[pd.value] = views
        return views

    def merge(self, k_out: int) -> Tuple[torch.Tensor, torch.Tensor]:
        with torch.cuda.device(self.device):
            out_d = torch.empty((self.nq, k_out), dtype=torch.float32, device=self.device)
            out_i = torch.empty((self.nq, k_out), dtype=torch.int64, device=self.device)
            self._native.check(self._lib.nwk_xchg_merge(self._h, k_out, out_d.data_ptr(), out_i.data_ptr(),
                                                        torch.cuda.current_stream(self.device).cuda_stream))
        return out_d, out_i

    def check(self) -> None:
        self._native.check(self._lib.nwk_xchg_check(self._h))

    def close(self) -> None:
        if self._h:
            self._views.clear()
            self._lib.nwk_xchg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


_peer: dict = {}
_peer_failed: Optional[str] = None


def peer_exchange(nq: int, k: int, device: torch.device, group=None) -> Optional[PeerExchange]:
    """The peer-memory exchange for this shape, or None when it cannot be set up here (no CUDA IPC between
    the ranks, more than 16 ranks, world * k > 2048): the packed all-gather serves the step then.
    ``NWK_EXCHANGE=allgather`` forces the collective."""
    global _peer_failed
    import os
    if os.environ.get("NWK_EXCHANGE", "") == "allgather" or _peer_failed is not None:
        return None
    if device.type != "cuda" or dist.get_backend(group) != "nccl":
        return None
    world = dist.get_world_size(group)
    if world > 16 or world * k > 2048:
        return None
    key = (nq, k, world, device.index)
    x = _peer.get(key)
    if x is None:
        for old in list(_peer.values()):
            old.close()
        _peer.clear()
        try:
            x = PeerExchange(nq, k, device, group)
        except Exception as e:  # noqa: BLE001 -- fall back to the collective, remember why
            _peer_failed = repr(e)[:300]
            return None
        _peer[key] = x
    return x


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
    global last_exchange
    kk = min(k, local_base.shape[0])
    peer = None
    if world == 1:
        if kk == k:
            return local_knn(local_base, queries, k, metric, row_offset)
        sbuf = torch.empty(_chunk_bytes(nq, k), dtype=torch.uint8, device=queries.device)
        d_loc, i_loc = _chunk_views(sbuf, nq, k)
    else:
        peer = peer_exchange(nq, k, queries.device, group) if merge is None else None
        if peer is not None:
            d_loc, i_loc = peer.slot()                 # the kernel writes into the exported buffer
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
    if peer is not None:
        last_exchange = "peer-memory gather+merge kernel"
        return peer.merge(k)
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
