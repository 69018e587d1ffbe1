"""Array-level operators: the calls the reference makes into RAPIDS / torch, served by libnwk.

``knn(dataset, queries, k)`` has the signature and result convention of
``pylibraft.neighbors.brute_force.knn`` as the reference uses it
(neighborhoodwatch/cu_knn.py:243-245, cu_knn_ds.py:213): row-major fp32 device arrays in,
``(distances f32[nq,k], indices i64[nq,k])`` out, ascending per row, indices local to
``dataset``.  The ``metric`` keyword selects what the reference selects with ``engine``
(cu_knn.py:242-262).

PyTorch is only the owner of device memory and streams here; every kernel is ours.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _native

ENGINE_METRIC = {"raft": "sqeuclidean", "cuvs": "cosine", "torch": "dot"}   # cu_knn.py:242-262
_ws_cache: dict = {}
last_stats: Optional[dict] = None


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """One grow-only scratch tensor per device (the caller owns scratch, include/nwk.h)."""
    key = (device.type, device.index)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        _ws_cache.pop(key, None)
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def release_workspace() -> None:
    _ws_cache.clear()


def _as_f32_cuda(x: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor on a CUDA device")
    if x.device.type != "cuda":
        raise RuntimeError(f"{name} is on {x.device}; this engine has no CPU path")
    if x.dim() != 2:
        raise ValueError(f"{name} must be 2-D [rows, d]")
    # the reference feeds `.astype(np.float32)` C-order arrays (cu_knn.py:243-244)
    return x.to(torch.float32).contiguous()


def knn(dataset: torch.Tensor, queries: torch.Tensor, k: int, metric: str = "sqeuclidean",
        id_offset: int = 0, force_exact_scan: bool = False, no_fallback: bool = False, profile: bool = False,
        tf32: bool = False, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exact k nearest neighbours; drop-in for brute_force.knn(dataset, queries, k)."""
    global last_stats
    lib = _native.load()
    dataset = _as_f32_cuda(dataset, "dataset")
    queries = _as_f32_cuda(queries, "queries")
    if dataset.shape[1] != queries.shape[1]:
        raise ValueError(f"dimension mismatch: dataset d={dataset.shape[1]}, queries d={queries.shape[1]}")
    if dataset.device != queries.device:
        raise ValueError("dataset and queries must live on the same device")
    # cu_knn.py:237 -- `assert (k <= len(dataset))`
    assert k <= len(dataset), f"k={k} exceeds dataset rows {len(dataset)}"
    metric_id = _native.METRIC_IDS[metric]
    nb, d = dataset.shape
    nq = queries.shape[0]
    dev = dataset.device
    flags = ((_native.FLAG_FORCE_EXACT_SCAN if force_exact_scan else 0) | (_native.FLAG_NO_FALLBACK if no_fallback else 0)
             | (_native.FLAG_PROFILE if profile else 0) | (_native.FLAG_TF32 if tf32 else 0))
    with torch.cuda.device(dev):
        if out is None:
            out_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
            out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        else:
            out_d, out_i = out
            assert out_d.shape == (nq, k) and out_i.shape == (nq, k) and out_d.is_contiguous() and out_i.is_contiguous()
        if nq == 0:
            return out_d, out_i
        ws_bytes = lib.nwk_knn_workspace_bytes(nb, nq, d, k, metric_id, flags)
        ws = _workspace(ws_bytes, dev)
        stats = _native.KnnStats()
        rc = lib.nwk_knn_f32(dataset.data_ptr(), nb, queries.data_ptr(), nq, d, k, metric_id, int(id_offset),
                             out_d.data_ptr(), out_i.data_ptr(), ws.data_ptr(), ws.numel(), flags,
                             ctypes.byref(stats), _stream_ptr(dev))
        _native.check(rc)
        last_stats = stats.as_dict()
    return out_d, out_i


def rows_from_columns(cols: torch.Tensor, rows: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``[d, ld]`` column-per-dimension staging (the parquet layout) -> row-major ``[rows, d]`` by our
    tiled transpose (``nwk_rows_from_columns_f32``; the reference: ``cupy.copy(order='C')``, cu_knn.py:211).
    ``out``: optional ``[>= rows, d]`` buffer whose first ``rows`` rows are returned."""
    lib = _native.load()
    if cols.device.type != "cuda":
        raise RuntimeError("rows_from_columns needs a CUDA tensor; this engine has no CPU path")
    assert cols.dim() == 2 and cols.dtype == torch.float32 and cols.stride(1) == 1
    d, ld = cols.shape[0], cols.stride(0)
    assert d == 1 or ld >= cols.shape[1]
    rows = cols.shape[1] if rows is None else int(rows)
    assert rows <= cols.shape[1]
    dev = cols.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((rows, d), dtype=torch.float32, device=dev)
        assert out.is_contiguous() and out.dtype == torch.float32 and out.shape[1] == d and out.shape[0] >= rows
        if rows:
            _native.check(lib.nwk_rows_from_columns_f32(cols.data_ptr(), max(ld, rows), d, rows, out.data_ptr(), _stream_ptr(dev)))
    return out[:rows]


def merge_topk(dist: torch.Tensor, idx: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge ``[P, nq, k_in]`` sorted partial lists into ``[nq, k]`` by (distance, index).

    Device version of the argsort fold at neighborhoodwatch/merge.py:73-104.
    """
    lib = _native.load()
    if dist.device.type != "cuda":
        raise RuntimeError("merge_topk needs CUDA tensors; this engine has no CPU path")
    assert dist.dim() == 3 and idx.shape == dist.shape
    dist = dist.to(torch.float32).contiguous()
    idx = idx.to(torch.int64).contiguous()
    P, nq, k_in = dist.shape
    if k > _native.MERGE_MAX_K:
        # the reference merge (merge.py:73-104) takes any k; the device merge keeps k records of every
        # 4096-record segment per round and needs k <= 2048 to converge
        raise ValueError(f"merge_topk supports k <= {_native.MERGE_MAX_K}, got {k}")
    assert k <= P * k_in, f"k={k} exceeds the {P * k_in} candidates per query"
    dev = dist.device
    with torch.cuda.device(dev):
        out_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        if nq == 0:
            return out_d, out_i
        ws_bytes = lib.nwk_merge_workspace_bytes(P, nq, k_in, k)
        ws = _workspace(ws_bytes, dev)
        rc = lib.nwk_merge_topk(P, dist.data_ptr(), idx.data_ptr(), nq, k_in, k, out_d.data_ptr(),
                                out_i.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        _native.check(rc)
    return out_d, out_i


def packed_chunk_bytes(nq: int, k: int) -> int:
    """Bytes of one rank's chunk of a packed exchange buffer: ``[nq*k f32 | pad to 16 B | nq*k i64]``."""
    return (nq * k * 4 + 15) // 16 * 16 + nq * k * 8


def packed_views(buf: torch.Tensor, nq: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(distances f32 [nq,k], indices i64 [nq,k]) views into one chunk (a uint8 tensor)."""
    off = (nq * k * 4 + 15) // 16 * 16
    return (buf[: nq * k * 4].view(torch.float32).view(nq, k), buf[off: off + nq * k * 8].view(torch.int64).view(nq, k))


def merge_topk_packed(buf: torch.Tensor, parts: int, nq: int, k_in: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge the ``parts`` chunks of a packed exchange buffer (``nwk_merge_topk_strided``): what one
    all-gather of ``packed_chunk_bytes`` per rank delivers, merged without unpacking."""
    lib = _native.load()
    assert buf.device.type == "cuda" and buf.dtype == torch.uint8 and buf.is_contiguous()
    chunk = packed_chunk_bytes(nq, k_in)
    assert buf.numel() == parts * chunk
    if k > _native.MERGE_MAX_K:
        raise ValueError(f"merge_topk supports k <= {_native.MERGE_MAX_K}, got {k}")
    dev = buf.device
    off = (nq * k_in * 4 + 15) // 16 * 16
    with torch.cuda.device(dev):
        out_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        ws = _workspace(lib.nwk_merge_workspace_bytes(parts, nq, k_in, k), dev)
        rc = lib.nwk_merge_topk_strided(parts, buf.data_ptr(), chunk // 4, buf.data_ptr() + off, chunk // 8, nq, k_in, k,
                                        out_d.data_ptr(), out_i.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        _native.check(rc)
    return out_d, out_i


class RunningTopK:
    """Running per-query top-k across base slabs (nwk_topk_state_*), see include/nwk.h."""

    def __init__(self, nq: int, k: int, device: torch.device):
        self._lib = _native.load()
        self.nq, self.k, self.device = nq, k, torch.device(device)
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _native.check(self._lib.nwk_topk_state_create(nq, k, ctypes.byref(self._handle)))

    def update(self, dist: torch.Tensor, idx: torch.Tensor) -> None:
        assert dist.shape == (self.nq, self.k) and idx.shape == (self.nq, self.k)
        dist = dist.to(torch.float32).contiguous()
        idx = idx.to(torch.int64).contiguous()
        with torch.cuda.device(self.device):
            _native.check(self._lib.nwk_topk_state_update(self._handle, dist.data_ptr(), idx.data_ptr(),
                                                          _stream_ptr(self.device)))
            # keep the inputs alive until the enqueued copies have run
            torch.cuda.current_stream(self.device).synchronize()

    def search(self, dataset: torch.Tensor, queries: torch.Tensor, metric: str = "sqeuclidean", id_offset: int = 0,
               profile: bool = False) -> None:
        """knn(dataset slab) folded into the state in one call (nwk_topk_state_search): the current
        k-th distances bound the search, so slabs after the first cost a fraction of the finish stage."""
        global last_stats
        dataset = _as_f32_cuda(dataset, "dataset")
        queries = _as_f32_cuda(queries, "queries")
        assert queries.shape[0] == self.nq and dataset.shape[1] == queries.shape[1]
        assert self.k <= len(dataset), f"k={self.k} exceeds dataset rows {len(dataset)}"      # cu_knn_ds.py:204
        metric_id = _native.METRIC_IDS[metric]
        nb, d = dataset.shape
        flags = _native.FLAG_PROFILE if profile else 0
        with torch.cuda.device(self.device):
            ws_bytes = self._lib.nwk_knn_workspace_bytes(nb, self.nq, d, self.k, metric_id, flags)
            ws = _workspace(ws_bytes, self.device)
            stats = _native.KnnStats()
            _native.check(self._lib.nwk_topk_state_search(self._handle, dataset.data_ptr(), nb, queries.data_ptr(), d,
                                                          metric_id, int(id_offset), ws.data_ptr(), ws.numel(), flags,
                                                          ctypes.byref(stats), _stream_ptr(self.device)))
            last_stats = stats.as_dict()

    def result(self) -> Tuple[torch.Tensor, torch.Tensor]:
        with torch.cuda.device(self.device):
            out_d = torch.empty((self.nq, self.k), dtype=torch.float32, device=self.device)
            out_i = torch.empty((self.nq, self.k), dtype=torch.int64, device=self.device)
            _native.check(self._lib.nwk_topk_state_finalize(self._handle, out_d.data_ptr(), out_i.data_ptr(),
                                                            _stream_ptr(self.device)))
        return out_d, out_i

    def reset(self) -> None:
        """Empty the state (enqueued on the current stream) so the handle serves another run."""
        with torch.cuda.device(self.device):
            _native.check(self._lib.nwk_topk_state_reset(self._handle, _stream_ptr(self.device)))

    def close(self) -> None:
        if self._handle:
            self._lib.nwk_topk_state_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiGpu:
    """Single-process multi-GPU exact kNN (``nwk_mg_*``): base shards resident on several devices of this
    process, queries from host memory, the global ``[nq, k]`` result back in host memory."""

    def __init__(self, devices):
        self._lib = _native.load()
        self.devices = [int(x) for x in devices]
        arr = (ctypes.c_int32 * len(self.devices))(*self.devices)
        self._h = ctypes.c_void_p()
        _native.check(self._lib.nwk_mg_create(arr, len(self.devices), ctypes.byref(self._h)))

    def knn(self, base_shards, queries, k: int, metric: str = "sqeuclidean", id_offset: int = 0):
        global last_stats
        import numpy as np
        assert len(base_shards) == len(self.devices)
        shards = []
        for dev, b in zip(self.devices, base_shards):
            assert b.device.type == "cuda" and b.device.index == dev and b.dtype == torch.float32 and b.is_contiguous()
            shards.append(b)
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq, d = q.shape
        assert all(b.shape[1] == d for b in shards)
        assert all(k <= b.shape[0] for b in shards), "every shard must hold at least k rows"      # cu_knn.py:237 per shard
        ptrs = (ctypes.c_void_p * len(shards))(*[b.data_ptr() for b in shards])
        rows = (ctypes.c_int64 * len(shards))(*[b.shape[0] for b in shards])
        out_d = np.empty((nq, k), dtype=np.float32)
        out_i = np.empty((nq, k), dtype=np.int64)
        stats = _native.KnnStats()
        for b in shards:
            torch.cuda.current_stream(b.device).synchronize()      # the shards are complete before the library's streams read them
        _native.check(self._lib.nwk_mg_knn_f32(self._h, ptrs, rows, q.ctypes.data, nq, d, k, _native.METRIC_IDS[metric],
                                               int(id_offset), out_d.ctypes.data, out_i.ctypes.data, 0, ctypes.byref(stats)))
        last_stats = stats.as_dict()
        return out_d, out_i

    def close(self):
        if self._h:
            self._lib.nwk_mg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class PinnedArray:
    """``[rows, d]`` fp32 pinned host memory from ``nwk_host_alloc`` (optionally write-combined), exposed as a
    numpy array (``.array``); freed with the object."""

    def __init__(self, shape, write_combined: bool = False, dtype="float32"):
        import numpy as np
        self._lib = _native.load()
        self.shape = tuple(int(x) for x in shape)
        nbytes = int(np.prod(self.shape)) * np.dtype(dtype).itemsize
        self._ptr = ctypes.c_void_p()
        _native.check(self._lib.nwk_host_alloc(nbytes, 1 if write_combined else 0, ctypes.byref(self._ptr)))
        buf = (ctypes.c_char * nbytes).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(self.shape)

    def close(self):
        if self._ptr:
            self.array = None
            self._lib.nwk_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def knn_host(base, queries, k: int, metric: str = "sqeuclidean", id_offset: int = 0, slab_rows: int = 0,
             out=None, force_exact_scan: bool = False):
    """Host-buffer entry point (nwk_knn_host_f32): numpy / pinned-torch arrays in, numpy out.

    The base is streamed through a two-deep device ring on a side stream while the previous
    slab is searched -- the loader design for bases that exceed HBM (cu_knn_ds.py:181-239).
    """
    global last_stats
    import numpy as np
    lib = _native.load()
    if not torch.cuda.is_available():
        raise RuntimeError("knn_host needs a CUDA device; this engine has no CPU path")

    def host_ptr(a):
        if isinstance(a, torch.Tensor):
            assert a.device.type == "cpu" and a.dtype == torch.float32 and a.is_contiguous()
            return a.data_ptr(), tuple(a.shape)
        a = np.ascontiguousarray(a, dtype=np.float32)
        return a.ctypes.data, a.shape, a

    bp = host_ptr(base)
    qp = host_ptr(queries)
    (nb, d), (nq, d2) = bp[1], qp[1]
    assert d == d2
    assert k <= nb, f"k={k} exceeds dataset rows {nb}"
    if out is None:
        out_d = np.empty((nq, k), dtype=np.float32)
        out_i = np.empty((nq, k), dtype=np.int64)
    else:
        out_d, out_i = out
        assert out_d.shape == (nq, k) and out_d.dtype == np.float32 and out_d.flags.c_contiguous
        assert out_i.shape == (nq, k) and out_i.dtype == np.int64 and out_i.flags.c_contiguous
    stats = _native.KnnStats()
    rc = lib.nwk_knn_host_f32(bp[0], nb, qp[0], nq, d, k, _native.METRIC_IDS[metric], int(id_offset),
                              int(slab_rows), out_d.ctypes.data, out_i.ctypes.data,
                              _native.FLAG_FORCE_EXACT_SCAN if force_exact_scan else 0, ctypes.byref(stats))
    _native.check(rc)
    last_stats = stats.as_dict()
    return out_d, out_i


def debug_tc_tile(dataset: torch.Tensor, queries: torch.Tensor, b0: int, q0: int, f16: bool = True) -> torch.Tensor:
    """Raw tensor-core dot products of one 128x256 tile pair (descriptor validation)."""
    lib = _native.load()
    dataset = _as_f32_cuda(dataset, "dataset")
    queries = _as_f32_cuda(queries, "queries")
    dev = dataset.device
    with torch.cuda.device(dev):
        out = torch.zeros((128, 256), dtype=torch.float32, device=dev)
        rc = lib.nwk_debug_tc_tile(dataset.data_ptr(), dataset.shape[0], queries.data_ptr(), queries.shape[0],
                                   dataset.shape[1], b0, q0, out.data_ptr(), 1 if f16 else 0, _stream_ptr(dev))
        _native.check(rc)
    return out


def maxsim_topk(query_tokens: torch.Tensor, doc_tokens: torch.Tensor, k: int, doc_len: Optional[torch.Tensor] = None,
                id_offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Document-level ColBERT MaxSim (nwk_maxsim_f32): ``score(Q, D) = sum_t max_j q_t . d_j``.

    ``query_tokens`` [nq, Tq, d] (Tq a power of two <= 128; pad with zero tokens), ``doc_tokens``
    [nd, Td, d] (Td <= 256), ``doc_len`` optional int32 [nd] valid tokens per document.
    Returns (scores f32 [nq, k] descending, document ids i64 [nq, k] + id_offset).
    """
    global last_stats
    lib = _native.load()
    for name, x in (("query_tokens", query_tokens), ("doc_tokens", doc_tokens)):
        if not isinstance(x, torch.Tensor) or x.device.type != "cuda":
            raise RuntimeError(f"{name} must be a CUDA tensor; this engine has no CPU path")
        if x.dim() != 3:
            raise ValueError(f"{name} must be 3-D [n, tokens, d]")
    q = query_tokens.to(torch.float32).contiguous()
    dd = doc_tokens.to(torch.float32).contiguous()
    nq, tq, d = q.shape
    nd, td, d2 = dd.shape
    if d != d2:
        raise ValueError(f"dimension mismatch: queries d={d}, documents d={d2}")
    assert k <= nd, f"k={k} exceeds the {nd} documents"
    dev = q.device
    dl_ptr = None
    if doc_len is not None:
        doc_len = doc_len.to(device=dev, dtype=torch.int32).contiguous()
        assert doc_len.shape == (nd,)
        dl_ptr = doc_len.data_ptr()
    with torch.cuda.device(dev):
        out_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
        ws_bytes = lib.nwk_maxsim_workspace_bytes(nq, tq, nd, td, d, k)
        ws = _workspace(ws_bytes, dev)
        stats = _native.KnnStats()
        rc = lib.nwk_maxsim_f32(q.data_ptr(), nq, tq, dd.data_ptr(), nd, td, dl_ptr, d, k, int(id_offset),
                                out_s.data_ptr(), out_i.data_ptr(), ws.data_ptr(), ws.numel(), ctypes.byref(stats),
                                _stream_ptr(dev))
        _native.check(rc)
        last_stats = stats.as_dict()
    return out_s, out_i
