"""CPU oracle for the exact brute-force kNN hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``neighborhoodwatch_b200``) never imports anything under ``oracle/``.

What it restates (reference = /root/reference, phact/neighborhoodwatch):

* ``knn_f64``      -- the array-level operator ``knn(dataset, queries, k)`` the reference
  calls at ``neighborhoodwatch/cu_knn.py:243-245`` (pylibraft brute_force.knn, default
  metric squared-L2), ``cu_knn.py:247-248`` (cuVS brute force, cosine distance) and
  ``cu_knn.py:256-257`` (``1 - q @ b.T`` then ``topk(largest=False)``), evaluated in
  float64 with a deterministic (distance, index) tie order.  This is the "truth" the
  CUDA path is compared with.
* ``knn_f32_reference_form`` -- the same three engines in the arithmetic the reference
  libraries use (fp32; expanded ``|q|^2 + |b|^2 - 2 q.b`` for squared L2 as libraft does,
  ``1 - q.b/(|q||b|)`` for cosine, ``1 - q.b`` for the torch engine) with
  ``argpartition`` + sort of the k survivors.  This is the CPU baseline that is timed.
* ``process_batches_oracle`` -- the batch / split / offset bookkeeping of
  ``cu_knn.py:193-287`` (base batches, query splits, ``+ batch_offset`` at ``:273``)
  on plain arrays, returning the list of per-batch partial results.

Pinning status: the kNN libraries (pylibraft 24.8.1, cuvs 24.8.0) are third-party
and absent from /root/reference, and the reference's tests hold no golden vectors for
this path.  The torch engine and the batch bookkeeping ARE
pinned: ``tests/golden/make_golden.py`` executes the reference's own
``process_batches(engine='torch')`` and ``merge_indices_and_distances`` (unmodified
source, GPU dataframe libraries replaced by numpy shims) and the resulting fixtures
are checked against this oracle in ``tests/test_oracle.py``.  The squared-L2 and
cosine definitions are pinned only through the relation the reference's own validator
states (``parquet_to_format.py:373``: ``distance / 2 == 1 - dot`` on unit vectors);
beyond that they are "parity unpinned".
"""
from __future__ import annotations

import math

import numpy as np

METRICS = ("sqeuclidean", "cosine", "dot")
# engine name used by the reference (cu_knn.py:242-262) -> metric it selects
ENGINE_METRIC = {"raft": "sqeuclidean", "cuvs": "cosine", "torch": "dot"}


def _check(dataset: np.ndarray, queries: np.ndarray, k: int, metric: str) -> None:
    if metric not in METRICS:
        raise ValueError(f"unknown metric {metric!r}")
    if dataset.ndim != 2 or queries.ndim != 2 or dataset.shape[1] != queries.shape[1]:
        raise ValueError("dataset and queries must be [n, d] with equal d")
    # cu_knn.py:237 -- the reference asserts k <= len(dataset) before every knn() call
    assert k <= len(dataset), f"k={k} exceeds dataset rows {len(dataset)}"


def distances_f64(dataset: np.ndarray, queries: np.ndarray, metric: str) -> np.ndarray:
    """Full [nq, nb] float64 distance matrix (small inputs only)."""
    b = np.asarray(dataset, dtype=np.float64)
    q = np.asarray(queries, dtype=np.float64)
    if metric == "sqeuclidean":
        # direct form: sum (q - b)^2, no cancellation
        out = np.empty((q.shape[0], b.shape[0]), dtype=np.float64)
        for i in range(q.shape[0]):
            diff = b - q[i]
            out[i] = np.einsum("ij,ij->i", diff, diff)
        return out
    dots = q @ b.T
    if metric == "dot":
        return 1.0 - dots
    qn = np.sqrt(np.einsum("ij,ij->i", q, q))
    bn = np.sqrt(np.einsum("ij,ij->i", b, b))
    return 1.0 - dots / (qn[:, None] * bn[None, :])


def _block_distances_f64(b: np.ndarray, q: np.ndarray, metric: str,
                         qsq: np.ndarray, qn: np.ndarray) -> np.ndarray:
    """Distances of all queries against one float64 base block, GEMM form.

    In float64 the expanded form loses ~1e-16 absolute, far below fp32 resolution of
    the values being compared, so it is safe as truth for fp32 inputs.
    """
    dots = q @ b.T
    if metric == "dot":
        return 1.0 - dots
    bsq = np.einsum("ij,ij->i", b, b)
    if metric == "sqeuclidean":
        dist = qsq[:, None] + bsq[None, :] - 2.0 * dots
        np.maximum(dist, 0.0, out=dist)
        return dist
    return 1.0 - dots / (qn[:, None] * np.sqrt(bsq)[None, :])


def _topk_sorted(dist: np.ndarray, ids: np.ndarray, k: int):
    """Rows of (dist, ids) -> k smallest, ordered by (distance, index)."""
    n = dist.shape[1]
    if n > k:
        part = np.argpartition(dist, k - 1, axis=1)[:, :k]
        # argpartition is not tie-stable at the k-th value: pull in every entry tied
        # with the k-th distance so the (distance, index) order decides
        kth = np.take_along_axis(dist, part, axis=1).max(axis=1)
        out_d = np.empty((dist.shape[0], k), dtype=dist.dtype)
        out_i = np.empty((dist.shape[0], k), dtype=np.int64)
        for r in range(dist.shape[0]):
            cand = np.nonzero(dist[r] <= kth[r])[0]
            order = np.lexsort((ids[cand] if ids.ndim == 1 else ids[r, cand], dist[r, cand]))[:k]
            sel = cand[order]
            out_d[r] = dist[r, sel]
            out_i[r] = ids[sel] if ids.ndim == 1 else ids[r, sel]
        return out_d, out_i
    out_d = np.empty_like(dist)
    out_i = np.empty(dist.shape, dtype=np.int64)
    for r in range(dist.shape[0]):
        rid = ids if ids.ndim == 1 else ids[r]
        order = np.lexsort((rid, dist[r]))
        out_d[r] = dist[r, order]
        out_i[r] = rid[order]
    return out_d, out_i


def knn_f64(dataset: np.ndarray, queries: np.ndarray, k: int, metric: str = "sqeuclidean",
            block: int = 65536, return_f64: bool = False):
    """Exact top-k in float64; ties broken by ascending index.

    Returns (distances float32 [nq,k], indices int64 [nq,k]) -- the operator contract of
    cu_knn.py:243-245 -- rows ascending by distance.
    """
    _check(dataset, queries, k, metric)
    q = np.asarray(queries, dtype=np.float64)
    qsq = np.einsum("ij,ij->i", q, q)
    qn = np.sqrt(qsq)
    nq, nb = q.shape[0], dataset.shape[0]
    best_d = np.empty((nq, 0), dtype=np.float64)
    best_i = np.empty((nq, 0), dtype=np.int64)
    for start in range(0, nb, block):
        b = np.asarray(dataset[start:start + block], dtype=np.float64)
        dist = _block_distances_f64(b, q, metric, qsq, qn)
        ids = np.arange(start, start + b.shape[0], dtype=np.int64)
        d_blk, i_blk = _topk_sorted(dist, ids, min(k, b.shape[0]))
        cat_d = np.concatenate([best_d, d_blk], axis=1)
        cat_i = np.concatenate([best_i, i_blk], axis=1)
        best_d, best_i = _topk_sorted(cat_d, cat_i, min(k, cat_d.shape[1]))
    if return_f64:
        return best_d, best_i
    return best_d.astype(np.float32), best_i


def knn_f32_reference_form(dataset: np.ndarray, queries: np.ndarray, k: int,
                           metric: str = "sqeuclidean", block: int = 32768):
    """fp32 brute force in the arithmetic the reference's engines use; the timed CPU baseline.

    sqeuclidean: expanded |q|^2 + |b|^2 - 2 q.b, one sgemm per base block (libraft's tiled
    brute force, k > 64 path, behind cu_knn.py:243).  cosine: 1 - q.b / (|q||b|)
    (cu_knn.py:247-248).  dot: 1 - q.b then smallest-k (cu_knn.py:256-257).
    """
    _check(dataset, queries, k, metric)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    qsq = np.einsum("ij,ij->i", q, q)
    nq, nb = q.shape[0], dataset.shape[0]
    best_d = np.empty((nq, 0), dtype=np.float32)
    best_i = np.empty((nq, 0), dtype=np.int64)
    rows = np.arange(nq)[:, None]
    for start in range(0, nb, block):
        b = np.ascontiguousarray(dataset[start:start + block], dtype=np.float32)
        dist = q @ b.T
        if metric == "dot":
            np.subtract(np.float32(1.0), dist, out=dist)
        else:
            bsq = np.einsum("ij,ij->i", b, b)
            if metric == "sqeuclidean":
                dist *= np.float32(-2.0)
                dist += qsq[:, None]
                dist += bsq[None, :]
                np.maximum(dist, np.float32(0.0), out=dist)
            else:
                dist /= (np.sqrt(qsq)[:, None] * np.sqrt(bsq)[None, :])
                np.subtract(np.float32(1.0), dist, out=dist)
        kk = min(k, b.shape[0])
        if b.shape[0] > kk:
            part = np.argpartition(dist, kk - 1, axis=1)[:, :kk]
        else:
            part = np.broadcast_to(np.arange(b.shape[0]), (nq, b.shape[0]))
        d_blk = dist[rows, part]
        i_blk = part.astype(np.int64) + start
        cat_d = np.concatenate([best_d, d_blk], axis=1)
        cat_i = np.concatenate([best_i, i_blk], axis=1)
        kk = min(k, cat_d.shape[1])
        if cat_d.shape[1] > kk:
            sel = np.argpartition(cat_d, kk - 1, axis=1)[:, :kk]
            cat_d, cat_i = cat_d[rows, sel], cat_i[rows, sel]
        order = np.argsort(cat_d, axis=1, kind="stable")
        best_d = np.take_along_axis(cat_d, order, axis=1)
        best_i = np.take_along_axis(cat_i, order, axis=1)
    return best_d, best_i


def process_batches_oracle(base: np.ndarray, queries: np.ndarray, batch_count: int,
                           batch_size: int, k: int, split: bool, engine: str = "raft"):
    """Array-level restatement of the loop of cu_knn.py:193-287.

    Returns ``[(distances_i, indices_i)]`` for ``i in range(batch_count)``, each ``[Q, k]``
    -- the content of ``partial/distances{i}.parquet`` / ``indices{i}.parquet``
    (cu_knn.py:284-285), indices already global (``+ batch_offset``, cu_knn.py:273).
    """
    metric = ENGINE_METRIC[engine]
    nq = len(queries)
    partials = []
    for start in range(batch_count):
        batch_offset = start * batch_size                                  # :202
        batch_length = batch_size if start != batch_count - 1 else len(base) - batch_offset  # :203
        dataset = base[batch_offset:batch_offset + batch_length]
        splits = batch_count if split else 1                               # :216-223
        rows_per_split = nq // (1 * batch_count)                           # :218 (split_factor = 1)
        d_parts, i_parts = [], []
        for i in range(splits):
            offset = i * rows_per_split                                    # :226
            length = rows_per_split if i != splits - 1 else nq - offset    # :227
            assert k <= len(dataset)                                       # :237
            d, idx = knn_f64(dataset, queries[offset:offset + length], k, metric)
            d_parts.append(d)
            i_parts.append(idx + batch_offset)                             # :273
        dist = np.concatenate(d_parts, axis=0)
        ind = np.concatenate(i_parts, axis=0)
        assert len(dist) == nq and len(ind) == nq                          # :281-282
        partials.append((dist, ind))
    return partials


def batch_plan(base_rows: int, batch_size: int, k: int):
    """cu_knn.py:176-178: number of base batches and the last-batch-size assertion."""
    batch_count = math.ceil(base_rows / batch_size)
    assert (base_rows % batch_size == 0) or k <= (base_rows % batch_size), (
        f"Cannot generate k of {k} with only {base_rows} rows and batch_size of {batch_size}.")
    return batch_count


# --------------------------------------------------------------------------------------
# comparison helper shared by the parity tests (tolerances from BASELINE.json north_star)
# --------------------------------------------------------------------------------------
def compare_knn(got_d: np.ndarray, got_i: np.ndarray, dataset: np.ndarray, queries: np.ndarray,
                metric: str, truth_d64: np.ndarray, truth_i: np.ndarray,
                rtol: float = 1e-5, atol: float = 1e-6, tie_rtol: float = 1e-6,
                id_offset: int = 0) -> dict:
    """Check a result against float64 truth.

    * distances: |got - truth| <= atol + rtol * |truth| for every slot
    * indices: identical, except where the competing true distances tie within
      ``tie_rtol`` relative (+ the same absolute floor scaled down); for a mismatching slot
      the returned neighbour's own float64 distance must tie with the truth's k-slot value.
    Returns counters; raises AssertionError on a real mismatch.
    """
    got_d = np.asarray(got_d)
    got_i = np.asarray(got_i)
    nq, k = truth_i.shape
    assert got_d.shape == (nq, k) and got_i.shape == (nq, k)
    err = np.abs(got_d.astype(np.float64) - truth_d64)
    tol = atol + rtol * np.abs(truth_d64)
    bad = np.argwhere(err > tol)
    assert bad.size == 0, (f"{len(bad)} distances outside tolerance, first at {bad[0]}: "
                           f"got {got_d[tuple(bad[0])]} truth {truth_d64[tuple(bad[0])]}")
    # rows ascending (merge.py:104 asserts this of every merged row)
    assert np.all(np.diff(got_d, axis=1) >= 0), "distances not ascending"
    mism = np.argwhere(got_i != truth_i)
    tie_swaps = 0
    if mism.size:
        q64 = np.asarray(queries, dtype=np.float64)
        for r, c in mism:
            j = int(got_i[r, c]) - id_offset
            assert 0 <= j < len(dataset), f"index out of range at {(r, c)}: {got_i[r, c]}"
            # (only the rows in question are widened: the dataset may be a multi-GB shard)
            dj = distances_f64(np.asarray(dataset[j:j + 1], dtype=np.float64), q64[r:r + 1], metric)[0, 0]
            ref = truth_d64[r, c]
            tie_tol = tie_rtol * max(abs(ref), abs(dj)) + atol * 1e-1
            assert abs(dj - ref) <= tie_tol, (
                f"index mismatch at {(int(r), int(c))} is not a tie: got id {got_i[r, c]} "
                f"(d={dj!r}) truth id {truth_i[r, c]} (d={ref!r})")
            tie_swaps += 1
        # a row must not repeat a neighbour
        for r in np.unique(mism[:, 0]):
            assert len(np.unique(got_i[r])) == k, f"duplicate neighbour ids in row {r}"
    return {"max_abs_err": float(err.max()) if err.size else 0.0,
            "index_mismatches_tied": tie_swaps, "slots": int(nq * k)}
