"""CPU oracle for the partial-result merge.  TEST INFRASTRUCTURE ONLY (see knn_oracle.py).

Restates ``neighborhoodwatch/merge.py:39-114`` on plain arrays: the reference folds the
partial ``[Q, k]`` lists pairwise -- column-concatenate the running result with the next
partial (``:75-76``), ``np.argsort(axis=1)`` (``:83``), ``take_along_axis`` on both arrays
(``:90-91``), keep the first k columns (``:101-102``) -- and asserts every row ascending
(``:104``).  ``get_file_count`` (``:15-27``) counts ``indices<N>.parquet`` files.

Pinned by ``tests/golden/merge_*.npz``: fixtures produced by running the reference's own
``merge_indices_and_distances`` on seeded partial files (``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import glob
import os
import re

import numpy as np


def file_count(partial_dir: str) -> int:
    """merge.py:15-27 -- number of files named ``indices<digits>.parquet``."""
    pattern = re.compile(re.escape(partial_dir) + r"/indices(\d+)\.parquet")
    return sum(1 for f in sorted(glob.glob(os.path.join(partial_dir, "indices*.parquet")))
               if pattern.match(f))


def merge_pairwise(partials, k: int, stable: bool = False):
    """Fold ``[(dist [Q,k_i], idx [Q,k_i]), ...]`` the way merge.py:56-101 does.

    ``stable=False`` mirrors the reference (numpy's default introsort -- tie order between
    equal distances is unspecified there); ``stable=True`` gives the deterministic order
    the CUDA merge implements (equal distances: the entry that appears first in the
    concatenation, i.e. lower partial number, wins; within one partial the given order).
    """
    final_d = final_i = None
    for dist, idx in partials:
        dist = np.asarray(dist)
        idx = np.asarray(idx)
        if final_d is None:                       # merge.py:67-69
            final_d, final_i = dist, idx
            continue
        cat_d = np.concatenate([final_d, dist], axis=1)    # :75
        cat_i = np.concatenate([final_i, idx], axis=1)     # :76
        order = np.argsort(cat_d, axis=1, kind="stable" if stable else None)   # :83
        final_d = np.take_along_axis(cat_d, order, axis=1)[:, :k]              # :90,101
        final_i = np.take_along_axis(cat_i, order, axis=1)[:, :k]              # :91,102
        assert np.all(np.diff(final_d, axis=1) >= 0)                           # :104
    return final_d, final_i


def merge_by_distance_then_index(partials, k: int):
    """Order-independent merge: global k smallest by (distance, index).

    This is the tie rule of the CUDA merge kernel; it equals ``merge_pairwise`` wherever
    distances within a row are distinct.
    """
    cat_d = np.concatenate([np.asarray(d) for d, _ in partials], axis=1)
    cat_i = np.concatenate([np.asarray(i) for _, i in partials], axis=1)
    out_d = np.empty((cat_d.shape[0], min(k, cat_d.shape[1])), dtype=cat_d.dtype)
    out_i = np.empty(out_d.shape, dtype=cat_i.dtype)
    for r in range(cat_d.shape[0]):
        order = np.lexsort((cat_i[r], cat_d[r]))[:out_d.shape[1]]
        out_d[r] = cat_d[r, order]
        out_i[r] = cat_i[r, order]
    return out_d, out_i
