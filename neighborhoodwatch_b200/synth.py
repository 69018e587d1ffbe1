"""Deterministic synthetic vectors in the reference's fixture style.

The reference's tests build their inputs as ``np.random.rand(n).astype(float32)`` per
``embedding_i`` column and L2-normalise each row (tests/test_02_cu_knn.py:21-28,
nw_utils.py:56-59).  This module produces vectors of the same distribution from a
counter-based hash of ``(seed, row, col)`` so that numpy (here) and CUDA
(csrc/nwk_synth.cu, ``nwk_synth_fill``) generate bit-identical data and any slab of a
very large base can be regenerated on demand.

    u(row, col) = splitmix64(seed * GOLDEN + row * d + col) >> 40          (24 bits)
    S(row)      = sum_col u^2                                              (exact, uint64)
    y(row, col) = float32( float64(u) / sqrt(float64(S)) )
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 1234
QUERY_SEED = 5678
_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _u24(seed: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = np.uint64(seed) * _GOLDEN + idx
        z = z + _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return z >> np.uint64(40)


def vectors(rows: int, d: int, seed: int, row0: int = 0, normalize: bool = True,
            block_rows: int = 16384) -> np.ndarray:
    """``[rows, d]`` float32 array holding synthetic rows ``row0 .. row0+rows-1``."""
    out = np.empty((rows, d), dtype=np.float32)
    cols = np.arange(d, dtype=np.uint64)[None, :]
    for start in range(0, rows, block_rows):
        n = min(block_rows, rows - start)
        r = (np.arange(row0 + start, row0 + start + n, dtype=np.uint64) * np.uint64(d))[:, None]
        u = _u24(seed, r + cols)
        if normalize:
            s = (u * u).sum(axis=1, dtype=np.uint64)
            norm = np.sqrt(s.astype(np.float64))
            out[start:start + n] = (u.astype(np.float64) / norm[:, None]).astype(np.float32)
        else:
            out[start:start + n] = u.astype(np.float32) * np.float32(2.0 ** -24)
    return out


def device_vectors(rows: int, d: int, seed: int, row0: int = 0, normalize: bool = True, device="cuda",
                   out=None):
    """The same rows generated on the GPU by ``nwk_synth_fill`` (bit-identical)."""
    import torch

    from . import _native
    lib = _native.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("device_vectors generates on a CUDA device; use vectors() for numpy")
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((rows, d), dtype=torch.float32, device=dev)
        assert out.shape == (rows, d) and out.is_contiguous() and out.dtype == torch.float32
        rc = lib.nwk_synth_fill(out.data_ptr(), rows, d, seed, row0, 1 if normalize else 0,
                                torch.cuda.current_stream(dev).cuda_stream)
        _native.check(rc)
    return out
