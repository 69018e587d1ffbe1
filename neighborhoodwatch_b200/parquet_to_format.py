"""fvec / ivec output and the post-validator -- the step after the hot path.

* ``write_ivec_fvec_from_dataframe`` keeps the reference's signature and byte layout
  (neighborhoodwatch/parquet_to_format.py:71-89: per row ``<int32 dim, little endian><dim x 4
  bytes>``, indices cast to int32, everything else to float32, a trailing extra column dropped)
  but writes whole blocks with one numpy structured array instead of a Python loop per row --
  the reference's loop is the largest stage of its only published run (197 s of 303 s,
  docs/cli.png).
* ``count_vectors`` / ``get_nth_vector`` mirror parquet_to_format.py:37-68.
* ``validate_files_v0`` applies the reference's acceptance rule (parquet_to_format.py:351-383:
  for every query and every returned index ``isclose(1 - dot(q, base[idx]), distance / 2)``,
  i.e. squared L2 on unit vectors) to all Q*k pairs on the GPU (``nwk_validate_pairs``).
"""
from __future__ import annotations

import ctypes
import os
import struct

import numpy as np

from .nw_utils import get_full_filename, output_dimension_validity_check

_BLOCK_ROWS = 1 << 16


def _as_matrix(df) -> np.ndarray:
    if hasattr(df, "to_numpy"):
        return df.to_numpy()
    return np.asarray(df)


def write_ivec_fvec_from_dataframe(data_dir, model_name, filename, df, type_char, num_columns):
    """parquet_to_format.py:71-89, vectorised; byte-identical output."""
    full_filename = get_full_filename(data_dir, filename)
    values = _as_matrix(df)
    if values.ndim != 2:
        raise ValueError("expected a 2-D frame")
    if values.shape[1] == num_columns + 1:          # potentially remove rownum field (:79-80)
        values = values[:, :-1]
    assert output_dimension_validity_check(model_name, num_columns, values.shape[1]), \
        f"Expected {num_columns} values, got {values.shape[1]} for model {model_name} [filename: {filename}]"
    dim = values.shape[1]
    item = np.dtype("<f4") if type_char == "f" else np.dtype("<i4")      # :84-86
    record = np.dtype([("dim", "<i4"), ("vec", item, (dim,))])
    with open(full_filename, "wb") as f:
        for start in range(0, len(values), _BLOCK_ROWS):
            block = values[start:start + _BLOCK_ROWS]
            out = np.empty(len(block), dtype=record)
            out["dim"] = dim
            out["vec"] = block.astype(item, copy=False)
            f.write(out.tobytes())


def count_vectors(data_dir, filename):
    """parquet_to_format.py:37-47: file size / record size (record size from the first header)."""
    full_filename = get_full_filename(data_dir, filename)
    with open(full_filename, "rb") as f:
        dimension = struct.unpack("<i", f.read(4))[0]
    return os.path.getsize(full_filename) // (4 + 4 * dimension)


def get_nth_vector(data_dir, filename, n):
    """parquet_to_format.py:50-68: the n-th record of an fvec (float32) / ivec (int32) file."""
    full_filename = get_full_filename(data_dir, filename)
    dtype = "<i4" if full_filename.endswith("ivec") else "<f4"
    with open(full_filename, "rb") as f:
        dimension = struct.unpack("<i", f.read(4))[0]
        f.seek(4 * n * (1 + dimension) + 4)
        return tuple(np.frombuffer(f.read(4 * dimension), dtype=dtype).tolist())


def read_vec_file(data_dir, filename) -> np.ndarray:
    """Whole fvec / ivec file as a ``[rows, dim]`` array."""
    full_filename = get_full_filename(data_dir, filename)
    item = "<i4" if full_filename.endswith("ivec") else "<f4"
    with open(full_filename, "rb") as f:
        dimension = struct.unpack("<i", f.read(4))[0]
    rec = np.dtype([("dim", "<i4"), ("vec", item, (dimension,))])
    data = np.fromfile(full_filename, dtype=rec)
    assert np.all(data["dim"] == dimension)
    return np.ascontiguousarray(data["vec"])


def validate_pairs(base, queries, indices, distances, rtol: float = 1e-5, atol: float = 1e-8) -> int:
    """Number of (query, neighbour) pairs violating ``isclose(1 - q.b, distance / 2)``; CUDA tensors."""
    import torch

    from . import _native
    lib = _native.load()
    if base.device.type != "cuda":
        raise RuntimeError("validate_pairs runs on a CUDA device; there is no CPU path")
    base = base.to(torch.float32).contiguous()
    queries = queries.to(torch.float32).contiguous()
    indices = indices.to(torch.int64).contiguous()
    distances = distances.to(torch.float32).contiguous()
    nq, k = indices.shape
    out = ctypes.c_int64(0)
    with torch.cuda.device(base.device):
        rc = lib.nwk_validate_pairs(base.data_ptr(), base.shape[0], queries.data_ptr(), nq, base.shape[1],
                                    indices.data_ptr(), distances.data_ptr(), k, rtol, atol, ctypes.byref(out),
                                    torch.cuda.current_stream(base.device).cuda_stream)
        _native.check(rc)
    return int(out.value)


def validate_files_v0(data_dir, query_vector_fvec, base_vector_fvec, indices_ivec, distances_fvec) -> int:
    """parquet_to_format.py:351-383 on the GPU: returns (and prints) the total mismatch count."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device())
    q = torch.from_numpy(read_vec_file(data_dir, query_vector_fvec)).to(dev)
    b = torch.from_numpy(read_vec_file(data_dir, base_vector_fvec)).to(dev)
    i = torch.from_numpy(read_vec_file(data_dir, indices_ivec).astype(np.int64)).to(dev)
    d = torch.from_numpy(read_vec_file(data_dir, distances_fvec)).to(dev)
    total = validate_pairs(b, q, i, d)
    print(f"Total mismatch count: {total}")
    return total
