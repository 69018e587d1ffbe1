#!/usr/bin/env python
"""Generate golden fixtures by running the REFERENCE's own code (read-only /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

What is executed unmodified from /root/reference/neighborhoodwatch:
  * ``cu_knn.process_batches(..., engine='torch')`` -- the batch / split / offset / concat /
    column-naming / parquet-writing loop of cu_knn.py:193-287 with the torch engine's own
    arithmetic (``1 - torch.matmul(q, b.T)``; ``torch.topk(largest=False)``, cu_knn.py:253-260)
    on CPU torch;
  * ``merge.merge_indices_and_distances`` (merge.py:39-114).

The GPU dataframe stack the reference imports (cudf / cupy / rmm / pylibraft / cuvs) and the
SaaS SDKs pulled in through nw_utils -> model_generator are not installable here; they are
replaced by the thin numpy / pandas shims below, which only move data (no arithmetic of the
hot path lives in them).  The pylibraft / cuvs engines themselves cannot be run, so the
'raft' and 'cuvs' metrics have no golden vectors (see oracle/knn_oracle.py header).

Outputs: tests/golden/*.npz (inputs are regenerated from seeds by
neighborhoodwatch_b200.synth; only results are stored).
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from unittest import mock

import numpy as np
import pandas as pd
import pyarrow as pa
import pyarrow.parquet as pq

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"
sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------------------
# shims for the libraries the reference imports but that only move data on this path
# ----------------------------------------------------------------------------------------
class FakeFrame(pd.DataFrame):
    """Stands in for cudf.DataFrame: a pandas frame with the three cuDF-only calls used."""

    @property
    def _constructor(self):
        return FakeFrame

    @classmethod
    def from_arrow(cls, table):
        if isinstance(table, pa.RecordBatch):
            table = pa.Table.from_batches([table])
        return cls(table.to_pandas())

    def to_arrow(self):
        return pa.Table.from_pandas(pd.DataFrame(self))

    def to_dlpack(self):
        # cuDF exports column-major; the reference then makes it C-order (cu_knn.py:211)
        return np.asfortranarray(self.to_numpy())


class FakeArray:
    """Stands in for cupy.ndarray."""

    def __init__(self, a):
        self.a = np.asarray(a)

    def copy(self, order="C"):
        return FakeArray(np.array(self.a, order=order, copy=True))

    def astype(self, dt):
        return FakeArray(self.a.astype(dt))

    def get(self):
        return self.a

    def toDlpack(self):
        return np.ascontiguousarray(self.a)

    def __add__(self, other):
        return FakeArray(self.a + other)

    def __len__(self):
        return len(self.a)

    @property
    def shape(self):
        return self.a.shape


def _install_shims():
    cudf = types.ModuleType("cudf")
    cudf.DataFrame = FakeFrame
    cudf.from_pandas = lambda df: FakeFrame(df)
    cudf.concat = lambda frames, ignore_index=False: FakeFrame(
        pd.concat([pd.DataFrame(f) for f in frames], ignore_index=ignore_index))
    cupy = types.ModuleType("cupy")

    def from_dlpack(x):
        if isinstance(x, np.ndarray):
            return FakeArray(x)
        return FakeArray(np.from_dlpack(x))
    cupy.from_dlpack = from_dlpack
    cupy.asarray = lambda x: x if isinstance(x, FakeArray) else FakeArray(x)
    rmm = mock.MagicMock()
    sys.modules.update({"cudf": cudf, "cupy": cupy, "rmm": rmm})
    for name in ["pylibraft", "pylibraft.neighbors", "pylibraft.neighbors.brute_force", "cuvs", "cuvs.neighbors",
                 "cuvs.distance", "datasets", "sentence_transformers", "vertexai", "vertexai.preview",
                 "vertexai.preview.language_models", "vertexai.language_models", "colbert", "colbert.infra",
                 "colbert.infra.config", "colbert.modeling", "colbert.modeling.checkpoint", "colbert.indexing",
                 "colbert.indexing.collection_encoder", "cohere", "voyageai", "openai", "spacy", "h5py",
                 "nvidia_smi", "google", "google.cloud", "google.cloud.aiplatform"]:
        sys.modules.setdefault(name, mock.MagicMock())
    sys.path.insert(0, REFERENCE)


def _write_vectors(path, vecs):
    pq.write_table(pa.table({f"embedding_{i}": pa.array(vecs[:, i]) for i in range(vecs.shape[1])}), path)


def _read(path):
    t = pq.read_table(path)
    return np.stack([t.column(n).to_numpy() for n in t.schema.names if n.isdigit()], axis=1), t.schema


def golden_process_batches(name, nb, nq, d, k, batch_size, split, base_seed=1234, query_seed=5678):
    import math

    from neighborhoodwatch import cu_knn as ref_knn
    from neighborhoodwatch import merge as ref_merge
    from neighborhoodwatch.nw_utils import get_partial_distances_filename, get_partial_indices_filename

    from neighborhoodwatch_b200 import synth
    base = synth.vectors(nb, d, base_seed)
    q = synth.vectors(nq, d, query_seed)
    with tempfile.TemporaryDirectory() as data_dir:
        os.makedirs(os.path.join(data_dir, "partial"))
        _write_vectors(os.path.join(data_dir, "b.parquet"), base)
        _write_vectors(os.path.join(data_dir, "q.parquet"), q)
        base_table = pq.read_table(os.path.join(data_dir, "b.parquet"))
        query_table = pq.read_table(os.path.join(data_dir, "q.parquet"))
        batch_count = math.ceil(nb / batch_size)
        ref_knn.process_batches(data_dir, base_table, query_table, batch_count, batch_size, k, split, engine="torch")
        out = {"kind": "process_batches", "engine": "torch", "nb": nb, "nq": nq, "d": d, "k": k,
               "batch_size": batch_size, "batch_count": batch_count, "split": split, "base_seed": base_seed,
               "query_seed": query_seed}
        schemas = []
        for i in range(batch_count):
            dist, s_d = _read(get_partial_distances_filename(data_dir, i))
            idx, s_i = _read(get_partial_indices_filename(data_dir, i))
            out[f"distances{i}"] = dist
            out[f"indices{i}"] = idx
            schemas.append((s_d.names, str(s_d.field(0).type), str(s_i.field(0).type)))
        assert all(s == schemas[0] for s in schemas)
        out["column_names"] = np.array(schemas[0][0])
        out["distance_type"] = schemas[0][1]
        out["index_type"] = schemas[0][2]
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        # and the merge of exactly these partial files by the reference's merge.py
        ref_merge.merge_indices_and_distances(data_dir, k)
        f_d, _ = _read(get_partial_distances_filename(data_dir, -1))
        f_i, _ = _read(get_partial_indices_filename(data_dir, -1))
        m = {"kind": "merge", "P": batch_count, "k": k, "final_distances": f_d, "final_indices": f_i}
        distinct = True
        for i in range(batch_count):
            m[f"distances{i}"] = out[f"distances{i}"]
            m[f"indices{i}"] = out[f"indices{i}"]
        cat = np.concatenate([out[f"distances{i}"] for i in range(batch_count)], axis=1)
        distinct = all(len(np.unique(row)) == len(row) for row in cat)
        m["distinct"] = distinct
        np.savez_compressed(os.path.join(HERE, f"merge_of_{name}.npz"), **m)
        print(name, "batches", batch_count, "distinct distances per row:", distinct)


def golden_merge_random(name, P, nq, k, seed, ties):
    from neighborhoodwatch import merge as ref_merge
    from neighborhoodwatch.nw_utils import get_partial_distances_filename, get_partial_indices_filename
    rng = np.random.default_rng(seed)
    with tempfile.TemporaryDirectory() as data_dir:
        os.makedirs(os.path.join(data_dir, "partial"))
        m = {"kind": "merge", "P": P, "k": k}
        for i in range(P):
            dist = np.sort(rng.random((nq, k)).astype(np.float32), axis=1)
            if ties:
                dist = np.round(dist * 16) / np.float32(16)            # heavy ties, still sorted
            idx = rng.integers(0, 10 ** 9, size=(nq, k), dtype=np.int64)
            cols = [str(j) for j in range(k)]
            pq.write_table(pa.Table.from_pandas(pd.DataFrame(dist, columns=cols)), get_partial_distances_filename(data_dir, i))
            pq.write_table(pa.Table.from_pandas(pd.DataFrame(idx, columns=cols)), get_partial_indices_filename(data_dir, i))
            m[f"distances{i}"] = dist
            m[f"indices{i}"] = idx
        ref_merge.merge_indices_and_distances(data_dir, k)
        m["final_distances"], _ = _read(get_partial_distances_filename(data_dir, -1))
        m["final_indices"], _ = _read(get_partial_indices_filename(data_dir, -1))
        cat = np.concatenate([m[f"distances{i}"] for i in range(P)], axis=1)
        m["distinct"] = all(len(np.unique(row)) == len(row) for row in cat)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **m)
        print(name, "distinct:", m["distinct"])


def golden_fvec_ivec(name):
    """Bytes written by the reference's own per-row writer (parquet_to_format.py:71-89)."""
    from neighborhoodwatch import parquet_to_format as ref_fmt
    rng = np.random.default_rng(7)
    vec = rng.random((23, 12)).astype(np.float32)
    idx = rng.integers(0, 2 ** 31 - 1, size=(23, 10), dtype=np.int64)
    dist = np.sort(rng.random((23, 10)).astype(np.float32), axis=1)
    with tempfile.TemporaryDirectory() as data_dir:
        out = {"kind": "fvec_ivec", "vectors": vec, "indices": idx, "distances": dist}
        ref_fmt.write_ivec_fvec_from_dataframe(data_dir, "m", "v.fvec", pd.DataFrame(vec), "f", 12)
        ref_fmt.write_ivec_fvec_from_dataframe(data_dir, "m", "i.ivec", pd.DataFrame(idx), "i", 10)
        ref_fmt.write_ivec_fvec_from_dataframe(data_dir, "m", "d.fvec", pd.DataFrame(dist), "f", 10)
        # a frame with a trailing RowNum column is trimmed by the reference (:79-80)
        with_rownum = pd.DataFrame(idx)
        with_rownum["RowNum"] = np.arange(len(idx))
        ref_fmt.write_ivec_fvec_from_dataframe(data_dir, "m", "r.ivec", with_rownum, "i", 10)
        for fn in ("v.fvec", "i.ivec", "d.fvec", "r.ivec"):
            out[fn.replace(".", "_")] = np.frombuffer(open(os.path.join(data_dir, fn), "rb").read(), dtype=np.uint8)
        out["count_v"] = ref_fmt.count_vectors(data_dir, "v.fvec")
        out["nth_v_5"] = np.array(ref_fmt.get_nth_vector(data_dir, "v.fvec", 5), dtype=np.float32)
        out["nth_i_7"] = np.array(ref_fmt.get_nth_vector(data_dir, "i.ivec", 7), dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, "ok")


if __name__ == "__main__":
    assert os.path.isdir(REFERENCE), "run this in the build container: /root/reference is not mounted here"
    _install_shims()
    # test_02_cu_knn scale (Q=100, B=1000, k=10), 4 base batches x 4 query splits
    golden_process_batches("pb_torch_split", nb=1000, nq=100, d=64, k=10, batch_size=300, split=True)
    # ck style: split=False, k=100, ragged last batch (>= k rows)
    golden_process_batches("pb_torch_nosplit", nb=700, nq=37, d=128, k=100, batch_size=250, split=False)
    golden_merge_random("merge_random_p5", P=5, nq=40, k=100, seed=1, ties=False)
    golden_merge_random("merge_ties_p3", P=3, nq=20, k=10, seed=2, ties=True)
    golden_fvec_ivec("fvec_ivec_writer")
