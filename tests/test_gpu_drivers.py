"""The reference-facing Python entry points on real parquet files.  -m gpu."""
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from neighborhoodwatch_b200 import synth
from oracle import knn_oracle as ko

pytestmark = pytest.mark.gpu


def write_vectors(path, vecs, prefix="embedding_", extra=True):
    """Fixture layout of tests/test_02_cu_knn.py:15-41: metadata columns + one column per dimension."""
    cols = {}
    if extra:
        cols["document_id_idx"] = pa.array(np.arange(len(vecs), dtype=np.int64))
        cols["text"] = pa.array([f"t{i}" for i in range(len(vecs))])
        cols["title"] = pa.array(["x"] * len(vecs))
        cols["document_id"] = pa.array(np.arange(len(vecs), dtype=np.int64))
    for j in range(vecs.shape[1]):
        cols[f"{prefix}{j}"] = pa.array(vecs[:, j])
    pq.write_table(pa.table(cols), path, row_group_size=257)


def read_matrix(path, dtype):
    t = pq.read_table(path)
    assert t.schema.names == [str(i) for i in range(t.num_columns)]
    return np.stack([t.column(j).to_numpy() for j in range(t.num_columns)], axis=1).astype(dtype, copy=False), t


@pytest.mark.parametrize("split", [True, False])
def test_compute_knn_and_merge(tmp_path, cuda_device, split):
    from neighborhoodwatch_b200 import cu_knn, merge
    from neighborhoodwatch_b200.nw_utils import get_partial_distances_filename, get_partial_indices_filename
    data_dir = str(tmp_path)
    nq, nb, d, k = 100, 1000, 384, 10                       # tests/test_02_cu_knn.py:67-71
    base, q = synth.vectors(nb, d, 1234), synth.vectors(nq, d, 5678)
    write_vectors(os.path.join(data_dir, "base.parquet"), base)
    write_vectors(os.path.join(data_dir, "query.parquet"), q)
    cu_knn.compute_knn(data_dir, "intfloat/e5-small-v2", d, "query.parquet", nq, "base.parquet", nb,
                       mem_tune=False, k=k, initial_batch_size=300, split=split)
    partials = ko.process_batches_oracle(base, q, 4, 300, k, split, engine="raft")
    for i in range(4):
        dist, t_d = read_matrix(get_partial_distances_filename(data_dir, i), np.float32)
        idx, t_i = read_matrix(get_partial_indices_filename(data_dir, i), np.int64)
        assert t_d.schema.field(0).type == pa.float32() and t_i.schema.field(0).type == pa.int64()
        np.testing.assert_array_equal(idx, partials[i][1])
        np.testing.assert_allclose(dist, partials[i][0], rtol=1e-5, atol=1e-6)
    merge.merge_indices_and_distances(data_dir, k)
    f_d, _ = read_matrix(get_partial_distances_filename(data_dir, -1), np.float32)
    f_i, _ = read_matrix(get_partial_indices_filename(data_dir, -1), np.int64)
    t_d, t_i = ko.knn_f64(base, q, k, "sqeuclidean")
    np.testing.assert_array_equal(f_i, t_i)
    np.testing.assert_allclose(f_d, t_d, rtol=1e-5, atol=1e-6)
    # dimension mismatch -> AssertionError (cu_knn.py:136)
    with pytest.raises(AssertionError):
        cu_knn.compute_knn(data_dir, "intfloat/e5-small-v2", d + 1, "query.parquet", nq, "base.parquet", nb, k=k)
    # last batch smaller than k -> AssertionError (cu_knn.py:177)
    with pytest.raises(AssertionError):
        cu_knn.compute_knn(data_dir, "intfloat/e5-small-v2", d, "query.parquet", nq, "base.parquet", nb, k=k,
                           initial_batch_size=333)


def test_compute_knn_ds_streaming(tmp_path, cuda_device):
    from neighborhoodwatch_b200 import cu_knn_ds, merge
    from neighborhoodwatch_b200.nw_utils import get_partial_distances_filename, get_partial_indices_filename
    data_dir = str(tmp_path)
    nq, nb, d, k = 64, 5000, 128, 100
    base, q = synth.vectors(nb, d, 1234), synth.vectors(nq, d, 5678)
    write_vectors(os.path.join(data_dir, "base.parquet"), base)
    write_vectors(os.path.join(data_dir, "query.parquet"), q)
    cu_knn_ds.compute_knn_ds(data_dir, d, "query.parquet", nq, "base.parquet", nb, k=k, initial_batch_size=1200)
    f_d, _ = read_matrix(get_partial_distances_filename(data_dir, -1), np.float32)
    f_i, _ = read_matrix(get_partial_indices_filename(data_dir, -1), np.int64)
    t_d, t_i = ko.knn_f64(base, q, k, "sqeuclidean")
    np.testing.assert_array_equal(f_i, t_i)
    np.testing.assert_allclose(f_d, t_d, rtol=1e-5, atol=1e-6)
    merge.merge_indices_and_distances(data_dir, k)         # no numbered partials: must be a no-op (merge.py:42)
    f_i2, _ = read_matrix(get_partial_indices_filename(data_dir, -1), np.int64)
    np.testing.assert_array_equal(f_i2, t_i)


def test_colbert_token_knn(tmp_path, cuda_device):
    from neighborhoodwatch_b200 import colbert_knn, merge
    from neighborhoodwatch_b200.nw_utils import get_partial_indices_filename
    data_dir = str(tmp_path)
    nq, nb, d, k = 96, 20000, 128, 100                      # 3 queries x 32 tokens vs ~111 docs x 180 tokens
    base, q = synth.vectors(nb, d, 1234), synth.vectors(nq, d, 5678)
    write_vectors(os.path.join(data_dir, "b.parquet"), base, prefix="token_embedding_", extra=False)
    write_vectors(os.path.join(data_dir, "q.parquet"), q, prefix="token_embedding_", extra=False)
    colbert_knn.process_knn_computation(data_dir, os.path.join(data_dir, "b.parquet"), nb,
                                        os.path.join(data_dir, "q.parquet"), nq, k=k)
    assert merge.get_file_count(os.path.join(data_dir, "partial")) == 3      # ceil(20000 / 9000)
    merge.merge_indices_and_distances(data_dir, k)
    f_i, _ = read_matrix(get_partial_indices_filename(data_dir, -1), np.int64)
    t_d, t_i = ko.knn_f64(base, q, k, "dot")
    np.testing.assert_array_equal(f_i, t_i)


@pytest.mark.parametrize("rows,d,ld_extra,shift", [
    (1000, 384, 0, 0),        # both sides vectorised (16-byte aligned, ld % 4 == 0, d % 4 == 0)
    (1000, 384, 24, 0),       # staging wider than the rows in use (the streaming driver's slabs)
    (257, 100, 0, 0),         # ld % 4 != 0: scalar reads, float4 writes
    (1001, 130, 3, 0),        # d % 4 != 0: float4 reads, scalar writes
    (63, 7, 0, 0),            # smaller than a tile on both sides
    (4100, 65, 0, 1),         # source 4 bytes off a 16-byte boundary
    (130, 1, 0, 0),
])
def test_rows_from_columns_shapes(cuda_device, rows, d, ld_extra, shift):
    """nwk_rows_from_columns_f32 (the reference's cupy.copy(order='C'), cu_knn.py:211,235) on ragged shapes:
    every combination of vectorised / scalar sides and partial tiles, bit-exact against numpy's transpose."""
    import torch
    from neighborhoodwatch_b200 import knn
    rng = np.random.default_rng(rows * 131 + d)
    ld = rows + ld_extra
    host = rng.standard_normal((d, ld)).astype(np.float32)
    flat = torch.empty(d * ld + 4, dtype=torch.float32, device=cuda_device)
    cols = flat[shift:shift + d * ld].view(d, ld)
    cols.copy_(torch.from_numpy(host))
    out = knn.rows_from_columns(cols, rows)
    assert out.shape == (rows, d)
    assert np.array_equal(out.cpu().numpy(), host[:, :rows].T)
    # into a caller's larger buffer: rows beyond `rows` stay untouched
    buf = torch.full((rows + 5, d), -7.0, dtype=torch.float32, device=cuda_device)
    got = knn.rows_from_columns(cols, rows, out=buf)
    assert np.array_equal(got.cpu().numpy(), host[:, :rows].T)
    assert bool((buf[rows:] == -7.0).all())
