"""Host-side logic that needs no GPU: naming, planning, slab iteration, parquet schema,
the sharded path under gloo with world_size 2."""
import math
import os
import socket
import sys

import numpy as np
import pyarrow as pa
import pyarrow.dataset as ds
import pyarrow.parquet as pq
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neighborhoodwatch_b200 import cu_knn, cu_knn_ds, merge, nw_utils, synth
from neighborhoodwatch_b200 import dist as nwdist
from oracle import knn_oracle as ko
from oracle import merge_oracle as mo


def test_filenames():
    assert nw_utils.get_partial_indices_filename("/d", 3) == "/d/partial/indices3.parquet"
    assert nw_utils.get_partial_indices_filename("/d", -1) == "/d/partial/final_indices.parquet"
    assert nw_utils.get_partial_distances_filename("/d", 0) == "/d/partial/distances0.parquet"
    assert nw_utils.get_partial_distances_filename("/d", -1) == "/d/partial/final_distances.parquet"
    assert nw_utils.get_full_filename("/d", "x.parquet") == "/d/x.parquet"
    assert nw_utils.get_full_filename("/d", "/d/x.parquet") == "/d/x.parquet"


def test_dimension_check_and_normalize():
    assert nw_utils.output_dimension_validity_check("m", 384, 384)
    assert not nw_utils.output_dimension_validity_check("m", 384, 48)
    assert nw_utils.output_dimension_validity_check("voyage-3-large", 1024, 128)
    v = nw_utils.normalize_vector(np.array([3.0, 4.0]))
    assert v.dtype == np.float32 and np.allclose(v, [0.6, 0.8])
    with pytest.raises(AssertionError):
        nw_utils.normalize_vector(np.zeros(4))


def test_synth_fixture_style():
    v = synth.vectors(100, 384, 1234)
    assert v.dtype == np.float32 and v.min() >= 0
    assert np.all(np.abs(np.linalg.norm(v.astype(np.float64), axis=1) - 1) < 2e-7)   # test_02_cu_knn.py:28
    # counter-based: any slab can be regenerated
    assert np.array_equal(synth.vectors(10, 384, 1234, row0=50), v[50:60])
    assert not np.array_equal(synth.vectors(10, 384, 1235), v[:10])


def test_tune_memory_matches_reference_search(monkeypatch):
    """cu_knn.py:54-109 restated on arithmetic: replay the reference loop with memory_used = rows * bytes."""
    total = 17179869184                                   # the A5500 of docs/cli.png
    monkeypatch.setattr(cu_knn, "_total_gpu_memory", lambda: total)
    table = pa.table({f"embedding_{i}": pa.array(np.zeros(10, dtype=np.float32)) for i in range(1536)})

    class Big:
        schema = table.schema

        def __len__(self):
            return 4_000_000
    got = cu_knn.tune_memory(Big(), 100000, 0.1)
    # docs/cli.png: 100 000-row frame = 614 400 000 B, factor 6, threshold 480 000
    assert math.ceil(total * .2 / 614400000) == 6
    assert got == 480000
    assert cu_knn.tune_memory(table, 100000, 0.1) == 10   # batch larger than the table -> all rows


def test_stream_to_parquet_schema(tmp_path):
    arr = np.arange(12, dtype=np.float32).reshape(4, 3)
    path = str(tmp_path / "d.parquet")
    cu_knn.stream_cudf_to_parquet(path, arr, 3)
    t = pq.read_table(path)
    assert t.schema.names == ["0", "1", "2"] and t.schema.field(0).type == pa.float32()
    assert pq.ParquetFile(path).num_row_groups == 2
    np.testing.assert_array_equal(np.stack([t.column(j).to_numpy() for j in range(3)], 1), arr)
    cu_knn.stream_cudf_to_parquet(path, torch.arange(6, dtype=torch.int64).reshape(2, 3), 100000)
    assert pq.read_table(path).schema.field(2).type == pa.int64()


def test_prep_table_keeps_only_embedding_columns(tmp_path):
    vec = synth.vectors(20, 8, 1)
    cols = {"text": pa.array(["a"] * 20), "document_id_idx": pa.array(np.arange(20))}
    cols.update({f"embedding_{i}": pa.array(vec[:, i]) for i in range(8)})
    pq.write_table(pa.table(cols), str(tmp_path / "b.parquet"))
    t = cu_knn.prep_table(str(tmp_path), "m", "b.parquet", 15, 8)
    assert t.schema.names == [f"embedding_{i}" for i in range(8)] and len(t) == 15
    with pytest.raises(AssertionError):
        cu_knn.prep_table(str(tmp_path), "m", "b.parquet", 15, 9)
    assert cu_knn.get_embedding_count(t) == 8
    assert cu_knn.drop_columns(pa.table(cols), ["text"]).schema.names == ["text"]
    assert cu_knn.float_column_names(pa.table(cols)) == [f"embedding_{i}" for i in range(8)]


def test_iter_slabs_exact_offsets(tmp_path):
    vec = synth.vectors(1000, 4, 1)
    pq.write_table(pa.table({f"embedding_{i}": pa.array(vec[:, i]) for i in range(4)}),
                   str(tmp_path / "b.parquet"), row_group_size=97)     # ragged record batches
    dset = ds.dataset(str(tmp_path / "b.parquet"), format="parquet")
    cols = [f"embedding_{i}" for i in range(4)]
    seen = []
    for off, rows, batches in cu_knn_ds._iter_slabs(dset, cols, 950, 300):
        stage = np.empty((4, 300), dtype=np.float32)
        cu_knn_ds._fill_pinned(stage, batches, cols)
        np.testing.assert_array_equal(stage[:, :rows].T, vec[off:off + rows])
        seen.append((off, rows))
    assert seen == [(0, 300), (300, 300), (600, 300), (900, 50)]


def test_iter_slabs_rank_shard_and_folded_tail(tmp_path):
    """A rank's shard [start, limit) in slabs; a tail shorter than min_rows joins the slab before it;
    row groups before the shard are skipped."""
    vec = synth.vectors(1000, 4, 1)
    pq.write_table(pa.table({f"embedding_{i}": pa.array(vec[:, i]) for i in range(4)}),
                   str(tmp_path / "b.parquet"), row_group_size=97)
    dset = ds.dataset(str(tmp_path / "b.parquet"), format="parquet")
    cols = [f"embedding_{i}" for i in range(4)]

    def collect(limit, slab, start, min_rows):
        seen = []
        for off, rows, batches in cu_knn_ds._iter_slabs(dset, cols, limit, slab, start=start, min_rows=min_rows):
            stage = np.empty((4, rows), dtype=np.float32)
            cu_knn_ds._fill_pinned(stage, batches, cols)
            np.testing.assert_array_equal(stage.T, vec[off:off + rows])
            seen.append((off, rows))
        return seen
    assert collect(1000, 300, 500, 1) == [(500, 300), (800, 200)]
    assert collect(950, 300, 333, 1) == [(333, 300), (633, 300), (933, 17)]
    assert collect(950, 300, 333, 100) == [(333, 300), (633, 317)]                # 17-row tail folded
    assert collect(420, 300, 400, 100) == [(400, 20)]                             # a whole shard may be short
    assert collect(500, 300, 500, 1) == []
    # every (world, rank) split of a base tiles it exactly
    for world in (2, 3, 8):
        covered = []
        for r in range(world):
            lo, hi = nwdist.shard_bounds(987, world, r)
            covered += collect(hi, 50, lo, 10)
        assert covered[0][0] == 0 and sum(n for _, n in covered) == 987
        assert all(covered[i][0] + covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
        assert all(n >= 10 for _, n in covered)


def test_parquet_rows_reads_only_what_a_slice_needs(tmp_path):
    vec = synth.vectors(1000, 6, 3)
    cols = {"text": pa.array(["a"] * 1000)}
    cols.update({f"embedding_{i}": pa.array(vec[:, i]) for i in range(6)})
    pq.write_table(pa.table(cols), str(tmp_path / "b.parquet"), row_group_size=128)
    names = [f"embedding_{i}" for i in range(6)]
    rows = cu_knn.ParquetRows(str(tmp_path / "b.parquet"), names, 900)
    assert len(rows) == 900 and rows.schema.names == names
    for off, n in [(0, 900), (0, 1), (127, 2), (128, 128), (500, 10 ** 6), (899, 1), (900, 5)]:
        t = rows.slice(off, n)
        want = vec[off:min(900, off + n)]
        assert t.schema.names == names and t.num_rows == len(want)
        got = np.stack([t.column(j).to_numpy() for j in range(6)], axis=1) if len(want) else want
        np.testing.assert_array_equal(got, want)


def test_plan_batch_size():
    """The planner that replaces tune_memory's trial loop: the reference's threshold criterion and the
    engine's real footprint, per GPU."""
    total = 180 << 30
    # 10M x 1536 base, 10k queries: 10 % of 180 GB = 19.3 GB of fp32 rows = 3.1M rows (the threshold binds)
    rows = cu_knn.plan_batch_size(10 ** 7, 1536, 10 ** 4, 100, 0.1, free_bytes=total, total_bytes=total)
    assert rows == int(0.1 * total) // (4 * 1536)
    # a generous threshold: the whole base in one batch when rows + workspace fit
    assert cu_knn.plan_batch_size(10 ** 7, 1536, 10 ** 4, 100, 0.9, free_bytes=total, total_bytes=total) == 10 ** 7
    # little free memory: the footprint binds, and what is planned does fit
    from neighborhoodwatch_b200 import _native
    lib = _native.load()
    free = 8 << 30
    rows = cu_knn.plan_batch_size(10 ** 7, 1536, 10 ** 4, 100, 0.9, free_bytes=free, total_bytes=total)
    need = lambda r: r * 4 * 1536 + lib.nwk_knn_workspace_bytes(r, 10 ** 4, 1536, 100, 0, 0)   # noqa: E731
    assert need(rows) <= 0.9 * free < need(rows + 1000)
    # eight ranks share a batch: eight times the rows
    one = cu_knn.plan_batch_size(10 ** 8, 1024, 10 ** 5, 100, 0.1, free_bytes=total, total_bytes=total)
    eight = cu_knn.plan_batch_size(10 ** 8, 1024, 10 ** 5, 100, 0.1, world=8, free_bytes=total, total_bytes=total)
    assert eight == 8 * one
    # never below k, never above the table
    assert cu_knn.plan_batch_size(50, 8, 5, 10, 0.1, free_bytes=total, total_bytes=total) == 50
    assert cu_knn.plan_batch_size(10 ** 6, 8, 5, 100, 1e-9, free_bytes=total, total_bytes=total) == 100


def test_merge_file_count(tmp_path):
    p = tmp_path / "partial"
    p.mkdir()
    for name in ["indices0.parquet", "indices1.parquet", "final_indices.parquet", "distances0.parquet"]:
        (p / name).write_bytes(b"")
    assert merge.get_file_count(str(p)) == 2 == mo.file_count(str(p))


def test_fvec_ivec_writer_is_byte_identical_to_the_reference(tmp_path):
    """Golden bytes come from the reference's own per-row writer (tests/golden/make_golden.py)."""
    import pandas as pd

    from neighborhoodwatch_b200 import parquet_to_format as fmt
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fvec_ivec_writer.npz"))
    d = str(tmp_path)
    fmt.write_ivec_fvec_from_dataframe(d, "m", "v.fvec", pd.DataFrame(g["vectors"]), "f", 12)
    fmt.write_ivec_fvec_from_dataframe(d, "m", "i.ivec", g["indices"], "i", 10)              # arrays work too
    fmt.write_ivec_fvec_from_dataframe(d, "m", "d.fvec", pd.DataFrame(g["distances"]), "f", 10)
    with_rownum = pd.DataFrame(g["indices"])
    with_rownum["RowNum"] = np.arange(len(with_rownum))
    fmt.write_ivec_fvec_from_dataframe(d, "m", "r.ivec", with_rownum, "i", 10)
    for fn in ("v.fvec", "i.ivec", "d.fvec", "r.ivec"):
        got = np.frombuffer(open(os.path.join(d, fn), "rb").read(), dtype=np.uint8)
        np.testing.assert_array_equal(got, g[fn.replace(".", "_")], err_msg=fn)
    assert fmt.count_vectors(d, "v.fvec") == int(g["count_v"]) == 23
    np.testing.assert_array_equal(np.array(fmt.get_nth_vector(d, "v.fvec", 5), dtype=np.float32), g["nth_v_5"])
    np.testing.assert_array_equal(np.array(fmt.get_nth_vector(d, "i.ivec", 7), dtype=np.int64), g["nth_i_7"])
    np.testing.assert_array_equal(fmt.read_vec_file(d, "d.fvec"), g["distances"])
    with pytest.raises(AssertionError):
        fmt.write_ivec_fvec_from_dataframe(d, "m", "x.fvec", g["vectors"], "f", 9)


def test_shard_bounds_cover_exactly():
    for total, world in [(10, 3), (10 ** 7, 8), (5, 8), (0, 2)]:
        spans = [nwdist.shard_bounds(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nb, nq, d, k, metric, out_dir):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = nwdist.shard_bounds(nb, world, rank)
        base = synth.vectors(hi - lo, d, 1234, row0=lo)
        q = synth.vectors(nq, d, 5678)

        def local_knn(b, qq, kk, m, off):           # CPU stand-in for the kernel (test only)
            dd, ii = ko.knn_f64(b.numpy(), qq.numpy(), kk, m)
            return torch.from_numpy(dd), torch.from_numpy(ii + off)

        def merge_fn(all_d, all_i, kk):
            md, mi = mo.merge_by_distance_then_index([(all_d[p].numpy(), all_i[p].numpy()) for p in range(all_d.shape[0])], kk)
            return torch.from_numpy(md), torch.from_numpy(mi)

        out_d, out_i = nwdist.sharded_knn(torch.from_numpy(base), torch.from_numpy(q), k, metric, lo,
                                          local_knn=local_knn, merge=merge_fn)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), d=out_d.numpy(), i=out_i.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nb,k", [(1001, 10), (15, 10)])
def test_sharded_knn_gloo_world2(tmp_path, nb, k):
    """N>1 path: shard, local top-k with global ids, all-gather, merge -- identical on all ranks
    and equal to the single-device answer.  (15 rows over 2 ranks: shards smaller than k.)"""
    nq, d, metric = 23, 16, "sqeuclidean"
    port = _free_port()
    mp.spawn(_worker, args=(2, port, nb, nq, d, k, metric, str(tmp_path)), nprocs=2, join=True)
    base, q = synth.vectors(nb, d, 1234), synth.vectors(nq, d, 5678)
    t_d, t_i = ko.knn_f64(base, q, k, metric)
    for r in range(2):
        g = np.load(tmp_path / f"r{r}.npz")
        np.testing.assert_array_equal(g["i"], t_i)
        np.testing.assert_array_equal(g["d"], t_d)


def _maxsim_worker(rank, world, port, nd, nq, k, out_dir):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import maxsim_oracle as mso
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        qt = rng.standard_normal((nq, 8, 16)).astype(np.float32)
        dt = rng.standard_normal((nd, 12, 16)).astype(np.float32)
        lo, hi = nwdist.shard_bounds(nd, world, rank)

        def local_maxsim(q, dd, kk, dl, off):       # CPU stand-in for the kernel (test only)
            ss, ii = mso.maxsim_topk_f64(q.numpy(), dd.numpy(), kk)
            return torch.from_numpy(ss.astype(np.float32)), torch.from_numpy(ii + off)

        def merge_fn(all_d, all_i, kk):
            md, mi = mo.merge_by_distance_then_index([(all_d[p].numpy(), all_i[p].numpy()) for p in range(all_d.shape[0])], kk)
            return torch.from_numpy(md), torch.from_numpy(mi)

        out_s, out_i = nwdist.sharded_maxsim(torch.from_numpy(qt), torch.from_numpy(dt[lo:hi]), k, lo,
                                             local_maxsim=local_maxsim, merge=merge_fn)
        np.savez(os.path.join(out_dir, f"m{rank}.npz"), s=out_s.numpy(), i=out_i.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nd,k", [(301, 10), (13, 10)])
def test_sharded_maxsim_gloo_world2(tmp_path, nd, k):
    """Documents sharded over two ranks: local top-k with global ids, all-gather, merge of the negated
    scores -- identical on both ranks and equal to the single-device answer."""
    from oracle import maxsim_oracle as mso
    nq = 9
    port = _free_port()
    mp.spawn(_maxsim_worker, args=(2, port, nd, nq, k, str(tmp_path)), nprocs=2, join=True)
    rng = np.random.default_rng(5)
    qt = rng.standard_normal((nq, 8, 16)).astype(np.float32)
    dt = rng.standard_normal((nd, 12, 16)).astype(np.float32)
    t_s, t_i = mso.maxsim_topk_f64(qt, dt, k)
    for r in range(2):
        g = np.load(tmp_path / f"m{r}.npz")
        np.testing.assert_array_equal(g["i"], t_i)
        np.testing.assert_allclose(g["s"], t_s, rtol=1e-6)
