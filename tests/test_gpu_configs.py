"""Parity at the arithmetic of BASELINE configs 3 and 4 (long rows, squared L2) and of the real
multi-GPU chain (local kernels + NCCL exchange + merge kernel).  -m gpu.

Config 3 = 10 000 queries x 10^7 base x d=1536, squared L2, base sharded over 8 GPUs: one GPU's
share is 1 250 000 rows.  Config 4 = 10^5 x 10^8 x d=1024, squared L2, streamed through
``compute_knn_ds`` in slabs.  Truth is ``oracle.knn_f64`` (float64 numpy) on sampled queries against
the whole base of the test, plus the on-device fp64 scan on a larger sample.
"""
import os
import socket
import sys

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from neighborhoodwatch_b200 import synth
from oracle import knn_oracle as ko

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dev(arr, device):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr)).to(device)


def _check(got_d, got_i, base, q, k, metric, id_offset=0):
    t_d, t_i = ko.knn_f64(base, q, k, metric, return_f64=True)
    return ko.compare_knn(np.asarray(got_d), np.asarray(got_i), base, q, metric, t_d, t_i + id_offset,
                          rtol=1e-5, atol=1e-6, tie_rtol=1e-6, id_offset=id_offset)


@pytest.mark.parametrize("metric", ko.METRICS)
@pytest.mark.parametrize("d", [1536, 1100, 1028, 2048])
def test_long_rows_against_the_oracle(cuda_device, metric, d):
    """Rows longer than 1024 elements leave baseprep's register cache (ROW_CACHE = 8 float4 per lane)
    and re-read their tail: norms, alpha / beta, the fp16 copy and the neighbours of every query are
    checked against the full oracle.  d = 1536 is config 3's; 1100 / 1028 end inside a 128-element
    stripe of the tail; 2048 is two full cache lengths.  Rows are NOT unit length (scaled per row by
    0.25 .. 4) so that a wrong |b|^2 of the tail changes the ranking."""
    from neighborhoodwatch_b200 import knn
    nq, nb, k = 257, 20_000, 100
    rng = np.random.default_rng(d)
    base = synth.vectors(nb, d, 1234) * rng.uniform(0.25, 4.0, size=(nb, 1)).astype(np.float32)
    # signed components in the tail only matter if the tail is read: flip signs there
    base[:, 1024:] *= rng.choice(np.array([-1.0, 1.0], dtype=np.float32), size=(nb, d - 1024))
    q = synth.vectors(nq, d, 5678)
    got_d, got_i = knn.knn(_dev(base, cuda_device), _dev(q, cuda_device), k, metric=metric, id_offset=11)
    st = dict(knn.last_stats)
    assert st["path"] == 1 and st["max_err_ratio"] <= 1.0, st
    res = _check(got_d.cpu().numpy(), got_i.cpu().numpy(), base, q, k, metric, id_offset=11)
    print(metric, d, res, st)


def test_config3_shard_full_size(cuda_device):
    """One GPU's share of BASELINE config 3 at full size: 10 000 queries x 1 250 000 rows x d=1536,
    squared L2, id_offset of rank 3.  64 sampled queries against oracle.knn_f64 over the whole shard,
    128 against the on-device fp64 scan (bit for bit), every row through size-independent properties."""
    import torch
    from neighborhoodwatch_b200 import knn
    nq, nb, d, k = 10_000, 1_250_000, 1536, 100
    off = 3 * nb
    bd = synth.device_vectors(nb, d, synth.BASE_SEED, row0=off, device=cuda_device)
    qd = synth.device_vectors(nq, d, synth.QUERY_SEED, device=cuda_device)
    got_d, got_i = knn.knn(bd, qd, k, metric="sqeuclidean", id_offset=off)
    st = dict(knn.last_stats)
    assert st["path"] == 1 and st["max_err_ratio"] <= 1.0, st
    assert st["n_fallback_queries"] <= 10, st
    assert torch.all(got_d[:, 1:] >= got_d[:, :-1])
    assert int(got_i.min()) >= off and int(got_i.max()) < off + nb
    srt = torch.sort(got_i, dim=1).values
    assert torch.all(srt[:, 1:] != srt[:, :-1])
    sel = torch.arange(0, nq, 78, device=cuda_device)[:128]
    ex_d, ex_i = knn.knn(bd, qd[sel], k, metric="sqeuclidean", id_offset=off, force_exact_scan=True)
    assert torch.equal(ex_i, got_i[sel]) and torch.equal(ex_d, got_d[sel])
    # oracle truth for 64 of them (the device generator is bit-identical to synth.vectors:
    # test_synth_bit_identical, so the shard is simply copied back)
    sel64 = sel[::2].cpu().numpy()
    base = bd.cpu().numpy()
    q = qd.cpu().numpy()[sel64]
    res = _check(got_d.cpu().numpy()[sel64], got_i.cpu().numpy()[sel64], base, q, k, "sqeuclidean", id_offset=off)
    # the reference's own acceptance rule (parquet_to_format.py:373) on every returned pair
    from neighborhoodwatch_b200 import parquet_to_format as fmt
    assert fmt.validate_pairs(bd, qd, got_i - off, got_d) == 0
    print(res, st)


def _write_base_parquet(path, rows_total, d, chunk, device, seed):
    """One column per dimension (generate_dataset.py:233-235), written chunk by chunk; returns the
    host copy of the vectors."""
    names = [f"embedding_{i}" for i in range(d)]
    schema = pa.schema([pa.field(n, pa.float32()) for n in names])
    host = np.empty((rows_total, d), dtype=np.float32)
    with pq.ParquetWriter(path, schema, compression="none") as w:
        for lo in range(0, rows_total, chunk):
            n = min(chunk, rows_total - lo)
            dev = synth.device_vectors(n, d, seed, row0=lo, device=device)
            host[lo:lo + n] = dev.cpu().numpy()
            cm = dev.t().contiguous().cpu().numpy()              # [d, n]: one contiguous array per column
            w.write_table(pa.Table.from_arrays([pa.array(cm[j]) for j in range(d)], schema=schema))
    return host


def test_config4_arithmetic_through_compute_knn_ds(tmp_path, cuda_device):
    """BASELINE config 4's arithmetic through the streaming entry point: d=1024, squared L2, the base
    (1 000 000 rows) arrives in 4 slabs of 250 000 rows from a parquet dataset through the pinned
    double buffer; every slab is a running search bounded by the state's k-th distances.  The final
    files are compared with oracle.knn_f64 over the whole base on 64 sampled queries and with the
    one-shot device result on all of them."""
    import torch
    from neighborhoodwatch_b200 import cu_knn_ds, knn
    from neighborhoodwatch_b200.nw_utils import get_partial_distances_filename, get_partial_indices_filename
    data_dir = str(tmp_path)
    nq, nb, d, k, slab = 2048, 1_000_000, 1024, 100, 250_000
    base = _write_base_parquet(os.path.join(data_dir, "base.parquet"), nb, d, 125_000, cuda_device, synth.BASE_SEED)
    q = _write_base_parquet(os.path.join(data_dir, "query.parquet"), nq, d, nq, cuda_device, synth.QUERY_SEED)
    cu_knn_ds.compute_knn_ds(data_dir, d, "query.parquet", nq, "base.parquet", nb, k=k, initial_batch_size=slab)

    def read(path, dtype):
        t = pq.read_table(path)
        assert t.schema.names == [str(i) for i in range(k)]
        return np.stack([t.column(j).to_numpy() for j in range(k)], axis=1).astype(dtype, copy=False)
    f_d = read(get_partial_distances_filename(data_dir, -1), np.float32)
    f_i = read(get_partial_indices_filename(data_dir, -1), np.int64)
    assert f_d.shape == (nq, k) and np.all(np.diff(f_d, axis=1) >= 0)
    sel = np.arange(0, nq, 32)                                              # 64 sampled queries
    res = _check(f_d[sel], f_i[sel], base, q[sel], k, "sqeuclidean")
    # all queries: the streamed result equals the one-shot result on the resident base (distances
    # exactly; ids as sets per row -- the running merge orders by (fp32 distance, id))
    one_d, one_i = knn.knn(_dev(base, cuda_device), _dev(q, cuda_device), k, metric="sqeuclidean")
    assert np.array_equal(one_d.cpu().numpy(), f_d)
    assert np.array_equal(np.sort(one_i.cpu().numpy(), axis=1), np.sort(f_i, axis=1))
    print(res)


# ---------------------------------------------------------------------------------------------
# real multi-GPU chain: libnwk kernels + NCCL exchange + merge kernel, two ranks
# ---------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _nccl_worker(rank, world, port, nb, nq, d, k, metric, out_dir, exchange):
    sys.path.insert(0, ROOT)
    if exchange == "allgather":
        os.environ["NWK_EXCHANGE"] = "allgather"
    import torch
    import torch.distributed as dist
    from neighborhoodwatch_b200 import dist as nwdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        lo, hi = nwdist.shard_bounds(nb, world, rank)
        base = synth.device_vectors(hi - lo, d, 1234, row0=lo, device=dev)
        q = synth.device_vectors(nq, d, 5678, device=dev)
        for it in range(3):                       # repeated steps reuse the exchange buffers
            out_d, out_i = nwdist.sharded_knn(base, q, k, metric, lo)
        torch.cuda.synchronize(dev)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), d=out_d.cpu().numpy(), i=out_i.cpu().numpy(),
                 exchange=np.array(nwdist.last_exchange or "none"), why=np.array(nwdist._peer_failed or ""))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["peer", "allgather"])
@pytest.mark.parametrize("metric,nb,d,k", [("sqeuclidean", 300_001, 1536, 100), ("cosine", 150, 64, 100)])
def test_sharded_knn_nccl_world2(tmp_path, metric, nb, d, k, exchange):
    """dist.sharded_knn on two GPUs: shard, local tcgen05 top-k with global ids, then either the
    peer-memory gather+merge kernel (lists read from the owners' HBM over NVLink, nwk_xchg_*) or one
    packed NCCL all-gather + the merge kernel -- identical on both ranks and equal to the oracle.
    (150 rows over 2 ranks: shards smaller than k are padded.)"""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    nq = 500
    port = _free_port()
    mp.spawn(_nccl_worker, args=(2, port, nb, nq, d, k, metric, str(tmp_path), exchange), nprocs=2, join=True)
    base, q = synth.vectors(nb, d, 1234), synth.vectors(nq, d, 5678)
    g0 = np.load(tmp_path / "r0.npz")
    g1 = np.load(tmp_path / "r1.npz")
    assert np.array_equal(g0["i"], g1["i"]) and np.array_equal(g0["d"], g1["d"])
    sel = np.arange(0, nq, 5)
    res = _check(g0["d"][sel], g0["i"][sel], base, q[sel], k, metric)
    print(res, str(g0["exchange"]), str(g0["why"]))
    want = "peer-memory gather+merge kernel" if exchange == "peer" else "packed-all-gather"
    assert str(g0["exchange"]) == want == str(g1["exchange"]), str(g0["why"])


def _write_vectors(path, vecs, row_group_size=257):
    cols = {"document_id_idx": pa.array(np.arange(len(vecs), dtype=np.int64)), "text": pa.array(["t"] * len(vecs))}
    for j in range(vecs.shape[1]):
        cols[f"embedding_{j}"] = pa.array(vecs[:, j])
    pq.write_table(pa.table(cols), path, row_group_size=row_group_size)


def _entry_worker(rank, world, port, data_dir, nq, nb, d, k):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from neighborhoodwatch_b200 import cu_knn, cu_knn_ds, merge
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        # table API: 4 base batches, each row-sharded over the two ranks; rank 0 writes the partial files
        cu_knn.compute_knn(os.path.join(data_dir, "a"), "intfloat/e5-small-v2", d, "query.parquet", nq, "base.parquet", nb,
                           k=k, initial_batch_size=1000, split=True)
        if rank == 0:
            merge.merge_indices_and_distances(os.path.join(data_dir, "a"), k)
        # streaming API: every rank streams its own half of the base in slabs of 400 rows
        cu_knn_ds.compute_knn_ds(os.path.join(data_dir, "b"), d, "query.parquet", nq, "base.parquet", nb, k=k,
                                 initial_batch_size=400)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_entry_points_are_rank_aware_world2(tmp_path):
    """compute_knn / compute_knn_ds under a two-rank NCCL process group write the files a single-GPU
    run writes: partial files per base batch (then merged) and the streamed final files, both equal to
    the oracle over the whole base."""
    import torch
    import torch.multiprocessing as mp
    from neighborhoodwatch_b200.nw_utils import get_partial_distances_filename, get_partial_indices_filename
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    nq, nb, d, k = 150, 3010, 96, 10
    base, q = synth.vectors(nb, d, 1234), synth.vectors(nq, d, 5678)
    for sub in ("a", "b"):
        os.makedirs(tmp_path / sub)
        _write_vectors(str(tmp_path / sub / "base.parquet"), base)
        _write_vectors(str(tmp_path / sub / "query.parquet"), q)
    mp.spawn(_entry_worker, args=(2, _free_port(), str(tmp_path), nq, nb, d, k), nprocs=2, join=True)
    t_d, t_i = ko.knn_f64(base, q, k, "sqeuclidean")

    def read(path, dtype):
        t = pq.read_table(path)
        return np.stack([t.column(j).to_numpy() for j in range(t.num_columns)], axis=1).astype(dtype, copy=False)
    partials = ko.process_batches_oracle(base, q, 4, 1000, k, True, engine="raft")
    for i in range(4):
        np.testing.assert_array_equal(read(get_partial_indices_filename(str(tmp_path / "a"), i), np.int64), partials[i][1])
    for sub in ("a", "b"):
        f_d = read(get_partial_distances_filename(str(tmp_path / sub), -1), np.float32)
        f_i = read(get_partial_indices_filename(str(tmp_path / sub), -1), np.int64)
        np.testing.assert_array_equal(f_i, t_i, err_msg=sub)
        np.testing.assert_allclose(f_d, t_d, rtol=1e-5, atol=1e-6, err_msg=sub)


def test_single_process_multi_gpu(cuda_device):
    """nwk_mg_*: one process, two devices, shards of unequal size; device 0 merges the peers' lists in place over
    NVLink.  Equal to the oracle over the whole base, twice (the handle's buffers are reused)."""
    import torch
    from neighborhoodwatch_b200 import knn
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    nq, nb, d, k = 300, 120_001, 256, 100
    base, q = synth.vectors(nb, d, 1234), synth.vectors(nq, d, 5678)
    cut = 70_000
    shards = [torch.from_numpy(base[:cut]).to("cuda:0"), torch.from_numpy(base[cut:]).to("cuda:1")]
    mg = knn.MultiGpu([0, 1])
    for metric in ("sqeuclidean", "cosine"):
        got_d, got_i = mg.knn(shards, q, k, metric=metric, id_offset=7)
        _check(got_d, got_i, base, q, k, metric, id_offset=7)
    mg.close()
    with pytest.raises(Exception):
        knn.MultiGpu([0, 1]).knn([shards[0][:50], shards[1]], q, k)      # a shard shorter than k
