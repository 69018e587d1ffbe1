"""CPU tests of the oracle itself: internal consistency, the reference's own relations and
the golden fixtures produced by running the reference's code (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

from neighborhoodwatch_b200 import synth
from oracle import knn_oracle as ko
from oracle import merge_oracle as mo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def brute(dataset, queries, k, metric):
    dist = ko.distances_f64(dataset, queries, metric)
    out_d = np.empty((len(queries), k))
    out_i = np.empty((len(queries), k), dtype=np.int64)
    for r in range(len(queries)):
        order = np.lexsort((np.arange(dist.shape[1]), dist[r]))[:k]
        out_d[r], out_i[r] = dist[r, order], order
    return out_d, out_i


@pytest.mark.parametrize("metric", ko.METRICS)
@pytest.mark.parametrize("nq,nb,d,k", [(17, 301, 24, 10), (5, 100, 7, 100), (33, 1000, 384, 100)])
def test_blocked_oracle_equals_direct(metric, nq, nb, d, k):
    base = synth.vectors(nb, d, 11)
    q = synth.vectors(nq, d, 22)
    d64, i64 = ko.knn_f64(base, q, k, metric, block=128, return_f64=True)
    bd, bi = brute(base, q, k, metric)
    np.testing.assert_array_equal(i64, bi)
    np.testing.assert_allclose(d64, bd, rtol=0, atol=1e-12)


def test_ties_resolved_by_index():
    base = np.tile(synth.vectors(4, 16, 3), (50, 1))           # every row appears 50 times
    q = synth.vectors(3, 16, 4)
    _, idx = ko.knn_f64(base, q, 60, "sqeuclidean", block=64)
    for r in range(3):
        first = idx[r, :50]
        assert np.all(np.diff(first) > 0) and len(set(first % 4)) == 1


@pytest.mark.parametrize("metric", ko.METRICS)
def test_reference_form_fp32_agrees_within_tolerance(metric):
    base = synth.vectors(5000, 384, 1234)
    q = synth.vectors(64, 384, 5678)
    t_d, t_i = ko.knn_f64(base, q, 100, metric, return_f64=True)
    r_d, r_i = ko.knn_f32_reference_form(base, q, 100, metric)
    res = ko.compare_knn(r_d, r_i, base, q, metric, t_d, t_i, rtol=1e-5, atol=2e-6, tie_rtol=1e-5)
    assert res["slots"] == 6400


def test_validator_relation_unit_vectors():
    """parquet_to_format.py:373 -- on unit vectors squared-L2 / 2 == 1 - dot."""
    base = synth.vectors(2000, 128, 5)
    q = synth.vectors(20, 128, 6)
    d_l2, i_l2 = ko.knn_f64(base, q, 10, "sqeuclidean", return_f64=True)
    for r in range(len(q)):
        dots = base[i_l2[r]].astype(np.float64) @ q[r].astype(np.float64)
        assert np.all(np.isclose(1 - dots, d_l2[r] / 2, rtol=1e-5, atol=1e-6))


@pytest.mark.parametrize("engine", ["raft", "cuvs", "torch"])
def test_batches_plus_merge_equals_one_shot(engine):
    base = synth.vectors(1000, 32, 7)
    q = synth.vectors(101, 32, 8)
    k, batch_size = 10, 300
    bc = ko.batch_plan(len(base), batch_size, k)
    partials = ko.process_batches_oracle(base, q, bc, batch_size, k, split=True, engine=engine)
    assert len(partials) == 4 and partials[0][0].shape == (101, 10)
    m_d, m_i = mo.merge_by_distance_then_index(partials, k)
    p_d, p_i = mo.merge_pairwise(partials, k, stable=True)
    t_d, t_i = ko.knn_f64(base, q, k, ko.ENGINE_METRIC[engine])
    np.testing.assert_array_equal(m_i, t_i)
    np.testing.assert_array_equal(p_i, t_i)
    np.testing.assert_array_equal(m_d, t_d)


def test_batch_plan_assertion():
    with pytest.raises(AssertionError):
        ko.batch_plan(1005, 100, 10)       # last batch would hold 5 < k rows (cu_knn.py:177)
    assert ko.batch_plan(1000, 100, 10) == 10


def test_file_count(tmp_path):
    for name in ["indices0.parquet", "indices1.parquet", "indices12.parquet", "final_indices.parquet",
                 "indicesX.parquet", "distances0.parquet"]:
        (tmp_path / name).write_bytes(b"")
    assert mo.file_count(str(tmp_path)) == 3


golden_files = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))


@pytest.mark.parametrize("path", golden_files, ids=[os.path.basename(p) for p in golden_files])
def test_golden_fixture(path):
    """Fixtures were produced by the reference's own code (see tests/golden/make_golden.py)."""
    g = np.load(path, allow_pickle=False)
    kind = str(g["kind"])
    if kind == "process_batches":
        base = synth.vectors(int(g["nb"]), int(g["d"]), int(g["base_seed"]))
        q = synth.vectors(int(g["nq"]), int(g["d"]), int(g["query_seed"]))
        k = int(g["k"])
        partials = ko.process_batches_oracle(base, q, int(g["batch_count"]), int(g["batch_size"]), k,
                                             bool(g["split"]), engine=str(g["engine"]))
        metric = ko.ENGINE_METRIC[str(g["engine"])]
        for i, (d_i, i_i) in enumerate(partials):
            ref_d, ref_i = g[f"distances{i}"], g[f"indices{i}"]
            assert ref_i.dtype == np.int64 and ref_d.dtype == np.float32
            # the reference computed in fp32: compare it against the float64 truth of the same batch
            lo = i * int(g["batch_size"])
            hi = min(lo + int(g["batch_size"]), len(base))
            t_d, t_i = ko.knn_f64(base[lo:hi], q, k, metric, return_f64=True)
            ko.compare_knn(ref_d, ref_i, base[lo:hi], q, metric, t_d, t_i + lo, rtol=1e-5, atol=2e-6,
                           tie_rtol=1e-5, id_offset=lo)
            np.testing.assert_array_equal(i_i, t_i + lo)
    elif kind == "merge":
        P = int(g["P"])
        partials = [(g[f"distances{i}"], g[f"indices{i}"]) for i in range(P)]
        k = int(g["k"])
        o_d, o_i = mo.merge_pairwise(partials, k)
        np.testing.assert_array_equal(o_d, g["final_distances"])
        if bool(g["distinct"]):
            np.testing.assert_array_equal(o_i, g["final_indices"])
            m_d, m_i = mo.merge_by_distance_then_index(partials, k)
            np.testing.assert_array_equal(m_i, g["final_indices"])
            np.testing.assert_array_equal(m_d, g["final_distances"])
    elif kind == "fvec_ivec":
        return      # checked against the product writer in tests/test_host_logic.py
    else:
        raise AssertionError(f"unknown golden kind {kind}")
