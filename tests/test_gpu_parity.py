"""Parity of the CUDA path against the oracle, through the C ABI (libnwk.so).  -m gpu."""
import os

import numpy as np
import pytest

from neighborhoodwatch_b200 import synth
from oracle import knn_oracle as ko
from oracle import merge_oracle as mo

pytestmark = pytest.mark.gpu

ENGINES = {"sqeuclidean": "raft", "cosine": "cuvs", "dot": "torch"}


def _dev(arr, device):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr)).to(device)


def _check(got_d, got_i, base, q, k, metric, id_offset=0):
    t_d, t_i = ko.knn_f64(base, q, k, metric, return_f64=True)
    # north_star tolerances: distances 1e-5 relative (abs floor 1e-6), ids exact except ties < 1e-6 rel
    return ko.compare_knn(got_d.cpu().numpy(), got_i.cpu().numpy(), base, q, metric, t_d, t_i + id_offset,
                          rtol=1e-5, atol=1e-6, tie_rtol=1e-6, id_offset=id_offset)


def test_synth_bit_identical(cuda_device):
    for rows, d, seed, row0 in [(257, 384, 1234, 0), (100, 100, 5678, 12345), (3, 1536, 9, 10 ** 8)]:
        dev = synth.device_vectors(rows, d, seed, row0, device=cuda_device).cpu().numpy()
        ref = synth.vectors(rows, d, seed, row0)
        assert np.array_equal(dev.view(np.uint32), ref.view(np.uint32))
    raw = synth.device_vectors(10, 16, 3, normalize=False, device=cuda_device).cpu().numpy()
    assert np.array_equal(raw, synth.vectors(10, 16, 3, normalize=False))


@pytest.mark.parametrize("metric", ko.METRICS)
@pytest.mark.parametrize("nq,nb,d,k", [(37, 1000, 96, 10), (8, 4097, 33, 100), (1, 100, 7, 100), (130, 9000, 128, 100)])
def test_exact_scan_parity(cuda_device, metric, nq, nb, d, k):
    from neighborhoodwatch_b200 import knn
    base = synth.vectors(nb, d, 1234)
    q = synth.vectors(nq, d, 5678)
    got_d, got_i = knn.knn(_dev(base, cuda_device), _dev(q, cuda_device), k, metric=metric, id_offset=77,
                           force_exact_scan=True)
    assert knn.last_stats["path"] == 0
    res = _check(got_d, got_i, base, q, k, metric, id_offset=77)
    assert res["index_mismatches_tied"] == 0      # the fp64 scan has the oracle's arithmetic


@pytest.mark.parametrize("f16", [True, False], ids=["fp16", "tf32"])
def test_tc_tile_matches_operand_model(cuda_device, f16):
    """Validates TMA map, UMMA descriptors and TMEM read-back in isolation, and the error
    constant the certification relies on."""
    from neighborhoodwatch_b200 import knn
    rng = np.random.default_rng(0)
    worst = 0.0
    for d in (32, 64, 384, 100, 1536):
        base = (rng.standard_normal((600, d)) * rng.uniform(0.01, 100.0, (600, 1))).astype(np.float32)
        q = rng.standard_normal((300, d)).astype(np.float32)
        out = knn.debug_tc_tile(_dev(base, cuda_device), _dev(q, cuda_device), 256, 128, f16=f16).cpu().numpy()
        exact = q[128:256].astype(np.float64) @ base[256:512].astype(np.float64).T
        scale = np.abs(q[128:256].astype(np.float64)) @ np.abs(base[256:512].astype(np.float64)).T
        err = np.abs(out - exact)
        if f16:
            c = 2.0 ** -10 * 1.001 + (d / 16 + 2) / 2.0 ** 22
        else:
            c = 2.0 ** -9 * 1.001 + (d / 8 + 2) / 2.0 ** 22
        ratio = (err / (c * scale + 1e-30)).max()
        worst = max(worst, ratio)
        assert ratio <= 1.0, f"d={d}: max err ratio {ratio}"
        # rows beyond the inputs read as zero (TMA out-of-bounds fill)
        out2 = knn.debug_tc_tile(_dev(base, cuda_device), _dev(q, cuda_device), 512, 256, f16=f16).cpu().numpy()
        assert np.all(out2[300 - 256:] == 0) and np.all(out2[:, 600 - 512:] == 0)
    print("worst error / bound:", worst)


def test_tc_operand_rounding_probe(cuda_device):
    """Documents how the tensor core treats fp32 operands under kind::tf32 (truncate vs round)."""
    from neighborhoodwatch_b200 import knn
    base = np.zeros((256, 32), dtype=np.float32)
    q = np.zeros((128, 32), dtype=np.float32)
    base[:, 0] = 1.0
    vals = [1 + 2.0 ** -11 + 2.0 ** -12, 1 + 2.0 ** -11, 1 + 2.0 ** -10 + 2.0 ** -11, 1 + 2.0 ** -12]
    for i, v in enumerate(vals):
        q[i, 0] = v
    out = knn.debug_tc_tile(_dev(base, cuda_device), _dev(q, cuda_device), 0, 0, f16=False).cpu().numpy()
    got = [float(out[i, 0]) for i in range(len(vals))]
    trunc = [1.0, 1.0, 1 + 2.0 ** -10, 1.0]
    rne = [1 + 2.0 ** -10, 1.0, 1 + 2.0 ** -9, 1.0]
    print("tf32 operand probe:", got, "truncation" if got == trunc else ("round-nearest-even" if got == rne else "other"))
    assert all(abs(g - v) <= 2.0 ** -10 * 1.001 for g, v in zip(got, vals))


@pytest.mark.parametrize("metric", ko.METRICS)
@pytest.mark.parametrize("nq,nb,d,k", [
    (1000, 10000, 384, 100),     # BASELINE config 1
    (333, 5003, 100, 100),       # ragged everything, d not a multiple of 32
    (5, 150, 64, 100),           # tiny base: everything is logged, no sampling
    (129, 2049, 768, 10),        # one row past a tile boundary on both sides
    (64, 40000, 128, 256),       # k at the tensor-core path's limit
    (50, 3000, 4, 10),           # the shortest rows the tensor-core path takes: one K-block, one MMA step
    (70, 9000, 12, 20),          # (squared L2's extra columns share that block)
    (1, 300, 8, 10),             # a single query against less than two tiles
    (257, 70000, 256, 100),      # three query tiles, four K-blocks (+ one for squared L2's columns): one resident buffer
])
def test_tc_knn_parity(cuda_device, metric, nq, nb, d, k):
    from neighborhoodwatch_b200 import knn
    base = synth.vectors(nb, d, 1234)
    q = synth.vectors(nq, d, 5678)
    got_d, got_i = knn.knn(_dev(base, cuda_device), _dev(q, cuda_device), k, metric=metric, id_offset=5)
    st = dict(knn.last_stats)
    assert st["path"] == 1, st
    res = _check(got_d, got_i, base, q, k, metric, id_offset=5)
    assert st["n_fallback_queries"] <= max(1, nq // 100), st
    assert 0.0 < st["max_err_ratio"] <= 1.0, st
    print(metric, (nq, nb, d, k), st, res)


@pytest.mark.parametrize("direct", ["0", "1"], ids=["alpha-beta-keys", "direct-keys"])
@pytest.mark.parametrize("pair,ares", [("0", "0"), ("0", "1"), ("1", "0")], ids=["single-cta", "single-cta-resident-queries", "cta-pair"])
@pytest.mark.parametrize("nq,nb,d,k,metric", [
    (333, 5003, 100, 100, "sqeuclidean"),    # odd number of query tiles: the last pair has an empty half
    (100, 3000, 64, 50, "cosine"),           # one query tile: the second CTA of the pair has no queries at all
    (1000, 10000, 384, 100, "dot"),           # six K-blocks: one resident query buffer, three base stages
    (700, 300000, 64, 100, "cosine"),        # sample large enough for the warm-up list and several parts per query tile
    (600, 40000, 256, 100, "sqeuclidean"),   # four K-blocks: one resident query buffer, four base stages
])
def test_tc_filter_variants(cuda_device, monkeypatch, pair, ares, direct, nq, nb, d, k, metric):
    """The CTA-pair (cta_group::2) and single-CTA filter kernels, with per-column alpha / beta keys
    or direct keys (one global base scale; squared L2 with |b|^2 in extra K columns), must all reproduce the oracle,
    whatever the planner would have picked for the shape (NWK_PAIR / NWK_DIRECT are the library's
    experiment switches)."""
    from neighborhoodwatch_b200 import knn
    monkeypatch.setenv("NWK_PAIR", pair)
    monkeypatch.setenv("NWK_DIRECT", direct)
    monkeypatch.setenv("NWK_ARES", ares)
    base = synth.vectors(nb, d, 4321)
    q = synth.vectors(nq, d, 8765)
    got_d, got_i = knn.knn(_dev(base, cuda_device), _dev(q, cuda_device), k, metric=metric, id_offset=11)
    st = dict(knn.last_stats)
    assert st["path"] == 1, st
    assert st["n_fallback_queries"] <= max(1, nq // 100), st
    assert st["max_candidates"] <= 4096, st
    if nb > 100000:     # too large for the numpy oracle on every query: check a sample of them
        sel = np.arange(0, nq, 7)
        t_d, t_i = ko.knn_f64(base, q[sel], k, metric, return_f64=True)
        res = ko.compare_knn(got_d.cpu().numpy()[sel], got_i.cpu().numpy()[sel], base, q[sel], metric, t_d, t_i + 11,
                             rtol=1e-5, atol=1e-6, tie_rtol=1e-6, id_offset=11)
    else:
        res = _check(got_d, got_i, base, q, k, metric, id_offset=11)
    print(pair, (nq, nb, d, k, metric), st, res)


@pytest.mark.parametrize("metric", ko.METRICS)
def test_tc_knn_parity_tf32_operands(cuda_device, metric):
    from neighborhoodwatch_b200 import knn
    base = synth.vectors(20000, 256, 1234)
    q = synth.vectors(300, 256, 5678)
    got_d, got_i = knn.knn(_dev(base, cuda_device), _dev(q, cuda_device), 100, metric=metric, tf32=True)
    st = dict(knn.last_stats)
    assert st["path"] == 2 and st["max_err_ratio"] <= 1.0, st
    _check(got_d, got_i, base, q, 100, metric)
    print(metric, st)


def test_tc_gaussian_unnormalised(cuda_device):
    """Non-unit, signed vectors: norms matter for sq-L2 / cosine ordering."""
    from neighborhoodwatch_b200 import knn
    rng = np.random.default_rng(5)
    base = (rng.standard_normal((20000, 256)) * rng.uniform(0.5, 2.0, (20000, 1))).astype(np.float32)
    q = rng.standard_normal((200, 256)).astype(np.float32)
    for metric in ko.METRICS:
        got_d, got_i = knn.knn(_dev(base, cuda_device), _dev(q, cuda_device), 100, metric=metric)
        assert knn.last_stats["path"] == 1
        _check(got_d, got_i, base, q, 100, metric)


@pytest.mark.parametrize("direct", ["0", "1"], ids=["alpha-beta-keys", "direct-keys"])
@pytest.mark.parametrize("metric", ["dot", "cosine", "sqeuclidean"])
def test_tc_signed_rows_with_unrankable_entries(cuda_device, monkeypatch, metric, direct):
    """Signed vectors of very different lengths, a ragged last tile, zero vectors (no cosine) and rows
    with inf / nan entries: such rows must never be returned, by either key form."""
    from neighborhoodwatch_b200 import knn
    monkeypatch.setenv("NWK_DIRECT", direct)
    rng = np.random.default_rng(11)
    nb, nq, d, k = 30011, 150, 96, 100
    base = (rng.standard_normal((nb, d)) * rng.uniform(0.01, 30.0, (nb, 1))).astype(np.float32)
    q = (rng.standard_normal((nq, d)) * 0.2).astype(np.float32)
    bad = [3, 777, 12345, nb - 1]
    base[bad[0]] = 0.0
    base[bad[1], 5] = np.inf if metric == "cosine" else np.nan      # (an inf entry gives a dot product of +-inf, which ranks)
    base[bad[2], 7] = np.nan
    base[bad[3]] = 0.0
    got_d, got_i = knn.knn(_dev(base, cuda_device), _dev(q, cuda_device), k, metric=metric)
    st = dict(knn.last_stats)
    assert st["path"] == 1 and st["max_err_ratio"] <= 1.0, st
    got_i = got_i.cpu().numpy()
    skip = bad if metric == "cosine" else bad[1:3]       # a zero vector has a dot product and a distance, but no cosine
    assert not np.isin(got_i, skip).any()
    keep = np.setdiff1d(np.arange(nb), skip)
    t_d, t_i = ko.knn_f64(base[keep], q, k, metric, return_f64=True)
    res = ko.compare_knn(got_d.cpu().numpy(), got_i, base, q, metric, t_d, keep[t_i], rtol=1e-5, atol=1e-6, tie_rtol=1e-6)
    print(metric, direct, st, res)


@pytest.mark.parametrize("ratio,expect_direct", [(4.0, True), (40.0, False)])
def test_l2_direct_keys_scale_overflow(cuda_device, ratio, expect_direct):
    """Direct squared-L2 keys take their scale from the query norms; base rows more than ~11 x longer
    than the longest query do not fit it and the call must repeat itself with alpha / beta keys."""
    from neighborhoodwatch_b200 import knn
    base = synth.vectors(30000, 128, 1234) * np.float32(ratio)
    base[::7] *= np.float32(0.25)                       # mixed lengths
    q = synth.vectors(200, 128, 5678)
    got_d, got_i = knn.knn(_dev(base, cuda_device), _dev(q, cuda_device), 100, metric="sqeuclidean", profile=True)
    st = dict(knn.last_stats)
    assert st["path"] == 1 and st["max_err_ratio"] <= 1.0, st
    # the repeated call launches the whole pipeline twice (a second-chance pass would add its own launches)
    assert (st["gpu_launches"] - (14 if st["n_retry_queries"] else 0) <= 14) == expect_direct, st
    _check(got_d, got_i, base, q, 100, "sqeuclidean")
    print(ratio, st)


def test_tc_duplicates_fall_back_and_stay_exact(cuda_device):
    """Many exact duplicates near the k-th: the candidate band cannot be certified; the fp64
    scan must take over and ties must come out in index order."""
    from neighborhoodwatch_b200 import knn
    uniq = synth.vectors(40, 64, 77)
    base = np.tile(uniq, (200, 1))                       # 8000 rows, each vector 200 times
    q = synth.vectors(16, 64, 78)
    got_d, got_i = knn.knn(_dev(base, cuda_device), _dev(q, cuda_device), 100, metric="sqeuclidean")
    st = dict(knn.last_stats)
    t_d, t_i = ko.knn_f64(base, q, 100, "sqeuclidean", return_f64=True)
    np.testing.assert_array_equal(got_i.cpu().numpy(), t_i)
    np.testing.assert_allclose(got_d.cpu().numpy(), t_d, rtol=1e-5, atol=1e-6)
    print(st)


def test_uncertifiable_queries_take_the_fp64_scan(cuda_device):
    """One huge-norm outlier row inflates the proven TF32 error bound beyond the logging
    threshold: no query can be certified, every one must be recomputed by the fp64 scan."""
    from neighborhoodwatch_b200 import knn
    base = synth.vectors(6000, 64, 1234).copy()
    base[17] *= 1000.0
    q = synth.vectors(40, 64, 5678)
    got_d, got_i = knn.knn(_dev(base, cuda_device), _dev(q, cuda_device), 100, metric="sqeuclidean")
    st = dict(knn.last_stats)
    assert st["path"] == 1 and st["n_fallback_queries"] == 40, st
    res = _check(got_d, got_i, base, q, 100, "sqeuclidean")
    assert res["index_mismatches_tied"] == 0


def test_merge_parity(cuda_device):
    import torch
    from neighborhoodwatch_b200 import knn
    rng = np.random.default_rng(1)
    for P, nq, k_in, k in [(2, 50, 100, 100), (9, 333, 100, 100), (64, 17, 100, 100), (3, 10, 7, 5), (50, 9, 100, 1000)]:
        k = min(k, P * k_in)
        d = np.sort(rng.random((P, nq, k_in)).astype(np.float32), axis=2)
        i = rng.permutation(P * nq * k_in).astype(np.int64).reshape(P, nq, k_in)
        out_d, out_i = knn.merge_topk(torch.from_numpy(d).to(cuda_device), torch.from_numpy(i).to(cuda_device), k)
        ref_d, ref_i = mo.merge_by_distance_then_index([(d[p], i[p]) for p in range(P)], k)
        np.testing.assert_array_equal(out_d.cpu().numpy(), ref_d)
        np.testing.assert_array_equal(out_i.cpu().numpy(), ref_i)
    # ties: equal distances resolve by index (deterministic, unlike np.argsort in merge.py:83)
    d = np.zeros((4, 3, 10), dtype=np.float32)
    i = rng.permutation(120).astype(np.int64).reshape(4, 3, 10)
    out_d, out_i = knn.merge_topk(torch.from_numpy(d).to(cuda_device), torch.from_numpy(i).to(cuda_device), 10)
    for r in range(3):
        assert list(out_i[r].cpu().numpy()) == sorted(i[:, r, :].ravel())[:10]
    # lists that are not sorted (the small-merge kernel ranks them by counting instead of searching)
    d = rng.random((5, 20, 50)).astype(np.float32)
    i = rng.permutation(5 * 20 * 50).astype(np.int64).reshape(5, 20, 50)
    out_d, out_i = knn.merge_topk(torch.from_numpy(d).to(cuda_device), torch.from_numpy(i).to(cuda_device), 60)
    ref_d, ref_i = mo.merge_by_distance_then_index([(d[p], i[p]) for p in range(5)], 60)
    np.testing.assert_array_equal(out_d.cpu().numpy(), ref_d)
    np.testing.assert_array_equal(out_i.cpu().numpy(), ref_i)
    # padded lists (+inf / -1, what a bounded running search emits): fewer finite entries than k_out
    d = np.sort(rng.random((3, 8, 40)).astype(np.float32), axis=2)
    i = rng.permutation(3 * 8 * 40).astype(np.int64).reshape(3, 8, 40)
    d[:, :, 10:] = np.inf
    i[:, :, 10:] = -1
    out_d, out_i = knn.merge_topk(torch.from_numpy(d).to(cuda_device), torch.from_numpy(i).to(cuda_device), 40)
    out_d, out_i = out_d.cpu().numpy(), out_i.cpu().numpy()
    ref_d, ref_i = mo.merge_by_distance_then_index([(d[p, :, :10], i[p, :, :10]) for p in range(3)], 30)
    np.testing.assert_array_equal(out_d[:, :30], ref_d)
    np.testing.assert_array_equal(out_i[:, :30], ref_i)
    assert np.all(np.isinf(out_d[:, 30:])) and np.all(out_i[:, 30:] == -1)


def test_running_topk_equals_one_shot(cuda_device):
    from neighborhoodwatch_b200 import knn
    base = synth.vectors(6000, 64, 1234)
    q = synth.vectors(100, 64, 5678)
    bd, qd = _dev(base, cuda_device), _dev(q, cuda_device)
    st = knn.RunningTopK(100, 20, cuda_device)
    for lo in range(0, 6000, 1500):
        d, i = knn.knn(bd[lo:lo + 1500], qd, 20, metric="sqeuclidean", id_offset=lo)
        st.update(d, i)
    rd, ri = st.result()
    _check(rd, ri, base, q, 20, "sqeuclidean")
    st.close()


@pytest.mark.parametrize("metric", ko.METRICS)
@pytest.mark.parametrize("nb,slab,d,k,order", [
    (6000, 1500, 64, 20, "as-is"),         # slabs small enough to be logged whole: the bound alone thins them out
    (160000, 40000, 128, 100, "as-is"),    # sampled slabs; later slabs pass fewer than k rows for most queries
    (160000, 40000, 128, 100, "far-first"),  # every slab holds nearer rows than the one before: the bound never bites
    (90000, 30000, 768, 100, "near-first"),  # the first slab already holds most of the answer (CTA-pair filter)
])
def test_running_search_equals_one_shot(cuda_device, metric, nb, slab, d, k, order):
    """nwk_topk_state_search (knn of a slab bounded by the state's k-th distances, folded into the
    state) over all slabs == the oracle on the whole base, whatever the order the rows arrive in."""
    from neighborhoodwatch_b200 import knn
    base = synth.vectors(nb, d, 1234)
    q = synth.vectors(64, d, 5678)
    if order != "as-is":
        # sort the rows by distance to the mean query, so slabs differ systematically
        key = ko.distances_f64(base, q.mean(axis=0, keepdims=True), metric)[0]
        perm = np.argsort(key if order == "near-first" else -key, kind="stable")
        base = np.ascontiguousarray(base[perm])
    bd, qd = _dev(base, cuda_device), _dev(q, cuda_device)
    st = knn.RunningTopK(len(q), k, cuda_device)
    cands = []
    for lo in range(0, nb, slab):
        st.search(bd[lo:lo + slab], qd, metric=metric, id_offset=lo)
        assert knn.last_stats["max_err_ratio"] <= 1.0, knn.last_stats
        cands.append(knn.last_stats["total_candidates"])
    rd, ri = st.result()
    res = _check(rd, ri, base, q, k, metric)
    st.close()
    if order != "far-first":
        assert cands[-1] < cands[0], cands           # the bound must actually thin out the later slabs
    print(metric, (nb, slab, d, k, order), cands, res)


def test_host_entry_point(cuda_device):
    from neighborhoodwatch_b200 import knn
    base = synth.vectors(30000, 128, 1234)
    q = synth.vectors(300, 128, 5678)
    import torch
    out_d, out_i = knn.knn_host(base, q, 100, metric="cosine", slab_rows=7000)     # 5 slabs, ragged last
    _check(torch.from_numpy(out_d), torch.from_numpy(out_i), base, q, 100, "cosine")
    out_d, out_i = knn.knn_host(base, q, 100, metric="dot")                          # one slab
    _check(torch.from_numpy(out_d), torch.from_numpy(out_i), base, q, 100, "dot")


def test_error_behaviour(cuda_device):
    from neighborhoodwatch_b200 import _native, knn
    import torch
    base = _dev(synth.vectors(50, 16, 1), cuda_device)
    q = _dev(synth.vectors(5, 16, 2), cuda_device)
    with pytest.raises(AssertionError):              # cu_knn.py:237
        knn.knn(base, q, 51)
    with pytest.raises(RuntimeError):
        knn.knn(base.cpu(), q.cpu(), 5)              # no CPU path
    lib = _native.load()
    rc = lib.nwk_knn_f32(base.data_ptr(), 50, q.data_ptr(), 5, 16, 10, 0, 0, q.data_ptr(), q.data_ptr(), None, 0, 0,
                         None, None)
    assert rc == 3 and b"workspace" in lib.nwk_last_error()
    d, i = knn.knn(base, q[:0], 5)
    assert d.shape == (0, 5)


def test_large_shape_properties(cuda_device):
    """BASELINE config-2 arithmetic (d=768, cosine) at a size whose truth the oracle still
    computes in seconds for a sample of queries; the rest is checked against the on-device
    fp64 scan and through size-independent properties."""
    import torch
    from neighborhoodwatch_b200 import knn
    nq, nb, d, k = 2048, 200_000, 768, 100
    bd = synth.device_vectors(nb, d, 1234, device=cuda_device)
    qd = synth.device_vectors(nq, d, 5678, device=cuda_device)
    got_d, got_i = knn.knn(bd, qd, k, metric="cosine")
    st = dict(knn.last_stats)
    assert st["path"] == 1
    assert torch.all(got_d[:, 1:] >= got_d[:, :-1])                       # ascending (merge.py:104)
    assert int(got_i.min()) >= 0 and int(got_i.max()) < nb
    srt = torch.sort(got_i, dim=1).values
    assert torch.all(srt[:, 1:] != srt[:, :-1])                           # no repeated neighbour
    sel = torch.arange(0, nq, 16, device=cuda_device)                     # 128 sampled queries
    ex_d, ex_i = knn.knn(bd, qd[sel], k, metric="cosine", force_exact_scan=True)
    assert torch.equal(ex_i, got_i[sel]) and torch.equal(ex_d, got_d[sel])
    # oracle truth for 32 of them
    base = synth.vectors(nb, d, 1234)
    q = synth.vectors(nq, d, 5678)[::64]
    _check(got_d[::64], got_i[::64], base, q, k, "cosine")
    # sharding property: top-k of the union == merge of the shards' top-k
    parts = [knn.knn(bd[lo:lo + 50_000], qd[:256], k, metric="cosine", id_offset=lo) for lo in range(0, nb, 50_000)]
    m_d, m_i = knn.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), k)
    assert torch.equal(m_d, got_d[:256])
    # the merge orders by (fp32 distance, id), the one-shot result by (fp64 distance, id): ids may
    # differ only inside runs of fp32-equal distances, and each row must hold the same id set
    diff = (m_i != got_i[:256])
    assert torch.equal(torch.sort(m_i, dim=1).values, torch.sort(got_i[:256], dim=1).values)
    left = torch.zeros_like(diff)
    left[:, 1:] = m_d[:, 1:] == m_d[:, :-1]
    right = torch.zeros_like(diff)
    right[:, :-1] = m_d[:, :-1] == m_d[:, 1:]
    assert bool(torch.all(~diff | left | right))
    assert float(diff.float().mean()) < 5e-3
    print(st)


@pytest.mark.parametrize("metric,d", [("dot", 128), ("sqeuclidean", 128), ("dot", 384)])
def test_short_row_shapes_against_the_fp64_scan(cuda_device, metric, d):
    """BASELINE config-5 / config-1 arithmetic (short rows: resident query tile, direct keys for dot)
    at a size with many query tiles and sample parts: sampled queries must equal the on-device fp64
    scan bit for bit, all rows must be ascending and free of repeats."""
    import torch
    from neighborhoodwatch_b200 import knn
    nq, nb, k = 4000, 600_000, 100
    bd = synth.device_vectors(nb, d, 1234, device=cuda_device)
    qd = synth.device_vectors(nq, d, 5678, device=cuda_device)
    got_d, got_i = knn.knn(bd, qd, k, metric=metric, id_offset=7)
    st = dict(knn.last_stats)
    assert st["path"] == 1 and st["max_err_ratio"] <= 1.0, st
    assert st["n_fallback_queries"] <= nq // 100, st
    assert torch.all(got_d[:, 1:] >= got_d[:, :-1])
    srt = torch.sort(got_i, dim=1).values
    assert torch.all(srt[:, 1:] != srt[:, :-1])
    sel = torch.arange(0, nq, 40, device=cuda_device)                     # 100 sampled queries
    ex_d, ex_i = knn.knn(bd, qd[sel], k, metric=metric, id_offset=7, force_exact_scan=True)
    assert torch.equal(ex_i, got_i[sel]) and torch.equal(ex_d, got_d[sel])
    print(metric, d, st)


def test_post_validator_reference_rule(cuda_device):
    """parquet_to_format.py:351-383: isclose(1 - dot, distance / 2) for every returned pair."""
    import torch
    from neighborhoodwatch_b200 import knn
    from neighborhoodwatch_b200 import parquet_to_format as fmt
    base = synth.vectors(20000, 128, 1234)
    q = synth.vectors(200, 128, 5678)
    q[7] = 0.0                                            # zero query vectors are skipped by the reference
    bd, qd = _dev(base, cuda_device), _dev(q, cuda_device)
    d, i = knn.knn(bd, qd, 100, metric="sqeuclidean")
    assert fmt.validate_pairs(bd, qd, i, d) == 0
    # the same check on the host for a few pairs, as the reference does it
    dn, inn = d.cpu().numpy(), i.cpu().numpy()
    for r in (0, 3, 199):
        for c in (0, 50, 99):
            dot = float(np.dot(q[r].astype(np.float64), base[inn[r, c]].astype(np.float64)))
            assert np.isclose(1 - dot, dn[r, c] / 2)
    bad = d.clone()
    bad[3, 5] *= 1.01
    bad[9, 0] += 1e-3
    assert fmt.validate_pairs(bd, qd, i, bad) == 2
    wrong = i.clone()
    wrong[11, 2] = (wrong[11, 2] + 1) % 20000
    assert fmt.validate_pairs(bd, qd, wrong, d) == 1


def test_config2_full_size(cuda_device):
    """BASELINE config 2 at full size (10 000 x 1 000 000 x 768, cosine, k=100): sampled queries
    against the on-device fp64 scan, the reference's validator rule on the squared-L2 twin, and
    size-independent properties."""
    import torch
    from neighborhoodwatch_b200 import knn
    from neighborhoodwatch_b200 import parquet_to_format as fmt
    nq, nb, d, k = 10_000, 1_000_000, 768, 100
    bd = synth.device_vectors(nb, d, synth.BASE_SEED, device=cuda_device)
    qd = synth.device_vectors(nq, d, synth.QUERY_SEED, device=cuda_device)
    got_d, got_i = knn.knn(bd, qd, k, metric="cosine")
    st = dict(knn.last_stats)
    assert st["path"] == 1 and st["max_err_ratio"] <= 1.0, st
    assert st["n_fallback_queries"] <= 10, st
    assert torch.all(got_d[:, 1:] >= got_d[:, :-1])
    assert int(got_i.min()) >= 0 and int(got_i.max()) < nb
    srt = torch.sort(got_i, dim=1).values
    assert torch.all(srt[:, 1:] != srt[:, :-1])
    sel = torch.arange(0, nq, 157, device=cuda_device)[:64]
    ex_d, ex_i = knn.knn(bd, qd[sel], k, metric="cosine", force_exact_scan=True)
    assert torch.equal(ex_i, got_i[sel]) and torch.equal(ex_d, got_d[sel])
    # unit vectors: cosine ranking == squared-L2 ranking, and the reference's validator rule holds
    l2_d, l2_i = knn.knn(bd, qd[:2000], k, metric="sqeuclidean")
    assert fmt.validate_pairs(bd, qd[:2000], l2_i, l2_d) == 0
    agree = (l2_i == got_i[:2000]).float().mean()
    assert float(agree) > 0.999        # identical up to fp64-level ties between the two formulas
    print(st)


def test_clustered_base_sorted_by_cluster(cuda_device):
    """Real embedding sets are clustered and often stored cluster by cluster: the strided sample
    must still give usable thresholds, and whatever falls back must stay exact."""
    from neighborhoodwatch_b200 import knn
    rng = np.random.default_rng(11)
    d, n_clusters, per = 96, 40, 1500
    centers = rng.standard_normal((n_clusters, d)).astype(np.float32)
    base = np.concatenate([c + 0.05 * rng.standard_normal((per, d)).astype(np.float32) for c in centers])
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    q = centers[rng.integers(0, n_clusters, 300)] + 0.05 * rng.standard_normal((300, d)).astype(np.float32)
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    for metric in ("sqeuclidean", "cosine"):
        got_d, got_i = knn.knn(_dev(base, cuda_device), _dev(q, cuda_device), 100, metric=metric)
        st = dict(knn.last_stats)
        _check(got_d, got_i, base, q, 100, metric)
        assert st["path"] == 1 and st["max_err_ratio"] <= 1.0
        print(metric, st)


@pytest.mark.parametrize("nq,tq,nd,td,d,k,ragged", [
    (24, 32, 3000, 180, 128, 100, False),     # the ColBERT shape of BASELINE config 5 (query_maxlen 32, doc_maxlen 180, d 128)
    (9, 32, 700, 180, 128, 50, True),         # ragged documents, an odd number of queries (the last query tile is mostly padding)
    (40, 8, 500, 64, 96, 10, True),           # short queries: 16 of them per 128-row tile; d not a multiple of 64
    (6, 64, 400, 256, 64, 20, False),         # queries of two warps' worth of tokens (summed by atomics), the longest documents
    (5, 32, 30, 33, 128, 30, True),           # every document is a candidate
])
def test_maxsim_parity(cuda_device, nq, tq, nd, td, d, k, ragged):
    """Document-level MaxSim through the C ABI against the float64 oracle (oracle/maxsim_oracle.py):
    scores within 1e-5 relative, document ids identical except where exact scores tie within 1e-6."""
    import torch
    from oracle import maxsim_oracle as mso
    from neighborhoodwatch_b200 import knn
    rng = np.random.default_rng(7)
    qt = rng.standard_normal((nq, tq, d)).astype(np.float32)
    qt /= np.linalg.norm(qt, axis=2, keepdims=True)
    dt = rng.standard_normal((nd, td, d)).astype(np.float32)
    dt /= np.linalg.norm(dt, axis=2, keepdims=True)
    doc_len = None
    if ragged:
        doc_len = rng.integers(1, td + 1, size=nd).astype(np.int32)
        doc_len[::97] = 0                                       # empty documents never rank
    got_s, got_i = knn.maxsim_topk(torch.from_numpy(qt).to(cuda_device), torch.from_numpy(dt).to(cuda_device), k,
                                   doc_len=None if doc_len is None else torch.from_numpy(doc_len).to(cuda_device), id_offset=3)
    st = dict(knn.last_stats)
    assert st["path"] == 3 and st["max_err_ratio"] <= 1.0, st
    got_s, got_i = got_s.cpu().numpy(), got_i.cpu().numpy()
    full = mso.maxsim_scores_f64(qt, dt, doc_len)
    t_s, t_i = mso.maxsim_topk_f64(qt, dt, k, doc_len)
    finite = np.isfinite(t_s)
    np.testing.assert_allclose(got_s[finite], t_s[finite], rtol=1e-5, atol=1e-6)
    assert np.all(np.diff(got_s, axis=1)[finite[:, 1:]] <= 0)
    mism = np.argwhere((got_i - 3 != t_i) & finite)
    for r, c in mism:                                           # only exact ties may swap
        assert abs(full[r, got_i[r, c] - 3] - t_s[r, c]) <= 1e-6 * max(1.0, abs(t_s[r, c])), (r, c)
    print((nq, tq, nd, td, d, k, ragged), st, len(mism))


def test_maxsim_uncertified_queries_are_rescored_exactly(cuda_device):
    """Every document occurs 700 times: far more than the 256 candidates tie with the k-th score, no
    query can be certified, and all of them must be re-scored against every document (ids in order)."""
    import torch
    from oracle import maxsim_oracle as mso
    from neighborhoodwatch_b200 import knn
    rng = np.random.default_rng(3)
    uniq = rng.standard_normal((3, 40, 64)).astype(np.float32)
    dt = np.tile(uniq, (700, 1, 1))                            # 2100 documents, each one of 3
    qt = rng.standard_normal((7, 32, 64)).astype(np.float32)
    got_s, got_i = knn.maxsim_topk(torch.from_numpy(qt).to(cuda_device), torch.from_numpy(dt).to(cuda_device), 20)
    st = dict(knn.last_stats)
    assert st["n_fallback_queries"] == 7, st
    t_s, t_i = mso.maxsim_topk_f64(qt, dt, 20)
    np.testing.assert_array_equal(got_i.cpu().numpy(), t_i)
    np.testing.assert_allclose(got_s.cpu().numpy(), t_s, rtol=1e-5, atol=1e-6)


def test_maxsim_error_behaviour(cuda_device):
    import torch
    from neighborhoodwatch_b200 import knn
    from neighborhoodwatch_b200._native import NwkError
    q = torch.zeros((2, 24, 64), device=cuda_device)           # 24 tokens: not a power of two
    d = torch.zeros((50, 40, 64), device=cuda_device)
    with pytest.raises(NwkError):
        knn.maxsim_topk(q, d, 5)
    with pytest.raises(AssertionError):
        knn.maxsim_topk(torch.zeros((2, 32, 64), device=cuda_device), d, 51)
    with pytest.raises(RuntimeError):
        knn.maxsim_topk(torch.zeros((2, 32, 64)), d.cpu(), 5)


def test_maxsim_shards_merge_to_the_one_shot_result(cuda_device):
    """Documents split into three shards (global ids through id_offset), merged as negated scores by the
    kNN reducer == the one-shot top-k (the multi-GPU path of dist.sharded_maxsim on one device)."""
    import torch
    from neighborhoodwatch_b200 import knn
    rng = np.random.default_rng(21)
    qt = torch.from_numpy(rng.standard_normal((16, 32, 128)).astype(np.float32)).to(cuda_device)
    dt = torch.from_numpy(rng.standard_normal((900, 100, 128)).astype(np.float32)).to(cuda_device)
    one_s, one_i = knn.maxsim_topk(qt, dt, 50)
    parts = [knn.maxsim_topk(qt, dt[lo:lo + 300], 50, id_offset=lo) for lo in range(0, 900, 300)]
    m_d, m_i = knn.merge_topk(torch.stack([-p[0] for p in parts]), torch.stack([p[1] for p in parts]), 50)
    assert torch.equal(-m_d, one_s)
    # the merge orders by (fp32 score, id), the one-shot result by (fp64 score, id): ids may differ only
    # inside runs of fp32-equal scores, and each row must hold the same id set
    assert torch.equal(torch.sort(m_i, dim=1).values, torch.sort(one_i, dim=1).values)
    diff = m_i != one_i
    left = torch.zeros_like(diff)
    left[:, 1:] = m_d[:, 1:] == m_d[:, :-1]
    right = torch.zeros_like(diff)
    right[:, :-1] = m_d[:, :-1] == m_d[:, 1:]
    assert bool(torch.all(~diff | left | right))


@pytest.mark.parametrize("metric,d", [("dot", 128), ("sqeuclidean", 768), ("cosine", 96)])
def test_second_chance_pass_stays_exact(cuda_device, monkeypatch, metric, d):
    """A low threshold rank (forced here: rank 8) leaves many queries uncertified; they must come back
    through the second tensor-core pass (stats.n_retry_queries) -- not the fp64 scan -- with results
    that equal the scan bit for bit, also when the call is a bounded running search."""
    import torch
    from neighborhoodwatch_b200 import knn
    monkeypatch.setenv("NWK_SAMPLE_M", "8")
    nq, nb, k = 3000, 400_000, 100
    bd = synth.device_vectors(nb, d, 1234, device=cuda_device)
    qd = synth.device_vectors(nq, d, 5678, device=cuda_device)
    got_d, got_i = knn.knn(bd, qd, k, metric=metric, id_offset=5)
    st = dict(knn.last_stats)
    assert st["path"] == 1 and st["max_err_ratio"] <= 1.0, st
    assert st["n_retry_queries"] > 0, st
    assert st["n_fallback_queries"] <= st["n_retry_queries"] // 4 + 2, st
    sel = torch.arange(0, nq, 10, device=cuda_device)
    ex_d, ex_i = knn.knn(bd, qd[sel], k, metric=metric, id_offset=5, force_exact_scan=True)
    assert torch.equal(ex_i, got_i[sel]) and torch.equal(ex_d, got_d[sel])
    _check(got_d[::50], got_i[::50], bd.cpu().numpy(), qd.cpu().numpy()[::50], k, metric, id_offset=5)
    # running search over two slabs: the second slab's thresholds are lowered by the state's bound
    state = knn.RunningTopK(nq, k, cuda_device)
    retried = 0
    for lo in (0, nb // 2):
        state.search(bd[lo:lo + nb // 2], qd, metric=metric, id_offset=5 + lo)
        retried += knn.last_stats["n_retry_queries"]
    r_d, r_i = state.result()
    state.close()
    assert retried > 0
    assert torch.equal(r_d, got_d)
    # (the running merge orders by (fp32 distance, id), the one-shot result by (fp64 distance, id): a tie in
    # fp32 that straddles the k-th slot may legitimately keep a different row)
    same_rows = (torch.sort(r_i, dim=1).values == torch.sort(got_i, dim=1).values).all(dim=1).float().mean()
    assert float(same_rows) > 0.995, float(same_rows)
    print(metric, d, st)
