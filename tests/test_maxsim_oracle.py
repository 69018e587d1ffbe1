"""The MaxSim oracle against a direct triple loop (tiny shapes) and its ordering rules."""
import numpy as np

from oracle import maxsim_oracle as mso


def test_maxsim_oracle_matches_loops():
    rng = np.random.default_rng(0)
    q = rng.standard_normal((3, 4, 5)).astype(np.float32)
    d = rng.standard_normal((6, 7, 5)).astype(np.float32)
    doc_len = np.array([7, 3, 0, 1, 7, 5], dtype=np.int32)
    got = mso.maxsim_scores_f64(q, d, doc_len)
    for a in range(3):
        for b in range(6):
            if doc_len[b] == 0:
                assert got[a, b] == -np.inf
                continue
            want = 0.0
            for t in range(4):
                want += max(float(np.dot(q[a, t].astype(np.float64), d[b, j].astype(np.float64))) for j in range(doc_len[b]))
            assert abs(got[a, b] - want) < 1e-12


def test_maxsim_topk_order_and_ties():
    q = np.ones((1, 2, 4), dtype=np.float32)
    d = np.zeros((5, 3, 4), dtype=np.float32)
    d[1] = 1.0
    d[3] = 1.0                      # documents 1 and 3 tie at the top: the smaller id first
    s, i = mso.maxsim_topk_f64(q, d, 3)
    assert list(i[0]) == [1, 3, 0]
    assert s[0, 0] == s[0, 1] == 8.0 and s[0, 2] == 0.0
