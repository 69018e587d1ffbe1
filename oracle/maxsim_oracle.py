"""CPU oracle for document-level ColBERT MaxSim.  TEST INFRASTRUCTURE ONLY.

    score(Q, D) = sum over the query's tokens t of  max over the document's valid tokens j of  q_t . d_j

Parity unpinned: the reference (phact/neighborhoodwatch, colbert_knn.py:62-126) computes a
token-level kNN over the flattened token embeddings and has no document-level score anywhere
(semantics note for config 5); this float64 numpy restatement of the
published ColBERT late-interaction score (Khattab & Zaharia 2020, eq. 1) is the only truth the
CUDA path ``nwk_maxsim_f32`` is compared with.  Order: (score descending, document id).
"""
from __future__ import annotations

import numpy as np


def maxsim_scores_f64(query_tokens: np.ndarray, doc_tokens: np.ndarray, doc_len=None) -> np.ndarray:
    """[nq, nd] float64 score matrix (small inputs only)."""
    q = np.asarray(query_tokens, dtype=np.float64)
    d = np.asarray(doc_tokens, dtype=np.float64)
    nq, nd, td = q.shape[0], d.shape[0], d.shape[1]
    out = np.empty((nq, nd), dtype=np.float64)
    for i in range(nd):
        n = td if doc_len is None else int(max(0, min(int(doc_len[i]), td)))
        if n == 0:
            out[:, i] = -np.inf
            continue
        sim = np.einsum("qtc,jc->qtj", q, d[i, :n])          # [nq, Tq, n]
        out[:, i] = sim.max(axis=2).sum(axis=1)
    return out


def maxsim_topk_f64(query_tokens, doc_tokens, k, doc_len=None):
    """(scores float64 [nq, k] descending, ids int64 [nq, k]), ties by document id."""
    s = maxsim_scores_f64(query_tokens, doc_tokens, doc_len)
    nq, nd = s.shape
    ids = np.arange(nd)
    out_s = np.empty((nq, k), dtype=np.float64)
    out_i = np.empty((nq, k), dtype=np.int64)
    for r in range(nq):
        order = np.lexsort((ids, -s[r]))[:k]
        out_s[r] = s[r, order]
        out_i[r] = order
    return out_s, out_i
