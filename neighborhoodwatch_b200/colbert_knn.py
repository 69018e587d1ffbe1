"""ColBERT token-kNN stage: ``process_knn_computation`` of neighborhoodwatch/colbert_knn.py:84-126.

Only the kNN entry point of ``ck`` is on the hot path; the
token-embedding dataset builder needs network + model weights and is out of scope.  As in
the reference this is *flattened token-level* kNN: every query-token vector is a query and
every document-token vector a base row, with the torch engine's ``1 - q.b`` distance
(cu_knn.py:256) -- not a document-level MaxSim.  ``maxsim_topk`` below adds that score as an extra
metric: the reference never computes it, so its parity is pinned by
``oracle/maxsim_oracle.py`` only.
"""
from __future__ import annotations

import math

import pyarrow.parquet as pq

from .cu_knn import process_batches, tune_memory


def process_knn_computation(data_dir,
                            base_filename,
                            base_count,
                            query_filename,
                            query_count,
                            mem_tune=False,
                            initial_batch_size=1000000,
                            max_memory_threshold=0.1,
                            k=100,
                            split=False,
                            engine='torch'):
    if engine == 'torch':
        if initial_batch_size > 100000:
            # colbert_knn.py:95-98.  The reference shrinks the batch to 9000 rows because its
            # torch engine materialises the [n_query_tokens, batch] distance matrix in HBM;
            # the fused kernel never does, but the batch size also fixes how many partial
            # files are written, so the rule is kept to stay file-for-file compatible.
            initial_batch_size = 9000

    # colbert_knn.py:102-104: token-embedding files are read whole, no column filtering
    query_table = pq.read_table(query_filename).slice(0, query_count)
    base_table = pq.read_table(base_filename).slice(0, base_count)

    batch_size = initial_batch_size
    if mem_tune:
        batch_size = tune_memory(base_table, batch_size, max_memory_threshold)

    batch_count = math.ceil(len(base_table) / batch_size)
    assert (len(base_table) % batch_size == 0) or k <= (len(base_table) % batch_size), \
        f"Cannot generate k of {k} with only {len(base_table)} rows and batch_size of {batch_size}."

    process_batches(data_dir, base_table, query_table, batch_count, batch_size, k, split, engine=engine)


def maxsim_topk(query_tokens, doc_tokens, k=100, doc_len=None, id_offset=0):
    """Top-k documents per query by ColBERT MaxSim, ``score(Q, D) = sum_t max_j q_t . d_j``.

    ``query_tokens`` [n_queries, query_maxlen, d] and ``doc_tokens`` [n_docs, doc_maxlen, d] are the
    per-token embeddings the reference flattens into its ``token_embedding_*`` files
    (colbert_knn.py:62-66,229); ``doc_len`` gives the valid tokens of ragged documents.
    """
    from . import knn as _knn
    return _knn.maxsim_topk(query_tokens, doc_tokens, k, doc_len=doc_len, id_offset=id_offset)
