"""Small end-to-end run of every kernel family, meant to be wrapped in compute-sanitizer."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neighborhoodwatch_b200 import knn, synth  # noqa: E402
from neighborhoodwatch_b200 import parquet_to_format as fmt  # noqa: E402

dev = torch.device("cuda", 0)
base = synth.device_vectors(20000, 128, 1234, device=dev)
q = synth.device_vectors(300, 128, 5678, device=dev)
for metric in ("sqeuclidean", "cosine", "dot"):
    d1, i1 = knn.knn(base, q, 100, metric=metric)
    print(metric, knn.last_stats["path"], knn.last_stats["n_fallback_queries"], knn.last_stats["max_err_ratio"])
d2, i2 = knn.knn(base, q[:40], 100, metric="dot", force_exact_scan=True)
assert torch.equal(i2, i1[:40])
d3, i3 = knn.knn(base, q, 100, metric="sqeuclidean", tf32=True)
parts = [knn.knn(base[lo:lo + 5000], q, 100, metric="sqeuclidean", id_offset=lo) for lo in range(0, 20000, 5000)]
m_d, m_i = knn.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), 100)
assert torch.equal(m_d, d3)
assert fmt.validate_pairs(base, q, m_i, m_d) == 0
o_d, o_i = knn.knn_host(base.cpu().numpy(), q.cpu().numpy(), 100, metric="cosine", slab_rows=6000)
# CTA-pair filter (d >= 512) with both key forms, resident query tile with alpha / beta keys, running search
b768 = synth.device_vectors(6000, 768, 1234, device=dev)
q768 = synth.device_vectors(300, 768, 5678, device=dev)
for metric in ("sqeuclidean", "cosine"):
    knn.knn(b768, q768, 100, metric=metric)
    print("pair", metric, knn.last_stats["path"], knn.last_stats["n_fallback_queries"])
st = knn.RunningTopK(300, 100, dev)
for lo in range(0, 20000, 5000):
    st.search(base[lo:lo + 5000], q, metric="sqeuclidean", id_offset=lo)
r_d, r_i = st.result()
assert torch.equal(r_d, d3)
st.close()
print("ok")
