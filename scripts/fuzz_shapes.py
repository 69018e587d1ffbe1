"""Random shapes / metrics: the tensor-core path must equal the on-device fp64 scan bit for bit
(both end in the same fp64 arithmetic and the same (distance, index) order).  Usage: fuzz_shapes.py [n] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from neighborhoodwatch_b200 import knn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda", 0)
bad = 0
t0 = time.time()
for it in range(n):
    d = int(rng.choice([4, 8, 12, 32, 64, 100, 128, 132, 200, 256, 384, 512, 768, 1024]))
    nq = int(rng.choice([1, 7, 128, 129, 300, 600, 1500]))
    nb = int(rng.choice([300, 2049, 5000, 33333, 120000]))
    if nb * d > 60_000_000:
        nb = 60_000_000 // d
    k = int(rng.choice([1, 10, 100, 256]))
    k = min(k, nb)
    metric = str(rng.choice(["sqeuclidean", "cosine", "dot"]))
    kind = str(rng.choice(["unit-positive", "gaussian", "scaled"]))
    g = torch.Generator(device=dev).manual_seed(it + 1000)
    if kind == "unit-positive":
        base = torch.rand((nb, d), generator=g, device=dev); q = torch.rand((nq, d), generator=g, device=dev)
        base /= base.norm(dim=1, keepdim=True); q /= q.norm(dim=1, keepdim=True)
    elif kind == "gaussian":
        base = torch.randn((nb, d), generator=g, device=dev); q = torch.randn((nq, d), generator=g, device=dev)
    else:
        base = torch.randn((nb, d), generator=g, device=dev) * (0.1 + 5 * torch.rand((nb, 1), generator=g, device=dev))
        q = torch.randn((nq, d), generator=g, device=dev) * 3
    got_d, got_i = knn.knn(base, q, k, metric=metric, id_offset=it)
    st = dict(knn.last_stats)
    sel = torch.arange(0, nq, max(1, nq // 64), device=dev)
    ex_d, ex_i = knn.knn(base, q[sel], k, metric=metric, id_offset=it, force_exact_scan=True)
    ok = torch.equal(ex_i, got_i[sel]) and torch.equal(ex_d, got_d[sel]) and st["max_err_ratio"] <= 1.0
    bad += 0 if ok else 1
    print(("ok  " if ok else "FAIL"), dict(nq=nq, nb=nb, d=d, k=k, metric=metric, kind=kind), "path", st["path"], "fallback",
          st["n_fallback_queries"], "err", round(st["max_err_ratio"], 3), flush=True)
print(f"{n - bad}/{n} shapes equal the fp64 scan in {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)
