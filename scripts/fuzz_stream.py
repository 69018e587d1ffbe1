"""Random shapes: a running search over random slab splits (nwk_topk_state_search, every slab bounded by the
k-th distances found so far) must give the one-shot result: same distances, same id sets (ids may swap only
inside runs of fp32-equal distances, and at a tie with the k-th distance either row may make the cut: the state
orders by fp32 distance, the one-shot call by fp64).
Usage: fuzz_stream.py [n] [seed]"""
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
    d = int(rng.choice([8, 64, 100, 128, 256, 768]))
    nq = int(rng.choice([1, 50, 200, 700]))
    nb = int(rng.choice([3000, 20000, 90000]))
    k = int(rng.choice([1, 10, 100]))
    metric = str(rng.choice(["sqeuclidean", "cosine", "dot"]))
    nslab = int(rng.choice([2, 3, 7]))
    order = str(rng.choice(["as-is", "near-first", "far-first"]))
    g = torch.Generator(device=dev).manual_seed(it + 5000)
    base = torch.randn((nb, d), generator=g, device=dev) * (0.5 + torch.rand((nb, 1), generator=g, device=dev))
    q = torch.randn((nq, d), generator=g, device=dev)
    if order != "as-is":
        key = (base - q.mean(dim=0, keepdim=True)).norm(dim=1)
        base = base[torch.argsort(key, descending=(order == "far-first"))].contiguous()
    one_d, one_i = knn.knn(base, q, k, metric=metric)
    cuts = sorted(set([0, nb] + [int(x) for x in rng.integers(k, nb - k, size=nslab - 1)]))
    cuts = [c for j, c in enumerate(cuts) if j == 0 or c - cuts[j - 1] >= k] if all(b - a >= k for a, b in zip(cuts, cuts[1:])) else [0, nb // 2, nb]
    st = knn.RunningTopK(nq, k, dev)
    for lo, hi in zip(cuts, cuts[1:]):
        st.search(base[lo:hi], q, metric=metric, id_offset=lo)
    r_d, r_i = st.result()
    st.close()
    ok = torch.equal(r_d, one_d)
    if ok and not torch.equal(torch.sort(r_i, dim=1).values, torch.sort(one_i, dim=1).values):
        # the id sets may differ only by rows that tie (in fp32) with the k-th distance: which of them
        # makes the cut is decided by fp64 in the one-shot call and by the id in the state
        for r in range(nq):
            a, b = set(r_i[r].tolist()), set(one_i[r].tolist())
            for ids, dd, ii in ((a - b, r_d[r], r_i[r]), (b - a, one_d[r], one_i[r])):
                for x in ids:
                    pos = int((ii == x).nonzero()[0])
                    ok = ok and bool(dd[pos] == dd[k - 1])
    bad += 0 if ok else 1
    print(("ok  " if ok else "FAIL"), dict(nq=nq, nb=nb, d=d, k=k, metric=metric, slabs=len(cuts) - 1, order=order), flush=True)
print(f"{n - bad}/{n} running searches equal the one-shot result in {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)
