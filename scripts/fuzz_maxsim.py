"""Random MaxSim shapes against the float64 oracle (scores 1e-5 relative; ids equal except at exact-score ties).
Usage: fuzz_maxsim.py [n] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from neighborhoodwatch_b200 import knn
from oracle import maxsim_oracle as mso

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda", 0)
bad = 0
t0 = time.time()
for it in range(n):
    tq = int(rng.choice([1, 2, 4, 8, 16, 32, 64, 128]))
    td = int(rng.integers(1, 257))
    d = int(rng.choice([4, 8, 20, 64, 96, 128, 200, 256]))
    nq = int(rng.choice([1, 3, 9, 20]))
    nd = int(rng.choice([1, 17, 300, 1200]))
    k = int(min(nd, rng.choice([1, 10, 100])))
    ragged = bool(rng.integers(0, 2))
    scale = float(rng.choice([1.0, 0.01, 30.0]))
    qt = (rng.standard_normal((nq, tq, d)) * scale).astype(np.float32)
    dt = rng.standard_normal((nd, td, d)).astype(np.float32)
    doc_len = rng.integers(0, td + 1, size=nd).astype(np.int32) if ragged else None
    got_s, got_i = knn.maxsim_topk(torch.from_numpy(qt).to(dev), torch.from_numpy(dt).to(dev), k,
                                   doc_len=None if doc_len is None else torch.from_numpy(doc_len).to(dev))
    st = dict(knn.last_stats)
    got_s, got_i = got_s.cpu().numpy(), got_i.cpu().numpy()
    full = mso.maxsim_scores_f64(qt, dt, doc_len)
    t_s, t_i = mso.maxsim_topk_f64(qt, dt, k, doc_len)
    fin = np.isfinite(t_s)
    ok = st["max_err_ratio"] <= 1.0 and np.allclose(got_s[fin], t_s[fin], rtol=1e-5, atol=1e-6 * max(1.0, scale))
    ok = ok and bool(np.all(np.isinf(got_s[~fin])))
    for r, c in np.argwhere((got_i != t_i) & fin):
        ok = ok and abs(full[r, got_i[r, c]] - t_s[r, c]) <= 1e-6 * max(1.0, abs(t_s[r, c]))
    bad += 0 if ok else 1
    print(("ok  " if ok else "FAIL"), dict(nq=nq, tq=tq, nd=nd, td=td, d=d, k=k, ragged=ragged, scale=scale), "fallback",
          st["n_fallback_queries"], "err", round(st["max_err_ratio"], 3), flush=True)
print(f"{n - bad}/{n} MaxSim shapes equal the oracle in {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)
