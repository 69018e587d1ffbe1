"""Document-level MaxSim at the shape of BASELINE config 5 (one GPU's share at 8 GPUs by default):
1000 queries x 32 tokens against 12 500 documents x 180 tokens, d = 128, k = 100."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neighborhoodwatch_b200 import knn, synth

nq, tq, nd, td, d, k, steps = [int(x) for x in (sys.argv[1:8] or "1000 32 12500 180 128 100 10".split())]
dev = torch.device("cuda", 0)
q = synth.device_vectors(nq * tq, d, synth.QUERY_SEED, device=dev).view(nq, tq, d)
docs = synth.device_vectors(nd * td, d, synth.BASE_SEED, device=dev).view(nd, td, d)
# signed token embeddings (the synthetic generator is all-positive): flip signs by a fixed pattern
sign = (torch.arange(d, device=dev) % 3 == 0).float() * -2 + 1
q = (q * sign).contiguous(); docs = (docs * sign.roll(1)).contiguous()
for _ in range(3):
    s, i = knn.maxsim_topk(q, docs, k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    s, i = knn.maxsim_topk(q, docs, k)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
pairs = nq * tq * nd * td
print(json.dumps({"workload": f"maxsim: {nq} queries x {tq} tokens vs {nd} documents x {td} tokens, d={d}, k={k}",
                  "ms_per_step": ms, "token_pairs_per_s": pairs / (ms * 1e-3), "tflops": 2 * pairs * d / (ms * 1e-3) / 1e12,
                  "stats": knn.last_stats}))
