"""Document-level MaxSim at the shape of BASELINE config 5 (one GPU's share at 8 GPUs by default):
1000 queries x 32 tokens against 12 500 documents x 180 tokens, d = 128, k = 100."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neighborhoodwatch_b200 import knn, synth

nq, tq, nd, td, d, k, steps = [int(x) for x in (sys.argv[1:8] or "1000 32 12500 180 128 100 10".split())]
# under torch.distributed.run: nd documents PER RANK (config 5 = 8 ranks x 12 500), sharded_maxsim = local top-k with
# global document ids + one exchange + merge; time = max over ranks
world, rank, lrank = (int(os.environ.get(v, d0)) for v, d0 in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
dev = torch.device("cuda", lrank)
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    from neighborhoodwatch_b200 import dist as nwdist
    dist.init_process_group("nccl", device_id=dev)
q = synth.device_vectors(nq * tq, d, synth.QUERY_SEED, device=dev).view(nq, tq, d)
docs = synth.device_vectors(nd * td, d, synth.BASE_SEED, row0=rank * nd * td, device=dev).view(nd, td, d)
# signed token embeddings (the synthetic generator is all-positive): flip signs by a fixed pattern
sign = (torch.arange(d, device=dev) % 3 == 0).float() * -2 + 1
q = (q * sign).contiguous(); docs = (docs * sign.roll(1)).contiguous()
def step():
    if world > 1:
        return nwdist.sharded_maxsim(q, docs, k, rank * nd)
    return knn.maxsim_topk(q, docs, k)


for _ in range(3):
    s, i = step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    s, i = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
if world > 1:
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    dist.destroy_process_group()
    if rank != 0:
        sys.exit(0)
nd = nd * world
pairs = nq * tq * nd * td
print(json.dumps({"n_gpus": world, "workload": f"maxsim: {nq} queries x {tq} tokens vs {nd} documents x {td} tokens, d={d}, k={k}",
                  "ms_per_step": ms, "token_pairs_per_s": pairs / (ms * 1e-3), "tflops": 2 * pairs * d / (ms * 1e-3) / 1e12,
                  "stats": knn.last_stats}))
