"""Host -> device copy rate with all ranks copying at once: default pinned memory vs write-combined.
Run under torch.distributed.run; rank 0 prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from neighborhoodwatch_b200 import knn

world, rank, lrank = (int(os.environ.get(v, d0)) for v, d0 in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
dev = torch.device("cuda", lrank)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rows, d, reps = 1_000_000, 1024, 6
dst = torch.empty((rows, d), dtype=torch.float32, device=dev)
out = {"n_gpus": world, "bytes_per_copy": rows * d * 4}
for name, wc in (("pinned", False), ("write_combined", True)):
    holder = knn.PinnedArray((rows, d), write_combined=wc)
    src = torch.from_numpy(holder.array)
    dst.copy_(src)              # touch
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)      # (torch asks the driver: cudaHostAlloc'ed memory counts as pinned)
    torch.cuda.synchronize()
    gbs = torch.tensor([rows * d * 4 * reps / (time.perf_counter() - t0) / 1e9], device=dev, dtype=torch.float64)
    lo, hi = gbs.clone(), gbs.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out[name + "_is_pinned"] = bool(src.is_pinned())
    out[name] = {"gb_per_s_per_gpu_min": float(lo.item()), "gb_per_s_per_gpu_max": float(hi.item())}
    del src
    holder.close()
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
