#!/usr/bin/env python
"""bench.py -- exact top-100 brute-force kNN throughput (query x base pairs / s).

    python bench.py --gpus N --steps K --warmup W            # our arm (libnwk, B200)
    python bench.py --impl reference ...                      # the reference's CPU-equivalent path

A "step" is one pass of the hot path over one batch of synthetic input: exact top-k of
every query against the whole base.  Workload (BASELINE.json configs[1], "c2"): 10 000
queries x 1 000 000 base rows per GPU, d=768, k=100, cosine, synthetic unit vectors in the
reference's fixture style.  With N GPUs the base is N x 1M rows, row-sharded (weak scaling),
each rank computes its exact local top-k, ONE all-gather exchanges the [Q,k] lists and a
k-way merge kernel reduces them (neighborhoodwatch_b200/dist.py).

JSON keys beyond the base contract:
  value        pairs/s with inputs resident in HBM, CUDA-event time of K steps, max over ranks
  e2e          the same metric through the host-buffer C-ABI entry point (nwk_knn_host_f32):
               base + queries start in pinned HOST memory every step, results end in host memory
  roofline     dominant kernel = tc_kernel<FILTER> (tcgen05 fp16 GEMM + fused threshold filter);
               achieved = 2*Q*B*d FLOP per launch / its CUDA-event duration (measured inside
               the library on the launching stream, NWK_FLAG_PROFILE)
  cpu_baseline numpy fp32 brute force in the reference's arithmetic (oracle/knn_oracle.py,
               kind "port": RAPIDS is not installable here and the reference has no CPU path of
               its own) on a bounded sample, rank 0, N=1 only
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (queries, base rows per GPU, d, k, metric)
    "c1": (1000, 10_000, 384, 100, "dot"),
    "c2": (10_000, 1_000_000, 768, 100, "cosine"),
    "c3shard": (10_000, 1_250_000, 1536, 100, "sqeuclidean"),     # one GPU's share of config 3 at W=8
    "c5shard": (32_000, 2_250_000, 128, 100, "dot"),              # one GPU's share of config 5 at W=8
    "c4slab": (100_000, 1_000_000, 1024, 100, "sqeuclidean"),     # one streamed slab of config 4 (12.5 such slabs per GPU at W=8)
    "c5l2": (32_000, 2_250_000, 128, 100, "sqeuclidean"),         # the config-5 shard shape under the reference's default metric
    "tiny": (512, 50_000, 128, 100, "cosine"),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock, power and throttle reasons during the timed region, polled through NVML every
    few milliseconds (the timed region of a short run is shorter than nvidia-smi's loop period);
    falls back to `nvidia-smi -lms` (the recipe in B200_PROFILING.md) when pynvml is missing."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []
        self.thread = None
        self.samples = []          # (sm_mhz, power_w, reasons bitmask)
        self.stop_flag = threading.Event()
        self.nvml = None
        self.sm_max = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical order; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu_index
            if vis and all(tok.strip().isdigit() for tok in vis.split(",")):
                idx = int(vis.split(",")[self.gpu_index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001 -- fall back to the CLI
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:  # noqa: BLE001
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((sm, pw, rs))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.004)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            nv = self.nvml
            names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm = [s[0] for s in self.samples]
            power = [s[1] for s in self.samples]
            reasons = sorted(n for n, bit in names.items() if any(s[2] & bit for s in self.samples))
            loaded = sm
            if power:
                cut = statistics.median(power)
                loaded = [c for c, pw in zip(sm, power) if pw >= cut] or sm
            return {"sm_mhz": statistics.median(loaded) if loaded else None, "sm_max_mhz": self.sm_max,
                    "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": reasons,
                    "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # samples under load: the upper half by power draw
        loaded = sm
        if power:
            cut = statistics.median(power)
            loaded = [s for s, p in zip(sm, power) if p >= cut] or sm
        return {"sm_mhz": statistics.median(loaded) if loaded else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def gpu_local_cpus(torch, dev_index):
    """CPUs on the NUMA node the GPU hangs off (sysfs local_cpulist), or None."""
    try:
        props = torch.cuda.get_device_properties(dev_index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as fh:
            text = fh.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        return cpus or None
    except Exception:  # noqa: BLE001 -- best effort
        return None


def cpu_reference_run(q_count, b_rows, d, k, metric, steps, warmup, budget_rows=None):
    """The reference's CPU-equivalent exact brute force (numpy sgemm + argpartition), all host cores."""
    import numpy as np

    from neighborhoodwatch_b200 import synth
    from oracle import knn_oracle as ko
    cores = os.cpu_count() or 1
    if budget_rows is None:
        # ~10-20 s of CPU work per measurement: ~2e10 multiply-adds per core-second is optimistic,
        # so bound the sample at 1e9 pairs (= 100 000 base rows at Q = 10 000)
        budget_rows = max(k, min(b_rows, int(1e9 // max(q_count, 1))))
    base = synth.vectors(budget_rows, d, synth.BASE_SEED)
    q = synth.vectors(q_count, d, synth.QUERY_SEED)
    for _ in range(warmup):
        ko.knn_f32_reference_form(base[: max(k, budget_rows // 8)], q, k, metric)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ko.knn_f32_reference_form(base, q, k, metric)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    pairs = q_count * budget_rows * steps
    return {"value": pairs / total, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"{q_count} queries x {budget_rows} base rows x d={d}, {metric}, k={k}; numpy fp32 "
                      f"sgemm + argpartition (oracle.knn_f32_reference_form), {steps} timed pass(es)",
            "ms_per_step": 1e3 * total / steps, "queries_per_s_at_full_base": (pairs / total) / b_rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    nq, nb_local, d, k, metric = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, 1)
    nb_total = nb_local * n_gpus
    config = {"workload": f"{args.workload}: {nq} queries x {nb_local} base rows per GPU x d={d}, k={k}, {metric}",
              "queries": nq, "base_rows_total": nb_total, "base_rows_per_gpu": nb_local, "d": d, "k": k,
              "metric": metric, "sharding": f"base rows over {n_gpus} GPU(s), all-gather + k-way merge" if n_gpus > 1 else "single GPU",
              "l2": "inputs_exceed_l2 (base shard %.2f GB vs 126 MB L2)" % (nb_local * d * 4 / 1e9)}

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 3))
        res = cpu_reference_run(nq, nb_total, d, k, metric, steps, min(args.warmup, 1))
        line = {"impl": "reference", "metric": "query_base_pairs_per_s", "value": res["value"], "unit": "pairs/s",
                "n_gpus": n_gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "queries_per_s": res["queries_per_s_at_full_base"],
                "cpu_baseline": {"value": res["value"], "unit": "pairs/s", "cores": res["cores"], "kind": "port",
                                 "sample": res["sample"]},
                "e2e": {"value": res["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "RAPIDS (pylibraft/cuvs/cudf) is not installable offline and the reference has no CPU "
                        "code path; this is the numpy restatement of its engines' arithmetic on the host cores"}
        print(json.dumps(line))
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from neighborhoodwatch_b200 import _native, knn, synth
    from neighborhoodwatch_b200 import dist as nwdist

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)
        assert world == n_gpus, f"--gpus {n_gpus} but WORLD_SIZE={world}"
    lo = rank * nb_local
    base = synth.device_vectors(nb_local, d, synth.BASE_SEED, row0=lo, device=dev)
    queries = synth.device_vectors(nq, d, synth.QUERY_SEED, device=dev)
    torch.cuda.synchronize(dev)
    stage_ms = {"prep": [], "sample": [], "filter": [], "finish": [], "fallback": [], "total": []}
    agg = {"fallback": 0, "max_candidates": 0, "max_rerank": 0, "cand": 0, "rerank": 0, "err_ratio": 0.0}

    def local_knn(b, q, kk, m, off):
        out = knn.knn(b, q, kk, metric=m, id_offset=off, profile=True)
        st = knn.last_stats
        for name in stage_ms:
            stage_ms[name].append(st[f"ms_{name}"])
        agg["fallback"] += st["n_fallback_queries"]
        agg["max_candidates"] = max(agg["max_candidates"], st["max_candidates"])
        agg["max_rerank"] = max(agg["max_rerank"], st["max_rerank"])
        agg["err_ratio"] = max(agg["err_ratio"], st["max_err_ratio"])
        agg["cand"] += st["total_candidates"]
        agg["rerank"] += st["total_rerank"]
        return out

    exch_events = []

    def merge_timed(all_d, all_i, kk):
        out = knn.merge_topk(all_d, all_i, kk)
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        exch_events[-1].append(e)
        return out

    def local_knn_timed(b, q, kk, m, off):
        out = local_knn(b, q, kk, m, off)
        e = torch.cuda.Event(enable_timing=True)
        e.record()                       # local top-k done; what follows is all-gather + merge
        exch_events.append([e])
        return out

    def step():
        return nwdist.sharded_knn(base, queries, k, metric, lo, local_knn=local_knn_timed, merge=merge_timed)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        out_d, out_i = step()
    for v in stage_ms.values():
        v.clear()
    for key in agg:
        agg[key] = 0
    exch_events.clear()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # two more untimed steps after the rendezvous (the GPUs idled through the barrier and the NVML
    # set-up: a 10-step multi-GPU run was once timed at +25 % because its first steps ran on
    # clocks that had dropped), then the timed region starts from a barrier again
    for _ in range(2):
        step()
    for v in stage_ms.values():
        v.clear()
    for key in agg:
        agg[key] = 0
    exch_events.clear()
    barrier()
    launches0 = _native.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out_d, out_i = step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = _native.launch_count() - launches0
    elapsed_ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    launches_t = torch.tensor([launches], device=dev, dtype=torch.int64)
    if distributed:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches_t, op=dist.ReduceOp.SUM)
    elapsed_s = float(elapsed_ms.item()) / 1e3
    value = nq * nb_total * args.steps / elapsed_s

    # ---- end to end through the host-buffer entry point -------------------------------------
    e2e = None
    numa = None
    if not args.no_e2e:
        # first-touch the pinned staging buffers from the GPU-local NUMA node: host->device copies
        # from the far socket run at roughly half the PCIe rate
        all_cpus = os.sched_getaffinity(0)
        local = gpu_local_cpus(torch, local_rank)
        if local:
            os.sched_setaffinity(0, local)
            numa = {"gpu_local_cpus": len(local), "of": len(all_cpus)}
        base_host = torch.empty((nb_local, d), dtype=torch.float32, pin_memory=True)
        q_host = torch.empty((nq, d), dtype=torch.float32, pin_memory=True)
        base_host.copy_(base)
        q_host.copy_(queries)
        out_d_host = torch.empty((nq, k), dtype=torch.float32, pin_memory=True)
        out_i_host = torch.empty((nq, k), dtype=torch.int64, pin_memory=True)
        torch.cuda.synchronize(dev)

        def e2e_step():
            d_h, i_h = knn.knn_host(base_host, q_host, k, metric=metric, id_offset=lo,
                                    out=(out_d_host.numpy(), out_i_host.numpy()))
            if not distributed:
                return d_h, i_h
            all_d, all_i = nwdist.gather_partials(out_d_host.to(dev, non_blocking=True), out_i_host.to(dev, non_blocking=True))
            m_d, m_i = knn.merge_topk(all_d, all_i, k)
            return m_d.cpu(), m_i.cpu()

        e2e_step()
        e2e_step()
        barrier()
        # every step is timed on its own (max over ranks); the reported value uses the MEDIAN
        # step so that one burst of host / PCIe contention on a shared box does not decide it --
        # all step times are listed next to it
        step_s = []
        for _ in range(args.e2e_steps):
            barrier()
            t0 = time.perf_counter()
            e_d, e_i = e2e_step()
            torch.cuda.synchronize(dev)
            t_step = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if distributed:
                dist.all_reduce(t_step, op=dist.ReduceOp.MAX)
            step_s.append(float(t_step.item()))
        t_med = statistics.median(step_s)
        t_e2e = torch.tensor([t_med * args.e2e_steps], device=dev, dtype=torch.float64)
        # slab results are merged by (fp32 distance, id), the resident result is ordered by
        # (fp64 distance, id): compare distances exactly and ids as per-row sets
        e_i_t = torch.as_tensor(e_i).to(dev)
        e_d_t = torch.as_tensor(e_d).to(dev)
        # (a tie that straddles the k-th slot may legitimately pick a different row)
        same_d = bool(torch.equal(e_d_t, out_d))
        rows_same = float((torch.sort(e_i_t, dim=1).values == torch.sort(out_i, dim=1).values).all(dim=1).float().mean())
        same = same_d and rows_same > 0.99
        e2e = {"value": nq * nb_total * args.e2e_steps / float(t_e2e.item()), "unit": "pairs/s",
               "h2d_bytes_per_step": (nb_local * d * 4 + nq * d * 4) * n_gpus,
               "d2h_bytes_per_step": nq * k * 12 * n_gpus, "steps": args.e2e_steps,
               "ms_per_step": 1e3 * t_med, "timing": "median of the listed steps, each timed wall-clock incl. H2D + D2H",
               "step_ms": [round(1e3 * x, 3) for x in step_s],
               "api": "nwk_knn_host_f32 (pinned host base+queries in, host results out; slabs double-buffered)",
               "matches_resident_result": same, "distances_identical": same_d,
               "rows_with_identical_id_sets": rows_same, "numa": numa,
               "h2d_gb_per_s": (nb_local * d * 4 + nq * d * 4) * args.e2e_steps / float(t_e2e.item()) / 1e9}
        del base_host
        if local:
            os.sched_setaffinity(0, all_cpus)

    exchange_ms = None
    if distributed:
        pairs_ev = [ev for ev in exch_events if len(ev) == 2]
        if pairs_ev:
            exchange_ms = statistics.mean(a.elapsed_time(b) for a, b in pairs_ev)
    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    filt = statistics.mean(stage_ms["filter"]) if stage_ms["filter"] else float("nan")
    flops = 2.0 * nq * nb_local * d
    achieved = flops / (filt * 1e-3) / 1e12 if filt == filt and filt > 0 else None
    peak = peaks["bf16_sustained"]
    roofline = {"bound": "tensor", "kernel": "tc_kernel<FILTER, fp16> (tcgen05 kind::f16 + fused threshold filter)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved else None,
                "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a step; fp16 and bf16 "
                               f"share the tensor-pipe rate)",
                "flop_per_launch": flops, "ms_per_launch": filt, "traffic": None,
                "stage_ms_mean": {name: (statistics.mean(v) if v else None) for name, v in stage_ms.items()}}
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tr = json.load(fh).get(args.workload)
        if tr:   # measured once with `ncu --set full` on this workload (profiles/), per launch
            roofline["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            roofline["traffic_source"] = tr["source"]
            roofline["algorithmic_dram_bytes"] = nb_local * d * 2 + nq * d * 2 + nb_local * 8
    line = {"metric": "query_base_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * elapsed_s / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands -> f32 accumulate (candidates) + f64 (rerank)",
            "data": "synthetic", "config": config, "queries_per_s": nq * args.steps / elapsed_s,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_t.item()), "roofline": roofline,
            "exchange_ms_per_step": exchange_ms,   # all-gather of the [Q,k] lists + merge kernel (rank 0), N > 1 only
            "exactness": {"fallback_queries_per_step": agg["fallback"] / max(args.steps, 1),
                          "mean_candidates_per_query": agg["cand"] / max(args.steps * nq, 1),
                          "mean_rerank_per_query": agg["rerank"] / max(args.steps * nq, 1),
                          "max_candidates": agg["max_candidates"], "max_rerank": agg["max_rerank"],
                          "max_err_over_proven_bound": agg["err_ratio"]}}
    if n_gpus == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_run(nq, nb_total, d, k, metric, 1, 1)
        line["cpu_baseline"] = {key: cb[key] for key in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
