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
        self.window = []
        self.load_end = float("inf")

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates in PCI order, CUDA need not: find this rank's GPU by its PCI address
            self.handle = None
            try:
                import torch
                pr = torch.cuda.get_device_properties(self.gpu_index)
                bdf = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
                self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(bdf.encode())
            except Exception:  # noqa: BLE001
                self.handle = None
            if self.handle is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = self.gpu_index
                if vis and all(tok.strip().isdigit() for tok in vis.split(",")):
                    idx = int(vis.split(",")[self.gpu_index])
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            # (the first NVML queries of a process take tens of milliseconds: pay that before the timed region)
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetPowerUsage(self.handle)
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
                self.samples.append((sm, pw, rs, time.perf_counter()))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.004)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Start (first call) / end (second call) of the timed region, for the per-window sample count."""
        self.window.append(time.perf_counter())

    def in_window(self) -> int:
        if len(self.window) < 2:
            return 0
        return sum(1 for s in self.samples if len(s) > 3 and self.window[0] <= s[3] <= self.window[1])

    def stop(self) -> dict:
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            nv = self.nvml
            names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            smp = self.samples
            if self.window:
                # under load: from 150 ms after the timed region began (NVML's readings lag) to the end of the load
                late = [x for x in smp if x[3] >= self.window[0] + 0.15 and x[3] <= self.load_end]
                smp = late or [x for x in smp if x[3] >= self.window[0]] or smp
            sm = [s[0] for s in smp]
            power = [s[1] for s in smp]
            reasons = sorted(n for n, bit in names.items() if any(s[2] & bit for s in smp))
            loaded = sm
            if power:
                cut = statistics.median(power)
                loaded = [c for c, pw in zip(sm, power) if pw >= cut] or sm
            return {"sm_mhz": statistics.median(loaded) if loaded else None, "sm_max_mhz": self.sm_max,
                    "power_w_max": max(power) if power else None, "samples": len(sm),
                    "samples_in_timed_region": self.in_window(), "reasons": reasons, "source": "nvml"}
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


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def blas_threads(limit=None):
    """(context manager that lifts the BLAS pools to `limit` threads, threads actually configured).
    torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the CPU arm must not inherit that."""
    import contextlib
    try:
        import threadpoolctl
        ctx = threadpoolctl.threadpool_limits(limits=limit) if limit else contextlib.nullcontext()
        return ctx, lambda: max([p.get("num_threads", 1) for p in threadpoolctl.threadpool_info()] or [1])
    except Exception:  # noqa: BLE001
        return contextlib.nullcontext(), lambda: None


def cpu_reference_run(q_count, b_rows, d, k, metric, steps, warmup, budget_rows=None):
    """The reference's CPU-equivalent exact brute force (numpy sgemm + argpartition), all host cores."""
    import numpy as np

    from neighborhoodwatch_b200 import synth
    from oracle import knn_oracle as ko
    cores = host_cores()
    ctx, used = blas_threads(cores)
    with ctx:
        res = _cpu_reference_run(np, synth, ko, cores, q_count, b_rows, d, k, metric, steps, warmup, budget_rows)
        n_used = used()
    if n_used:
        res["cores"] = int(n_used)          # the threads the BLAS pool really ran with
    return res


def _cpu_reference_run(np, synth, ko, cores, q_count, b_rows, d, k, metric, steps, warmup, budget_rows):
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


class ShardedRun:
    """One workload resident on this rank: a base shard + the replicated queries.  step() = local
    tcgen05 top-k written straight into the exchange chunk, ONE packed all-gather, merge kernel."""

    def __init__(self, torch, dev, nq, nb_local, lo, d, k, metric, distributed):
        from neighborhoodwatch_b200 import synth
        self.torch, self.dev, self.distributed = torch, dev, distributed
        self.nq, self.nb_local, self.lo, self.d, self.k, self.metric = nq, nb_local, lo, d, k, metric
        self.base = synth.device_vectors(nb_local, d, synth.BASE_SEED, row0=lo, device=dev)
        self.queries = synth.device_vectors(nq, d, synth.QUERY_SEED, device=dev)
        torch.cuda.synchronize(dev)
        self.reset()

    def reset(self):
        self.stage_ms = {"prep": [], "sample": [], "filter": [], "finish": [], "fallback": [], "total": []}
        self.agg = {"fallback": 0, "retry": 0, "max_candidates": 0, "max_rerank": 0, "cand": 0, "rerank": 0, "err_ratio": 0.0}
        self.exch = []

    def _local(self, b, q, kk, m, off, out=None, **kw):
        from neighborhoodwatch_b200 import knn
        res = knn.knn(b, q, kk, metric=m, id_offset=off, profile=True, out=out, **kw)
        st = knn.last_stats
        for name in self.stage_ms:
            self.stage_ms[name].append(st[f"ms_{name}"])
        a = self.agg
        a["fallback"] += st["n_fallback_queries"]
        a["retry"] += st.get("n_retry_queries", 0)
        a["max_candidates"] = max(a["max_candidates"], st["max_candidates"])
        a["max_rerank"] = max(a["max_rerank"], st["max_rerank"])
        a["err_ratio"] = max(a["err_ratio"], st["max_err_ratio"])
        a["cand"] += st["total_candidates"]
        a["rerank"] += st["total_rerank"]
        e = self.torch.cuda.Event(enable_timing=True)
        e.record()                       # local top-k done; what follows is exchange + merge
        self.exch.append([e])
        return res

    def step(self, queries=None, **kw):
        from neighborhoodwatch_b200 import dist as nwdist
        q = self.queries if queries is None else queries

        def local(b, qq, kk, m, off, out=None):
            return self._local(b, qq, kk, m, off, out=out, **kw)
        res = nwdist.sharded_knn(self.base, q, self.k, self.metric, self.lo, local_knn=local)
        if self.distributed:
            e = self.torch.cuda.Event(enable_timing=True)
            e.record()
            self.exch[-1].append(e)
        return res

    def barrier(self):
        if self.distributed:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed(self, steps, warmup, sampler=None):
        """W warm-up steps, two settle steps after the rendezvous, then exactly `steps` timed steps
        between barriers; CUDA-event time, max over ranks."""
        import torch.distributed as dist
        torch = self.torch
        from neighborhoodwatch_b200 import _native
        for _ in range(warmup):
            out = self.step()
        self.barrier()
        if sampler:
            sampler.start()
        # (the GPUs idle through the barrier and the NVML set-up: a 10-step multi-GPU run was once
        #  timed at +25 % because its first steps ran on clocks that had dropped)
        for _ in range(2):
            self.step()
        self.reset()
        self.barrier()
        launches0 = _native.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler:
            sampler.mark()
        ev0.record()
        for _ in range(steps):
            out = self.step()
        ev1.record()
        self.barrier()
        if sampler:
            sampler.mark()
        launches = _native.launch_count() - launches0
        stage_snap = {n: list(v) for n, v in self.stage_ms.items()}
        agg_snap = dict(self.agg)
        exch_snap = list(self.exch)
        # NVML's clock / power readings lag the load by ~0.1 s: a timed region shorter than 0.6 s is followed by
        # the same steps (untimed) until the sampler has seen 0.6 s of load, and the reported clock is the
        # median of the samples taken from 150 ms into that window
        short = bool(sampler) and (sampler.window[1] - sampler.window[0]) < 0.6
        extra = torch.tensor([1 if short else 0], device=self.dev)
        if self.distributed:
            dist.broadcast(extra, src=0)
        if int(extra.item()):
            t_end = (sampler.window[0] + 0.6) if sampler else (time.perf_counter() + 0.5)
            while True:
                self.step()
                go = torch.tensor([1 if time.perf_counter() < t_end else 0], device=self.dev)
                if self.distributed:
                    dist.broadcast(go, src=0)
                if not int(go.item()):
                    break
            self.torch.cuda.synchronize(self.dev)
        if sampler:
            sampler.load_end = time.perf_counter()
        clocks = sampler.stop() if sampler else None
        if clocks is not None and int(extra.item()):
            clocks["note"] = ("timed region shorter than 0.6 s: the same steps were repeated (untimed) until 0.6 s of load had "
                              "been sampled; sm_mhz = median of the samples from 150 ms into the load")
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=self.dev, dtype=torch.float64)
        lt = torch.tensor([launches], device=self.dev, dtype=torch.int64)
        local_ms = statistics.mean(stage_snap["total"]) if stage_snap["total"] else 0.0
        lmax = torch.tensor([local_ms], device=self.dev, dtype=torch.float64)
        lmin = lmax.clone()
        if self.distributed:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(lt, op=dist.ReduceOp.SUM)
            dist.all_reduce(lmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(lmin, op=dist.ReduceOp.MIN)
        exchange_ms = None
        if self.distributed:
            pairs_ev = [ev for ev in exch_snap if len(ev) == 2]
            if pairs_ev:
                exchange_ms = statistics.mean(a.elapsed_time(b) for a, b in pairs_ev)
        from neighborhoodwatch_b200 import dist as nwdist
        return {"out": out, "elapsed_s": float(ms.item()) / 1e3, "launches": int(lt.item()), "clocks": clocks,
                "exchange": (nwdist.last_exchange if self.distributed else None) if not nwdist._peer_failed
                else f"{nwdist.last_exchange} (peer memory unavailable: {nwdist._peer_failed})",
                "exchange_ms": exchange_ms, "local_ms_max": float(lmax.item()), "local_ms_min": float(lmin.item()),
                "stage_ms_mean": {n: (statistics.mean(v) if v else None) for n, v in stage_snap.items()},
                "agg": agg_snap}

    def free(self):
        from neighborhoodwatch_b200 import dist as nwdist
        from neighborhoodwatch_b200 import knn
        self.base = self.queries = None
        nwdist._xbuf.clear()
        knn.release_workspace()
        self.torch.cuda.empty_cache()


def torch_engine_baseline(torch, base, queries, k, ours_idx, batch_rows=100_000, steps=3):
    """The reference's own GPU arithmetic where RAPIDS is absent: its torch engine (cu_knn.py:253-260:
    ``1 - q @ b.T`` then ``topk(k, largest=False)``, ``+ batch_offset`` :273) restated with the image's
    torch / cuBLAS in the reference's batch shape (100 000 base rows per batch, its default
    ``initial_batch_size``, all queries per call), partial lists folded on the device (the reference
    folds them on the CPU, merge.py:73-101 -- not charged here).  Library kernels: an ungraded
    comparator, never part of the product path."""
    nb = base.shape[0]

    def one_pass():
        parts_d, parts_i = [], []
        for lo in range(0, nb, batch_rows):
            dataset = base[lo:lo + batch_rows]
            distances = 1 - torch.matmul(queries, dataset.T)
            dd, ii = torch.topk(distances, min(k, dataset.shape[0]), largest=False)
            parts_d.append(dd)
            parts_i.append(ii + lo)
            del distances
        cat_d, cat_i = torch.cat(parts_d, 1), torch.cat(parts_i, 1)
        dd, sel = torch.topk(cat_d, k, largest=False)
        return dd, torch.gather(cat_i, 1, sel)

    out = {}
    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            one_pass()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                dd, ii = one_pass()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            agree = float((torch.sort(ii, dim=1).values == torch.sort(ours_idx, dim=1).values).float().mean())
            out[name] = {"ms_per_step": ms, "value": queries.shape[0] * nb / (ms * 1e-3), "unit": "pairs/s",
                         "ids_equal_to_ours": agree}
            del dd, ii
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return {"kind": "reference torch engine restated (cu_knn.py:253-260) on this GPU: torch.matmul (cuBLAS) + torch.topk",
            "batch_rows": batch_rows, "steps": steps, "value": out["fp32"]["value"], "unit": "pairs/s",
            "ms_per_step": out["fp32"]["ms_per_step"], "allow_tf32": False, "fp32": out["fp32"], "tf32": out["tf32"],
            "note": "1 - q.b ranks unit vectors like cosine / squared L2; torch's default fp32 matmul (allow_tf32=False) "
                    "is the reference's setting, the tf32 line is the same code with allow_tf32=True"}


def run_e2e(torch, run, args, n_gpus, local_rank):
    """The same metric through the host-buffer C-ABI entry point: base + queries start in pinned HOST
    memory every step, results end in host memory."""
    import torch.distributed as dist

    from neighborhoodwatch_b200 import dist as nwdist
    from neighborhoodwatch_b200 import knn
    dev, nq, nb_local, d, k, metric, lo = run.dev, run.nq, run.nb_local, run.d, run.k, run.metric, run.lo
    distributed = run.distributed
    ref_d, ref_i = run.step()
    base_host = torch.empty((nb_local, d), dtype=torch.float32, pin_memory=True)
    q_host = torch.empty((nq, d), dtype=torch.float32, pin_memory=True)
    base_host.copy_(run.base)
    q_host.copy_(run.queries)
    out_d_host = torch.empty((nq, k), dtype=torch.float32, pin_memory=True)
    out_i_host = torch.empty((nq, k), dtype=torch.int64, pin_memory=True)
    torch.cuda.synchronize(dev)
    if distributed:
        sbuf, gbuf = nwdist.exchange_buffers(nq, k, n_gpus, dev)
        s_d, s_i = nwdist._chunk_views(sbuf, nq, k)

    def e2e_step():
        d_h, i_h = knn.knn_host(base_host, q_host, k, metric=metric, id_offset=lo,
                                out=(out_d_host.numpy(), out_i_host.numpy()))
        if not distributed:
            return d_h, i_h
        s_d.copy_(out_d_host, non_blocking=True)
        s_i.copy_(out_i_host, non_blocking=True)
        dist.all_gather_into_tensor(gbuf, sbuf)
        m_d, m_i = knn.merge_topk_packed(gbuf, n_gpus, nq, k, k)
        return m_d.cpu(), m_i.cpu()

    e2e_step()
    e2e_step()
    run.barrier()
    # every step is timed on its own (max over ranks); the reported value uses the MEDIAN step and
    # all step times are listed next to it
    step_s = []
    for _ in range(args.e2e_steps):
        run.barrier()
        t0 = time.perf_counter()
        e_d, e_i = e2e_step()
        torch.cuda.synchronize(dev)
        t_step = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if distributed:
            dist.all_reduce(t_step, op=dist.ReduceOp.MAX)
        step_s.append(float(t_step.item()))
    t_med = statistics.median(step_s)
    # slab results are merged by (fp32 distance, id), the resident result is ordered by
    # (fp64 distance, id): compare distances exactly and ids as per-row sets
    e_i_t = torch.as_tensor(e_i).to(dev)
    e_d_t = torch.as_tensor(e_d).to(dev)
    same_d = bool(torch.equal(e_d_t, ref_d))
    rows_same = float((torch.sort(e_i_t, dim=1).values == torch.sort(ref_i, dim=1).values).all(dim=1).float().mean())
    st = knn.last_stats or {}
    del base_host
    nb_total = nb_local * n_gpus
    return {"value": nq * nb_total / t_med, "unit": "pairs/s",
            "h2d_bytes_per_step": (nb_local * d * 4 + nq * d * 4) * n_gpus,
            "d2h_bytes_per_step": nq * k * 12 * n_gpus, "steps": args.e2e_steps,
            "ms_per_step": 1e3 * t_med, "timing": "median of the listed steps, each timed wall-clock incl. H2D + D2H",
            "step_ms": [round(1e3 * x, 3) for x in step_s], "step_ms_max_over_median": max(step_s) / t_med,
            "api": "nwk_knn_host_f32 (pinned host base+queries in, host results out; slabs double-buffered, "
                   "staging ring / workspace / running state cached per device)",
            "matches_resident_result": same_d and rows_same > 0.99, "distances_identical": same_d,
            "rows_with_identical_id_sets": rows_same, "gpu_launches_per_step": st.get("gpu_launches"),
            "h2d_gb_per_s_per_gpu": (nb_local * d * 4 + nq * d * 4) / t_med / 1e9}


def run_config3(torch, dev, n_gpus, rank, distributed, peaks, steps=5, warmup=3):
    """BASELINE config 3 (the north-star target): 10 000 queries x 10 000 000 base x d=1536, squared L2,
    k=100, the base STRONG-sharded over the N GPUs of this run (10^7 / N rows each), recorded next to
    the weak-scaling headline.  Verified in-run: 32 sampled queries against the sharded fp64 scan."""
    from neighborhoodwatch_b200 import dist as nwdist
    nq, nb_total, d, k, metric = 10_000, 10_000_000, 1536, 100, "sqeuclidean"
    lo, hi = nwdist.shard_bounds(nb_total, n_gpus, rank)
    run = ShardedRun(torch, dev, nq, hi - lo, lo, d, k, metric, distributed)
    r = run.timed(steps, warmup)
    out_d, out_i = r["out"]
    sel = torch.arange(0, nq, nq // 32, device=dev)[:32]
    ex_d, ex_i = run.step(queries=run.queries[sel].contiguous(), force_exact_scan=True)
    same = bool(torch.equal(ex_d, out_d[sel])) and bool(torch.equal(ex_i, out_i[sel]))
    run.free()
    ms = 1e3 * r["elapsed_s"] / steps
    flops = 2.0 * nq * nb_total * d
    tf = flops / (ms * 1e-3) / 1e12
    filt = r["stage_ms_mean"]["filter"]
    return {"workload": f"c3: {nq} queries x {nb_total} base x d={d}, k={k}, {metric}; base strong-sharded over "
                        f"{n_gpus} GPU(s) ({hi - lo} rows on rank {rank})",
            "scaling": "strong", "steps": steps, "warmup": warmup, "ms_per_step": ms,
            "pairs_per_s": nq * nb_total / (ms * 1e-3), "queries_per_s": nq / (ms * 1e-3),
            "aggregate_tflops": tf, "frac_of_n_x_burst": tf / (n_gpus * peaks["bf16_burst"]),
            "frac_of_n_x_sustained": tf / (n_gpus * peaks["bf16_sustained"]),
            "filter_ms": filt, "filter_tflops_per_gpu": 2.0 * nq * (hi - lo) * d / (filt * 1e-3) / 1e12 if filt else None,
            "exchange_ms": r["exchange_ms"], "exchange": r.get("exchange"), "local_ms_max_over_ranks": r["local_ms_max"],
            "local_ms_min_over_ranks": r["local_ms_min"], "stage_ms_mean": r["stage_ms_mean"],
            "gpu_launches": r["launches"], "fallback_queries_per_step": r["agg"]["fallback"] / steps,
            "second_chance_queries_per_step": r["agg"]["retry"] / steps,
            "sampled_queries_equal_fp64_scan": same}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c4"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config3", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--config4-min-gpus", type=int, default=8, help="run the config-4 leg from this many GPUs on (tests: 2)")
    ap.add_argument("--c4-queries", type=int, default=100_000)
    ap.add_argument("--c4-rows-per-gpu", type=int, default=12_500_000)
    ap.add_argument("--c4-slab-rows", type=int, default=1_000_000)
    ap.add_argument("--c4-host-wc", action="store_true", help="keep the pinned host shard in write-combined memory")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, 1)
    if args.workload == "c4":
        from scripts import bench_c4
        return bench_c4.main(args, n_gpus, world, rank, local_rank, load_peaks(), ClockSampler)
    nq, nb_local, d, k, metric = WORKLOADS[args.workload]
    nb_total = nb_local * n_gpus
    config = {"workload": f"{args.workload}: {nq} queries x {nb_local} base rows per GPU x d={d}, k={k}, {metric}",
              "queries": nq, "base_rows_total": nb_total, "base_rows_per_gpu": nb_local, "d": d, "k": k,
              "metric": metric, "sharding": f"base rows over {n_gpus} GPU(s), gather + k-way merge in one kernel over peer memory (NVLink)" if n_gpus > 1 else "single GPU",
              "l2": "inputs_exceed_l2 (base shard %.2f GB vs 126 MB L2)" % (nb_local * d * 4 / 1e9)}

    if args.impl == "reference":
        if rank != 0:
            return
        # torch.distributed.run exports OMP_NUM_THREADS=1; numpy is not imported yet
        for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[var] = str(host_cores())
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
                        "code path; this is the numpy restatement of its engines' arithmetic on the host cores "
                        "(BLAS threads set explicitly: torchrun's OMP_NUM_THREADS=1 is not inherited)"}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    from neighborhoodwatch_b200 import _native, knn  # noqa: F401

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)
        assert world == n_gpus, f"--gpus {n_gpus} but WORLD_SIZE={world}"
    peaks = load_peaks()
    run = ShardedRun(torch, dev, nq, nb_local, rank * nb_local, d, k, metric, distributed)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    r = run.timed(args.steps, max(args.warmup, 3), sampler)
    out_d, out_i = r["out"]
    elapsed_s = r["elapsed_s"]
    value = nq * nb_total * args.steps / elapsed_s

    e2e = None if args.no_e2e else run_e2e(torch, run, args, n_gpus, local_rank)
    gpu_baseline = None
    if n_gpus == 1 and not args.no_gpu_baseline:
        try:
            gpu_baseline = torch_engine_baseline(torch, run.base, run.queries, k, out_i)
        except Exception as e:  # noqa: BLE001 -- a side leg must not take the bench down
            gpu_baseline = {"unavailable": repr(e)[:300]}
    run.free()
    _native.load().nwk_host_release()
    config3 = None
    if not args.no_config3 and args.workload == "c2":
        try:
            config3 = run_config3(torch, dev, n_gpus, rank, distributed, peaks)
        except Exception as e:  # noqa: BLE001
            if distributed:
                raise
            config3 = {"unavailable": repr(e)[:300]}
    # BASELINE config 4 (streamed from pinned host memory across 8 B200) rides along in the 8-GPU line
    config4 = None
    if n_gpus >= args.config4_min_gpus and distributed and args.workload == "c2" and not args.no_config4:
        from scripts import bench_c4
        knn.release_workspace()
        torch.cuda.empty_cache()
        c4 = bench_c4.measure(args, n_gpus, world, rank, local_rank, dev, distributed, peaks, ClockSampler, steps=2)
        if rank == 0 and c4 is not None:
            config4 = {key: c4.get(key) for key in ("config", "e2e", "resident", "fill_seconds", "clocks", "unavailable")
                       if c4.get(key) is not None}
    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    filt = r["stage_ms_mean"]["filter"] or float("nan")
    flops = 2.0 * nq * nb_local * d
    achieved = flops / (filt * 1e-3) / 1e12 if filt == filt and filt > 0 else None
    peak = peaks["bf16_sustained"]
    roofline = {"bound": "tensor", "kernel": "tc_kernel<FILTER, fp16> (tcgen05 kind::f16 + fused threshold filter)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved else None,
                "frac_of_burst": (achieved / peaks["bf16_burst"]) if achieved else None,
                "whole_step_frac_of_burst": flops / (1e-3 * 1e3 * elapsed_s / args.steps) / 1e12 / peaks["bf16_burst"],
                "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a step; fp16 and bf16 "
                               f"share the tensor-pipe rate); burst = {peaks['bf16_burst']}",
                "flop_per_launch": flops, "ms_per_launch": filt, "traffic": None,
                "launch_note": "the filter pass of a step is up to three launches of this kernel (parts of the base, on two "
                               "streams, each behind its part of the preparation); flop_per_launch / ms_per_launch are the "
                               "step's total FLOPs and the CUDA-event time from the first launch to the join of the last",
                "stage_ms_mean": r["stage_ms_mean"]}
    for tname in ("r02_traffic.json", "r01_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            with open(tpath) as fh:
                tr = json.load(fh).get(args.workload)
            if tr:   # measured once with `ncu --set full` on this workload (profiles/), per launch
                roofline["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
                roofline["traffic_source"] = tr["source"]
                roofline["algorithmic_dram_bytes"] = nb_local * d * 2 + nq * d * 2 + nb_local * 8
                break
    agg = r["agg"]
    line = {"metric": "query_base_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * elapsed_s / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands -> f32 accumulate (candidates) + f64 (rerank)",
            "data": "synthetic", "config": config, "queries_per_s": nq * args.steps / elapsed_s,
            "clocks": r["clocks"], "e2e": e2e, "gpu_launches": r["launches"], "roofline": roofline,
            # exchange = one packed all-gather of the [Q,k] lists + the merge kernel, from the moment this rank's
            # local top-k is done (so it includes waiting for the slowest rank: the local_ms spread says how much)
            "exchange_ms_per_step": r["exchange_ms"], "exchange": r.get("exchange"),
            "local_knn_ms": {"max_over_ranks": r["local_ms_max"], "min_over_ranks": r["local_ms_min"]},
            "exactness": {"fallback_queries_per_step": agg["fallback"] / max(args.steps, 1),
                          "second_chance_queries_per_step": agg["retry"] / max(args.steps, 1),
                          "mean_candidates_per_query": agg["cand"] / max(args.steps * nq, 1),
                          "mean_rerank_per_query": agg["rerank"] / max(args.steps * nq, 1),
                          "max_candidates": agg["max_candidates"], "max_rerank": agg["max_rerank"],
                          "max_err_over_proven_bound": agg["err_ratio"]},
            "config3": config3, "config4": config4, "gpu_baseline": gpu_baseline}
    if n_gpus == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_run(nq, nb_total, d, k, metric, 1, 1)
        line["cpu_baseline"] = {key: cb[key] for key in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
