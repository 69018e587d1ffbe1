"""Running search over resident slabs (the arithmetic of config 4 without the PCIe feed):
nwk_topk_state_search (bounded by the state's k-th distances) vs knn() + update() per slab."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neighborhoodwatch_b200 import knn, synth

nq, slab_rows, nslab, d, k, metric = [int(x) if x.isdigit() else x for x in (sys.argv[1:7] or "100000 1000000 4 1024 100 sqeuclidean".split())]
dev = torch.device("cuda", 0)
q = synth.device_vectors(nq, d, synth.QUERY_SEED, device=dev)
slabs = [synth.device_vectors(slab_rows, d, synth.BASE_SEED, row0=i * slab_rows, device=dev) for i in range(nslab)]
out = {}
for mode in ("update", "search", "update", "search"):
    st = knn.RunningTopK(nq, k, dev)
    per = []
    torch.cuda.synchronize()
    for i, b in enumerate(slabs):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if mode == "search":
            st.search(b, q, metric=metric, id_offset=i * slab_rows, profile=True)
        else:
            dd, ii = knn.knn(b, q, k, metric=metric, id_offset=i * slab_rows, profile=True)
            st.update(dd, ii)
        e1.record()
        torch.cuda.synchronize()
        s = knn.last_stats
        per.append({"ms": round(e0.elapsed_time(e1), 2), "filter": round(s["ms_filter"], 2), "finish": round(s["ms_finish"], 2),
                    "cand_per_q": round(s["total_candidates"] / nq, 1), "rerank_per_q": round(s["total_rerank"] / nq, 1),
                    "fallback": s["n_fallback_queries"]})
    rd, ri = st.result()
    torch.cuda.synchronize()
    st.close()
    out[mode] = {"total_ms": round(sum(p["ms"] for p in per), 2), "slabs": per}
    res = (rd, ri) if mode == "update" else res if "res" in dir() else (rd, ri)
    if mode == "search":
        same_d = bool(torch.equal(rd, res[0]))
        same_i = float((ri == res[1]).float().mean())
        out["search_equals_update"] = {"distances_identical": same_d, "ids_identical_fraction": same_i}
print(json.dumps({"shape": [nq, slab_rows, nslab, d, k, metric], **out}))
