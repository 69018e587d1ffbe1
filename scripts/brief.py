"""Print the few numbers of a bench.py JSON line that matter when comparing builds."""
import json
import sys

for path in sys.argv[1:]:
    try:
        line = [l for l in open(path) if l.startswith("{")][-1]
        j = json.loads(line)
    except Exception as e:  # noqa: BLE001
        print(path, "unreadable:", e)
        continue
    r = j.get("roofline") or {}
    st = {k: round(v, 3) for k, v in (r.get("stage_ms_mean") or {}).items() if v is not None}
    ex = j.get("exactness") or {}
    print(f"{path}: ms/step {j.get('ms_per_step'):.3f} filter TF/s {r.get('achieved') or 0:.0f} stages {st} "
          f"fallback {ex.get('fallback_queries_per_step')} retry {ex.get('second_chance_queries_per_step')} cand {ex.get('mean_candidates_per_query')} "
          f"rerank {ex.get('mean_rerank_per_query')} clocks {(j.get('clocks') or {}).get('sm_mhz')}")
