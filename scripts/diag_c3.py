"""Diagnostic: one shard of config 3 (rows row0 .. row0+rows of the synthetic base), stats per variant."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from neighborhoodwatch_b200 import knn, synth
    row0, rows = int(sys.argv[2]), int(sys.argv[3])
    dev = torch.device("cuda", 0)
    b = synth.device_vectors(rows, 1536, synth.BASE_SEED, row0=row0, device=dev)
    q = synth.device_vectors(10000, 1536, synth.QUERY_SEED, device=dev)
    for _ in range(3):
        knn.knn(b, q, 100, metric="sqeuclidean", id_offset=row0, profile=True)
        st = knn.last_stats
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
else:
    for row0, rows in ((5_000_000, 5_000_000), (0, 5_000_000)):
        for env in ({"NWK_SAMPLE_M": "32", "NWK_PARTS": "1"}, {"NWK_SAMPLE_M": "16", "NWK_PARTS": "1"},
                    {"NWK_SAMPLE_M": "16", "NWK_PARTS": "3"}, {"NWK_SAMPLE_M": "32", "NWK_PARTS": "3"}):
            e = dict(os.environ)
            e.update(env)
            r = subprocess.run([sys.executable, __file__, "child", str(row0), str(rows)], env=e, capture_output=True, text=True)
            print(row0, env, r.stdout.strip()[-700:], r.stderr.strip()[-300:])
