"""Closeness (bit-parallel regime) on the bench graph BA(1M, 8): python scripts/bench_closeness_ba.py [sources]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rustworkx_b200 as rx
n = int(os.environ.get("N", 1_000_000)); S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = rx.barabasi_albert_graph(n, 8, seed=1)
dg = rx.device_snapshot(g)
src = np.random.default_rng(5).choice(n, size=S, replace=False).astype(np.uint32)
for rep in range(3):
    t0 = time.perf_counter(); out = dg.closeness_partial(src, True); wall = time.perf_counter() - t0
    st = dg.stats()
    print(json.dumps({"case": f"closeness BA({n},8), {S} sources", "wall_ms": round(wall * 1e3, 1),
                      "device_ms": round(st["ms_total"], 1), "gteps": round(st["te"] / st["ms_total"] / 1e6, 1),
                      "levels": st["levels"], "launches": st["launches"], "batches": st["batches"]}), flush=True)
