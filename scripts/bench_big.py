"""Config 5: betweenness on a 4M-node power-law (BA) graph, sampled sources, vs oracle on a tiny sample."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import rustworkx_b200 as rx
from rustworkx_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
t0 = time.perf_counter(); g = rx.barabasi_albert_graph(n, 8, seed=5); print("gen s", time.perf_counter() - t0)
t0 = time.perf_counter(); dg = rx.device_snapshot(g); print("snapshot s", time.perf_counter() - t0)
src = np.random.default_rng(3).choice(n, size=S, replace=False).astype(np.uint32)
for rep in range(3):
    t0 = time.perf_counter(); out = dg.betweenness_partial(src, False); wall = time.perf_counter() - t0
    st = dg.stats()
    print(json.dumps({"n": n, "sources": S, "wall_s": round(wall, 3), "device_ms": round(st["ms_total"], 1),
                      "gteps": round(st["te"] / st["ms_total"] / 1e6, 1), "fwd": round(st["ms_forward"], 1),
                      "bwd": round(st["ms_backward"], 1), "batches": st["batches"], "levels": st["levels"]}), flush=True)
if os.environ.get("CHECK"):
    import oracle
    og = oracle.OracleGraph.from_graph(g)
    k = 16
    want, _ = og.betweenness_partial(src[:k], False, threads=oracle.max_threads())
    got = dg.betweenness_partial(src[:k], False)
    err = np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-12))
    print("max rel err vs oracle on", k, "sources:", err)
