"""Wall clock of rxc_betweenness_source_order on the bench graph (host sort + two small kernels)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rustworkx_b200 as rx
g = rx.barabasi_albert_graph(1_000_000, 8, seed=1)
dg = rx.device_snapshot(g)
plan = np.random.default_rng(12345).permutation(g.num_nodes()).astype(np.uint32)
for k in (1024, 65536, 1_000_000):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); o = dg.betweenness_source_order(plan[:k]); best = min(best, time.perf_counter() - t0)
    assert sorted(o.tolist()) == sorted(plan[:k].tolist())
    print(f"source_order k={k}: {best*1e3:.1f} ms", flush=True)
