"""One Newman weighted closeness batch on BA(1M, 8) for ncu: python scripts/ncu_wsp.py [sources]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rustworkx_b200 as rx
n = int(os.environ.get("N", 1_000_000)); S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = rx.barabasi_albert_graph(n, 8, seed=1)
dg = rx.device_snapshot(g)
rng = np.random.default_rng(7)
sources = rng.choice(n, size=S, replace=False).astype(np.uint32)
_, eb = g._bounds()
w = rng.uniform(0.05, 2.0, size=eb)
for rep in range(int(os.environ.get("REPS", 1))):
    dg.weighted_closeness_partial(w, sources)
    st = dg.stats()
    print({k: st[k] for k in ("ms_total", "ms_forward", "ms_backward", "levels", "launches", "te", "te_dag", "v")})
