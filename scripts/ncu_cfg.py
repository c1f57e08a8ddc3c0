"""One call of another BASELINE config for ncu:  python scripts/ncu_cfg.py grid|hex|gnm_edge|ba4m"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rustworkx_b200 as rx
which = sys.argv[1]
if which == "grid":      # config 2: closeness on grid_graph(1000, 1000), one wave of sources
    g = rx.generators.grid_graph(1000, 1000); dg = rx.device_snapshot(g)
    src = np.random.default_rng(1).choice(g.num_nodes(), size=592, replace=False).astype(np.uint32)
    dg.closeness_partial(src, True)
elif which == "hex":     # config 2: closeness on heavy_hex_graph(51), all sources
    g = rx.generators.heavy_hex_graph(51); dg = rx.device_snapshot(g)
    dg.closeness(True)
elif which == "gnm_edge":  # config 4: edge betweenness on directed gnm(2^20, 2^24), one batch
    g = rx.directed_gnm_random_graph(2**20, 2**24, seed=7); dg = rx.device_snapshot(g)
    plan = np.random.default_rng(12345).permutation(g.num_nodes()).astype(np.uint32)
    dg.edge_betweenness_partial(plan[:2048])
elif which == "ba4m":    # config 5: betweenness on BA(4M, 8), one batch (4 groups)
    g = rx.barabasi_albert_graph(4_000_000, 8, seed=3); dg = rx.device_snapshot(g)
    plan = np.random.default_rng(12345).permutation(g.num_nodes()).astype(np.uint32)
    dg.betweenness_partial(plan[:512], False)
st = dg.stats()
print(which, {k: st[k] for k in ("te", "ms_total", "ms_forward", "ms_backward", "launches", "levels")})
