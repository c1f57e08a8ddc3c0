"""One betweenness call on the bench graph for ncu (no warm-up: every launch of the call is captured):
ncu --set full --clock-control none --import-source on -k regex:"bc_(forward|backward)(_look)?_kernel" -c 40 \
    -o gpurun_out/r02_step python scripts/ncu_step.py [sources]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rustworkx_b200 as rx
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
g = (rx.directed_gnm_random_graph(2**20, 2**24, seed=7) if os.environ.get("GRAPH") == "gnm"
     else rx.barabasi_albert_graph(1_000_000, 8, seed=1))
dg = rx.device_snapshot(g)
plan = np.random.default_rng(12345).permutation(g.num_nodes()).astype(np.uint32)
mode = sys.argv[2] if len(sys.argv) > 2 else "node"
out = dg.edge_betweenness_partial(plan[:S]) if mode == "edge" else dg.betweenness_partial(plan[:S], False)
st = dg.stats()
print({k: st[k] for k in ("te", "te_dag", "v", "ms_total", "ms_forward", "ms_backward", "launches", "levels")})
