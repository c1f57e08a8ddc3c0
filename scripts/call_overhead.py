"""Wall clock vs device time of rxc_betweenness_partial calls (diagnostic)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rustworkx_b200 as rx
from rustworkx_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
g = rx.barabasi_albert_graph(n, 8, seed=1)
snap = g._snapshot_arrays()
t0 = time.perf_counter()
dg = _lib.DeviceGraph(snap["n_bound"], snap["live_nodes"], snap["e_bound"], snap["edge_src"],
                      snap["edge_dst"], snap["edge_id"], snap["directed"])
print("snapshot wall ms", 1e3 * (time.perf_counter() - t0), dg.stats()["ms_snapshot"])
plan = np.random.default_rng(1).permutation(n).astype(np.uint32)
for i in range(6):
    src = plan[i * S:(i + 1) * S]
    t0 = time.perf_counter()
    out = dg.betweenness_partial(src, False)
    wall = 1e3 * (time.perf_counter() - t0)
    st = dg.stats()
    print(f"call {i}: wall {wall:.1f} ms  device {st['ms_total']:.1f}  fwd {st['ms_forward']:.1f} "
          f"bwd {st['ms_backward']:.1f} d2h {st['ms_d2h']:.1f} batches {st['batches']} launches {st['launches']}")
t0 = time.perf_counter()
dg.close()
print("close wall ms", 1e3 * (time.perf_counter() - t0))
