"""Wall clock of snapshot + one partial call + free, repeated (diagnostic for the e2e number)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rustworkx_b200 as rx
from rustworkx_b200 import _lib
n = 1_000_000; S = 1024
g = rx.barabasi_albert_graph(n, 8, seed=1)
snap = g._snapshot_arrays()
plan = np.random.default_rng(1).permutation(n).astype(np.uint32)
for i in range(6):
    t0 = time.perf_counter()
    dg = _lib.DeviceGraph(snap["n_bound"], snap["live_nodes"], snap["e_bound"], snap["edge_src"],
                          snap["edge_dst"], snap["edge_id"], snap["directed"])
    t1 = time.perf_counter()
    out = dg.betweenness_partial(plan[i * S:(i + 1) * S], False)
    t2 = time.perf_counter()
    st = dg.stats()
    dg.close()
    t3 = time.perf_counter()
    print(f"iter {i}: snapshot {1e3*(t1-t0):.1f} ms (lib {st['ms_snapshot']:.1f})  call {1e3*(t2-t1):.1f} ms (device {st['ms_total']:.1f})  close {1e3*(t3-t2):.1f} ms")
