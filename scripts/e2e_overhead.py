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
PINNED = os.environ.get("PINNED", "0") == "1"   # inputs / output in rxc_host_alloc memory
NOIDS = os.environ.get("NOIDS", "0") == "1"     # edge_id = NULL (the position in the list)
arr = {k: (_lib.pinned_copy(snap[k]) if PINNED else snap[k]) for k in ("live_nodes", "edge_src", "edge_dst", "edge_id")}
if NOIDS:
    arr["edge_id"] = None
buf = _lib.pinned_empty(n, np.float64) if PINNED else None
for i in range(6):
    t0 = time.perf_counter()
    dg = _lib.DeviceGraph(snap["n_bound"], arr["live_nodes"], snap["e_bound"], arr["edge_src"],
                          arr["edge_dst"], arr["edge_id"], snap["directed"])
    t1 = time.perf_counter()
    out = dg.betweenness_partial(plan[i * S:(i + 1) * S], False, out=buf)
    t2 = time.perf_counter()
    st = dg.stats()
    dg.close()
    t3 = time.perf_counter()
    print(f"iter {i}: snapshot {1e3*(t1-t0):.1f} ms (lib {st['ms_snapshot']:.1f})  call {1e3*(t2-t1):.1f} ms (device {st['ms_total']:.1f})  close {1e3*(t3-t2):.1f} ms")
