"""A/B: hub deferral on/off, groups per batch (diagnostic)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rustworkx_b200 as rx
from rustworkx_b200 import _lib

n = 1_000_000
g = rx.barabasi_albert_graph(n, 8, seed=1)
dg = rx.device_snapshot(g)
plan = np.random.default_rng(1).permutation(n).astype(np.uint32)
ref = None
for hub in (1 << 30, 2048, 512):
    _lib.set_option("hub_degree", hub)
    for rep in range(3):
        out = dg.betweenness_partial(plan[:256], False)
        st = dg.stats()
    if ref is None:
        ref = out
    err = np.max(np.abs(out - ref) / np.maximum(np.abs(ref), 1e-9))
    print(f"hub_degree {hub}: device {st['ms_total']:.1f} fwd {st['ms_forward']:.1f} bwd {st['ms_backward']:.1f} launches {st['launches']} relerr_vs_first {err:.2e}")
