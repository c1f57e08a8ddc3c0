"""A/B of a library option on the BC bench graph: python scripts/opt_ab.py name v1 v2 ..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rustworkx_b200 as rx
from rustworkx_b200 import _lib
name = sys.argv[1]; vals = [int(v) for v in sys.argv[2:]]
n = int(os.environ.get("N", 1_000_000)); S = int(os.environ.get("S", 256))
g = rx.barabasi_albert_graph(n, 8, seed=1)
dg = rx.device_snapshot(g)
plan = np.random.default_rng(1).permutation(n).astype(np.uint32)
ref = None
for v in vals:
    _lib.set_option(name, v)
    for rep in range(3):
        out = dg.betweenness_partial(plan[:S], False)
        st = dg.stats()
    if ref is None:
        ref = out
    err = np.max(np.abs(out - ref) / np.maximum(np.abs(ref), 1e-9))
    print(f"{name}={v}: device {st['ms_total']:.1f} fwd {st['ms_forward']:.1f} bwd {st['ms_backward']:.1f} "
          f"launches {st['launches']} gteps {st['te']/st['ms_total']/1e6:.1f} relerr_vs_first {err:.2e}")
