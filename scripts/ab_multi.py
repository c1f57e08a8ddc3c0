"""A/B of option sets on the BC bench graph, one line per configuration:
python scripts/ab_multi.py "S=16384;bc_order=1" "S=16384;bc_order=2;bc_groups=16" ...
(S = sources per call; every other key is an rxc_set_option name; LIB=path picks another build)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rustworkx_b200 as rx
from rustworkx_b200 import _lib
n = int(os.environ.get("N", 1_000_000)); reps = int(os.environ.get("REPS", 2))
g = (rx.directed_gnm_random_graph(2**20, 2**24, seed=7) if os.environ.get("GRAPH") == "gnm"
     else rx.barabasi_albert_graph(n, 8, seed=1))
n = g.num_nodes()
dg = rx.device_snapshot(g)
plan = np.random.default_rng(12345).permutation(n).astype(np.uint32)
DEFAULTS = {"bc_order": 2, "bc_groups": 0, "bc_sigma": 0, "hub_degree": 512, "bc_sparse_div": 16, "bc_siblings": 2, "bc_bwd_ctas": 32768, "bc_look": 1}
ref = {}
for spec in sys.argv[1:]:
    kv = dict(x.split("=") for x in spec.split(";") if x)
    S = int(kv.pop("S", 1024)); mode = kv.pop("mode", "node")
    for k, v in DEFAULTS.items():
        _lib.set_option(k, v)
    for k, v in kv.items():
        _lib.set_option(k, int(v))
    best = None
    for rep in range(reps + 1):
        out = dg.edge_betweenness_partial(plan[:S]) if mode == "edge" else dg.betweenness_partial(plan[:S], mode == "ep")
        st = dg.stats()
        if rep and (best is None or st["ms_total"] < best["ms_total"]):
            best = st
    key = (S, mode)
    ref.setdefault(key, out)
    err = float(np.max(np.abs(out - ref[key]) / np.maximum(np.abs(ref[key]), 1e-9)))
    print(json.dumps({"spec": spec, "ms_total": round(best["ms_total"], 2), "fwd": round(best["ms_forward"], 2),
                      "bwd": round(best["ms_backward"], 2), "gteps": round(best["te"] / best["ms_total"] / 1e6, 1),
                      "launches": best["launches"], "batches": best["batches"], "relerr_vs_first": err}), flush=True)
