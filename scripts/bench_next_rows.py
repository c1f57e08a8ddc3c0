"""SURVEY 8f rows on the bench graph BA(1M, 8): group betweenness and Newman weighted closeness, with
the CPU oracle timed beside them on a bounded sample (oracle = test infrastructure, used as checker
and CPU baseline only).  One JSON line per case."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import rustworkx_b200 as rx

n = int(os.environ.get("N", 1_000_000)); S = int(os.environ.get("S", 1024)); CPU_S = int(os.environ.get("CPU_S", 64))
g = rx.barabasi_albert_graph(n, 8, seed=1)
dg = rx.device_snapshot(g)
rng = np.random.default_rng(7)
sources = rng.choice(n, size=S, replace=False).astype(np.uint32)
group = rng.choice(n, size=100, replace=False).astype(np.uint32); group[:2] = [0, 1]
_, eb = g._bounds()
w = rng.uniform(0.05, 2.0, size=eb)
import oracle
og = oracle.OracleGraph.from_graph(g)
threads = oracle.max_threads()

def line(case, fn, cpu_fn, check):
    for _ in range(3):
        t0 = time.perf_counter(); got = fn(); wall = time.perf_counter() - t0
    st = dg.stats()
    t0 = time.perf_counter(); want, cst = cpu_fn(); cpu_s = time.perf_counter() - t0
    print(json.dumps({"case": case, "sources": S, "device_ms": round(st["ms_total"], 2), "wall_ms": round(wall * 1e3, 2),
                      "te": st["te"], "gteps_device": round(st["te"] / st["ms_total"] / 1e6, 1),
                      "rounds_or_levels": st["levels"], "launches": st["launches"],
                      "cpu": {"kind": "port", "cores": threads, "sample_sources": CPU_S, "seconds": round(cpu_s, 2),
                              "gteps": round(cst["te"] / cpu_s / 1e9, 3)},
                      "parity": check(got, want)}), flush=True)

# group betweenness: partial sums over S sources (both BFSs of every source count into TE)
line("group_betweenness_centrality, |group|=100",
     lambda: dg.group_betweenness_partial(group, sources),
     lambda: og.group_betweenness_partial(group, sources[:CPU_S], threads=threads),
     lambda got, want: {"rel_err_on_cpu_sample": abs(dg.group_betweenness_partial(group, sources[:CPU_S]) - want) / abs(want)})
# Newman weighted closeness of S sources, uniform(0.05, 2) weights
def cpu_newman():
    out, st = og.weighted_closeness_partial(w, sources[:CPU_S], threads=threads)
    return out[sources[:CPU_S]], st
line("newman_weighted_closeness_centrality, uniform weights",
     lambda: dg.weighted_closeness_partial(w, sources),
     cpu_newman,
     lambda got, want: {"max_rel_err_on_cpu_sample": float(np.max(np.abs(got[:CPU_S] - want) / want))})
