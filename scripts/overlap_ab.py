"""Two handles on ONE device driven by two host threads (their streams overlap) against one handle:
does running the issue-bound forward pass of one batch next to the DRAM-bound dependency sweep of
another pay?  python scripts/overlap_ab.py"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rustworkx_b200 as rx

n = int(os.environ.get("N", 1_000_000)); S = int(os.environ.get("S", 1024)); REP = int(os.environ.get("REP", 4))
g = rx.barabasi_albert_graph(n, 8, seed=1)
plan = np.random.default_rng(1).permutation(n).astype(np.uint32)
rx.set_snapshot_cache(False)
h = [rx.device_snapshot(g) for _ in range(int(os.environ.get("H", 2)))]


def run(parts):
    """parts: list of (handle, sources); all run concurrently; returns wall ms, TE"""
    outs = [None] * len(parts)
    def work(i):
        outs[i] = parts[i][0].betweenness_partial(parts[i][1], False)
    th = [threading.Thread(target=work, args=(i,)) for i in range(len(parts))]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    ms = (time.perf_counter() - t0) * 1e3
    te = sum(p[0].stats()["te"] for p in parts)
    return ms, te, outs

for total in (S, 2 * S):
    ref = None
    for k in range(1, len(h) + 1):
        src = plan[:total]
        parts = [(h[i], src[i::k].copy()) for i in range(k)]
        best = None
        for rep in range(REP):
            ms, te, outs = run(parts)
            best = ms if best is None else min(best, ms)
        tot = sum(outs)
        if ref is None: ref = tot
        err = np.max(np.abs(tot - ref) / np.maximum(np.abs(ref), 1e-9))
        print(f"sources {total} over {k} handle(s): wall {best:.1f} ms  {te / best / 1e6:.1f} GTEPS  relerr {err:.1e}", flush=True)
