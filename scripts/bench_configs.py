"""Timing of the other BASELINE.json configs (closeness / distance_matrix / edge BC) on one GPU."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rustworkx_b200 as rx
from rustworkx_b200 import _lib

def timed(label, fn, dg, te_bytes=None):
    t0 = time.perf_counter(); out = fn(); wall = time.perf_counter() - t0
    st = dg.stats()
    gteps = st["te"] / (st["ms_total"] * 1e-3) / 1e9 if st["ms_total"] else 0
    print(json.dumps({"case": label, "wall_s": round(wall, 4), "device_ms": round(st["ms_total"], 3),
                      "gteps_device": round(gteps, 2), "te": st["te"], "v": st["v"], "levels": st["levels"],
                      "launches": st["launches"], "batches": st["batches"], "d2h_ms": round(st["ms_d2h"], 2)}), flush=True)
    return out

for kv in os.environ.get("OPT", "").split(","):
    if "=" in kv:
        _lib.set_option(kv.split("=")[0], int(kv.split("=")[1]))
which = sys.argv[1:] or ["hh", "grid", "gnm"]
if "hh" in which:
    g = rx.generators.heavy_hex_graph(51)
    dg = rx.device_snapshot(g)
    for rep in range(2):
        timed("heavy_hex(51) closeness", lambda: dg.closeness(True), dg)
    for rep in range(2):
        timed("heavy_hex(51) distance_matrix", lambda: dg.distance_matrix(True, 0.0), dg)
if "grid" in which:
    side = int(os.environ.get("GRID", "1000"))
    g = rx.generators.grid_graph(side, side)
    dg = rx.device_snapshot(g)
    nsrc = int(os.environ.get("GRID_SOURCES", "131072"))
    src = np.random.default_rng(0).choice(side * side, size=min(nsrc, side * side), replace=False).astype(np.uint32)
    src.sort()
    for rep in range(2):
        timed(f"grid({side},{side}) closeness, {len(src)} sampled sources", lambda: dg.closeness_partial(src, True), dg)
if "gnm" in which:
    g = rx.directed_gnm_random_graph(2**20, 2**24, seed=7)
    dg = rx.device_snapshot(g)
    src = np.random.default_rng(0).choice(2**20, size=512, replace=False).astype(np.uint32)
    for rep in range(2):
        timed("directed_gnm(2^20,2^24) edge betweenness, 512 sources", lambda: dg.edge_betweenness_partial(src), dg)
    for rep in range(2):
        timed("directed_gnm(2^20,2^24) node betweenness, 512 sources", lambda: dg.betweenness_partial(src), dg)
