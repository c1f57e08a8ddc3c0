"""Sanity timings of betweenness on deep graphs and small graphs (not BASELINE configs)."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rustworkx_b200 as rx
def run(label, g, fn):
    dg = rx.device_snapshot(g)
    for rep in range(2):
        t0 = time.perf_counter(); fn(g); wall = time.perf_counter() - t0
    st = dg.stats()
    print(json.dumps({"case": label, "n": g.num_nodes(), "wall_s": round(wall, 4), "device_ms": round(st["ms_total"], 2),
                      "levels": st["levels"], "launches": st["launches"], "batches": st["batches"],
                      "gteps": round(st["te"] / max(st["ms_total"], 1e-9) / 1e6, 2)}), flush=True)
run("heavy_hex(51) betweenness", rx.generators.heavy_hex_graph(51), lambda g: rx.betweenness_centrality(g))
run("grid(200,200) betweenness", rx.generators.grid_graph(200, 200), lambda g: rx.betweenness_centrality(g))
run("gnp(5000,0.002) betweenness [config 1]", rx.undirected_gnp_random_graph(5000, 0.002, seed=42), lambda g: rx.betweenness_centrality(g))
run("gnp(5000,0.002) edge betweenness", rx.undirected_gnp_random_graph(5000, 0.002, seed=42), lambda g: rx.edge_betweenness_centrality(g))
run("BA(100000,8) closeness", rx.barabasi_albert_graph(100000, 8, seed=1), lambda g: rx.closeness_centrality(g))
run("path(2000) betweenness", rx.generators.path_graph(2000), lambda g: rx.betweenness_centrality(g))
