"""Generates tests/golden/nx_golden_next_rows.json with networkx 3.6.1: independent vectors for the
SURVEY 8f rows (group centralities, weighted closeness, all-pairs Dijkstra lengths, average shortest
path length) on the graphs of nx_golden.json.  The reference's own tests take their expected values
for these functions from networkx as well (tests/graph/test_centrality.py:415-423,515-567).

    python tests/golden/make_golden_next_rows.py

Group betweenness is computed from its definition by enumerating shortest paths
(nx.all_shortest_paths): networkx's own group_betweenness_centrality raises or is wrong on graphs with
unreachable pairs; on the connected undirected cases the two are asserted equal here.
"""
import json
import os

import networkx as nx
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    base = json.load(open(os.path.join(HERE, "nx_golden.json")))
    cases = []
    for ci, c in enumerate(base["cases"]):
        directed, n, edges = c["directed"], c["n"], [tuple(e) for e in c["edges"]]
        G = nx.DiGraph() if directed else nx.Graph()
        G.add_nodes_from(range(n))
        G.add_edges_from(edges)
        rng = np.random.default_rng(500 + ci)
        group = sorted(int(x) for x in rng.choice(n, size=3, replace=False))
        # group betweenness from the definition, by enumerating shortest paths (independent of both
        # the oracle and networkx's own group algorithm, which is wrong or raises on graphs with
        # unreachable pairs: case 5 is such a digraph -- networkx says 88.83, the enumeration 98.83)
        C = set(group)
        tot = 0.0
        for s in G:
            if s in C:
                continue
            for t in G:
                if t in C or t == s:
                    continue
                try:
                    paths = list(nx.all_shortest_paths(G, s, t))
                except nx.NetworkXNoPath:
                    continue
                tot += sum(1 for p in paths if C & set(p[1:-1])) / len(paths)
        if not directed:
            tot /= 2.0
        rest = n - len(C)
        scale = (1.0 if directed else 2.0) / (rest * (rest - 1))  # centrality.rs:2088-2098
        gbc = {"False": tot, "True": tot * scale}
        if not directed and nx.is_connected(G):
            for normalized in (True, False):
                ref = nx.group_betweenness_centrality(G, group, normalized=normalized, endpoints=False)
                assert abs(ref - gbc[str(normalized)]) <= 1e-12 * max(1.0, abs(ref)), (ci, ref, gbc)
        gcc = nx.group_closeness_centrality(G, group)
        w = rng.uniform(0.5, 3.0, size=len(edges))
        for (a, b), ww in zip(edges, w):
            G[a][b]["weight"] = float(ww)
            G[a][b]["cost"] = 1.0 / float(ww)
        # Newman: closeness over cost = 1/weight along incoming paths (networkx reverses digraphs too)
        ncc = nx.closeness_centrality(G, distance="cost", wf_improved=True)
        dl = dict(nx.all_pairs_dijkstra_path_length(G, weight="weight"))
        rows = [[(dl[s][t] if (t in dl[s] and t != s) else None) for t in range(n)] for s in range(n)]
        # unweighted average shortest path length: sum of finite distances / n(n-1) and / connected pairs
        hops = dict(nx.all_pairs_shortest_path_length(G))
        total = sum(d for s in range(n) for t, d in hops[s].items() if t != s)
        pairs = sum(1 for s in range(n) for t in hops[s] if t != s)
        cases.append({"case": ci, "group": group, "group_betweenness": gbc, "group_closeness": gcc,
                      "weights": [float(x) for x in w], "newman_closeness_wf": [ncc[i] for i in range(n)],
                      "dijkstra_lengths": rows, "distance_total": int(total), "connected_pairs": int(pairs)})
    out = os.path.join(HERE, "nx_golden_next_rows.json")
    with open(out, "w") as f:
        json.dump({"generator": "networkx " + nx.__version__, "graphs": "nx_golden.json", "cases": cases}, f)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
