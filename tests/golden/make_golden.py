"""Generates tests/golden/nx_golden.json with networkx 3.6.1 (an implementation independent of both
the oracle and the CUDA path; the reference's own tests use networkx the same way, e.g.
tests/graph/test_centrality.py:415-423).  Run here, commit the output:

    python tests/golden/make_golden.py

The reference itself (Rust) cannot be imported in this image, so these vectors complement -- not
replace -- the reference's hard-coded golden values pinned in tests/test_oracle_golden.py.
Simple graphs only (networkx Graph/DiGraph have no parallel edges); edges keep insertion order so
edge indices line up.
"""
import json
import os

import networkx as nx
import numpy as np


def case(directed, n, edges, normalized, endpoints):
    G = nx.DiGraph() if directed else nx.Graph()
    G.add_nodes_from(range(n))
    G.add_edges_from(edges)
    bc = nx.betweenness_centrality(G, normalized=normalized, endpoints=endpoints)
    ebc = nx.edge_betweenness_centrality(G, normalized=normalized)
    if not normalized and not directed:
        pass  # networkx already halves undirected scores, as the reference does (scale 0.5)
    cc = nx.closeness_centrality(G, wf_improved=True)  # incoming distances, like the reference
    dm = np.full((n, n), -1.0)
    for s, dd in nx.all_pairs_shortest_path_length(G):
        for t, d in dd.items():
            dm[s, t] = float(d)
    def e_val(a, b):
        return ebc[(a, b)] if (a, b) in ebc else ebc[(b, a)]
    return {
        "directed": directed, "n": n, "edges": [list(e) for e in edges], "normalized": normalized,
        "endpoints": endpoints, "betweenness": [bc[i] for i in range(n)],
        "edge_betweenness": [e_val(a, b) for a, b in edges],
        "closeness_wf": [cc[i] for i in range(n)], "distance_matrix": dm.tolist(),
    }


def random_simple(directed, n, m, seed):
    rng = np.random.default_rng(seed)
    seen, edges = set(), []
    while len(edges) < m:
        a, b = (int(x) for x in rng.integers(0, n, size=2))
        key = (a, b) if directed else (min(a, b), max(a, b))
        if a != b and key not in seen:
            seen.add(key)
            edges.append((a, b))
    return edges


def main():
    cases = []
    for i, (directed, n, m) in enumerate([(False, 30, 45), (True, 30, 80), (False, 60, 70),
                                           (True, 60, 150), (False, 25, 200), (True, 40, 60)]):
        edges = random_simple(directed, n, m, 100 + i)
        cases.append(case(directed, n, edges, normalized=(i % 2 == 0), endpoints=(i % 3 == 0)))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nx_golden.json")
    with open(out, "w") as f:
        json.dump({"generator": "networkx " + nx.__version__, "cases": cases}, f)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
