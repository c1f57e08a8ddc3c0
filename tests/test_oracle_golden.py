"""Pins the CPU oracle against every golden vector the reference's own tests hold for the hot path
(SURVEY.md section 8c).  Exact `==` where the reference uses assert_eq / assertEqual, tolerance where
the reference uses assert_almost_equal / assertAlmostEqual.
"""

import math

import numpy as np
import pytest


# ----------------------------------------------------------------------------- node betweenness
def test_bc_core_doctest(oracle):
    # rustworkx-core/src/centrality.rs:58-66
    g = oracle.OracleGraph.from_edges(False, [(0, 4), (1, 2), (2, 3), (3, 4), (1, 4)])
    assert g.betweenness_centrality(True, True, 200) == [0.4, 0.5, 0.45, 0.5, 0.75]


def _hole_graph(oracle):
    # rustworkx-core/src/centrality.rs:640-658 and :674-691
    g = oracle.OracleGraph(False)
    n0 = g.add_node(); d0 = g.add_node(); n1 = g.add_node(); d1 = g.add_node()
    n2 = g.add_node(); d2 = g.add_node(); n3 = g.add_node()
    g.remove_node(d0); g.remove_node(d1); g.remove_node(d2)
    g.add_edge(n0, n1); g.add_edge(n1, n2); g.add_edge(n2, n3); g.add_edge(n3, n0)
    g.remove_edge(1)
    return g


def test_bc_stable_graph_with_removed_nodes_and_edges(oracle):
    # rustworkx-core/src/centrality.rs:672-696
    g = _hole_graph(oracle)
    assert g.betweenness_centrality(False, False, 50) == [2.0, None, 0.0, None, 0.0, None, 2.0]


@pytest.mark.parametrize("threshold", [50, 1])
@pytest.mark.parametrize("directed", [False, True])
@pytest.mark.parametrize("deleted", [False, True])
def test_bc_python_path4(oracle, threshold, directed, deleted):
    # tests/graph/test_centrality.py:20-103, tests/digraph/test_centrality.py:20-132
    g = oracle.OracleGraph(directed)
    a, b, c = g.add_node(), g.add_node(), g.add_node()
    c0 = g.add_node() if deleted else None
    d = g.add_node()
    g.add_edge(a, b); g.add_edge(b, c); g.add_edge(c, d)
    if deleted:
        g.remove_node(c0)
    keys = [0, 1, 2, 4] if deleted else [0, 1, 2, 3]

    def as_map(v):
        return {i: x for i, x in enumerate(v) if x is not None}

    if directed:
        exp_norm = [0.0, 0.3333333333333333, 0.3333333333333333, 0.0]
        exp_end = [0.25, 0.41666666666666663, 0.41666666666666663, 0.25]
    else:
        exp_norm = [0.0, 0.6666666666666666, 0.6666666666666666, 0.0]
        exp_end = [0.5, 0.8333333333333333, 0.8333333333333333, 0.5]
    exp_un = [0.0, 2.0, 2.0, 0.0]
    # python order (normalized, endpoints) -> core order (include_endpoints, normalized)
    assert as_map(g.betweenness_centrality(False, True, threshold)) == dict(zip(keys, exp_norm))
    assert as_map(g.betweenness_centrality(True, True, threshold)) == dict(zip(keys, exp_end))
    assert as_map(g.betweenness_centrality(False, False, threshold)) == dict(zip(keys, exp_un))


# ----------------------------------------------------------------------------- edge betweenness
def test_ebc_core_doctest(oracle):
    # rustworkx-core/src/centrality.rs:164-170
    g = oracle.OracleGraph.from_edges(False, [(0, 4), (1, 2), (1, 3), (2, 3), (3, 4), (1, 4)])
    assert g.edge_betweenness_centrality(False, 200) == [4.0, 2.0, 1.0, 2.0, 3.0, 3.0]


EBC_CASES = [
    # rustworkx-core/src/centrality.rs:528-556
    (False, True,
     [(0, 6), (0, 4), (0, 1), (0, 5), (1, 6), (1, 7), (1, 3), (1, 4), (2, 6), (2, 3), (3, 5),
      (3, 7), (3, 6), (4, 5), (5, 6)],
     [0.1023809, 0.0547619, 0.0922619, 0.05654762, 0.09940476, 0.125, 0.09940476, 0.12440476,
      0.12857143, 0.12142857, 0.13511905, 0.125, 0.06547619, 0.08869048, 0.08154762]),
    # :558-582
    (False, False,
     [(0, 2), (0, 4), (0, 1), (1, 3), (1, 5), (1, 7), (2, 7), (2, 3), (3, 5), (3, 6), (4, 6),
      (5, 7)],
     [3.83333, 5.5, 5.33333, 3.5, 2.5, 3.0, 3.5, 4.0, 3.66667, 6.5, 3.5, 2.16667]),
    # :584-604
    (True, True,
     [(0, 1), (1, 0), (1, 3), (1, 2), (1, 4), (2, 3), (2, 4), (2, 1), (3, 2), (4, 3)],
     [0.2, 0.2, 0.1, 0.1, 0.1, 0.05, 0.1, 0.3, 0.35, 0.2]),
    # :606-626
    (True, False,
     [(0, 4), (1, 0), (1, 3), (2, 3), (2, 4), (2, 0), (3, 4), (3, 2), (3, 1), (4, 1)],
     [4.5, 3.0, 6.5, 1.5, 1.5, 1.5, 1.5, 4.5, 2.0, 7.5]),
]


@pytest.mark.parametrize("directed,normalized,edges,expected", EBC_CASES)
def test_ebc_core_unit(oracle, directed, normalized, edges, expected):
    g = oracle.OracleGraph.from_edges(directed, edges)
    out = g.edge_betweenness_centrality(normalized, 50)
    assert len(out) == len(expected)
    for x, y in zip(out, expected):
        assert abs(x - y) < 1e-4


def test_ebc_stable_graph_with_removed_edges(oracle):
    # rustworkx-core/src/centrality.rs:628-636
    g = oracle.OracleGraph.from_edges(False, [(0, 1), (1, 2), (2, 3), (3, 0)])
    g.remove_edge(1)
    assert g.edge_betweenness_centrality(False, 50) == [3.0, None, 3.0, 4.0]


def test_ebc_stable_graph_with_removed_nodes_and_edges(oracle):
    # rustworkx-core/src/centrality.rs:638-662
    g = _hole_graph(oracle)
    assert g.edge_betweenness_centrality(False, 50) == [3.0, None, 3.0, 4.0]


def _gen(oracle, graph):
    return oracle.OracleGraph.from_graph(graph)


def test_ebc_python_goldens(oracle):
    # tests/graph/test_centrality.py:194-248 and tests/digraph/test_centrality.py:227-259
    import rustworkx_b200.generators as gen

    def vals(g, normalized=True):
        return _gen(oracle, g).edge_betweenness_centrality(normalized, 50)

    for v in vals(gen.mesh_graph(5)):
        assert v == pytest.approx(0.1, abs=1e-7)
    assert vals(gen.path_graph(5)) == pytest.approx([0.4, 0.6, 0.6, 0.4], abs=1e-7)
    for v in vals(gen.cycle_graph(5)):
        assert v == pytest.approx(0.3, abs=1e-7)
    assert vals(gen.full_rary_tree(2, 7), False) == pytest.approx([12, 12, 6, 6, 6, 6], abs=1e-7)
    assert vals(gen.path_graph(5), False) == pytest.approx([4, 6, 6, 4], abs=1e-7)
    from rustworkx_b200 import PyGraph

    g = PyGraph()
    g.add_nodes_from(range(10))
    g.add_edges_from([(0, 1, 1), (0, 2, 1), (0, 3, 1), (0, 4, 1), (3, 5, 1), (4, 6, 1), (5, 7, 1),
                      (6, 8, 1), (7, 8, 1), (8, 9, 1)])
    assert vals(g, False) == pytest.approx([9, 9, 12, 15, 11, 14, 10, 13, 9, 9], abs=1e-7)
    # directed
    for v in vals(gen.directed_mesh_graph(5)):
        assert v == pytest.approx(0.05, abs=1e-7)
    assert vals(gen.directed_path_graph(5)) == pytest.approx([0.2, 0.3, 0.3, 0.2], abs=1e-7)
    for v in vals(gen.directed_cycle_graph(5)):
        assert v == pytest.approx(0.5, abs=1e-7)
    assert vals(gen.full_rary_tree(2, 7).to_directed(), False) == pytest.approx(
        [12, 12, 12, 12, 6, 6, 6, 6, 6, 6, 6, 6], abs=1e-7)
    assert vals(gen.directed_path_graph(5), False) == pytest.approx([4, 6, 6, 4], abs=1e-7)


# ----------------------------------------------------------------------------- closeness
def test_closeness_core_doctest(oracle):
    # rustworkx-core/src/centrality.rs:1146-1164
    edges = [(0, 4), (1, 2), (2, 3), (3, 4), (1, 4)]
    g = oracle.OracleGraph.from_edges(False, edges)
    assert g.closeness_centrality(True, 200) == [1.0 / 2.0, 2.0 / 3.0, 4.0 / 7.0, 2.0 / 3.0, 4.0 / 5.0]
    dg = oracle.OracleGraph.from_edges(True, edges)
    assert dg.closeness_centrality(True, 200) == [0.0, 0.0, 1.0 / 4.0, 1.0 / 3.0, 4.0 / 5.0]


@pytest.mark.parametrize("threshold", [200, 1])
def test_closeness_core_integration(oracle, threshold):
    # rustworkx-core/tests/centrality.rs:17-120
    g = oracle.OracleGraph.from_edges(False, [(1, 2), (2, 3), (3, 4), (1, 4)])
    assert g.closeness_centrality(True, threshold) == [0.0, 0.5625, 0.5625, 0.5625, 0.5625]
    g = oracle.OracleGraph.from_edges(False, [(0, 1), (1, 2), (2, 3), (4, 5), (5, 6)])
    assert g.closeness_centrality(True, threshold) == [
        1.0 / 4.0, 3.0 / 8.0, 3.0 / 8.0, 1.0 / 4.0, 2.0 / 9.0, 1.0 / 3.0, 2.0 / 9.0]
    assert g.closeness_centrality(False, threshold) == [
        1.0 / 2.0, 3.0 / 4.0, 3.0 / 4.0, 1.0 / 2.0, 2.0 / 3.0, 1.0, 2.0 / 3.0]
    g = oracle.OracleGraph.from_edges(True, [(0, 1), (1, 2)])
    assert g.closeness_centrality(True, threshold) == [0.0, 1.0 / 2.0, 2.0 / 3.0]
    # Reversed(&g): same as the digraph with every edge flipped
    gr = oracle.OracleGraph.from_edges(True, [(1, 0), (2, 1)])
    assert gr.closeness_centrality(True, threshold) == [2.0 / 3.0, 1.0 / 2.0, 0.0]
    k5 = [(0, 1), (0, 2), (0, 3), (0, 4), (1, 2), (1, 3), (1, 4), (2, 3), (2, 4), (3, 4)]
    assert oracle.OracleGraph.from_edges(False, k5).closeness_centrality(True, threshold) == [1.0] * 5
    g = oracle.OracleGraph.from_edges(False, [(0, 1), (1, 2)])
    assert g.closeness_centrality(True, threshold) == [2.0 / 3.0, 1.0, 2.0 / 3.0]


@pytest.mark.parametrize("threshold", [50, 1])
def test_closeness_python_deleted_node(oracle, threshold):
    # tests/graph/test_centrality.py:105-120 and tests/digraph/test_centrality.py:134-149
    for directed in (False, True):
        g = oracle.OracleGraph(directed)
        a, b, c, c0, d = (g.add_node() for _ in range(5))
        g.add_edge(a, b); g.add_edge(b, c); g.add_edge(c, d)
        g.remove_node(c0)
        wf = g.closeness_centrality(True, threshold)
        nowf = g.closeness_centrality(False, threshold)
        if directed:
            assert wf == [0.0, 1.0 / 3.0, 4.0 / 9.0, None, 0.5]
            assert nowf == [0.0, 1.0, 2.0 / 3.0, None, 0.5]
        else:
            assert wf == [0.5, 0.75, 0.75, None, 0.5]
            assert nowf == [0.5, 0.75, 0.75, None, 0.5]


# ----------------------------------------------------------------------------- distance matrix
CYCLE7 = [(0, 1), (0, 6), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6)]
CYCLE7_UNDIRECTED = np.array([
    [0.0, 1.0, 2.0, 3.0, 3.0, 2.0, 1.0],
    [1.0, 0.0, 1.0, 2.0, 3.0, 3.0, 2.0],
    [2.0, 1.0, 0.0, 1.0, 2.0, 3.0, 3.0],
    [3.0, 2.0, 1.0, 0.0, 1.0, 2.0, 3.0],
    [3.0, 3.0, 2.0, 1.0, 0.0, 1.0, 2.0],
    [2.0, 3.0, 3.0, 2.0, 1.0, 0.0, 1.0],
    [1.0, 2.0, 3.0, 3.0, 2.0, 1.0, 0.0],
])
CYCLE7_DIRECTED = np.array([
    [0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 1.0],
    [0.0, 0.0, 1.0, 2.0, 3.0, 4.0, 5.0],
    [0.0, 0.0, 0.0, 1.0, 2.0, 3.0, 4.0],
    [0.0, 0.0, 0.0, 0.0, 1.0, 2.0, 3.0],
    [0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 2.0],
    [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0],
    [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
])


def _nan_expected():
    e = np.full((8, 8), np.nan)
    e[:7, :7] = CYCLE7_UNDIRECTED
    e[7, 7] = 0.0
    return e


@pytest.mark.parametrize("threshold", [300, 5])
def test_distance_matrix_goldens(oracle, threshold):
    # distance_matrix.rs:56-69; tests/graph/test_dist_matrix.py:21-95;
    # tests/digraph/test_dist_matrix.py:21-132
    g = oracle.OracleGraph.from_edges(False, CYCLE7)
    assert np.array_equal(g.distance_matrix(threshold, True, 0.0), CYCLE7_UNDIRECTED)
    assert np.array_equal(g.distance_matrix(threshold, False, 0.0), CYCLE7_UNDIRECTED)
    dg = oracle.OracleGraph.from_edges(True, CYCLE7)
    assert np.array_equal(dg.distance_matrix(threshold, False, 0.0), CYCLE7_DIRECTED)
    assert np.array_equal(dg.distance_matrix(threshold, True, 0.0), CYCLE7_UNDIRECTED)
    for directed in (False, True):
        g = oracle.OracleGraph.from_edges(directed, CYCLE7)
        g.add_node()
        got = g.distance_matrix(threshold, True, math.nan)
        assert np.array_equal(got, _nan_expected(), equal_nan=True)


def test_distance_matrix_node_hole(oracle):
    # tests/graph/test_dist_matrix.py:97-102, tests/digraph/test_dist_matrix.py:135-140
    g = oracle.OracleGraph.from_edges(False, [(0, 1), (1, 2), (2, 3)])
    g.remove_node(0)
    assert np.array_equal(g.distance_matrix(300, True, 0.0),
                          np.array([[0.0, 1.0, 2.0], [1.0, 0.0, 1.0], [2.0, 1.0, 0.0]]))
    dg = oracle.OracleGraph.from_edges(True, [(0, 1), (1, 2), (2, 3)])
    dg.remove_node(0)
    assert np.array_equal(dg.distance_matrix(300, False, 0.0),
                          np.array([[0.0, 1.0, 2.0], [0.0, 0.0, 1.0], [0.0, 0.0, 0.0]]))


# ----------------------------------------------------------------------------- generators
def test_heavy_hex_fixture(oracle):
    # rustworkx-core/src/generators/heavy_hex_graph.rs:76-105 (doctest edge list, d=3)
    import rustworkx_b200.generators as gen

    expected = [(0, 13), (1, 13), (1, 14), (2, 14), (3, 15), (4, 15), (4, 16), (5, 16), (6, 17),
                (7, 17), (7, 18), (8, 18), (0, 9), (3, 9), (5, 12), (8, 12), (10, 14), (10, 16),
                (11, 15), (11, 17)]
    g = gen.heavy_hex_graph(3)
    assert list(g.edge_list()) == expected
    for d in (3, 5, 7, 51):
        g = gen.heavy_hex_graph(d)
        assert g.num_nodes() == (5 * d * d - 2 * d - 1) // 2
        assert g.num_edges() == 2 * d * (d - 1) + (d + 1) * (d - 1)


def test_grid_fixture(oracle):
    # rustworkx-core/src/generators/grid_graph.rs doctest: 3x3 grid edge order
    import rustworkx_b200.generators as gen

    g = gen.grid_graph(3, 3)
    assert list(g.edge_list()) == [(0, 3), (0, 1), (1, 4), (1, 2), (2, 5), (3, 6), (3, 4), (4, 7),
                                   (4, 5), (5, 8), (6, 7), (7, 8)]
    g = gen.grid_graph(1000, 1000)
    assert g.num_nodes() == 10**6 and g.num_edges() == 1998000


# ----------------------------------------------------------------------------- average path length
def test_average_shortest_path_length_goldens(oracle):
    # tests/graph/test_avg_shortest_path.py:20-100
    import rustworkx_b200.generators as gen

    def val(g, **kw):
        return _gen(oracle, g).average_shortest_path_length(**kw)

    g = oracle.OracleGraph.from_edges(False, [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (3, 6), (6, 7)])
    assert g.average_shortest_path_length() == pytest.approx(2.5714285714285716, abs=1e-7)
    assert val(gen.cycle_graph(7)) == pytest.approx(2, abs=1e-7)
    assert val(gen.path_graph(5)) == pytest.approx(2, abs=1e-7)
    assert val(gen.grid_graph(30, 11)) == pytest.approx(13.666666666666666, abs=1e-7)
    assert math.isnan(oracle.OracleGraph(False).average_shortest_path_length())
    one = oracle.OracleGraph(False)
    a = one.add_node()
    assert math.isnan(one.average_shortest_path_length())
    one.add_edge(a, a)
    assert math.isnan(one.average_shortest_path_length())
    iso = oracle.OracleGraph(False)
    for _ in range(32):
        iso.add_node()
    assert math.isinf(iso.average_shortest_path_length())
    assert math.isnan(iso.average_shortest_path_length(disconnected=True))
    part = gen.cycle_graph(32)
    part.add_nodes_from(list(range(32)))
    assert math.isinf(val(part))
    assert val(part, disconnected=True) == pytest.approx(8192 / 992, abs=1e-7)
    assert val(gen.cycle_graph(32)) == pytest.approx(8192 / 992, abs=1e-7)
    # directed: tests/digraph/test_avg_shortest_path.py
    assert val(gen.directed_cycle_graph(7)) == pytest.approx(3.5, abs=1e-7)
    assert math.isinf(val(gen.directed_path_graph(5)))
    assert val(gen.directed_path_graph(5), as_undirected=True) == pytest.approx(2, abs=1e-7)


# ----------------------------------------------------------------------------- group centralities
def test_group_closeness_goldens(oracle):
    # rustworkx-core/src/centrality.rs:2155-2211; tests/graph/test_centrality.py:341-497;
    # tests/digraph/test_centrality.py:374-495
    from group_cases import CLOSENESS

    for name, directed, n, edges, group, expected in CLOSENESS:
        g = oracle.OracleGraph.from_edges(directed, edges, num_nodes=n)
        assert abs(g.group_closeness_centrality(group) - expected) < 1e-10, name


@pytest.mark.parametrize("threshold", [50, 1])
def test_group_betweenness_goldens(oracle, threshold):
    # rustworkx-core/src/centrality.rs:2213-2360; tests/graph/test_centrality.py:366-567;
    # tests/digraph/test_centrality.py:396-565
    from group_cases import BETWEENNESS

    for name, directed, n, edges, group, normalized, expected in BETWEENNESS:
        g = oracle.OracleGraph.from_edges(directed, edges, num_nodes=n)
        got = g.group_betweenness_centrality(group, normalized, threshold)
        assert abs(got - expected) < 1e-10, name


def test_group_betweenness_singleton_matches_betweenness(oracle):
    # rustworkx-core/src/centrality.rs:2342-2360: a one-node group equals that node's betweenness
    g = oracle.OracleGraph.from_edges(True, [(0, 1), (1, 2), (2, 3)])
    per_node = g.betweenness_centrality(False, False, 200)
    for v in range(4):
        assert abs(g.group_betweenness_centrality([v], False, 200) - per_node[v]) < 1e-10


# ----------------------------------------------------------------------------- weighted shortest paths
def _weighted_oracle_graph(oracle, directed, edges, remove=None):
    g = oracle.OracleGraph.from_edges(directed, [(a, b) for a, b, _ in edges])
    w = [float(x) for _, _, x in edges]
    if remove is not None:
        g.remove_node(remove)
    return g, w


def test_newman_weighted_closeness_goldens(oracle):
    # rustworkx-core/src/centrality.rs:1382-1760
    from weighted_cases import NEWMAN, NEWMAN_UNIT

    for name, directed, edges, wf, expected in NEWMAN:
        g, w = _weighted_oracle_graph(oracle, directed, edges)
        for threshold in (200, 1):
            got = g.newman_weighted_closeness_centrality(w, wf, threshold)
            assert len(got) == len(expected), name
            for a, b in zip(got, expected):
                assert abs(a - b) < 1e-4, (name, got)
    for name, directed, edges in NEWMAN_UNIT:
        g, w = _weighted_oracle_graph(oracle, directed, edges)
        for wf in (False, True):
            assert g.newman_weighted_closeness_centrality(w, wf, 200) == g.closeness_centrality(wf, 200), name


def test_all_pairs_dijkstra_goldens(oracle):
    # tests/graph/test_dijkstra.py:129-176, tests/digraph/test_dijkstra.py:198-258
    from weighted_cases import ALL_PAIRS, DIJKSTRA_EDGES

    for name, directed, remove, expected in ALL_PAIRS:
        g, w = _weighted_oracle_graph(oracle, directed, DIJKSTRA_EDGES, remove)
        live = sorted(expected)
        rows = g.dijkstra_lengths_rows(w, live)
        got = {s: {t: rows[i, t] for t in range(rows.shape[1]) if t != s and not math.isnan(rows[i, t])}
               for i, s in enumerate(live)}
        assert got == expected, name


# ---------------------------------------------------------------------------------------------
# multigraph semantics from first principles
# ---------------------------------------------------------------------------------------------
def _definition_by_matrix_powers(directed, n, edges, endpoints):
    """Betweenness from its definition, with exact integer path counts: sigma(s,t) is the (s,t) entry
    of A^d(s,t), A holding edge MULTIPLICITIES (a neighbour listed k times adds k paths,
    centrality.rs:433-436; a self-loop is on no shortest path).  Independent of Brandes' recursion."""
    A = np.zeros((n, n), dtype=object)
    for a, b in edges:
        if a == b:
            continue
        A[a, b] += 1
        if not directed:
            A[b, a] += 1
    INF = 10**9
    dist = np.full((n, n), INF, dtype=np.int64)
    sig = np.zeros((n, n), dtype=object)
    P = np.eye(n, dtype=object)
    for i in range(n):
        dist[i, i] = 0
        sig[i, i] = 1
    for d in range(1, n):
        P = P.dot(A)
        new = (dist == INF) & (P != 0)
        if not new.any():
            break
        dist[new] = d
        sig[new] = P[new]
    bc = np.zeros(n)
    ebc = np.zeros(len(edges))
    for s in range(n):
        for t in range(n):
            if s == t or dist[s, t] >= INF:
                continue
            for v in range(n):
                if v not in (s, t) and dist[s, v] + dist[v, t] == dist[s, t]:
                    bc[v] += float(sig[s, v] * sig[v, t]) / float(sig[s, t])
            if endpoints:  # centrality.rs:281-290: the source gets |S|-1, every reached vertex +1
                bc[s] += 1.0
                bc[t] += 1.0
            for e, (a, b) in enumerate(edges):
                if a == b:
                    continue
                for x, y in (((a, b),) if directed else ((a, b), (b, a))):
                    if (dist[s, x] < INF and dist[y, t] < INF
                            and dist[s, x] + 1 + dist[y, t] == dist[s, t]):
                        ebc[e] += float(sig[s, x] * sig[y, t]) / float(sig[s, t])
    if not directed:  # _rescale with normalized = false: 0.5 for undirected graphs, centrality.rs:236-241
        bc /= 2.0
        ebc /= 2.0
    return bc, ebc


@pytest.mark.parametrize("directed", [False, True])
def test_multigraph_semantics_from_first_principles(oracle, directed):
    """Parallel edges count as distinct shortest paths, self-loops never matter, per-edge scores of
    parallel edges are separate: the oracle against the matrix-power definition on random
    multigraphs (the reference pins these semantics only on a handful of tiny cases)."""
    rng = np.random.default_rng(7 + int(directed))
    for _ in range(10):
        n = int(rng.integers(5, 14))
        m = int(rng.integers(n, 3 * n))
        edges = [(int(a), int(b)) for a, b in rng.integers(0, n, size=(m, 2))]
        edges += [edges[int(i)] for i in rng.integers(0, m, size=max(1, m // 4))]  # parallel copies
        g = oracle.OracleGraph.from_edges(directed, edges, num_nodes=n)
        for endpoints in (False, True):
            want_bc, want_ebc = _definition_by_matrix_powers(directed, n, edges, endpoints)
            got = np.array(g.betweenness_centrality(endpoints, False, 1 << 30), dtype=float)
            assert np.allclose(got, want_bc, rtol=1e-12, atol=1e-12)
        got_e = np.array(g.edge_betweenness_centrality(False, 1 << 30), dtype=float)
        assert np.allclose(got_e, want_ebc, rtol=1e-12, atol=1e-12)


def _hop_distances(directed, n, edges, as_undirected=False):
    """All-pairs hop distances by repeated boolean matrix products (no BFS): -1 where unreachable."""
    A = np.zeros((n, n), dtype=bool)
    for a, b in edges:
        A[a, b] = True
        if not directed or as_undirected:
            A[b, a] = True
    dist = np.full((n, n), -1, dtype=np.int64)
    np.fill_diagonal(dist, 0)
    reach = np.eye(n, dtype=bool)
    for d in range(1, n + 1):
        reach = (reach.astype(np.int64) @ A.astype(np.int64)) > 0
        new = reach & (dist < 0)
        if not new.any():
            break
        dist[new] = d
    return dist


@pytest.mark.parametrize("directed", [False, True])
def test_distance_functions_from_first_principles_with_holes(oracle, directed):
    """closeness (incoming distances, Wasserman-Faust scaling), distance_matrix (compacted order,
    null value, as_undirected), group closeness and the average-path-length sums on random
    multigraphs from which nodes were removed -- against distances obtained by boolean matrix
    powers, with the reference's own floating-point expression order (centrality.rs:1209-1220)."""
    rng = np.random.default_rng(31 + int(directed))
    for _ in range(8):
        n0 = int(rng.integers(6, 16))
        m = int(rng.integers(n0, 3 * n0))
        edges0 = [(int(a), int(b)) for a, b in rng.integers(0, n0, size=(m, 2))]
        g = oracle.OracleGraph.from_edges(directed, edges0, num_nodes=n0)
        removed = sorted(int(x) for x in rng.choice(n0, size=2, replace=False))
        for r in removed:
            g.remove_node(r)
        live = [v for v in range(n0) if v not in removed]
        pos = {v: i for i, v in enumerate(live)}
        n = len(live)
        edges = [(pos[a], pos[b]) for a, b in edges0 if a in pos and b in pos]
        dist = _hop_distances(directed, n, edges)
        # closeness of v: distances TO v (Reversed graph, centrality.rs:1208)
        for wf in (True, False):
            got = g.closeness_centrality(wf, 1 << 30)
            assert [x is None for x in got] == [v not in pos for v in range(len(got))]
            for v in live:
                col = dist[:, pos[v]]
                reach = int((col >= 0).sum())
                total = int(col[col >= 0].sum())
                if reach == 1:
                    want = 0.0
                else:
                    want = float(reach - 1) / float(total)
                    if wf:
                        want = want * float(reach - 1) / float(n - 1)
                assert got[v] == want, (v, got[v], want)
        for as_undirected in (False, True):
            d2 = _hop_distances(directed, n, edges, as_undirected)
            want = np.where(d2 >= 0, d2, -7).astype(np.float64)
            np.fill_diagonal(want, 0.0)
            assert np.array_equal(g.distance_matrix(300, as_undirected, -7.0), want)
            off = ~np.eye(n, dtype=bool)
            assert g.distance_sum(300, as_undirected) == (int(d2[off & (d2 > 0)].sum()),
                                                          int((off & (d2 > 0)).sum()))
        group = [live[0], live[n // 2]]
        to_group = np.stack([dist[:, pos[x]] for x in group])          # incoming distances again
        to_group = np.where(to_group < 0, 10**9, to_group).min(axis=0)
        others = [i for i in range(n) if live[i] not in group]
        total = sum(int(to_group[i]) for i in others if to_group[i] < 10**9)
        want = 0.0 if total == 0 else float(len(others)) / float(total)
        assert g.group_closeness_centrality(group) == want
