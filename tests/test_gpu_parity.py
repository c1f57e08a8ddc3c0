"""GPU parity: the CUDA path (through the C-ABI, via the rustworkx-shaped Python API) against the CPU
oracle on the same seeded inputs, plus the reference's own golden vectors.

Tolerances (BASELINE.json north_star): distance_matrix and closeness bit-exact; betweenness and edge
betweenness within 1e-9 relative (1e-12 absolute floor for zeros).
"""

import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-9
ATOL = 1e-12


@pytest.fixture(scope="module")
def rx():
    import rustworkx_b200 as rx
    from rustworkx_b200 import _lib

    if _lib.device_count() == 0:
        pytest.fail("no CUDA device: the gpu-marked tests need a B200 (there is no CPU fallback)")
    return rx


def assert_close(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    err = np.abs(got - want)
    tol = ATOL + RTOL * np.abs(want)
    bad = np.flatnonzero(~(err <= tol))
    assert bad.size == 0, f"{bad.size} mismatches, first at {bad[:5]}: got {got[bad[:5]]} want {want[bad[:5]]}"


def random_graph(rx, directed, n, m, seed, holes=0, multi=True, loops=True):
    """Random multigraph with optional self-loops, parallel edges, removed nodes and removed edges."""
    rng = np.random.default_rng(seed)
    g = rx.PyDiGraph() if directed else rx.PyGraph()
    g.add_nodes_from(range(n))
    src = rng.integers(0, n, size=m)
    dst = rng.integers(0, n, size=m)
    if not loops:
        keep = src != dst
        src, dst = src[keep], dst[keep]
    g.add_edges_from_no_data(list(zip(src.tolist(), dst.tolist())))
    if multi and m:
        k = max(1, m // 10)
        idx = rng.integers(0, len(src), size=k)
        g.add_edges_from_no_data(list(zip(src[idx].tolist(), dst[idx].tolist())))
    for node in rng.choice(n, size=min(holes, n), replace=False).tolist():
        g.remove_node(int(node))
    eids = list(g.edge_indices())
    for e in rng.choice(eids, size=min(holes, len(eids)), replace=False).tolist() if eids else []:
        g.remove_edge_from_index(int(e))
    # re-use some freed slots (index reuse after removal)
    for _ in range(holes // 2):
        g.add_node(None)
    live = list(g.node_indices())
    for _ in range(holes):
        a, b = rng.choice(live, size=2).tolist()
        g.add_edge(int(a), int(b), None)
    return g


def oracle_graph(oracle, g):
    return oracle.OracleGraph.from_graph(g)


def dense_from_mapping(mapping, bound):
    out = np.zeros(bound)
    present = np.zeros(bound, dtype=bool)
    for k, v in mapping.items():
        out[k] = v
        present[k] = True
    return out, present


# ------------------------------------------------------------------------------------ goldens
def test_reference_goldens_betweenness(rx):
    # tests/graph/test_centrality.py:20-103 and tests/digraph/test_centrality.py:20-132
    for cls, deleted in [(rx.PyGraph, False), (rx.PyGraph, True), (rx.PyDiGraph, False),
                         (rx.PyDiGraph, True)]:
        g = cls()
        a, b, c = g.add_node("A"), g.add_node("B"), g.add_node("C")
        c0 = g.add_node("C0") if deleted else None
        d = g.add_node("D")
        g.add_edges_from([(a, b, 1), (b, c, 1), (c, d, 1)])
        if deleted:
            g.remove_node(c0)
        keys = [0, 1, 2, 4] if deleted else [0, 1, 2, 3]
        if cls is rx.PyDiGraph:
            norm = [0.0, 0.3333333333333333, 0.3333333333333333, 0.0]
            endp = [0.25, 0.41666666666666663, 0.41666666666666663, 0.25]
        else:
            norm = [0.0, 0.6666666666666666, 0.6666666666666666, 0.0]
            endp = [0.5, 0.8333333333333333, 0.8333333333333333, 0.5]
        for threshold in (50, 1):
            assert rx.betweenness_centrality(g, parallel_threshold=threshold) == dict(zip(keys, norm))
            assert rx.betweenness_centrality(g, endpoints=True, parallel_threshold=threshold) == dict(zip(keys, endp))
            assert rx.betweenness_centrality(g, endpoints=False, normalized=False,
                                             parallel_threshold=threshold) == dict(zip(keys, [0.0, 2.0, 2.0, 0.0]))


def test_reference_goldens_core_doctests(rx):
    # rustworkx-core/src/centrality.rs:58-66, :164-170, :1146-1164
    g = rx.PyGraph()
    g.extend_from_edge_list([(0, 4), (1, 2), (2, 3), (3, 4), (1, 4)])
    assert list(rx.betweenness_centrality(g, normalized=True, endpoints=True).values()) == pytest.approx(
        [0.4, 0.5, 0.45, 0.5, 0.75], rel=1e-12)
    assert rx.closeness_centrality(g) == {0: 1 / 2, 1: 2 / 3, 2: 4 / 7, 3: 2 / 3, 4: 4 / 5}
    dg = rx.PyDiGraph()
    dg.extend_from_edge_list([(0, 4), (1, 2), (2, 3), (3, 4), (1, 4)])
    assert rx.closeness_centrality(dg) == {0: 0.0, 1: 0.0, 2: 1 / 4, 3: 1 / 3, 4: 4 / 5}
    g = rx.PyGraph()
    g.extend_from_edge_list([(0, 4), (1, 2), (1, 3), (2, 3), (3, 4), (1, 4)])
    assert rx.edge_betweenness_centrality(g, normalized=False) == {0: 4.0, 1: 2.0, 2: 1.0, 3: 2.0, 4: 3.0, 5: 3.0}


def test_reference_goldens_holes(rx):
    # rustworkx-core/src/centrality.rs:628-696
    g = rx.PyGraph()
    n0, d0, n1, d1, n2, d2, n3 = (g.add_node(None) for _ in range(7))
    for d in (d0, d1, d2):
        g.remove_node(d)
    g.add_edges_from_no_data([(n0, n1), (n1, n2), (n2, n3), (n3, n0)])
    g.remove_edge_from_index(1)
    assert rx.betweenness_centrality(g, normalized=False) == {0: 2.0, 2: 0.0, 4: 0.0, 6: 2.0}
    assert rx.edge_betweenness_centrality(g, normalized=False) == {0: 3.0, 2: 3.0, 3: 4.0}


def test_reference_goldens_edge_betweenness(rx):
    # tests/graph/test_centrality.py:194-248, tests/digraph/test_centrality.py:227-259
    gen = rx.generators

    def check(g, expected, normalized=True):
        got = rx.edge_betweenness_centrality(g, normalized=normalized)
        assert list(got.keys()) == list(range(len(expected)))
        for k, v in got.items():
            assert v == pytest.approx(expected[k], abs=1e-7)

    check(gen.mesh_graph(5), [0.1] * 10)
    check(gen.path_graph(5), [0.4, 0.6, 0.6, 0.4])
    check(gen.cycle_graph(5), [0.3] * 5)
    check(gen.full_rary_tree(2, 7), [12, 12, 6, 6, 6, 6], False)
    check(gen.path_graph(5), [4, 6, 6, 4], False)
    g = rx.PyGraph()
    g.add_nodes_from(range(10))
    g.add_edges_from([(0, 1, 1), (0, 2, 1), (0, 3, 1), (0, 4, 1), (3, 5, 1), (4, 6, 1), (5, 7, 1),
                      (6, 8, 1), (7, 8, 1), (8, 9, 1)])
    check(g, [9, 9, 12, 15, 11, 14, 10, 13, 9, 9], False)
    check(gen.directed_mesh_graph(5), [0.05] * 20)
    check(gen.directed_path_graph(5), [0.2, 0.3, 0.3, 0.2])
    check(gen.directed_cycle_graph(5), [0.5] * 5)
    check(gen.full_rary_tree(2, 7).to_directed(), [12, 12, 12, 12, 6, 6, 6, 6, 6, 6, 6, 6], False)
    check(gen.directed_path_graph(5), [4, 6, 6, 4], False)


def test_reference_goldens_closeness(rx):
    # rustworkx-core/tests/centrality.rs:17-120; tests/{graph,digraph}/test_centrality.py closeness
    def g_from(cls, edges, n=None):
        g = cls()
        g.add_nodes_from(range(n if n is not None else max(max(e) for e in edges) + 1))
        g.add_edges_from_no_data(edges)
        return g

    g = g_from(rx.PyGraph, [(1, 2), (2, 3), (3, 4), (1, 4)])
    assert rx.closeness_centrality(g) == {0: 0.0, 1: 0.5625, 2: 0.5625, 3: 0.5625, 4: 0.5625}
    g = g_from(rx.PyGraph, [(0, 1), (1, 2), (2, 3), (4, 5), (5, 6)])
    assert rx.closeness_centrality(g) == {0: 1 / 4, 1: 3 / 8, 2: 3 / 8, 3: 1 / 4, 4: 2 / 9, 5: 1 / 3, 6: 2 / 9}
    assert rx.closeness_centrality(g, wf_improved=False) == {0: 1 / 2, 1: 3 / 4, 2: 3 / 4, 3: 1 / 2,
                                                             4: 2 / 3, 5: 1.0, 6: 2 / 3}
    g = g_from(rx.PyDiGraph, [(0, 1), (1, 2)])
    assert rx.closeness_centrality(g) == {0: 0.0, 1: 1 / 2, 2: 2 / 3}
    assert rx.closeness_centrality(rx.generators.mesh_graph(5)) == {i: 1.0 for i in range(5)}
    assert rx.closeness_centrality(rx.generators.path_graph(3)) == {0: 2 / 3, 1: 1.0, 2: 2 / 3}
    for cls in (rx.PyGraph, rx.PyDiGraph):
        g = cls()
        a, b, c, c0, d = (g.add_node(x) for x in "ABCXD")
        g.add_edges_from([(a, b, 1), (b, c, 1), (c, d, 1)])
        g.remove_node(c0)
        for thr in (50, 1):
            if cls is rx.PyGraph:
                assert rx.closeness_centrality(g, parallel_threshold=thr) == {0: 0.5, 1: 0.75, 2: 0.75, 4: 0.5}
                assert rx.closeness_centrality(g, wf_improved=False) == {0: 0.5, 1: 0.75, 2: 0.75, 4: 0.5}
            else:
                assert rx.closeness_centrality(g, parallel_threshold=thr) == {0: 0.0, 1: 1 / 3, 2: 4 / 9, 4: 0.5}
                assert rx.closeness_centrality(g, wf_improved=False) == {0: 0.0, 1: 1.0, 2: 2 / 3, 4: 0.5}


def test_reference_goldens_distance_matrix(rx):
    # tests/graph/test_dist_matrix.py:21-102, tests/digraph/test_dist_matrix.py:21-140
    from test_oracle_golden import CYCLE7, CYCLE7_DIRECTED, CYCLE7_UNDIRECTED, _nan_expected

    g = rx.PyGraph()
    g.add_nodes_from(range(7))
    g.add_edges_from_no_data(CYCLE7)
    assert np.array_equal(rx.graph_distance_matrix(g), CYCLE7_UNDIRECTED)
    assert np.array_equal(rx.graph_distance_matrix(g, parallel_threshold=5), CYCLE7_UNDIRECTED)
    g.add_node(7)
    assert np.array_equal(rx.graph_distance_matrix(g, null_value=np.nan), _nan_expected(), equal_nan=True)
    dg = rx.PyDiGraph()
    dg.add_nodes_from(range(7))
    dg.add_edges_from_no_data(CYCLE7)
    assert np.array_equal(rx.digraph_distance_matrix(dg), CYCLE7_DIRECTED)
    assert np.array_equal(rx.digraph_distance_matrix(dg, as_undirected=True), CYCLE7_UNDIRECTED)
    dg.add_node(7)
    assert np.array_equal(rx.distance_matrix(dg, as_undirected=True, null_value=np.nan),
                          _nan_expected(), equal_nan=True)
    g = rx.generators.path_graph(4)
    g.remove_node(0)
    assert np.array_equal(rx.graph_distance_matrix(g), np.array([[0., 1, 2], [1, 0, 1], [2, 1, 0]]))
    dg = rx.generators.directed_path_graph(4)
    dg.remove_node(0)
    assert np.array_equal(rx.digraph_distance_matrix(dg), np.array([[0., 1, 2], [0, 0, 1], [0, 0, 0]]))


def test_dispatch_and_errors(rx):
    # tests/test_dispatch.py:29-38,99-101
    with pytest.raises(TypeError):
        rx.betweenness_centrality("not a graph")
    with pytest.raises(TypeError):
        rx.distance_matrix(rx.PyGraph(), as_undirected=True)  # graph_distance_matrix has no such kwarg
    with pytest.raises(OverflowError):
        rx.betweenness_centrality(rx.PyGraph(), parallel_threshold=-1)
    with pytest.raises(TypeError):
        rx.graph_betweenness_centrality(rx.PyDiGraph())
    assert rx.betweenness_centrality(rx.PyGraph()) == {}
    assert rx.closeness_centrality(rx.PyDiGraph()) == {}
    assert rx.distance_matrix(rx.PyGraph()).shape == (0, 0)
    g = rx.PyGraph()
    g.add_node(None)
    assert rx.betweenness_centrality(g) == {0: 0.0}
    assert rx.closeness_centrality(g) == {0: 0.0}
    assert np.array_equal(rx.distance_matrix(g), np.zeros((1, 1)))


# ------------------------------------------------------------------------------------ vs oracle
CASES = [
    # directed, n, m, seed, holes
    (False, 40, 60, 1, 0),
    (True, 40, 90, 2, 0),
    (False, 97, 300, 3, 7),
    (True, 97, 400, 4, 9),
    (False, 300, 450, 5, 20),   # sparse: several components
    (True, 300, 700, 6, 20),
    (False, 1000, 4000, 7, 0),
    (True, 1000, 6000, 8, 30),
]


@pytest.mark.parametrize("directed,n,m,seed,holes", CASES)
def test_betweenness_vs_oracle(rx, oracle, directed, n, m, seed, holes):
    g = random_graph(rx, directed, n, m, seed, holes)
    og = oracle_graph(oracle, g)
    nb, _ = g._bounds()
    for normalized in (True, False):
        for endpoints in (False, True):
            want, present = og.betweenness_centrality(endpoints, normalized, 1 << 30, as_arrays=True)
            got = rx.betweenness_centrality(g, normalized=normalized, endpoints=endpoints)
            gd, gp = dense_from_mapping(got, nb)
            assert np.array_equal(gp, present.astype(bool))
            assert list(got.keys()) == list(g.node_indices())
            assert_close(gd, want)


@pytest.mark.parametrize("directed,n,m,seed,holes", CASES)
def test_edge_betweenness_vs_oracle(rx, oracle, directed, n, m, seed, holes):
    g = random_graph(rx, directed, n, m, seed, holes)
    og = oracle_graph(oracle, g)
    _, eb = g._bounds()
    for normalized in (True, False):
        want, present = og.edge_betweenness_centrality(normalized, 1 << 30, as_arrays=True)
        got = rx.edge_betweenness_centrality(g, normalized=normalized)
        gd, gp = dense_from_mapping(got, eb)
        assert np.array_equal(gp, present.astype(bool))
        assert list(got.keys()) == list(g.edge_indices())
        assert_close(gd, want)


@pytest.mark.parametrize("directed,n,m,seed,holes", CASES)
def test_closeness_vs_oracle_bit_exact(rx, oracle, directed, n, m, seed, holes):
    g = random_graph(rx, directed, n, m, seed, holes)
    og = oracle_graph(oracle, g)
    nb, _ = g._bounds()
    for wf in (True, False):
        want, present = og.closeness_centrality(wf, 1 << 30, as_arrays=True)
        got = rx.closeness_centrality(g, wf_improved=wf)
        gd, gp = dense_from_mapping(got, nb)
        assert np.array_equal(gp, present.astype(bool))
        assert np.array_equal(gd, want)  # bit exact


@pytest.mark.parametrize("directed,n,m,seed,holes", CASES)
def test_distance_matrix_vs_oracle_bit_exact(rx, oracle, directed, n, m, seed, holes):
    g = random_graph(rx, directed, n, m, seed, holes)
    og = oracle_graph(oracle, g)
    for null in (0.0, math.nan, -1.0):
        if directed:
            for und in (False, True):
                want = og.distance_matrix(300, und, null)
                got = rx.digraph_distance_matrix(g, as_undirected=und, null_value=null)
                assert got.dtype == np.float64 and got.flags.c_contiguous
                assert np.array_equal(got, want, equal_nan=True)
        else:
            want = og.distance_matrix(300, True, null)
            got = rx.graph_distance_matrix(g, null_value=null)
            assert np.array_equal(got, want, equal_nan=True)


def test_config1_gnp5000_betweenness(rx, oracle):
    # BASELINE.json configs[0]: betweenness on undirected_gnp_random_graph(5000, 0.002, seed=42)
    g = rx.undirected_gnp_random_graph(5000, 0.002, seed=42)
    og = oracle_graph(oracle, g)
    for normalized, endpoints in ((True, False), (False, True)):
        want, _ = og.betweenness_centrality(endpoints, normalized, 50, as_arrays=True)
        got = rx.betweenness_centrality(g, normalized=normalized, endpoints=endpoints)
        assert_close(np.array(list(got.values())), want)


def test_config2_heavy_hex(rx, oracle):
    # BASELINE.json configs[1]: distance_matrix + closeness on heavy_hex_graph(d); d=15 on the CPU
    # side keeps the oracle's O(n^2/64)-per-level bitset BFS within seconds, d=51 uses invariants
    g = rx.generators.heavy_hex_graph(15)
    og = oracle_graph(oracle, g)
    assert np.array_equal(rx.distance_matrix(g), og.distance_matrix(300, True, 0.0))
    for wf in (True, False):
        want, _ = og.closeness_centrality(wf, 50, as_arrays=True)
        assert np.array_equal(np.array(list(rx.closeness_centrality(g, wf_improved=wf).values())), want)
    g = rx.generators.heavy_hex_graph(51)
    n = g.num_nodes()
    dm = rx.distance_matrix(g)
    assert dm.shape == (n, n)
    assert np.array_equal(dm, dm.T) and np.all(np.diag(dm) == 0.0)
    assert np.all(dm[~np.eye(n, dtype=bool)] >= 1.0)
    src, dst = g._esrc[: g._elen], g._edst[: g._elen]
    assert np.all(dm[src, dst] == 1.0)
    # closeness must equal (n-1)/rowsum * (n-1)/(n-1) computed from the matrix (connected graph)
    cc = np.array(list(rx.closeness_centrality(g).values()))
    rowsum = dm.sum(axis=0)
    assert np.array_equal(cc, ((n - 1) / rowsum) * (n - 1) / (n - 1))
    sample = np.arange(0, n, 97, dtype=np.uint32)
    want, _ = oracle_graph(oracle, g).closeness_partial(sample, True, threads=oracle.max_threads())
    assert np.array_equal(cc[sample], want[sample])


def test_sampled_sources_large_ba(rx, oracle):
    # config 3 shape at reduced size: BA graph, un-rescaled partial sums over sampled sources
    from rustworkx_b200 import device_snapshot

    g = rx.barabasi_albert_graph(20000, 8, seed=1)
    og = oracle_graph(oracle, g)
    rng = np.random.default_rng(0)
    sources = rng.choice(20000, size=96, replace=False).astype(np.uint32)
    dg = device_snapshot(g)
    want, st = og.betweenness_partial(sources, False, threads=oracle.max_threads())
    got = dg.betweenness_partial(sources, False)
    assert_close(got, want)
    gst = dg.stats()
    assert gst["te"] == st["te"] and gst["te_dag"] == st["te_dag"] and gst["v"] == st["v"]
    want_e, _ = og.edge_betweenness_partial(sources, threads=oracle.max_threads())
    assert_close(dg.edge_betweenness_partial(sources), want_e)
    # duplicates in the source list are counted twice, as additive contributions
    dup = np.concatenate([sources[:5], sources[:5]])
    assert_close(dg.betweenness_partial(dup, True), 2 * og.betweenness_partial(sources[:5], True)[0])


def test_grid_closeness_sample(rx, oracle):
    # config 2 high-diameter shape at reduced size, bit exact against the oracle
    g = rx.generators.grid_graph(120, 90)
    og = oracle_graph(oracle, g)
    want, _ = og.closeness_centrality(True, 50, as_arrays=True)
    got = np.array(list(rx.closeness_centrality(g).values()))
    assert np.array_equal(got, want)


def test_networkx_golden_fixture(rx):
    # committed vectors from an independent implementation (tests/golden/make_golden.py)
    import json
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nx_golden.json")
    for case in json.load(open(path))["cases"]:
        g = rx.PyDiGraph() if case["directed"] else rx.PyGraph()
        g.add_nodes_from(range(case["n"]))
        g.add_edges_from_no_data([tuple(e) for e in case["edges"]])
        bc = rx.betweenness_centrality(g, normalized=case["normalized"], endpoints=case["endpoints"])
        assert_close(list(bc.values()), case["betweenness"])
        ebc = rx.edge_betweenness_centrality(g, normalized=case["normalized"])
        assert_close(list(ebc.values()), case["edge_betweenness"])
        cc = rx.closeness_centrality(g)
        assert np.allclose(list(cc.values()), case["closeness_wf"], rtol=4e-16, atol=0)
        if case["directed"]:
            dm = rx.digraph_distance_matrix(g, null_value=-1.0)
        else:
            dm = rx.graph_distance_matrix(g, null_value=-1.0)
        assert np.array_equal(dm, np.array(case["distance_matrix"]))


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_distance_regimes_agree_with_oracle(rx, oracle, mode):
    """The distance-only regimes (1: 32-source words top-down, 2: CTA-per-source in shared memory,
    3: 128-source masks bottom-up) are forced in turn and must be bit-exact; auto mode picks between
    them with a pilot group.  (The grid forces regime 3 through its level-queue overflow path.)"""
    from rustworkx_b200 import _lib

    _lib.set_option("bfs_mode", mode)
    try:
        for g in (rx.generators.grid_graph(37, 23), rx.generators.heavy_hex_graph(7),
                  random_graph(rx, True, 500, 1500, 11, holes=10),
                  random_graph(rx, False, 700, 900, 12, holes=5)):
            og = oracle_graph(oracle, g)
            nb, _ = g._bounds()
            for wf in (True, False):
                want, present = og.closeness_centrality(wf, 1 << 30, as_arrays=True)
                gd, gp = dense_from_mapping(rx.closeness_centrality(g, wf_improved=wf), nb)
                assert np.array_equal(gd, want) and np.array_equal(gp, present.astype(bool))
            if isinstance(g, rx.PyDiGraph):
                for und in (False, True):
                    assert np.array_equal(rx.digraph_distance_matrix(g, as_undirected=und, null_value=-1.0),
                                          og.distance_matrix(300, und, -1.0))
                    assert rx.digraph_unweighted_average_shortest_path_length(g, as_undirected=und, disconnected=True) \
                        == og.average_shortest_path_length(300, und, True)
            else:
                assert np.array_equal(rx.graph_distance_matrix(g, null_value=math.nan),
                                      og.distance_matrix(300, True, math.nan), equal_nan=True)
        # duplicates in a partial call are computed once and copied
        g = random_graph(rx, True, 300, 900, 21, holes=4)
        og = oracle_graph(oracle, g)
        live = np.asarray(g.node_indices(), dtype=np.uint32)
        src = np.concatenate([live[:150], live[:20], live[100:140]])
        want, _ = og.closeness_partial(np.unique(src), True)
        from rustworkx_b200 import device_snapshot
        assert np.array_equal(device_snapshot(g).closeness_partial(src, True), want[src])
    finally:
        _lib.set_option("bfs_mode", 0)


def test_device_and_host_csr_builders_agree(rx, oracle):
    """rxc_graph_snapshot builds the CSR on the device by default; the host counting sort is kept as
    a cross-check.  Results must be identical for the exact ops and within tolerance for BC."""
    from rustworkx_b200 import _lib

    for g in (random_graph(rx, True, 3000, 20000, 21, holes=40),
              random_graph(rx, False, 3000, 12000, 22, holes=40),
              rx.barabasi_albert_graph(30000, 6, seed=2)):
        res = {}
        for flag in (1, 0):
            _lib.set_option("device_csr", flag)
            try:
                dg = _lib.DeviceGraph.from_graph(g)
                src = np.asarray(g.node_indices()[:64], dtype=np.uint32)
                res[flag] = (dg.closeness(True), dg.betweenness_partial(src, True),
                             dg.edge_betweenness_partial(src),
                             dg.distance_matrix_rows(0, 8, as_undirected=True, null_value=-1.0))
                dg.close()
            finally:
                _lib.set_option("device_csr", 1)
        assert np.array_equal(res[1][0], res[0][0])
        assert_close(res[1][1], res[0][1])
        assert_close(res[1][2], res[0][2])
        assert np.array_equal(res[1][3], res[0][3])
    # invalid inputs are rejected by the device builder too
    with pytest.raises(_lib.RxCudaError):
        _lib.DeviceGraph(4, np.array([0, 1, 2], np.uint32), 1, np.array([0], np.uint32),
                         np.array([3], np.uint32), np.array([0], np.uint32), False)


def test_pinned_host_buffers_and_positional_edge_ids(rx, oracle):
    """rxc_host_alloc memory is copied by direct DMA (no staging) and edge_id = NULL means "the
    position in the list": both must give what pageable buffers with explicit ids give, and the
    oracle's numbers."""
    from rustworkx_b200 import _lib

    g = rx.barabasi_albert_graph(20000, 5, seed=11)
    s = g._snapshot_arrays()
    assert s["edge_id_is_position"]
    src = np.arange(0, 20000, 157, dtype=np.uint32)
    plain = _lib.DeviceGraph(s["n_bound"], s["live_nodes"], s["e_bound"], s["edge_src"], s["edge_dst"],
                             s["edge_id"], s["directed"])
    pinned = _lib.DeviceGraph(s["n_bound"], _lib.pinned_copy(s["live_nodes"]), s["e_bound"],
                              _lib.pinned_copy(s["edge_src"]), _lib.pinned_copy(s["edge_dst"]), None,
                              s["directed"])
    assert np.array_equal(pinned.edge_id, s["edge_id"])
    out = _lib.pinned_empty(s["n_bound"], np.float64)
    out[:] = -1.0
    got = pinned.betweenness_partial(src, False, out=out)
    assert got.base is not None and np.shares_memory(got, out)
    assert_close(got, plain.betweenness_partial(src, False))
    assert_close(pinned.edge_betweenness_partial(src), plain.edge_betweenness_partial(src))
    assert np.array_equal(pinned.closeness(True), plain.closeness(True))
    og = oracle.OracleGraph.from_graph(g)
    want, _ = og.closeness_centrality(True, 50, as_arrays=True)
    assert np.array_equal(pinned.closeness(True), want)
    # a graph that lost an edge keeps explicit ids
    g.remove_edge_from_index(3)
    assert not g._snapshot_arrays()["edge_id_is_position"]
    with pytest.raises(ValueError):
        pinned.betweenness_partial(src, False, out=np.zeros(5))
    plain.close()
    pinned.close()


def test_full_size_invariant_ba_1m(rx):
    """BASELINE.json configs[2] at full size (BA(10^6, 8)): size-independent property instead of the
    oracle.  Every shortest s-t path of length d has d-1 interior vertices whose pair-dependencies
    sum to d-1, so  sum_v partial_bc[v] = sum_{s in S} sum_t (d(s,t) - 1);  the right-hand side comes
    from the (bit-exact) distance kernel via closeness = (r-1)/sum_d on a connected graph."""
    from rustworkx_b200 import device_snapshot

    n = 1_000_000
    g = rx.barabasi_albert_graph(n, 8, seed=1)
    dg = device_snapshot(g)
    src = np.random.default_rng(5).choice(n, size=256, replace=False).astype(np.uint32)
    partial = dg.betweenness_partial(src, False)
    st = dg.stats()
    assert st["v"] == 256 * n  # connected: every source reaches every vertex
    assert st["te"] == 256 * 2 * g.num_edges()
    cc = dg.closeness_partial(src, wf_improved=False)
    dsum = np.rint((n - 1) / cc)  # integers well below 2^53
    want = float(np.sum(dsum - (n - 1)))
    got = float(np.sum(partial))
    assert abs(got - want) <= 1e-9 * want
    # endpoints add exactly (reached - 1) to the source and 1 to every other reached vertex
    with_ep = dg.betweenness_partial(src[:128], True)
    without = dg.betweenness_partial(src[:128], False)
    assert abs(float(np.sum(with_ep - without)) - 2.0 * 128 * (n - 1)) <= 1e-6 * 128 * n
    # edge betweenness: every s-t path of length d crosses d edges
    epart = dg.edge_betweenness_partial(src[:128])
    want_e = float(np.sum(dsum[:128]))
    assert abs(float(np.sum(epart)) - want_e) <= 1e-9 * want_e


def test_average_shortest_path_length(rx, oracle):
    # tests/graph/test_avg_shortest_path.py, tests/digraph/test_avg_shortest_path.py
    gen = rx.generators
    g = rx.PyGraph()
    g.extend_from_edge_list([(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (3, 6), (6, 7)])
    assert rx.graph_unweighted_average_shortest_path_length(g) == pytest.approx(2.5714285714285716, abs=1e-7)
    assert rx.unweighted_average_shortest_path_length(gen.cycle_graph(7)) == pytest.approx(2, abs=1e-7)
    assert rx.unweighted_average_shortest_path_length(gen.grid_graph(30, 11)) == pytest.approx(
        13.666666666666666, abs=1e-7)
    assert math.isnan(rx.unweighted_average_shortest_path_length(rx.PyGraph()))
    iso = rx.PyGraph()
    iso.add_nodes_from(range(32))
    assert math.isinf(rx.unweighted_average_shortest_path_length(iso))
    assert math.isnan(rx.unweighted_average_shortest_path_length(iso, disconnected=True))
    part = gen.cycle_graph(32)
    part.add_nodes_from(range(32))
    assert math.isinf(rx.unweighted_average_shortest_path_length(part))
    assert rx.unweighted_average_shortest_path_length(part, disconnected=True) == pytest.approx(8192 / 992, abs=1e-7)
    assert math.isinf(rx.unweighted_average_shortest_path_length(gen.directed_path_graph(5)))
    assert rx.digraph_unweighted_average_shortest_path_length(
        gen.directed_path_graph(5), as_undirected=True) == pytest.approx(2, abs=1e-7)
    # integer sums are exact, so the ratio is bit-identical to the oracle
    for directed, seed in ((False, 31), (True, 32)):
        g = random_graph(rx, directed, 800, 2500, seed, holes=15)
        og = oracle_graph(oracle, g)
        for disc in (False, True):
            for und in ((False, True) if directed else (True,)):
                want = og.average_shortest_path_length(300, und, disc)
                if directed:
                    got = rx.digraph_unweighted_average_shortest_path_length(g, as_undirected=und, disconnected=disc)
                else:
                    got = rx.unweighted_average_shortest_path_length(g, disconnected=disc)
                assert (math.isnan(want) and math.isnan(got)) or want == got


def test_cta_per_source_with_global_bitmap(rx, oracle):
    """n > 1.3 M: the visited bitmap no longer fits in shared memory and moves to a per-CTA slab in
    global memory; results stay bit-exact (grid 1200 x 1200, a few sampled sources)."""
    from rustworkx_b200 import _lib, device_snapshot

    g = rx.generators.grid_graph(1200, 1200)
    og = oracle_graph(oracle, g)
    src = np.array([0, 7, 1199, 720600, 1439999, 333333], dtype=np.uint32)
    want, _ = og.closeness_partial(src, True, threads=oracle.max_threads())
    dg = device_snapshot(g)
    for mode in (0, 2):
        _lib.set_option("bfs_mode", mode)
        try:
            got = dg.closeness_partial(src, True)
        finally:
            _lib.set_option("bfs_mode", 0)
        assert np.array_equal(got, want[src])


# ------------------------------------------------------------------------------------ group centralities (8f #2)
def _graph_from_case(rx, directed, n, edges):
    g = rx.PyDiGraph() if directed else rx.PyGraph()
    g.add_nodes_from(range(n))
    g.add_edges_from_no_data(edges)
    return g


def test_reference_goldens_group_closeness(rx):
    from group_cases import CLOSENESS

    for name, directed, n, edges, group, expected in CLOSENESS:
        g = _graph_from_case(rx, directed, n, edges)
        typed = rx.digraph_group_closeness_centrality if directed else rx.graph_group_closeness_centrality
        assert abs(typed(g, group) - expected) < 1e-10, name
        assert rx.group_closeness_centrality(g, group) == typed(g, group)


def test_reference_goldens_group_betweenness(rx):
    from group_cases import BETWEENNESS

    for name, directed, n, edges, group, normalized, expected in BETWEENNESS:
        g = _graph_from_case(rx, directed, n, edges)
        typed = rx.digraph_group_betweenness_centrality if directed else rx.graph_group_betweenness_centrality
        for threshold in (50, 1):
            got = typed(g, group, normalized=normalized, parallel_threshold=threshold)
            assert abs(got - expected) < 1e-10, name
        assert abs(rx.group_betweenness_centrality(g, group, normalized) - expected) < 1e-10, name
    # rustworkx-core/src/centrality.rs:2338-2360: a one-node group equals that node's betweenness
    g = _graph_from_case(rx, True, 4, [(0, 1), (1, 2), (2, 3)])
    for normalized in (False, True):
        per_node = rx.betweenness_centrality(g, normalized=normalized)
        for v in range(4):
            assert abs(rx.group_betweenness_centrality(g, [v], normalized) - per_node[v]) < 1e-10


def test_group_errors(rx):
    # tests/graph/test_centrality.py:362-365,407-410; src/centrality.rs:1216-1222
    g = rx.generators.path_graph(3)
    with pytest.raises(ValueError):
        rx.graph_group_closeness_centrality(g, [10])
    with pytest.raises(ValueError):
        rx.graph_group_betweenness_centrality(g, [10])
    g.remove_node(1)
    with pytest.raises(ValueError):
        rx.group_betweenness_centrality(g, [1])
    with pytest.raises(TypeError):
        rx.graph_group_betweenness_centrality(rx.PyDiGraph(), [0])
    with pytest.raises(OverflowError):
        rx.group_closeness_centrality(g, [-1])
    assert rx.group_betweenness_centrality(g, [], normalized=False) == 0.0
    assert rx.group_closeness_centrality(g, [0, 2]) == 0.0  # group covers every node


@pytest.mark.parametrize("directed,n,m,seed,holes", CASES)
def test_group_centralities_vs_oracle(rx, oracle, directed, n, m, seed, holes):
    g = random_graph(rx, directed, n, m, seed, holes)
    og = oracle_graph(oracle, g)
    live = np.asarray(g.node_indices())
    rng = np.random.default_rng(seed + 100)
    for size in (1, 2, 5, max(2, len(live) // 3)):
        group = rng.choice(live, size=size, replace=False).tolist()
        group = group + group[:1]  # a duplicate counts once
        assert rx.group_closeness_centrality(g, group) == og.group_closeness_centrality(group)  # bit exact
        for normalized in (False, True):
            want = og.group_betweenness_centrality(group, normalized, 1 << 30)
            got = rx.group_betweenness_centrality(g, group, normalized=normalized)
            assert abs(got - want) <= ATOL + RTOL * abs(want), (size, normalized, got, want)


def test_group_betweenness_sampled_sources_ba(rx, oracle):
    """BA(100k, 8): un-normalised partial sums over 200 sampled sources against the oracle, work
    counters included (TE and V count both BFSs of every source)."""
    from rustworkx_b200 import device_snapshot

    g = rx.barabasi_albert_graph(100_000, 8, seed=3)
    og = oracle_graph(oracle, g)
    rng = np.random.default_rng(5)
    group = rng.choice(100_000, size=50, replace=False).astype(np.uint32)
    group[:3] = [0, 1, 2]  # include the oldest hubs
    sources = rng.choice(100_000, size=200, replace=False).astype(np.uint32)
    want, wst = og.group_betweenness_partial(group, sources, threads=oracle.max_threads())
    dg = device_snapshot(g)
    got = dg.group_betweenness_partial(group, sources)
    assert abs(got - want) <= RTOL * abs(want)
    st = dg.stats()
    assert st["te"] == wst["te"] and st["v"] == wst["v"]


# ------------------------------------------------------------------------------------ weighted shortest paths (8f #4)
def _weighted_graph(rx, directed, edges, remove=None):
    g = rx.PyDiGraph() if directed else rx.PyGraph()
    n = max(max(a, b) for a, b, _ in edges) + 1
    g.add_nodes_from(range(n))
    g.add_edges_from([(a, b, w) for a, b, w in edges])
    if remove is not None:
        g.remove_node(remove)
    return g


def test_reference_goldens_newman_weighted_closeness(rx):
    from weighted_cases import NEWMAN, NEWMAN_UNIT

    for name, directed, edges, wf, expected in NEWMAN:
        g = _weighted_graph(rx, directed, edges)
        for threshold in (200, 1):
            got = rx.newman_weighted_closeness_centrality(g, weight_fn=float, wf_improved=wf,
                                                          parallel_threshold=threshold)
            assert list(got.keys()) == list(range(len(expected))), name
            for a, b in zip(got.values(), expected):
                assert abs(a - b) < 1e-4, (name, list(got.values()))
    for name, directed, edges in NEWMAN_UNIT:
        g = _weighted_graph(rx, directed, edges)
        for wf in (False, True):
            assert rx.newman_weighted_closeness_centrality(g, float, wf) == rx.closeness_centrality(g, wf), name
    # tests/graph/test_centrality.py:122-130, tests/digraph/test_centrality.py:151-158
    for directed in (False, True):
        g = rx.PyDiGraph() if directed else rx.PyGraph()
        a, b, c = g.add_node("A"), g.add_node("B"), g.add_node("C")
        c0 = g.add_node("C0")
        d = g.add_node("D")
        g.add_edges_from([(a, b, 1), (b, c, 1), (c, d, 1)])
        g.remove_node(c0)
        assert rx.closeness_centrality(g, parallel_threshold=1) == rx.newman_weighted_closeness_centrality(
            g, default_weight=1.0, parallel_threshold=1)


def test_reference_goldens_all_pairs_dijkstra(rx):
    from weighted_cases import ALL_PAIRS, DIJKSTRA_EDGES

    for name, directed, remove, expected in ALL_PAIRS:
        g = _weighted_graph(rx, directed, DIJKSTRA_EDGES, remove)
        typed = rx.digraph_all_pairs_dijkstra_path_lengths if directed else rx.graph_all_pairs_dijkstra_path_lengths
        got = typed(g, float)
        assert got == expected, (name, got)
        assert rx.all_pairs_dijkstra_path_lengths(g, float) == expected
        assert list(got.keys()) == sorted(expected)
    # tests/graph/test_dijkstra.py:189-205, :262-270; tests/test_dispatch.py:88-90
    assert rx.graph_all_pairs_dijkstra_path_lengths(rx.PyGraph(), float) == {}
    g = rx.PyGraph()
    g.add_nodes_from(list(range(1000)))
    assert rx.graph_all_pairs_dijkstra_path_lengths(g, float) == {x: {} for x in range(1000)}
    g = rx.generators.path_graph(2)
    for bad in (float("nan"), -1):
        with pytest.raises(ValueError):
            rx.graph_all_pairs_dijkstra_path_lengths(g, lambda _: bad)
    assert rx.all_pairs_dijkstra_path_lengths(g, lambda _: 1) == {0: {1: 1.0}, 1: {0: 1.0}}


def _random_weights(g, seed, kind):
    rng = np.random.default_rng(seed)
    _, eb = g._bounds()
    if kind == "uniform":
        w = rng.uniform(0.05, 2.0, size=max(eb, 1))
    elif kind == "ints":  # many ties between different paths
        w = rng.integers(1, 4, size=max(eb, 1)).astype(np.float64)
    else:  # wide dynamic range: rounding of the left-to-right sums matters
        w = np.exp(rng.uniform(-12, 12, size=max(eb, 1)))
    return w


@pytest.mark.parametrize("kind", ["uniform", "ints", "wide"])
@pytest.mark.parametrize("directed,n,m,seed,holes", CASES)
def test_weighted_vs_oracle(rx, oracle, directed, n, m, seed, holes, kind):
    from rustworkx_b200 import device_snapshot

    g = random_graph(rx, directed, n, m, seed, holes)
    og = oracle_graph(oracle, g)
    w = _random_weights(g, seed, kind)
    dg = device_snapshot(g)
    live = np.asarray(g.node_indices(), dtype=np.uint32)
    # all-pairs Dijkstra lengths: bit exact, same reachability
    want = og.dijkstra_lengths_rows(w, live, threads=oracle.max_threads())[:, live]
    got = dg.all_pairs_dijkstra_lengths_rows(w)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])
    # Newman closeness: distances are exact, the per-source sum is tiled -> a few ulp
    for wf in (True, False):
        cw, present = og.newman_weighted_closeness_centrality(w, wf, 1 << 30, as_arrays=True)
        cg = dg.newman_weighted_closeness(w, wf)
        assert np.allclose(cg, cw, rtol=1e-12, atol=0.0)
        assert np.array_equal(cg == 0.0, cw == 0.0)


def test_weighted_infinite_costs(rx, oracle):
    """Newman weight 0 -> cost +inf; an infinite edge cost for Dijkstra: the vertex behind it is in the
    map with distance +inf (dijkstra.rs:139-141), an unreachable one is not."""
    from rustworkx_b200 import device_snapshot

    for directed in (False, True):
        g = rx.PyDiGraph() if directed else rx.PyGraph()
        g.add_nodes_from(range(6))
        g.add_edges_from([(0, 1, 1.0), (1, 2, 0.0), (2, 3, 2.0), (4, 0, 0.5)])  # node 5 isolated
        og = oracle_graph(oracle, g)
        w = np.array([1.0, 0.0, 2.0, 0.5])
        dg = device_snapshot(g)
        for wf in (True, False):
            cw, _ = og.newman_weighted_closeness_centrality(w, wf, 1 << 30, as_arrays=True)
            assert np.array_equal(dg.newman_weighted_closeness(w, wf), cw)
        cost = np.array([1.0, np.inf, 2.0, 0.5])
        want = og.dijkstra_lengths_rows(cost, np.arange(6, dtype=np.uint32))
        got = dg.all_pairs_dijkstra_lengths_rows(cost)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])


def test_weighted_closeness_sampled_sources_ba(rx, oracle):
    """BA(100k, 8) with random weights: Newman closeness of 256 sampled sources against the oracle."""
    from rustworkx_b200 import device_snapshot

    g = rx.barabasi_albert_graph(100_000, 8, seed=4)
    og = oracle_graph(oracle, g)
    w = _random_weights(g, 9, "uniform")
    sources = np.random.default_rng(2).choice(100_000, size=256, replace=False).astype(np.uint32)
    want, _ = og.weighted_closeness_partial(w, sources, threads=oracle.max_threads())
    got = device_snapshot(g).weighted_closeness_partial(w, sources)
    assert np.allclose(got, want[sources], rtol=1e-12, atol=0.0)


# ------------------------------------------------------------------------------------ sigma row formats
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("directed,n,m,seed,holes", CASES[2:6])
def test_sigma_row_formats_vs_oracle(rx, oracle, directed, n, m, seed, holes, mode):
    """bc_sigma = 1: f64 sigma rows; 2: uint32 path counts (the default tries uint32 first)."""
    from rustworkx_b200 import _lib

    g = random_graph(rx, directed, n, m, seed, holes)
    og = oracle_graph(oracle, g)
    nb, eb = g._bounds()
    _lib.set_option("bc_sigma", mode)
    try:
        for endpoints in (False, True):
            want, _ = og.betweenness_centrality(endpoints, True, 1 << 30, as_arrays=True)
            gd, _ = dense_from_mapping(rx.betweenness_centrality(g, endpoints=endpoints), nb)
            assert_close(gd, want)
        want, _ = og.edge_betweenness_centrality(False, 1 << 30, as_arrays=True)
        gd, _ = dense_from_mapping(rx.edge_betweenness_centrality(g, normalized=False), eb)
        assert_close(gd, want)
        group = list(g.node_indices())[:3]
        want = og.group_betweenness_centrality(group, True, 1 << 30)
        assert abs(rx.group_betweenness_centrality(g, group) - want) <= ATOL + RTOL * abs(want)
    finally:
        _lib.set_option("bc_sigma", 0)


@pytest.mark.parametrize("directed", [False, True])
def test_source_order_does_not_change_results(rx, oracle, directed):
    """bc_order = 1 (default) sorts the sources of a call by two-hop neighbourhood size and spreads
    consecutive ones over a lane's sector; bc_order = 0 keeps the caller's order.  Contributions are
    additive over sources, so both must agree with the oracle -- including a last group that is not
    full, a source list in descending order and the endpoints / edge variants."""
    from rustworkx_b200 import _lib, device_snapshot

    g = random_graph(rx, directed, 2500, 15000, 77, holes=25)
    og = oracle_graph(oracle, g)
    nb, eb = g._bounds()
    nodes = np.asarray(g.node_indices(), dtype=np.uint32)
    src = nodes[::-1][: 3 * 128 + 37].copy()
    res = {}
    for flag in (1, 0):
        _lib.set_option("bc_order", flag)
        try:
            dg = device_snapshot(g)
            res[flag] = (dg.betweenness_partial(src, False), dg.betweenness_partial(src, True),
                         dg.edge_betweenness_partial(src), dg.stats()["te"], dg.stats()["v"])
        finally:
            _lib.set_option("bc_order", 2)
    for a, b in zip(res[1][:3], res[0][:3]):
        assert_close(a, b)
    assert res[1][3:] == res[0][3:]
    assert_close(res[1][0], og.betweenness_partial(src, False)[0])
    assert_close(res[1][1], og.betweenness_partial(src, True)[0])
    assert_close(res[1][2], og.edge_betweenness_partial(src)[0])
    # all sources, through the public API, against the oracle
    for endpoints in (False, True):
        want, _ = og.betweenness_centrality(endpoints, True, 1 << 30, as_arrays=True)
        gd, _ = dense_from_mapping(rx.betweenness_centrality(g, endpoints=endpoints), nb)
        assert_close(gd, want)


def test_sigma_overflow_falls_back_to_f64(rx, oracle):
    """A 40x40 grid has C(78,39) ~ 2.7e22 shortest paths corner to corner: the uint32 path counts
    overflow, the batch is redone with f64 sigma rows, and the handle stays wide."""
    from rustworkx_b200 import _lib, device_snapshot

    g = rx.generators.grid_graph(40, 40)
    og = oracle_graph(oracle, g)
    want, _ = og.betweenness_centrality(False, True, 1 << 30, as_arrays=True)
    got = np.array(list(rx.betweenness_centrality(g).values()))
    assert_close(got, want)
    dg = device_snapshot(g)
    st = dg.stats()
    assert st["sources"] == 1600  # the redone batch is not counted twice
    want_e, _ = og.edge_betweenness_centrality(True, 1 << 30, as_arrays=True)
    assert_close(np.array(list(rx.edge_betweenness_centrality(g).values())), want_e)
    group = [0, 820, 1599]
    want_g = og.group_betweenness_centrality(group, True, 1 << 30)
    assert abs(rx.group_betweenness_centrality(g, group) - want_g) <= RTOL * abs(want_g)
    _lib.set_option("bc_sigma", 2)
    try:
        with pytest.raises(rx.RxCudaError):
            device_snapshot(rx.generators.grid_graph(40, 41)).betweenness(False, True)
    finally:
        _lib.set_option("bc_sigma", 0)


def test_level_queue_overflow_falls_back_to_worst_case_queues(rx, oracle):
    """On a path graph every vertex sits on a different level for each of the 128 sources of a group:
    the typical-case level queues (16 entries per vertex) overflow, the batch is redone with
    worst-case queues, and the results still match the oracle."""
    from rustworkx_b200 import device_snapshot

    g = rx.generators.path_graph(600)
    og = oracle_graph(oracle, g)
    want, _ = og.betweenness_centrality(False, False, 1 << 30, as_arrays=True)
    got = np.array(list(rx.betweenness_centrality(g, normalized=False).values()))
    assert_close(got, want)
    assert device_snapshot(g).stats()["sources"] == 600
    want_e, _ = og.edge_betweenness_centrality(False, 1 << 30, as_arrays=True)
    assert_close(np.array(list(rx.edge_betweenness_centrality(g, normalized=False).values())), want_e)
    h = rx.generators.directed_path_graph(300)
    oh = oracle_graph(oracle, h)
    group = [10, 150]
    want_g = oh.group_betweenness_centrality(group, False, 1 << 30)
    assert abs(rx.group_betweenness_centrality(h, group, normalized=False) - want_g) <= RTOL * abs(want_g)
