"""Differential fuzzing of the CUDA path against the oracle on small arbitrary graphs -- the shape of
the reference's own fuzz target (rustworkx-core/fuzz/fuzz_targets/test_fuzz_centrality.rs:32-60:
arbitrary node count and edge list, out-of-range endpoints dropped), extended to every function of
the path, directed and undirected, with holes.  Deterministic (derandomize) so a failure reproduces."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rx():
    import rustworkx_b200 as rx
    from rustworkx_b200 import _lib

    if _lib.device_count() == 0:
        pytest.fail("no CUDA device: the gpu-marked tests need a B200 (there is no CPU fallback)")
    return rx


graphs = st.tuples(
    st.booleans(),                                     # directed
    st.integers(min_value=1, max_value=70),            # node_count
    st.lists(st.tuples(st.integers(0, 80), st.integers(0, 80)), max_size=160),
    st.lists(st.integers(0, 69), max_size=4),          # nodes to remove
    st.integers(0, 2**31 - 1),                         # seed for groups / weights
)


@settings(max_examples=150, deadline=None, derandomize=True,
          suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(graphs)
def test_fuzz_all_functions_vs_oracle(rx, oracle, case):
    directed, n, edges, removed, seed = case
    g = rx.PyDiGraph() if directed else rx.PyGraph()
    g.add_nodes_from(range(n))
    g.add_edges_from_no_data([(u, v) for u, v in edges if u < n and v < n])
    for r in set(removed):
        if r < n and g.num_nodes() > 1:
            g.remove_node(r)
    og = oracle.OracleGraph.from_graph(g)
    nb, eb = g._bounds()
    live = np.asarray(g.node_indices(), dtype=np.int64)

    def dense(mapping, bound):
        out = np.zeros(bound)
        for k, v in mapping.items():
            out[k] = v
        return out

    for normalized in (True, False):
        want, _ = og.betweenness_centrality(True, normalized, 1 << 30, as_arrays=True)
        got = dense(rx.betweenness_centrality(g, normalized=normalized, endpoints=True), nb)
        assert np.allclose(got, want, rtol=1e-9, atol=1e-12)
        want, _ = og.edge_betweenness_centrality(normalized, 1 << 30, as_arrays=True)
        got = dense(rx.edge_betweenness_centrality(g, normalized=normalized), eb)
        assert np.allclose(got, want, rtol=1e-9, atol=1e-12)
    for wf in (True, False):
        want, _ = og.closeness_centrality(wf, 1 << 30, as_arrays=True)
        assert np.array_equal(dense(rx.closeness_centrality(g, wf_improved=wf), nb), want)
    nv = -1.5
    want = og.distance_matrix(1 << 30, not directed, nv)
    got = rx.distance_matrix(g, null_value=nv)
    assert np.array_equal(got, want)
    a, b = rx.unweighted_average_shortest_path_length(g), og.average_shortest_path_length(300, not directed, False)
    assert (a == b) or (a != a and b != b)

    rng = np.random.default_rng(seed)
    group = rng.choice(live, size=min(len(live), int(rng.integers(1, 5))), replace=False).tolist()
    assert rx.group_closeness_centrality(g, group) == og.group_closeness_centrality(group)
    want = og.group_betweenness_centrality(group, True, 1 << 30)
    assert abs(rx.group_betweenness_centrality(g, group) - want) <= 1e-12 + 1e-9 * abs(want)

    w = rng.uniform(0.1, 3.0, size=max(eb, 1))
    from rustworkx_b200 import device_snapshot

    dg = device_snapshot(g)
    want, _ = og.newman_weighted_closeness_centrality(w, True, 1 << 30, as_arrays=True)
    assert np.allclose(dg.newman_weighted_closeness(w, True), want, rtol=1e-12, atol=0.0)
    if g.num_edges():
        want = og.dijkstra_lengths_rows(w, live.astype(np.uint32))[:, live]
        got = dg.all_pairs_dijkstra_lengths_rows(w)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])
