"""CPU-side tests: the C-ABI library loads and exports every declared symbol, the host logic (graph
stand-in, mappings, generators, sharding + reduce over gloo) behaves like the reference, and the
product path fails loudly without a GPU.  No compute calls on a device here.
"""

import ctypes
import os
import pickle
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rx():
    import __graft_entry__ as ge

    if not os.path.exists(os.path.join(ROOT, "rustworkx_b200", "librxcuda.so")):
        ge.build()
    import rustworkx_b200 as rx

    return rx


# --------------------------------------------------------------------------------- C-ABI surface
def test_stats_struct_mirror_matches_the_header(rx):
    """rxc_stats is filled by the library and read through ctypes: same fields, same order, 8 bytes each."""
    from rustworkx_b200 import _lib

    header = open(os.path.join(ROOT, "include", "rxcuda.h")).read()
    body = header[header.index("typedef struct {"):header.index("} rxc_stats;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    declared = re.findall(r"\b(uint64_t|double)\s+(\w+)\s*;", body)
    assert [name for _, name in declared] == [name for name, _ in _lib.Stats._fields_]
    for (ctype, name), (_, field) in zip(declared, _lib.Stats._fields_):
        assert field is (ctypes.c_uint64 if ctype == "uint64_t" else ctypes.c_double), name
    assert ctypes.sizeof(_lib.Stats) == 8 * len(declared)


def test_header_symbols_are_exported(rx):
    from rustworkx_b200 import _lib

    header = open(os.path.join(ROOT, "include", "rxcuda.h")).read()
    declared = set(re.findall(r"\b(rxc_[a-z0-9_]+)\s*\(", header))
    declared -= {"rxc_status", "rxc_graph", "rxc_stats"}
    assert len(declared) >= 25
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} is declared in include/rxcuda.h but not exported"
    assert declared == set(_lib.SYMBOLS), "ctypes table and header disagree"
    assert b"sm_100a" in L.rxc_version()


def test_no_cpu_fallback(rx):
    from rustworkx_b200 import _lib

    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    g = rx.generators.path_graph(4)
    for call in (lambda: rx.betweenness_centrality(g), lambda: rx.closeness_centrality(g),
                 lambda: rx.edge_betweenness_centrality(g), lambda: rx.distance_matrix(g)):
        with pytest.raises(rx.RxCudaError):
            call()


def test_pinned_allocation_needs_a_device(rx):
    """rxc_host_alloc page-locks memory through the CUDA runtime: without a device it fails loudly
    instead of quietly handing out pageable memory."""
    from rustworkx_b200 import _lib

    if _lib.device_count() > 0:
        a = _lib.pinned_copy(np.arange(10, dtype=np.uint32))
        assert a.dtype == np.uint32 and a.tolist() == list(range(10))
        b = _lib.pinned_empty((3, 4), np.float64)
        assert b.shape == (3, 4) and b.flags.c_contiguous
        return
    with pytest.raises(rx.RxCudaError):
        _lib.pinned_empty(16, np.float64)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "rustworkx_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "rx_oracle" not in text and "import oracle" not in text, f


def test_rescale_matches_reference_rule(rx):
    # rustworkx-core/src/centrality.rs:221-252
    from rustworkx_b200 import _lib

    x = np.array([1.0, 2.0, 3.0])
    assert np.array_equal(_lib.rescale(x.copy(), 5, True, False, False), x * (1.0 / 12.0))
    assert np.array_equal(_lib.rescale(x.copy(), 5, True, True, True), x * (1.0 / 20.0))
    assert np.array_equal(_lib.rescale(x.copy(), 2, True, False, False), x)
    assert np.array_equal(_lib.rescale(x.copy(), 1, True, False, True), x)
    assert np.array_equal(_lib.rescale(x.copy(), 5, False, False, False), x * 0.5)
    assert np.array_equal(_lib.rescale(x.copy(), 5, False, True, False), x)


# --------------------------------------------------------------------------------- containers
def test_stable_graph_index_semantics(rx):
    g = rx.PyGraph()
    a, b, c, d = (g.add_node(x) for x in "abcd")
    e0 = g.add_edge(a, b, "x")
    e1 = g.add_edge(b, c, "y")
    e2 = g.add_edge(c, d, "z")
    g.remove_node(b)  # removes e0, e1
    assert list(g.node_indices()) == [0, 2, 3] and list(g.edge_indices()) == [e2]
    assert g._bounds() == (4, 3)
    g.remove_node(99)  # silently ignored, src/graph.rs:883-888
    assert g.add_node("new") == b  # LIFO reuse of the vacated slot
    # petgraph frees b's outgoing list then its incoming list; the last freed edge is reused first
    assert g.add_edge(a, c, None) in (e0, e1)
    assert g.num_nodes() == 4 and g.num_edges() == 2
    g.remove_node(d)
    assert g._bounds()[0] == 3  # node_bound = highest live index + 1


def test_multigraph_false_updates_in_place(rx):
    g = rx.PyGraph(multigraph=False)
    g.add_nodes_from(range(3))
    e = g.add_edge(0, 1, "a")
    assert g.add_edge(1, 0, "b") == e and g.num_edges() == 1 and g.get_edge_data(0, 1) == "b"
    dg = rx.PyDiGraph(multigraph=False)
    dg.add_nodes_from(range(3))
    e = dg.add_edge(0, 1, "a")
    assert dg.add_edge(1, 0, "b") != e and dg.num_edges() == 2
    mg = rx.PyGraph()
    mg.add_nodes_from(range(2))
    mg.add_edge(0, 1, None)
    mg.add_edge(0, 1, None)
    assert mg.num_edges() == 2  # parallel edges are kept: path multiplicity matters


def test_snapshot_arrays(rx):
    g = rx.generators.directed_path_graph(5)
    g.remove_node(2)
    s = g._snapshot_arrays()
    assert s["n_bound"] == 5 and s["directed"] and list(s["live_nodes"]) == [0, 1, 3, 4]
    assert list(zip(s["edge_src"], s["edge_dst"])) == [(0, 1), (3, 4)] and list(s["edge_id"]) == [0, 3]
    assert not s["edge_id_is_position"]


def test_snapshot_arrays_without_holes_are_views(rx):
    """No slot ever vacated: no scans, no copies -- read-only views of the container's storage and
    "the edge index is the position in the list" (the C-ABI then gets edge_id = NULL)."""
    g = rx.generators.path_graph(6)
    s = g._snapshot_arrays()
    assert s["edge_id_is_position"] and s["n_bound"] == 6 and s["e_bound"] == 5
    assert list(s["live_nodes"]) == list(range(6)) and list(s["edge_id"]) == list(range(5))
    assert np.shares_memory(s["edge_src"], g._esrc) and not s["edge_src"].flags.writeable
    assert g._esrc.flags.writeable  # the container itself stays mutable
    assert list(zip(s["edge_src"].tolist(), s["edge_dst"].tolist())) == [(i, i + 1) for i in range(5)]
    assert g._bounds() == (6, 5)
    # vacate a slot: explicit ids; re-use it: positional again, with the new endpoints in place
    g.remove_edge_from_index(1)
    s = g._snapshot_arrays()
    assert not s["edge_id_is_position"] and list(s["edge_id"]) == [0, 2, 3, 4] and s["e_bound"] == 5
    assert not np.shares_memory(s["edge_src"], g._esrc)
    assert g.add_edge(0, 5, None) == 1
    s = g._snapshot_arrays()
    assert s["edge_id_is_position"] and (int(s["edge_src"][1]), int(s["edge_dst"][1])) == (0, 5)
    # the last node removed: the bound shrinks to the highest live index + 1
    g.remove_node(5)
    assert g._bounds()[0] == 5 and list(g._snapshot_arrays()["live_nodes"]) == [0, 1, 2, 3, 4]


def test_centrality_mapping_behaviour(rx):
    # src/iterators.rs:1199-1316
    m = rx.CentralityMapping._from_arrays([0, 2, 5], [0.5, 1.0, 0.25])
    assert m == {0: 0.5, 2: 1.0, 5: 0.25} and m != {0: 0.5}
    assert list(m.keys()) == [0, 2, 5] and list(m.values()) == [0.5, 1.0, 0.25]
    assert list(m.items()) == [(0, 0.5), (2, 1.0), (5, 0.25)] and list(m) == [0, 2, 5]
    assert len(m) == 3 and 2 in m and 3 not in m and m[5] == 0.25
    with pytest.raises(IndexError, match="No node found for index"):
        m[1]
    with pytest.raises(NotImplementedError):
        m < {}
    assert str(m) == "CentralityMapping{0: 0.5, 2: 1.0, 5: 0.25}"
    assert pickle.loads(pickle.dumps(m)) == m and hash(m) == hash(pickle.loads(pickle.dumps(m)))
    # look-ups before the dict exists are binary searches over the ascending keys
    big = rx.CentralityMapping._from_arrays(np.arange(0, 200000, 2), np.arange(100000, dtype=np.float64))
    assert big[19998] == 9999.0 and 19999 not in big and big._dict is None
    with pytest.raises(IndexError, match="No node found for index"):
        big[200001]
    with pytest.raises(OverflowError):
        big[-1]
    with pytest.raises(TypeError):
        big[1.0]
    e = rx.EdgeCentralityMapping._from_arrays([1], [2.0])
    assert str(e) == "EdgeCentralityMapping{1: 2.0}" and rx.CentralityMapping() == {}


def test_argument_errors_before_any_gpu_work(rx):
    g = rx.PyGraph()
    with pytest.raises(OverflowError):
        rx.betweenness_centrality(g, parallel_threshold=-1)
    with pytest.raises(TypeError):
        rx.betweenness_centrality(g, normalized=1)
    with pytest.raises(TypeError):
        rx.closeness_centrality(g, wf_improved="yes")
    with pytest.raises(TypeError):
        rx.betweenness_centrality(object())
    with pytest.raises(TypeError):
        rx.distance_matrix(g, as_undirected=True)  # tests/test_dispatch.py:33-36
    with pytest.raises(TypeError):
        rx.digraph_closeness_centrality(g)
    with pytest.raises(TypeError):
        rx.graph_betweenness_centrality(graph=g)  # positional-only, as in the pyo3 signature


# --------------------------------------------------------------------------------- generators
def test_random_generator_counts(rx):
    # the reference pins only counts: tests/test_random.py:104-107,133-140,440-443
    g = rx.directed_gnm_random_graph(20, 100, seed=10)
    assert g.num_nodes() == 20 and g.num_edges() == 100
    el = list(g.edge_list())
    assert len(set(el)) == 100 and all(a != b for a, b in el)
    g = rx.undirected_gnm_random_graph(20, 400, seed=1)
    assert g.num_edges() == 190  # saturates at the complete graph
    g = rx.undirected_gnp_random_graph(20, 1.0, seed=1)
    assert g.num_edges() == 190
    assert rx.directed_gnp_random_graph(20, 1.0).num_edges() == 380
    assert rx.undirected_gnp_random_graph(20, 0.0).num_edges() == 0
    g = rx.undirected_gnp_random_graph(5000, 0.002, seed=42)
    assert g.num_nodes() == 5000 and abs(g.num_edges() - 24995) < 1000
    assert list(g.edge_list()) == list(rx.undirected_gnp_random_graph(5000, 0.002, seed=42).edge_list())
    g = rx.barabasi_albert_graph(20, 12, seed=42)
    assert g.num_nodes() == 20 and g.num_edges() == 107  # random_graph.rs doctest :752-753
    g = rx.barabasi_albert_graph(10**5, 8, seed=1)
    assert g.num_edges() == 7 + 8 * (10**5 - 8)
    deg = np.bincount(np.concatenate([g._esrc[: g._elen], g._edst[: g._elen]]))
    assert deg.min() >= 1 and deg.max() > 300  # preferential attachment grows hubs


def test_oracle_vs_networkx_fixture(oracle):
    # tests/golden/nx_bc_gnp300.json was generated with networkx 3.6.1 (tests/golden/make_golden.py)
    import json

    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "nx_golden.json")))
    for case in fx["cases"]:
        g = oracle.OracleGraph.from_edges(case["directed"], [tuple(e) for e in case["edges"]],
                                          num_nodes=case["n"])
        bc = g.betweenness_centrality(case["endpoints"], case["normalized"], 1 << 30)
        assert np.allclose(bc, case["betweenness"], rtol=1e-12, atol=1e-15)
        ebc = g.edge_betweenness_centrality(case["normalized"], 1 << 30)
        assert np.allclose(ebc, case["edge_betweenness"], rtol=1e-12, atol=1e-15)
        cc = g.closeness_centrality(True, 1 << 30)
        assert np.allclose(cc, case["closeness_wf"], rtol=4e-16, atol=0)
        dm = g.distance_matrix(300, False, -1.0)
        assert np.array_equal(dm, np.array(case["distance_matrix"]))


def test_oracle_next_rows_vs_networkx_fixture(oracle):
    """The SURVEY 8f restatements against independent networkx vectors
    (tests/golden/make_golden_next_rows.py): group centralities, Newman weighted closeness,
    all-pairs Dijkstra lengths and the sums behind unweighted_average_shortest_path_length."""
    import json

    base = json.load(open(os.path.join(ROOT, "tests", "golden", "nx_golden.json")))["cases"]
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "nx_golden_next_rows.json")))
    assert len(fx["cases"]) == len(base)
    checked_gbc = 0
    for c, b in zip(fx["cases"], base):
        n = b["n"]
        g = oracle.OracleGraph.from_edges(b["directed"], [tuple(e) for e in b["edges"]], num_nodes=n)
        for normalized in (True, False):
            want = c["group_betweenness"][str(normalized)]  # enumerated shortest paths
            got = g.group_betweenness_centrality(c["group"], normalized, 1 << 30)
            assert abs(got - want) <= 1e-12 * max(1.0, abs(want))
            checked_gbc += 1
        assert abs(g.group_closeness_centrality(c["group"]) - c["group_closeness"]) <= 1e-15
        w = np.array(c["weights"])
        got, _ = g.newman_weighted_closeness_centrality(w, True, 1 << 30, as_arrays=True)
        assert np.allclose(got, c["newman_closeness_wf"], rtol=1e-13, atol=0)
        rows = g.dijkstra_lengths_rows(w, np.arange(n, dtype=np.uint32))
        for s in range(n):
            for t in range(n):
                want = c["dijkstra_lengths"][s][t]
                if want is None:
                    assert s == t or np.isnan(rows[s, t])
                else:
                    assert rows[s, t] == want  # same left-to-right sums: bit-identical
        assert g.distance_sum(300, False) == (c["distance_total"], c["connected_pairs"])
    assert checked_gbc == 2 * len(base)


# --------------------------------------------------------------------------------- multi-process
GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle"))
import numpy as np, torch.distributed as dist
import rustworkx_b200 as rx
from rustworkx_b200 import distributed as rxd
import oracle
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
g = rx.barabasi_albert_graph(400, 3, seed=5)
g.remove_node(7); g.remove_node(100)
og = oracle.OracleGraph.from_graph(g)
node = rxd.betweenness_centrality(g, True, True,
        compute_partial=lambda s, ep: og.betweenness_partial(s, ep)[0])
edge = rxd.edge_betweenness_centrality(g, False,
        compute_partial=lambda s, ep: og.edge_betweenness_partial(s)[0])
group = [3, 50, 50, 211]
gb = rxd.group_betweenness_centrality(g, group, True,
        compute_partial=lambda grp, s: og.group_betweenness_partial(grp, s)[0])
_, eb = g._bounds()
w = np.random.default_rng(3).uniform(0.1, 2.0, size=eb)
nw = rxd.newman_weighted_closeness_centrality(g, w, True,
        compute_partial=lambda ww, s: og.weighted_closeness_partial(ww, s)[0][s])
cc = rxd.closeness_centrality(g, True, compute_partial=lambda s, wf: og.closeness_partial(s, wf)[0][s])
full_dm, _ = og.distance_matrix(1 << 30, False, -1.0, return_stats=True)
dm = rxd.distance_matrix(g, False, -1.0, compute_rows=lambda lo, hi, und, nv: full_dm[lo:hi])
if dist.get_rank() == 0:
    want_cc, _ = og.closeness_centrality(True, 1 << 30, as_arrays=True)
    assert np.array_equal(cc, want_cc)  # per-source values: the sharded result is bit-identical
    assert np.array_equal(dm, full_dm) and dm.shape == (g.num_nodes(), g.num_nodes())
    assert abs(gb - og.group_betweenness_centrality(group, True, 1 << 30)) <= 1e-12 * abs(gb)
    want_nw, _ = og.newman_weighted_closeness_centrality(w, True, 1 << 30, as_arrays=True)
    assert np.array_equal(nw, want_nw)
    want, _ = og.betweenness_centrality(True, True, 1 << 30, as_arrays=True)
    want_e, _ = og.edge_betweenness_centrality(False, 1 << 30, as_arrays=True)
    assert np.allclose(node, want, rtol=1e-12, atol=1e-15), np.abs(node - want).max()
    assert np.allclose(edge, want_e, rtol=1e-12, atol=1e-15)
    print("GLOO_OK")
else:
    assert node is None and edge is None and gb is None and nw is None and cc is None and dm is None
dist.destroy_process_group()
"""


def test_sharded_betweenness_over_gloo(rx, oracle, tmp_path):
    """world_size-2 run of the N>1 host path (shard -> partial -> one reduce -> rescale on rank 0);
    the oracle stands in for the GPU partial so this runs on CPU."""
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "GLOO_OK" in outs[0]


def test_shard_sources_partition(rx):
    from rustworkx_b200.distributed import shard_sources

    src = np.arange(1003, dtype=np.uint32)
    for world in (1, 2, 4, 8):
        parts = [shard_sources(src, r, world) for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == src.tolist()
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_sources(src, 2, 2)


def test_group_argument_validation_happens_before_any_device_work():
    # src/centrality.rs:1216-1222: the binding checks the group first -> ValueError even without a GPU
    import rustworkx_b200 as rx

    g = rx.generators.path_graph(3)
    with pytest.raises(ValueError, match="Node index 10 is not in the graph"):
        rx.graph_group_closeness_centrality(g, [10])
    with pytest.raises(ValueError):
        rx.group_betweenness_centrality(g, [0, 7])
    with pytest.raises(TypeError):
        rx.group_betweenness_centrality(g, "01")
    with pytest.raises(TypeError):
        rx.digraph_group_closeness_centrality(g, [0])


def test_all_pairs_path_length_mapping_behaviour():
    """AllPairsPathLengthMapping / PathLengthMapping (src/iterators.rs, custom_hash_map_iter_impl!):
    every live node is a key, the inner mapping holds the reached nodes except the node itself."""
    import pickle as pk

    from rustworkx_b200.iterators import AllPairsPathLengthMapping, PathLengthMapping

    nan = float("nan")
    rows = np.array([[0.0, 1.5, nan], [nan, 0.0, nan], [2.0, 3.5, 0.0]])
    m = AllPairsPathLengthMapping([0, 2, 5], rows)
    assert m == {0: {2: 1.5}, 2: {}, 5: {0: 2.0, 2: 3.5}}
    assert list(m.keys()) == [0, 2, 5] and len(m) == 3 and 5 in m and 1 not in m
    assert isinstance(m[5], PathLengthMapping) and m[5][2] == 3.5
    with pytest.raises(IndexError):
        m[1]
    with pytest.raises(IndexError):
        m[0][5]
    assert m != {0: {2: 1.5}, 2: {}, 5: {0: 2.0}}
    assert str(m).startswith("AllPairsPathLengthMapping{0: PathLengthMapping{2: 1.5}")
    assert pk.loads(pk.dumps(m)) == m
    assert [k for k, _ in m.items()] == [0, 2, 5]
    empty = AllPairsPathLengthMapping([1, 3], None)  # a graph without edges
    assert empty == {1: {}, 3: {}}
    assert AllPairsPathLengthMapping() == {}


def test_weight_validation_matches_is_valid_weight():
    # src/lib.rs:303-313 (negative incl. -0.0 and NaN rejected, +inf and 0 accepted); extraction errors
    import rustworkx_b200 as rx

    g = rx.PyGraph()
    g.add_nodes_from(range(3))
    g.add_edges_from([(0, 1, 2.0), (1, 2, "x")])
    for bad in (-1.0, -0.0, float("nan")):
        with pytest.raises(ValueError):
            rx.all_pairs_dijkstra_path_lengths(g, lambda _w, b=bad: b)
    with pytest.raises(TypeError):
        rx.all_pairs_dijkstra_path_lengths(g, lambda w: w)  # "x" is not a float
    with pytest.raises(TypeError):
        rx.all_pairs_dijkstra_path_lengths(g, 3)  # not callable
    with pytest.raises(ValueError):
        rx.newman_weighted_closeness_centrality(g, weight_fn=lambda _w: -2.0)
    with pytest.raises(TypeError):
        rx.newman_weighted_closeness_centrality(rx.PyGraph(), wf_improved=1)
    with pytest.raises(TypeError):
        rx.graph_newman_weighted_closeness_centrality(rx.PyDiGraph())
    # graphs without nodes / edges never reach the device
    assert rx.all_pairs_dijkstra_path_lengths(rx.PyDiGraph(), float) == {}
    h = rx.PyGraph()
    h.add_nodes_from(range(4))
    assert rx.all_pairs_dijkstra_path_lengths(h, float) == {i: {} for i in range(4)}


def test_stand_in_container_and_oracle_container_agree_under_mutation(rx, oracle):
    """Two independent restatements of petgraph's StableGraph index discipline -- the numpy-backed
    stand-in (rustworkx_b200/graph.py) that feeds the device and the intrusive-list one inside the C
    oracle -- driven by the same random add / remove sequence: every returned index, the counts and
    the bounds must agree (vacated slots are re-used LIFO; removing a node removes its edges)."""
    for directed in (False, True):
        rng = np.random.default_rng(99 + int(directed))
        g = rx.PyDiGraph() if directed else rx.PyGraph()
        og = oracle.OracleGraph(directed)
        live_nodes, live_edges = [], []
        for step in range(600):
            op = rng.random()
            if op < 0.30 or len(live_nodes) < 2:
                a, b = g.add_node(None), og.add_node()
                assert a == b
                live_nodes.append(a)
            elif op < 0.75:
                u, v = (int(x) for x in rng.choice(live_nodes, size=2))
                a, b = g.add_edge(u, v, None), og.add_edge(u, v)
                assert a == b
                live_edges.append(a)
            elif op < 0.88 and live_edges:
                e = live_edges.pop(int(rng.integers(0, len(live_edges))))
                g.remove_edge_from_index(e)
                og.remove_edge(e)
            elif live_nodes:
                x = live_nodes.pop(int(rng.integers(0, len(live_nodes))))
                g.remove_node(x)
                og.remove_node(x)
                live_edges = [int(e) for e in g.edge_indices()]
            assert (g.num_nodes(), g.num_edges()) == (og.node_count(), og.edge_count())
            assert g._bounds() == (og.node_bound(), og.edge_bound())
        s = g._snapshot_arrays()
        assert [int(x) for x in s["live_nodes"]] == sorted(live_nodes)
        for e, a, b in zip(s["edge_id"].tolist(), s["edge_src"].tolist(), s["edge_dst"].tolist()):
            pa, pb = ctypes.c_uint32(), ctypes.c_uint32()
            assert og._L.rxo_edge_endpoints(og._h, e, ctypes.byref(pa), ctypes.byref(pb)) == 0
            assert (pa.value, pb.value) == (a, b)


# --------------------------------------------------------------------------------- snapshot cache
class _FakeDeviceGraph:
    """Counts snapshots; stands in for DeviceGraph where no GPU exists (host logic only)."""

    made = 0

    @classmethod
    def from_graph(cls, graph, cached=False):
        cls.made += 1
        dg = cls()
        dg.nodes = graph.num_nodes()
        dg.edges = graph.num_edges()
        return dg


def test_snapshot_cache_never_resurrects_a_cleared_graph(rx, monkeypatch):
    """clear() used to reset the mutation count to 0 while the old snapshot stayed
    attached, so a rebuilt graph that reached the old count was answered from the old CSR."""
    monkeypatch.setattr(rx, "DeviceGraph", _FakeDeviceGraph)
    for cls in (rx.PyGraph, rx.PyDiGraph):
        g = cls()
        g.add_nodes_from(range(3))
        g.add_edges_from_no_data([(0, 1), (1, 2)])
        first = rx.device_snapshot(g)
        assert rx.device_snapshot(g) is first  # unchanged graph: cache hit
        stamp = g._version
        g.clear()
        assert g._version != stamp and g.num_nodes() == 0
        g.add_nodes_from(range(5))  # the same NUMBER of mutations as before
        second = rx.device_snapshot(g)
        assert second is not first and second.nodes == 5
        g.add_edge(0, 4, None)
        assert rx.device_snapshot(g) is not second  # any mutation invalidates
    # stamps are never handed out twice, not even across graphs
    a, b = rx.PyGraph(), rx.PyGraph()
    assert a._version != b._version


def test_graphs_pickle_and_deepcopy_after_a_snapshot(rx, monkeypatch):
    """the cached device handle used to live in the graph's __dict__, which made
    pickle / deepcopy fail after any centrality call.  The reference containers support both."""
    import copy
    import pickle

    monkeypatch.setattr(rx, "DeviceGraph", _FakeDeviceGraph)
    g = rx.PyDiGraph()
    g.add_nodes_from(["a", "b", "c"])
    g.add_edges_from([(0, 1, 1.5), (1, 2, 2.5)])
    dg = rx.device_snapshot(g)
    assert not any("rxc" in k for k in g.__dict__)
    for clone in (pickle.loads(pickle.dumps(g)), copy.deepcopy(g), g.copy()):
        assert clone.num_nodes() == 3 and list(clone.edge_list()) == [(0, 1), (1, 2)]
        assert rx.device_snapshot(clone) is not dg  # a clone never shares the handle
    assert rx.device_snapshot(g) is dg
    rx.set_snapshot_cache(False)
    try:
        assert rx.device_snapshot(g) is not dg
    finally:
        rx.set_snapshot_cache(True)


def test_generator_library_is_separate(rx):
    """The benchmark input generators live in librxgen.so (include/rxgen.h), not in the product
    library: bench.py's reference arm builds its graph without mapping librxcuda.so."""
    from rustworkx_b200 import _lib

    header = open(os.path.join(ROOT, "include", "rxgen.h")).read()
    declared = set(re.findall(r"\b(rxc_gen_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.GEN_SYMBOLS) and len(declared) == 3
    G = _lib.gen_lib()
    L = _lib.lib()
    for name in declared:
        assert hasattr(G, name)
        assert not hasattr(L, name), f"{name} must not be exported by librxcuda.so"
    g = rx.barabasi_albert_graph(500, 3, seed=5)
    assert g.num_nodes() == 500 and g.num_edges() == 2 + 3 * 497


def test_group_shard_partitions_an_ordered_list(rx):
    """Multi-process sharding of the ORDERED source list: whole chunks of 256 dealt round-robin -- a
    partition of the list, clusters (consecutive runs) kept together, ranks balanced."""
    from rustworkx_b200.distributed import SHARD_CHUNK, group_shard

    order = np.random.default_rng(0).permutation(70000).astype(np.uint32)
    for world in (1, 2, 4, 8):
        parts = [group_shard(order, r, world) for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == sorted(order.tolist())
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= SHARD_CHUNK
        for r, p in enumerate(parts):  # every chunk is a run of the ordered list
            first = p[:SHARD_CHUNK]
            start = r * SHARD_CHUNK
            assert np.array_equal(first, order[start:start + SHARD_CHUNK])
    with pytest.raises(ValueError):
        group_shard(order, 3, 3)
    assert group_shard(order[:100], 1, 2).shape == (0,)


# ------------------------------------------------------------------------------ bench.py contract
def test_bench_reference_arm_prints_one_contract_line(rx, oracle):
    """`bench.py --impl reference` on a small instance of the workload: ONE JSON line on stdout with the
    driver's keys, the CPU arm's own cpu_baseline / e2e blocks, and no GPU launches."""
    import json

    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "1", "--nodes", "3000", "--m", "4", "--cpu-sample", "16"],
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, proc.stdout
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "all_sources_betweenness_gteps"
    assert line["unit"] == "GTEPS" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and line["cpu_baseline"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "GTEPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["gpu_launches"] == 0


def test_bench_roofline_inputs_parse(rx):
    """The committed evidence bench.py reads for `roofline.traffic` / `line_requests` is well formed."""
    sys.path.insert(0, ROOT)
    import bench

    traffic, src = bench.ncu_traffic()
    assert {"bc_forward_kernel", "bc_backward_kernel"} <= set(traffic) and "profiles/" in src
    assert all(v["dram_bytes_per_source"] > 1e6 for v in traffic.values())
    pl = bench.probe_lines()
    assert pl and 40.0 < pl["ceiling_glines_s"] < 80.0
    for k in ("bc_forward_kernel", "bc_backward_kernel"):
        ps = pl["per_source"][k]
        assert ps["lines128"] <= ps["sectors32"] <= 4 * ps["lines128"] and ps["rows"] < ps["lines128"]


def test_bench_reference_arm_under_torchrun_prints_once(rx, oracle):
    """N > 1: the driver launches the reference arm under torchrun too; rank 0 alone runs and prints,
    the other rank exits 0 without work -- and OMP_NUM_THREADS=1 (torchrun's default) must not
    shrink the CPU arm to one thread (round 1's 20x handicap)."""
    import json
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
           "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--nodes", "3000",
           "--cpu-sample", "16"]  # (no --m: torchrun's own parser takes it for an abbreviation)
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1, proc.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2
    assert line["cpu_baseline"]["cores"] == max(1, len(os.sched_getaffinity(0)))
