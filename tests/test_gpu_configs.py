"""GPU parity at BASELINE.json's FULL configuration sizes (missing #2, #4, #5; weak #5).

The CPU oracle cannot finish these graphs for all sources, so every test compares (a) a small sample
of sources against the oracle on the same explicit edge list and (b) a size-independent invariant
whose right-hand side comes from the bit-exact distance kernels:

    sum_v  partial_bc[v]   = sum_s sum_{t reachable, t != s} (d(s,t) - 1)
    sum_e  partial_ebc[e]  = sum_s sum_{t reachable}          d(s,t)

(every shortest s-t path of length d has d-1 interior vertices and d edges, and the pair
dependencies of one (s,t) pair sum to 1 over each of them).
"""

import threading

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


def host_threads():
    import os

    return max(1, len(os.sched_getaffinity(0)))


def assert_close(got, want, rtol=RTOL):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    bad = np.flatnonzero(~(np.abs(got - want) <= ATOL + rtol * np.abs(want)))
    assert bad.size == 0, f"{bad.size} mismatches, first at {bad[:5]}: got {got[bad[:5]]} want {want[bad[:5]]}"


# ------------------------------------------------------------------------------------- config 4
def test_config4_edge_betweenness_directed_gnm_full_size(rx, oracle):
    """BASELINE.json configs[3]: edge_betweenness_centrality on directed_gnm_random_graph(2**20,
    2**24, seed=7) -- 32 sources against the oracle (centrality.rs:174-219,301-327, 459-511) and the
    edge / node invariants with distances from the distance-matrix kernel (distance_matrix.rs:71-159)."""
    from rustworkx_b200 import device_snapshot

    n, m = 2**20, 2**24
    g = rx.directed_gnm_random_graph(n, m, seed=7)
    assert g.num_nodes() == n and g.num_edges() == m
    dg = device_snapshot(g)
    og = oracle.OracleGraph.from_graph(g)
    b = 54321
    src = np.arange(b, b + 32, dtype=np.uint32)  # G(n,m): every vertex is statistically alike
    got = dg.edge_betweenness_partial(src)
    st = dg.stats()
    want, ost = og.edge_betweenness_partial(src, threads=host_threads())
    assert (st["te"], st["te_dag"], st["v"]) == (ost["te"], ost["te_dag"], ost["v"])
    assert_close(got, want)
    # invariants: hop distances of the same sources, null value 0 (unreachable pairs add nothing)
    rows = dg.distance_matrix_rows(b, b + 32, as_undirected=False, null_value=0.0)
    dsum = float(rows.sum())
    reached = float(np.count_nonzero(rows))  # t != s with a path
    assert abs(float(got.sum()) - dsum) <= RTOL * dsum
    node = dg.betweenness_partial(src, False)
    want_n, _ = og.betweenness_partial(src[:8], False, threads=host_threads())
    assert_close(dg.betweenness_partial(src[:8], False), want_n)
    assert abs(float(node.sum()) - (dsum - reached)) <= RTOL * dsum
    # the public entry point hands back one score per edge, keyed by edge index (no holes here)
    assert got.shape == (m,) and np.all(np.isfinite(got)) and got.min() >= 0.0


# ------------------------------------------------------------------------------------- config 5
def test_config5_betweenness_4m_node_power_law_full_size(rx, oracle):
    """BASELINE.json configs[4]: betweenness on a 4M-node power-law graph (BA(4*10^6, 8)): 8
    sources against the oracle, matching work counters, and the sum invariant on 128 sources."""
    from rustworkx_b200 import device_snapshot

    n = 4_000_000
    g = rx.barabasi_albert_graph(n, 8, seed=3)
    dg = device_snapshot(g)
    og = oracle.OracleGraph.from_graph(g)
    rng = np.random.default_rng(11)
    src = rng.choice(n, size=128, replace=False).astype(np.uint32)
    src[0] = 0  # the oldest vertex: the biggest hub (bottom-up stress)
    want, ost = og.betweenness_partial(src[:8], False, threads=host_threads())
    got = dg.betweenness_partial(src[:8], False)
    st = dg.stats()
    assert (st["te"], st["te_dag"], st["v"]) == (ost["te"], ost["te_dag"], ost["v"])
    assert_close(got, want)
    want_ep, _ = og.betweenness_partial(src[:2], True, threads=2)
    assert_close(dg.betweenness_partial(src[:2], True), want_ep)
    part = dg.betweenness_partial(src, False)
    assert dg.stats()["v"] == 128 * n  # connected
    cc = dg.closeness_partial(src, wf_improved=False)
    rhs = float(np.sum(np.rint((n - 1) / cc) - (n - 1)))
    assert abs(float(part.sum()) - rhs) <= RTOL * rhs


# ------------------------------------------------------------------------------------- config 2
def test_config2_grid_1000x1000_closeness_sample_is_bit_exact(rx, oracle):
    """BASELINE.json configs[1]: closeness on grid_graph(1000, 1000) (1998 BFS levels): a sample of
    sources against the oracle, `==` (integer sums, then the reference's float expression order,
    centrality.rs:1209-1220), both wf_improved settings."""
    from rustworkx_b200 import device_snapshot

    g = rx.generators.grid_graph(1000, 1000)
    n = g.num_nodes()
    assert n == 1_000_000 and g.num_edges() == 1_998_000
    dg = device_snapshot(g)
    og = oracle.OracleGraph.from_graph(g)
    rng = np.random.default_rng(2)
    src = np.concatenate([[0, 999, n - 1, 500 * 1000 + 500], rng.choice(n, size=60, replace=False)]).astype(np.uint32)
    for wf in (True, False):
        want, _ = og.closeness_partial(src, wf, threads=host_threads())
        got = dg.closeness_partial(src, wf_improved=wf)
        assert np.array_equal(got, want[src])


# ------------------------------------------------------------------------- f64 sigma overflow
def test_sigma_overflow_to_infinity_matches_the_reference_nan_pattern(rx, oracle):
    """The reference keeps sigma in f64 (centrality.rs:415,433-436).  On a 540x540 lattice the path
    counts pass 1.8e308 and become +inf; `(1 + delta) / sigma` is then 0 and `sigma * coeff`
    (centrality.rs:272-279) turns into inf * 0 = NaN, which spreads to every ancestor.  The CUDA path
    (uint32 counts -> f64 fall-back rows) must reproduce that, position by position: NaN where the
    oracle has NaN, and the finite entries within 1e-9."""
    from rustworkx_b200 import device_snapshot

    side = 540
    g = rx.generators.grid_graph(side, side)
    n = side * side
    dg = device_snapshot(g)
    og = oracle.OracleGraph.from_graph(g)
    seen_partial_nan = False
    for s in (n // 2 + side // 2, 17, 0):  # centre: all finite; near a corner: mixed; corner: all NaN
        src = np.array([s], dtype=np.uint32)
        want, _ = og.betweenness_partial(src, False)
        got = dg.betweenness_partial(src, False)
        assert np.array_equal(np.isnan(got), np.isnan(want)), f"source {s}: NaN positions differ"
        assert np.array_equal(np.isinf(got), np.isinf(want))
        fin = np.isfinite(want)
        assert_close(got[fin], want[fin])
        seen_partial_nan |= bool(0 < np.isnan(want).sum() < n - 2)
        want_ep, _ = og.betweenness_partial(src, True)
        got_ep = dg.betweenness_partial(src, True)
        assert np.array_equal(np.isnan(got_ep), np.isnan(want_ep))
        fin = np.isfinite(want_ep)
        assert_close(got_ep[fin], want_ep[fin])
    assert seen_partial_nan  # the mixed case is really exercised
    src = np.array([17], dtype=np.uint32)
    want_e, _ = og.edge_betweenness_partial(src)
    got_e = dg.edge_betweenness_partial(src)
    assert np.array_equal(np.isnan(got_e), np.isnan(want_e))
    fin = np.isfinite(want_e)
    assert_close(got_e[fin], want_e[fin])
    assert dg.stats()["sources"] == 1


# ------------------------------------------------------------------------------ source ordering
def test_hub_sibling_source_order_matches_the_oracle(rx, oracle):
    """bc_order = 2 (calls with >= 4096 sources are sorted by (highest-degree neighbour, two-hop
    key), so that BFS-tree siblings share a lane): only the grouping of sources changes, so the
    sums must match the oracle and the default order, node / endpoints / edge variants."""
    from rustworkx_b200 import _lib, device_snapshot

    g = rx.barabasi_albert_graph(6000, 4, seed=9)
    og = oracle.OracleGraph.from_graph(g)
    src = np.asarray(g.node_indices(), dtype=np.uint32)
    assert src.shape[0] >= 4096
    res = {}
    for flag in (2, 1):
        _lib.set_option("bc_order", flag)
        try:
            dg = device_snapshot(g)
            res[flag] = (dg.betweenness_partial(src, False), dg.betweenness_partial(src, True),
                         dg.edge_betweenness_partial(src), dg.stats()["te"], dg.stats()["te_dag"])
        finally:
            _lib.set_option("bc_order", 2)
    assert res[1][3:] == res[2][3:]
    for a, b in zip(res[1][:3], res[2][:3]):
        assert_close(a, b)
    threads = host_threads()
    assert_close(res[2][0], og.betweenness_partial(src, False, threads=threads)[0])
    assert_close(res[2][1], og.betweenness_partial(src, True, threads=threads)[0])
    assert_close(res[2][2], og.edge_betweenness_partial(src, threads=threads)[0])


# ---------------------------------------------------------------------- C-ABI snapshot cache
def test_content_keyed_snapshot_cache(rx, oracle):
    """SURVEY 8f #3 behind the C-ABI: the same arrays give the same handle back (nothing is copied or
    rebuilt, the scratch is re-used); any change of content gives a new one; results are unaffected."""
    from rustworkx_b200 import _lib

    L = _lib.lib()
    L.rxc_snapshot_cache_clear()
    g = rx.barabasi_albert_graph(20000, 5, seed=4)
    s = g._snapshot_arrays()
    args = (s["n_bound"], s["live_nodes"], s["e_bound"], s["edge_src"], s["edge_dst"], s["edge_id"], s["directed"])
    a = _lib.DeviceGraph(*args, cached=True)
    assert not a.cache_hit
    src = np.arange(300, dtype=np.uint32)
    first = a.betweenness_partial(src, False)
    b = _lib.DeviceGraph(*args, cached=True)
    assert b.cache_hit and b._h.value == a._h.value  # the SAME handle: CSR, scratch and tuning re-used
    assert np.array_equal(b.closeness_partial(src), a.closeness_partial(src))
    a.close()  # one reference gone; b still works
    assert_close(b.betweenness_partial(src, False), first)
    want, _ = oracle.OracleGraph.from_graph(g).betweenness_partial(src, False, threads=host_threads())
    assert_close(first, want)
    # one endpoint changed -> different content -> a new snapshot
    dst2 = np.array(s["edge_dst"], copy=True)
    dst2[7] = (dst2[7] + 1) % 20000
    c = _lib.DeviceGraph(s["n_bound"], s["live_nodes"], s["e_bound"], s["edge_src"], dst2, s["edge_id"],
                         s["directed"], cached=True)
    assert not c.cache_hit and c._h.value != b._h.value
    # direction is part of the key
    d = _lib.DeviceGraph(s["n_bound"], s["live_nodes"], s["e_bound"], s["edge_src"], s["edge_dst"],
                         s["edge_id"], True, cached=True)
    assert not d.cache_hit
    # least recently used entries leave the cache (2 entries by default); b's own reference keeps it alive
    e = _lib.DeviceGraph(*args, cached=True)
    assert not e.cache_hit  # evicted by c and d
    assert_close(b.betweenness_partial(src, False), first)
    for h in (b, c, d, e):
        h.close()
    assert L.rxc_snapshot_cache_clear() == 2
    assert L.rxc_snapshot_cache_clear() == 0


def test_partial_reduce_without_a_communicator_is_the_partial_call(rx):
    from rustworkx_b200 import device_snapshot

    g = rx.barabasi_albert_graph(5000, 4, seed=2)
    dg = device_snapshot(g)
    src = np.arange(0, 5000, 7, dtype=np.uint32)
    assert_close(dg.betweenness_partial_reduce(src, False, root=0), dg.betweenness_partial(src, False))
    assert_close(dg.betweenness_partial_reduce(src, True, root=0), dg.betweenness_partial(src, True))
    assert_close(dg.edge_betweenness_partial_reduce(src, root=0), dg.edge_betweenness_partial(src))
    with pytest.raises(rx.RxCudaError):
        dg.betweenness_partial_reduce(src, False, root=3)  # no such rank


def test_options_are_snapshotted_per_call(rx, oracle):
    """tunables used to be one unsynchronised global read in the middle of a call.  A
    second thread flips every option that changes the kernel path while calls run; each call works
    on its own copy, so every result still matches the oracle."""
    from rustworkx_b200 import _lib, device_snapshot

    g = rx.barabasi_albert_graph(30000, 6, seed=8)
    og = oracle.OracleGraph.from_graph(g)
    src = np.arange(0, 30000, 29, dtype=np.uint32)
    want, _ = og.betweenness_partial(src, False, threads=host_threads())
    dg = device_snapshot(g)
    stop = threading.Event()

    def flip():
        i = 0
        while not stop.is_set():
            _lib.set_option("bc_sigma", i % 2)        # uint32 <-> f64 sigma rows
            _lib.set_option("bc_order", i % 3)
            _lib.set_option("bc_groups", 1 + i % 4)
            _lib.set_option("hub_degree", 64 if i % 2 else 512)
            i += 1

    t = threading.Thread(target=flip)
    t.start()
    try:
        for _ in range(6):
            assert_close(dg.betweenness_partial(src, False), want)
    finally:
        stop.set()
        t.join()
        for name, val in (("bc_sigma", 0), ("bc_order", 2), ("bc_groups", 0), ("hub_degree", 512)):
            _lib.set_option(name, val)


def test_source_order_is_a_permutation_and_stable_on_slices(rx, oracle):
    """rxc_betweenness_source_order: what a multi-process run cuts into contiguous per-rank slices.
    It must be a permutation of the input (holes respected), ordering a slice again must keep it a
    permutation of the slice (the per-rank call does that), and the sliced sums must add up to the
    single-call result."""
    from rustworkx_b200 import device_snapshot
    from rustworkx_b200.distributed import group_shard

    g = rx.barabasi_albert_graph(9000, 4, seed=21)
    for node in (5, 77, 4000):
        g.remove_node(node)
    dg = device_snapshot(g)
    live = np.asarray(g.node_indices(), dtype=np.uint32)
    order = dg.betweenness_source_order(live)
    assert sorted(order.tolist()) == live.tolist()
    assert not np.array_equal(order, live)  # it does reorder
    total = np.zeros(dg.n_bound)
    for rank in range(3):
        mine = group_shard(order, rank, 3)
        assert sorted(dg.betweenness_source_order(mine).tolist()) == sorted(mine.tolist())
        total += dg.betweenness_partial(mine, False)
    assert_close(total, dg.betweenness_partial(live, False))
    want, _ = oracle.OracleGraph.from_graph(g).betweenness_partial(live, False, threads=host_threads())
    assert_close(total, want)
