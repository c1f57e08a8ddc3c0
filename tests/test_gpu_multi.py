"""N>1 on real GPUs: two ranks (torchrun-style, NCCL) shard the sources, reduce once, and rank 0
compares with the oracle.  Skipped when fewer than two GPUs are visible."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle"))
import numpy as np, torch, torch.distributed as dist
import rustworkx_b200 as rx
from rustworkx_b200 import _lib, distributed as rxd
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank); _lib.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
g = rx.barabasi_albert_graph(5000, 4, seed=9)
g.remove_node(17)
node = rxd.betweenness_centrality(g, True, True)
edge = rxd.edge_betweenness_centrality(g, False)
group = [0, 5, 1234]
gb = rxd.group_betweenness_centrality(g, group, True)
_, eb = g._bounds()
w = np.random.default_rng(3).uniform(0.1, 2.0, size=eb)
nw = rxd.newman_weighted_closeness_centrality(g, w, True)
cc = rxd.closeness_centrality(g, True)
h = rx.generators.heavy_hex_graph(7)
dm = rxd.distance_matrix(h, False, -1.0)
if rank == 0:
    import oracle
    og = oracle.OracleGraph.from_graph(g)
    want, _ = og.betweenness_centrality(True, True, 50, as_arrays=True)
    want_e, _ = og.edge_betweenness_centrality(False, 50, as_arrays=True)
    assert np.allclose(node, want, rtol=1e-9, atol=1e-12)
    assert np.allclose(edge, want_e, rtol=1e-9, atol=1e-12)
    assert abs(gb - og.group_betweenness_centrality(group, True, 50)) <= 1e-9 * abs(gb)
    want_nw, _ = og.newman_weighted_closeness_centrality(w, True, 50, as_arrays=True)
    assert np.allclose(nw, want_nw, rtol=1e-12, atol=0.0)
    want_cc, _ = og.closeness_centrality(True, 50, as_arrays=True)
    assert np.array_equal(cc, want_cc)  # per-source values: sharding keeps them bit-identical
    assert np.array_equal(dm, oracle.OracleGraph.from_graph(h).distance_matrix(300, False, -1.0))
    print("NCCL_OK")
else:
    assert node is None and edge is None and cc is None and dm is None
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_nccl_betweenness(tmp_path):
    from rustworkx_b200 import _lib

    if _lib.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:]
    assert "NCCL_OK" in proc.stdout


def test_in_process_replicas(oracle):
    """rxc_set_devices: one process, the snapshot replicated on two GPUs, full calls sharded over
    them with one host thread per device -- same results as the single-device path / the oracle."""
    import numpy as np

    import rustworkx_b200 as rx
    from rustworkx_b200 import _lib

    if _lib.device_count() < 2:
        pytest.skip("needs two GPUs")
    rx.set_devices([0, 1])
    try:
        g = rx.barabasi_albert_graph(6000, 5, seed=4)
        g.remove_node(11)
        og = oracle.OracleGraph.from_graph(g)
        want, _ = og.betweenness_centrality(True, True, 50, as_arrays=True)
        got = rx.betweenness_centrality(g, endpoints=True)
        keys = np.array(list(got.keys()))
        assert np.allclose(np.array(list(got.values())), want[keys], rtol=1e-9, atol=1e-12)
        st = rx.device_snapshot(g).stats()
        assert st["sources"] == g.num_nodes() and st["batches"] >= 2  # both shards worked
        want_e, _ = og.edge_betweenness_centrality(False, 50, as_arrays=True)
        got_e = rx.edge_betweenness_centrality(g, normalized=False)
        assert np.allclose(np.array(list(got_e.values())), want_e[np.array(list(got_e.keys()))],
                           rtol=1e-9, atol=1e-12)
        want_c, _ = og.closeness_centrality(True, 50, as_arrays=True)
        assert np.array_equal(np.array(list(rx.closeness_centrality(g).values())), want_c[keys])
        h = rx.generators.heavy_hex_graph(9)
        oh = oracle.OracleGraph.from_graph(h)
        assert np.array_equal(rx.distance_matrix(h, null_value=-1.0), oh.distance_matrix(300, True, -1.0))
        assert rx.unweighted_average_shortest_path_length(h) == oh.average_shortest_path_length(300, True, False)
        # 8f #2 / #4 through the sharded full calls
        group = [3, 77, 4000]
        assert abs(rx.group_betweenness_centrality(g, group) - og.group_betweenness_centrality(group, True, 50)) \
            <= 1e-9 * og.group_betweenness_centrality(group, True, 50)
        _, eb = g._bounds()
        w = np.random.default_rng(1).uniform(0.1, 2.0, size=eb)
        dg = rx.device_snapshot(g)
        want_nw, _ = og.newman_weighted_closeness_centrality(w, True, 50, as_arrays=True)
        assert np.allclose(dg.newman_weighted_closeness(w, True), want_nw, rtol=1e-12, atol=0.0)
        live = np.asarray(g.node_indices(), dtype=np.uint32)
        want_rows = og.dijkstra_lengths_rows(w, live[:300], threads=4)[:, live]
        got_rows = dg.all_pairs_dijkstra_lengths_rows(w, 0, 300)
        assert np.array_equal(np.isnan(got_rows), np.isnan(want_rows))
        assert np.array_equal(got_rows[~np.isnan(got_rows)], want_rows[~np.isnan(want_rows)])
    finally:
        rx.set_devices([])
