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
if rank == 0:
    import oracle
    og = oracle.OracleGraph.from_graph(g)
    want, _ = og.betweenness_centrality(True, True, 50, as_arrays=True)
    want_e, _ = og.edge_betweenness_centrality(False, 50, as_arrays=True)
    assert np.allclose(node, want, rtol=1e-9, atol=1e-12)
    assert np.allclose(edge, want_e, rtol=1e-9, atol=1e-12)
    print("NCCL_OK")
else:
    assert node is None and edge is None
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
