"""BASELINE.json configs[2] in full: all-sources betweenness of BA(1_000_000, 8, seed=1), sources
sharded over the ranks of `torchrun`, one NCCL reduce, rescale on rank 0 -- then checked with the
size-independent invariant  sum_v BC_unnormalised[v] * 2 = sum_s sum_t (d(s,t) - 1)  (undirected
scores are halved) using all-sources distance sums from the bit-exact distance kernel, and against the
CPU oracle on a few sampled sources (un-rescaled partial sums)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, torch.distributed as dist
import rustworkx_b200 as rx
from rustworkx_b200 import _lib, distributed as rxd

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); _lib.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rxd.init_nccl_comm()
n = int(os.environ.get("N", 1_000_000))
g = rx.barabasi_albert_graph(n, 8, seed=1)
dg = rx.device_snapshot(g)
if world > 1:
    dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
bc = rxd.betweenness_centrality(g, normalized=False, endpoints=False)  # shard -> partial -> reduce -> rescale
if world > 1:
    dist.barrier(); torch.cuda.synchronize()
t_bc = time.perf_counter() - t0
st = dg.stats()
# all-sources distance sums (this rank's shard), reduced the same way
mine = rxd.shard_sources(np.arange(n, dtype=np.uint32), rank, world)
cc = dg.closeness_partial(mine, wf_improved=False)
dsum_local = np.array([float(np.sum(np.rint((n - 1) / cc) - (n - 1)))])
total = rxd.reduce_partial(dsum_local)
if rank == 0:
    te = float(n) * 2.0 * g.num_edges()
    lhs, rhs = float(np.sum(bc)) * 2.0, float(total[0])
    import oracle
    og = oracle.OracleGraph.from_graph(g)
    sample = np.random.default_rng(11).choice(n, size=16, replace=False).astype(np.uint32)
    want, _ = og.betweenness_partial(sample, False, threads=oracle.max_threads())
    got = dg.betweenness_partial(sample, False)
    err = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-12)))
    print(json.dumps({"config": f"betweenness_centrality(barabasi_albert_graph({n}, 8, seed=1))", "n_gpus": world,
                      "seconds": round(t_bc, 3), "gteps": round(te / t_bc / 1e9, 1),
                      "rank0_device_ms": round(st["ms_total"], 1), "rank0_sources": st["sources"],
                      "invariant_rel_err": abs(lhs - rhs) / rhs, "oracle_sample_max_rel_err": err,
                      "bc_max": float(bc.max()), "bc_argmax": int(bc.argmax())}), flush=True)
if world > 1:
    dist.destroy_process_group()
