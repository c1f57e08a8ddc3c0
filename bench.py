#!/usr/bin/env python
"""bench.py -- all-sources betweenness GTEPS on B200 (BASELINE.json metric), driver contract.

Workload (config.workload): ``betweenness_centrality`` on ``barabasi_albert_graph(1_000_000, 8,
seed=1)`` -- BASELINE.json configs[2], the configuration the metric is quoted on; it fits one GPU.

A *step* is one pass of the hot path over one batch of synthetic input: a FIXED set of
``--sources-per-step`` sources (default 65536, a seeded uniform sample of the 10^6 sources of the
full job), whatever the number of GPUs -- **strong scaling**.  The step's sources are put into the
library's grouping order, cut into chunks of 256 (whole kernel groups) and dealt round-robin to the
N ranks (one process per GPU, CSR
replicated); every rank runs forward BFS + sigma
and the dependency sweep for its share and accumulates a partial centrality vector in HBM; ONE
in-place ``ncclReduce`` over NVLink combines the partial vectors and rank 0 reads the result back
(``rxc_betweenness_partial_reduce``).  That is exactly the shape of the full job, cut to 6.5 % of it.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm
    python bench.py --impl reference [...]                        # the reference's CPU algorithm

* ``value``  : GTEPS with the graph snapshot already resident in HBM: TE of the K steps / device time
               (CUDA events on the library's own stream around each call's kernels + the reduce,
               summed over the steps, max over ranks).  TE = (source, arc) pairs traversed, counted
               once per source although Brandes sweeps each twice.
* ``e2e``    : the same metric through the C-ABI with HOST buffers: every step re-snapshots the graph
               on every rank from edge arrays in pinned host memory (H2D + device CSR build), runs the
               sharded step incl. the reduce and reads the vector back into pinned host memory on
               rank 0; wall clock between barriers, max over ranks.
* ``roofline``: the dominant kernel against the measured HBM copy peak (MEASURED_PEAKS.json), twice:
               ``frac`` on ALGORITHMIC bytes (SURVEY 8d model) and ``frac_dram`` on the DRAM bytes ncu
               counted for the same kernel (committed capture under profiles/), both over the kernel's
               live CUDA-event time.
* ``parity`` : correctness of THIS run's numbers, checked on rank 0 outside the timed region:
               sum_v reduced[v] = sum_s sum_t (d(s,t) - 1) with the right-hand side from the bit-exact
               distance kernel; at N > 1 the reduced vector against a single-GPU recomputation of the
               same step; and a 4-source sample against the CPU oracle.
* ``cpu_baseline``: the oracle port of the reference algorithm on the box's host cores (rank 0).
* ``work.full_run_s`` (N = 8, or ``--full-run``): the whole job -- all 10^6 sources -- wall clock
               between barriers incl. snapshot on every rank, reduce and read-back, with its invariant.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "all_sources_betweenness_gteps"
UNIT = "GTEPS"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
TRAFFIC_FILES = ("r02_traffic_step.json", "r01_traffic_batch1024.json")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nodes", type=int, default=1_000_000)
    ap.add_argument("--m", type=int, default=8)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--sources-per-step", type=int, default=65536,
                    help="sources of one step, summed over all GPUs (strong scaling)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="sources in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--full-run", action="store_true",
                    help="also run the whole job (every vertex a source); default at --gpus 8")
    return ap.parse_args()


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def host_threads():
    """Host cores this process may run on.  NOT omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1, which made round 1's N >= 2 reference arm single-threaded."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def build_graph(args):
    # librxgen.so only: the generators are not in the product library
    from rustworkx_b200.random_graph import barabasi_albert_graph

    return barabasi_albert_graph(args.nodes, args.m, seed=args.seed)


def source_plan(n, seed=12345):
    return np.random.default_rng(seed).permutation(n).astype(np.uint32)


def step_sources(plan, step, total):
    n = plan.shape[0]
    idx = np.arange(step * total, (step + 1) * total) % n
    return plan[idx]


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per SOURCE and kernel from the committed `ncu --set full` capture of this workload
    (profiles/, made with scripts/ncu_traffic.py): {kernel: bytes_per_source}, provenance."""
    for name in TRAFFIC_FILES:
        path = os.path.join(ROOT, "profiles", name)
        try:
            with open(path) as f:
                rec = json.load(f)
            sources = int(rec.get("_sources", 1024))
            out = {k: {"dram_bytes_per_source": v["dram_bytes"] / sources,
                       "launches_nonempty": v.get("launches_nonempty")}
                   for k, v in rec.items() if isinstance(v, dict) and "dram_bytes" in v}
            return out, f"profiles/{name} (ncu --set full, {sources}-source capture of this workload)"
        except Exception:
            continue
    return {}, "no ncu capture committed"


def probe_lines():
    """128-byte line requests per SOURCE and kernel, counted by the RXC_PROBE build on this workload
    (profiles/r02_probe_lines.json, scripts/probe_lines.py), and what the memory system delivers for
    such requests (profiles/r02_sector_gather.txt, scripts/micro/sector_gather.cu: random 1 KB rows out of
    24 GB, any 1-4 sectors of a 128-byte line cost the same line request)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_probe_lines.json")) as f:
            rec = json.load(f)
    except Exception:
        return None
    ceiling = None
    try:
        rates = []
        with open(os.path.join(ROOT, "profiles", "r02_sector_gather.txt")) as f:
            for ln in f:
                if ln.startswith("24 GB") and "Gblocks/s" in ln:
                    toks = ln.split()
                    blocks = float(toks[toks.index("Gblocks/s") - 1])
                    sectors = int(ln.split("sectors/block")[1].split(":")[0])
                    rates.append(blocks * (2 if sectors == 8 else 1))  # 8 sectors = both lines of a block
        ceiling = max(rates) if rates else None
    except Exception:
        pass
    return {"per_source": {k: rec[k] for k in ("bc_forward_kernel", "bc_backward_kernel") if k in rec},
            "probe_sources": rec.get("sources"), "ceiling_glines_s": ceiling}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, val in zip(names, r[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU legs: the oracle port of the reference algorithm (the only place bench.py executes oracle/)
# --------------------------------------------------------------------------------------------------
def workload_config(args, graph):
    """`config`, identical on both arms (the driver compares it key by key)"""
    return {
        "workload": (f"betweenness_centrality on barabasi_albert_graph({args.nodes}, {args.m}, "
                     f"seed={args.seed}) [BASELINE.json configs[2]]"),
        "nodes": graph.num_nodes(), "edges": graph.num_edges(),
        "sources": "seeded uniform sample of the graph's sources (numpy default_rng(12345) permutation)",
    }


def cpu_oracle_graph(graph):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle

    return oracle, oracle.OracleGraph.from_graph(graph)


def cpu_sample_run(og, sources, threads):
    t0 = time.perf_counter()
    vec, st = og.betweenness_partial(sources, False, threads=threads)
    dt = time.perf_counter() - t0
    return st["te"] / dt / 1e9, dt, st, vec


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    graph = build_graph(args)
    oracle, og = cpu_oracle_graph(graph)
    threads = host_threads()
    plan = source_plan(graph.num_nodes())
    per_step = args.cpu_sample or 2 * threads
    times, te = [], 0
    for step in range(args.warmup + args.steps):
        src = plan[-(step + 1) * per_step: plan.shape[0] - step * per_step]
        gteps, dt, st, _ = cpu_sample_run(og, src, threads)
        if step >= args.warmup:
            times.append(dt)
            te += st["te"]
    total = sum(times)
    value = te / total / 1e9
    sample = (f"{per_step} sources/step x {args.steps} steps (uniform random sample of the same graph, "
              f"{total:.1f} s of wall clock on {threads} threads)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, graph),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample, "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "work": {"sources_per_step": per_step, "timing": "wall clock around the oracle call, all host threads"},
        "note": "reference Rust crate cannot be built here (no cargo/rustc); this is oracle/ (C "
                "restatement at -O3, OpenMP over sources with an explicit thread count, one global "
                "lock as in centrality.rs:107,280).  A real rustworkx install, were one importable, has no "
                "source-subset entry point: its betweenness_centrality always runs all 10^6 sources "
                "(>= 17 h here), so a bounded sample of this workload can only be timed through the port",
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def rel_err(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-12))) if got.size else 0.0


def run_b200(args):
    rank, local_rank, world = dist_env()
    from rustworkx_b200 import _lib

    torch = None
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if _lib.device_count() == 0:
        raise RuntimeError("bench.py needs a CUDA device: rustworkx_b200 has no CPU fallback")
    _lib.set_device(local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def allreduce(vals, op):
        if dist is None:
            return list(vals)
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return t.tolist()

    graph = build_graph(args)
    n = graph.num_nodes()
    snap = graph._snapshot_arrays()
    plan = source_plan(n)
    T = args.sources_per_step

    def my_sources(step, handle=None):
        """This rank's share of the step: chunks of the library's grouping order of the step's
        sources (every rank computes the same order), so sibling clusters stay on one rank."""
        from rustworkx_b200.distributed import group_shard

        ordered = (handle or dg).betweenness_source_order(step_sources(plan, step, T))
        return group_shard(ordered, rank, world)

    # Host inputs live in page-locked memory (rxc_host_alloc), as a binding that assembles the edge
    # arrays itself would hold them: the snapshot's H2D is then one DMA per array.  A graph from which
    # no edge was removed passes edge_id = NULL ("the position in the list").
    ids_are_positions = bool(snap.get("edge_id_is_position"))
    host_in = {k: _lib.pinned_copy(snap[k]) for k in ("live_nodes", "edge_src", "edge_dst")}
    host_in["edge_id"] = None if ids_are_positions else _lib.pinned_copy(snap["edge_id"])
    host_out = _lib.pinned_empty(max(int(snap["n_bound"]), 1), np.float64)

    def make_device_graph():
        return _lib.DeviceGraph(snap["n_bound"], host_in["live_nodes"], snap["e_bound"],
                                host_in["edge_src"], host_in["edge_dst"], host_in["edge_id"],
                                snap["directed"])

    # multi-GPU: our own NCCL communicator for the one collective of the path
    if world > 1:
        ident = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            _lib.check(_lib.lib().rxc_comm_unique_id(_lib.p_u8(ident)))
        t = torch.from_numpy(ident).cuda()
        dist.broadcast(t, 0)
        ident = t.cpu().numpy().copy()
        _lib.check(_lib.lib().rxc_comm_init(rank, world, _lib.p_u8(ident)))

    dg = make_device_graph()  # inputs resident in HBM before the timed region

    order_ms = [0.0]

    def run_step(handle, step, out):
        """one sharded step: order the step's sources, take this rank's slice -> partial vector in
        HBM -> ncclReduce -> rank 0"""
        mine = my_sources(step, handle)
        order_ms[0] = handle.stats()["ms_total"]  # device time of the ordering kernels
        return handle.betweenness_partial_reduce(mine, False, root=0, out=out)

    # warm-up: also pays NCCL's lazy connection set-up for a vector of the real length
    for step in range(args.warmup):
        run_step(dg, step, host_out)

    sampler = ClockSampler(local_rank)
    sampler.start()
    agg = {"te": 0, "te_dag": 0, "v": 0, "launches": 0, "ms_total": 0.0, "ms_forward": 0.0,
           "ms_backward": 0.0, "ms_reduce": 0.0, "ms_reduce_wait": 0.0, "levels": 0, "batches": 0, "sources": 0}
    last_vec = None
    barrier()
    wall0 = time.perf_counter()
    for step in range(args.warmup, args.warmup + args.steps):
        vec = run_step(dg, step, host_out)
        st = dg.stats()
        for k in agg:
            agg[k] += st[k]
        agg["ms_total"] += order_ms[0]
        if rank == 0 and step == args.warmup + args.steps - 1:
            last_vec = np.array(vec, copy=True)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()

    # CUDA-event time of the K steps on this rank: kernels + waiting for the slowest rank + the reduce
    device_ms = agg["ms_total"] + agg["ms_reduce_wait"] + agg["ms_reduce"]
    device_ms_max, wall_ms_max = allreduce([device_ms, wall * 1e3], "MAX")
    te_all, dag_all, v_all, launches_all, src_all = allreduce(
        [agg["te"], agg["te_dag"], agg["v"], agg["launches"], agg["sources"]], "SUM")
    value = te_all / (device_ms_max * 1e-3) / 1e9

    # ---- e2e: C-ABI with host buffers; snapshot + sharded step + reduce + readback per step --------
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, args.steps)
        for step in range(2):  # warm-up: the pinned staging slots and the pool's second workspace
            g2 = make_device_graph()
            run_step(g2, step, host_out)
            g2.close()
        barrier()
        t0 = time.perf_counter()
        te_e2e = 0
        phase = [0.0, 0.0, 0.0]  # wall clock of snapshot / call / free, summed over the steps
        for step in range(e2e_steps):
            ta = time.perf_counter()
            g2 = make_device_graph()
            tb = time.perf_counter()
            run_step(g2, args.warmup + step, host_out)
            tc_ = time.perf_counter()
            te_e2e += g2.stats()["te"]
            g2.close()
            td = time.perf_counter()
            phase[0] += tb - ta
            phase[1] += tc_ - tb
            phase[2] += td - tc_
        barrier()
        dt = time.perf_counter() - t0
        (dt,) = allreduce([dt], "MAX")
        (te_e2e,) = allreduce([float(te_e2e)], "SUM")
        # per rank: live nodes + (src, dst[, id]) per edge + its share of the sources
        h2d = 4 * n + (8 if ids_are_positions else 12) * graph.num_edges() + 4 * len(my_sources(0))
        e2e = {"value": te_e2e / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d * world),
               "d2h_bytes_per_step": int(8 * n), "steps": e2e_steps,
               "ms_per_step": {"snapshot": 1e3 * phase[0] / e2e_steps, "call": 1e3 * phase[1] / e2e_steps,
                               "free": 1e3 * phase[2] / e2e_steps, "total": 1e3 * dt / e2e_steps},
               "includes": "on every rank H2D of the edge list from pinned host memory + device CSR build "
                           "+ kernels; the in-HBM ncclReduce; D2H of the vector into pinned host memory "
                           "on rank 0"}

    # ---- the whole job (every vertex a source), N = 8 or --full-run ------------------------------
    full = None
    if args.full_run or world == 8:
        from rustworkx_b200.distributed import group_shard

        barrier()
        t0 = time.perf_counter()
        g3 = make_device_graph()
        mine = group_shard(g3.betweenness_source_order(np.arange(n, dtype=np.uint32)), rank, world)
        vec = g3.betweenness_partial_reduce(mine, False, root=0, out=host_out)
        barrier()
        full_s = time.perf_counter() - t0
        st = g3.stats()
        (full_dev_ms,) = allreduce([st["ms_total"] + st["ms_reduce_wait"] + st["ms_reduce"]], "MAX")
        (full_te,) = allreduce([float(st["te"])], "SUM")
        # invariant: sum_v bc[v] = sum_s sum_t (d(s,t) - 1), right-hand side from the distance kernel
        cc = g3.closeness_partial(mine, wf_improved=False)
        rhs_part = float(np.sum(np.rint((n - 1) / cc) - (n - 1)))
        (rhs,) = allreduce([rhs_part], "SUM")
        g3.close()
        if rank == 0:
            got = float(np.sum(vec))
            full = {"full_run_s": full_s, "full_run_device_s": full_dev_ms * 1e-3,
                    "full_run_gteps_wall": full_te / full_s / 1e9, "full_run_sources": n,
                    "full_run_invariant_rel_err": abs(got - rhs) / rhs,
                    "full_run_includes": "snapshot on every rank, all kernels, ncclReduce, D2H on rank 0"}

    # ---- parity of this run (outside the timed region) -------------------------------------------
    parity = None
    if not args.no_parity:
        last = args.warmup + args.steps - 1
        cc = dg.closeness_partial(my_sources(last), wf_improved=False)  # bit-exact distance kernel
        (rhs,) = allreduce([float(np.sum(np.rint((n - 1) / cc) - (n - 1)))], "SUM")
        if rank == 0:
            parity = {"invariant": "sum_v reduced[v] == sum_s sum_t (d(s,t)-1) over the last step's sources",
                      "invariant_rel_err": abs(float(np.sum(last_vec)) - rhs) / rhs, "n_ranks": world}
            if world > 1:  # the reduced vector of N ranks against one GPU doing the whole step alone
                alone = dg.betweenness_partial(step_sources(plan, last, T), False)
                parity["vs_single_gpu_max_rel_err"] = rel_err(last_vec, alone)
            oracle, og = cpu_oracle_graph(graph)
            src4 = np.ascontiguousarray(step_sources(plan, last, T)[:4])
            want, _ = og.betweenness_partial(src4, False, threads=min(4, host_threads()))
            parity["oracle_sample"] = {"sources": 4, "max_rel_err": rel_err(dg.betweenness_partial(src4, False), want)}
            errs = [parity["invariant_rel_err"], parity["oracle_sample"]["max_rel_err"],
                    parity.get("vs_single_gpu_max_rel_err", 0.0)]
            parity["tolerance"] = 1e-9
            parity["ok"] = bool(all(e <= 1e-9 for e in errs))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (algorithmic bytes, SURVEY.md 8d / DESIGN.md) ------------
    peak, peak_src = hbm_peak()
    te, dag, v = agg["te"], agg["te_dag"], agg["v"]  # this rank's counts / times
    bytes_fwd = 8 * te + 8 * dag + 20 * v
    bytes_bwd = 8 * te + 16 * dag + 24 * v
    useful_fwd = 4 * te + 4 * dag + 8 * v     # col; one uint32 path count per DAG arc; own count + mask
    useful_bwd = 4 * te + 8 * dag + 20 * v    # col; one f64 coefficient per DAG arc; own sigma, row, score
    traffic, traffic_src = ncu_traffic()
    kernels = {
        "bc_forward_kernel": {"ms": agg["ms_forward"], "bytes": bytes_fwd, "useful": useful_fwd},
        "bc_backward_kernel": {"ms": agg["ms_backward"], "bytes": bytes_bwd, "useful": useful_bwd},
    }
    for k, kv in kernels.items():
        per_src = traffic.get(k, {}).get("dram_bytes_per_source")
        kv["dram"] = per_src * agg["sources"] if per_src else None
    dom = max(kernels, key=lambda k: kernels[k]["ms"])

    def gbs(b, ms):
        return b / (ms * 1e-3) / 1e9 if (b is not None and ms) else None

    achieved = gbs(kernels[dom]["bytes"], kernels[dom]["ms"])
    dram_rate = gbs(kernels[dom]["dram"], kernels[dom]["ms"])
    whole = gbs(bytes_fwd + bytes_bwd, agg["ms_total"])
    dram_all = None if any(kv["dram"] is None for kv in kernels.values()) else sum(kv["dram"] for kv in kernels.values())
    launches_dom = max(1, agg["levels"])
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak,
        "frac_dram": dram_rate / peak if dram_rate else None,
        "achieved_dram": dram_rate,
        "traffic": kernels[dom]["dram"] / launches_dom if kernels[dom]["dram"] else None,
        "traffic_source": traffic_src,
        "algorithmic_bytes_per_launch": kernels[dom]["bytes"] / launches_dom,
        "useful_bytes_per_launch": kernels[dom]["useful"] / launches_dom,
        "launches": launches_dom,
        "peak_source": peak_src,
        "note": "frac = SURVEY 8d algorithmic bytes / live CUDA-event time / peak (a MODEL fraction: the "
                "uint32 path counts and 128-source rows move fewer bytes than the per-source model); "
                "frac_dram = DRAM bytes ncu counted for this kernel (per source, committed capture) / the "
                "same live time / peak = how busy HBM really is; useful = bytes of values actually consumed",
        "whole_step": {"achieved": whole, "frac": whole / peak,
                       "frac_dram": gbs(dram_all, agg["ms_total"]) / peak if dram_all else None,
                       "bytes_model": "16*TE + 24*TE_dag + 44*V"},
        "frac_of_8TBs_nominal": achieved / 8000.0,
        "per_kernel": {k: {"ms": kv["ms"], "GBs_model": gbs(kv["bytes"], kv["ms"]),
                           "GBs_dram": gbs(kv["dram"], kv["ms"]), "GBs_useful": gbs(kv["useful"], kv["ms"])}
                       for k, kv in kernels.items()},
    }
    pl = probe_lines()
    if pl:  # the memory system counts 128-byte lines, not bytes
        lines = {}
        for k, kv in kernels.items():
            ps = pl["per_source"].get(k)
            if ps and kv["ms"]:
                per_s = agg["sources"] / (kv["ms"] * 1e-3) / 1e9
                lines[k] = {"row_lines_G_per_s": ps["lines128"] * per_s, "lookups_G_per_s": ps["lookups"] * per_s,
                            "sectors_per_row_line": ps["sectors32"] / ps["lines128"]}
        roofline["line_requests"] = {
            "per_kernel": lines, "ceiling_G_per_s": pl["ceiling_glines_s"],
            "source": f"profiles/r02_probe_lines.json (RXC_PROBE build, {pl['probe_sources']} sources of this workload) "
                      "over this run's CUDA-event times; ceiling = profiles/r02_sector_gather.txt (random-row gather "
                      "microbenchmark, 24 GB)",
            "note": "128-byte lines requested by the row gathers and 32-byte level-mask look-ups (one line each) per "
                    "second, averaged over all launches of the kernel; L2 hits are not subtracted"}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        oracle, og = cpu_oracle_graph(graph)
        threads = host_threads()
        k = args.cpu_sample or 2 * threads
        gteps, dt, st, _ = cpu_sample_run(og, plan[-k:], threads)
        if dt < 8.0:  # grow the sample to roughly 10-30 s of CPU work
            k2 = int(min(n, max(k, k * 15.0 / max(dt, 1e-3))))
            k = max(threads, (k2 // threads) * threads)
            gteps, dt, st, _ = cpu_sample_run(og, plan[-k:], threads)
        cpu_baseline = {"value": gteps, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{k} uniformly sampled sources of the same graph, {dt:.1f} s",
                        "extrapolated_full_run_s": n * 2.0 * graph.num_edges() / (gteps * 1e9)}

    work = {"te": te_all, "te_dag": dag_all, "v": v_all, "levels": agg["levels"],
            "batches": agg["batches"], "wall_ms": wall_ms_max, "reduce_ms_per_step": agg["ms_reduce"] / max(args.steps, 1),
            "reduce_wait_ms_per_step_rank0": agg["ms_reduce_wait"] / max(args.steps, 1),
            "sources_per_step": T, "sources_per_step_per_gpu": int(src_all / max(args.steps, 1) / world),
            "host_gap_ms_per_step": (wall_ms_max - device_ms_max) / max(args.steps, 1),
            "sharding": f"each step's {T} sources in the library's grouping order, chunks of 256 dealt round-robin "
                        f"to {world} rank(s), CSR replicated, one in-HBM ncclReduce per step",
            "l2": "inputs larger than L2 (per-batch state >= 2 GB vs 126 MB L2)",
            "timing": "CUDA events on the library stream (rxc_stats.ms_total + ms_reduce_wait + ms_reduce: kernels, the "
                      "device-side rendezvous in front of the reduce, the ncclReduce), max over ranks"}
    if full:
        work.update(full)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": device_ms_max / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, graph),
        "e2e": e2e, "gpu_launches": int(launches_all), "clocks": clocks, "roofline": roofline,
        "cpu_baseline": cpu_baseline, "parity": parity, "work": work,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


_JSON_FD = None


def emit(line):
    """The ONE JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse_args()
    # Libraries print to stdout behind our back (NCCL's "NCCL version ..." banner on rank 0): keep the
    # real stdout for the JSON line and send everything else written to fd 1 to stderr.
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
