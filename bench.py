#!/usr/bin/env python
"""bench.py -- all-sources betweenness GTEPS on B200 (BASELINE.json metric), driver contract.

Workload (config.workload): ``betweenness_centrality`` on ``barabasi_albert_graph(1_000_000, 8,
seed=1)`` -- BASELINE.json configs[2], the configuration the metric is quoted on; it fits one GPU.
A *step* is one pass of the hot path over one batch of synthetic input: ``--sources-per-step``
sources per GPU (forward BFS + sigma, backward dependency sweep, accumulation into the partial
centrality vector).  Sources are sharded across ranks (weak scaling: per-GPU work is fixed), each rank
accumulates a partial vector, and ONE NCCL reduce at the end of the timed region combines them.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm
    python bench.py --impl reference [...]                        # the reference's CPU algorithm

* ``value``  : GTEPS with the graph snapshot already resident in HBM (device time, CUDA events on the
               library's own stream, max over ranks).  TE = (source, arc) pairs traversed, counted
               once per source although Brandes sweeps each twice.
* ``e2e``    : the same metric through the C-ABI with HOST buffers: every step re-snapshots the graph
               from edge arrays in pinned host memory (H2D + device CSR build), runs the step and
               reads the partial vector back into pinned host memory (D2H); wall clock around the calls.
* ``roofline``: algorithmic bytes of the dominant kernel / its CUDA-event time, against the measured
               HBM copy peak in MEASURED_PEAKS.json.
* ``cpu_baseline``: the oracle port of the reference algorithm on the box's host cores (rank 0, N=1).
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nodes", type=int, default=1_000_000)
    ap.add_argument("--m", type=int, default=8)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--sources-per-step", type=int, default=1024)
    ap.add_argument("--cpu-sample", type=int, default=0, help="sources in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def build_graph(args):
    import rustworkx_b200 as rx

    return rx.barabasi_albert_graph(args.nodes, args.m, seed=args.seed)


def source_plan(n, seed=12345):
    return np.random.default_rng(seed).permutation(n).astype(np.uint32)


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, sources_per_step):
    """dram__bytes_read.sum + dram__bytes_write.sum per (non-empty) launch of `kernel`, from the
    committed `ncu --set full` capture of one 1024-source batch of this same workload (profiles/,
    made with scripts/ncu_traffic.py)."""
    name = "r01_traffic_batch1024.json"
    path = os.path.join(ROOT, "profiles", name)
    try:
        with open(path) as f:
            rec = json.load(f)[kernel]
        return (rec["dram_bytes"] / max(1, rec["launches_nonempty"]),
                f"profiles/{name} (ncu --set full, per non-empty launch of a 1024-source batch)")
    except Exception:
        return None, "no ncu capture committed"


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
def workload_name(args):
    """config.workload, identical on both arms"""
    return (f"betweenness_centrality on barabasi_albert_graph({args.nodes}, {args.m}, "
            f"seed={args.seed}) [BASELINE.json configs[2]]")


def cpu_oracle_graph(graph):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle

    return oracle, oracle.OracleGraph.from_graph(graph)


def cpu_sample_run(oracle, og, sources, threads):
    t0 = time.perf_counter()
    _, st = og.betweenness_partial(sources, False, threads=threads)
    dt = time.perf_counter() - t0
    return st["te"] / dt / 1e9, dt, st


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    graph = build_graph(args)
    oracle, og = cpu_oracle_graph(graph)
    threads = oracle.max_threads()
    plan = source_plan(graph.num_nodes())
    per_step = args.cpu_sample or 2 * threads
    times, te = [], 0
    for step in range(args.warmup + args.steps):
        src = plan[step * per_step:(step + 1) * per_step]
        gteps, dt, st = cpu_sample_run(oracle, og, src, threads)
        if step >= args.warmup:
            times.append(dt)
            te += st["te"]
    total = sum(times)
    value = te / total / 1e9
    sample = (f"{per_step} sources/step x {args.steps} steps of BA({args.nodes},{args.m}) "
              f"(uniform random sample, seed 12345)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "nodes": graph.num_nodes(),
                   "edges": graph.num_edges(), "sources_per_step": per_step,
                   "timing": "wall clock around the oracle call, all host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference Rust crate cannot be built here (no cargo/rustc); this is oracle/ (C "
                "restatement, OpenMP over sources, one global lock as in centrality.rs:107,280)",
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
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

    graph = build_graph(args)
    n = graph.num_nodes()
    snap = graph._snapshot_arrays()
    plan = source_plan(n)
    S = args.sources_per_step

    def sources_for(step):
        i = (step * world + rank) * S
        idx = np.arange(i, i + S) % n
        return np.ascontiguousarray(plan[idx])

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

    # multi-GPU: our own NCCL communicator for the single end-of-run reduce
    if world > 1:
        ident = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            _lib.check(_lib.lib().rxc_comm_unique_id(_lib.p_u8(ident)))
        t = torch.from_numpy(ident).cuda()
        dist.broadcast(t, 0)
        ident = t.cpu().numpy().copy()
        _lib.check(_lib.lib().rxc_comm_init(rank, world, _lib.p_u8(ident)))
        # The first collective of a given size class pays NCCL's lazy connection set-up (hundreds of
        # ms; a small warm-up does not cover the protocol the 8 MB reduce uses): warm the communicator
        # with a vector of the real length, before the timed region.
        warm = np.zeros(max(int(snap["n_bound"]), 1), dtype=np.float64)
        for _ in range(2):
            _lib.check(_lib.lib().rxc_comm_reduce_sum_f64(_lib.p_f64(warm), warm.shape[0], 0))

    dg = make_device_graph()  # inputs resident in HBM before the timed region
    partial = np.zeros(snap["n_bound"], dtype=np.float64)

    for step in range(args.warmup):
        dg.betweenness_partial(sources_for(step), False)

    sampler = ClockSampler(local_rank)
    sampler.start()
    agg = {"te": 0, "te_dag": 0, "v": 0, "launches": 0, "ms_total": 0.0, "ms_forward": 0.0,
           "ms_backward": 0.0, "levels": 0, "batches": 0}
    barrier()
    wall0 = time.perf_counter()
    for step in range(args.warmup, args.warmup + args.steps):
        partial += dg.betweenness_partial(sources_for(step), False)
        st = dg.stats()
        for k in agg:
            agg[k] += st[k]
    reduce_ms = 0.0
    if world > 1:
        t0 = time.perf_counter()
        _lib.check(_lib.lib().rxc_comm_reduce_sum_f64(_lib.p_f64(partial), partial.shape[0], 0))
        reduce_ms = 1e3 * (time.perf_counter() - t0)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()

    device_ms = agg["ms_total"] + reduce_ms  # CUDA-event time of the K steps (+ the reduce)
    if world > 1:
        tm = torch.tensor([device_ms, wall * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        cnt = torch.tensor([agg["te"], agg["te_dag"], agg["v"], agg["launches"]], dtype=torch.float64,
                           device="cuda")
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        device_ms_max, wall_ms_max = tm.tolist()
        te_all, dag_all, v_all, launches_all = cnt.tolist()
    else:
        device_ms_max, wall_ms_max = device_ms, wall * 1e3
        te_all, dag_all, v_all, launches_all = agg["te"], agg["te_dag"], agg["v"], agg["launches"]
    value = te_all / (device_ms_max * 1e-3) / 1e9

    # ---- e2e: C-ABI with host buffers, snapshot + step + readback per step -----------------------
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, args.steps)
        for step in range(2):  # warm-up: the pinned staging slots and the pool's second workspace
            g2 = make_device_graph()
            g2.betweenness_partial(sources_for(step), False, out=host_out)
            g2.close()
        barrier()
        t0 = time.perf_counter()
        te_e2e = 0
        phase = [0.0, 0.0, 0.0]  # wall clock of snapshot / call / free, summed over the steps
        for step in range(e2e_steps):
            src_step = sources_for(args.warmup + step)
            ta = time.perf_counter()
            g2 = make_device_graph()
            tb = time.perf_counter()
            g2.betweenness_partial(src_step, False, out=host_out)
            tc_ = time.perf_counter()
            te_e2e += g2.stats()["te"]
            g2.close()
            td = time.perf_counter()
            phase[0] += tb - ta
            phase[1] += tc_ - tb
            phase[2] += td - tc_
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = tt.item()
            tc = torch.tensor([float(te_e2e)], dtype=torch.float64, device="cuda")
            dist.all_reduce(tc, op=dist.ReduceOp.SUM)
            te_e2e = tc.item()
        # live nodes + (src, dst[, id]) per edge + sources
        h2d = 4 * n + (8 if ids_are_positions else 12) * graph.num_edges() + 4 * S
        e2e = {"value": te_e2e / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(8 * n), "steps": e2e_steps,
               "ms_per_step": {"snapshot": 1e3 * phase[0] / e2e_steps, "call": 1e3 * phase[1] / e2e_steps,
                               "free": 1e3 * phase[2] / e2e_steps, "total": 1e3 * dt / e2e_steps},
               "includes": "H2D of the edge list from pinned host memory + device CSR build + kernels "
                           "+ D2H of the partial vector into pinned host memory"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (algorithmic bytes, SURVEY.md 8d / DESIGN.md) ------------
    peak, peak_src = hbm_peak()
    te, dag, v = agg["te"], agg["te_dag"], agg["v"]  # this rank's counts / times
    bytes_fwd = 8 * te + 8 * dag + 20 * v
    bytes_bwd = 8 * te + 16 * dag + 24 * v
    kernels = {
        "bc_forward_kernel": {"ms": agg["ms_forward"], "bytes": bytes_fwd},
        "bc_backward_kernel": {"ms": agg["ms_backward"], "bytes": bytes_bwd},
    }
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    achieved = kernels[dom]["bytes"] / (kernels[dom]["ms"] * 1e-3) / 1e9
    whole = (bytes_fwd + bytes_bwd) / (agg["ms_total"] * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(dom, S)
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
        "algorithmic_bytes_per_launch": kernels[dom]["bytes"] / max(1, agg["levels"]),
        "peak_source": peak_src,
        "whole_step": {"achieved": whole, "frac": whole / peak,
                       "bytes_model": "16*TE + 24*TE_dag + 44*V"},
        "frac_of_8TBs_nominal": achieved / 8000.0,
        "per_kernel": {k: {"GBs": kv["bytes"] / (kv["ms"] * 1e-3) / 1e9 if kv["ms"] else None,
                           "ms": kv["ms"]} for k, kv in kernels.items()},
    }

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        oracle, og = cpu_oracle_graph(graph)
        threads = oracle.max_threads()
        k = args.cpu_sample or 2 * threads
        src = plan[-k:]
        gteps, dt, st = cpu_sample_run(oracle, og, src, threads)
        if dt < 8.0:  # grow the sample to roughly 10-30 s of CPU work
            k2 = int(min(n, max(k, k * 15.0 / max(dt, 1e-3))))
            k2 = max(threads, (k2 // threads) * threads)
            src = plan[-k2:]
            gteps, dt, st = cpu_sample_run(oracle, og, src, threads)
            k = k2
        cpu_baseline = {"value": gteps, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{k} uniformly sampled sources of the same graph, {dt:.1f} s",
                        "extrapolated_full_run_s": n * 2.0 * graph.num_edges() / (gteps * 1e9)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": device_ms_max / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": workload_name(args),
            "nodes": n, "edges": graph.num_edges(), "sources_per_step_per_gpu": S,
            "sharding": f"sources sharded over {world} rank(s), CSR replicated, one NCCL reduce at end",
            "l2": "inputs larger than L2 (per-batch state >= 2 GB vs 126 MB L2)",
            "timing": "CUDA events on the library stream (rxc_stats.ms_total), max over ranks",
        },
        "e2e": e2e, "gpu_launches": int(launches_all), "clocks": clocks, "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "work": {"te": te_all, "te_dag": dag_all, "v": v_all, "levels": agg["levels"],
                 "batches": agg["batches"], "wall_ms": wall_ms_max, "reduce_ms": reduce_ms},
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
