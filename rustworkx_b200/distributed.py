"""Multi-GPU host logic: one process per GPU, sources sharded, ONE reduce at the end.

The reference is single-process: rayon workers take sources from a shared iterator and add into one
vector behind an ``RwLock`` (``rustworkx-core/src/centrality.rs:107-121,280,317``).  Brandes'
contributions are additive over sources, so here every rank

1. snapshots the whole graph onto its own GPU (the CSR is small: 68 MB for the 1M-node config),
2. runs ``rxc_*_partial`` on its round-robin shard of the live sources,
3. takes part in a single sum-reduce of the partial vector: on GPUs the fused
   ``rxc_*_partial_reduce`` (the vector stays in HBM, one in-place ncclReduce over NVLink on the
   handle's stream, only the root reads it back); with a ``gloo`` process group -- the CPU tests --
   ``torch.distributed.reduce`` of the host vector,

and rank 0 applies ``_rescale`` (``centrality.rs:221-252``).  Closeness and distance-matrix outputs
are disjoint per source, so they need no reduction, only a gather of slices.

``torch.distributed`` is plumbing only (rendezvous, the 128-byte NCCL id, the gloo test path).
"""

from __future__ import annotations

import numpy as np

from . import _lib


def shard_sources(sources, rank, world):
    """Round-robin shard: rank r takes sources[r::world] (balances BA's age-correlated degrees)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad shard {rank}/{world}")
    return np.ascontiguousarray(np.asarray(sources)[rank::world])


SHARD_CHUNK = 256  # sources per dealt chunk: whole groups of the kernels (128 or 256 sources)


def group_shard(ordered, rank, world, chunk=SHARD_CHUNK):
    """Rank r's share of an ORDERED source list (``rxc_betweenness_source_order``): the list is cut
    into chunks of ``chunk`` consecutive sources -- whole kernel groups, so the clusters of sibling
    sources the order forms stay together -- and the chunks are dealt in alternating direction.  Dealing (rather
    than one contiguous slice per rank) balances the ranks: a clustered group is cheaper than a
    group from the unclustered tail of the order (measured at N = 2 with one slice per rank: one
    rank waited 240 ms of a 2 s step for the other)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad shard {rank}/{world}")
    ordered = np.asarray(ordered)
    k = ordered.shape[0]
    nchunks = (k + chunk - 1) // chunk
    # boustrophedon: rounds of `world` chunks go out left to right, then right to left, ... -- the cost
    # of a chunk drifts along the order (clusters first, the unclustered tail last), and a plain
    # round-robin would hand rank 0 the cheapest chunk of every round and the last rank the dearest
    mine = [c for c in range(nchunks)
            if (c % world if (c // world) % 2 == 0 else world - 1 - c % world) == rank]
    parts = [ordered[c * chunk: (c + 1) * chunk] for c in mine]
    if not parts:
        return np.ascontiguousarray(ordered[:0])
    return np.ascontiguousarray(np.concatenate(parts))


def _dist():
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return None
    return dist


_COMM_READY = False


def init_nccl_comm():
    """Create librxcuda's own NCCL communicator; the unique id travels over torch.distributed."""
    global _COMM_READY
    dist = _dist()
    if dist is None or dist.get_world_size() == 1 or _COMM_READY:
        return
    import torch

    ident = np.zeros(128, dtype=np.uint8)
    if dist.get_rank() == 0:
        _lib.check(_lib.lib().rxc_comm_unique_id(_lib.p_u8(ident)))
    obj = [ident.tobytes()]
    dist.broadcast_object_list(obj, src=0)
    ident = np.frombuffer(obj[0], dtype=np.uint8).copy()
    _lib.check(_lib.lib().rxc_comm_init(dist.get_rank(), dist.get_world_size(), _lib.p_u8(ident)))
    _COMM_READY = True
    del torch


def reduce_partial(partial, root=0):
    """Sum-reduce a host f64 vector to ``root``.  NCCL (ours) when a CUDA process group is up,
    gloo otherwise (CPU tests).  Returns the reduced vector on root, the local one elsewhere."""
    dist = _dist()
    partial = np.ascontiguousarray(partial, dtype=np.float64)
    if dist is None or dist.get_world_size() == 1:
        return partial
    if dist.get_backend() == "nccl":
        init_nccl_comm()
        _lib.check(_lib.lib().rxc_comm_reduce_sum_f64(_lib.p_f64(partial), partial.shape[0], root))
        return partial
    import torch

    t = torch.from_numpy(partial)
    dist.reduce(t, dst=root, op=dist.ReduceOp.SUM)
    return t.numpy()


def _nccl_group():
    dist = _dist()
    return dist is not None and dist.get_world_size() > 1 and dist.get_backend() == "nccl"


def _reduced_on_device(graph, kind, live, rank, world, endpoints):
    """The fused call: shard -> partial vector in HBM -> in-place ncclReduce -> root reads it back
    (``rxc_*_partial_reduce``); the vector never visits the host before it is combined."""
    from . import device_snapshot

    init_nccl_comm()
    dg = device_snapshot(graph)
    # every rank computes the same grouping order and takes a contiguous slice of it
    mine = group_shard(dg.betweenness_source_order(live), rank, world)
    if kind == "node":
        return dg.betweenness_partial_reduce(mine, endpoints, root=0)
    return dg.edge_betweenness_partial_reduce(mine, root=0)


def _device_partial(graph, kind):
    from . import device_snapshot

    dg = device_snapshot(graph)
    if kind == "node":
        return lambda sources, endpoints: dg.betweenness_partial(sources, endpoints)
    return lambda sources, endpoints: dg.edge_betweenness_partial(sources)


def betweenness_centrality(graph, normalized=True, endpoints=False, *, rank=None, world=None,
                           compute_partial=None):
    """All-sources betweenness with the sources sharded over the process group.

    Returns the dense, rescaled vector (original node indexing) on rank 0 and ``None`` elsewhere.
    ``compute_partial(sources, endpoints) -> f64[n_bound]`` defaults to the CUDA path; the CPU tests
    inject the oracle to exercise the sharding / reduce logic without a GPU.
    """
    dist = _dist()
    if rank is None:
        rank = dist.get_rank() if dist else 0
    if world is None:
        world = dist.get_world_size() if dist else 1
    snap = graph._snapshot_arrays()
    mine = shard_sources(snap["live_nodes"], rank, world)
    if compute_partial is None and _nccl_group():
        total = _reduced_on_device(graph, "node", snap["live_nodes"], rank, world, bool(endpoints))
    else:
        fn = compute_partial or _device_partial(graph, "node")
        total = reduce_partial(fn(mine, bool(endpoints)))
    if rank != 0:
        return None
    return _lib.rescale(total.copy(), snap["live_nodes"].shape[0], normalized, snap["directed"],
                        endpoints)


def edge_betweenness_centrality(graph, normalized=True, *, rank=None, world=None,
                                compute_partial=None):
    """Edge variant of :func:`betweenness_centrality` (dense over original edge indices)."""
    dist = _dist()
    if rank is None:
        rank = dist.get_rank() if dist else 0
    if world is None:
        world = dist.get_world_size() if dist else 1
    snap = graph._snapshot_arrays()
    mine = shard_sources(snap["live_nodes"], rank, world)
    if compute_partial is None and _nccl_group():
        total = _reduced_on_device(graph, "edge", snap["live_nodes"], rank, world, False)
    else:
        fn = compute_partial or _device_partial(graph, "edge")
        total = reduce_partial(fn(mine, False))
    if rank != 0:
        return None
    # edge betweenness always rescales with include_endpoints=true (centrality.rs:211-217)
    return _lib.rescale(total.copy(), snap["live_nodes"].shape[0], normalized, snap["directed"], True)


def _world(rank, world):
    dist = _dist()
    if rank is None:
        rank = dist.get_rank() if dist else 0
    if world is None:
        world = dist.get_world_size() if dist else 1
    return rank, world


def group_betweenness_centrality(graph, group, normalized=True, *, rank=None, world=None,
                                 compute_partial=None):
    """Group betweenness (``rustworkx-core/src/centrality.rs:1963-2101``) with the non-group sources
    sharded over the process group: the per-source sums are additive, so one scalar is reduced.
    ``compute_partial(group, sources) -> float`` defaults to ``rxc_group_betweenness_partial``.
    Returns the value on rank 0 and ``None`` elsewhere."""
    rank, world = _world(rank, world)
    snap = graph._snapshot_arrays()
    members = sorted(set(int(x) for x in group))
    node_count = int(snap["live_nodes"].shape[0])
    if not members or node_count <= 1:
        return 0.0 if rank == 0 else None
    mine = shard_sources(snap["live_nodes"], rank, world)
    if compute_partial is None:
        from . import device_snapshot

        dg = device_snapshot(graph)
        compute_partial = dg.group_betweenness_partial
    part = np.array([compute_partial(np.asarray(members, dtype=np.uint32), mine)], dtype=np.float64)
    total = float(reduce_partial(part)[0])
    if rank != 0:
        return None
    if not snap["directed"]:
        total /= 2.0
    if normalized:
        non_group = node_count - len(members)
        if non_group > 1:
            norm = non_group * (non_group - 1) if snap["directed"] else (non_group * (non_group - 1)) // 2
            total /= float(norm)
    return total


def newman_weighted_closeness_centrality(graph, edge_weight, wf_improved=True, *, rank=None, world=None,
                                         compute_partial=None):
    """Newman's weighted closeness (``centrality.rs:1297-1357``), sources sharded: every rank fills the
    entries of its own sources in a zero vector, one sum-reduce assembles them (entries are disjoint).
    ``edge_weight``: dense vector by original edge index.  ``compute_partial(edge_weight, sources) ->
    f64[len(sources)]`` defaults to ``rxc_weighted_closeness_partial``.  Dense vector over original
    node indices on rank 0, ``None`` elsewhere."""
    rank, world = _world(rank, world)
    snap = graph._snapshot_arrays()
    mine = shard_sources(snap["live_nodes"], rank, world)
    if compute_partial is None:
        from . import device_snapshot

        dg = device_snapshot(graph)
        compute_partial = lambda w, s: dg.weighted_closeness_partial(w, s, True, True, wf_improved)  # noqa: E731
    dense = np.zeros(max(int(snap["n_bound"]), 1), dtype=np.float64)
    if mine.size:
        dense[mine.astype(np.int64)] = compute_partial(edge_weight, mine)
    total = reduce_partial(dense)
    return total[: int(snap["n_bound"])] if rank == 0 else None


def closeness_centrality(graph, wf_improved=True, *, rank=None, world=None, compute_partial=None):
    """``closeness_centrality`` (``centrality.rs:1166-1224``) with the sources sharded: closeness is a
    per-source value, so every rank fills the entries of its own sources in a zero vector and one
    sum-reduce assembles them (x + 0 is exact: the result stays bit-identical to a single-GPU call).
    ``compute_partial(sources, wf_improved) -> f64[len(sources)]`` defaults to
    ``rxc_closeness_partial``.  Dense vector over original node indices on rank 0, ``None`` elsewhere."""
    rank, world = _world(rank, world)
    snap = graph._snapshot_arrays()
    mine = shard_sources(snap["live_nodes"], rank, world)
    if compute_partial is None:
        from . import device_snapshot

        dg = device_snapshot(graph)
        compute_partial = lambda s, wf: dg.closeness_partial(s, wf)  # noqa: E731
    dense = np.zeros(max(int(snap["n_bound"]), 1), dtype=np.float64)
    if mine.size:
        dense[mine.astype(np.int64)] = compute_partial(mine, bool(wf_improved))
    total = reduce_partial(dense)
    return total[: int(snap["n_bound"])] if rank == 0 else None


def distance_matrix(graph, as_undirected=False, null_value=0.0, *, rank=None, world=None,
                    compute_rows=None):
    """Unweighted ``distance_matrix`` (``distance_matrix.rs:71-159``) with the ROWS sharded: rank r
    computes the contiguous block of rows ``[r n / world, (r + 1) n / world)`` of the compacted matrix
    (``rxc_distance_matrix_rows``) and rank 0 gathers the blocks -- outputs are disjoint, there is
    nothing to reduce.  ``compute_rows(row_begin, row_end, as_undirected, null_value) ->
    f64[rows][n]`` defaults to the CUDA path.  The n x n matrix on rank 0, ``None`` elsewhere."""
    rank, world = _world(rank, world)
    n = int(graph._snapshot_arrays()["live_nodes"].shape[0])
    lo, hi = n * rank // world, n * (rank + 1) // world
    if compute_rows is None:
        from . import device_snapshot

        dg = device_snapshot(graph)
        compute_rows = dg.distance_matrix_rows
    block = np.ascontiguousarray(compute_rows(lo, hi, bool(as_undirected), float(null_value)),
                                 dtype=np.float64).reshape(hi - lo, n)
    dist = _dist()
    if dist is None or world == 1:
        return block
    import torch

    cuda = dist.get_backend() == "nccl"
    mine = torch.from_numpy(block)
    if cuda:
        mine = mine.cuda()
    if rank == 0:
        parts = [torch.empty((n * (r + 1) // world - n * r // world, n), dtype=torch.float64,
                             device=mine.device) for r in range(world)]
        parts[0] = mine
        ops = [dist.irecv(parts[r], src=r) for r in range(1, world)]
        for op in ops:
            op.wait()
        return torch.cat(parts, 0).cpu().numpy()
    dist.send(mine, dst=0)
    return None
