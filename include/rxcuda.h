/*
 * rxcuda.h -- C-ABI of librxcuda.so: the B200 (sm_100a) implementation of rustworkx's
 * all-sources unweighted BFS hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces one generic
 * function of the reference's `rustworkx-core` crate, i.e. exactly what a `rustworkx-cuda` crate's
 * `extern "C"` block would bind (see INTEGRATION.md for the Rust and ctypes stubs):
 *
 *   rxc_betweenness*       <- rustworkx-core/src/centrality.rs:70-132   betweenness_centrality
 *   rxc_edge_betweenness*  <- rustworkx-core/src/centrality.rs:174-219  edge_betweenness_centrality
 *   rxc_closeness*         <- rustworkx-core/src/centrality.rs:1166-1224 closeness_centrality
 *   rxc_distance_matrix    <- rustworkx-core/src/shortest_path/distance_matrix.rs:71-159
 *   rxc_rescale            <- rustworkx-core/src/centrality.rs:221-252  _rescale
 *   rxc_graph_snapshot     <- the petgraph StableGraph borrowed as `&graph.graph` by the pyo3
 *                             wrappers at src/centrality.rs:83-98,151-166,310-323,371-384,592-606,
 *                             654-668 and src/shortest_path/mod.rs:1559-1573,1601-1614
 *
 * Plain pointers and sizes only; no torch / numpy types.  All functions return 0 on success or a
 * negative rxc_status; rxc_last_error() gives the message for the calling thread.  There is NO CPU
 * fallback: without a CUDA device every compute entry point fails with RXC_ERR_CUDA.
 *
 * Index conventions follow the reference: node-keyed outputs are dense over [0, n_bound) in ORIGINAL
 * node indices (holes left at 0.0 -- the caller drops them, src/centrality.rs:91-97), edge-keyed
 * outputs are dense over [0, e_bound) in original edge indices, and the distance matrix is
 * n_live x n_live over live nodes in ascending-index (compacted) order.
 */
#ifndef RXCUDA_H
#define RXCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    RXC_OK = 0,
    RXC_ERR_INVALID = -1, /* bad argument */
    RXC_ERR_CUDA = -2,    /* CUDA runtime error / no device */
    RXC_ERR_OOM = -3,     /* device or host allocation failed */
    RXC_ERR_NCCL = -4,    /* NCCL not loadable or NCCL error */
    RXC_ERR_LIMIT = -5    /* graph exceeds supported index width */
} rxc_status;

typedef struct rxc_graph rxc_graph;

/* Work and timing counters of the last compute call on a handle. */
typedef struct {
    uint64_t te;          /* traversed (source, arc) pairs: sum_s sum_{v reached} deg(v)        */
    uint64_t te_dag;      /* (source, arc) pairs on shortest-path DAGs (betweenness only)        */
    uint64_t v;           /* (source, vertex) pairs reached                                      */
    uint64_t sources;     /* sources processed                                                   */
    uint64_t levels;      /* BFS level launches (sum over batches)                               */
    uint64_t launches;    /* kernels launched by the call                                        */
    uint64_t batches;     /* source batches                                                      */
    double ms_total;      /* device time, first to last kernel of the call (CUDA events)         */
    double ms_forward;    /* device time in BFS / sigma kernels                                  */
    double ms_backward;   /* device time in dependency kernels                                   */
    double ms_snapshot;   /* host CSR build + H2D of the snapshot (wall clock)                   */
    double ms_d2h;        /* device-to-host result copies                                        */
    double ms_reduce;     /* rxc_*_partial_reduce: the in-place ncclReduce on the handle's stream
                             (CUDA events), after a device-side rendezvous of the ranks          */
    double ms_reduce_wait; /* that rendezvous: how long this rank waited for the slowest one     */
} rxc_stats;

/* ---- library / device ------------------------------------------------------------------------- */
const char *rxc_version(void);
const char *rxc_last_error(void);
int rxc_device_count(void);       /* 0 when no CUDA device is usable */
int rxc_set_device(int device);   /* device used by subsequent rxc_graph_snapshot calls */
int rxc_get_device(void);
/* In-process multi-GPU: subsequent snapshots replicate the CSR on every listed device and the full
 * calls (rxc_betweenness, ..._edge_betweenness, ..._closeness, ..._distance_matrix*, ..._distance_sums)
 * shard the live sources over them, one host thread per device.  n = 0 returns to single-device. */
int rxc_set_devices(const int *devices, int n);

/* Page-locked host buffers.  Every host pointer of this API may be pageable (it is then staged
 * through the library's own pinned slots, ~16 GB/s) or page-locked -- from rxc_host_alloc,
 * cudaHostAlloc or cudaHostRegister -- in which case the DMA engine reads / writes it directly at
 * PCIe speed.  A binding that owns its edge arrays (the snapshot of a StableGraph is assembled by
 * the binding anyway) should assemble them in rxc_host_alloc memory. */
int rxc_host_alloc(size_t bytes, void **out);
void rxc_host_free(void *p);

/* ---- snapshot: StableGraph -> compact int32 CSR/CSC in HBM ---------------------------------------
 * live_nodes[n_live]: ascending original indices of live nodes (PyGraph.node_indices()).
 * edge_src/edge_dst/edge_id[n_edges]: endpoints (original node indices) and original edge index of
 * every live edge (PyGraph.edge_list() / edge_indices()); edge_id == NULL means "the position in the
 * list" (a graph from which no edge was ever removed) and saves a third of the copy.  Parallel edges are kept (path
 * multiplicity, centrality.rs:433-436); self-loops are dropped (no effect on BFS).  The snapshot is
 * a copy: the caller's graph may change afterwards. */
int rxc_graph_snapshot(uint32_t n_bound, const uint32_t *live_nodes, uint32_t n_live,
                       uint32_t e_bound, const uint32_t *edge_src, const uint32_t *edge_dst,
                       const uint32_t *edge_id, uint64_t n_edges, int directed,
                       rxc_graph **out_graph);
void rxc_graph_free(rxc_graph *g);

/* Content-keyed snapshot cache.  The reference re-reads the StableGraph on every
 * call (the bindings borrow `&graph.graph`, src/centrality.rs:89; a caller outside the crate feeds
 * `edge_list()` / `extend_from_edge_list`, src/graph.rs:815-826,983-998) and has no mutation counter
 * a binding could key on, so the key is a 128-bit hash of exactly what rxc_graph_snapshot receives
 * (bounds, live nodes, edge arrays, direction, device set).  A hit returns the SAME handle again
 * (*hit = 1; reference counted): its CSR in HBM, its scratch and its tuned state (sigma row format,
 * queue sizing) are re-used, and nothing is copied.  Every returned handle is released with
 * rxc_graph_free; the cache keeps `snapshot_cache_entries` (option, default 2) handles alive,
 * most recently used first.  rxc_snapshot_cache_clear drops them and returns how many it held. */
int rxc_graph_snapshot_cached(uint32_t n_bound, const uint32_t *live_nodes, uint32_t n_live,
                              uint32_t e_bound, const uint32_t *edge_src, const uint32_t *edge_dst,
                              const uint32_t *edge_id, uint64_t n_edges, int directed,
                              rxc_graph **out_graph, int *hit);
int rxc_snapshot_cache_clear(void);
int rxc_graph_info(const rxc_graph *g, uint32_t *n_bound, uint32_t *n_live, uint32_t *e_bound,
                   uint64_t *n_arcs, int *directed);

/* ---- full calls (all live sources, final scaling applied; outputs are HOST buffers) -------------- */
int rxc_betweenness(rxc_graph *g, int include_endpoints, int normalized,
                    double *out /* n_bound */);
int rxc_edge_betweenness(rxc_graph *g, int normalized, double *out /* e_bound */);
int rxc_closeness(rxc_graph *g, int wf_improved, double *out /* n_bound */);
int rxc_distance_matrix(rxc_graph *g, int as_undirected, double null_value,
                        double *out /* n_live * n_live, row-major */);

/* compute_distance_sum (src/shortest_path/average_length.rs:22-77), the kernel under
 * graph_/digraph_unweighted_average_shortest_path_length (src/shortest_path/mod.rs:1657-1740):
 * sum of all finite shortest-path lengths and the number of connected ordered pairs. */
int rxc_distance_sums(rxc_graph *g, int as_undirected, uint64_t *total_distance,
                      uint64_t *connected_pairs);

/* group centralities.  `group` holds ORIGINAL indices of live nodes (anything else is
 * RXC_ERR_INVALID -- the binding raises ValueError, src/centrality.rs:1216-1222,1263-1269,1318-1324,
 * 1371-1377); duplicates count once.
 *   rxc_group_betweenness <- rustworkx-core/src/centrality.rs:1963-2101 group_betweenness_centrality
 *   rxc_group_closeness   <- rustworkx-core/src/centrality.rs:1860-1916 group_closeness_centrality */
int rxc_group_betweenness(rxc_graph *g, const uint32_t *group, uint64_t n_group, int normalized,
                          double *out /* 1 */);
int rxc_group_closeness(rxc_graph *g, const uint32_t *group, uint64_t n_group, double *out /* 1 */);
/* the per-source sums of centrality.rs:1985-2080 over a subset of sources (members of the group are
 * skipped), before the /2 of undirected graphs and the normalisation; additive over source shards */
int rxc_group_betweenness_partial(rxc_graph *g, const uint32_t *group, uint64_t n_group,
                                  const uint32_t *sources, uint64_t n_sources, double *out_sum);

/* weighted all-sources shortest path lengths.  `edge_weight` holds e_bound doubles by
 * ORIGINAL edge index (slots of removed edges are ignored); weights must be non-negative and not NaN
 * (the bindings reject them: is_valid_weight, src/lib.rs:303-313).  Distances are bit-identical to
 * the reference's binary-heap Dijkstra (min over paths of the left-to-right f64 sum).
 *   rxc_newman_weighted_closeness       <- rustworkx-core/src/centrality.rs:1297-1357 (cost = 1/weight
 *                                          over Reversed(&graph); per-source sum in vertex order)
 *   rxc_all_pairs_dijkstra_lengths_rows <- src/shortest_path/all_pairs_dijkstra.rs:36-105: rows
 *                                          [row_begin, row_end) of the n_live x n_live length matrix in
 *                                          compacted order; `unreachable_value` where there is no path */
int rxc_newman_weighted_closeness(rxc_graph *g, const double *edge_weight, int wf_improved,
                                  double *out /* n_bound */);
int rxc_all_pairs_dijkstra_lengths_rows(rxc_graph *g, const double *edge_weight, double unreachable_value,
                                        uint32_t row_begin, uint32_t row_end,
                                        double *out /* (row_end-row_begin) * n_live */);
/* closeness of the listed nodes from weighted distances; invert: cost = 1/weight; reversed: follow
 * incoming edges.  (invert=1, reversed=1) is Newman's closeness; out[i] for sources[i]. */
int rxc_weighted_closeness_partial(rxc_graph *g, const double *edge_weight, int invert, int reversed,
                                   const uint32_t *sources, uint64_t n_sources, int wf_improved,
                                   double *out /* n_sources */);

/* ---- partial calls: a subset of sources (ORIGINAL node indices), no final scaling.  This is the
 * unit a rank runs when sources are sharded across GPUs; partial vectors are additive.
 * `out` is zero-initialised by the call. ----------------------------------------------------------- */
int rxc_betweenness_partial(rxc_graph *g, const uint32_t *sources, uint64_t n_sources,
                            int include_endpoints, double *out /* n_bound */);
int rxc_edge_betweenness_partial(rxc_graph *g, const uint32_t *sources, uint64_t n_sources,
                                 double *out /* e_bound */);
/* closeness of the listed nodes only: out[i] for sources[i] */
int rxc_closeness_partial(rxc_graph *g, const uint32_t *sources, uint64_t n_sources,
                          int wf_improved, double *out /* n_sources */);
/* rows of the distance matrix for live-node ranks [row_begin, row_end) */
int rxc_distance_matrix_rows(rxc_graph *g, int as_undirected, double null_value,
                             uint32_t row_begin, uint32_t row_end,
                             double *out /* (row_end-row_begin) * n_live */);

/* _rescale (centrality.rs:221-252): multiply by the precomputed reciprocal, on the host. */
void rxc_rescale(double *x, uint64_t len, uint64_t node_count, int normalized, int directed,
                 int include_endpoints);

/* ---- multi-GPU: one process per GPU, sources sharded, ONE reduce at the end ------------------------
 * NCCL is dlopen()ed (libnccl.so.2); the 128-byte unique id travels over the caller's own
 * rendezvous (torch.distributed / MPI / a file). */
int rxc_comm_unique_id(uint8_t id_out[128]);
int rxc_comm_init(int rank, int world, const uint8_t id[128]);
int rxc_comm_reduce_sum_f64(double *host_inout, uint64_t len, int root);
/* The sharded call of a multi-process run (one process per GPU): this rank's share of
 * the sources goes through the betweenness kernels, the partial vector STAYS in HBM, one in-place
 * ncclReduce(sum, f64) on the handle's stream combines the ranks over NVLink, and only `root` copies
 * the result to `out` (n_bound / e_bound doubles, original indexing, un-rescaled; ignored on the
 * other ranks, may be NULL there).  Replaces the reference's shared Vec behind an RwLock
 * (rustworkx-core/src/centrality.rs:107,280,317).  Every rank of rxc_comm_init must make the call;
 * without a communicator it is rxc_*_partial.  rxc_stats.ms_reduce times the collective. */
int rxc_betweenness_partial_reduce(rxc_graph *g, const uint32_t *sources, uint64_t n_sources,
                                   int include_endpoints, int root, double *out);
/* The order in which the betweenness calls group a source list (option bc_order): sources whose
 * searches have a similar level structure -- siblings of a hub first -- end up next to each other and
 * share a group and a DRAM sector.  Contributions are additive over sources, so the order never
 * changes results.  A multi-process run orders the whole list once and gives every rank a CONTIGUOUS
 * slice of it (rustworkx_b200/distributed.py), which keeps the clusters together on one rank (the
 * per-rank call orders its slice again and finds the same clusters, except one that a slice
 * boundary cut below the cluster size).  ordered[n_sources]: original node indices. */
int rxc_betweenness_source_order(rxc_graph *g, const uint32_t *sources, uint64_t n_sources,
                                 uint32_t *ordered);
int rxc_edge_betweenness_partial_reduce(rxc_graph *g, const uint32_t *sources, uint64_t n_sources,
                                        int root, double *out);
int rxc_comm_barrier(void);
void rxc_comm_destroy(void);

/* ---- instrumentation --------------------------------------------------------------------------- */
int rxc_get_stats(const rxc_graph *g, rxc_stats *out);
/* Tunables (tests and A/B measurements; the defaults are the measured optimum):
 *   bc_groups N        groups of 128 sources per betweenness batch (0: 16M / n_live, at most 64)
 *   bc_sigma 0|1|2     sigma rows: 0 uint32 path counts with f64 fall-back on overflow, 1 f64, 2 uint32
 *   bc_sparse_div N    forward levels with <= n/N queue entries use the frontier filter (0: off)
 *   hub_degree N       vertices above N arcs are handled by a whole CTA (512)
 *   bfs_mode 0|1|2|3   distance-only regime: 0 auto (pilot), 1 32-source words top-down,
 *                      2 CTA-per-source in shared memory, 3 128-source masks bottom-up
 *   bfs_depth_switch N auto: pilot BFS deeper than N levels -> regime 2 (48)
 *   bfs_ell 0|1        regime 2: one uint4 per vertex when no vertex has more than 4 arcs (1)
 *   bc_order 0|1|2     betweenness: which sources share a group and a DRAM sector.  0: caller's
 *                      order; 1: sorted by two-hop neighbourhood size (similar level structure);
 *                      2 (default): sources that hang off the same hub (at least 2 lanes' worth of
 *                      them in the call) form clusters that come first, the rest in order 1
 *   bc_siblings N      bc_order 2: a hub named by at least N sources of the call forms a cluster (2)
 *   bc_bwd_ctas N      dependency sweep: CTAs of one launch, all groups together (32768; 0: a warp per
 *                      1..8 queue entries)
 *   bc_look 0|1|2      dependency sweep, look-ahead kernel (a warp prefetches for its next three queue
 *                      entries): 0 never, 1 (default) on leaf-heavy levels -- the next level has less
 *                      than a quarter of this level's entries --, 2 on every level (node mode only)
 *   snapshot_cache_entries N  handles rxc_graph_snapshot_cached keeps alive (2)
 *   bfs_threads N      regime 2: threads per CTA (0: 128 / 256 / 512 / 1024 for n < 2^14 / 2^16 / 2^18 / above)
 *   bfs_groups N       regime 1: groups of 32 sources per batch (0: up to 4096)
 *   device_csr 0|1     build the CSR on the device (1) or with the host counting sort (0)
 *   mem_fraction_pct N share of free HBM the scratch may take (70)
 * RXC_TRACE=1 in the environment prints the phases of a snapshot and scratch allocations to stderr. */
int rxc_set_option(const char *name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* RXCUDA_H */
