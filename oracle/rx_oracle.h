/*
 * rx_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT THE PRODUCT).
 *
 * A plain-C restatement of the reference's all-sources unweighted BFS path
 * (rustworkx v0.18.0).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (rustworkx_b200 / librxcuda.so) never links, imports or calls it.
 *
 * Parity status: PINNED against every golden vector the reference's own tests
 * hold for this path; see tests/test_oracle_golden.py.
 * The reference itself (Rust) cannot be compiled in this image (no cargo /
 * rustc), so oracle/_ref does not exist; the oracle is a "port".
 *
 * Each function cites the reference lines it follows (paths relative to the
 * reference checkout).
 */
#ifndef RX_ORACLE_H
#define RX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RXO_END 0xFFFFFFFFu

typedef struct rxo_graph rxo_graph;

/* Work counters: TE = sum over sources of out-degrees of
 * reached vertices; TE_dag = arcs on shortest-path DAGs; V = reached vertices. */
typedef struct {
    uint64_t te;
    uint64_t te_dag;
    uint64_t v;
    uint64_t sources;
    uint64_t max_level;
} rxo_stats;

/* ---- petgraph-0.8 StableGraph-like container (adjacency = two intrusive
 *      linked lists per node, head insertion, index reuse via free lists) ---- */
rxo_graph *rxo_graph_new(int directed);
void rxo_graph_free(rxo_graph *g);
uint32_t rxo_add_node(rxo_graph *g);
uint32_t rxo_add_edge(rxo_graph *g, uint32_t a, uint32_t b);
int rxo_remove_edge(rxo_graph *g, uint32_t e);
int rxo_remove_node(rxo_graph *g, uint32_t a);
size_t rxo_node_count(const rxo_graph *g);
size_t rxo_edge_count(const rxo_graph *g);
size_t rxo_node_bound(const rxo_graph *g);
size_t rxo_edge_bound(const rxo_graph *g);
int rxo_is_directed(const rxo_graph *g);
int rxo_node_alive(const rxo_graph *g, uint32_t a);
int rxo_edge_alive(const rxo_graph *g, uint32_t e);
/* endpoints of a live edge; returns 0 on success */
int rxo_edge_endpoints(const rxo_graph *g, uint32_t e, uint32_t *a, uint32_t *b);

/* Rebuild a graph from a snapshot: slots [0,n_bound) with alive flags, edge
 * slots [0,e_bound) with endpoints/alive flags and an insertion sequence
 * number; live edges are linked in ascending seq order so that adjacency
 * iteration order equals petgraph's (most recently added edge first). */
rxo_graph *rxo_graph_from_snapshot(int directed, size_t n_bound, const uint8_t *node_alive,
                                   size_t e_bound, const uint32_t *esrc, const uint32_t *edst,
                                   const uint8_t *edge_alive, const uint64_t *eseq);

/* ---- the four hot-path functions.  Output arrays are dense over the index
 *      bound; `present[i]`=0 marks a hole (reference: None). ---- */

/* rustworkx-core/src/centrality.rs:70-132 */
void rxo_betweenness_centrality(const rxo_graph *g, int include_endpoints, int normalized,
                                size_t parallel_threshold, double *out, uint8_t *present,
                                rxo_stats *stats);
/* rustworkx-core/src/centrality.rs:174-219 */
void rxo_edge_betweenness_centrality(const rxo_graph *g, int normalized,
                                     size_t parallel_threshold, double *out, uint8_t *present,
                                     rxo_stats *stats);
/* rustworkx-core/src/centrality.rs:1166-1224 */
void rxo_closeness_centrality(const rxo_graph *g, int wf_improved, size_t parallel_threshold,
                              double *out, uint8_t *present, rxo_stats *stats);
/* rustworkx-core/src/shortest_path/distance_matrix.rs:71-159; out is
 * node_count x node_count, row-major, compacted live-node order. */
void rxo_distance_matrix(const rxo_graph *g, size_t parallel_threshold, int as_undirected,
                         double null_value, double *out, rxo_stats *stats);

/* src/shortest_path/average_length.rs:22-77 and src/shortest_path/mod.rs:1657-1740 */
void rxo_distance_sum(const rxo_graph *g, size_t parallel_threshold, int as_undirected,
                      uint64_t *sum_out, uint64_t *pairs_out);
double rxo_average_shortest_path_length(const rxo_graph *g, size_t parallel_threshold,
                                        int as_undirected, int disconnected);

/* ---- partial (source-subset) variants used for sampled parity on large
 *      graphs and as the timed CPU baseline: same per-source code, no final
 *      _rescale; accumulate into `out` (caller zero-initialises). ---- */
void rxo_betweenness_partial(const rxo_graph *g, const uint32_t *sources, size_t k,
                             int include_endpoints, int threads, double *out, rxo_stats *stats);
void rxo_edge_betweenness_partial(const rxo_graph *g, const uint32_t *sources, size_t k,
                                  int threads, double *out, rxo_stats *stats);
void rxo_closeness_partial(const rxo_graph *g, const uint32_t *sources, size_t k,
                           int wf_improved, int threads, double *out, rxo_stats *stats);

/* ---- group centralities.  `group` holds ORIGINAL node indices of live nodes
 *      (the binding validates them, src/centrality.rs:1216-1222); duplicates count once. ---- */
/* rustworkx-core/src/centrality.rs:1860-1916 */
double rxo_group_closeness_centrality(const rxo_graph *g, const uint32_t *group, size_t k);
/* rustworkx-core/src/centrality.rs:1963-2101 */
double rxo_group_betweenness_centrality(const rxo_graph *g, const uint32_t *group, size_t k,
                                        int normalized, size_t parallel_threshold);
/* the per-source sums of :1985-2080 over a subset of sources, before /2 and normalisation */
double rxo_group_betweenness_partial(const rxo_graph *g, const uint32_t *group, size_t k,
                                     const uint32_t *sources, size_t n_sources, int threads,
                                     rxo_stats *stats);

/* ---- weighted shortest path lengths.  Weights / costs are dense over the edge
 *      bound by ORIGINAL edge index; non-negative, not NaN. ---- */
/* rustworkx-core/src/centrality.rs:1297-1357 */
void rxo_newman_weighted_closeness_centrality(const rxo_graph *g, const double *weight_by_edge,
                                              int wf_improved, size_t parallel_threshold,
                                              double *out, uint8_t *present, rxo_stats *stats);
void rxo_weighted_closeness_partial(const rxo_graph *g, const double *weight_by_edge, int invert,
                                    int rev, const uint32_t *sources, size_t k, int wf_improved,
                                    int threads, double *out, rxo_stats *stats);
/* src/shortest_path/all_pairs_dijkstra.rs:36-105 + rustworkx-core/src/shortest_path/dijkstra.rs:100-160;
 * out[i * node_bound + v], NAN where the reference's map has no entry */
void rxo_dijkstra_lengths_rows(const rxo_graph *g, const double *cost_by_edge, const uint32_t *sources,
                               size_t k, int threads, double *out, rxo_stats *stats);

/* _rescale, rustworkx-core/src/centrality.rs:221-252, on a dense vector */
void rxo_rescale(double *x, size_t len, size_t node_count, int normalized, int directed,
                 int include_endpoints);

int rxo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
