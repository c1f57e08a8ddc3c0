/*
 * rx_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT THE PRODUCT).
 * See rx_oracle.h for the rules on who may use this file.
 *
 * Plain-C restatement of rustworkx v0.18.0's all-sources unweighted BFS path:
 *   rustworkx-core/src/centrality.rs:70-327,397-511,1166-1224
 *   rustworkx-core/src/shortest_path/distance_matrix.rs:71-159
 * over a restatement of petgraph 0.8.3's StableGraph adjacency (Cargo.lock;
 * not vendored in the reference tree -- behaviour restated from its public
 * documentation: intrusive per-node out/in edge lists with head insertion,
 * LIFO free lists for vacated node/edge slots).
 *
 * The per-source allocation pattern (predecessor vectors, sigma, distance,
 * delta sized node_bound) and the single global lock around accumulation are
 * kept on purpose: this file is also the timed "reference CPU path" (kind
 * "port") in bench.py.
 */
#include "rx_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* petgraph-like container                                             */
/* ------------------------------------------------------------------ */

typedef struct {
    uint32_t next[2]; /* heads of outgoing / incoming edge lists (or free-list links) */
    uint8_t alive;
} rxo_node_t;

typedef struct {
    uint32_t next[2]; /* next edge in source's outgoing list / target's incoming list */
    uint32_t node[2]; /* [source, target] */
    void *weight;     /* the reference stores a Py<PyAny> here; kept for struct size/alive */
} rxo_edge_t;

struct rxo_graph {
    rxo_node_t *nodes;
    size_t nodes_len, nodes_cap;
    rxo_edge_t *edges;
    size_t edges_len, edges_cap;
    size_t node_count, edge_count;
    uint32_t free_node, free_edge;
    int directed;
};

rxo_graph *rxo_graph_new(int directed) {
    rxo_graph *g = (rxo_graph *)calloc(1, sizeof(rxo_graph));
    g->free_node = RXO_END;
    g->free_edge = RXO_END;
    g->directed = directed ? 1 : 0;
    return g;
}

void rxo_graph_free(rxo_graph *g) {
    if (!g) return;
    free(g->nodes);
    free(g->edges);
    free(g);
}

size_t rxo_node_count(const rxo_graph *g) { return g->node_count; }
size_t rxo_edge_count(const rxo_graph *g) { return g->edge_count; }
int rxo_is_directed(const rxo_graph *g) { return g->directed; }
int rxo_node_alive(const rxo_graph *g, uint32_t a) { return a < g->nodes_len && g->nodes[a].alive; }
int rxo_edge_alive(const rxo_graph *g, uint32_t e) {
    return e < g->edges_len && g->edges[e].weight != NULL;
}
int rxo_edge_endpoints(const rxo_graph *g, uint32_t e, uint32_t *a, uint32_t *b) {
    if (!rxo_edge_alive(g, e)) return -1;
    *a = g->edges[e].node[0];
    *b = g->edges[e].node[1];
    return 0;
}

/* StableGraph::node_bound: highest live index + 1 */
size_t rxo_node_bound(const rxo_graph *g) {
    size_t i = g->nodes_len;
    while (i > 0 && !g->nodes[i - 1].alive) i--;
    return i;
}
size_t rxo_edge_bound(const rxo_graph *g) {
    size_t i = g->edges_len;
    while (i > 0 && g->edges[i - 1].weight == NULL) i--;
    return i;
}

static void *const RXO_SOME = (void *)(uintptr_t)1;

uint32_t rxo_add_node(rxo_graph *g) {
    if (g->free_node != RXO_END) {
        /* occupy_vacant_node: doubly linked free list through next[0]/next[1] */
        uint32_t idx = g->free_node;
        rxo_node_t *slot = &g->nodes[idx];
        uint32_t previous = slot->next[1], next = slot->next[0];
        slot->next[0] = slot->next[1] = RXO_END;
        slot->alive = 1;
        if (previous != RXO_END) g->nodes[previous].next[0] = next;
        if (next != RXO_END) g->nodes[next].next[1] = previous;
        if (g->free_node == idx) g->free_node = next;
        g->node_count++;
        return idx;
    }
    if (g->nodes_len == g->nodes_cap) {
        g->nodes_cap = g->nodes_cap ? g->nodes_cap * 2 : 16;
        g->nodes = (rxo_node_t *)realloc(g->nodes, g->nodes_cap * sizeof(rxo_node_t));
    }
    rxo_node_t *n = &g->nodes[g->nodes_len];
    n->next[0] = n->next[1] = RXO_END;
    n->alive = 1;
    g->node_count++;
    return (uint32_t)g->nodes_len++;
}

static uint32_t rxo_link_edge(rxo_graph *g, uint32_t e, uint32_t a, uint32_t b) {
    rxo_edge_t *edge = &g->edges[e];
    edge->node[0] = a;
    edge->node[1] = b;
    edge->weight = RXO_SOME;
    if (a == b) {
        rxo_node_t *an = &g->nodes[a];
        edge->next[0] = an->next[0];
        edge->next[1] = an->next[1];
        an->next[0] = e;
        an->next[1] = e;
    } else {
        rxo_node_t *an = &g->nodes[a], *bn = &g->nodes[b];
        edge->next[0] = an->next[0];
        edge->next[1] = bn->next[1];
        an->next[0] = e;
        bn->next[1] = e;
    }
    g->edge_count++;
    return e;
}

uint32_t rxo_add_edge(rxo_graph *g, uint32_t a, uint32_t b) {
    if (!rxo_node_alive(g, a) || !rxo_node_alive(g, b)) return RXO_END;
    uint32_t e;
    if (g->free_edge != RXO_END) {
        e = g->free_edge;
        g->free_edge = g->edges[e].next[0];
    } else {
        if (g->edges_len == g->edges_cap) {
            g->edges_cap = g->edges_cap ? g->edges_cap * 2 : 16;
            g->edges = (rxo_edge_t *)realloc(g->edges, g->edges_cap * sizeof(rxo_edge_t));
        }
        e = (uint32_t)g->edges_len++;
    }
    return rxo_link_edge(g, e, a, b);
}

int rxo_remove_edge(rxo_graph *g, uint32_t e) {
    if (!rxo_edge_alive(g, e)) return -1;
    rxo_edge_t ed = g->edges[e];
    /* change_edge_links */
    for (int k = 0; k < 2; k++) {
        rxo_node_t *node = &g->nodes[ed.node[k]];
        uint32_t fst = node->next[k];
        if (fst == e) {
            node->next[k] = ed.next[k];
        } else {
            uint32_t cur = fst;
            while (cur != RXO_END) {
                rxo_edge_t *ce = &g->edges[cur];
                if (ce->next[k] == e) {
                    ce->next[k] = ed.next[k];
                    break;
                }
                cur = ce->next[k];
            }
        }
    }
    rxo_edge_t *edge = &g->edges[e];
    edge->next[0] = g->free_edge;
    edge->next[1] = RXO_END;
    edge->node[0] = edge->node[1] = RXO_END;
    edge->weight = NULL;
    g->free_edge = e;
    g->edge_count--;
    return 0;
}

int rxo_remove_node(rxo_graph *g, uint32_t a) {
    if (!rxo_node_alive(g, a)) return -1;
    g->nodes[a].alive = 0;
    for (int k = 0; k < 2; k++) {
        for (;;) {
            uint32_t next = g->nodes[a].next[k];
            if (next == RXO_END) break;
            rxo_remove_edge(g, next);
        }
    }
    rxo_node_t *slot = &g->nodes[a];
    slot->next[0] = g->free_node;
    slot->next[1] = RXO_END;
    if (g->free_node != RXO_END) g->nodes[g->free_node].next[1] = a;
    g->free_node = a;
    g->node_count--;
    return 0;
}

typedef struct {
    uint64_t seq;
    uint32_t idx;
} rxo_seq_t;

static int rxo_seq_cmp(const void *x, const void *y) {
    const rxo_seq_t *a = (const rxo_seq_t *)x, *b = (const rxo_seq_t *)y;
    return (a->seq > b->seq) - (a->seq < b->seq);
}

rxo_graph *rxo_graph_from_snapshot(int directed, size_t n_bound, const uint8_t *node_alive,
                                   size_t e_bound, const uint32_t *esrc, const uint32_t *edst,
                                   const uint8_t *edge_alive, const uint64_t *eseq) {
    rxo_graph *g = rxo_graph_new(directed);
    g->nodes_cap = n_bound ? n_bound : 1;
    g->nodes = (rxo_node_t *)malloc(g->nodes_cap * sizeof(rxo_node_t));
    g->nodes_len = n_bound;
    for (size_t i = 0; i < n_bound; i++) {
        g->nodes[i].next[0] = g->nodes[i].next[1] = RXO_END;
        g->nodes[i].alive = node_alive ? (node_alive[i] != 0) : 1;
        if (g->nodes[i].alive) g->node_count++;
    }
    /* vacant node slots: chain them into the free list (ascending push => LIFO pops highest) */
    for (size_t i = 0; i < n_bound; i++) {
        if (!g->nodes[i].alive) {
            g->nodes[i].next[0] = g->free_node;
            g->nodes[i].next[1] = RXO_END;
            if (g->free_node != RXO_END) g->nodes[g->free_node].next[1] = (uint32_t)i;
            g->free_node = (uint32_t)i;
        }
    }
    g->edges_cap = e_bound ? e_bound : 1;
    g->edges = (rxo_edge_t *)malloc(g->edges_cap * sizeof(rxo_edge_t));
    g->edges_len = e_bound;
    size_t live = 0;
    for (size_t e = 0; e < e_bound; e++) {
        g->edges[e].next[0] = g->edges[e].next[1] = RXO_END;
        g->edges[e].node[0] = g->edges[e].node[1] = RXO_END;
        g->edges[e].weight = NULL;
        if (!edge_alive || edge_alive[e]) live++;
    }
    rxo_seq_t *order = (rxo_seq_t *)malloc((live ? live : 1) * sizeof(rxo_seq_t));
    size_t k = 0;
    for (size_t e = 0; e < e_bound; e++) {
        if (!edge_alive || edge_alive[e]) {
            order[k].seq = eseq ? eseq[e] : (uint64_t)e;
            order[k].idx = (uint32_t)e;
            k++;
        } else {
            g->edges[e].next[0] = g->free_edge;
            g->free_edge = (uint32_t)e;
        }
    }
    if (eseq) qsort(order, live, sizeof(rxo_seq_t), rxo_seq_cmp);
    for (size_t i = 0; i < live; i++) {
        uint32_t e = order[i].idx;
        rxo_link_edge(g, e, esrc[e], edst[e]);
    }
    free(order);
    return g;
}

/* ------------------------------------------------------------------ */
/* small helpers: Vec<usize>-like growable array, FIFO queue           */
/* ------------------------------------------------------------------ */

typedef struct {
    size_t *p;
    uint32_t len, cap;
} rxo_vec;

static inline void rxo_vec_push(rxo_vec *v, size_t x) {
    if (v->len == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 4; /* Rust's RawVec: first allocation holds 4 usizes */
        v->p = (size_t *)realloc(v->p, (size_t)v->cap * sizeof(size_t));
    }
    v->p[v->len++] = x;
}

int rxo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ */
/* betweenness: forward BFS                                            */
/* ------------------------------------------------------------------ */

typedef struct {
    uint32_t *verts_sorted_by_distance; /* reversed BFS order (a stack) */
    size_t n_verts;
    rxo_vec *predecessors;      /* per node index */
    rxo_vec *predecessor_edges; /* per node index (edge variant only) */
    double *sigma;
    uint64_t te;
    uint64_t max_level;
} rxo_spd;

#define RXO_NONE ((size_t)-1) /* Option<usize>::None */

/* shortest_path_for_centrality, centrality.rs:406-445, and
 * shortest_path_for_edge_centrality, centrality.rs:459-511 (with_edges=1).
 * Neighbour iteration = petgraph Neighbors/Edges: outgoing list first, then
 * (undirected only) the incoming list with self-loops skipped. */
static void rxo_shortest_paths(const rxo_graph *g, uint32_t s, int with_edges, rxo_spd *r) {
    size_t c = g->node_count;
    size_t max_index = rxo_node_bound(g);
    r->verts_sorted_by_distance = (uint32_t *)malloc((c ? c : 1) * sizeof(uint32_t));
    r->n_verts = 0;
    r->predecessors = (rxo_vec *)calloc(max_index ? max_index : 1, sizeof(rxo_vec));
    r->predecessor_edges =
        with_edges ? (rxo_vec *)calloc(max_index ? max_index : 1, sizeof(rxo_vec)) : NULL;
    r->sigma = (double *)calloc(max_index ? max_index : 1, sizeof(double));
    size_t *distance = (size_t *)malloc((max_index ? max_index : 1) * sizeof(size_t));
    for (size_t i = 0; i < max_index; i++) distance[i] = RXO_NONE;
    uint32_t *Q = (uint32_t *)malloc((c ? c : 1) * sizeof(uint32_t));
    size_t qh = 0, qt = 0;
    uint64_t te = 0, maxl = 0;

    r->sigma[s] = 1.0;
    distance[s] = 0;
    Q[qt++] = s;
    while (qh < qt) {
        uint32_t v = Q[qh++];
        r->verts_sorted_by_distance[r->n_verts++] = v;
        size_t distance_v = distance[v];
        if (distance_v > maxl) maxl = distance_v;
        const rxo_node_t *vn = &g->nodes[v];
        /* outgoing list */
        for (uint32_t e = vn->next[0]; e != RXO_END;) {
            const rxo_edge_t *ed = &g->edges[e];
            uint32_t w = ed->node[1];
            te++;
            if (distance[w] == RXO_NONE) {
                Q[qt++] = w;
                distance[w] = distance_v + 1;
            }
            if (distance[w] == distance_v + 1) {
                r->sigma[w] += r->sigma[v];
                rxo_vec_push(&r->predecessors[w], v);
                if (with_edges) rxo_vec_push(&r->predecessor_edges[w], e);
            }
            e = ed->next[0];
        }
        if (!g->directed) {
            /* incoming list, skipping self-loops already seen above */
            for (uint32_t e = vn->next[1]; e != RXO_END;) {
                const rxo_edge_t *ed = &g->edges[e];
                uint32_t w = ed->node[0];
                uint32_t enext = ed->next[1];
                if (w != v) {
                    te++;
                    if (distance[w] == RXO_NONE) {
                        Q[qt++] = w;
                        distance[w] = distance_v + 1;
                    }
                    if (distance[w] == distance_v + 1) {
                        r->sigma[w] += r->sigma[v];
                        rxo_vec_push(&r->predecessors[w], v);
                        if (with_edges) rxo_vec_push(&r->predecessor_edges[w], e);
                    }
                }
                e = enext;
            }
        }
    }
    /* verts_sorted_by_distance.reverse() */
    if (r->n_verts > 1) {
        for (size_t i = 0, j = r->n_verts - 1; i < j; i++, j--) {
            uint32_t t = r->verts_sorted_by_distance[i];
            r->verts_sorted_by_distance[i] = r->verts_sorted_by_distance[j];
            r->verts_sorted_by_distance[j] = t;
        }
    }
    r->te = te;
    r->max_level = maxl;
    free(distance);
    free(Q);
}

static void rxo_spd_free(rxo_spd *r, size_t max_index) {
    for (size_t i = 0; i < max_index; i++) {
        free(r->predecessors[i].p);
        if (r->predecessor_edges) free(r->predecessor_edges[i].p);
    }
    free(r->predecessors);
    free(r->predecessor_edges);
    free(r->sigma);
    free(r->verts_sorted_by_distance);
}

/* _accumulate_vertices, centrality.rs:254-299.  The write lock of :280 is the
 * omp critical section; `betweenness` holds values for live slots only. */
static uint64_t rxo_accumulate_vertices(double *betweenness, size_t max_index, rxo_spd *pc,
                                        uint32_t node_s, int include_endpoints) {
    double *delta = (double *)calloc(max_index ? max_index : 1, sizeof(double));
    uint64_t dag = 0;
    for (size_t i = 0; i < pc->n_verts; i++) {
        uint32_t iw = pc->verts_sorted_by_distance[i];
        double coeff = (1.0 + delta[iw]) / pc->sigma[iw];
        const rxo_vec *p_w = &pc->predecessors[iw];
        dag += p_w->len;
        for (uint32_t j = 0; j < p_w->len; j++) {
            size_t iv = p_w->p[j];
            delta[iv] += pc->sigma[iv] * coeff;
        }
    }
#pragma omp critical(rxo_betweenness_lock)
    {
        if (include_endpoints) {
            betweenness[node_s] = betweenness[node_s] + (double)(pc->n_verts - 1);
            for (size_t i = 0; i < pc->n_verts; i++) {
                uint32_t iw = pc->verts_sorted_by_distance[i];
                if (iw != node_s) betweenness[iw] = betweenness[iw] + delta[iw] + 1.0;
            }
        } else {
            for (size_t i = 0; i < pc->n_verts; i++) {
                uint32_t iw = pc->verts_sorted_by_distance[i];
                if (iw != node_s) betweenness[iw] = betweenness[iw] + delta[iw];
            }
        }
    }
    free(delta);
    return dag;
}

/* accumulate_edges, centrality.rs:301-327 (lock taken per w, :317) */
static uint64_t rxo_accumulate_edges(double *betweenness, size_t max_index, rxo_spd *pc) {
    double *delta = (double *)calloc(max_index ? max_index : 1, sizeof(double));
    uint64_t dag = 0;
    for (size_t i = 0; i < pc->n_verts; i++) {
        uint32_t iw = pc->verts_sorted_by_distance[i];
        double coeff = (1.0 + delta[iw]) / pc->sigma[iw];
        const rxo_vec *p_w = &pc->predecessors[iw];
        const rxo_vec *e_w = &pc->predecessor_edges[iw];
        dag += p_w->len;
#pragma omp critical(rxo_betweenness_lock)
        {
            for (uint32_t j = 0; j < p_w->len; j++) {
                size_t iv = p_w->p[j];
                size_t ie = e_w->p[j];
                double c = pc->sigma[iv] * coeff;
                betweenness[ie] = betweenness[ie] + c;
                delta[iv] += c;
            }
        }
    }
    free(delta);
    return dag;
}

/* _rescale, centrality.rs:221-252: multiply by a precomputed reciprocal */
void rxo_rescale(double *x, size_t len, size_t node_count, int normalized, int directed,
                 int include_endpoints) {
    int do_scale = 1;
    double scale = 1.0;
    if (normalized) {
        if (include_endpoints) {
            if (node_count < 2) do_scale = 0;
            else scale = 1.0 / (double)(node_count * (node_count - 1));
        } else if (node_count <= 2) {
            do_scale = 0;
        } else {
            scale = 1.0 / (double)((node_count - 1) * (node_count - 2));
        }
    } else if (!directed) {
        scale = 0.5;
    } else {
        do_scale = 0;
    }
    if (do_scale)
        for (size_t i = 0; i < len; i++) x[i] = x[i] * scale;
}

static void rxo_stats_add(rxo_stats *st, uint64_t te, uint64_t dag, uint64_t v, uint64_t maxl) {
    if (!st) return;
#pragma omp critical(rxo_stats_lock)
    {
        st->te += te;
        st->te_dag += dag;
        st->v += v;
        st->sources += 1;
        if (maxl > st->max_level) st->max_level = maxl;
    }
}

static void rxo_bc_sources(const rxo_graph *g, const uint32_t *sources, size_t k, int edge_mode,
                           int include_endpoints, int threads, double *acc, rxo_stats *st) {
    size_t max_index = rxo_node_bound(g);
    long kk = (long)k;
    (void)threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1)
    for (long i = 0; i < kk; i++) {
        uint32_t s = sources[i];
        rxo_spd pc;
        rxo_shortest_paths(g, s, edge_mode, &pc);
        uint64_t dag;
        if (edge_mode) dag = rxo_accumulate_edges(acc, max_index, &pc);
        else dag = rxo_accumulate_vertices(acc, max_index, &pc, s, include_endpoints);
        rxo_stats_add(st, pc.te, dag, pc.n_verts, pc.max_level);
        rxo_spd_free(&pc, max_index);
    }
}

static uint32_t *rxo_live_nodes(const rxo_graph *g, size_t *count) {
    uint32_t *ids = (uint32_t *)malloc((g->node_count ? g->node_count : 1) * sizeof(uint32_t));
    size_t k = 0;
    for (size_t i = 0; i < g->nodes_len; i++)
        if (g->nodes[i].alive) ids[k++] = (uint32_t)i;
    *count = k;
    return ids;
}

/* betweenness_centrality, centrality.rs:70-132 */
void rxo_betweenness_centrality(const rxo_graph *g, int include_endpoints, int normalized,
                                size_t parallel_threshold, double *out, uint8_t *present,
                                rxo_stats *stats) {
    size_t max_index = rxo_node_bound(g);
    if (stats) memset(stats, 0, sizeof(*stats));
    for (size_t i = 0; i < max_index; i++) {
        out[i] = 0.0;
        present[i] = g->nodes[i].alive;
    }
    size_t k;
    uint32_t *ids = rxo_live_nodes(g, &k);
    int threads = (g->node_count >= parallel_threshold) ? rxo_max_threads() : 1;
    rxo_bc_sources(g, ids, k, 0, include_endpoints, threads, out, stats);
    free(ids);
    rxo_rescale(out, max_index, g->node_count, normalized, g->directed, include_endpoints);
}

/* edge_betweenness_centrality, centrality.rs:174-219 */
void rxo_edge_betweenness_centrality(const rxo_graph *g, int normalized,
                                     size_t parallel_threshold, double *out, uint8_t *present,
                                     rxo_stats *stats) {
    size_t e_bound = rxo_edge_bound(g);
    if (stats) memset(stats, 0, sizeof(*stats));
    for (size_t i = 0; i < e_bound; i++) {
        out[i] = 0.0;
        present[i] = g->edges[i].weight != NULL;
    }
    size_t k;
    uint32_t *ids = rxo_live_nodes(g, &k);
    int threads = (g->node_count >= parallel_threshold) ? rxo_max_threads() : 1;
    rxo_bc_sources(g, ids, k, 1, 0, threads, out, stats);
    free(ids);
    rxo_rescale(out, e_bound, g->node_count, normalized, g->directed, 1);
}

void rxo_betweenness_partial(const rxo_graph *g, const uint32_t *sources, size_t k,
                             int include_endpoints, int threads, double *out, rxo_stats *stats) {
    if (stats) memset(stats, 0, sizeof(*stats));
    rxo_bc_sources(g, sources, k, 0, include_endpoints, threads, out, stats);
}

void rxo_edge_betweenness_partial(const rxo_graph *g, const uint32_t *sources, size_t k,
                                  int threads, double *out, rxo_stats *stats) {
    if (stats) memset(stats, 0, sizeof(*stats));
    rxo_bc_sources(g, sources, k, 1, 0, threads, out, stats);
}

/* ------------------------------------------------------------------ */
/* closeness                                                           */
/* ------------------------------------------------------------------ */

/* The closure at centrality.rs:1190-1202: petgraph Bfs over Reversed(&graph)
 * (incoming edges; all incident edges when undirected).  The reference keeps a
 * HashMap<NodeId,usize>; only its len() and the sum of its values are used
 * (:1209-1210), so this restatement keeps a dense distance array. */
static void rxo_closeness_one(const rxo_graph *g, uint32_t s, size_t max_index, size_t *reach,
                              size_t *dsum, uint64_t *te_out, uint64_t *maxl_out) {
    size_t *dist = (size_t *)malloc((max_index ? max_index : 1) * sizeof(size_t));
    for (size_t i = 0; i < max_index; i++) dist[i] = RXO_NONE;
    uint32_t *Q = (uint32_t *)malloc((g->node_count ? g->node_count : 1) * sizeof(uint32_t));
    size_t qh = 0, qt = 0, reachable = 0, sum = 0;
    uint64_t te = 0, maxl = 0;
    dist[s] = 0;
    Q[qt++] = s;
    while (qh < qt) {
        uint32_t v = Q[qh++];
        size_t d = dist[v];
        reachable++;
        sum += d;
        if (d > maxl) maxl = d;
        const rxo_node_t *vn = &g->nodes[v];
        if (g->directed) {
            for (uint32_t e = vn->next[1]; e != RXO_END; e = g->edges[e].next[1]) {
                uint32_t w = g->edges[e].node[0];
                te++;
                if (dist[w] == RXO_NONE) {
                    dist[w] = d + 1;
                    Q[qt++] = w;
                }
            }
        } else {
            for (uint32_t e = vn->next[0]; e != RXO_END; e = g->edges[e].next[0]) {
                uint32_t w = g->edges[e].node[1];
                te++;
                if (dist[w] == RXO_NONE) {
                    dist[w] = d + 1;
                    Q[qt++] = w;
                }
            }
            for (uint32_t e = vn->next[1]; e != RXO_END; e = g->edges[e].next[1]) {
                uint32_t w = g->edges[e].node[0];
                if (w == v) continue;
                te++;
                if (dist[w] == RXO_NONE) {
                    dist[w] = d + 1;
                    Q[qt++] = w;
                }
            }
        }
    }
    free(dist);
    free(Q);
    *reach = reachable;
    *dsum = sum;
    *te_out = te;
    *maxl_out = maxl;
}

/* centrality.rs:1209-1220: op order is (c * (r-1)) / (N-1), left to right */
static double rxo_closeness_value(size_t reachable, size_t dists_sum, int wf_improved,
                                  size_t node_count) {
    if (reachable == 1) return 0.0;
    double c = (double)(reachable - 1) / (double)dists_sum;
    if (wf_improved) c = c * (double)(reachable - 1) / (double)(node_count - 1);
    return c;
}

void rxo_closeness_partial(const rxo_graph *g, const uint32_t *sources, size_t k,
                           int wf_improved, int threads, double *out, rxo_stats *stats) {
    size_t max_index = rxo_node_bound(g);
    if (stats) memset(stats, 0, sizeof(*stats));
    long kk = (long)k;
    (void)threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1)
    for (long i = 0; i < kk; i++) {
        uint32_t s = sources[i];
        size_t reach, dsum;
        uint64_t te, maxl;
        rxo_closeness_one(g, s, max_index, &reach, &dsum, &te, &maxl);
        out[s] = rxo_closeness_value(reach, dsum, wf_improved, g->node_count);
        rxo_stats_add(stats, te, 0, reach, maxl);
    }
}

/* closeness_centrality, centrality.rs:1166-1224 */
void rxo_closeness_centrality(const rxo_graph *g, int wf_improved, size_t parallel_threshold,
                              double *out, uint8_t *present, rxo_stats *stats) {
    size_t max_index = rxo_node_bound(g);
    for (size_t i = 0; i < max_index; i++) {
        out[i] = 0.0;
        present[i] = g->nodes[i].alive;
    }
    size_t k;
    uint32_t *ids = rxo_live_nodes(g, &k);
    int threads = (g->node_count >= parallel_threshold) ? rxo_max_threads() : 1;
    rxo_closeness_partial(g, ids, k, wf_improved, threads, out, stats);
    free(ids);
}

/* compute_distance_sum, src/shortest_path/average_length.rs:22-77: per source a level-by-level
 * BFS over outgoing (and, for a directed graph with as_undirected, incoming) neighbours; returns
 * (sum of distances, connected ordered pairs). */
void rxo_distance_sum(const rxo_graph *g, size_t parallel_threshold, int as_undirected,
                      uint64_t *sum_out, uint64_t *pairs_out) {
    size_t bound = rxo_node_bound(g);
    size_t k;
    uint32_t *ids = rxo_live_nodes(g, &k);
    int threads = (g->node_count >= parallel_threshold) ? rxo_max_threads() : 1;
    uint64_t tot_sum = 0, tot_pairs = 0;
    long kk = (long)k;
    (void)threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1) \
    reduction(+ : tot_sum, tot_pairs)
    for (long i = 0; i < kk; i++) {
        size_t *dist = (size_t *)malloc((bound ? bound : 1) * sizeof(size_t));
        for (size_t j = 0; j < bound; j++) dist[j] = RXO_NONE;
        uint32_t *Q = (uint32_t *)malloc((g->node_count ? g->node_count : 1) * sizeof(uint32_t));
        size_t qh = 0, qt = 0;
        uint64_t count = 0, found = 0;
        dist[ids[i]] = 0;
        Q[qt++] = ids[i];
        while (qh < qt) {
            uint32_t v = Q[qh++];
            size_t d = dist[v];
            count += d;
            found++;
            const rxo_node_t *vn = &g->nodes[v];
            for (uint32_t e = vn->next[0]; e != RXO_END; e = g->edges[e].next[0]) {
                uint32_t w = g->edges[e].node[1];
                if (dist[w] == RXO_NONE) { dist[w] = d + 1; Q[qt++] = w; }
            }
            if (!g->directed || as_undirected) {
                for (uint32_t e = vn->next[1]; e != RXO_END; e = g->edges[e].next[1]) {
                    uint32_t w = g->edges[e].node[0];
                    if (dist[w] == RXO_NONE) { dist[w] = d + 1; Q[qt++] = w; }
                }
            }
        }
        tot_sum += count;
        tot_pairs += found - 1;
        free(dist);
        free(Q);
    }
    free(ids);
    *sum_out = tot_sum;
    *pairs_out = tot_pairs;
}

/* graph_/digraph_unweighted_average_shortest_path_length, src/shortest_path/mod.rs:1657-1740 */
double rxo_average_shortest_path_length(const rxo_graph *g, size_t parallel_threshold,
                                        int as_undirected, int disconnected) {
    size_t n = g->node_count;
    if (n <= 1) return NAN;
    uint64_t sum, conn_pairs;
    rxo_distance_sum(g, parallel_threshold, as_undirected, &sum, &conn_pairs);
    uint64_t tot_pairs = (uint64_t)n * (n - 1);
    if (disconnected && conn_pairs == 0) return NAN;
    if (!disconnected && conn_pairs < tot_pairs) return INFINITY;
    return (double)sum / (double)conn_pairs;
}

/* ------------------------------------------------------------------ */
/* distance_matrix                                                     */
/* ------------------------------------------------------------------ */

typedef struct {
    uint64_t *w;
    size_t nwords;
} rxo_bitset;

static inline void rxo_bs_put(uint64_t *w, size_t i) { w[i >> 6] |= (uint64_t)1 << (i & 63); }

/* distance_matrix, distance_matrix.rs:71-159: per-node FixedBitSet adjacency
 * rows over compacted live-node positions, level-synchronous bitset BFS. */
void rxo_distance_matrix(const rxo_graph *g, size_t parallel_threshold, int as_undirected,
                         double null_value, double *out, rxo_stats *stats) {
    size_t n = g->node_count;
    size_t bound = rxo_node_bound(g);
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n == 0) return;
    /* node_map / node_map_inv (:82-105): rank among live indices */
    size_t *node_map = (size_t *)malloc((bound ? bound : 1) * sizeof(size_t));
    uint32_t *node_map_inv = (uint32_t *)malloc(n * sizeof(uint32_t));
    size_t k = 0;
    for (size_t i = 0; i < bound; i++) {
        node_map[i] = RXO_NONE;
        if (g->nodes[i].alive) {
            node_map[i] = k;
            node_map_inv[k++] = (uint32_t)i;
        }
    }
    for (size_t i = 0; i < n * n; i++) out[i] = null_value; /* :106 */
    size_t nwords = (n + 63) / 64;
    uint64_t *neighbors = (uint64_t *)calloc(n * nwords, sizeof(uint64_t)); /* :107-126 */
    uint64_t arcs = 0;
    for (size_t idx = 0; idx < n; idx++) {
        uint32_t v = node_map_inv[idx];
        uint64_t *row = neighbors + idx * nwords;
        const rxo_node_t *vn = &g->nodes[v];
        if (g->directed) {
            if (as_undirected)
                for (uint32_t e = vn->next[1]; e != RXO_END; e = g->edges[e].next[1])
                    rxo_bs_put(row, node_map[g->edges[e].node[0]]);
            for (uint32_t e = vn->next[0]; e != RXO_END; e = g->edges[e].next[0])
                rxo_bs_put(row, node_map[g->edges[e].node[1]]);
        } else {
            for (uint32_t e = vn->next[0]; e != RXO_END; e = g->edges[e].next[0])
                rxo_bs_put(row, node_map[g->edges[e].node[1]]);
            for (uint32_t e = vn->next[1]; e != RXO_END; e = g->edges[e].next[1])
                rxo_bs_put(row, node_map[g->edges[e].node[0]]);
        }
        for (size_t wi = 0; wi < nwords; wi++) arcs += (uint64_t)__builtin_popcountll(row[wi]);
    }
    (void)arcs;
    int threads = (n >= parallel_threshold) ? rxo_max_threads() : 1;
    uint64_t tot_v = 0, tot_te = 0, tot_maxl = 0;
    long nn = (long)n;
    (void)threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1) \
    reduction(+ : tot_v, tot_te) reduction(max : tot_maxl)
    for (long start = 0; start < nn; start++) {
        /* bfs_traversal, :127-144 */
        double *row = out + (size_t)start * n;
        double distance = 0.0;
        uint64_t *seen = (uint64_t *)calloc(nwords, sizeof(uint64_t));
        uint64_t *next = (uint64_t *)calloc(nwords, sizeof(uint64_t));
        uint64_t *cur = (uint64_t *)calloc(nwords, sizeof(uint64_t));
        rxo_bs_put(cur, (size_t)start);
        uint64_t lvl = 0;
        for (;;) {
            int clear = 1;
            for (size_t wi = 0; wi < nwords; wi++)
                if (cur[wi]) {
                    clear = 0;
                    break;
                }
            if (clear) break;
            memset(next, 0, nwords * sizeof(uint64_t));
            for (size_t wi = 0; wi < nwords; wi++) {
                uint64_t bits = cur[wi];
                while (bits) {
                    size_t found = wi * 64 + (size_t)__builtin_ctzll(bits);
                    bits &= bits - 1;
                    row[found] = distance;
                    const uint64_t *nb = neighbors + found * nwords;
                    for (size_t wj = 0; wj < nwords; wj++) {
                        next[wj] |= nb[wj];
                        tot_te += (uint64_t)__builtin_popcountll(nb[wj]);
                    }
                    tot_v++;
                }
            }
            for (size_t wi = 0; wi < nwords; wi++) {
                seen[wi] |= cur[wi];
                next[wi] &= ~seen[wi];
            }
            if (lvl > tot_maxl) tot_maxl = lvl;
            lvl++;
            distance += 1.0;
            uint64_t *t = cur;
            cur = next;
            next = t;
        }
        free(seen);
        free(next);
        free(cur);
    }
    if (stats) {
        stats->v = tot_v;
        stats->te = tot_te;
        stats->sources = n;
        stats->max_level = tot_maxl;
    }
    free(neighbors);
    free(node_map);
    free(node_map_inv);
}

/* ------------------------------------------------------------------ */
/* group centralities                                   */
/* ------------------------------------------------------------------ */

/* Adjacency cursor.  rev=0: graph.neighbors(v) -- the outgoing list, then (undirected only) the
 * incoming list with self-loops skipped.  rev=1: Reversed(&graph).edges(v) -- the incoming list of a
 * directed graph; every incident edge of an undirected one. */
typedef struct {
    const rxo_graph *g;
    uint32_t v, e;
    uint32_t last_e; /* edge index of the arc returned last */
    int phase; /* 0: first list, 1: second list, 2: done */
    int rev;
} rxo_cursor;

static void rxo_cursor_init(rxo_cursor *c, const rxo_graph *g, uint32_t v, int rev) {
    c->g = g;
    c->v = v;
    c->rev = rev;
    if (g->directed) {
        c->phase = 1; /* a single list */
        c->e = g->nodes[v].next[rev ? 1 : 0];
    } else {
        c->phase = 0;
        c->e = g->nodes[v].next[0];
    }
}

static int rxo_cursor_next(rxo_cursor *c, uint32_t *w) {
    const rxo_graph *g = c->g;
    for (;;) {
        if (c->e == RXO_END) {
            if (c->phase == 0) {
                c->phase = 1;
                c->e = g->nodes[c->v].next[1];
                continue;
            }
            return 0;
        }
        const rxo_edge_t *ed = &g->edges[c->e];
        c->last_e = c->e;
        if (g->directed) {
            const int k = c->rev ? 1 : 0;
            *w = ed->node[1 - k];
            c->e = ed->next[k];
            return 1;
        }
        if (c->phase == 0) {
            *w = ed->node[1];
            c->e = ed->next[0];
            return 1;
        }
        c->e = ed->next[1];
        if (ed->node[0] == c->v) continue; /* self-loop: already seen in the outgoing list */
        *w = ed->node[0];
        return 1;
    }
}

/* group_set: HashSet<usize> of the group, as a dense flag array; returns its size */
static size_t rxo_group_flags(const rxo_graph *g, const uint32_t *group, size_t k, uint8_t **flags) {
    size_t max_index = rxo_node_bound(g);
    size_t top = max_index;
    for (size_t i = 0; i < k; i++)
        if ((size_t)group[i] + 1 > top) top = (size_t)group[i] + 1;
    uint8_t *f = (uint8_t *)calloc(top ? top : 1, 1);
    size_t sz = 0;
    for (size_t i = 0; i < k; i++)
        if (!f[group[i]]) {
            f[group[i]] = 1;
            sz++;
        }
    *flags = f;
    return sz;
}

/* group_closeness_centrality, centrality.rs:1860-1916: one multi-source BFS over Reversed(&graph) */
double rxo_group_closeness_centrality(const rxo_graph *g, const uint32_t *group, size_t k) {
    size_t node_count = g->node_count;
    uint8_t *in_group;
    size_t group_size = rxo_group_flags(g, group, k, &in_group);
    if (group_size >= node_count) {
        free(in_group);
        return 0.0;
    }
    size_t max_index = rxo_node_bound(g);
    size_t *distance = (size_t *)malloc((max_index ? max_index : 1) * sizeof(size_t));
    for (size_t i = 0; i < max_index; i++) distance[i] = RXO_NONE;
    uint32_t *Q = (uint32_t *)malloc((node_count ? node_count : 1) * sizeof(uint32_t));
    size_t qh = 0, qt = 0;
    for (size_t i = 0; i < max_index; i++) /* iteration order of the set does not affect distances */
        if (in_group[i]) {
            distance[i] = 0;
            Q[qt++] = (uint32_t)i;
        }
    while (qh < qt) {
        uint32_t v = Q[qh++];
        size_t dist_v = distance[v];
        rxo_cursor c;
        uint32_t w;
        rxo_cursor_init(&c, g, v, 1);
        while (rxo_cursor_next(&c, &w)) {
            if (distance[w] == RXO_NONE) {
                distance[w] = dist_v + 1;
                Q[qt++] = w;
            }
        }
    }
    size_t dist_sum = 0;
    for (size_t i = 0; i < max_index; i++) {
        if (!g->nodes[i].alive || in_group[i]) continue;
        if (distance[i] != RXO_NONE) dist_sum += distance[i];
    }
    free(distance);
    free(Q);
    free(in_group);
    if (dist_sum == 0) return 0.0;
    return (double)(node_count - group_size) / (double)dist_sum;
}

/* BFS with path counts from s; `blocked` (may be NULL) removes vertices (centrality.rs:1994-2044) */
static uint64_t rxo_group_bfs(const rxo_graph *g, uint32_t s, const uint8_t *blocked, double *sigma,
                              size_t *dist, uint32_t *Q, size_t max_index, uint64_t *reached) {
    for (size_t i = 0; i < max_index; i++) {
        sigma[i] = 0.0;
        dist[i] = RXO_NONE;
    }
    size_t qh = 0, qt = 0;
    uint64_t te = 0;
    sigma[s] = 1.0;
    dist[s] = 0;
    Q[qt++] = s;
    while (qh < qt) {
        uint32_t v = Q[qh++];
        size_t dist_v = dist[v];
        rxo_cursor c;
        uint32_t w;
        rxo_cursor_init(&c, g, v, 0);
        while (rxo_cursor_next(&c, &w)) {
            te++;
            if (blocked && blocked[w]) continue;
            if (dist[w] == RXO_NONE) {
                dist[w] = dist_v + 1;
                Q[qt++] = w;
            }
            if (dist[w] == dist_v + 1) sigma[w] += sigma[v];
        }
    }
    *reached = qt;
    return te;
}

/* Sum over the listed non-group sources of sum_t (sigma_full - paths_avoiding)/sigma_full
 * (centrality.rs:1985-2080), before the /2 and the normalisation. */
double rxo_group_betweenness_partial(const rxo_graph *g, const uint32_t *group, size_t k,
                                     const uint32_t *sources, size_t n_sources, int threads,
                                     rxo_stats *stats) {
    size_t max_index = rxo_node_bound(g);
    uint8_t *in_group;
    (void)rxo_group_flags(g, group, k, &in_group);
    if (stats) memset(stats, 0, sizeof(*stats));
    double total = 0.0;
    long kk = (long)n_sources;
    (void)threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1) reduction(+ : total)
    for (long i = 0; i < kk; i++) {
        uint32_t s = sources[i];
        if (in_group[s]) continue;
        size_t m = max_index ? max_index : 1;
        double *sigma_full = (double *)malloc(m * sizeof(double));
        double *sigma_no_group = (double *)malloc(m * sizeof(double));
        size_t *dist_full = (size_t *)malloc(m * sizeof(size_t));
        size_t *dist_no_group = (size_t *)malloc(m * sizeof(size_t));
        uint32_t *Q = (uint32_t *)malloc((g->node_count ? g->node_count : 1) * sizeof(uint32_t));
        uint64_t r1, r2;
        uint64_t te = rxo_group_bfs(g, s, NULL, sigma_full, dist_full, Q, max_index, &r1);
        te += rxo_group_bfs(g, s, in_group, sigma_no_group, dist_no_group, Q, max_index, &r2);
        double source_group_betweenness = 0.0;
        for (size_t t = 0; t < max_index; t++) {
            if (!g->nodes[t].alive || t == s || in_group[t]) continue;
            if (sigma_full[t] == 0.0) continue;
            double paths_avoiding = (dist_no_group[t] == dist_full[t]) ? sigma_no_group[t] : 0.0;
            source_group_betweenness += (sigma_full[t] - paths_avoiding) / sigma_full[t];
        }
        total += source_group_betweenness;
        rxo_stats_add(stats, te, 0, r1 + r2, 0);
        free(sigma_full);
        free(sigma_no_group);
        free(dist_full);
        free(dist_no_group);
        free(Q);
    }
    free(in_group);
    return total;
}

/* group_betweenness_centrality, centrality.rs:1963-2101 */
double rxo_group_betweenness_centrality(const rxo_graph *g, const uint32_t *group, size_t k,
                                        int normalized, size_t parallel_threshold) {
    size_t node_count = g->node_count;
    uint8_t *in_group;
    size_t group_size = rxo_group_flags(g, group, k, &in_group);
    free(in_group);
    if (group_size == 0 || node_count <= 1) return 0.0;
    size_t n;
    uint32_t *ids = rxo_live_nodes(g, &n);
    int threads = (node_count >= parallel_threshold) ? rxo_max_threads() : 1;
    double gb = rxo_group_betweenness_partial(g, group, k, ids, n, threads, NULL);
    free(ids);
    if (!g->directed) gb /= 2.0;
    if (normalized) {
        size_t non_group = node_count - group_size;
        if (non_group > 1) {
            double norm = g->directed ? (double)(non_group * (non_group - 1))
                                      : (double)((non_group * (non_group - 1)) / 2);
            gb /= norm;
        }
    }
    return gb;
}


/* ------------------------------------------------------------------ */
/* weighted shortest path lengths                       */
/* ------------------------------------------------------------------ */

typedef struct {
    double score;
    uint32_t node;
} rxo_scored;

/* BinaryHeap<MinScored<f64, NodeId>>: a binary min-heap on the score (ties do not affect the
 * distances) */
static void rxo_heap_push(rxo_scored **h, size_t *len, size_t *cap, rxo_scored x) {
    if (*len == *cap) {
        *cap = *cap ? *cap * 2 : 64;
        *h = (rxo_scored *)realloc(*h, *cap * sizeof(rxo_scored));
    }
    size_t i = (*len)++;
    while (i > 0) {
        size_t parent = (i - 1) / 2;
        if ((*h)[parent].score <= x.score) break;
        (*h)[i] = (*h)[parent];
        i = parent;
    }
    (*h)[i] = x;
}

static rxo_scored rxo_heap_pop(rxo_scored *h, size_t *len) {
    rxo_scored top = h[0];
    rxo_scored last = h[--(*len)];
    size_t i = 0, n = *len;
    for (;;) {
        size_t c = 2 * i + 1;
        if (c >= n) break;
        if (c + 1 < n && h[c + 1].score < h[c].score) c++;
        if (last.score <= h[c].score) break;
        h[i] = h[c];
        i = c;
    }
    if (n) h[i] = last;
    return top;
}

/* dijkstra without a goal: rustworkx-core/src/shortest_path/dijkstra.rs:100-160 (all_pairs) and
 * petgraph 0.8.3 algo::dijkstra (Newman; same loop, published algorithm -- petgraph is not in the
 * tree).  cost[e] by edge index; rev: Reversed(&graph).  score[i] = NAN marks "not in the map". */
static uint64_t rxo_dijkstra(const rxo_graph *g, uint32_t start, const double *cost, int rev,
                             double *score, uint8_t *visited, size_t max_index) {
    for (size_t i = 0; i < max_index; i++) {
        score[i] = NAN;
        visited[i] = 0;
    }
    rxo_scored *heap = NULL;
    size_t len = 0, cap = 0;
    uint64_t te = 0;
    score[start] = 0.0;
    rxo_heap_push(&heap, &len, &cap, (rxo_scored){0.0, start});
    while (len) {
        rxo_scored top = rxo_heap_pop(heap, &len);
        uint32_t node = top.node;
        if (visited[node]) continue;
        rxo_cursor c;
        uint32_t next;
        rxo_cursor_init(&c, g, node, rev);
        while (rxo_cursor_next(&c, &next)) {
            te++;
            if (visited[next]) continue;
            double next_score = top.score + cost[c.last_e];
            if (score[next] == score[next]) { /* Some(current_score) */
                if (next_score < score[next]) {
                    score[next] = next_score;
                    rxo_heap_push(&heap, &len, &cap, (rxo_scored){next_score, next});
                }
            } else {
                score[next] = next_score;
                rxo_heap_push(&heap, &len, &cap, (rxo_scored){next_score, next});
            }
        }
        visited[node] = 1;
    }
    free(heap);
    return te;
}

/* closeness from weighted distances for a subset of sources: out[sources[i]].  invert: cost = 1/w
 * (centrality.rs:1319-1320).  The reference sums a HashMap's values (:1343, hash order); here the
 * sum runs in ascending node index. */
void rxo_weighted_closeness_partial(const rxo_graph *g, const double *weight_by_edge, int invert,
                                    int rev, const uint32_t *sources, size_t k, int wf_improved,
                                    int threads, double *out, rxo_stats *stats) {
    size_t max_index = rxo_node_bound(g), eb = rxo_edge_bound(g);
    double *cost = (double *)malloc((eb ? eb : 1) * sizeof(double));
    for (size_t e = 0; e < eb; e++) cost[e] = invert ? 1.0 / weight_by_edge[e] : weight_by_edge[e];
    if (stats) memset(stats, 0, sizeof(*stats));
    long kk = (long)k;
    (void)threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1)
    for (long i = 0; i < kk; i++) {
        uint32_t s = sources[i];
        double *score = (double *)malloc((max_index ? max_index : 1) * sizeof(double));
        uint8_t *visited = (uint8_t *)malloc(max_index ? max_index : 1);
        uint64_t te = rxo_dijkstra(g, s, cost, rev, score, visited, max_index);
        size_t reachable = 0;
        double dists_sum = 0.0;
        for (size_t v = 0; v < max_index; v++)
            if (score[v] == score[v]) {
                reachable++;
                dists_sum += score[v];
            }
        double c = 0.0;
        if (reachable != 1) {
            c = (double)(reachable - 1) / dists_sum;
            if (wf_improved) c = c * (double)(reachable - 1) / (double)(g->node_count - 1);
        }
        out[s] = c;
        rxo_stats_add(stats, te, 0, reachable, 0);
        free(score);
        free(visited);
    }
    free(cost);
}

/* newman_weighted_closeness_centrality, centrality.rs:1297-1357 */
void rxo_newman_weighted_closeness_centrality(const rxo_graph *g, const double *weight_by_edge,
                                              int wf_improved, size_t parallel_threshold,
                                              double *out, uint8_t *present, rxo_stats *stats) {
    size_t max_index = rxo_node_bound(g);
    for (size_t i = 0; i < max_index; i++) {
        out[i] = 0.0;
        present[i] = g->nodes[i].alive;
    }
    size_t k;
    uint32_t *ids = rxo_live_nodes(g, &k);
    int threads = (g->node_count >= parallel_threshold) ? rxo_max_threads() : 1;
    rxo_weighted_closeness_partial(g, weight_by_edge, 1, 1, ids, k, wf_improved, threads, out, stats);
    free(ids);
}

/* all_pairs_dijkstra_path_lengths, src/shortest_path/all_pairs_dijkstra.rs:36-105, for a subset of
 * sources: out[i * node_bound + v] = length sources[i] -> v, NAN where the map has no entry. */
void rxo_dijkstra_lengths_rows(const rxo_graph *g, const double *cost_by_edge, const uint32_t *sources,
                               size_t k, int threads, double *out, rxo_stats *stats) {
    size_t max_index = rxo_node_bound(g);
    if (stats) memset(stats, 0, sizeof(*stats));
    long kk = (long)k;
    (void)threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1)
    for (long i = 0; i < kk; i++) {
        uint8_t *visited = (uint8_t *)malloc(max_index ? max_index : 1);
        uint64_t te = rxo_dijkstra(g, sources[i], cost_by_edge, 0, out + (size_t)i * max_index, visited,
                                   max_index);
        rxo_stats_add(stats, te, 0, 0, 0);
        free(visited);
    }
}
