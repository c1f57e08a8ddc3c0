// snapshot.hpp -- host side of rxc_graph_snapshot: StableGraph edge list -> compact int32 CSR/CSC.
//
// Replaces the petgraph linked-list adjacency the reference walks (`graph.neighbors(v)` at
// rustworkx-core/src/centrality.rs:427, `graph.edges(v)` at :489, `Reversed` at :1208) and the
// dense FixedBitSet rows of distance_matrix.rs:107-126.  Live nodes are renumbered 0..n_live-1 in
// ascending original index (the order of `node_identifiers()`, which is also distance_matrix's
// compaction order, distance_matrix.rs:82-105); the inverse map is kept to scatter results back.
#pragma once

#include <stdint.h>

#include <stdexcept>
#include <vector>

namespace rxc {

struct HostCSR {
    std::vector<uint32_t> row;  // n + 1
    std::vector<uint32_t> col;  // arcs
    std::vector<uint32_t> eid;  // arcs: original edge index (may be empty)
};

struct HostSnapshot {
    uint32_t n_bound = 0, n_live = 0, e_bound = 0;
    uint64_t n_edges = 0;  // live edges given (self-loops included)
    bool directed = false;
    bool identity = false;                  // no holes: compact id == original index, maps left empty
    std::vector<uint32_t> orig_of_compact;  // n_live
    std::vector<uint32_t> compact_of_orig;  // n_bound, 0xFFFFFFFF for holes
    uint32_t orig_of(uint32_t compact) const { return identity ? compact : orig_of_compact[compact]; }
    // 0xFFFFFFFF when `orig` is not a live node
    uint32_t compact_of(uint32_t orig) const {
        if (orig >= n_bound) return 0xFFFFFFFFu;
        return identity ? orig : compact_of_orig[orig];
    }
    HostCSR out;  // successors (undirected: all incident arcs)
    HostCSR in;   // predecessors (empty when undirected: same as `out`)
};

struct LimitError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// counting sort of arcs (a -> b) by a; stable in edge order
inline void build_csr(uint32_t n, const std::vector<uint32_t> &a, const std::vector<uint32_t> &b,
                      const std::vector<uint32_t> &id, HostCSR &csr) {
    const size_t m = a.size();
    csr.row.assign((size_t)n + 1, 0);
    for (size_t i = 0; i < m; i++) csr.row[(size_t)a[i] + 1]++;
    for (uint32_t v = 0; v < n; v++) csr.row[v + 1] += csr.row[v];
    csr.col.resize(m);
    csr.eid.resize(id.empty() ? 0 : m);
    std::vector<uint32_t> cur(csr.row.begin(), csr.row.end() - 1);
    for (size_t i = 0; i < m; i++) {
        const uint32_t pos = cur[a[i]]++;
        csr.col[pos] = b[i];
        if (!id.empty()) csr.eid[pos] = id[i];
    }
}

inline HostSnapshot make_snapshot(uint32_t n_bound, const uint32_t *live_nodes, uint32_t n_live,
                                  uint32_t e_bound, const uint32_t *esrc, const uint32_t *edst,
                                  const uint32_t *eid, uint64_t n_edges, bool directed) {
    HostSnapshot s;
    s.n_bound = n_bound;
    s.n_live = n_live;
    s.e_bound = e_bound;
    s.n_edges = n_edges;
    s.directed = directed;
    s.orig_of_compact.assign(live_nodes, live_nodes + n_live);
    s.compact_of_orig.assign(n_bound, 0xFFFFFFFFu);
    for (uint32_t i = 0; i < n_live; i++) {
        const uint32_t o = live_nodes[i];
        if (o >= n_bound) throw std::invalid_argument("live node index >= n_bound");
        if (i > 0 && live_nodes[i - 1] >= o)
            throw std::invalid_argument("live_nodes must be strictly ascending");
        s.compact_of_orig[o] = i;
    }
    const uint64_t arcs_max = directed ? n_edges : 2 * n_edges;
    if (arcs_max >= 0xFFFFFFFFull) throw LimitError("more than 2^32-2 arcs are not supported");
    std::vector<uint32_t> a, b, id;
    a.reserve(arcs_max);
    b.reserve(arcs_max);
    id.reserve(arcs_max);
    for (uint64_t i = 0; i < n_edges; i++) {
        const uint32_t so = esrc[i], to = edst[i];
        if (so >= n_bound || to >= n_bound) throw std::invalid_argument("edge endpoint >= n_bound");
        const uint32_t u = s.compact_of_orig[so], v = s.compact_of_orig[to];
        if (u == 0xFFFFFFFFu || v == 0xFFFFFFFFu)
            throw std::invalid_argument("edge endpoint is not a live node");
        const uint32_t e = eid ? eid[i] : (uint32_t)i;
        if (e >= e_bound) throw std::invalid_argument("edge index >= e_bound");
        if (u == v) continue;  // self-loop: distance[w] == d+1 can never hold (centrality.rs:433)
        a.push_back(u);
        b.push_back(v);
        id.push_back(e);
        if (!directed) {
            a.push_back(v);
            b.push_back(u);
            id.push_back(e);
        }
    }
    build_csr(n_live, a, b, id, s.out);
    if (directed) build_csr(n_live, b, a, id, s.in);
    return s;
}

// union of out- and in-arcs of a directed snapshot (distance_matrix as_undirected=true,
// distance_matrix.rs:108-116); duplicates are harmless for distances
inline HostCSR make_symmetric(const HostSnapshot &s) {
    const uint32_t n = s.n_live;
    HostCSR r;
    r.row.assign((size_t)n + 1, 0);
    for (uint32_t v = 0; v < n; v++)
        r.row[v + 1] = r.row[v] + (s.out.row[v + 1] - s.out.row[v]) + (s.in.row[v + 1] - s.in.row[v]);
    r.col.resize(r.row[n]);
    for (uint32_t v = 0; v < n; v++) {
        uint32_t pos = r.row[v];
        for (uint32_t e = s.in.row[v]; e < s.in.row[v + 1]; e++) r.col[pos++] = s.in.col[e];
        for (uint32_t e = s.out.row[v]; e < s.out.row[v + 1]; e++) r.col[pos++] = s.out.col[e];
    }
    return r;
}

}  // namespace rxc
