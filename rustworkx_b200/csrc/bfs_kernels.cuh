// bfs_kernels.cuh -- distance-only batched multi-source BFS (closeness, distance_matrix).
//
// Replaces:
//   closeness per-node BFS over Reversed(&graph) : rustworkx-core/src/centrality.rs:1190-1220
//   dense bitset BFS + row fill                  : rustworkx-core/src/shortest_path/distance_matrix.rs:127-157
//
// Bit-parallel: 32 sources share one 32-bit word per vertex (the sparse analogue of the reference's
// FixedBitSet rows).  Per group g:
//   visited[g][v]  bit b = source b has reached v
//   nextA/nextB[g][v]  bits pushed into v for the level being built (double buffered)
// A level is ONE launch: the owner thread of each queued vertex finalises its new bits (counting,
// distance rows), then pushes them along its arcs with atomicOr; the first push into a vertex
// enqueues it for the next level.  Work per level is proportional to the frontier, which is what
// the 200..2000-level graphs of config 2 (heavy-hex, grid) need.
#pragma once

#include "common.cuh"

namespace rxc {

struct BfsParams {
    const uint32_t *row;  // adjacency to expand along (in-arcs for closeness, out-arcs for distances)
    const uint32_t *col;
    uint32_t *visited;    // [NG][n]
    uint32_t *next0;      // [NG][n]
    uint32_t *next1;      // [NG][n]
    uint32_t *q0;         // [NG][n] vertex queue, even levels
    uint32_t *q1;         // [NG][n] vertex queue, odd levels
    uint32_t *cnt;        // [3][NG] queue lengths, rotating by level % 3
    uint32_t *lvl_total;  // [lcap]
    unsigned long long *reach;  // [NG][32] vertices reached per source (incl. the source)
    unsigned long long *dsum;   // [NG][32] sum of distances per source
    unsigned long long *stats;  // [ST_COUNT]
    double *rows;         // [NG*32][n_cols] distance rows of this batch, or nullptr
    const uint32_t *src;  // [NG][32] compact source ids, NO_SOURCE padded
    uint64_t n_cols;
    uint32_t n;
    uint32_t ngroups;
};

__global__ void __launch_bounds__(BLOCK) bfs_fill_f64_kernel(double *x, uint64_t len, double v) {
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < len;
         i += (uint64_t)gridDim.x * BLOCK)
        x[i] = v;
}

// visited = ~valid, next0 = next1 = 0 for every vertex of every group
__global__ void __launch_bounds__(BLOCK) bfs_init_kernel(BfsParams p) {
    const uint32_t g = blockIdx.y;
    const uint32_t s = p.src[g * GW + lane_id()];
    const uint32_t valid = __ballot_sync(FULL, s != NO_SOURCE);
    const uint32_t w = blockIdx.x * BLOCK + threadIdx.x;
    if (w < p.n) {
        const size_t i = (size_t)g * p.n + w;
        p.visited[i] = ~valid;
        p.next0[i] = 0;
        p.next1[i] = 0;
    }
}

// one warp per group: queue the sources as level 0 with their bit already "pushed"
__global__ void __launch_bounds__(32) bfs_seed_kernel(BfsParams p) {
    const uint32_t g = blockIdx.x;
    const int lane = lane_id();
    const uint32_t s = p.src[g * GW + lane];
    const bool ok = s != NO_SOURCE;
    const uint32_t valid = __ballot_sync(FULL, ok);
    if (ok) {
        const size_t gn = (size_t)g * p.n;
        p.next0[gn + s] = 1u << lane;  // sources of one group are distinct vertices
        p.q0[gn + __popc(valid & lanemask_lt())] = s;
    }
    p.reach[g * GW + lane] = 0ull;
    p.dsum[g * GW + lane] = 0ull;
    if (lane == 0) {
        p.cnt[0 * p.ngroups + g] = __popc(valid);
        p.cnt[1 * p.ngroups + g] = 0;
        p.cnt[2 * p.ngroups + g] = 0;
        atomicAdd(&p.lvl_total[0], __popc(valid));
    }
}

// group closeness (rustworkx-core/src/centrality.rs:1860-1916): ONE multi-source BFS.
// Slot 0 of group 0 is seeded at every member of the group; src[0] only marks the slot as valid.
__global__ void __launch_bounds__(BLOCK)
    bfs_seed_multi_kernel(BfsParams p, const uint32_t *members, uint32_t nmem) {
    const uint32_t i = blockIdx.x * BLOCK + threadIdx.x;
    if (i < nmem) {
        p.next0[members[i]] = 1u;  // members are distinct
        p.q0[i] = members[i];
    }
    if (i < GW) {
        p.reach[i] = 0ull;
        p.dsum[i] = 0ull;
    }
    if (i == 0) {
        p.cnt[0] = nmem;
        p.cnt[1] = 0;
        p.cnt[2] = 0;
        p.lvl_total[0] = nmem;
    }
}

// One BFS level for every group of the batch (thread per queued vertex).
template <bool ROWS>
__global__ void __launch_bounds__(BLOCK) bfs_level_kernel(BfsParams p, uint32_t level) {
    const uint32_t g = blockIdx.y;
    const uint32_t *cnt_in = p.cnt + (size_t)(level % 3) * p.ngroups;
    uint32_t *cnt_out = p.cnt + (size_t)((level + 1) % 3) * p.ngroups;
    uint32_t *cnt_clr = p.cnt + (size_t)((level + 2) % 3) * p.ngroups;
    const uint32_t cnt = cnt_in[g];
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt_clr[g] = 0;
    if (cnt == 0) return;

    const int lane = lane_id();
    const size_t gn = (size_t)g * p.n;
    uint32_t *visited = p.visited + gn;
    uint32_t *nextA = ((level & 1) ? p.next1 : p.next0) + gn;
    uint32_t *nextB = ((level & 1) ? p.next0 : p.next1) + gn;
    const uint32_t *qin = ((level & 1) ? p.q1 : p.q0) + gn;
    uint32_t *qout = ((level & 1) ? p.q0 : p.q1) + gn;

    __shared__ uint32_t s_enq;
    __shared__ unsigned long long s_reach[GW], s_te, s_v;
    if (threadIdx.x == 0) {
        s_enq = 0;
        s_te = 0;
        s_v = 0;
    }
    if (threadIdx.x < GW) s_reach[threadIdx.x] = 0;
    __syncthreads();

    unsigned long long my_reach = 0, my_te = 0, my_v = 0;
    uint32_t my_enq = 0;
    const uint32_t cnt_round = (cnt + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * BLOCK + threadIdx.x; i < cnt_round; i += gridDim.x * BLOCK) {
        uint32_t w = 0, nm = 0;
        if (i < cnt) {
            w = qin[i];
            nm = nextA[w];
            nextA[w] = 0;
            const uint32_t vis = visited[w];
            nm &= ~vis;
            if (nm) visited[w] = vis | nm;
        }
        // per-source counts: transpose the warp's 32 words with ballots
        if (__ballot_sync(FULL, nm != 0)) {
#pragma unroll 8
            for (int b = 0; b < 32; b++) {
                const uint32_t c = __popc(__ballot_sync(FULL, (nm >> b) & 1u));
                if (lane == b) my_reach += c;
            }
        }
        if (nm) {
            if (ROWS) {
                uint32_t bits = nm;
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    p.rows[(size_t)(g * GW + b) * p.n_cols + w] = (double)level;
                }
            }
            const uint32_t e0 = p.row[w], e1 = p.row[w + 1];
            my_te += (unsigned long long)__popc(nm) * (e1 - e0);
            my_v += __popc(nm);
            for (uint32_t e = e0; e < e1; e++) {
                const uint32_t x = p.col[e];
                const uint32_t pm = nm & ~visited[x];
                if (pm) {
                    const uint32_t old = atomicOr(&nextB[x], pm);
                    if (old == 0) {
                        const uint32_t pos = atomicAdd(&cnt_out[g], 1u);
                        qout[pos] = x;
                        my_enq++;
                    }
                }
            }
        }
    }
    if (my_reach) atomicAdd(&s_reach[lane], my_reach);
    if (my_enq) atomicAdd(&s_enq, my_enq);
    if (my_te) atomicAdd(&s_te, my_te);
    if (my_v) atomicAdd(&s_v, my_v);
    __syncthreads();
    if (threadIdx.x < GW && s_reach[threadIdx.x]) {
        atomicAdd(&p.reach[g * GW + threadIdx.x], s_reach[threadIdx.x]);
        atomicAdd(&p.dsum[g * GW + threadIdx.x], s_reach[threadIdx.x] * level);
    }
    if (threadIdx.x == 0) {
        if (s_enq) atomicAdd(&p.lvl_total[level + 1], s_enq);
        if (s_te) stat_add(p.stats, ST_TE, s_te);
        if (s_v) stat_add(p.stats, ST_V, s_v);
    }
}

// closeness_finalize: centrality.rs:1209-1220 in the reference's operation order, every step
// explicitly round-to-nearest so the result is bit-identical to the CPU.
__global__ void __launch_bounds__(BLOCK)
    closeness_finalize_kernel(const unsigned long long *reach, const unsigned long long *dsum,
                              const uint32_t *src, uint32_t nslots, int wf_improved,
                              unsigned long long node_count, double *out /* [nslots] */) {
    const uint32_t i = blockIdx.x * BLOCK + threadIdx.x;
    if (i >= nslots) return;
    if (src[i] == NO_SOURCE) return;
    const unsigned long long r = reach[i];
    double c = 0.0;
    if (r != 1) {
        c = __ddiv_rn((double)(r - 1), (double)dsum[i]);
        if (wf_improved) {
            c = __dmul_rn(c, (double)(r - 1));
            c = __ddiv_rn(c, (double)(node_count - 1));
        }
    }
    out[i] = c;
}



// ================================================================================================
// One CTA per source with the visited bitmap in shared memory -- the high-diameter regime.
//
// On lattices and other long, thin graphs (grid 1000x1000: 1998 levels, heavy-hex: ~200) the
// frontiers of different sources never coincide, so bit-parallel words hold one or two sources and
// every level launch touches gigabytes of per-group state for a few thousand vertices.  Here a
// source's whole BFS stays on one SM: the visited bitmap (n/8 bytes; 125 KB at n = 10^6) and the
// first QS entries of both frontier queues live in shared memory, only the CSR is read from
// L2/HBM, and a level costs two __syncthreads instead of a kernel launch.  CTAs are persistent and
// take sources from a global counter.
// ================================================================================================
constexpr uint32_t SB_QS = 8192;  // most frontier entries kept in shared memory per queue

struct SmemBfsParams {
    const uint32_t *row;
    const uint32_t *col;
    const uint4 *ell;              // [n] adjacency when every vertex has <= 4 arcs, else nullptr
    const uint32_t *sources;       // compact ids
    uint32_t *qspill;              // [gridDim.x][2][n] queue entries beyond SB_QS
    uint32_t *gbitmap;             // [gridDim.x][words] visited bitmaps when they do not fit in smem
    uint32_t *next_source;         // dynamic source counter
    unsigned long long *reach;     // [nsrc]
    unsigned long long *dsum;      // [nsrc]
    uint32_t *depth;               // [nsrc] number of BFS levels of each source, or nullptr
    unsigned long long *stats;
    double *rows;                  // [nsrc][n_cols] or nullptr
    uint64_t n_cols;
    uint32_t n;
    uint32_t nsrc;
    uint32_t words;                // ceil(n / 32)
    uint32_t qs;                   // frontier entries per queue held in shared memory (<= SB_QS)
};

// ELL: every vertex has at most 4 arcs (grids, heavy-hex and the other coupling-map lattices) and
// the adjacency is one uint4 per vertex (0xFFFFFFFF padded): one dependent load per level instead
// of two (CSR offsets, then targets).
template <bool ROWS, bool GLOBAL_BITMAP, bool ELL>
__global__ void __launch_bounds__(1024, 1) bfs_smem_kernel(SmemBfsParams p) {
    extern __shared__ uint32_t sb_smem[];
    // the bitmap lives in shared memory when it fits (n <= ~1.3 M), else in this CTA's slab of
    // global memory (L2-resident: 148 slabs of n/8 bytes)
    uint32_t *vis = GLOBAL_BITMAP ? p.gbitmap + (size_t)blockIdx.x * p.words : sb_smem;
    uint32_t *q0 = GLOBAL_BITMAP ? sb_smem : sb_smem + p.words;
    uint32_t *q1 = q0 + p.qs;
    __shared__ uint32_t s_src, s_tail[2];
    __shared__ unsigned long long s_te;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    uint32_t *spill0 = p.qspill + (size_t)blockIdx.x * 2 * p.n;
    uint32_t *spill1 = spill0 + p.n;

    for (;;) {
        if (tid == 0) {
            s_src = atomicAdd(p.next_source, 1u);
            s_tail[0] = 0;
            s_tail[1] = 0;
            s_te = 0;
        }
        __syncthreads();
        const uint32_t si = s_src;
        if (si >= p.nsrc) break;
        const uint32_t s = p.sources[si];
        for (uint32_t i = tid; i < p.words; i += nthr) vis[i] = 0;
        __syncthreads();
        if (tid == 0) {
            vis[s >> 5] = 1u << (s & 31);
            q0[0] = s;
        }
        __syncthreads();
        uint32_t size = 1, level = 0;
        unsigned long long reach = 0, dsum = 0, te = 0;
        uint32_t *qc = q0, *qn = q1, *sc = spill0, *sn = spill1;
        while (size) {
            reach += size;
            dsum += (unsigned long long)size * level;
            uint32_t *tail = &s_tail[level & 1];
            for (uint32_t i = tid; i < size; i += nthr) {
                const bool active = true;
                uint32_t v = 0;
                if (active) {
                    v = (i < p.qs) ? qc[i] : sc[i];
                    if (ROWS) p.rows[(size_t)si * p.n_cols + v] = (double)level;
                }
                if (ELL) {
                    uint4 nb = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                    if (active) nb = p.ell[v];
                    const uint32_t w[4] = {nb.x, nb.y, nb.z, nb.w};
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        if (w[u] == 0xFFFFFFFFu) continue;
                        te++;
                        const uint32_t bit = 1u << (w[u] & 31);
                        if (!(vis[w[u] >> 5] & bit)) {
                            const uint32_t old = atomicOr(&vis[w[u] >> 5], bit);
                            if (!(old & bit)) {
                                // per-thread append: a warp-aggregated one (scan + one atomic per
                                // warp) was measured slower, 235 vs 287 GTEPS on the grid
                                const uint32_t pos = atomicAdd(tail, 1u);
                                if (pos < p.qs) qn[pos] = w[u];
                                else sn[pos] = w[u];
                            }
                        }
                    }
                } else if (active) {
                    // general degrees: per-thread appends (measured: making this loop a warp
                    // collective for an aggregated append costs more than the atomics it saves,
                    // 173 -> 150 GTEPS on the grid)
                    const uint32_t e0 = p.row[v], e1 = p.row[v + 1];
                    te += e1 - e0;
                    for (uint32_t e = e0; e < e1; e += 4) {
                        // four independent target loads in flight before any is used
                        uint32_t w[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) w[u] = (e + u < e1) ? p.col[e + u] : 0xFFFFFFFFu;
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (w[u] == 0xFFFFFFFFu) continue;
                            const uint32_t bit = 1u << (w[u] & 31);
                            if (!(vis[w[u] >> 5] & bit)) {
                                const uint32_t old = atomicOr(&vis[w[u] >> 5], bit);
                                if (!(old & bit)) {
                                    const uint32_t pos = atomicAdd(tail, 1u);
                                    if (pos < p.qs) qn[pos] = w[u];
                                    else sn[pos] = w[u];
                                }
                            }
                        }
                    }
                }
            }
            __syncthreads();
            size = *tail;
            if (tid == 0) s_tail[(level + 1) & 1] = 0;  // the counter the next level appends to
            __syncthreads();
            uint32_t *t = qc; qc = qn; qn = t;
            t = sc; sc = sn; sn = t;
            level++;
        }
        if (te) atomicAdd(&s_te, te);
        __syncthreads();
        if (tid == 0) {
            p.reach[si] = reach;
            p.dsum[si] = dsum;
            if (p.depth) p.depth[si] = level;
            stat_add(p.stats, ST_TE, s_te);
            stat_add(p.stats, ST_V, reach);
        }
    }
}

// ELL adjacency of a CSR whose rows have at most 4 arcs
__global__ void __launch_bounds__(BLOCK) csr_to_ell4_kernel(const uint32_t *row, const uint32_t *col,
                                                            uint32_t n, uint4 *ell) {
    const uint32_t v = blockIdx.x * BLOCK + threadIdx.x;
    if (v >= n) return;
    const uint32_t e0 = row[v], e1 = row[v + 1];
    uint32_t w[4];
#pragma unroll
    for (int u = 0; u < 4; u++) w[u] = (e0 + u < e1) ? col[e0 + u] : 0xFFFFFFFFu;
    ell[v] = make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(BLOCK) csr_max_degree_kernel(const uint32_t *row, uint32_t n,
                                                               uint32_t *out) {
    const uint32_t v = blockIdx.x * BLOCK + threadIdx.x;
    uint32_t d = (v < n) ? row[v + 1] - row[v] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d = max(d, __shfl_xor_sync(FULL, d, o));
    if (lane_id() == 0 && d) atomicMax(out, d);
}

// (reach, sum of distances) per source as exact doubles (both < 2^53) for
// unweighted_average_shortest_path_length, src/shortest_path/average_length.rs:22-77
__global__ void __launch_bounds__(BLOCK)
    distance_sums_kernel(const unsigned long long *reach, const unsigned long long *dsum,
                         uint32_t nslots, double *out /* [nslots][2] */) {
    const uint32_t i = blockIdx.x * BLOCK + threadIdx.x;
    if (i >= nslots) return;
    out[2 * (size_t)i] = (double)reach[i];
    out[2 * (size_t)i + 1] = (double)dsum[i];
}

// closeness_finalize for the CTA-per-source path (no NO_SOURCE padding)
__global__ void __launch_bounds__(BLOCK)
    closeness_finalize_flat_kernel(const unsigned long long *reach, const unsigned long long *dsum,
                                   uint32_t nsrc, int wf_improved, unsigned long long node_count,
                                   double *out) {
    const uint32_t i = blockIdx.x * BLOCK + threadIdx.x;
    if (i >= nsrc) return;
    const unsigned long long r = reach[i];
    double c = 0.0;
    if (r != 1) {
        c = __ddiv_rn((double)(r - 1), (double)dsum[i]);
        if (wf_improved) {
            c = __dmul_rn(c, (double)(r - 1));
            c = __ddiv_rn(c, (double)(node_count - 1));
        }
    }
    out[i] = c;
}
}  // namespace rxc
