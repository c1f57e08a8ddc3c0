// wsp_kernels.cuh -- all-sources WEIGHTED shortest path lengths as batched
// multi-source label correcting on sm_100a.
//
// Replaces the per-source binary-heap Dijkstra loops of the reference:
//   newman_weighted_closeness_centrality : rustworkx-core/src/centrality.rs:1297-1357
//                                          (dijkstra over Reversed(&graph), cost = 1 / weight)
//   all_pairs_dijkstra_path_lengths      : src/shortest_path/all_pairs_dijkstra.rs:36-105
//                                          (rustworkx-core/src/shortest_path/dijkstra.rs:100-160)
//
// Why the results are bit-identical to Dijkstra: with non-negative costs Dijkstra returns, for every
// vertex, the minimum over all paths of the path's cost summed LEFT TO RIGHT in f64 (rounded
// addition is monotone and never decreases a non-negative partial sum, so the greedy argument
// holds for the rounded sums too).  A label-correcting iteration  d[w] = min(d[w], d[v] + c(v,w))
// converges to exactly that fixed point, whatever the order of the relaxations.
//
// Layout: the betweenness layout (bc_kernels.cuh).  Groups of 128 sources; per vertex one 1 KB row
// dist[g][v][0..127], lane l owning the sector of sources {l, 32+l, 64+l, 96+l}; per vertex a tagged
// 128-bit mask "the sources whose distance at v decreased in the previous round" (the P entries of
// the betweenness kernels); a per-group bitmap `maybe` of the vertices that have an in-neighbour
// which changed.  One round = mark (push the changed set's out-neighbours into `maybe`) + relax
// (every marked vertex PULLS d[v] + c over its in-arcs from the changed neighbours, rows staged
// through shared memory with cp.async exactly like the sigma gathers).  Rows are only written by
// the warp that owns the vertex; a row read while its owner updates it yields an older or a newer
// valid path length, both harmless (the newer one is flagged as changed again next round).
#pragma once

#include "bc_kernels.cuh"

namespace rxc {

struct WspParams {
    const uint32_t *row;   // pull adjacency (in-arcs of the traversal direction)
    const uint32_t *col;
    const double *cost;    // [arcs of the pull adjacency] non-negative, may be +inf
    const uint32_t *srow;  // push adjacency (out-arcs of the traversal direction), for marking
    const uint32_t *scol;
    double *dist;          // [NG][n][128]
    PEntry *C0;            // [NG][n] tagged change masks, even rounds
    PEntry *C1;            // [NG][n] odd rounds
    uint32_t *maybe;       // [NG][mwords]
    uint32_t *qv;          // [NG][2][n] vertices changed in the round, by round parity
    uint32_t *lvlcnt;      // [NG][lcap] changed vertices per round
    uint32_t *lvl_total;   // [lcap] summed over groups
    unsigned long long *stats;
    const uint32_t *src;   // [NG][128]
    uint64_t qstride;      // entries per group in qv
    uint32_t n;
    uint32_t lcap;
    uint32_t mwords;
    uint32_t ngroups;
    uint32_t tagbase;
};

// row slots + arc table + the per-arc costs, per warp
// Newman closeness, BA(10^6,8), 1024 sources: 4 slots (4 CTAs/SM) 143.1 ms, 5 slots 143.0, 6 slots
// (3 CTAs/SM) 138.6
#ifndef RXC_WSTAGE
#define RXC_WSTAGE 6
#endif
constexpr int WSTAGE = RXC_WSTAGE;  // f64 row slots per warp in wsp_relax_kernel
constexpr int WSP_SMEM_BYTES = WARPS * (WSTAGE * BG * 8 + ARCTAB_BYTES + ARCTAB_AUX_BYTES);

__device__ __forceinline__ PEntry *wsp_C(const WspParams &p, uint32_t round) {
    return (round & 1) ? p.C1 : p.C0;
}
__device__ __forceinline__ double wsp_inf() { return __longlong_as_double(0x7FF0000000000000ll); }

// dist = +inf everywhere
__global__ void __launch_bounds__(BLOCK) wsp_fill_kernel(double *x, uint64_t len) {
    const double inf = wsp_inf();
    for (uint64_t i = ((uint64_t)blockIdx.x * BLOCK + threadIdx.x) * 2; i < len;
         i += (uint64_t)gridDim.x * BLOCK * 2)
        *reinterpret_cast<double2 *>(x + i) = make_double2(inf, inf);
}

// round 0: every source is "changed" with distance 0
__global__ void __launch_bounds__(32) wsp_seed_kernel(WspParams p) {
    const uint32_t g = blockIdx.x;
    const int lane = lane_id();
    uint32_t base = 0;
#pragma unroll
    for (int k = 0; k < BK; k++) {
        const uint32_t s = p.src[(size_t)g * BG + k * 32 + lane];
        const uint32_t valid = __ballot_sync(FULL, s != NO_SOURCE);
        if (s != NO_SOURCE) {
            const size_t gv = (size_t)g * p.n + s;
            SrcMask bit = mask_zero();
            bit.w[k] = 1u << lane;
            bc_store_level(p.C0 + (size_t)g * p.n, s, p.tagbase + 1, bit);  // sources of a group are distinct
            p.dist[gv * BG + lane * BK + k] = 0.0;
            p.qv[(size_t)g * p.qstride + base + __popc(valid & lanemask_lt())] = s;
        }
        base += __popc(valid);
    }
    if (lane == 0) {
        p.lvlcnt[(size_t)g * p.lcap] = base;
        atomicAdd(&p.lvl_total[0], base);
    }
}

// maybe[w] = 1 for every out-neighbour w of a vertex that changed in `round`
__global__ void __launch_bounds__(BLOCK) wsp_mark_kernel(WspParams p, uint32_t round) {
    const uint32_t g = blockIdx.y;
    const uint32_t cnt = p.lvlcnt[(size_t)g * p.lcap + round];
    if (cnt == 0) return;
    const uint32_t *qv = p.qv + (size_t)g * p.qstride + (size_t)(round & 1) * p.n;
    uint32_t *maybe = p.maybe + (size_t)g * p.mwords;
    const int lane = lane_id();
    for (uint32_t i = blockIdx.x * WARPS + (threadIdx.x >> 5); i < cnt; i += gridDim.x * WARPS) {
        const uint32_t v = qv[i];
        const uint32_t e0 = p.srow[v], e1 = p.srow[v + 1];
        for (uint32_t e = e0 + lane; e < e1; e += 32) {
            const uint32_t w = p.scol[e];
            const uint32_t bit = 1u << (w & 31);
            if (!(maybe[w >> 5] & bit)) atomicOr(&maybe[w >> 5], bit);
        }
    }
}

// One relaxation round: every marked vertex pulls d[v] + c(v, w) from the in-neighbours that
// changed in `round`, for the sources that changed there.
__global__ void __launch_bounds__(BLOCK, RXC_MINBLOCKS) wsp_relax_kernel(WspParams p, uint32_t round) {
    const uint32_t g = blockIdx.y;
    uint32_t *lc = p.lvlcnt + (size_t)g * p.lcap;
    if (lc[round] == 0) return;  // this group has converged

    __shared__ uint32_t s_w[BLOCK], s_e0[BLOCK], s_e1[BLOCK], s_ov[BLOCK];
    __shared__ uint32_t s_ncand, s_next, s_nout, s_base;
    extern __shared__ __align__(16) char s_dyn_stage[];
    char *stage = s_dyn_stage + (size_t)(threadIdx.x >> 5) * (WSP_SMEM_BYTES / WARPS);
    const int tid = threadIdx.x, lane = lane_id();
    if (tid == 0) {
        s_ncand = 0;
        s_next = 0;
        s_nout = 0;
    }
    __syncthreads();

    const size_t gn = (size_t)g * p.n;
    const PEntry *Ccur = wsp_C(p, round) + gn;
    PEntry *Cnext = wsp_C(p, round + 1) + gn;
    double *rows = p.dist + gn * BG;
    const uint32_t tag_cur = p.tagbase + round + 1, tag_next = tag_cur + 1;

    // 1. this CTA's 256 vertices: consume the `maybe` word, compact the marked ones with their offsets
    {
        const uint32_t w = blockIdx.x * BLOCK + tid;
        uint32_t *word = p.maybe + (size_t)g * p.mwords + (w >> 5);
        const uint32_t bits = (w < p.n) ? *word : 0u;  // one word per warp
        const bool look = (bits >> lane) & 1u;
        __syncwarp();
        if (lane == 0 && bits) *word = 0;
        uint32_t e0 = 0, e1 = 0;
        if (look) {
            e0 = p.row[w];
            e1 = p.row[w + 1];
        }
        const uint32_t bal = __ballot_sync(FULL, look);
        uint32_t wbase = 0;
        if (lane == 0 && bal) wbase = atomicAdd(&s_ncand, __popc(bal));
        wbase = __shfl_sync(FULL, wbase, 0);
        if (look) {
            const uint32_t i = wbase + __popc(bal & lanemask_lt());
            s_w[i] = w;
            s_e0[i] = e0;
            s_e1[i] = e1;
        }
    }
    __syncthreads();
    const uint32_t ncand = s_ncand;

    // 2. warps take marked vertices dynamically
    unsigned long long relaxed = 0;
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(&s_next, 1);
        i = __shfl_sync(FULL, i, 0);
        if (i >= ncand) break;
        const uint32_t w = s_w[i], e0 = s_e0[i], e1 = s_e1[i];
        double best[BK];
#pragma unroll
        for (int k = 0; k < BK; k++) best[k] = wsp_inf();
        SrcMask touched = mask_zero();
        for (uint32_t e = e0; e < e1; e += 32) {
            const uint32_t idx = e + lane;
            uint32_t v = 0;
            double c = 0.0;
            SrcMask m = mask_zero();
            if (idx < e1) {
                v = p.col[idx];
                c = p.cost[idx];
                m = bc_level_mask(Ccur, v, tag_cur);
            }
            const uint32_t nz = __ballot_sync(FULL, mask_any(m));
            if (nz) {  // once per look-up instead of once per row
#pragma unroll
                for (int k = 0; k < BK; k++) touched.w[k] |= __reduce_or_sync(FULL, m.w[k]);
                relaxed += mask_popc(m);  // per lane; summed over the warp at the end
                // next_score = node_score + cost, dijkstra.rs:128 -- through the arc table of the
                // betweenness gathers, the arc's cost riding along as the per-arc scalar
                bc_gather<WSTAGE, double, double, GatherMinPlus>(rows, v, m, nz, lane, stage, best, c);
            }
        }
        if (!mask_any(touched)) continue;
        // 3. compare with the vertex's own row; keep strict improvements (dijkstra.rs:131)
        const bool mine = (mask_or(touched) >> lane) & 1u;
        const RowK own = row_load(rows, w, lane, mine);
        uint32_t imp[BK];
#pragma unroll
        for (int k = 0; k < BK; k++) {
            const bool better = mine && ((touched.w[k] >> lane) & 1u) && best[k] < own.x[k];
            if (better) rows[(size_t)w * BG + lane * BK + k] = best[k];
            imp[k] = __ballot_sync(FULL, better);
        }
        uint32_t imp_any = 0;
#pragma unroll
        for (int k = 0; k < BK; k++) imp_any |= imp[k];
        if (imp_any) {
            if (lane == 0) {
                SrcMask nm;
#pragma unroll
                for (int k = 0; k < BK; k++) nm.w[k] = imp[k];
                bc_store_level(Cnext, w, tag_next, nm);
                s_ov[atomicAdd(&s_nout, 1u)] = w;
            }
        }
    }
    __syncthreads();

    // 4. append the vertices this CTA improved to the group's queue of the next round
    const uint32_t nout = s_nout;
    if (tid == 0) {
        uint32_t pos = 0;
        if (nout) {
            pos = atomicAdd(&lc[round + 1], nout);
            atomicAdd(&p.lvl_total[round + 1], nout);
        }
        s_base = pos;
    }
    __syncthreads();
    if ((uint32_t)tid < nout)
        p.qv[(size_t)g * p.qstride + (size_t)((round + 1) & 1) * p.n + s_base + tid] = s_ov[tid];
    relaxed = warp_sum_u64(relaxed);
    if (lane == 0 && relaxed) stat_add(p.stats, ST_DAG, relaxed);
}

// ---------------------------------------------------------------------------------------------
// results
// ---------------------------------------------------------------------------------------------
constexpr uint32_t WSP_TILE = 2048;  // vertices per CTA in the reductions below

// Per source: number of vertices with a finite distance, the sum of those distances (in vertex
// order inside a tile, tiles combined in order by wsp_sums_finish_kernel: deterministic), and the
// work counter TE = sum of out-degrees of the reached vertices.
__global__ void __launch_bounds__(BG) wsp_sums_tile_kernel(WspParams p, double *part_sum,
                                                           unsigned long long *part_cnt) {
    const uint32_t g = blockIdx.y, tile = blockIdx.x, slot = threadIdx.x;
    // slot t of the row is source 32*(t % 4) + t / 4 (lane t/4, k = t%4)
    const double *rows = p.dist + (size_t)g * p.n * BG;
    const uint32_t v0 = tile * WSP_TILE, v1 = min(v0 + WSP_TILE, p.n);
    double sum = 0.0;
    unsigned long long cnt = 0, te = 0;
    for (uint32_t v = v0; v < v1; v++) {
        const double d = rows[(size_t)v * BG + slot];
        if (d < wsp_inf()) {
            sum += d;
            cnt++;
            te += p.srow[v + 1] - p.srow[v];
        }
    }
    const size_t o = ((size_t)g * gridDim.x + tile) * BG + slot;
    part_sum[o] = sum;
    part_cnt[o] = cnt;
    const bool valid = p.src[(size_t)g * BG + 32 * (slot % BK) + slot / BK] != NO_SOURCE;
    te = valid ? te : 0;
    cnt = valid ? cnt : 0;
    te = warp_sum_u64(te);
    cnt = warp_sum_u64(cnt);
    if ((slot & 31) == 0 && (te | cnt)) {
        stat_add(p.stats, ST_TE, te);
        stat_add(p.stats, ST_V, cnt);
    }
}

// closeness of every source from the tile partials (centrality.rs:1336-1350), by slot ORDER of src
__global__ void __launch_bounds__(BG)
    wsp_sums_finish_kernel(WspParams p, const double *part_sum, const unsigned long long *part_cnt,
                           uint32_t ntiles, int wf_improved, unsigned long long node_count,
                           double *out /* [NG][128] indexed like src */) {
    const uint32_t g = blockIdx.x, slot = threadIdx.x;
    double sum = 0.0;
    unsigned long long cnt = 0;
    for (uint32_t t = 0; t < ntiles; t++) {
        const size_t o = ((size_t)g * ntiles + t) * BG + slot;
        sum += part_sum[o];
        cnt += part_cnt[o];
    }
    const uint32_t si = 32 * (slot % BK) + slot / BK;  // index into src of this row slot
    if (p.src[(size_t)g * BG + si] == NO_SOURCE) return;
    double c = 0.0;
    if (cnt > 1) {
        c = __ddiv_rn((double)(cnt - 1), sum);
        if (wf_improved) {
            c = __dmul_rn(c, (double)(cnt - 1));
            c = __ddiv_rn(c, (double)(node_count - 1));
        }
    }
    out[(size_t)g * BG + si] = c;
    // exact (reached, sum) pairs for callers that need the reached count (inf-cost edges)
    out[(size_t)gridDim.x * BG + ((size_t)g * BG + si) * 2] = (double)cnt;
    out[(size_t)gridDim.x * BG + ((size_t)g * BG + si) * 2 + 1] = sum;
}

// rows of path lengths: out[(g*128 + si) * n_cols + v] = dist[g][v][slot(si)]  (32x32 smem transpose)
__global__ void __launch_bounds__(BLOCK) wsp_rows_kernel(WspParams p, double *out, uint64_t n_cols) {
    __shared__ double tile[32][33];
    const uint32_t g = blockIdx.z;
    const uint32_t v0 = blockIdx.x * 32, t0 = blockIdx.y * 32;  // 32 vertices x 32 row slots
    const double *rows = p.dist + (size_t)g * p.n * BG;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const uint32_t v = v0 + r;
        tile[r][tx] = (v < p.n) ? rows[(size_t)v * BG + t0 + tx] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const uint32_t slot = t0 + r;
        const uint32_t si = 32 * (slot % BK) + slot / BK;
        const uint32_t v = v0 + tx;
        if (v < p.n && p.src[(size_t)g * BG + si] != NO_SOURCE)
            out[((size_t)g * BG + si) * n_cols + v] = tile[tx][r];
    }
}

// cost[a] = w[eid[a]] or 1 / w[eid[a]] (Newman: stronger ties are shorter, centrality.rs:1319-1320)
__global__ void __launch_bounds__(BLOCK)
    wsp_arc_costs_kernel(const uint32_t *eid, const double *w_by_edge, uint64_t arcs, int invert,
                         double *cost) {
    for (uint64_t a = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; a < arcs; a += (uint64_t)gridDim.x * BLOCK) {
        const double w = w_by_edge[eid[a]];
        cost[a] = invert ? __ddiv_rn(1.0, w) : w;
    }
}

}  // namespace rxc
