// bc_kernels.cuh -- Brandes betweenness (node + edge) as batched multi-source BFS on sm_100a.
//
// Replaces the per-source loops of the reference:
//   forward  BFS + sigma : rustworkx-core/src/centrality.rs:406-445 (node), :459-511 (edge)
//   backward dependency  : rustworkx-core/src/centrality.rs:254-299 (node), :301-327 (edge)
//
// Layout: sources are processed in GROUPS of 32 -- one warp lane per source.  For a
// group g, every per-(vertex, source) f64 lives in a 256-byte row  sc[g][v][0..31]  so that the
// random gather a BFS does per arc moves eight full 32-byte sectors instead of one 8-byte value per
// sector.  Which of the 32 sources an arc matters for is decided from 32-bit masks:
//   visited[g][v]   bit b = source b has reached v
//   P[par][g][v]    (tag << 32) | mask of sources for which v sits at the level the tag names
// The tag makes the two P arrays reusable across levels and batches without clearing.
// Every level also appends its (vertex, mask) pairs to a per-group queue; the backward sweep replays
// the queue levels in reverse and pulls from successors (no atomics on the f64 rows).
#pragma once

#include "common.cuh"

namespace rxc {

struct BcParams {
    const uint32_t *row;  // CSR offsets of the adjacency this kernel walks
    const uint32_t *col;  // CSR targets (compact vertex ids)
    const uint32_t *eid;  // arc -> original edge index (edge betweenness only)
    double *sc;           // [NG][n][32] sigma on the way down, (1+delta)/sigma on the way up
    uint64_t *P0;         // [NG][n] tagged level masks, even levels
    uint64_t *P1;         // [NG][n] tagged level masks, odd levels
    uint32_t *visited;    // [NG][n]
    uint32_t *qv;         // [NG][qcap] level queues: vertex
    uint32_t *qm;         // [NG][qcap] level queues: source mask
    uint32_t *loff;       // [NG][lcap] queue offset of each level
    uint32_t *lvlcnt;     // [NG][lcap] entries of each level
    uint32_t *lvl_total;  // [lcap] entries of each level summed over groups
    unsigned long long *reach;  // [NG][32] vertices reached per source, excluding the source
    unsigned long long *stats;  // [ST_COUNT]
    double *acc;          // node mode: [n] by compact vertex; edge mode: [e_bound] by edge index
    const uint32_t *src;  // [NG][32] compact source ids, NO_SOURCE padded
    uint64_t qcap;
    uint32_t n;
    uint32_t lcap;
    uint32_t tagbase;     // tag of level l is tagbase + l + 1
};

__device__ __forceinline__ uint64_t *bc_P(const BcParams &p, uint32_t level) {
    return (level & 1) ? p.P1 : p.P0;
}

// visited[g][*] = ~valid(g): lanes without a source count as "reached everywhere".
__global__ void __launch_bounds__(BLOCK) bc_init_kernel(BcParams p) {
    const uint32_t g = blockIdx.y;
    const uint32_t s = p.src[g * GW + lane_id()];
    const uint32_t valid = __ballot_sync(FULL, s != NO_SOURCE);
    const uint32_t w = blockIdx.x * BLOCK + threadIdx.x;
    if (w < p.n) p.visited[(size_t)g * p.n + w] = ~valid;
}

// Level 0: one warp per group seeds its (up to) 32 sources.
__global__ void __launch_bounds__(32) bc_seed_kernel(BcParams p) {
    const uint32_t g = blockIdx.x;
    const int lane = lane_id();
    const uint32_t s = p.src[g * GW + lane];
    const bool ok = s != NO_SOURCE;
    const uint32_t valid = __ballot_sync(FULL, ok);
    const uint32_t cnt = __popc(valid);
    if (ok) {
        const size_t gv = (size_t)g * p.n + s;
        const uint32_t bit = 1u << lane;
        p.visited[gv] = ~valid | bit;
        p.P0[gv] = ((uint64_t)(p.tagbase + 1) << 32) | bit;
        p.sc[gv * GW + lane] = 1.0;
        const uint32_t i = __popc(valid & lanemask_lt());
        p.qv[(size_t)g * p.qcap + i] = s;
        p.qm[(size_t)g * p.qcap + i] = bit;
    }
    p.reach[g * GW + lane] = 0ull;
    if (lane == 0) {
        p.loff[(size_t)g * p.lcap] = 0;
        p.lvlcnt[(size_t)g * p.lcap] = cnt;
        atomicAdd(&p.lvl_total[0], cnt);
    }
}

// Gather rows[vj][lane] for up to four arcs per round so that four 256-byte rows are in flight
// per warp.  v_lane/m_lane hold one arc per lane; nz = ballot of lanes whose mask is non-zero.
__device__ __forceinline__ void bc_gather(const double *__restrict__ rows, uint32_t v_lane,
                                          uint32_t m_lane, uint32_t nz, int lane, double &acc,
                                          uint32_t &seen, unsigned long long &dag) {
    while (nz) {
        uint32_t vj[4], mj[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const bool ok = nz != 0;
            const int j = ok ? (__ffs(nz) - 1) : 0;
            nz &= nz - 1;
            vj[u] = __shfl_sync(FULL, v_lane, j);
            const uint32_t mm = __shfl_sync(FULL, m_lane, j);
            mj[u] = ok ? mm : 0u;
        }
        double x[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
            x[u] = ((mj[u] >> lane) & 1u) ? rows[(size_t)vj[u] * GW + lane] : 0.0;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            acc += x[u];
            seen |= mj[u];
            dag += __popc(mj[u]);
        }
    }
}

// One BFS level for every group of the batch: vertices not yet reached by all 32 sources scan
// their in-arcs, pick up the sources whose frontier (level `level`) holds the neighbour, and sum the
// neighbours' sigma rows -- the bottom-up form of centrality.rs:427-437.  No early exit: every
// frontier parent contributes to sigma.
__global__ void __launch_bounds__(BLOCK) bc_forward_kernel(BcParams p, uint32_t level) {
    const uint32_t g = blockIdx.y;
    uint32_t *lc = p.lvlcnt + (size_t)g * p.lcap;
    uint32_t *lo = p.loff + (size_t)g * p.lcap;
    const uint32_t fcnt = lc[level];
    if (fcnt == 0) return;  // this group's BFS has finished

    __shared__ uint32_t s_w[BLOCK], s_unv[BLOCK], s_ov[BLOCK], s_om[BLOCK];
    __shared__ uint32_t s_ncand, s_next, s_nout, s_base;
    const int tid = threadIdx.x, lane = lane_id();
    if (tid == 0) {
        s_ncand = 0;
        s_next = 0;
        s_nout = 0;
    }
    __syncthreads();

    const size_t gn = (size_t)g * p.n;
    uint32_t *visited = p.visited + gn;
    const uint64_t *Pcur = bc_P(p, level) + gn;
    uint64_t *Pnext = bc_P(p, level + 1) + gn;
    double *rows = p.sc + gn * GW;
    const uint32_t tag_cur = p.tagbase + level + 1, tag_next = tag_cur + 1;

    // 1. coalesced read of this CTA's 256 visited words; compact the unfinished vertices
    {
        const uint32_t w = blockIdx.x * BLOCK + tid;
        const uint32_t unv = (w < p.n) ? ~visited[w] : 0u;
        const uint32_t bal = __ballot_sync(FULL, unv != 0);
        uint32_t wbase = 0;
        if (lane == 0 && bal) wbase = atomicAdd(&s_ncand, __popc(bal));
        wbase = __shfl_sync(FULL, wbase, 0);
        if (unv) {
            const uint32_t i = wbase + __popc(bal & lanemask_lt());
            s_w[i] = w;
            s_unv[i] = unv;
        }
    }
    __syncthreads();
    const uint32_t ncand = s_ncand;

    // 2. warps take candidates dynamically (degree skew)
    unsigned long long dag = 0;
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(&s_next, 1);
        i = __shfl_sync(FULL, i, 0);
        if (i >= ncand) break;
        const uint32_t w = s_w[i], unv = s_unv[i];
        const uint32_t e0 = p.row[w], e1 = p.row[w + 1];
        double acc = 0.0;
        uint32_t newm = 0;
        for (uint32_t e = e0; e < e1; e += 32) {
            const uint32_t idx = e + lane;
            uint32_t v = 0, m = 0;
            if (idx < e1) {
                v = p.col[idx];
                const uint64_t pv = Pcur[v];
                if ((uint32_t)(pv >> 32) == tag_cur) m = (uint32_t)pv & unv;
            }
            const uint32_t nz = __ballot_sync(FULL, m != 0);
            bc_gather(rows, v, m, nz, lane, acc, newm, dag);
        }
        if (newm) {
            if ((newm >> lane) & 1u) rows[(size_t)w * GW + lane] = acc;
            if (lane == 0) {
                visited[w] = ~unv | newm;
                Pnext[w] = ((uint64_t)tag_next << 32) | newm;
                const uint32_t j = atomicAdd(&s_nout, 1);
                s_ov[j] = w;
                s_om[j] = newm;
            }
        }
    }
    __syncthreads();

    // 3. append this CTA's discoveries to the group's level queue
    const uint32_t nout = s_nout;
    if (tid == 0) {
        const uint32_t base = lo[level] + fcnt;
        if (blockIdx.x == 0) lo[level + 1] = base;
        uint32_t pos = 0;
        if (nout) {
            pos = atomicAdd(&lc[level + 1], nout);
            atomicAdd(&p.lvl_total[level + 1], nout);
        }
        s_base = base + pos;
    }
    __syncthreads();
    if ((uint32_t)tid < nout) {
        const size_t q = (size_t)g * p.qcap + s_base + tid;
        p.qv[q] = s_ov[tid];
        p.qm[q] = s_om[tid];
    }
    (void)dag;
}

// One level of the dependency sweep (deepest level first).  A warp owns one queue entry (v, fm):
// delta[v][b] = sigma[v][b] * sum_{w in N+(v), depth_b(w) = level+1} (1+delta[w][b])/sigma[w][b]
// which regroups centrality.rs:272-279 by predecessor (pull instead of push; no f64 atomics on the
// rows).  The row of v is then overwritten with (1+delta)/sigma for the level above to pull.
//   node mode : acc[v] += sum_b delta[v][b] (+1 per reached source when endpoints), :281-298
//   edge mode : acc[eid(v->w)] += sum_b sigma[v][b] * coef[w][b], :318-325
template <bool EDGE, bool ENDPOINTS>
__global__ void __launch_bounds__(BLOCK) bc_backward_kernel(BcParams p, uint32_t level) {
    const uint32_t g = blockIdx.y;
    const uint32_t cnt = p.lvlcnt[(size_t)g * p.lcap + level];
    const uint32_t off = p.loff[(size_t)g * p.lcap + level];
    const int lane = lane_id();
    const size_t gn = (size_t)g * p.n;
    const uint64_t *Pnext = bc_P(p, level + 1) + gn;
    uint64_t *Pcur = bc_P(p, level) + gn;
    double *rows = p.sc + gn * GW;
    const uint32_t tag_cur = p.tagbase + level + 1, tag_next = tag_cur + 1;
    const uint32_t *qv = p.qv + (size_t)g * p.qcap + off;
    const uint32_t *qm = p.qm + (size_t)g * p.qcap + off;

    unsigned long long te = 0, vv = 0, reach = 0;
    unsigned long long dag = 0;
    const uint32_t warp = blockIdx.x * WARPS + (threadIdx.x >> 5);
    for (uint32_t i = warp; i < cnt; i += gridDim.x * WARPS) {
        const uint32_t v = qv[i], fm = qm[i];
        const uint32_t e0 = p.row[v], e1 = p.row[v + 1];
        const bool mine = (fm >> lane) & 1u;
        const double sig = mine ? rows[(size_t)v * GW + lane] : 0.0;
        double acc = 0.0;
        uint32_t seen = 0;
        for (uint32_t e = e0; e < e1; e += 32) {
            const uint32_t idx = e + lane;
            uint32_t w = 0, m = 0, ed = 0;
            if (idx < e1) {
                w = p.col[idx];
                const uint64_t pw = Pnext[w];
                if ((uint32_t)(pw >> 32) == tag_next) m = (uint32_t)pw & fm;
                if (EDGE && m) ed = p.eid[idx];
            }
            uint32_t nz = __ballot_sync(FULL, m != 0);
            if (!EDGE) {
                bc_gather(rows, w, m, nz, lane, acc, seen, dag);
            } else {
                while (nz) {
                    const int j = __ffs(nz) - 1;
                    nz &= nz - 1;
                    const uint32_t wj = __shfl_sync(FULL, w, j);
                    const uint32_t mj = __shfl_sync(FULL, m, j);
                    const uint32_t ej = __shfl_sync(FULL, ed, j);
                    const double coef = ((mj >> lane) & 1u) ? rows[(size_t)wj * GW + lane] : 0.0;
                    const double c = sig * coef;  // sigma[v] * coeff, centrality.rs:322
                    acc += c;
                    dag += __popc(mj);
                    const double tot = warp_sum(c);
                    if (lane == 0) atomicAdd(&p.acc[ej], tot);
                }
            }
        }
        double contrib = 0.0;
        if (mine) {
            const double delta = EDGE ? acc : sig * acc;
            rows[(size_t)v * GW + lane] = (1.0 + delta) / sig;
            if (!EDGE && level > 0) contrib = ENDPOINTS ? delta + 1.0 : delta;
        }
        if (!EDGE && level > 0) {
            const double tot = warp_sum(contrib);
            if (lane == 0) atomicAdd(&p.acc[v], tot);
        }
        if (lane == 0) Pcur[v] = ((uint64_t)tag_cur << 32) | fm;
        if (level > 0 && mine) reach++;
        if (lane == 0) {
            te += (unsigned long long)__popc(fm) * (e1 - e0);
            vv += __popc(fm);
        }
    }
    if (ENDPOINTS && reach) atomicAdd(&p.reach[g * GW + lane], reach);
    if (lane == 0 && (te | vv | dag)) {  // dag is warp-uniform
        atomicAdd(&p.stats[ST_TE], te);
        atomicAdd(&p.stats[ST_V], vv);
        atomicAdd(&p.stats[ST_DAG], (unsigned long long)dag);
    }
}

// endpoints: the source itself receives |S| - 1 (centrality.rs:282-284)
__global__ void __launch_bounds__(BLOCK) bc_endpoint_sources_kernel(BcParams p, uint32_t ngroups) {
    const uint32_t i = blockIdx.x * BLOCK + threadIdx.x;
    if (i >= ngroups * GW) return;
    const uint32_t s = p.src[i];
    if (s != NO_SOURCE) atomicAdd(&p.acc[s], (double)p.reach[i]);
}

}  // namespace rxc
