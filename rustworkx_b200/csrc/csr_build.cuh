// csr_build.cuh -- device side of rxc_graph_snapshot: edge list -> compact int32 CSR/CSC in HBM.
//
// Replaces walking petgraph's per-node linked edge lists (`graph.neighbors(v)` at
// rustworkx-core/src/centrality.rs:427, `graph.edges(v)` at :489, `Reversed` at :1208) and the dense
// FixedBitSet adjacency of distance_matrix.rs:107-126.  The raw edge arrays (12 bytes per edge) are
// copied to the device once; histogram -> exclusive scan -> scatter builds the row offsets, targets
// and per-arc original edge ids without touching the host again.  Self-loops are dropped (they never
// satisfy distance[w] == d+1, centrality.rs:433), parallel edges are kept (path multiplicity).
// The order of arcs inside a row follows the scatter's atomics; it only permutes f64 additions.
#pragma once

#include "common.cuh"

namespace rxc {

constexpr uint32_t CSR_BAD = 0xFFFFFFFFu;
enum CsrError { CSR_ERR_ENDPOINT_RANGE = 1, CSR_ERR_ENDPOINT_DEAD = 2, CSR_ERR_EDGE_ID = 4 };

__global__ void __launch_bounds__(BLOCK) csr_fill_u32_kernel(uint32_t *x, uint64_t len, uint32_t v) {
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < len;
         i += (uint64_t)gridDim.x * BLOCK)
        x[i] = v;
}

// compact_of_orig[live[i]] = i
__global__ void __launch_bounds__(BLOCK)
    csr_map_kernel(const uint32_t *live, uint32_t n_live, uint32_t n_bound, uint32_t *cmap,
                   uint32_t *err) {
    const uint32_t i = blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n_live) return;
    const uint32_t o = live[i];
    if (o >= n_bound) {
        atomicOr(err, (uint32_t)CSR_ERR_ENDPOINT_RANGE);
        return;
    }
    cmap[o] = i;
}

// Translate endpoints to compact ids (in place) and count arcs per row.
//   by_src[u]++ for u -> v;  by_dst[v]++ as well (same array when undirected: both directions).
__global__ void __launch_bounds__(BLOCK)
    csr_count_kernel(uint32_t *src, uint32_t *dst, const uint32_t *eid, uint64_t n_edges,
                     const uint32_t *cmap, uint32_t n_bound, uint32_t e_bound, uint32_t *by_src,
                     uint32_t *by_dst, uint32_t *err) {
    for (uint64_t e = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; e < n_edges;
         e += (uint64_t)gridDim.x * BLOCK) {
        const uint32_t so = src[e], to = dst[e];
        uint32_t u = CSR_BAD, v = CSR_BAD;
        if (so >= n_bound || to >= n_bound) {
            atomicOr(err, (uint32_t)CSR_ERR_ENDPOINT_RANGE);
        } else {
            u = cmap[so];
            v = cmap[to];
            if (u == CSR_BAD || v == CSR_BAD) {
                atomicOr(err, (uint32_t)CSR_ERR_ENDPOINT_DEAD);
                u = v = CSR_BAD;
            }
        }
        if (eid && eid[e] >= e_bound) atomicOr(err, (uint32_t)CSR_ERR_EDGE_ID);
        if (u == v) u = v = CSR_BAD;  // self-loop (or invalid): no arc
        src[e] = u;
        dst[e] = v;
        if (u != CSR_BAD) {
            atomicAdd(&by_src[u], 1u);
            atomicAdd(&by_dst[v], 1u);
        }
    }
}

// ---- exclusive scan of n u32 values (three small kernels; n is a few million at most) ----------
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = BLOCK * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t x, uint32_t *total) {
    __shared__ uint32_t s_warp[WARPS];
    const int lane = lane_id(), wid = threadIdx.x >> 5;
    uint32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < WARPS ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < WARPS; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, w, o);
            if (lane >= o) w += y;
        }
        if (lane < WARPS) s_warp[lane] = w;
    }
    __syncthreads();
    const uint32_t base = wid ? s_warp[wid - 1] : 0;
    if (total) *total = s_warp[WARPS - 1];
    __syncthreads();
    return base + incl - x;
}

__global__ void __launch_bounds__(BLOCK)
    scan_tile_kernel(const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *tile_sums) {
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        sum += v[k];
    }
    uint32_t total;
    uint32_t ex = block_excl_scan(sum, &total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of the tile sums in place; writes the grand total to *total_out
__global__ void __launch_bounds__(BLOCK)
    scan_sums_kernel(uint32_t *tile_sums, uint32_t ntiles, uint32_t *total_out) {
    uint32_t carry = 0;
    for (uint32_t t0 = 0; t0 < ntiles; t0 += BLOCK) {
        const uint32_t i = t0 + threadIdx.x;
        const uint32_t x = i < ntiles ? tile_sums[i] : 0;
        uint32_t total;
        const uint32_t ex = block_excl_scan(x, &total);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(BLOCK)
    scan_add_kernel(uint32_t *out, uint32_t n, const uint32_t *tile_sums, const uint32_t *total) {
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    const uint32_t add = tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) out[base + k] += add;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *total;  // row[n] = number of arcs
}

// deg[v] = out-degree + in-degree (rows of the symmetrised structure of a directed graph)
__global__ void __launch_bounds__(BLOCK)
    csr_degree_sum_kernel(const uint32_t *row_a, const uint32_t *row_b, uint32_t n, uint32_t *deg) {
    const uint32_t v = blockIdx.x * BLOCK + threadIdx.x;
    if (v < n) deg[v] = (row_a[v + 1] - row_a[v]) + (row_b[v + 1] - row_b[v]);
    if (v == n) deg[v] = 0;
}

// Scatter arcs into their rows.  mode 0: key = src (successors); mode 1: key = dst (predecessors);
// mode 2: both directions into one structure (undirected graph, or symmetrised directed graph).
__global__ void __launch_bounds__(BLOCK)
    csr_scatter_kernel(const uint32_t *src, const uint32_t *dst, const uint32_t *eid,
                       uint64_t n_edges, int mode, const uint32_t *row, uint32_t *cursor,
                       uint32_t *col, uint32_t *arc_eid) {
    for (uint64_t e = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; e < n_edges;
         e += (uint64_t)gridDim.x * BLOCK) {
        const uint32_t u = src[e], v = dst[e];
        if (u == CSR_BAD) continue;
        const uint32_t id = eid ? eid[e] : (uint32_t)e;
        if (mode == 0 || mode == 2) {
            const uint32_t pos = row[u] + atomicAdd(&cursor[u], 1u);
            col[pos] = v;
            if (arc_eid) arc_eid[pos] = id;
        }
        if (mode == 1 || mode == 2) {
            const uint32_t pos = row[v] + atomicAdd(&cursor[v], 1u);
            col[pos] = u;
            if (arc_eid) arc_eid[pos] = id;
        }
    }
}

// Restore a deterministic arc order inside every row: ascending (original edge id, target).
// One warp per row: rows of up to 32 arcs are ranked in registers, rows of up to CSR_SORT_MID arcs
// are sorted by the warp in shared memory (BA(10^6,8) has 60 000 rows of 33..256 arcs: sending them
// all to the CTA-per-row kernel below made it 2.4 ms of a 5 ms snapshot), longer rows are queued for
// the CTA-per-row bitonic sort.
constexpr uint32_t CSR_SORT_MID = 256;
__global__ void __launch_bounds__(BLOCK)
    csr_sort_rows_kernel(const uint32_t *row, uint32_t n, uint32_t *col, uint32_t *arc_eid,
                         uint32_t *long_rows, uint32_t *n_long) {
    __shared__ unsigned long long s_mid[WARPS][CSR_SORT_MID];
    const uint32_t r = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (r >= n) return;
    const int lane = lane_id();
    const uint32_t e0 = row[r], e1 = row[r + 1], deg = e1 - e0;
    if (deg <= 1) return;
    if (deg > CSR_SORT_MID) {
        if (lane == 0) long_rows[atomicAdd(n_long, 1u)] = r;
        return;
    }
    if (deg > 32) {
        unsigned long long *s = s_mid[threadIdx.x >> 5];
        uint32_t N = 64;
        while (N < deg) N <<= 1;
        for (uint32_t i = lane; i < N; i += 32)
            s[i] = i < deg ? (((unsigned long long)arc_eid[e0 + i] << 32) | col[e0 + i]) : ~0ull;
        __syncwarp();
        for (uint32_t k = 2; k <= N; k <<= 1) {
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t i = lane; i < N; i += 32) {
                    const uint32_t ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned long long a = s[i], b = s[ixj];
                        if ((a > b) == ((i & k) == 0)) {
                            s[i] = b;
                            s[ixj] = a;
                        }
                    }
                }
                __syncwarp();
            }
        }
        for (uint32_t i = lane; i < deg; i += 32) {
            col[e0 + i] = (uint32_t)s[i];
            arc_eid[e0 + i] = (uint32_t)(s[i] >> 32);
        }
        return;
    }
    const bool ok = (uint32_t)lane < deg;
    const uint32_t c = ok ? col[e0 + lane] : 0xFFFFFFFFu;
    const uint32_t k = ok ? arc_eid[e0 + lane] : 0xFFFFFFFFu;
    const unsigned long long key = ((unsigned long long)k << 32) | c;
    uint32_t rank = 0;
    for (int j = 0; j < 32; j++) {
        const unsigned long long other = __shfl_sync(FULL, key, j);
        if ((uint32_t)j < deg && (other < key || (other == key && j < lane))) rank++;
    }
    if (ok) {
        col[e0 + rank] = c;
        arc_eid[e0 + rank] = k;
    }
}

// edge_id == NULL in rxc_graph_snapshot: the edge index is the position in the list
__global__ void __launch_bounds__(BLOCK) csr_iota_kernel(uint32_t *out, uint64_t len) {
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < len; i += (uint64_t)gridDim.x * BLOCK)
        out[i] = (uint32_t)i;
}

constexpr uint32_t CSR_SORT_MAX = 16384;  // keys a CTA sorts in shared memory (128 KB)

// Bitonic sort of one long row per CTA in shared memory.  Rows beyond CSR_SORT_MAX arcs keep the
// scatter order (their f64 sums are then order-nondeterministic, within the 1e-9 bar).
constexpr uint32_t CSR_SORT_THREADS = 1024;  // 128 KB of keys: one CTA per SM, so make it a full one
__global__ void __launch_bounds__(CSR_SORT_THREADS)
    csr_sort_long_kernel(const uint32_t *row, const uint32_t *long_rows, const uint32_t *n_long,
                         uint32_t *col, uint32_t *arc_eid) {
    extern __shared__ unsigned long long s_keys[];
    const uint32_t nl = *n_long;
    for (uint32_t li = blockIdx.x; li < nl; li += gridDim.x) {
        const uint32_t r = long_rows[li];
        const uint32_t e0 = row[r], deg = row[r + 1] - e0;
        if (deg > CSR_SORT_MAX) continue;
        uint32_t N = 64;
        while (N < deg) N <<= 1;
        for (uint32_t i = threadIdx.x; i < N; i += CSR_SORT_THREADS)
            s_keys[i] = i < deg ? (((unsigned long long)arc_eid[e0 + i] << 32) | col[e0 + i]) : ~0ull;
        __syncthreads();
        for (uint32_t k = 2; k <= N; k <<= 1) {
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t i = threadIdx.x; i < N; i += CSR_SORT_THREADS) {
                    const uint32_t ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned long long a = s_keys[i], b = s_keys[ixj];
                        const bool asc = (i & k) == 0;
                        if ((a > b) == asc) {
                            s_keys[i] = b;
                            s_keys[ixj] = a;
                        }
                    }
                }
                __syncthreads();
            }
        }
        for (uint32_t i = threadIdx.x; i < deg; i += CSR_SORT_THREADS) {
            col[e0 + i] = (uint32_t)s_keys[i];
            arc_eid[e0 + i] = (uint32_t)(s_keys[i] >> 32);
        }
        __syncthreads();
    }
}

}  // namespace rxc
