// bc_kernels.cuh -- Brandes betweenness (node + edge) as batched multi-source BFS on sm_100a.
//
// Replaces the per-source loops of the reference:
//   forward  BFS + sigma : rustworkx-core/src/centrality.rs:406-445 (node), :459-511 (edge)
//   backward dependency  : rustworkx-core/src/centrality.rs:254-299 (node), :301-327 (edge)
//
// Layout.  Sources are processed in GROUPS of 128.  For a group g every
// per-(vertex, source) f64 lives in one 1 KB row  sc[g][v][0..127]; warp lane l owns the 32-byte
// sector sc[g][v][4l..4l+3], i.e. FOUR sources (slot k of lane l is source 32k+l of the group).
// The irregular access of a BFS -- "for each arc (v,w) read something of v" -- therefore moves whole
// DRAM sectors, one per lane, and a lane only issues its load when one of its four sources needs
// the arc.  Which sources need it is decided from 128-bit masks (word k, bit l <-> lane l, slot k):
//   visited[g][v]   source has reached v
//   P[par][g][v]    {tag, mask}: sources for which v sits on the level the tag names (one 32-byte
//                   entry = one sector per look-up; the tag makes clearing unnecessary)
// Every level appends its (vertex, mask) pairs to a per-group queue; the backward sweep replays the
// queue levels in reverse and pulls from successors (no atomics on the f64 rows).
//
// Three users share this layout: betweenness proper (forward + backward), group betweenness (two
// slots per source in one forward pass, end of this file) and the distance-only BFS of shallow
// graphs (bfs128_level_kernel: the forward kernel without rows).  By default the forward pass keeps
// sigma as exact uint32 path counts in 512-byte rows ("narrow sigma rows" below) and falls back to
// the f64 rows described here when a count overflows.
//
// Why 128: ncu on the 32-source version showed ~40 issued instructions per gathered row and half
// of all issue slots busy; with four sources per lane the per-arc look-up, shuffles and address
// arithmetic are amortised over 4x the sources and every load carries 32 bytes per lane.
// Vertices above `hub_degree` are deferred to a CTA-per-vertex kernel so that one warp never
// serialises a hub's thousands of arcs.
#pragma once

#include "common.cuh"

namespace rxc {

// Group width.  RXC_BK = sources per lane: 4 (groups of 128: a lane's f64 values are one 32-byte
// sector) or 8 (groups of 256: two sectors of f64, one of uint32 path counts).  Wider groups gather
// fewer rows, look up fewer level masks and queue fewer entries per source at the same sector bytes
// (scripts/sim_group_width.py: -43 % / -46 % / -47 % on BA graphs) for more registers per thread.
#ifndef RXC_BK
#define RXC_BK 4
#endif
constexpr int BK = RXC_BK;   // sources per lane
constexpr int BG = 32 * BK;  // sources per betweenness group
constexpr int MQ = BK / 4;   // 16-byte quads per source mask
static_assert(BK == 4 || BK == 8, "RXC_BK must be 4 or 8");

struct __align__(16) SrcMask {  // one bit per source of a group: word k, bit l <-> lane l, slot k
    uint32_t w[BK];
};

struct __align__(8 * BK) PEntry {  // 32 bytes (BK = 4) / 64 bytes (BK = 8)
    uint32_t tag;
    uint32_t pad[BK - 1];
    SrcMask m;
};

__device__ __forceinline__ SrcMask mask_zero() {
    SrcMask r;
#pragma unroll
    for (int k = 0; k < BK; k++) r.w[k] = 0;
    return r;
}
__device__ __forceinline__ uint32_t mask_or(const SrcMask &a) {
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < BK; k++) r |= a.w[k];
    return r;
}
__device__ __forceinline__ bool mask_any(const SrcMask &a) { return mask_or(a) != 0; }
__device__ __forceinline__ uint32_t mask_popc(const SrcMask &a) {
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < BK; k++) r += __popc(a.w[k]);
    return r;
}
__device__ __forceinline__ bool mask_eq(const SrcMask &a, const SrcMask &b) {
    uint32_t d = 0;
#pragma unroll
    for (int k = 0; k < BK; k++) d |= a.w[k] ^ b.w[k];
    return d == 0;
}
__device__ __forceinline__ SrcMask mask_load(const SrcMask *p) {
    SrcMask r;
#pragma unroll
    for (int q = 0; q < MQ; q++) {
        const uint4 t = reinterpret_cast<const uint4 *>(p)[q];
        r.w[4 * q] = t.x;
        r.w[4 * q + 1] = t.y;
        r.w[4 * q + 2] = t.z;
        r.w[4 * q + 3] = t.w;
    }
    return r;
}
__device__ __forceinline__ void mask_store(SrcMask *p, const SrcMask &a) {
#pragma unroll
    for (int q = 0; q < MQ; q++)
        reinterpret_cast<uint4 *>(p)[q] = make_uint4(a.w[4 * q], a.w[4 * q + 1], a.w[4 * q + 2], a.w[4 * q + 3]);
}
__device__ __forceinline__ SrcMask mask_shfl(const SrcMask &a, int src) {
    SrcMask r;
#pragma unroll
    for (int k = 0; k < BK; k++) r.w[k] = __shfl_sync(FULL, a.w[k], src);
    return r;
}

struct BcParams {
    const uint32_t *row;  // CSR offsets of the adjacency this kernel walks
    const uint32_t *col;  // CSR targets (compact vertex ids)
    const uint32_t *eid;  // arc -> original edge index (edge betweenness only)
    const uint32_t *srow; // successor CSR (frontier marking in the forward pass)
    const uint32_t *scol;
    uint32_t *maybe;      // [NG][mwords] bit v: v has a neighbour on the current frontier (sparse levels)
    double *sc;           // [NG][n][128] sigma on the way down, (1+delta)/sigma on the way up
    uint32_t *sig32;      // [NG][n][128] narrow mode: sigma as exact 32-bit path counts (sc holds only
                          // (1+delta)/sigma then); lane l owns the 16-byte chunk [4l..4l+3]
    uint32_t *overflow;   // bit 0: a path count does not fit 32 bits (narrow mode); bit 1: a level
                          // queue outgrew qcap (the host redoes the batch with worst-case queues)
    PEntry *P0;           // [NG][n] tagged level masks, even levels
    PEntry *P1;           // [NG][n] tagged level masks, odd levels
    SrcMask *visited;     // [NG][n]
    uint32_t *qv;         // [NG][qcap] level queues: vertex
    SrcMask *qm;          // [NG][qcap] level queues: source mask
    uint2 *qe;            // [NG][qcap] level queues: the vertex's arc range in the SUCCESSOR structure, so
                          // that the dependency sweep starts its arc scan without a dependent look-up
    uint32_t *loff;       // [NG][lcap] queue offset of each level
    uint32_t *lvlcnt;     // [NG][lcap] entries of each level
    uint32_t *lvl_total;  // [lcap] entries of each level summed over groups
    uint32_t *hubv;       // [NG][hubcap] deferred high-degree work items: vertex
    SrcMask *hubm;        // [NG][hubcap] deferred high-degree work items: mask
    uint32_t *hubcnt;     // [2][NG] number of deferred items, by level parity
    unsigned long long *reach;  // [NG][128] vertices reached per source, excluding the source
    unsigned long long *stats;  // [ST_COUNT]
    double *acc;          // node mode: [n] by compact vertex; edge mode: [e_bound] by edge index
    const uint32_t *src;  // [NG][128] compact source ids, NO_SOURCE padded; slot 32k+l = lane l, k
    const uint32_t *blocked;  // group betweenness: bitmap of the group's vertices, else nullptr
    unsigned long long *dsum;   // distance-only mode: [NG][128] sum of distances per source
    double *drows;              // distance-only mode: [NG*128][n_cols] distance rows, or nullptr
    uint64_t n_cols;
    uint64_t qcap;
    uint32_t n;
    uint32_t lcap;
    uint32_t hubcap;
    uint32_t hub_degree;  // arcs; above this a vertex is handled by a whole CTA
    uint32_t mwords;      // words per group in `maybe`: ceil(n / 32)
    uint32_t sparse_max;  // a level with at most this many queue entries is filtered through `maybe`
    uint32_t ngroups;     // groups in this batch (stride of hubcnt)
    uint32_t tagbase;     // tag of level l is tagbase + l + 1
};

__device__ __forceinline__ PEntry *bc_P(const BcParams &p, uint32_t level) {
    return (level & 1) ? p.P1 : p.P0;
}

__device__ __forceinline__ uint32_t *bc_hubcnt(const BcParams &p, uint32_t level, uint32_t g) {
    return p.hubcnt + (size_t)(level & 1) * p.ngroups + g;
}

// Mask of vertex v if it sits on the level `tag` names, else empty; one sector per look-up.
__device__ __forceinline__ SrcMask bc_level_mask(const PEntry *P, uint32_t v, uint32_t tag) {
    const uint4 *q = reinterpret_cast<const uint4 *>(&P[v]);
    const uint4 t = q[0];
    SrcMask m = mask_load(reinterpret_cast<const SrcMask *>(q + MQ));  // the mask sits BK words in
    const bool ok = t.x == tag;
#pragma unroll
    for (int k = 0; k < BK; k++) m.w[k] = ok ? m.w[k] : 0u;
    return m;
}

__device__ __forceinline__ void bc_store_level(PEntry *P, uint32_t v, uint32_t tag,
                                               const SrcMask &m) {
    uint4 *q = reinterpret_cast<uint4 *>(&P[v]);
    q[0] = make_uint4(tag, 0, 0, 0);
    mask_store(reinterpret_cast<SrcMask *>(q + MQ), m);
}

// Defer a high-degree work item (warp-uniform v, mask) to the CTA-per-vertex kernel; called by
// the whole warp, appended once by lane 0.  False when the list is full.
__device__ __forceinline__ bool bc_defer_hub(const BcParams &p, uint32_t level, uint32_t g,
                                             uint32_t v, const SrcMask &mask) {
    int ok = 0;
    if (lane_id() == 0) {
        const uint32_t h = atomicAdd(bc_hubcnt(p, level, g), 1u);
        if (h < p.hubcap) {
            p.hubv[(size_t)g * p.hubcap + h] = v;
            mask_store(&p.hubm[(size_t)g * p.hubcap + h], mask);
            ok = 1;
        }
    }
    return __shfl_sync(FULL, ok, 0) != 0;
}

__device__ __forceinline__ SrcMask bc_valid_mask(const BcParams &p, uint32_t g) {
    SrcMask valid;
#pragma unroll
    for (int k = 0; k < BK; k++)
        valid.w[k] = __ballot_sync(FULL, p.src[(size_t)g * BG + k * 32 + lane_id()] != NO_SOURCE);
    return valid;
}

// visited[g][*] = ~valid(g): slots without a source count as "reached everywhere".
__global__ void __launch_bounds__(BLOCK) bc_init_kernel(BcParams p) {
    const uint32_t g = blockIdx.y;
    const SrcMask valid = bc_valid_mask(p, g);
    const uint32_t w = blockIdx.x * BLOCK + threadIdx.x;
    if (w < p.n) {
        SrcMask nv;
#pragma unroll
        for (int k = 0; k < BK; k++) nv.w[k] = ~valid.w[k];
        mask_store(&p.visited[(size_t)g * p.n + w], nv);
    }
}

// Level 0: one warp per group seeds its (up to) 128 sources.
template <bool NARROW>
__global__ void __launch_bounds__(32) bc_seed_kernel(BcParams p) {
    const uint32_t g = blockIdx.x;
    const int lane = lane_id();
    const SrcMask valid = bc_valid_mask(p, g);
    uint32_t base = 0;
#pragma unroll
    for (int k = 0; k < BK; k++) {
        const uint32_t s = p.src[(size_t)g * BG + k * 32 + lane];
        if (s != NO_SOURCE) {
            const size_t gv = (size_t)g * p.n + s;
            SrcMask bit = mask_zero();
            bit.w[k] = 1u << lane;
            SrcMask vis;
#pragma unroll
            for (int j = 0; j < BK; j++) vis.w[j] = ~valid.w[j] | bit.w[j];
            mask_store(&p.visited[gv], vis);  // sources of a group are distinct vertices
            bc_store_level(p.P0 + (size_t)g * p.n, s, p.tagbase + 1, bit);
            if (NARROW) p.sig32[gv * BG + lane * BK + k] = 1u;
            else p.sc[gv * BG + lane * BK + k] = 1.0;
            const uint32_t i = base + __popc(valid.w[k] & lanemask_lt());
            p.qv[(size_t)g * p.qcap + i] = s;
            mask_store(&p.qm[(size_t)g * p.qcap + i], bit);
            p.qe[(size_t)g * p.qcap + i] = make_uint2(p.srow[s], p.srow[s + 1]);
        }
        base += __popc(valid.w[k]);
        p.reach[(size_t)g * BG + k * 32 + lane] = 0ull;
    }
    if (lane == 0) {
        p.loff[(size_t)g * p.lcap] = 0;
        p.lvlcnt[(size_t)g * p.lcap] = base;
        p.hubcnt[g] = 0;
        p.hubcnt[p.ngroups + g] = 0;
        atomicAdd(&p.lvl_total[0], base);
    }
}

// ---------------------------------------------------------------------------------------------
// row gathers: lane l owns the BK consecutive values [BK*l, BK*l + BK) of every row
// ---------------------------------------------------------------------------------------------
struct RowK {
    double x[BK];
};

__device__ __forceinline__ RowK row_load(const double *rows, uint32_t v, int lane, bool on) {
    RowK r;
#pragma unroll
    for (int k = 0; k < BK; k++) r.x[k] = 0.0;
    if (on) {
        const double2 *q = reinterpret_cast<const double2 *>(rows + (size_t)v * BG + lane * BK);
#pragma unroll
        for (int c = 0; c < BK / 2; c++) {
            const double2 a = q[c];
            r.x[2 * c] = a.x;
            r.x[2 * c + 1] = a.y;
        }
    }
    return r;
}

// Staging area: every warp owns a few row slots in (dynamic) shared memory and fills them with
// cp.async (LDGSTS, 16-byte, L1-bypassing), so several KB per warp are in flight without holding
// registers; sectors none of whose sources needs the arc are not fetched.  Measured on BA(1M,8),
// 1024 sources, groups of 128, f64 rows (GTEPS): STAGE 8 x 3 CTAs/SM 169, 4 x 4 176, 4 x 5 172,
// 4 x 6 162, 2 x 6 165, 2 x 8 148 with the first slot layout; the same ranking with the
// conflict-free layout below (4 x 4: 204).
#ifndef RXC_STAGE
#define RXC_STAGE 4
#endif
#ifndef RXC_MINBLOCKS
#define RXC_MINBLOCKS (BK == 4 ? 4 : 2)
#endif
#ifndef RXC_FWD_MINBLOCKS
#define RXC_FWD_MINBLOCKS (BK == 4 ? 4 : 2)
#endif
#ifndef RXC_FWD32_MINBLOCKS
#define RXC_FWD32_MINBLOCKS (BK == 4 ? 5 : 3)  // BK = 4: 46 registers, no spills: 28.0 -> 25.9 ms per 1024 sources
#endif
#ifndef RXC_FWD32_SLOTS
#define RXC_FWD32_SLOTS 6  // BK = 4: 6 slots + the arc table keep the narrow forward kernel at 5 CTAs per SM
#endif
// Arc table: the arcs a look-up found on the frontier are compacted into a small per-warp table in
// shared memory -- (vertex, sector-need) and the source mask per arc -- and the staging loops walk
// that dense list with broadcast LDS reads.  ncu on the first form (arc a of the look-up lives in lane
// a: find-first-set on the ballot, clear the bit, 2 shuffles to stage the row, 4 more to fetch its
// mask when the row is consumed) counted 78 issued instructions per gathered row with the issue slots
// 80 % busy in the big forward launches; the table removes the bit scans and 6 of the 6.2 shuffles
// per row.
constexpr int ARCTAB_BYTES = 32 * 8 + (BK == 8 ? 32 * 4 : 0) + 32 * 16 * MQ;  // per warp
constexpr int STAGE = RXC_STAGE;  // f64 row slots per warp in the forward / edge / weighted kernels
constexpr int BC_SMEM_BYTES = WARPS * (STAGE * BG * 8 + ARCTAB_BYTES);  // BK = 4: 32 KB + tables per CTA
constexpr int FWD32_SLOTS = RXC_FWD32_SLOTS;  // uint32 row slots per warp in the narrow forward kernels
constexpr int BC_FWD32_SMEM_BYTES = WARPS * (FWD32_SLOTS * BG * 4 + ARCTAB_BYTES);
// The dependency kernels have no static shared memory, so 6 row slots per warp still fit 4 CTAs per
// SM (4 x 48 KB) with groups of 128: measured 45.3 -> 43.3 ms per 1024 sources; the forward kernels
// (10 KB static) would drop to 3 CTAs per SM with 6 slots and lose 3 ms.
#ifndef RXC_BSTAGE
#define RXC_BSTAGE (BK == 4 ? 6 : 5)
#endif
constexpr int BSTAGE = RXC_BSTAGE;
constexpr int BC_BWD_SMEM_BYTES = WARPS * (BSTAGE * BG * 8 + ARCTAB_BYTES);
// edge mode: rows in flight per warp, and arcs per pass of the cross-lane exchange network
#ifndef RXC_ESTAGE
#define RXC_ESTAGE BSTAGE
#endif
constexpr int ESTAGE = RXC_ESTAGE;
constexpr int ER = 4;
static_assert(ESTAGE <= BSTAGE, "edge mode stages into the dependency kernels' slots");

// Geometry of one staged row of BG elements: CPL 16-byte chunks per lane (1, 2 or 4), i.e. CPL
// warp-wide copy instructions whose global side is 512 contiguous bytes each.
//   CPL = 1 (uint32 rows, BK = 4): a lane copies and reads back its own chunk; nothing to swizzle.
//   CPL = 2 (f64 rows, BK = 4; uint32 rows, BK = 8): a lane's values are one 32-byte sector.
//   CPL = 4 (f64 rows, BK = 8): a lane's values are two sectors (slots 0..3 and 4..7).
// Copy instruction j moves chunk 32 j + lane; lane l reads back chunks CPL l .. CPL l + CPL - 1.
// Chunk c is stored at position (c & ~7) | ((c + (c >> 3)) & 7) -- a rotation inside its 128-byte
// line by the line number -- which makes the 8 lanes of every quarter warp hit 8 different 16-byte
// bank groups for the writes (8 consecutive chunks of one line) and for every read (CPL = 2: 8 even
// or 8 odd chunks of two neighbouring lines; CPL = 4: two chunks 4 apart in each of four
// consecutive lines).  ncu on the first layout (lane l copying its own chunks): 12.4 shared
// wavefronts per LDGSTS against an ideal of 4-5, shared-memory pipe at 65-85 % in the big launches.
// A chunk is only fetched when its SECTOR is needed (CPL = 1: its half sector); stale bytes of
// skipped chunks are never used, because a lane only adds values whose mask bit is set.
template <class Elem>
struct RowGeom {
    static constexpr int ROW_BYTES = BG * (int)sizeof(Elem);
    static constexpr int CPL = ROW_BYTES / 512;
    static constexpr int SPL = CPL >= 2 ? CPL / 2 : 1;  // sector-need words per arc (see row_need)
    static constexpr int VPC = 16 / (int)sizeof(Elem);  // values per chunk
    static_assert(CPL == 1 || CPL == 2 || CPL == 4, "row geometry");
};

__device__ __forceinline__ int stage_pos(int c) { return (c & ~7) | ((c + (c >> 3)) & 7); }

// Which lanes' sectors an arc with source mask m needs: need[h] bit l <-> sector h of lane l.
// RXC_NEEDLINES = 4 or 8 (experiment): round the need up to whole groups of that many
// lanes -- 4 lanes of an f64 row / 8 lanes of a uint32 row are one 128-byte line.
#ifndef RXC_NEEDLINES
#define RXC_NEEDLINES 0
#endif
template <int SPL>
__device__ __forceinline__ void row_need(const SrcMask &m, uint32_t (&need)[SPL]) {
#pragma unroll
    for (int h = 0; h < SPL; h++) {
        need[h] = 0;
#pragma unroll
        for (int k = h * (BK / SPL); k < (h + 1) * (BK / SPL); k++) need[h] |= m.w[k];
#if RXC_NEEDLINES == 4
        uint32_t t = need[h] | need[h] >> 1 | need[h] >> 2 | need[h] >> 3;
        need[h] = (t & 0x11111111u) * 15u;
#elif RXC_NEEDLINES == 8
        uint32_t t = need[h] | need[h] >> 1;
        t |= t >> 2;
        t |= t >> 4;
        need[h] = (t & 0x01010101u) * 255u;
#elif RXC_NEEDLINES == 32
        need[h] = need[h] ? 0xFFFFFFFFu : 0u;
#endif
    }
}

// lane_rows = (const char *)rows + 16 * lane: this lane's chunk of row 0, formed once by the caller,
// so that a row's address is one multiply-add on it (ncu: the base used to be re-assembled from two
// partial 64-bit sums for every row)
template <class Elem>
__device__ __forceinline__ void cp_async_row(uint32_t slot_smem, int lane, const char *lane_rows, uint32_t v,
                                             const uint32_t (&need)[RowGeom<Elem>::SPL]) {
    constexpr int CPL = RowGeom<Elem>::CPL;
    const char *grow = lane_rows + (size_t)v * RowGeom<Elem>::ROW_BYTES;
#pragma unroll
    for (int j = 0; j < CPL; j++) {
        const int c = 32 * j + lane;
        uint32_t on, dst;
        if (CPL == 1) {
            on = (need[0] >> lane) & 1u;
            dst = slot_smem + 16u * lane;
        } else if (CPL == 2) {
            on = (need[0] >> (c >> 1)) & 1u;
            dst = slot_smem + 16u * stage_pos(c);
        } else {
            on = ((((c >> 1) & 1) ? need[RowGeom<Elem>::SPL - 1] : need[0]) >> (c >> 2)) & 1u;
            dst = slot_smem + 16u * stage_pos(c);
        }
        asm volatile(
            "{\n"
            ".reg .pred p1;\n"
            "setp.ne.u32 p1, %2, 0;\n"
            "@p1 cp.async.cg.shared.global [%0], [%1], 16;\n"
            "}\n" ::"r"(dst),
            "l"(grow + 512 * j), "r"(on)
            : "memory");
    }
}

// the BK values of this lane out of a staged row
template <class Elem>
__device__ __forceinline__ void stage_read(const char *slot, int lane, Elem (&x)[BK]) {
    constexpr int CPL = RowGeom<Elem>::CPL, VPC = RowGeom<Elem>::VPC;
#pragma unroll
    for (int i = 0; i < CPL; i++) {
        const int c = CPL * lane + i;
        const char *at = slot + 16 * (CPL == 1 ? c : stage_pos(c));
        if constexpr (VPC == 4) {
            const uint4 q = *reinterpret_cast<const uint4 *>(at);
            x[4 * i] = q.x;
            x[4 * i + 1] = q.y;
            x[4 * i + 2] = q.z;
            x[4 * i + 3] = q.w;
        } else {
            const double2 q = *reinterpret_cast<const double2 *>(at);
            x[2 * i] = q.x;
            x[2 * i + 1] = q.y;
        }
    }
}

__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.wait_all;\n" ::: "memory");
}

// acc[k] += rows[v][lane][k] over the arcs in nz (lanes of v_lane / m_lane with a non-empty mask)
// for the sources the arc's mask names, arc order preserved; up to NS rows in flight per warp.
// Per-row bookkeeping is kept out of this loop (ncu: ~80 issued instructions per gathered row made
// the big launches co-limited by issue slots): the union of the masks and the DAG-arc count are
// formed once per look-up from the lanes' own masks by the callers.
// Batches of NS rows.  (A ring that tops itself up row by row, so that NS rows stay in flight for
// the whole arc list, was measured: 80 -> 86 ms per 1024 sources -- its bookkeeping costs more issue
// slots than the shorter waits give back.)
// What a gathered value does to the running value of its source: betweenness sums (sigma, then
// (1+delta)/sigma); the weighted shortest paths take min(dist + cost of the arc) -- the arc's cost is a
// per-arc scalar that travels through the arc table.
struct GatherSum {
    static constexpr bool HAS_AUX = false;
    template <class Acc, class Elem>
    static __device__ __forceinline__ void apply(Acc &acc, Elem x, double) { acc += x; }
};
struct GatherMinPlus {
    static constexpr bool HAS_AUX = true;
    static __device__ __forceinline__ void apply(double &acc, double x, double c) { acc = fmin(acc, x + c); }
};
constexpr int ARCTAB_AUX_BYTES = 32 * 8;  // per warp, behind the arc table (GatherMinPlus only)

template <int NS, class Elem, class Acc, class Op = GatherSum>
__device__ __forceinline__ void bc_gather(const Elem *rows, uint32_t v_lane, const SrcMask &m_lane,
                                          uint32_t nz, int lane, char *stage, Acc (&acc)[BK],
                                          double aux_lane = 0.0) {
    using G = RowGeom<Elem>;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stage);
    const char *lane_rows = reinterpret_cast<const char *>(rows) + 16 * lane;
    asm("" : "+l"(lane_rows));  // keep it one 64-bit value (else the compiler re-adds a uniform part per row)
    uint32_t need_lane[G::SPL];
    row_need<G::SPL>(m_lane, need_lane);
    // the table sits behind this warp's NS row slots: vn[32] = (vertex, need[0]), [need1[32]], m[32]
    char *tab = stage + NS * G::ROW_BYTES;
    uint2 *t_vn = reinterpret_cast<uint2 *>(tab);
    uint32_t *t_need1 = reinterpret_cast<uint32_t *>(tab + 32 * 8);
    SrcMask *t_m = reinterpret_cast<SrcMask *>(tab + 32 * 8 + (BK == 8 ? 32 * 4 : 0));
    double *t_aux = reinterpret_cast<double *>(tab + ARCTAB_BYTES);
    if (nz == 0) return;  // warp-uniform
    const int total = __popc(nz);
#if RXC_PROBE
    {
        // 32-byte sectors and 128-byte lines this look-up's row gathers touch
        const uint32_t nd = ((nz >> lane) & 1u) ? need_lane[0] : 0u;
        uint32_t sectors, lines;
        if (sizeof(Elem) == 8) {  // lane = one sector, 4 lanes = one line
            sectors = __popc(nd);
            lines = __popc((nd | nd >> 1 | nd >> 2 | nd >> 3) & 0x11111111u);
        } else {  // lane = half a sector, 8 lanes = one line
            sectors = __popc((nd | nd >> 1) & 0x55555555u);
            uint32_t f = nd | nd >> 1;
            f |= f >> 2;
            f |= f >> 4;
            lines = __popc(f & 0x01010101u);
        }
        sectors = __reduce_add_sync(FULL, sectors);
        lines = __reduce_add_sync(FULL, lines);
        if (lane == 0) {
            const int o = sizeof(Elem) == 8 ? 4 : 0;
            atomicAdd(&g_probe[PR_ROWS32 + o], (unsigned long long)total);
            atomicAdd(&g_probe[PR_SECTORS32 + o], (unsigned long long)sectors);
            atomicAdd(&g_probe[PR_LINES32 + o], (unsigned long long)lines);
        }
    }
#endif
    if ((nz >> lane) & 1u) {
        const int pos = __popc(nz & lanemask_lt());
        t_vn[pos] = make_uint2(v_lane, need_lane[0]);
        if (G::SPL == 2) t_need1[pos] = need_lane[G::SPL - 1];
        mask_store(&t_m[pos], m_lane);
        if (Op::HAS_AUX) t_aux[pos] = aux_lane;
    }
    __syncwarp();
    for (int j0 = 0; j0 < total; j0 += NS) {
        const int cnt = min(NS, total - j0);
        // Straight-line blocks of 2 or 3 slots: the table reads of a block, then its predicated copies;
        // a slot past the end of the list re-reads the last table entry with every lane switched
        // off.  (ptxas puts three padding LDS in front of every LDGSTS that follows an LDS or opens a
        // basic block; with a branch and a table read per slot that was 6 of the ~45 issue slots a
        // gathered row costs.)
        constexpr int NB = NS % 3 == 0 ? 3 : (NS % 2 == 0 ? 2 : 1);  // slots per block
#pragma unroll
        for (int b = 0; b < NS; b += NB) {
            if (b < cnt) {  // warp-uniform
                uint2 e[NB];
                uint32_t need1[NB];
#pragma unroll
                for (int i = 0; i < NB; i++) {
                    const int j = min(j0 + b + i, total - 1);
                    e[i] = t_vn[j];  // broadcast reads, all of them before the first copy (see above)
                    if (G::SPL == 2) need1[i] = t_need1[j];
                }
#pragma unroll
                for (int i = 0; i < NB; i++) {
                    uint32_t need[G::SPL];
                    need[0] = b + i < cnt ? e[i].y : 0u;
                    if (G::SPL == 2) need[G::SPL - 1] = b + i < cnt ? need1[i] : 0u;
                    cp_async_row<Elem>(sbase + (b + i) * G::ROW_BYTES, lane, lane_rows, e[i].x, need);
                }
            }
        }
        cp_async_wait_all();
        if (G::CPL > 1) __syncwarp();  // the chunks a lane reads were copied by other lanes
#pragma unroll
        for (int i = 0; i < NS; i++) {
            if (i < cnt) {
                const SrcMask mj = mask_load(&t_m[j0 + i]);  // broadcast read
                const double aux = Op::HAS_AUX ? t_aux[j0 + i] : 0.0;
                Elem x[BK];
                stage_read<Elem>(stage + i * G::ROW_BYTES, lane, x);
#pragma unroll
                for (int k = 0; k < BK; k++)
                    if ((mj.w[k] >> lane) & 1u) Op::apply(acc[k], x[k], aux);
            }
        }
        if (G::CPL > 1) __syncwarp();  // the slots may be refilled by other lanes from here on
    }
    __syncwarp();  // every lane is done with the table before the next look-up rewrites it
}

// ---------------------------------------------------------------------------------------------
// narrow sigma rows
//
// sigma counts shortest paths: exact integers.  On small-world graphs they are tiny (BA(1M,8): below
// 10^5), so the forward pass keeps them as uint32 -- half the row bytes, 8 sources per DRAM sector
// instead of 4 -- and accumulates in uint64.  That halves the bytes, the LDGSTS/LDS wavefronts and
// the copy instructions of every gathered row, and leaves room for twice the rows in flight per
// warp in the same shared memory.  The dependency sweep needs (1+delta)/sigma in f64 and writes it
// to the f64 rows.  A path count that does not fit 32 bits (lattices: binomially many paths) raises
// `overflow`; the host then redoes the batch with f64 sigma rows (the "wide" mode, which is the
// layout described at the top of this file) and stays wide for the handle.
// ---------------------------------------------------------------------------------------------
template <bool NARROW>
struct SigmaRows;
template <>
struct SigmaRows<false> {
    using Elem = double;
    using Acc = double;
    static constexpr int SLOTS = STAGE;
    static __device__ __forceinline__ Elem *rows(const BcParams &p, size_t gn) { return p.sc + gn * BG; }
};
template <>
struct SigmaRows<true> {
    using Elem = uint32_t;
    using Acc = unsigned long long;
    static constexpr int SLOTS = FWD32_SLOTS;  // half-size rows: more slots in the same shared memory
    static __device__ __forceinline__ Elem *rows(const BcParams &p, size_t gn) { return p.sig32 + gn * BG; }
};

// sigma of a newly discovered (vertex, source); narrow mode flags counts beyond 32 bits
template <bool NARROW>
__device__ __forceinline__ void bc_store_sigma(const BcParams &p, typename SigmaRows<NARROW>::Elem *dst,
                                               typename SigmaRows<NARROW>::Acc v) {
    if (NARROW) {
        if ((unsigned long long)v >> 32) atomicOr(p.overflow, 1u);
        *dst = (typename SigmaRows<NARROW>::Elem)v;
    } else {
        *dst = (typename SigmaRows<NARROW>::Elem)v;
    }
}

// own sigma of the BK sources of a lane, as f64
template <bool NARROW>
__device__ __forceinline__ RowK bc_load_sigma(const BcParams &p, size_t gn, uint32_t v, int lane, bool on) {
    if (NARROW) {
        RowK r;
#pragma unroll
        for (int k = 0; k < BK; k++) r.x[k] = 0.0;
        if (on) {
            const uint4 *q = reinterpret_cast<const uint4 *>(p.sig32 + (gn + v) * BG + lane * BK);
#pragma unroll
            for (int c = 0; c < MQ; c++) {
                const uint4 t = q[c];
                r.x[4 * c] = (double)t.x;
                r.x[4 * c + 1] = (double)t.y;
                r.x[4 * c + 2] = (double)t.z;
                r.x[4 * c + 3] = (double)t.w;
            }
        }
        return r;
    }
    return row_load(p.sc + gn * BG, v, lane, on);
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------

// Sparse levels: when a group's frontier is small, only vertices adjacent to it can be discovered.
// One warp per frontier entry sets the bit of every successor in `maybe`; the forward kernel then
// skips every vertex whose bit is clear (and clears the words it consumed), so the cost of an early
// BFS level -- or of any level of a long, thin graph -- is proportional to the frontier instead of n.
__global__ void __launch_bounds__(BLOCK) bc_mark_kernel(BcParams p, uint32_t level) {
    const uint32_t g = blockIdx.y;
    const uint32_t cnt = p.lvlcnt[(size_t)g * p.lcap + level];
    if (cnt == 0 || cnt > p.sparse_max) return;
    const uint32_t off = p.loff[(size_t)g * p.lcap + level];
    if ((uint64_t)off + cnt > p.qcap) return;  // queue overflow (flagged by the forward kernel): batch is redone
    const uint32_t *qv = p.qv + (size_t)g * p.qcap + off;
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


// One warp: scan the in-arcs of a vertex (chunks of 32 arcs starting at e0+first, stride `step`),
// pick up the sources whose frontier holds the neighbour, and sum the neighbours' sigma rows.
template <class Elem, class Acc>
__device__ __forceinline__ void bc_fwd_scan(const BcParams &p, const Elem *rows, const PEntry *Pcur,
                                            uint32_t tag_cur, uint32_t e0, uint32_t e1,
                                            uint32_t first, uint32_t step, const SrcMask &unv,
                                            int lane, char *stage, Acc (&acc)[BK],
                                            SrcMask &newm) {
    for (uint32_t e = e0 + first; e < e1; e += step) {
        const uint32_t idx = e + lane;
        uint32_t v = 0;
        SrcMask m = mask_zero();
        if (idx < e1) {
            v = p.col[idx];
            m = bc_level_mask(Pcur, v, tag_cur);
#pragma unroll
            for (int k = 0; k < BK; k++) m.w[k] &= unv.w[k];
        }
        const uint32_t nz = __ballot_sync(FULL, mask_any(m));
#if RXC_PROBE
        if (lane == 0) atomicAdd(&g_probe[sizeof(Elem) == 8 ? PR_LOOKUPS64 : PR_LOOKUPS32], (unsigned long long)min(32u, e1 - e));
#endif
        if (nz) {
#pragma unroll
            for (int k = 0; k < BK; k++) newm.w[k] |= __reduce_or_sync(FULL, m.w[k]);
            bc_gather<SigmaRows<sizeof(Elem) == 4>::SLOTS>(rows, v, m, nz, lane, stage, acc);
        }
    }
}

// One BFS level for every group of the batch: vertices not yet reached by all 128 sources scan
// their in-arcs -- the bottom-up form of centrality.rs:427-437.  No early exit: every frontier
// parent contributes to sigma.
template <bool NARROW>
__global__ void __launch_bounds__(BLOCK, (NARROW ? RXC_FWD32_MINBLOCKS : RXC_FWD_MINBLOCKS)) bc_forward_kernel(BcParams p, uint32_t level) {
    using Acc = typename SigmaRows<NARROW>::Acc;
    const uint32_t g = blockIdx.y;
    uint32_t *lc = p.lvlcnt + (size_t)g * p.lcap;
    uint32_t *lo = p.loff + (size_t)g * p.lcap;
    if (blockIdx.x == 0 && threadIdx.x == 0) *bc_hubcnt(p, level + 1, g) = 0;  // for the next launch
    const uint32_t fcnt = lc[level];
    if (fcnt == 0) return;  // this group's BFS has finished

    __shared__ SrcMask s_unv[BLOCK], s_om[BLOCK];
    __shared__ uint32_t s_w[BLOCK], s_ov[BLOCK];
    __shared__ uint2 s_e[BLOCK];  // arc range of every candidate, read coalesced while compacting
    __shared__ uint32_t s_ncand, s_next, s_nout, s_base;
    extern __shared__ __align__(16) char s_dyn_stage[];
    char *stage = s_dyn_stage + (size_t)(threadIdx.x >> 5) * ((NARROW ? BC_FWD32_SMEM_BYTES : BC_SMEM_BYTES) / WARPS);
    const int tid = threadIdx.x, lane = lane_id();
    if (tid == 0) {
        s_ncand = 0;
        s_next = 0;
        s_nout = 0;
    }
    __syncthreads();

    const size_t gn = (size_t)g * p.n;
    SrcMask *visited = p.visited + gn;
    const PEntry *Pcur = bc_P(p, level) + gn;
    PEntry *Pnext = bc_P(p, level + 1) + gn;
    auto *rows = SigmaRows<NARROW>::rows(p, gn);
    const uint32_t tag_cur = p.tagbase + level + 1, tag_next = tag_cur + 1;

    // 1. coalesced read of this CTA's 256 visited masks; compact the unfinished vertices
    {
        const uint32_t w = blockIdx.x * BLOCK + tid;
        SrcMask unv = mask_zero();
        bool look = w < p.n;
        if (fcnt <= p.sparse_max) {  // sparse level: only neighbours of the frontier can be reached
            uint32_t *word = p.maybe + (size_t)g * p.mwords + (w >> 5);
            const uint32_t bits = (w < p.n) ? *word : 0u;  // one word per warp (256 | 32)
            look = (bits >> lane) & 1u;
            __syncwarp();
            if (lane == 0 && bits) *word = 0;  // consumed: clean for the next level
        }
        if (look) {
            const SrcMask vis = mask_load(&visited[w]);
#pragma unroll
            for (int k = 0; k < BK; k++) unv.w[k] = ~vis.w[k];
        }
        const bool any = mask_any(unv);
        const uint32_t bal = __ballot_sync(FULL, any);
        uint32_t wbase = 0;
        if (lane == 0 && bal) wbase = atomicAdd(&s_ncand, __popc(bal));
        wbase = __shfl_sync(FULL, wbase, 0);
        if (any) {
            const uint32_t i = wbase + __popc(bal & lanemask_lt());
            s_w[i] = w;
            s_unv[i] = unv;
            // consecutive vertices: the offsets come in with coalesced loads here instead of one
            // dependent look-up per candidate at the head of the warp's chain below
            s_e[i] = make_uint2(p.row[w], p.row[w + 1]);
        }
    }
    __syncthreads();
    const uint32_t ncand = s_ncand;

    // 2. warps take candidates dynamically (degree skew)
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(&s_next, 1);
        i = __shfl_sync(FULL, i, 0);
        if (i >= ncand) break;
        const uint32_t w = s_w[i];
        const SrcMask unv = s_unv[i];
        const uint32_t e0 = s_e[i].x, e1 = s_e[i].y;
        if (e1 - e0 > p.hub_degree && bc_defer_hub(p, level, g, w, unv)) continue;
        Acc acc[BK] = {};
        SrcMask newm = mask_zero();
        bc_fwd_scan(p, rows, Pcur, tag_cur, e0, e1, 0, 32, unv, lane, stage, acc, newm);
        if (mask_any(newm)) {
#pragma unroll
            for (int k = 0; k < BK; k++)
                if ((newm.w[k] >> lane) & 1u) bc_store_sigma<NARROW>(p, &rows[(size_t)w * BG + lane * BK + k], acc[k]);
            if (lane == 0) {
                SrcMask vis;
#pragma unroll
                for (int k = 0; k < BK; k++) vis.w[k] = ~unv.w[k] | newm.w[k];
                mask_store(&visited[w], vis);
                bc_store_level(Pnext, w, tag_next, newm);
                const uint32_t j = atomicAdd(&s_nout, 1u);
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
        const uint64_t pos = (uint64_t)s_base + tid;
        if (pos < p.qcap) {
            const size_t q = (size_t)g * p.qcap + pos;
            const uint32_t w = s_ov[tid];
            p.qv[q] = w;
            mask_store(&p.qm[q], s_om[tid]);
            p.qe[q] = make_uint2(p.srow[w], p.srow[w + 1]);
        } else {
            atomicOr(p.overflow, 2u);
        }
    }
}

// Deferred high-degree vertices of the forward level: one CTA per vertex, warps split the arcs.
template <bool NARROW>
__global__ void __launch_bounds__(BLOCK) bc_forward_hub_kernel(BcParams p, uint32_t level) {
    using Acc = typename SigmaRows<NARROW>::Acc;
    const uint32_t g = blockIdx.y;
    uint32_t *lc = p.lvlcnt + (size_t)g * p.lcap;
    uint32_t *lo = p.loff + (size_t)g * p.lcap;
    const uint32_t fcnt = lc[level];
    const uint32_t nh = min(*bc_hubcnt(p, level, g), p.hubcap);
    if (fcnt == 0 || nh == 0) return;
    const int tid = threadIdx.x, lane = lane_id(), wid = tid >> 5;
    const size_t gn = (size_t)g * p.n;
    SrcMask *visited = p.visited + gn;
    const PEntry *Pcur = bc_P(p, level) + gn;
    PEntry *Pnext = bc_P(p, level + 1) + gn;
    auto *rows = SigmaRows<NARROW>::rows(p, gn);
    const uint32_t tag_cur = p.tagbase + level + 1, tag_next = tag_cur + 1;
    __shared__ Acc s_acc[WARPS][BG];
    __shared__ SrcMask s_new[WARPS];
    extern __shared__ __align__(16) char s_dyn_stage[];
    char *stage = s_dyn_stage + (size_t)(threadIdx.x >> 5) * ((NARROW ? BC_FWD32_SMEM_BYTES : BC_SMEM_BYTES) / WARPS);

    for (uint32_t h = blockIdx.x; h < nh; h += gridDim.x) {
        const uint32_t w = p.hubv[(size_t)g * p.hubcap + h];
        const SrcMask unv = mask_load(&p.hubm[(size_t)g * p.hubcap + h]);
        const uint32_t e0 = p.row[w], e1 = p.row[w + 1];
        Acc acc[BK] = {};
        SrcMask newm = mask_zero();
        bc_fwd_scan(p, rows, Pcur, tag_cur, e0, e1, wid * 32, WARPS * 32, unv, lane, stage, acc, newm);
#pragma unroll
        for (int k = 0; k < BK; k++) s_acc[wid][lane * BK + k] = acc[k];
        if (lane == 0) s_new[wid] = newm;
        __syncthreads();
        if (wid == 0) {
            Acc tot[BK] = {};
            SrcMask nm = mask_zero();
            for (int q = 0; q < WARPS; q++) {
#pragma unroll
                for (int k = 0; k < BK; k++) {
                    tot[k] += s_acc[q][lane * BK + k];
                    nm.w[k] |= s_new[q].w[k];
                }
            }
            if (mask_any(nm)) {
#pragma unroll
                for (int k = 0; k < BK; k++)
                    if ((nm.w[k] >> lane) & 1u) bc_store_sigma<NARROW>(p, &rows[(size_t)w * BG + lane * BK + k], tot[k]);
                if (lane == 0) {
                    SrcMask vis;
#pragma unroll
                    for (int k = 0; k < BK; k++) vis.w[k] = ~unv.w[k] | nm.w[k];
                    mask_store(&visited[w], vis);
                    bc_store_level(Pnext, w, tag_next, nm);
                    const uint32_t pos = atomicAdd(&lc[level + 1], 1u);
                    atomicAdd(&p.lvl_total[level + 1], 1u);
                    const uint64_t qpos = (uint64_t)lo[level] + fcnt + pos;
                    if (qpos < p.qcap) {
                        const size_t q = (size_t)g * p.qcap + qpos;
                        p.qv[q] = w;
                        mask_store(&p.qm[q], nm);
                        p.qe[q] = make_uint2(p.srow[w], p.srow[w + 1]);
                    } else {
                        atomicOr(p.overflow, 2u);
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------

// Successor pull of one warp over chunks of 32 arcs starting at e0+first with stride `step`.
// Node mode sums coef rows into acc; edge mode adds sigma[v][b]*coef[w][b], summed over b, to the
// arc's edge and the same products into acc (centrality.rs:318-325).
template <bool EDGE, int NS = BSTAGE>
__device__ __forceinline__ void bc_bwd_scan(const BcParams &p, const double *rows,
                                            const PEntry *Pnext, uint32_t tag_next, uint32_t e0,
                                            uint32_t e1, uint32_t first, uint32_t step,
                                            const SrcMask &fm, const double (&sig)[BK], int lane,
                                            char *stage, double (&acc)[BK],
                                            unsigned long long &dag) {
    for (uint32_t e = e0 + first; e < e1; e += step) {
        const uint32_t idx = e + lane;
        uint32_t w = 0, ed = 0;
        SrcMask m = mask_zero();
        if (idx < e1) {
            w = p.col[idx];
            m = bc_level_mask(Pnext, w, tag_next);
#pragma unroll
            for (int k = 0; k < BK; k++) m.w[k] &= fm.w[k];
            if (EDGE && mask_any(m)) ed = p.eid[idx];
        }
        uint32_t nz = __ballot_sync(FULL, mask_any(m));
#if RXC_PROBE
        if (lane == 0) atomicAdd(&g_probe[PR_LOOKUPS64], (unsigned long long)min(32u, e1 - e));
#endif
        dag += mask_popc(m);  // per lane: the sources of this lane's own arc (summed over the warp at the end)
        if (!EDGE) {
            bc_gather<NS>(rows, w, m, nz, lane, stage, acc);
        } else {
            static_assert(!EDGE || NS == BSTAGE, "edge mode stages into the default slots");
            using G = RowGeom<double>;
            const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stage);
            const char *lane_rows = reinterpret_cast<const char *>(rows) + 16 * lane;
            asm("" : "+l"(lane_rows));
            uint32_t need_lane[G::SPL];
            row_need<G::SPL>(m, need_lane);
            while (nz) {
                // stage up to ESTAGE successor rows, then form per-arc sums over all sources of the
                // group, ER arcs per pass of the exchange network
                uint32_t todo = nz;
                int cnt = 0;
                // blocks of 2 or 3 slots: the shuffles of a block, then its predicated copies in one
                // straight line (see bc_gather); a slot past the last arc copies nothing
                constexpr int NB = ESTAGE % 3 == 0 ? 3 : (ESTAGE % 2 == 0 ? 2 : 1);
#pragma unroll
                for (int b = 0; b < ESTAGE; b += NB) {
                    if (nz) {  // warp-uniform
                        uint32_t wj[NB], need[NB][G::SPL];
#pragma unroll
                        for (int i = 0; i < NB; i++) {
                            const bool live = nz != 0;
                            const int a = (__ffs(nz) - 1) & 31;
                            nz &= nz - 1;
                            wj[i] = __shfl_sync(FULL, w, a);
#pragma unroll
                            for (int h = 0; h < G::SPL; h++) {
                                const uint32_t nd = __shfl_sync(FULL, need_lane[h], a);
                                need[i][h] = live ? nd : 0u;
                            }
                            cnt += live;
                        }
#pragma unroll
                        for (int i = 0; i < NB; i++)
                            cp_async_row<double>(sbase + (b + i) * G::ROW_BYTES, lane, lane_rows, wj[i], need[i]);
                    }
                }
                cp_async_wait_all();
                __syncwarp();  // the chunks a lane reads were copied by other lanes
#pragma unroll
                for (int b = 0; b < ESTAGE; b += ER) {
                    if (b < cnt) {  // warp-uniform
                        double c[ER];
                        uint32_t my_e = 0;
#pragma unroll
                        for (int i = 0; i < ER; i++) {
                            c[i] = 0.0;
                            if (b + i < ESTAGE && b + i < cnt) {
                                const int a = __ffs(todo) - 1;
                                todo &= todo - 1;
                                const SrcMask mj = mask_shfl(m, a);
                                const uint32_t ej = __shfl_sync(FULL, ed, a);
                                if (lane / (32 / ER) == i) my_e = ej;
                                double xs[BK];
                                stage_read<double>(stage + (b + i) * G::ROW_BYTES, lane, xs);
#pragma unroll
                                for (int k = 0; k < BK; k++) {
                                    if ((mj.w[k] >> lane) & 1u) {
                                        const double ck = sig[k] * xs[k];  // sigma[v]*coeff, centrality.rs:322
                                        acc[k] += ck;
                                        c[i] += ck;
                                    }
                                }
                            }
                        }
                        // reduce the ER per-arc sums over the warp at once: log2(ER) halving exchanges
                        // leave lane l with arc l / (32/ER); the remaining butterfly steps finish the sum
                        int off = 16;
#pragma unroll
                        for (int half = ER / 2; half >= 1; half >>= 1, off >>= 1) {
                            const bool upper = lane & off;
#pragma unroll
                            for (int j = 0; j < half; j++) {
                                const double send = upper ? c[j] : c[j + half];
                                const double recv = __shfl_xor_sync(FULL, send, off);
                                c[j] = (upper ? c[j + half] : c[j]) + recv;
                            }
                        }
                        double tot = c[0];
#pragma unroll
                        for (; off >= 1; off >>= 1) tot += __shfl_xor_sync(FULL, tot, off);
                        if (lane % (32 / ER) == 0 && b + lane / (32 / ER) < cnt) atomicAdd(&p.acc[my_e], tot);
                    }
                }
                __syncwarp();  // the slots may be refilled by other lanes from here on
            }
        }
    }
}

// Finish one queue entry (v, fm): rewrite the row with (1+delta)/sigma, add the vertex score,
// restore P for the level above.  centrality.rs:274-279 (coeff) and :281-298 (accumulation).
template <bool EDGE, bool ENDPOINTS>
__device__ __forceinline__ void bc_bwd_finish(const BcParams &p, uint32_t v, const SrcMask &fm,
                                              const double (&sig)[BK], const double (&acc)[BK],
                                              double *rows, PEntry *Pcur, uint32_t tag_cur,
                                              uint32_t level, int lane) {
    // The sums start at -0.0 and every gathered term is +0.0, positive or NaN, so a sum that is still
    // -0.0 had NO successor on the next level: the reference never touches delta[v] then and it stays
    // 0 (centrality.rs:271-279) -- even when sigma[v] overflowed to +inf, where sigma * 0 would be NaN
    // (tests: test_sigma_overflow_to_infinity_matches_the_reference_nan_pattern).
    double contrib = 0.0;
#pragma unroll
    for (int k = 0; k < BK; k++) {
        if ((fm.w[k] >> lane) & 1u) {
            const bool leaf = acc[k] == 0.0 && signbit(acc[k]);
            const double delta = leaf ? 0.0 : (EDGE ? acc[k] : sig[k] * acc[k]);
            rows[(size_t)v * BG + lane * BK + k] = (1.0 + delta) / sig[k];
            if (!EDGE && level > 0) contrib += ENDPOINTS ? delta + 1.0 : delta;
        }
    }
    // (an entry without a successor for any of its sources contributes nothing: no reduction, no atomic --
    // most entries of the deepest levels)
    if (!EDGE && level > 0 && __any_sync(FULL, contrib != 0.0)) {
        const double tot = warp_sum(contrib);
        if (lane == 0) atomicAdd(&p.acc[v], tot);
    }
    if (lane == 0) bc_store_level(Pcur, v, tag_cur, fm);
}

// One level of the dependency sweep (deepest level first).  A warp owns one queue entry (v, fm):
//   delta[v][b] = sigma[v][b] * sum_{w in N+(v), depth_b(w) = level+1} (1+delta[w][b])/sigma[w][b]
// which regroups centrality.rs:272-279 by predecessor (pull instead of push; no f64 atomics on the
// rows).  The row of v is then overwritten with (1+delta)/sigma for the level above to pull.
template <bool EDGE, bool ENDPOINTS, bool NARROW>
__global__ void __launch_bounds__(BLOCK, RXC_MINBLOCKS) bc_backward_kernel(BcParams p, uint32_t level) {
    const uint32_t g = blockIdx.y;
    if (blockIdx.x == 0 && threadIdx.x == 0) *bc_hubcnt(p, level + 1, g) = 0;  // for the next launch
    const uint32_t cnt = p.lvlcnt[(size_t)g * p.lcap + level];
    const uint32_t off = p.loff[(size_t)g * p.lcap + level];
    const int lane = lane_id();
    const size_t gn = (size_t)g * p.n;
    const PEntry *Pnext = bc_P(p, level + 1) + gn;
    PEntry *Pcur = bc_P(p, level) + gn;
    double *rows = p.sc + gn * BG;
    const uint32_t tag_cur = p.tagbase + level + 1, tag_next = tag_cur + 1;
    const uint32_t *qv = p.qv + (size_t)g * p.qcap + off;
    const SrcMask *qm = p.qm + (size_t)g * p.qcap + off;
    const uint2 *qe = p.qe + (size_t)g * p.qcap + off;

    extern __shared__ __align__(16) char s_dyn_stage[];
    char *stage = s_dyn_stage + (size_t)(threadIdx.x >> 5) * (BC_BWD_SMEM_BYTES / WARPS);
    unsigned long long te = 0, vv = 0, dag = 0;
    unsigned long long reach[BK] = {};
    const uint32_t warp = blockIdx.x * WARPS + (threadIdx.x >> 5);
    for (uint32_t i = warp; i < cnt; i += gridDim.x * WARPS) {
        const uint32_t v = qv[i];
        const SrcMask fm = mask_load(&qm[i]);
        const uint2 ee = qe[i];  // = (row[v], row[v + 1]), stored at discovery
        const uint32_t e0 = ee.x, e1 = ee.y;
        if (lane == 0) {
            te += (unsigned long long)mask_popc(fm) * (e1 - e0);
            vv += mask_popc(fm);
        }
        if (ENDPOINTS && level > 0) {
#pragma unroll
            for (int k = 0; k < BK; k++) reach[k] += (fm.w[k] >> lane) & 1u;
        }
        if (e1 - e0 > p.hub_degree && bc_defer_hub(p, level, g, v, fm)) continue;
        const RowK sg = bc_load_sigma<NARROW>(p, gn, v, lane, (mask_or(fm) >> lane) & 1u);
        double acc[BK];  // -0.0: "nothing added yet", see bc_bwd_finish
#pragma unroll
        for (int k = 0; k < BK; k++) acc[k] = -0.0;
        bc_bwd_scan<EDGE>(p, rows, Pnext, tag_next, e0, e1, 0, 32, fm, sg.x, lane, stage, acc, dag);
        bc_bwd_finish<EDGE, ENDPOINTS>(p, v, fm, sg.x, acc, rows, Pcur, tag_cur, level, lane);
    }
    if (ENDPOINTS) {
#pragma unroll
        for (int k = 0; k < BK; k++)
            if (reach[k]) atomicAdd(&p.reach[(size_t)g * BG + k * 32 + lane], reach[k]);
    }
    dag = warp_sum_u64(dag);
    if (lane == 0 && (te | vv | dag)) {
        stat_add(p.stats, ST_TE, te);
        stat_add(p.stats, ST_V, vv);
        stat_add(p.stats, ST_DAG, dag);
    }
}

// ---------------------------------------------------------------------------------------------
// Look-ahead variant for LEAF-HEAVY levels (node mode).  On the deepest big level of a small-world
// graph most entries have no successor on the next level: an entry is a header, 16 look-ups that find
// nothing, four divisions and a row store -- ~370 instructions and FOUR dependent memory round trips
// (header -> targets + own sigma -> level masks -> nothing to gather), and the launch runs at a third
// of the DRAM rate with the issue slots 38 % busy.  Here a warp works on four entries at once: while
// entry t is finished, the level masks of entry t+1's targets, the targets and the own sigma of entry
// t+2 and the header of entry t+3 are on their way into lane-private shared-memory slots (cp.async),
// so the round trips overlap across entries.  The buffers cost two row slots (4 instead of 6), which
// is why the gather-heavy levels keep the plain kernel: applied to every level the look-ahead was
// neutral (561.0 against 561.9 ms per 16384 sources, profiles/r02_ab_staging_variants.txt); the host
// picks it per level (rxcuda.cu: the next level has less than a quarter of this level's entries).
// ---------------------------------------------------------------------------------------------
constexpr int LSTAGE = 4;                 // row slots per warp
constexpr int LOOK_HDR = 16 * MQ + 16;    // mask, arc range, vertex of one queue entry
constexpr int LOOK_BYTES = 4 * LOOK_HDR + 2 * 128 + 128 + 32 * 16 * MQ + 2 * 32 * 4 * BK;  // per warp
constexpr int BC_BWD_LOOK_SMEM_BYTES = WARPS * (LSTAGE * BG * 8 + ARCTAB_BYTES + LOOK_BYTES);

template <bool ENDPOINTS, bool NARROW>
__global__ void __launch_bounds__(BLOCK, RXC_MINBLOCKS) bc_backward_look_kernel(BcParams p, uint32_t level) {
    const uint32_t g = blockIdx.y;
    if (blockIdx.x == 0 && threadIdx.x == 0) *bc_hubcnt(p, level + 1, g) = 0;  // for the next launch
    const uint32_t cnt = p.lvlcnt[(size_t)g * p.lcap + level];
    const uint32_t off = p.loff[(size_t)g * p.lcap + level];
    const int lane = lane_id();
    const size_t gn = (size_t)g * p.n;
    const PEntry *Pnext = bc_P(p, level + 1) + gn;
    PEntry *Pcur = bc_P(p, level) + gn;
    double *rows = p.sc + gn * BG;
    const uint32_t tag_cur = p.tagbase + level + 1, tag_next = tag_cur + 1;
    const uint32_t *qv = p.qv + (size_t)g * p.qcap + off;
    const SrcMask *qm = p.qm + (size_t)g * p.qcap + off;
    const uint2 *qe = p.qe + (size_t)g * p.qcap + off;

    extern __shared__ __align__(16) char s_dyn_stage[];
    char *stage = s_dyn_stage + (size_t)(threadIdx.x >> 5) * (BC_BWD_LOOK_SMEM_BYTES / WARPS);
    // look-ahead buffers of this warp, behind its row slots and arc table
    char *look = stage + LSTAGE * BG * 8 + ARCTAB_BYTES;
    char *l_hdr = look;                                                   // [4][LOOK_HDR]
    uint32_t *l_col = reinterpret_cast<uint32_t *>(look + 4 * LOOK_HDR);  // [2][32]
    uint32_t *l_tag = l_col + 64;                                         // [32]
    SrcMask *l_mask = reinterpret_cast<SrcMask *>(l_tag + 32);            // [32]
    uint32_t *l_sig = reinterpret_cast<uint32_t *>(l_mask + 32);          // [2][32][BK]: own uint32 counts
    const uint32_t a_hdr = (uint32_t)__cvta_generic_to_shared(l_hdr);
    const uint32_t a_col = (uint32_t)__cvta_generic_to_shared(l_col) + 4u * lane;
    const uint32_t a_tag = (uint32_t)__cvta_generic_to_shared(l_tag) + 4u * lane;
    const uint32_t a_mask = (uint32_t)__cvta_generic_to_shared(l_mask) + (uint32_t)sizeof(SrcMask) * lane;
    const uint32_t a_sig = (uint32_t)__cvta_generic_to_shared(l_sig) + 4u * BK * lane;

    unsigned long long te = 0, vv = 0, dag = 0;
    unsigned long long reach[BK] = {};
    const uint32_t warp = blockIdx.x * WARPS + (threadIdx.x >> 5);
    const uint32_t stride = gridDim.x * WARPS;
    const uint32_t nw = cnt > warp ? (cnt - warp + stride - 1) / stride : 0;  // entries of this warp
    // step k: header of entry k; targets and own sigma of entry k-1; level masks of entry k-2; entry
    // k-3 is finished.  Every slot is written and read by the same lane, except the headers, which
    // the __syncwarp below publishes.
    for (uint32_t k = 0; k < nw + 3; k++) {
        cp_async_wait_all();
        __syncwarp();
        uint32_t v = 0, e0 = 0, e1 = 0, w0 = 0;
        SrcMask fm = mask_zero(), m0 = mask_zero();
        RowK sg;
#pragma unroll
        for (int j = 0; j < BK; j++) sg.x[j] = 0.0;
        if (k >= 3) {
            const char *h = l_hdr + ((k - 3) & 3) * LOOK_HDR;
            fm = mask_load(reinterpret_cast<const SrcMask *>(h));
            const uint2 ee = *reinterpret_cast<const uint2 *>(h + 16 * MQ);
            v = *reinterpret_cast<const uint32_t *>(h + 16 * MQ + 8);
            e0 = ee.x;
            e1 = ee.y;
            if (e0 + lane < e1) {
                w0 = l_col[((k - 3) & 1) * 32 + lane];
                m0 = mask_load(&l_mask[lane]);
                const bool ok = l_tag[lane] == tag_next;
#pragma unroll
                for (int j = 0; j < BK; j++) m0.w[j] = ok ? (m0.w[j] & fm.w[j]) : 0u;
            }
            if (NARROW && ((mask_or(fm) >> lane) & 1u)) {
#pragma unroll
                for (int q = 0; q < MQ; q++) {
                    const uint4 t = reinterpret_cast<const uint4 *>(l_sig + ((k - 3) & 1) * 32 * BK + lane * BK)[q];
                    sg.x[4 * q] = (double)t.x;
                    sg.x[4 * q + 1] = (double)t.y;
                    sg.x[4 * q + 2] = (double)t.z;
                    sg.x[4 * q + 3] = (double)t.w;
                }
            }
        }
        if (k >= 2 && k - 2 < nw) {  // level masks of entry k-2's first 32 targets
            const char *h = l_hdr + ((k - 2) & 3) * LOOK_HDR;
            const uint2 ee = *reinterpret_cast<const uint2 *>(h + 16 * MQ);
            if (ee.x + lane < ee.y) {
                const PEntry *pe = Pnext + l_col[((k - 2) & 1) * 32 + lane];
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(a_tag), "l"(pe) : "memory");
#pragma unroll
                for (int q = 0; q < MQ; q++)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(a_mask + 16u * q),
                                 "l"(reinterpret_cast<const uint4 *>(pe) + MQ + q)
                                 : "memory");
            }
        }
        if (k >= 1 && k - 1 < nw) {  // first 32 targets and own sigma of entry k-1
            const char *h = l_hdr + ((k - 1) & 3) * LOOK_HDR;
            const uint2 ee = *reinterpret_cast<const uint2 *>(h + 16 * MQ);
            if (ee.x + lane < ee.y)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(a_col + 128u * ((k - 1) & 1)),
                             "l"(p.col + ee.x + lane)
                             : "memory");
            if (NARROW) {
                const SrcMask f1 = mask_load(reinterpret_cast<const SrcMask *>(h));
                if ((mask_or(f1) >> lane) & 1u) {
                    const uint32_t v1 = *reinterpret_cast<const uint32_t *>(h + 16 * MQ + 8);
                    const uint4 *src = reinterpret_cast<const uint4 *>(p.sig32 + (gn + v1) * BG + lane * BK);
#pragma unroll
                    for (int q = 0; q < MQ; q++)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(
                                         a_sig + 4u * 32 * BK * ((k - 1) & 1) + 16u * q),
                                     "l"(src + q)
                                     : "memory");
                }
            }
        }
        if (k < nw) {  // header of entry k
            const size_t i = (size_t)warp + (size_t)k * stride;
            const uint32_t dst = a_hdr + (k & 3) * LOOK_HDR;
            if (lane < MQ)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dst + 16u * lane),
                             "l"(reinterpret_cast<const uint4 *>(&qm[i]) + lane)
                             : "memory");
            else if (lane == MQ)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst + 16u * MQ), "l"(&qe[i]) : "memory");
            else if (lane == MQ + 1)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst + 16u * MQ + 8u), "l"(&qv[i]) : "memory");
        }
        if (k < 3) continue;

        if (lane == 0) {
            te += (unsigned long long)mask_popc(fm) * (e1 - e0);
            vv += mask_popc(fm);
        }
        if (ENDPOINTS && level > 0) {
#pragma unroll
            for (int j = 0; j < BK; j++) reach[j] += (fm.w[j] >> lane) & 1u;
        }
        if (e1 - e0 > p.hub_degree && bc_defer_hub(p, level, g, v, fm)) continue;
        if (!NARROW) sg = bc_load_sigma<false>(p, gn, v, lane, (mask_or(fm) >> lane) & 1u);
        double acc[BK];  // -0.0: "nothing added yet", see bc_bwd_finish
#pragma unroll
        for (int j = 0; j < BK; j++) acc[j] = -0.0;
        const uint32_t nz = __ballot_sync(FULL, mask_any(m0));
#if RXC_PROBE
        if (lane == 0 && e1 > e0) atomicAdd(&g_probe[PR_LOOKUPS64], (unsigned long long)min(32u, e1 - e0));
#endif
        dag += mask_popc(m0);
        bc_gather<LSTAGE>(rows, w0, m0, nz, lane, stage, acc);
        if (e1 - e0 > 32)  // warp-uniform: the remaining chunks the plain way
            bc_bwd_scan<false, LSTAGE>(p, rows, Pnext, tag_next, e0, e1, 32, 32, fm, sg.x, lane, stage, acc, dag);
        bc_bwd_finish<false, ENDPOINTS>(p, v, fm, sg.x, acc, rows, Pcur, tag_cur, level, lane);
    }
    if (ENDPOINTS) {
#pragma unroll
        for (int j = 0; j < BK; j++)
            if (reach[j]) atomicAdd(&p.reach[(size_t)g * BG + j * 32 + lane], reach[j]);
    }
    dag = warp_sum_u64(dag);
    if (lane == 0 && (te | vv | dag)) {
        stat_add(p.stats, ST_TE, te);
        stat_add(p.stats, ST_V, vv);
        stat_add(p.stats, ST_DAG, dag);
    }
}

// Deferred high-degree entries of a backward level: one CTA per entry, warps split the arcs.
template <bool EDGE, bool ENDPOINTS, bool NARROW>
__global__ void __launch_bounds__(BLOCK) bc_backward_hub_kernel(BcParams p, uint32_t level) {
    const uint32_t g = blockIdx.y;
    const uint32_t nh = min(*bc_hubcnt(p, level, g), p.hubcap);
    if (nh == 0) return;
    const int tid = threadIdx.x, lane = lane_id(), wid = tid >> 5;
    const size_t gn = (size_t)g * p.n;
    const PEntry *Pnext = bc_P(p, level + 1) + gn;
    PEntry *Pcur = bc_P(p, level) + gn;
    double *rows = p.sc + gn * BG;
    const uint32_t tag_cur = p.tagbase + level + 1, tag_next = tag_cur + 1;
    __shared__ double s_acc[WARPS][BG];
    __shared__ unsigned long long s_dag[WARPS];
    extern __shared__ __align__(16) char s_dyn_stage[];
    char *stage = s_dyn_stage + (size_t)(threadIdx.x >> 5) * (BC_BWD_SMEM_BYTES / WARPS);

    for (uint32_t h = blockIdx.x; h < nh; h += gridDim.x) {
        const uint32_t v = p.hubv[(size_t)g * p.hubcap + h];
        const SrcMask fm = mask_load(&p.hubm[(size_t)g * p.hubcap + h]);
        const uint32_t e0 = p.row[v], e1 = p.row[v + 1];
        const RowK sg = bc_load_sigma<NARROW>(p, gn, v, lane, (mask_or(fm) >> lane) & 1u);
        double acc[BK];
#pragma unroll
        for (int k = 0; k < BK; k++) acc[k] = -0.0;
        unsigned long long dag = 0;
        bc_bwd_scan<EDGE>(p, rows, Pnext, tag_next, e0, e1, wid * 32, WARPS * 32, fm, sg.x, lane,
                          stage, acc, dag);
        __syncthreads();  // every warp has read row v before warp 0 rewrites it
#pragma unroll
        for (int k = 0; k < BK; k++) s_acc[wid][lane * BK + k] = acc[k];
        dag = warp_sum_u64(dag);
        if (lane == 0) s_dag[wid] = dag;
        __syncthreads();
        if (wid == 0) {
            double tot[BK];  // stays -0.0 only if every warp's sum is -0.0
#pragma unroll
            for (int k = 0; k < BK; k++) tot[k] = -0.0;
            unsigned long long d = 0;
            for (int q = 0; q < WARPS; q++) {
#pragma unroll
                for (int k = 0; k < BK; k++) tot[k] += s_acc[q][lane * BK + k];
                d += s_dag[q];
            }
            bc_bwd_finish<EDGE, ENDPOINTS>(p, v, fm, sg.x, tot, rows, Pcur, tag_cur, level, lane);
            if (lane == 0 && d) stat_add(p.stats, ST_DAG, d);
        }
        __syncthreads();
    }
}

// Ordering key of a source: the size of its two-hop neighbourhood along the expansion direction,
// sum of deg(w) over its arcs (s,w).  On BA(200k,8) it correlates -0.985 with the mean BFS distance
// of the source, i.e. with WHEN its search reaches the bulk of the graph (rxcuda.cu, bc_order_sources).
__global__ void __launch_bounds__(BLOCK)
    bc_source_keys_kernel(const uint32_t *row, const uint32_t *col, const uint32_t *src, uint32_t k,
                          unsigned long long *keys) {
    const uint32_t i = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (i >= k) return;
    const uint32_t s = src[i];
    unsigned long long sum = 0;
    for (uint32_t e = row[s] + lane_id(); e < row[s + 1]; e += 32) {
        const uint32_t w = col[e];
        sum += row[w + 1] - row[w];
    }
    sum = warp_sum_u64(sum);
    if (lane_id() == 0) keys[i] = sum;
}

// bc_order = 2 (experimental): the highest-degree neighbour of each source.
// Sources that hang off the same hub are siblings in the BFS tree of that hub and agree on the level
// of almost every vertex; sorting by (hub, key) puts them into the same lane.
__global__ void __launch_bounds__(BLOCK)
    bc_source_hubs_kernel(const uint32_t *row, const uint32_t *col, const uint32_t *src, uint32_t k,
                          uint32_t *hubs) {
    const uint32_t i = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (i >= k) return;
    const uint32_t s = src[i];
    // packed (degree, ~id): the maximum is the highest degree, ties to the lowest id
    unsigned long long best = 0;
    for (uint32_t e = row[s] + lane_id(); e < row[s + 1]; e += 32) {
        const uint32_t w = col[e];
        const unsigned long long cand = ((unsigned long long)(row[w + 1] - row[w]) << 32) | (0xFFFFFFFFu - w);
        best = cand > best ? cand : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(FULL, best, o);
        best = other > best ? other : best;
    }
    if (lane_id() == 0) hubs[i] = best ? 0xFFFFFFFFu - (uint32_t)best : s;  // isolated source: itself
}

// endpoints: the source itself receives |S| - 1 (centrality.rs:282-284)
__global__ void __launch_bounds__(BLOCK) bc_endpoint_sources_kernel(BcParams p, uint32_t ngroups) {
    const uint32_t i = blockIdx.x * BLOCK + threadIdx.x;
    if (i >= ngroups * BG) return;
    const uint32_t s = p.src[i];
    if (s != NO_SOURCE) atomicAdd(&p.acc[s], (double)p.reach[i]);
}

// ---------------------------------------------------------------------------------------------
// group betweenness (rustworkx-core/src/centrality.rs:1963-2101)
//
// The reference runs two sigma-BFSs per non-group source: one on the full graph and one on the
// graph with the group's vertices removed.  Here both run in the SAME forward pass as a pair of
// slots: slot j (< 64) of a group is the full BFS of source j, slot 64+j the BFS of the same source
// with the group removed -- lane l holds the pairs (l, 64+l) and (32+l, 96+l) in ONE sector, elements
// {0,2} and {1,3}.  "Removed" is a preset: the group's vertices start out visited in slots 64..127,
// so they are never discovered there and never act as parents.  A queue entry (v, mask) of level L
// holds bit j iff dist_full(s_j, v) = L and bit 64+j iff dist_no_group(s_j, v) = L, so the
// reference's test `dist_no_group[t] == dist_full[t]` (:2066) is "both bits in the same entry".
// ---------------------------------------------------------------------------------------------
constexpr int BGP = BG / 2;  // sources per group in pair mode

// visited[g][m] |= slots 64..127 for every member m of the group
__global__ void __launch_bounds__(BLOCK)
    bc_group_block_kernel(BcParams p, const uint32_t *members, uint32_t nmem) {
    const uint32_t g = blockIdx.y;
    const uint32_t i = blockIdx.x * BLOCK + threadIdx.x;
    if (i >= nmem) return;
    SrcMask *q = &p.visited[(size_t)g * p.n + members[i]];
    SrcMask m = mask_load(q);
#pragma unroll
    for (int k = BK / 2; k < BK; k++) m.w[k] = 0xFFFFFFFFu;
    mask_store(q, m);
}

// Level 0 in pair mode: source j seeds slots j and 64+j with one queue entry carrying both bits.
template <bool NARROW>
__global__ void __launch_bounds__(32) bc_seed_pair_kernel(BcParams p) {
    const uint32_t g = blockIdx.x;
    const int lane = lane_id();
    const SrcMask valid = bc_valid_mask(p, g);
    uint32_t base = 0;
#pragma unroll
    for (int k = 0; k < BK / 2; k++) {
        const uint32_t s = p.src[(size_t)g * BG + k * 32 + lane];
        if (s != NO_SOURCE) {
            const size_t gv = (size_t)g * p.n + s;
            SrcMask bit = mask_zero();
            bit.w[k] = bit.w[k + BK / 2] = 1u << lane;
            SrcMask vis;
#pragma unroll
            for (int j = 0; j < BK; j++) vis.w[j] = ~valid.w[j] | bit.w[j];
            mask_store(&p.visited[gv], vis);  // sources are distinct vertices outside the group
            bc_store_level(p.P0 + (size_t)g * p.n, s, p.tagbase + 1, bit);
            if (NARROW) {
                p.sig32[gv * BG + lane * BK + k] = 1u;
                p.sig32[gv * BG + lane * BK + k + BK / 2] = 1u;
            } else {
                p.sc[gv * BG + lane * BK + k] = 1.0;
                p.sc[gv * BG + lane * BK + k + BK / 2] = 1.0;
            }
            const uint32_t i = base + __popc(valid.w[k] & lanemask_lt());
            p.qv[(size_t)g * p.qcap + i] = s;
            mask_store(&p.qm[(size_t)g * p.qcap + i], bit);
            p.qe[(size_t)g * p.qcap + i] = make_uint2(p.srow[s], p.srow[s + 1]);
        }
        base += __popc(valid.w[k]);
    }
    if (lane == 0) {
        p.loff[(size_t)g * p.lcap] = 0;
        p.lvlcnt[(size_t)g * p.lcap] = base;
        p.hubcnt[g] = 0;
        p.hubcnt[p.ngroups + g] = 0;
        atomicAdd(&p.lvl_total[0], base);
    }
}

// After the forward pass: for every queue entry (t, mask) beyond level 0 with t outside the group,
// add (sigma_full - paths_avoiding) / sigma_full for each source whose full BFS found t on this
// level (centrality.rs:2054-2078).  One warp per entry; one atomicAdd per warp.
template <bool NARROW>
__global__ void __launch_bounds__(BLOCK) bc_group_finalize_kernel(BcParams p, uint32_t nlevels) {
    const uint32_t g = blockIdx.y;
    const uint32_t *lc = p.lvlcnt + (size_t)g * p.lcap;
    __shared__ uint32_t s_end;
    if (threadIdx.x == 0) s_end = 0;
    __syncthreads();
    uint32_t part = 0;
    for (uint32_t l = threadIdx.x; l < nlevels; l += BLOCK) part += lc[l];
    part = (uint32_t)warp_sum_u64(part);
    if (lane_id() == 0 && part) atomicAdd(&s_end, part);
    __syncthreads();
    const uint32_t end = s_end, first = lc[0];  // entries [0, first) are the sources themselves
    const int lane = lane_id();
    const uint32_t *qv = p.qv + (size_t)g * p.qcap;
    const SrcMask *qm = p.qm + (size_t)g * p.qcap;
    double sum = 0.0;
    unsigned long long te = 0, vv = 0;
    for (uint32_t i = blockIdx.x * WARPS + (threadIdx.x >> 5); i < end; i += gridDim.x * WARPS) {
        const uint32_t t = qv[i];
        const SrcMask fm = mask_load(&qm[i]);
        if (lane == 0) {
            te += (unsigned long long)mask_popc(fm) * (p.srow[t + 1] - p.srow[t]);
            vv += mask_popc(fm);
        }
        if (i < first || ((p.blocked[t >> 5] >> (t & 31)) & 1u)) continue;
        uint32_t full = 0;
#pragma unroll
        for (int k = 0; k < BK / 2; k++) full |= fm.w[k];
        const RowK r = bc_load_sigma<NARROW>(p, (size_t)g * p.n, t, lane, (full >> lane) & 1u);
#pragma unroll
        for (int k = 0; k < BK / 2; k++) {
            if ((fm.w[k] >> lane) & 1u) {
                const double sf = r.x[k];
                const double sa = ((fm.w[k + BK / 2] >> lane) & 1u) ? r.x[k + BK / 2] : 0.0;
                sum += (sf - sa) / sf;
            }
        }
    }
    sum = warp_sum(sum);
    if (lane == 0) {
        if (sum != 0.0) atomicAdd(&p.acc[0], sum);
        if (te | vv) {
            stat_add(p.stats, ST_TE, te);
            stat_add(p.stats, ST_V, vv);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// distance-only BFS in the betweenness layout (closeness, distance_matrix, average path length on
// shallow graphs; rustworkx-core/src/centrality.rs:1190-1220, shortest_path/distance_matrix.rs:127-157)
//
// Same bottom-up level kernel as bc_forward_kernel without the sigma rows: an unfinished vertex
// scans its arcs of the pull adjacency, ORs the tagged 128-bit level masks of the neighbours that sit
// on the frontier, and stops as soon as every missing source has been found.  One random access
// moves a whole 32-byte sector and serves 128 sources; the 32-source top-down kernel it replaces on
// small-world graphs (bfs_level_kernel: one 4-byte word and one atomicOr per arc and group) ran at
// 131 GTEPS on BA(1M,8).  Per-source reach counts and distance sums are accumulated per lane
// (lane l owns bit l of the four mask words), reduced in shared memory and added once per CTA.
// ---------------------------------------------------------------------------------------------
template <bool ROWS>
__global__ void __launch_bounds__(BLOCK, (BK == 4 ? 5 : 3)) bfs128_level_kernel(BcParams p, uint32_t level) {
    const uint32_t g = blockIdx.y;
    uint32_t *lc = p.lvlcnt + (size_t)g * p.lcap;
    uint32_t *lo = p.loff + (size_t)g * p.lcap;
    const uint32_t fcnt = lc[level];
    if (fcnt == 0) return;  // this group's BFS has finished

    __shared__ SrcMask s_unv[BLOCK];
    __shared__ uint32_t s_w[BLOCK], s_ov[BLOCK], s_cnt[BG];
    __shared__ uint32_t s_ncand, s_next, s_nout, s_base;
    __shared__ unsigned long long s_te;
    const int tid = threadIdx.x, lane = lane_id();
    if (tid == 0) {
        s_ncand = 0;
        s_next = 0;
        s_nout = 0;
        s_te = 0;
    }
    if (tid < BG) s_cnt[tid] = 0;
    __syncthreads();

    const size_t gn = (size_t)g * p.n;
    SrcMask *visited = p.visited + gn;
    const PEntry *Pcur = bc_P(p, level) + gn;
    PEntry *Pnext = bc_P(p, level + 1) + gn;
    const uint32_t tag_cur = p.tagbase + level + 1, tag_next = tag_cur + 1;

    // 1. this CTA's 256 vertices: compact the unfinished ones (sparse levels: only marked ones)
    {
        const uint32_t w = blockIdx.x * BLOCK + tid;
        SrcMask unv = mask_zero();
        bool look = w < p.n;
        if (fcnt <= p.sparse_max) {
            uint32_t *word = p.maybe + (size_t)g * p.mwords + (w >> 5);
            const uint32_t bits = (w < p.n) ? *word : 0u;
            look = (bits >> lane) & 1u;
            __syncwarp();
            if (lane == 0 && bits) *word = 0;
        }
        if (look) {
            const SrcMask vis = mask_load(&visited[w]);
#pragma unroll
            for (int k = 0; k < BK; k++) unv.w[k] = ~vis.w[k];
        }
        const bool any = mask_any(unv);
        const uint32_t bal = __ballot_sync(FULL, any);
        uint32_t wbase = 0;
        if (lane == 0 && bal) wbase = atomicAdd(&s_ncand, __popc(bal));
        wbase = __shfl_sync(FULL, wbase, 0);
        if (any) {
            const uint32_t i = wbase + __popc(bal & lanemask_lt());
            s_w[i] = w;
            s_unv[i] = unv;
        }
    }
    __syncthreads();
    const uint32_t ncand = s_ncand;

    // 2. warps take candidates dynamically
    uint32_t cnt[BK] = {};
    unsigned long long te = 0;
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(&s_next, 1);
        i = __shfl_sync(FULL, i, 0);
        if (i >= ncand) break;
        const uint32_t w = s_w[i];
        const SrcMask unv = s_unv[i];
        const uint32_t e0 = p.row[w], e1 = p.row[w + 1];
        SrcMask newm = mask_zero();
        for (uint32_t e = e0; e < e1; e += 32) {
            const uint32_t idx = e + lane;
            SrcMask m = mask_zero();
            if (idx < e1) {
                m = bc_level_mask(Pcur, p.col[idx], tag_cur);
#pragma unroll
                for (int k = 0; k < BK; k++) m.w[k] &= unv.w[k];
            }
            if (__any_sync(FULL, mask_any(m))) {
#pragma unroll
                for (int k = 0; k < BK; k++) newm.w[k] |= __reduce_or_sync(FULL, m.w[k]);
                // every missing source found: the remaining arcs cannot add anything
                if (mask_eq(newm, unv)) break;
            }
        }
        if (!mask_any(newm)) continue;
#pragma unroll
        for (int k = 0; k < BK; k++) {
            const uint32_t bit = (newm.w[k] >> lane) & 1u;
            cnt[k] += bit;
            if (ROWS && bit) p.drows[((size_t)g * BG + 32 * k + lane) * p.n_cols + w] = (double)(level + 1);
        }
        if (lane == 0) {
            SrcMask vis;
#pragma unroll
            for (int k = 0; k < BK; k++) vis.w[k] = ~unv.w[k] | newm.w[k];
            mask_store(&visited[w], vis);
            bc_store_level(Pnext, w, tag_next, newm);
            s_ov[atomicAdd(&s_nout, 1u)] = w;
            te += (unsigned long long)mask_popc(newm) * (p.srow[w + 1] - p.srow[w]);
        }
    }
#pragma unroll
    for (int k = 0; k < BK; k++)
        if (cnt[k]) atomicAdd(&s_cnt[32 * k + lane], cnt[k]);
    if (lane == 0 && te) atomicAdd(&s_te, te);
    __syncthreads();

    // 3. queue the discovered vertices (the frontier-marking kernel of the next level reads them),
    //    add this CTA's per-source counts
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
        if (s_te) stat_add(p.stats, ST_TE, s_te);
    }
    __syncthreads();
    if ((uint32_t)tid < nout) {
        const uint64_t pos = (uint64_t)s_base + tid;
        if (pos < p.qcap) p.qv[(size_t)g * p.qcap + pos] = s_ov[tid];
        else atomicOr(p.overflow, 2u);
    }
    if (tid < BG && s_cnt[tid]) {
        atomicAdd(&p.reach[(size_t)g * BG + tid], (unsigned long long)s_cnt[tid]);
        atomicAdd(&p.dsum[(size_t)g * BG + tid], (unsigned long long)s_cnt[tid] * (level + 1));
        stat_add(p.stats, ST_V, (unsigned long long)s_cnt[tid]);
    }
}

// the sources themselves: reached at distance 0 (closeness counts them, centrality.rs:1209)
__global__ void __launch_bounds__(BLOCK) bfs128_sources_kernel(BcParams p, uint32_t ngroups, int rows) {
    const uint32_t i = blockIdx.x * BLOCK + threadIdx.x;
    if (i >= ngroups * BG) return;
    const uint32_t s = p.src[i];
    p.dsum[i] = 0ull;
    p.reach[i] = s != NO_SOURCE ? 1ull : 0ull;
    if (s == NO_SOURCE) return;
    if (rows) p.drows[(size_t)i * p.n_cols + s] = 0.0;
    stat_add(p.stats, ST_TE, (unsigned long long)(p.srow[s + 1] - p.srow[s]));
    stat_add(p.stats, ST_V, 1ull);
}

}  // namespace rxc
