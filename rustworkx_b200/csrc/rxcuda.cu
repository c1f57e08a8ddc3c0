// rxcuda.cu -- C-ABI (include/rxcuda.h) and host-side batch scheduler of librxcuda.so.
//
// Host code here plays the role the north star gives to a `rustworkx-cuda` crate: it snapshots the
// StableGraph into device CSR/CSC, cuts the live sources into batches of 32-source groups sized for
// HBM, drives the level-synchronous kernels of bc_kernels.cuh / bfs_kernels.cuh on one stream, and
// scatters results back to original indices.  No PyTorch, no Triton, no CPU fallback.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <chrono>
#include <exception>
#include <limits>
#include <memory>
#include <mutex>
#include <thread>
#include <type_traits>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rxcuda.h"
#include "bc_kernels.cuh"
#include "bfs_kernels.cuh"
#include "common.cuh"
#include "csr_build.cuh"
#include "wsp_kernels.cuh"
#include "snapshot.hpp"

namespace rxc {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int comm_reduce_device(double *d_inout, uint64_t len, int root, cudaStream_t st);  // comm.cpp
int comm_rendezvous_device(cudaStream_t st);
int comm_rank();
int comm_world();

struct OomError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct NcclError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// Device buffers come from the device's stream-ordered memory pool with an unlimited release
// threshold: freeing a multi-GB workspace (rxc_graph_free) hands it back to the pool, and the next
// snapshot gets it again without a round trip through the driver's physical allocator.
static void pool_setup(int device) {
    static bool done[64] = {false};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t threshold = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    cudaGetLastError();
    done[device] = true;
}

// The betweenness scratch is tens of GB.  Fresh memory from the stream-ordered pool costs ~50 ms per
// GB the first time (measured: 1.7 s for the 30 GB of a 2048-source batch at n = 10^6, 0.4 s for four
// bare cudaMallocAsync calls of that size), one cudaMalloc of the same size 10-35 ms
// (scripts/micro/alloc_time.cu).  So the scratch is ONE cudaMalloc slab, carved into the arrays; a
// released slab is parked per device (cudaFree synchronises the device and the next handle usually
// wants the same size) and only freed when a larger one is needed or rxc_snapshot_cache_clear runs.
struct Slab {
    char *base = nullptr;
    size_t bytes = 0;
    int dev = -1;
};
static std::mutex g_slab_mu;
static Slab g_slab_parked[64];

static size_t slab_parked_bytes() {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_slab_mu);
    return g_slab_parked[dev & 63].bytes;
}

static Slab slab_acquire(size_t bytes) {
    Slab out;
    cudaGetDevice(&out.dev);
    {
        std::lock_guard<std::mutex> lock(g_slab_mu);
        Slab &parked = g_slab_parked[out.dev & 63];
        if (parked.base && parked.bytes >= bytes) {
            out = parked;
            parked = Slab{};
            return out;
        }
        if (parked.base) {
            cudaFree(parked.base);
            parked = Slab{};
        }
    }
    cudaError_t e = cudaMalloc((void **)&out.base, bytes);
    if (e == cudaErrorMemoryAllocation) {
        // the budget counts memory the stream-ordered pool holds in reserve (released buffers of earlier
        // calls) as available: hand it back to the driver and try once more
        cudaGetLastError();
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, out.dev) == cudaSuccess) {
            cudaDeviceSynchronize();
            cudaMemPoolTrimTo(pool, 0);
        }
        e = cudaMalloc((void **)&out.base, bytes);
    }
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        throw OomError("device allocation of " + std::to_string(bytes) + " bytes failed");
    }
    RXC_CUDA(e);
    out.bytes = bytes;
    return out;
}

// the caller has synchronised the stream that used the slab
static void slab_release(Slab &sl) {
    if (!sl.base) return;
    std::lock_guard<std::mutex> lock(g_slab_mu);
    Slab &parked = g_slab_parked[sl.dev & 63];
    int cur = sl.dev;
    cudaGetDevice(&cur);
    if (cur != sl.dev) cudaSetDevice(sl.dev);
    if (parked.base && parked.bytes >= sl.bytes) {
        cudaFree(sl.base);  // keep the larger one
    } else {
        if (parked.base) cudaFree(parked.base);
        parked = sl;
    }
    if (cur != sl.dev) cudaSetDevice(cur);
    sl = Slab{};
}

static void slab_free_parked() {
    std::lock_guard<std::mutex> lock(g_slab_mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (Slab &parked : g_slab_parked) {
        if (!parked.base) continue;
        cudaSetDevice(parked.dev);
        cudaFree(parked.base);
        parked = Slab{};
    }
    cudaSetDevice(cur);
}

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    int dev = -1;  // device the block came from: frees go back to that device's pool
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    bool owned = true;  // false: a view into a Slab (the slab's owner frees it)
    void view(void *at, size_t count) {
        release();
        p = reinterpret_cast<T *>(at);
        n = count;
        owned = false;
    }
    void alloc(size_t count) {
        release();
        if (count) {
            cudaGetDevice(&dev);
            cudaError_t e = cudaMallocAsync((void **)&p, count * sizeof(T), 0);
            if (e == cudaSuccess) e = cudaStreamSynchronize(0);
            if (e == cudaErrorMemoryAllocation && slab_parked_bytes()) {  // a parked scratch slab is in the way
                cudaGetLastError();
                slab_free_parked();
                e = cudaMallocAsync((void **)&p, count * sizeof(T), 0);
                if (e == cudaSuccess) e = cudaStreamSynchronize(0);
            }
            if (e == cudaErrorMemoryAllocation) {
                cudaGetLastError();
                p = nullptr;
                throw OomError("device allocation of " + std::to_string(count * sizeof(T)) + " bytes failed");
            }
            RXC_CUDA(e);
        }
        n = count;
    }
    // Callers release buffers whose consumer stream has been synchronised; on an exception the
    // C-ABI wrapper (`guarded`) synchronises the device before anything can be allocated again.
    void release() {
        if (p && owned) {
            int cur = dev;
            cudaGetDevice(&cur);
            if (cur != dev) cudaSetDevice(dev);
            cudaFreeAsync(p, 0);
            if (cur != dev) cudaSetDevice(cur);
        }
        p = nullptr;
        n = 0;
        owned = true;
    }
    size_t bytes() const { return n * sizeof(T); }
    ~DevBuf() { release(); }
};


struct DeviceCSR {
    DevBuf<uint32_t> row, col, eid;
    DevBuf<uint4> ell;      // one uint4 per vertex when max_degree <= 4 (built on first use)
    int64_t max_degree = -1;  // -1: not computed yet
    uint64_t arcs = 0;
    void upload(const HostCSR &h, cudaStream_t s) {
        row.alloc(h.row.size());
        col.alloc(h.col.size() ? h.col.size() : 1);
        arcs = h.col.size();
        RXC_CUDA(cudaMemcpyAsync(row.p, h.row.data(), h.row.size() * 4, cudaMemcpyHostToDevice, s));
        if (arcs)
            RXC_CUDA(cudaMemcpyAsync(col.p, h.col.data(), arcs * 4, cudaMemcpyHostToDevice, s));
        if (!h.eid.empty()) {
            eid.alloc(h.eid.size());
            RXC_CUDA(cudaMemcpyAsync(eid.p, h.eid.data(), arcs * 4, cudaMemcpyHostToDevice, s));
        }
    }
};

struct Options {
    int64_t bc_groups = 0;    // groups of 32 sources per betweenness batch (0 = auto)
    int64_t bfs_groups = 0;   // groups per closeness / distance batch (0 = auto)
    int64_t mem_fraction_pct = 70;
    int64_t hub_degree = 512;    // vertices above this degree get a whole CTA
    int64_t bfs_mode = 0;        // 0 auto, 1 bit-parallel level launches, 2 CTA-per-source in smem
    int64_t bc_sparse_div = 16;  // forward levels with <= n/this queue entries use the frontier filter (0: off)
    int64_t device_csr = 1;      // build the CSR on the device (0: host counting sort)
    int64_t bfs_depth_switch = 48;  // auto: pilot BFS deeper than this many levels -> mode 2
    int64_t bc_sigma = 0;           // betweenness sigma rows: 0 auto (uint32, f64 on overflow), 1 f64, 2 uint32
    int64_t bfs_threads = 0;        // CTA-per-source BFS: threads per CTA (0: 128 / 256 / 512 / 1024 by n)
    int64_t bfs_ell = 1;            // CTA-per-source BFS: ELL adjacency when every vertex has <= 4 arcs
    int64_t bc_look = 1;            // dependency sweep, look-ahead kernel: 0 never, 1 on leaf-heavy levels, 2 always (node mode)
    int64_t bc_bwd_ctas = 32768;    // dependency sweep: CTAs of one launch, all groups together (0: one warp per 1..8 entries).
                                    // Measured on BA(10^6,8), 16384 sources: 8192 -> 920 ms, 32768 -> 900, 131072 -> 907, 0 -> 915;
                                    // BA(125k,8), 8192 sources: 53.0 / 49.9 / 50.2 / 57.3 ms
    int64_t bc_siblings = 2;        // bc_order 2: a hub named by at least this many sources of the call forms a cluster
    int64_t bc_order = 2;           // betweenness source grouping: 0 off, 1 two-hop key, 2 key + hub siblings for >= 4096 sources
};
static Options g_opt;      // process-wide defaults (rxc_set_option), guarded by g_opt_mu
static std::mutex g_opt_mu;  // every call works on the copy begin_call() takes into its handle
static Options options_now() {
    std::lock_guard<std::mutex> lock(g_opt_mu);
    return g_opt;
}
constexpr uint32_t HUB_CAP = 4096;  // deferred high-degree work items per group and level
constexpr uint32_t HUB_CTAS = 32;   // CTAs per group in the hub kernels
constexpr uint32_t MARK_CTAS = 64;  // CTAs per group in the frontier-marking kernel
static int g_device = 0;
static std::vector<int> g_devices;  // rxc_set_devices: replicas for in-process multi-GPU (empty: single)

static double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// Per-handle scratch in HBM, kept across calls (cudaMalloc/cudaFree of multi-GB buffers would
// otherwise cost more than a batch of kernels).
struct BcWorkspace {
    DevBuf<double> sc;
    DevBuf<uint32_t> sig32, overflow;  // narrow sigma rows (bc_kernels.cuh) and their overflow flag
    DevBuf<PEntry> P0, P1;
    DevBuf<SrcMask> visited, qm, hubm;
    DevBuf<uint2> qe;
    DevBuf<uint32_t> qv, loff, lvlcnt, lvl_total, src, hubv, hubcnt, maybe;
    DevBuf<unsigned long long> reach;
    uint64_t ng = 0;        // capacity in groups
    uint64_t qper = 0;      // level-queue entries per vertex and group the queues were sized for
    uint64_t wanted = 0;    // groups asked for when `ng` was sized (may have been clamped by memory)
    uint32_t tagbase = 0;   // tags already used in P0/P1
    uint32_t dirty_levels = 0;  // prefix of the level tables the last batch wrote
    Slab slab;              // every array above is a view into this one allocation
    void release() {
        sc.release(); sig32.release(); overflow.release(); P0.release(); P1.release(); visited.release(); qv.release(); qm.release(); qe.release();
        loff.release(); lvlcnt.release(); lvl_total.release(); src.release(); reach.release();
        hubv.release(); hubm.release(); hubcnt.release(); maybe.release();
        slab_release(slab);
        ng = 0;
        qper = 0;
        wanted = 0;
    }
};

// Large device-to-host results (distance-matrix rows, path-length rows) go through two pinned staging
// buffers: the DMA engine fills one at PCIe speed while host threads copy the other into the
// caller's pageable buffer.  A plain cudaMemcpyAsync into pageable memory ran at ~4.5 GB/s here and
// was 98 % of distance_matrix's end-to-end time on heavy_hex(51) (333 MB).
struct PinnedStage {
    static constexpr size_t SLOT = 16u << 20;
    char *buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool used[2] = {false, false};  // an event has been recorded for the slot
    ~PinnedStage() {
        for (int i = 0; i < 2; i++) {
            if (buf[i]) cudaFreeHost(buf[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
        }
    }
    void ensure() {
        for (int i = 0; i < 2; i++) {
            if (!buf[i]) RXC_CUDA(cudaHostAlloc((void **)&buf[i], SLOT, cudaHostAllocDefault));
            if (!ev[i]) RXC_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
    }
};

// one staging pair per device, shared by every handle (pinning 64 MB costs more than a snapshot)
struct PinnedSlot {
    PinnedStage stage;
    std::mutex mu;
};
static PinnedSlot &pinned_slot(int device) {
    static PinnedSlot slots[64];
    return slots[device & 63];
}
static PinnedStage &pinned_for(int device) { return pinned_slot(device).stage; }

static void host_copy_parallel(char *dst, const char *src, size_t bytes) {
    const size_t T = bytes >= (8u << 20) ? 8 : (bytes >= (2u << 20) ? 4 : 1);
    if (T == 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t part = (bytes / T + 4095) & ~(size_t)4095;
    for (size_t t = 0; t < T; t++) {
        const size_t b = t * part, e = std::min(bytes, b + part);
        if (b < e) th.emplace_back([=] { memcpy(dst + b, src + b, e - b); });
    }
    for (auto &x : th) x.join();
}

// Page-locked caller buffers (rxc_host_alloc, cudaHostAlloc, cudaHostRegister) are DMA targets
// themselves: no staging, the copy runs at PCIe speed.
static bool host_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// dst: host memory; src: device memory; everything queued on `st` before the call is complete when
// it returns
static void d2h_staged(PinnedStage &ps, void *dst, const void *src, size_t bytes, cudaStream_t st) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(pinned_slot(dev).mu);
    if (bytes < (4u << 20) || host_is_pinned(dst)) {
        RXC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
        RXC_CUDA(cudaStreamSynchronize(st));
        return;
    }
    ps.ensure();
    const size_t nchunks = (bytes + PinnedStage::SLOT - 1) / PinnedStage::SLOT;
    auto issue = [&](size_t c) {
        const size_t off = c * PinnedStage::SLOT, len = std::min(PinnedStage::SLOT, bytes - off);
        RXC_CUDA(cudaMemcpyAsync(ps.buf[c & 1], (const char *)src + off, len, cudaMemcpyDeviceToHost, st));
        RXC_CUDA(cudaEventRecord(ps.ev[c & 1], st));
        ps.used[c & 1] = true;
    };
    issue(0);
    for (size_t c = 0; c < nchunks; c++) {
        if (c + 1 < nchunks) issue(c + 1);  // slot (c+1)&1 was drained in the previous iteration
        RXC_CUDA(cudaEventSynchronize(ps.ev[c & 1]));
        const size_t off = c * PinnedStage::SLOT, len = std::min(PinnedStage::SLOT, bytes - off);
        host_copy_parallel((char *)dst + off, ps.buf[c & 1], len);
    }
    RXC_CUDA(cudaStreamSynchronize(st));
}

// src: host memory; dst: device memory.  Pageable sources: returns with every copy queued on `st`;
// the staging slots are reused only after their DMA has completed (events), so the caller's buffer
// may be released as soon as this returns.  Page-locked sources are copied straight from the
// caller's buffer and only QUEUED: the caller must synchronise `st` before handing the buffer back
// (rxc_graph_snapshot does, at the end of the build).
static void h2d_staged(PinnedStage &ps, void *dst, const void *src, size_t bytes, cudaStream_t st) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (host_is_pinned(src)) {
        RXC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return;
    }
    std::lock_guard<std::mutex> lock(pinned_slot(dev).mu);
    if (bytes < (4u << 20)) {
        RXC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        RXC_CUDA(cudaStreamSynchronize(st));  // pageable source: keep the documented "copy on return"
        return;
    }
    ps.ensure();
    const size_t nchunks = (bytes + PinnedStage::SLOT - 1) / PinnedStage::SLOT;
    for (size_t c = 0; c < nchunks; c++) {
        const size_t off = c * PinnedStage::SLOT, len = std::min(PinnedStage::SLOT, bytes - off);
        if (c >= 2 || ps.used[c & 1]) RXC_CUDA(cudaEventSynchronize(ps.ev[c & 1]));  // slot's last DMA done
        host_copy_parallel(ps.buf[c & 1], (const char *)src + off, len);
        RXC_CUDA(cudaMemcpyAsync((char *)dst + off, ps.buf[c & 1], len, cudaMemcpyHostToDevice, st));
        RXC_CUDA(cudaEventRecord(ps.ev[c & 1], st));
        ps.used[c & 1] = true;
    }
}

struct BfsWorkspace {
    DevBuf<uint32_t> visited, next0, next1, q0, q1, cnt, lvl_total, src;
    DevBuf<unsigned long long> reach, dsum;
    DevBuf<double> vals, rows;
    uint64_t ng = 0, ng_rows = 0, wanted = 0;
    uint32_t dirty_levels = 0;
    void release() {
        visited.release(); next0.release(); next1.release(); q0.release(); q1.release();
        cnt.release(); lvl_total.release(); src.release(); reach.release(); dsum.release();
        vals.release(); rows.release();
        ng = ng_rows = 0;
        wanted = 0;
    }
};

}  // namespace rxc

using namespace rxc;

struct rxc_graph {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // [4],[5]: whole call; [6],[8],[7]: rendezvous, reduce
    HostSnapshot host;
    DeviceCSR out, in, sym;
    bool has_sym = false;
    DevBuf<uint32_t> d_esrc, d_edst;  // compact endpoints of every edge (device build; directed only)
    uint64_t n_arcs = 0;
    DevBuf<unsigned long long> d_stats;
    rxc_stats stats;
    double ms_snapshot = 0.0;
    uint64_t last_bfs_entries = 0;  // queue entries of the bit-parallel batches of this call
    BcWorkspace bc_ws;
    BfsWorkspace bfs_ws;
    DevBuf<double> acc_buf;  // partial centrality vector of the current call
    bool sigma_wide = false;  // a path count overflowed 32 bits on this graph: keep f64 sigma rows
    bool queue_wide = false;  // a level queue outgrew the typical-case size: keep worst-case queues
    std::vector<std::unique_ptr<rxc_graph>> replicas;  // same snapshot on the other devices of rxc_set_devices
    std::mutex mu;
    Options opt;    // the tunables this call runs with: copied from the process defaults by begin_call()
    int sms = 148;  // cudaDevAttrMultiProcessorCount of `device`
    uint32_t last_levels = 0;         // run_levels: levels launched by the previous level loop
    std::vector<uint32_t> order_cnt;  // bc_order_sources: children per hub, all zero between calls
    uint64_t cache_key[2] = {0, 0};  // rxc_graph_snapshot_cached: content hash (0,0: not cached)
    int refs = 1;                    // rxc_graph_snapshot_cached / rxc_graph_free

    const DeviceCSR &succ() const { return out; }
    const DeviceCSR &pred() const { return host.directed ? in : out; }
    ~rxc_graph() {
        replicas.clear();  // each replica switches to its own device while it goes away
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);  // nothing in flight touches the buffers below
        bc_ws.release();
        bfs_ws.release();
        acc_buf.release();
        for (auto &e : ev)
            if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace rxc {

template <class F>
static int guarded(F &&f) {
    // On every error path the device is drained before the status goes back: buffers released while
    // the stack unwound went to the pool with work possibly still in flight on the handle's stream.
    struct Drain {
        bool armed = true;
        ~Drain() {
            if (armed) {
                cudaDeviceSynchronize();
                cudaGetLastError();
            }
        }
    } drain;
    try {
        f();
        drain.armed = false;
        return RXC_OK;
    } catch (const CudaError &e) {
        set_error(std::string("CUDA error: ") + cudaGetErrorString(e.code) + " at " + e.file + ":" +
                  std::to_string(e.line));
        cudaGetLastError();
        return RXC_ERR_CUDA;
    } catch (const OomError &e) {
        set_error(std::string("out of device memory: ") + e.what());
        return RXC_ERR_OOM;
    } catch (const NcclError &e) {
        set_error(e.what());
        return RXC_ERR_NCCL;
    } catch (const std::bad_alloc &) {
        set_error("out of host memory");
        return RXC_ERR_OOM;
    } catch (const LimitError &e) {
        set_error(e.what());
        return RXC_ERR_LIMIT;
    } catch (const std::invalid_argument &e) {
        set_error(std::string("invalid argument: ") + e.what());
        return RXC_ERR_INVALID;
    } catch (const std::exception &e) {
        set_error(e.what());
        return RXC_ERR_INVALID;
    }
}

static void require_device() {
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        throw CudaError{e == cudaSuccess ? cudaErrorNoDevice : e, __FILE__, __LINE__};
    }
}

static size_t device_budget(int64_t mem_fraction_pct) {
    size_t free_b = 0, total_b = 0;
    RXC_CUDA(cudaMemGetInfo(&free_b, &total_b));
    // memory parked in the pool (reserved, not in use) is available to us as well
    int dev = 0;
    cudaGetDevice(&dev);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t reserved = 0, used = 0;
        if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess &&
            reserved > used)
            free_b += (size_t)(reserved - used);
    }
    cudaGetLastError();
    return (size_t)((double)free_b * (double)mem_fraction_pct / 100.0);
}

// ------------------------------------------------------------------------------------------------
// device CSR build (csr_build.cuh)
// ------------------------------------------------------------------------------------------------
static void device_scan(const uint32_t *in, uint32_t *out_row, uint32_t n, DevBuf<uint32_t> &tiles,
                        uint32_t *d_total, cudaStream_t st) {
    const uint32_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles.n < ntiles + 1) tiles.alloc(ntiles + 1);
    scan_tile_kernel<<<std::max(ntiles, 1u), BLOCK, 0, st>>>(in, out_row, n, tiles.p);
    scan_sums_kernel<<<1, BLOCK, 0, st>>>(tiles.p, ntiles, d_total);
    scan_add_kernel<<<std::max(ntiles, 1u), BLOCK, 0, st>>>(out_row, n, tiles.p, d_total);
}

static void device_sort_rows(DeviceCSR &csr, uint32_t n, DevBuf<uint32_t> &long_rows,
                             uint32_t *d_nlong, cudaStream_t st, int sms) {
    if (n == 0 || !csr.eid.p) return;
    static bool attr_done[64] = {false};  // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        RXC_CUDA(cudaFuncSetAttribute(csr_sort_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(CSR_SORT_MAX * 8)));
        attr_done[dev & 63] = true;
    }
    RXC_CUDA(cudaMemsetAsync(d_nlong, 0, 4, st));
    csr_sort_rows_kernel<<<(n + WARPS - 1) / WARPS, BLOCK, 0, st>>>(csr.row.p, n, csr.col.p, csr.eid.p,
                                                                   long_rows.p, d_nlong);
    csr_sort_long_kernel<<<sms, CSR_SORT_THREADS, CSR_SORT_MAX * 8, st>>>(csr.row.p, long_rows.p, d_nlong,
                                                              csr.col.p, csr.eid.p);
}

// mode 0 successors, 1 predecessors, 2 both directions in one structure
static void device_build_csr(DeviceCSR &csr, uint32_t n, const uint32_t *d_src, const uint32_t *d_dst,
                             const uint32_t *d_eid, uint64_t n_edges, int mode,
                             const uint32_t *d_deg, DevBuf<uint32_t> &cursor, DevBuf<uint32_t> &tiles,
                             uint32_t *d_total, cudaStream_t st, int sms) {
    const uint64_t cap = mode == 2 ? 2 * n_edges : n_edges;
    csr.row.alloc((size_t)n + 1);
    csr.col.alloc(cap ? cap : 1);
    if (d_eid) csr.eid.alloc(cap ? cap : 1);
    device_scan(d_deg, csr.row.p, n, tiles, d_total, st);
    RXC_CUDA(cudaMemsetAsync(cursor.p, 0, ((size_t)n + 1) * 4, st));
    if (n_edges) {
        const uint32_t blocks = (uint32_t)std::min<uint64_t>((n_edges + BLOCK - 1) / BLOCK, (uint64_t)sms * 32);
        csr_scatter_kernel<<<blocks, BLOCK, 0, st>>>(d_src, d_dst, d_eid, n_edges, mode, csr.row.p,
                                                     cursor.p, csr.col.p, d_eid ? csr.eid.p : nullptr);
    }
}

static void device_snapshot(rxc_graph *g, uint32_t n_bound, const uint32_t *live_nodes,
                            uint32_t n_live, uint32_t e_bound, const uint32_t *esrc,
                            const uint32_t *edst, const uint32_t *eid, uint64_t n_edges,
                            bool directed) {
    HostSnapshot &h = g->host;
    const bool trace = getenv("RXC_TRACE") != nullptr;
    const double tr0 = now_ms();
    auto tp = [&](const char *what) {
        if (trace) {
            cudaStreamSynchronize(g->stream);
            fprintf(stderr, "[rxc snapshot] %-12s %.2f ms\n", what, now_ms() - tr0);
        }
    };
    h.n_bound = n_bound;
    h.n_live = n_live;
    h.e_bound = e_bound;
    h.n_edges = n_edges;
    h.directed = directed;
    if ((directed ? n_edges : 2 * n_edges) >= 0xFFFFFFFFull)
        throw LimitError("more than 2^32-2 arcs are not supported");
    cudaStream_t st = g->stream;
    const uint32_t n = n_live;
    DevBuf<uint32_t> d_live, d_cmap, d_eid, d_deg_a, d_deg_b, cursor, tiles, d_small, long_rows;
    DevBuf<uint32_t> &d_src = g->d_esrc, &d_dst = g->d_edst;
    d_live.alloc(std::max<uint32_t>(n_live, 1));
    d_cmap.alloc(std::max<uint32_t>(n_bound, 1));
    d_src.alloc(std::max<uint64_t>(n_edges, 1));
    d_dst.alloc(std::max<uint64_t>(n_edges, 1));
    d_eid.alloc(std::max<uint64_t>(n_edges, 1));
    d_deg_a.alloc((size_t)n + 1);
    if (directed) d_deg_b.alloc((size_t)n + 1);
    cursor.alloc((size_t)n + 1);
    long_rows.alloc(std::max<uint32_t>(n, 1));
    d_small.alloc(4);  // [0] error flags, [1] scan total, [2] long-row count
    RXC_CUDA(cudaMemsetAsync(d_small.p, 0, 16, st));
    tp("allocs");
    if (n_live) RXC_CUDA(cudaMemcpyAsync(d_live.p, live_nodes, (size_t)n_live * 4, cudaMemcpyHostToDevice, st));
    if (n_edges) {
        // the 12 bytes per edge go through the pinned staging slots (a pageable cudaMemcpy of the
        // 100 MB of config 3 took 8 of the snapshot's 13 ms)
        h2d_staged(pinned_for(g->device), d_src.p, esrc, n_edges * 4, st);
        h2d_staged(pinned_for(g->device), d_dst.p, edst, n_edges * 4, st);
        if (eid) {
            h2d_staged(pinned_for(g->device), d_eid.p, eid, n_edges * 4, st);
        } else {  // edge ids default to the position in the list: nothing to copy
            const uint32_t blocks = (uint32_t)std::min<uint64_t>((n_edges + BLOCK - 1) / BLOCK, (uint64_t)g->sms * 32);
            csr_iota_kernel<<<blocks, BLOCK, 0, st>>>(d_eid.p, n_edges);
        }
    }
    tp("h2d");
    // host-side checks and index maps, while the DMA engine works on the copies queued above
    bool ascending = true, in_range = true;
    for (uint32_t i = 0; i < n_live; i++) {
        in_range &= live_nodes[i] < n_bound;
        ascending &= i == 0 || live_nodes[i - 1] < live_nodes[i];
    }
    if (!in_range) throw std::invalid_argument("live node index >= n_bound");
    if (!ascending) throw std::invalid_argument("live_nodes must be strictly ascending");
    // strictly ascending, below n_bound, n_bound of them: the identity -- no host maps needed
    h.identity = n_live == n_bound;
    if (!h.identity) {
        h.orig_of_compact.assign(live_nodes, live_nodes + n_live);
        h.compact_of_orig.assign(n_bound, 0xFFFFFFFFu);
        for (uint32_t i = 0; i < n_live; i++) h.compact_of_orig[live_nodes[i]] = i;
    }
    tp("host maps");
    RXC_CUDA(cudaMemsetAsync(d_cmap.p, 0xFF, d_cmap.bytes(), st));
    RXC_CUDA(cudaMemsetAsync(d_deg_a.p, 0, d_deg_a.bytes(), st));
    if (directed) RXC_CUDA(cudaMemsetAsync(d_deg_b.p, 0, d_deg_b.bytes(), st));
    if (n_live)
        csr_map_kernel<<<(n_live + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(d_live.p, n_live, n_bound, d_cmap.p, d_small.p);
    const uint32_t *d_eid_p = d_eid.p;
    if (n_edges) {
        const uint32_t blocks = (uint32_t)std::min<uint64_t>((n_edges + BLOCK - 1) / BLOCK, (uint64_t)g->sms * 32);
        csr_count_kernel<<<blocks, BLOCK, 0, st>>>(d_src.p, d_dst.p, d_eid_p, n_edges, d_cmap.p, n_bound,
                                                   e_bound, d_deg_a.p, directed ? d_deg_b.p : d_deg_a.p,
                                                   d_small.p);
    }
    if (directed) {
        device_build_csr(g->out, n, d_src.p, d_dst.p, d_eid_p, n_edges, 0, d_deg_a.p,
                         cursor, tiles, d_small.p + 1, st, g->sms);
        device_build_csr(g->in, n, d_src.p, d_dst.p, d_eid_p, n_edges, 1, d_deg_b.p,
                         cursor, tiles, d_small.p + 1, st, g->sms);
    } else {
        device_build_csr(g->out, n, d_src.p, d_dst.p, d_eid_p, n_edges, 2, d_deg_a.p,
                         cursor, tiles, d_small.p + 1, st, g->sms);
    }
    tp("count+build");
    device_sort_rows(g->out, n, long_rows, d_small.p + 2, st, g->sms);
    if (directed) device_sort_rows(g->in, n, long_rows, d_small.p + 2, st, g->sms);
    uint32_t h_err = 0, h_arcs = 0;
    RXC_CUDA(cudaMemcpyAsync(&h_err, d_small.p, 4, cudaMemcpyDeviceToHost, st));
    RXC_CUDA(cudaMemcpyAsync(&h_arcs, g->out.row.p + n, 4, cudaMemcpyDeviceToHost, st));
    RXC_CUDA(cudaStreamSynchronize(st));
    RXC_CUDA(cudaGetLastError());
    tp("sort");
    if (h_err & CSR_ERR_ENDPOINT_RANGE) throw std::invalid_argument("edge endpoint >= n_bound");
    if (h_err & CSR_ERR_ENDPOINT_DEAD) throw std::invalid_argument("edge endpoint is not a live node");
    if (h_err & CSR_ERR_EDGE_ID) throw std::invalid_argument("edge index >= e_bound");
    g->out.arcs = h_arcs;
    g->in.arcs = directed ? h_arcs : 0;
    g->n_arcs = h_arcs;
    if (!directed) {  // the symmetric structure is `out` itself
        g->d_esrc.release();
        g->d_edst.release();
    }
}

static void reset_stats(rxc_graph *g) {
    memset(&g->stats, 0, sizeof(g->stats));
    g->stats.ms_snapshot = g->ms_snapshot;
    RXC_CUDA(cudaMemsetAsync(g->d_stats.p, 0, g->d_stats.bytes(), g->stream));
}

static void begin_call(rxc_graph *g) {
    RXC_CUDA(cudaSetDevice(g->device));
    g->opt = options_now();  // a concurrent rxc_set_option never changes a running call
    g->last_bfs_entries = 0;
    reset_stats(g);
    RXC_CUDA(cudaEventRecord(g->ev[4], g->stream));
}

// ms_total = device time from the first to the last kernel of the call, host gaps included
static void end_call(rxc_graph *g) {
    RXC_CUDA(cudaEventRecord(g->ev[5], g->stream));
    RXC_CUDA(cudaEventSynchronize(g->ev[5]));
    float ms = 0.f;
    RXC_CUDA(cudaEventElapsedTime(&ms, g->ev[4], g->ev[5]));
    g->stats.ms_total = ms;
}

static void fetch_stats(rxc_graph *g) {
    unsigned long long h[ST_STRIPES * ST_COUNT];
    RXC_CUDA(cudaMemcpyAsync(h, g->d_stats.p, sizeof(h), cudaMemcpyDeviceToHost, g->stream));
    RXC_CUDA(cudaStreamSynchronize(g->stream));
    unsigned long long sum[ST_COUNT] = {0, 0, 0, 0};
    for (int s = 0; s < ST_STRIPES; s++)
        for (int c = 0; c < ST_COUNT; c++) sum[c] += h[s * ST_COUNT + c];
    g->stats.te = sum[ST_TE];
    g->stats.te_dag = sum[ST_DAG];
    g->stats.v = sum[ST_V];
#if RXC_PROBE
    unsigned long long pr[PR_COUNT], zero[PR_COUNT] = {};
    RXC_CUDA(cudaMemcpyFromSymbol(pr, g_probe, sizeof(pr)));
    RXC_CUDA(cudaMemcpyToSymbol(g_probe, zero, sizeof(zero)));
    fprintf(stderr,
            "[rxc probe] uint32 rows (forward): lookups %llu rows %llu sectors %llu lines128 %llu | f64 rows "
            "(dependency sweep; forward when wide): lookups %llu rows %llu sectors %llu lines128 %llu\n",
            pr[0], pr[1], pr[2], pr[3], pr[4], pr[5], pr[6], pr[7]);
#endif
}

// Split a source list into rounds of pairwise-distinct compact ids (a vertex listed twice must
// not share a 32-source group with itself; contributions are additive, so extra rounds are exact).
static std::vector<std::vector<uint32_t>> distinct_rounds(const std::vector<uint32_t> &src,
                                                          uint32_t n) {
    std::vector<std::vector<uint32_t>> rounds;
    if (src.empty()) return rounds;
    if (src.size() <= n / 16) {  // short list: a sort is cheaper than touching an n-sized table
        std::vector<uint32_t> sorted(src);
        std::sort(sorted.begin(), sorted.end());
        if (std::adjacent_find(sorted.begin(), sorted.end()) == sorted.end()) {
            rounds.push_back(src);
            return rounds;
        }
    }
    std::vector<uint32_t> seen(n, 0);
    for (uint32_t s : src) {
        const uint32_t r = seen[s]++;
        if (rounds.size() <= r) rounds.emplace_back();
        rounds[r].push_back(s);
    }
    return rounds;
}

// Launch `launch(level)` for level = 0, 1, ... in growing speculative chunks and return the number
// of BFS levels (first level whose total frontier is empty).  lvl_total[l] is final once the launch
// for level l-1 has completed; launches past the end find empty frontiers and return at once.
// The first chunk is what the previous level loop on the handle needed, plus one: batches of one call
// have the same depth give or take a level, so the usual batch costs one chunk, one read-back and
// one launch past its last level.
template <class F>
static uint32_t run_levels(rxc_graph *g, uint32_t *d_lvl_total, uint32_t n, F &&launch,
                           std::vector<uint32_t> &h_totals) {
    uint32_t level = 0, checked = 1;
    const bool adaptive = g->last_levels != 0;
    uint32_t chunk = adaptive ? std::min<uint32_t>(g->last_levels + 1, 64) : 4;
    const uint32_t first = chunk;
    h_totals.assign(1, 0);
    for (;;) {
        uint32_t launched = 0;
        while (launched < chunk && level < n) {
            launch(level);
            level++;
            launched++;
        }
        h_totals.resize((size_t)level + 1);
        RXC_CUDA(cudaMemcpyAsync(h_totals.data() + checked, d_lvl_total + checked,
                                 (size_t)(level + 1 - checked) * 4, cudaMemcpyDeviceToHost,
                                 g->stream));
        RXC_CUDA(cudaStreamSynchronize(g->stream));
        for (uint32_t l = checked; l <= level; l++) {
            if (h_totals[l] == 0) {
                g->last_levels = l;
                return l;
            }
        }
        checked = level + 1;
        if (level >= n) {
            g->last_levels = level;
            return level;
        }
        // the guess was short: carry on in small chunks that double
        chunk = (adaptive && checked == level + 1 && level == first) ? 4 : std::min<uint32_t>(chunk * 2, 64);
    }
}

// ------------------------------------------------------------------------------------------------
// betweenness (node / edge)
// ------------------------------------------------------------------------------------------------
static void bc_enable_smem() {
    static bool done_dev[64] = {false};  // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    bool &done = done_dev[dev & 63];
    if (done) return;
    auto attr = [&](const void *fn) {
        RXC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      std::max(std::max(BC_SMEM_BYTES, BC_FWD32_SMEM_BYTES), std::max(BC_BWD_SMEM_BYTES, BC_BWD_LOOK_SMEM_BYTES))));
    };
    attr((const void *)bc_forward_kernel<false>);
    attr((const void *)bc_forward_kernel<true>);
    attr((const void *)bc_forward_hub_kernel<false>);
    attr((const void *)bc_forward_hub_kernel<true>);
    attr((const void *)bc_backward_kernel<true, false, false>);
    attr((const void *)bc_backward_kernel<false, true, false>);
    attr((const void *)bc_backward_kernel<false, false, false>);
    attr((const void *)bc_backward_kernel<true, false, true>);
    attr((const void *)bc_backward_kernel<false, true, true>);
    attr((const void *)bc_backward_kernel<false, false, true>);
    attr((const void *)bc_backward_look_kernel<false, false>);
    attr((const void *)bc_backward_look_kernel<false, true>);
    attr((const void *)bc_backward_look_kernel<true, false>);
    attr((const void *)bc_backward_look_kernel<true, true>);
    attr((const void *)bc_backward_hub_kernel<true, false, false>);
    attr((const void *)bc_backward_hub_kernel<false, true, false>);
    attr((const void *)bc_backward_hub_kernel<false, false, false>);
    attr((const void *)bc_backward_hub_kernel<true, false, true>);
    attr((const void *)bc_backward_hub_kernel<false, true, true>);
    attr((const void *)bc_backward_hub_kernel<false, false, true>);
    done = true;
}

// Size (or re-use) the handle's betweenness scratch for up to `total_groups` groups of 128 sources;
// returns the number of groups a batch can hold.  Shared by the weighted shortest-path driver.
static uint64_t bc_ws_ensure(rxc_graph *g, uint64_t total_groups) {
    const uint32_t n = g->host.n_live;
    const uint32_t lcap = n + 2;
    // Level queues: a vertex enters a group's queue once per DISTINCT level it has among the 128
    // sources -- 128 in the worst case (a path graph), 3-5 on small-world graphs.  The queues are
    // sized for 16 and the kernels flag an overflow, after which the handle keeps worst-case queues
    // (mapping 33 GB of mostly untouched queue memory cost 2.4 s on the first call at n = 10^6).
    const uint64_t qper = g->queue_wide ? BG : std::min<uint64_t>(BG, 16);
    const uint64_t qcap = qper * n;
    if (qcap >= 0xFFFFFFFFull)  // queue offsets are 32-bit (unreachable below ~140 GB per group)
        throw LimitError("level queues beyond 2^32 entries per group are not supported");
    // bytes per group: f64 + uint32 rows, queues, two tagged mask arrays, visited, level tables
    const double per_group = (double)n * ((double)BG * 12.0 + (double)(12 + sizeof(SrcMask)) * (double)qper +
                                          2.0 * sizeof(PEntry) + (double)sizeof(SrcMask)) +
                             (double)lcap * 8.0 + 4096.0;
    // default batch: 2048 sources at n = 2^20 (24 GB of rows), fewer on larger graphs
    uint64_t ng = g->opt.bc_groups > 0 ? (uint64_t)g->opt.bc_groups
                                      : std::max<uint64_t>(1, ((16ull << 20) / std::max<uint32_t>(n, 1)) * 128 / BG);
    ng = std::min<uint64_t>(ng, 64);
    ng = std::min<uint64_t>(ng, total_groups);
    BcWorkspace &ws = g->bc_ws;
    cudaStream_t st = g->stream;
    if (ws.qper < qper) ws.release();  // queues too small for this handle now
    if (ws.ng > 0 && ws.ng < ng && ws.wanted >= ng) ng = ws.ng;  // already clamped by memory for this request
    if (ws.ng < ng) {
        const double t_alloc0 = now_ms();
        struct TraceAlloc {
            double t0;
            ~TraceAlloc() {
                if (getenv("RXC_TRACE")) fprintf(stderr, "[rxc workspace] alloc %.1f ms\n", now_ms() - t0);
            }
        } trace_alloc{t_alloc0};
        ws.release();
        g->bfs_ws.release();  // one scratch set at a time keeps the footprint predictable
        ws.wanted = ng;
        // (a parked slab is ours to re-use or to free: it counts as available)
        ng = std::min<uint64_t>(ng, std::max<uint64_t>(1, (uint64_t)(((double)device_budget(g->opt.mem_fraction_pct) +
                                                                       (double)slab_parked_bytes()) / per_group)));
        // one slab, carved into 256-byte aligned arrays (two passes: measure, then place)
        size_t off = 0;
        char *base = nullptr;
        auto carve = [&](auto &buf, size_t count) {
            using Elem = std::remove_pointer_t<decltype(buf.p)>;
            if (base) buf.view(base + off, count);
            off = (off + count * sizeof(Elem) + 255) & ~(size_t)255;
        };
        auto carve_all = [&]() {
            off = 0;
            carve(ws.sc, (size_t)ng * n * BG);
            carve(ws.sig32, (size_t)ng * n * BG);
            carve(ws.overflow, 1);
            carve(ws.P0, (size_t)ng * n);
            carve(ws.P1, (size_t)ng * n);
            carve(ws.visited, (size_t)ng * n);
            carve(ws.qv, (size_t)ng * qcap);
            carve(ws.qm, (size_t)ng * qcap);
            carve(ws.qe, (size_t)ng * qcap);
            carve(ws.loff, (size_t)ng * lcap);
            carve(ws.lvlcnt, (size_t)ng * lcap);
            carve(ws.lvl_total, lcap);
            carve(ws.src, (size_t)ng * BG);
            carve(ws.reach, (size_t)ng * BG);
            carve(ws.hubv, (size_t)ng * HUB_CAP);
            carve(ws.hubm, (size_t)ng * HUB_CAP);
            carve(ws.hubcnt, (size_t)2 * ng);
            carve(ws.maybe, (size_t)ng * ((n + 31) / 32));
        };
        carve_all();
        ws.slab = slab_acquire(off);
        base = ws.slab.base;
        carve_all();
        RXC_CUDA(cudaMemsetAsync(ws.maybe.p, 0, ws.maybe.bytes(), st));
        ws.ng = ng;
        ws.qper = qper;
        ws.tagbase = 0;
        ws.dirty_levels = lcap;
        RXC_CUDA(cudaMemsetAsync(ws.P0.p, 0, ws.P0.bytes(), st));
        RXC_CUDA(cudaMemsetAsync(ws.P1.p, 0, ws.P1.bytes(), st));
    }
    return std::min<uint64_t>(ng, ws.ng);
}

// Which sources share a group -- and, inside a group, a DRAM sector -- is free: contributions are
// additive over sources.  A row gather moves the sectors in which ANY of a lane's four sources needs
// the arc, and a vertex enters a group's level queue once per DISTINCT level it has among the 128
// sources, so sources whose searches reach the bulk of the graph at the same level cost fewer
// gathered rows, fewer queue entries and fewer bytes per useful value.  The sources are therefore
// sorted by the size of their two-hop neighbourhood (bc_source_keys_kernel; a proxy for the mean BFS
// distance, correlation -0.985 on BA graphs); consecutive sources of the sorted list become the four
// sources of one lane (run_batch).  Simulated on BA(200k,8), 512 random sources: 14 % fewer gathered
// rows, 9 % fewer bytes per useful value in the dependency sweep, 12 % in the forward pass.
static std::vector<uint32_t> bc_order_sources(rxc_graph *g, const std::vector<uint32_t> &sources) {
    const size_t k = sources.size();
    if (!g->opt.bc_order || k <= (size_t)BK) return sources;
    cudaStream_t st = g->stream;
    DevBuf<uint32_t> d_src, d_hubs;
    DevBuf<unsigned long long> d_keys;
    d_src.alloc(k);
    d_keys.alloc(k);
    RXC_CUDA(cudaMemcpyAsync(d_src.p, sources.data(), k * 4, cudaMemcpyHostToDevice, st));
    bc_source_keys_kernel<<<(uint32_t)((k + WARPS - 1) / WARPS), BLOCK, 0, st>>>(
        g->succ().row.p, g->succ().col.p, d_src.p, (uint32_t)k, d_keys.p);
    g->stats.launches++;
    std::vector<unsigned long long> keys(k);
    RXC_CUDA(cudaMemcpyAsync(keys.data(), d_keys.p, k * 8, cudaMemcpyDeviceToHost, st));
    // bc_order = 2 (default): BFS-tree siblings -- children of the same hub agree on the level of
    // almost every vertex -- share a lane.  Every source names its highest-degree neighbour; a hub
    // named by at least `bc_siblings` sources of the call forms a cluster, clusters come first (by hub,
    // then key), everything else follows in key order.  On a small random pool no hub has that many
    // children and this IS the key order; on a full run nearly every source sits in a cluster.  Ranks
    // of a multi-process run take contiguous slices of this order (rxc_betweenness_source_order):
    // ordering a slice again finds the same clusters, except one cut small by a slice boundary.
    // Measured on BA(10^6,8): 16384 sources 976 -> 951 ms, 65536 sources 3911 -> 3667 ms.
    std::vector<uint32_t> hubs;
    if (g->opt.bc_order == 2) {
        d_hubs.alloc(k);
        bc_source_hubs_kernel<<<(uint32_t)((k + WARPS - 1) / WARPS), BLOCK, 0, st>>>(
            g->succ().row.p, g->succ().col.p, d_src.p, (uint32_t)k, d_hubs.p);
        g->stats.launches++;
        hubs.resize(k);
        RXC_CUDA(cudaMemcpyAsync(hubs.data(), d_hubs.p, k * 4, cudaMemcpyDeviceToHost, st));
    }
    RXC_CUDA(cudaStreamSynchronize(st));  // also: before the device buffers go back to the pool
    const uint32_t SIBLINGS = (uint32_t)std::max<int64_t>(1, g->opt.bc_siblings);
    // Host side of the ordering, written for k = 10^6 (a full run orders every vertex): children per
    // hub are counted in a table kept in the handle (zeroed again below: O(k), not O(n)), the sort
    // runs on packed 16-byte records -- (cluster, inverted key, source), plain lexicographic order --
    // and on up to 8 threads with pairwise merges.  The first version (a binary search per source
    // for its cluster size, an index sort through three arrays) took 0.85 s for 10^6 sources on one
    // core -- 5 % of the whole job's wall clock on 8 GPUs; this one 0.1 s.
    struct Rec {
        uint32_t hub, src;
        unsigned long long invkey;
    };
    auto less = [](const Rec &a, const Rec &b) {
        if (a.hub != b.hub) return a.hub < b.hub;  // 0xFFFFFFFF (no cluster) last
        if (a.invkey != b.invkey) return a.invkey < b.invkey;  // larger key first
        return a.src < b.src;
    };
    std::vector<Rec> rec(k);
    if (!hubs.empty()) {  // hubs with too few children in this call do not form a cluster
        std::vector<uint32_t> &cnt = g->order_cnt;
        if (cnt.size() < g->host.n_live) cnt.assign(g->host.n_live, 0);
        for (size_t i = 0; i < k; i++) cnt[hubs[i]]++;
        for (size_t i = 0; i < k; i++)
            rec[i] = Rec{cnt[hubs[i]] < SIBLINGS ? 0xFFFFFFFFu : hubs[i], sources[i], ~keys[i]};
        for (size_t i = 0; i < k; i++) cnt[hubs[i]] = 0;
    } else {
        for (size_t i = 0; i < k; i++) rec[i] = Rec{0xFFFFFFFFu, sources[i], ~keys[i]};
    }
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int T = k < 32768 ? 1 : (int)std::min<unsigned>(8, hw);
    if (T == 1) {
        std::sort(rec.begin(), rec.end(), less);
    } else {
        std::vector<size_t> cut(T + 1);
        for (int t = 0; t <= T; t++) cut[t] = k * (size_t)t / T;
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t] { std::sort(rec.begin() + cut[t], rec.begin() + cut[t + 1], less); });
        for (auto &x : th) x.join();
        for (int w = 1; w < T; w *= 2) {
            th.clear();
            for (int t = 0; t + w < T; t += 2 * w)
                th.emplace_back([&, t, w] {
                    std::inplace_merge(rec.begin() + cut[t], rec.begin() + cut[t + w],
                                       rec.begin() + cut[std::min(T, t + 2 * w)], less);
                });
            for (auto &x : th) x.join();
        }
    }
    std::vector<uint32_t> out(k);
    for (size_t i = 0; i < k; i++) out[i] = rec[i].src;
    return out;
}

// Group betweenness (bc_kernels.cuh, "pair mode"): the group's vertices on the device.
struct GroupSpec {
    const uint32_t *d_members;  // compact ids, distinct
    uint32_t n_members;
    const uint32_t *d_blocked;  // bitmap over compact ids
};

static void bc_run(rxc_graph *g, const std::vector<uint32_t> &sources_in, bool endpoints,
                   bool edge_mode, double *d_acc, const GroupSpec *grp = nullptr) {
    bc_enable_smem();
    const uint32_t n = g->host.n_live;
    if (sources_in.empty() || n == 0) return;
    // pair mode keeps the caller's order (its slot layout is fixed by the pairing)
    const std::vector<uint32_t> sources = grp ? sources_in : bc_order_sources(g, sources_in);
    const uint32_t per = grp ? BGP : BG;  // sources per group (pair mode: two slots per source)
    const uint64_t total_groups = (sources.size() + per - 1) / per;
    const uint32_t lcap = n + 2;
    const uint64_t ng = bc_ws_ensure(g, total_groups);
    BcWorkspace &ws = g->bc_ws;
    const uint64_t qcap = ws.qper * n;
    cudaStream_t st = g->stream;

    BcParams p{};
    p.sc = ws.sc.p;
    p.sig32 = ws.sig32.p;
    p.overflow = ws.overflow.p;
    p.P0 = ws.P0.p;
    p.P1 = ws.P1.p;
    p.visited = ws.visited.p;
    p.qv = ws.qv.p;
    p.qm = ws.qm.p;
    p.qe = ws.qe.p;
    p.loff = ws.loff.p;
    p.lvlcnt = ws.lvlcnt.p;
    p.lvl_total = ws.lvl_total.p;
    p.reach = ws.reach.p;
    p.stats = g->d_stats.p;
    p.acc = d_acc;
    p.src = ws.src.p;
    p.hubv = ws.hubv.p;
    p.hubm = ws.hubm.p;
    p.hubcnt = ws.hubcnt.p;
    p.hubcap = HUB_CAP;
    p.maybe = ws.maybe.p;
    p.mwords = (n + 31) / 32;
    p.sparse_max = (uint32_t)std::min<int64_t>(g->opt.bc_sparse_div > 0 ? n / g->opt.bc_sparse_div : 0, 0x7FFFFFFF);
    p.srow = g->succ().row.p;
    p.scol = g->succ().col.p;
    p.hub_degree = (uint32_t)std::min<int64_t>(g->opt.hub_degree, 0x7FFFFFFF);
    p.blocked = grp ? grp->d_blocked : nullptr;
    p.qcap = qcap;
    p.n = n;
    p.lcap = lcap;
    uint32_t &tagbase = ws.tagbase;

    std::vector<uint32_t> h_src((size_t)ng * BG), h_totals, h_lvlcnt;
    const uint32_t vblocks = (n + BLOCK - 1) / BLOCK;
    uint32_t &clear_levels = ws.dirty_levels;  // prefix of the level tables the previous batch dirtied

    // One batch; false when a path count overflowed the narrow sigma rows (nothing accumulated yet).
    auto run_batch = [&](uint64_t g0, bool narrow) -> int {
        const rxc_stats stats_before = g->stats;
        const uint32_t bg = (uint32_t)std::min<uint64_t>(ng, total_groups - g0);
        std::fill(h_src.begin(), h_src.end(), NO_SOURCE);
        const size_t s0 = (size_t)g0 * per;
        const size_t cnt = std::min(sources.size() - s0, (size_t)bg * per);
        if (!grp) {
            // slot 32k + l is source k of lane l: consecutive (similar) sources share a lane's sector
            for (size_t i = 0; i < cnt; i++)
                h_src[(i / BG) * BG + (i % BK) * 32 + (i % BG) / BK] = sources[s0 + i];
        } else {  // slot j and slot 64+j of a group carry the same source
            for (size_t i = 0; i < cnt; i++) {
                const size_t slot = (i / BGP) * BG + i % BGP;
                h_src[slot] = h_src[slot + BGP] = sources[s0 + i];
            }
        }
        if ((uint64_t)tagbase + lcap + 16 > 0xFFFFFFF0ull) {
            RXC_CUDA(cudaMemsetAsync(ws.P0.p, 0, ws.P0.bytes(), st));
            RXC_CUDA(cudaMemsetAsync(ws.P1.p, 0, ws.P1.bytes(), st));
            tagbase = 0;
        }
        p.tagbase = tagbase;
        p.ngroups = bg;
        RXC_CUDA(cudaEventRecord(g->ev[0], st));
        RXC_CUDA(cudaMemcpyAsync(ws.src.p, h_src.data(), (size_t)bg * BG * 4,
                                 cudaMemcpyHostToDevice, st));
        RXC_CUDA(cudaMemset2DAsync(ws.lvlcnt.p, (size_t)lcap * 4, 0, (size_t)clear_levels * 4, ws.ng, st));
        RXC_CUDA(cudaMemsetAsync(ws.lvl_total.p, 0, (size_t)clear_levels * 4, st));
        RXC_CUDA(cudaMemsetAsync(ws.overflow.p, 0, 4, st));
        bc_init_kernel<<<dim3(vblocks, bg), BLOCK, 0, st>>>(p);
        if (grp) {
            if (grp->n_members)
                bc_group_block_kernel<<<dim3((grp->n_members + BLOCK - 1) / BLOCK, bg), BLOCK, 0, st>>>(
                    p, grp->d_members, grp->n_members);
            if (narrow) bc_seed_pair_kernel<true><<<bg, 32, 0, st>>>(p);
            else bc_seed_pair_kernel<false><<<bg, 32, 0, st>>>(p);
            g->stats.launches++;
        } else if (narrow) {
            bc_seed_kernel<true><<<bg, 32, 0, st>>>(p);
        } else {
            bc_seed_kernel<false><<<bg, 32, 0, st>>>(p);
        }
        g->stats.launches += 2;

        // forward: in-arcs (pull from predecessors)
        p.row = g->pred().row.p;
        p.col = g->pred().col.p;
        p.eid = nullptr;
        uint32_t launched_levels = 0;
        const uint32_t L = run_levels(
            g, ws.lvl_total.p, n,
            [&](uint32_t level) {
                if (p.sparse_max) bc_mark_kernel<<<dim3(MARK_CTAS, bg), BLOCK, 0, st>>>(p, level);
                if (narrow) {
                    bc_forward_kernel<true><<<dim3(vblocks, bg), BLOCK, BC_FWD32_SMEM_BYTES, st>>>(p, level);
                    bc_forward_hub_kernel<true><<<dim3(HUB_CTAS, bg), BLOCK, BC_FWD32_SMEM_BYTES, st>>>(p, level);
                } else {
                    bc_forward_kernel<false><<<dim3(vblocks, bg), BLOCK, BC_SMEM_BYTES, st>>>(p, level);
                    bc_forward_hub_kernel<false><<<dim3(HUB_CTAS, bg), BLOCK, BC_SMEM_BYTES, st>>>(p, level);
                }
                launched_levels++;
            },
            h_totals);
        g->stats.launches += (p.sparse_max ? 3 : 2) * launched_levels;
        g->stats.levels += L;
        clear_levels = std::min<uint32_t>(lcap, launched_levels + 2);
        RXC_CUDA(cudaMemsetAsync(ws.hubcnt.p, 0, ws.hubcnt.bytes(), st));
        RXC_CUDA(cudaEventRecord(g->ev[1], st));
        // the forward pass only wrote scratch: on overflow of the 32-bit path counts nothing has
        // been accumulated yet and the caller redoes the batch with f64 sigma rows
        uint32_t h_overflow = 0;
        RXC_CUDA(cudaMemcpyAsync(&h_overflow, ws.overflow.p, 4, cudaMemcpyDeviceToHost, st));

        if (grp) {  // no dependency sweep: the sums come straight from the two sigma rows
            RXC_CUDA(cudaStreamSynchronize(st));
            if (h_overflow) {
                g->stats = stats_before;
                tagbase += launched_levels + 8;
                return (int)h_overflow;
            }
            if (narrow) bc_group_finalize_kernel<true><<<dim3(g->sms * 2, bg), BLOCK, 0, st>>>(p, std::min(L + 1, lcap));
            else bc_group_finalize_kernel<false><<<dim3(g->sms * 2, bg), BLOCK, 0, st>>>(p, std::min(L + 1, lcap));
            g->stats.launches++;
            RXC_CUDA(cudaEventRecord(g->ev[2], st));
            RXC_CUDA(cudaStreamSynchronize(st));
            RXC_CUDA(cudaGetLastError());
            float f = 0.f, b = 0.f;
            RXC_CUDA(cudaEventElapsedTime(&f, g->ev[0], g->ev[1]));
            RXC_CUDA(cudaEventElapsedTime(&b, g->ev[1], g->ev[2]));
            g->stats.ms_forward += f;
            g->stats.ms_backward += b;
            g->stats.batches++;
            g->stats.sources += cnt;
            tagbase += launched_levels + 8;
            return 0;
        }

        // backward: out-arcs (pull from successors), deepest level first
        h_lvlcnt.resize((size_t)bg * L);
        RXC_CUDA(cudaMemcpy2DAsync(h_lvlcnt.data(), (size_t)L * 4, ws.lvlcnt.p, (size_t)lcap * 4,
                                   (size_t)L * 4, bg, cudaMemcpyDeviceToHost, st));
        RXC_CUDA(cudaStreamSynchronize(st));
        if (h_overflow) {
            g->stats = stats_before;
            tagbase += launched_levels + 8;
            return (int)h_overflow;
        }
        p.row = g->succ().row.p;
        p.col = g->succ().col.p;
        p.eid = g->succ().eid.p;
        // level 0 (the sources) adds nothing to node scores but is swept for the work counters
        for (uint32_t l = L; l-- > 0;) {
            uint32_t maxcnt = 0;
            for (uint32_t gi = 0; gi < bg; gi++)
                maxcnt = std::max(maxcnt, h_lvlcnt[(size_t)gi * L + l]);
            if (maxcnt == 0) continue;
            uint32_t gx = std::min<uint32_t>((maxcnt + WARPS - 1) / WARPS, 16384);
            if (g->opt.bc_bwd_ctas > 0)  // cap on the CTAs of one launch, all groups together
                gx = std::min<uint32_t>(gx, (uint32_t)std::max<int64_t>(1, g->opt.bc_bwd_ctas / bg));
            const dim3 grid(gx, bg), hgrid(HUB_CTAS, bg);
            // leaf-heavy level (the next level has less than a quarter of this one's entries): most
            // entries gather nothing and are bound by their chain of round trips -> look-ahead kernel
            uint64_t cnt_here = 0, cnt_next = 0;
            for (uint32_t gi = 0; gi < bg; gi++) {
                cnt_here += h_lvlcnt[(size_t)gi * L + l];
                if (l + 1 < L) cnt_next += h_lvlcnt[(size_t)gi * L + l + 1];
            }
            const bool look = !edge_mode && (g->opt.bc_look == 2 || (g->opt.bc_look == 1 && cnt_next * 4 < cnt_here &&
                                                                     cnt_here >= (uint64_t)g->sms * 256));
#define RXC_BWD(E, P, N)                                                                \
    do {                                                                                \
        if (look) bc_backward_look_kernel<P, N><<<grid, BLOCK, BC_BWD_LOOK_SMEM_BYTES, st>>>(p, l); \
        else bc_backward_kernel<E, P, N><<<grid, BLOCK, BC_BWD_SMEM_BYTES, st>>>(p, l);      \
        bc_backward_hub_kernel<E, P, N><<<hgrid, BLOCK, BC_BWD_SMEM_BYTES, st>>>(p, l); \
    } while (0)
            if (edge_mode) {
                if (narrow) RXC_BWD(true, false, true);
                else RXC_BWD(true, false, false);
            } else if (endpoints) {
                if (narrow) RXC_BWD(false, true, true);
                else RXC_BWD(false, true, false);
            } else {
                if (narrow) RXC_BWD(false, false, true);
                else RXC_BWD(false, false, false);
            }
#undef RXC_BWD
            g->stats.launches += 2;
        }
        if (endpoints && !edge_mode) {
            bc_endpoint_sources_kernel<<<(bg * BG + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(p, bg);
            g->stats.launches++;
        }
        RXC_CUDA(cudaEventRecord(g->ev[2], st));
        RXC_CUDA(cudaStreamSynchronize(st));
        RXC_CUDA(cudaGetLastError());
        float f = 0.f, b = 0.f;
        RXC_CUDA(cudaEventElapsedTime(&f, g->ev[0], g->ev[1]));
        RXC_CUDA(cudaEventElapsedTime(&b, g->ev[1], g->ev[2]));
        g->stats.ms_forward += f;
        g->stats.ms_backward += b;
        g->stats.batches++;
        g->stats.sources += cnt;
        tagbase += launched_levels + 8;
        return 0;
    };

    for (uint64_t g0 = 0; g0 < total_groups; g0 += ng) {
        for (;;) {
            const bool narrow = g->opt.bc_sigma == 2 || (g->opt.bc_sigma == 0 && !g->sigma_wide);
            const int over = run_batch(g0, narrow);
            if (over == 0) break;
            if (over & 2) {  // a level queue outgrew its typical-case size: worst-case queues from here on
                if (g->queue_wide) throw std::runtime_error("level queue overflow with worst-case queues");
                g->queue_wide = true;
                const std::vector<uint32_t> rest(sources.begin() + (size_t)g0 * per, sources.end());
                bc_run(g, rest, endpoints, edge_mode, d_acc, grp);  // re-sizes the scratch
                return;
            }
            if (g->opt.bc_sigma == 2) throw std::runtime_error("path count exceeds 32 bits (bc_sigma=2 forces uint32 sigma rows)");
            g->sigma_wide = true;
        }
    }
}

static std::vector<uint32_t> all_sources(const rxc_graph *g) {
    std::vector<uint32_t> s(g->host.n_live);
    for (uint32_t i = 0; i < g->host.n_live; i++) s[i] = i;
    return s;
}

static std::vector<uint32_t> to_compact(const rxc_graph *g, const uint32_t *sources, uint64_t k) {
    std::vector<uint32_t> s(k);
    for (uint64_t i = 0; i < k; i++) {
        const uint32_t o = sources[i];
        const uint32_t c = g->host.compact_of(o);
        if (c == 0xFFFFFFFFu) throw std::invalid_argument("source " + std::to_string(o) + " is not a live node");
        s[i] = c;
    }
    return s;
}

// un-rescaled betweenness sums over `sources` (compact ids) into out (original indexing).
// reduce_root >= 0: the partial vector is first sum-reduced over the ranks of rxc_comm_init, in
// place in HBM on the handle's stream, and only the root reads it back (out untouched elsewhere).
static void bc_partial(rxc_graph *g, const std::vector<uint32_t> &sources, bool endpoints,
                       bool edge_mode, double *out, int reduce_root = -1) {
    std::lock_guard<std::mutex> lock(g->mu);
    begin_call(g);
    const uint32_t n = g->host.n_live;
    const size_t out_len = edge_mode ? g->host.e_bound : g->host.n_bound;
    const size_t acc_len = edge_mode ? g->host.e_bound : n;
    const bool dense = edge_mode || g->host.n_bound == n;  // no holes: device order == original order
    if (out && (!dense || acc_len == 0)) std::fill(out, out + out_len, 0.0);
    if (acc_len == 0) {
        end_call(g);
        return;
    }
    DevBuf<double> &acc = g->acc_buf;  // kept across calls: cudaMalloc/cudaFree would serialise the device
    if (acc.n < acc_len) acc.alloc(acc_len);
    RXC_CUDA(cudaMemsetAsync(acc.p, 0, acc_len * sizeof(double), g->stream));
    for (const auto &round : distinct_rounds(sources, n)) bc_run(g, round, endpoints, edge_mode, acc.p);
    end_call(g);
    if (reduce_root >= 0 && comm_world() > 1) {
        // rendezvous first, so that the time a rank waits for the slowest one (ms_reduce_wait) is told
        // apart from the collective itself (ms_reduce: 8 MB over NVLink at config 3)
        RXC_CUDA(cudaEventRecord(g->ev[6], g->stream));
        if (comm_rendezvous_device(g->stream) != RXC_OK) throw NcclError(g_err);
        RXC_CUDA(cudaEventRecord(g->ev[8], g->stream));
        if (comm_reduce_device(acc.p, acc_len, reduce_root, g->stream) != RXC_OK)
            throw NcclError(g_err);
        RXC_CUDA(cudaEventRecord(g->ev[7], g->stream));
        RXC_CUDA(cudaEventSynchronize(g->ev[7]));
        float ms = 0.f, wait = 0.f;
        RXC_CUDA(cudaEventElapsedTime(&wait, g->ev[6], g->ev[8]));
        RXC_CUDA(cudaEventElapsedTime(&ms, g->ev[8], g->ev[7]));
        g->stats.ms_reduce = ms;
        g->stats.ms_reduce_wait = wait;
        if (comm_rank() != reduce_root) {
            fetch_stats(g);
            return;
        }
    }
    const double t0 = now_ms();
    if (dense) {  // 134 MB for the edge scores of config 4: through the pinned staging slots
        d2h_staged(pinned_for(g->device), out, acc.p, acc_len * sizeof(double), g->stream);
    } else {
        std::vector<double> h(n);
        RXC_CUDA(cudaMemcpyAsync(h.data(), acc.p, acc_len * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
        RXC_CUDA(cudaStreamSynchronize(g->stream));
        for (uint32_t i = 0; i < n; i++) out[g->host.orig_of(i)] = h[i];
    }
    fetch_stats(g);
    g->stats.ms_d2h += now_ms() - t0;
}

// Distinct compact ids of a group given in ORIGINAL indices (the reference collects a HashSet,
// centrality.rs:1869,1982); every index must name a live node (src/centrality.rs:1216-1222).
static std::vector<uint32_t> group_members(const rxc_graph *g, const uint32_t *group, uint64_t k) {
    std::vector<uint32_t> m = to_compact(g, group, k);
    std::sort(m.begin(), m.end());
    m.erase(std::unique(m.begin(), m.end()), m.end());
    return m;
}

// sum over the listed sources (compact ids; members of the group are skipped, :1987-1989) of
// sum_t (sigma_full - paths_avoiding) / sigma_full, before the /2 and the normalisation
static double group_bc_partial(rxc_graph *g, const std::vector<uint32_t> &members,
                               const std::vector<uint32_t> &sources) {
    std::lock_guard<std::mutex> lock(g->mu);
    begin_call(g);
    const uint32_t n = g->host.n_live;
    double total = 0.0;
    if (n && !sources.empty()) {
        cudaStream_t st = g->stream;
        std::vector<uint32_t> h_blocked((n + 31) / 32, 0);
        for (uint32_t m : members) h_blocked[m >> 5] |= 1u << (m & 31);
        DevBuf<uint32_t> d_members, d_blocked;
        d_members.alloc(std::max<size_t>(members.size(), 1));
        d_blocked.alloc(h_blocked.size());
        if (!members.empty())
            RXC_CUDA(cudaMemcpyAsync(d_members.p, members.data(), members.size() * 4, cudaMemcpyHostToDevice, st));
        RXC_CUDA(cudaMemcpyAsync(d_blocked.p, h_blocked.data(), h_blocked.size() * 4, cudaMemcpyHostToDevice, st));
        std::vector<uint32_t> src;
        src.reserve(sources.size());
        for (uint32_t s : sources)
            if (!((h_blocked[s >> 5] >> (s & 31)) & 1u)) src.push_back(s);
        DevBuf<double> &acc = g->acc_buf;
        if (acc.n < 1) acc.alloc(1);
        RXC_CUDA(cudaMemsetAsync(acc.p, 0, sizeof(double), st));
        const GroupSpec spec{d_members.p, (uint32_t)members.size(), d_blocked.p};
        for (const auto &round : distinct_rounds(src, n)) bc_run(g, round, false, false, acc.p, &spec);
        end_call(g);
        RXC_CUDA(cudaMemcpyAsync(&total, acc.p, sizeof(double), cudaMemcpyDeviceToHost, st));
        RXC_CUDA(cudaStreamSynchronize(st));
        fetch_stats(g);
    } else {
        end_call(g);
    }
    return total;
}

// ------------------------------------------------------------------------------------------------
// closeness / distance matrix
// ------------------------------------------------------------------------------------------------
enum class BfsMode { Closeness, Rows, Sums };  // Sums: out[i] = {reached, sum of distances}

// sources: compact ids.  Closeness: out[i] = closeness of sources[i].  Rows: out = rows of the
// distance matrix for sources[i] (n_live columns each).
static void bfs_run(rxc_graph *g, const DeviceCSR &csr, const std::vector<uint32_t> &sources,
                    BfsMode mode, bool wf_improved, double null_value, double *out) {
    const uint32_t n = g->host.n_live;
    if (sources.empty() || n == 0) return;
    const bool rows_mode = mode == BfsMode::Rows;
    const uint64_t total_groups = (sources.size() + GW - 1) / GW;
    const uint32_t lcap = n + 2;
    const double per_group = (double)n * 20.0 + (rows_mode ? (double)n * GW * 8.0 : 0.0) + 1024.0;
    uint64_t ng = g->opt.bfs_groups > 0 ? (uint64_t)g->opt.bfs_groups : 4096;
    {
        const uint64_t want = std::min<uint64_t>(ng, total_groups);
        BfsWorkspace &w0 = g->bfs_ws;
        const bool have = w0.ng >= want && (!rows_mode || w0.ng_rows >= want);
        if (!have && w0.wanted >= want && w0.ng > 0 && (!rows_mode || w0.ng_rows > 0)) {
            ng = rows_mode ? std::min(w0.ng, w0.ng_rows) : w0.ng;  // clamped by memory earlier
        } else if (!have) {
            w0.wanted = want;
            ng = std::min<uint64_t>(ng, std::max<uint64_t>(1, (uint64_t)((double)device_budget(g->opt.mem_fraction_pct) / per_group)));
        }
    }
    if (rows_mode)  // keep one batch of rows under 8 GiB
        ng = std::min<uint64_t>(ng, std::max<uint64_t>(1, (8ull << 30) / ((uint64_t)n * GW * 8)));
    ng = std::min<uint64_t>(ng, total_groups);
    ng = std::min<uint64_t>(ng, 65535);

    BfsWorkspace &ws = g->bfs_ws;
    cudaStream_t st = g->stream;
    if (ws.ng < ng) {
        ws.release();
        g->bc_ws.release();
        ws.visited.alloc((size_t)ng * n);
        ws.next0.alloc((size_t)ng * n);
        ws.next1.alloc((size_t)ng * n);
        ws.q0.alloc((size_t)ng * n);
        ws.q1.alloc((size_t)ng * n);
        ws.cnt.alloc((size_t)3 * ng);
        ws.lvl_total.alloc(lcap);
        ws.src.alloc((size_t)ng * GW);
        ws.reach.alloc((size_t)ng * GW);
        ws.dsum.alloc((size_t)ng * GW);
        ws.vals.alloc((size_t)ng * GW * 2);
        ws.ng = ng;
        ws.dirty_levels = lcap;
    }
    if (rows_mode && ws.ng_rows < ng) {
        ws.rows.alloc((size_t)ng * GW * n);
        ws.ng_rows = ng;
    }

    BfsParams p{};
    p.row = csr.row.p;
    p.col = csr.col.p;
    p.visited = ws.visited.p;
    p.next0 = ws.next0.p;
    p.next1 = ws.next1.p;
    p.q0 = ws.q0.p;
    p.q1 = ws.q1.p;
    p.cnt = ws.cnt.p;
    p.lvl_total = ws.lvl_total.p;
    p.reach = ws.reach.p;
    p.dsum = ws.dsum.p;
    p.stats = g->d_stats.p;
    p.rows = rows_mode ? ws.rows.p : nullptr;
    p.src = ws.src.p;
    p.n_cols = n;
    p.n = n;

    std::vector<uint32_t> h_src((size_t)ng * GW), h_totals;
    const size_t vstride = mode == BfsMode::Sums ? 2 : 1;
    std::vector<double> h_vals((size_t)ng * GW * 2);
    const uint32_t vblocks = (n + BLOCK - 1) / BLOCK;
    uint32_t &clear_levels = ws.dirty_levels;

    for (uint64_t g0 = 0; g0 < total_groups; g0 += ng) {
        const uint32_t bg = (uint32_t)std::min<uint64_t>(ng, total_groups - g0);
        p.ngroups = bg;
        std::fill(h_src.begin(), h_src.end(), NO_SOURCE);
        const size_t s0 = (size_t)g0 * GW;
        const size_t cnt = std::min(sources.size() - s0, (size_t)bg * GW);
        std::copy(sources.begin() + s0, sources.begin() + s0 + cnt, h_src.begin());
        RXC_CUDA(cudaEventRecord(g->ev[0], st));
        RXC_CUDA(cudaMemcpyAsync(ws.src.p, h_src.data(), (size_t)bg * GW * 4,
                                 cudaMemcpyHostToDevice, st));
        RXC_CUDA(cudaMemsetAsync(ws.lvl_total.p, 0, (size_t)clear_levels * 4, st));
        bfs_init_kernel<<<dim3(vblocks, bg), BLOCK, 0, st>>>(p);
        bfs_seed_kernel<<<bg, 32, 0, st>>>(p);
        g->stats.launches += 2;
        if (rows_mode) {
            const uint64_t len = (uint64_t)bg * GW * n;
            const uint32_t fb = (uint32_t)std::min<uint64_t>((len + BLOCK - 1) / BLOCK, (uint64_t)g->sms * 16);
            bfs_fill_f64_kernel<<<fb, BLOCK, 0, st>>>(ws.rows.p, len, null_value);
            g->stats.launches++;
        }
        uint32_t launched_levels = 0;
        // frontier sizes are only known a chunk late: size the grid from the last total seen
        // (the kernel grid-strides, so an underestimate only costs parallelism)
        uint32_t est = GW;
        const uint32_t gx_fill = (g->sms * 4 + bg - 1) / bg;
        const uint32_t L = run_levels(
            g, ws.lvl_total.p, n,
            [&](uint32_t level) {
                if (level < h_totals.size() && h_totals[level]) est = h_totals[level] / bg + 1;
                uint32_t gx = (4 * est + BLOCK - 1) / BLOCK;
                gx = std::max<uint32_t>(1, std::min(std::max(gx, gx_fill), vblocks));
                if (rows_mode)
                    bfs_level_kernel<true><<<dim3(gx, bg), BLOCK, 0, st>>>(p, level);
                else
                    bfs_level_kernel<false><<<dim3(gx, bg), BLOCK, 0, st>>>(p, level);
                launched_levels++;
            },
            h_totals);
        g->stats.launches += launched_levels;
        g->stats.levels += L;
        for (uint32_t l = 0; l < L && l < h_totals.size(); l++) g->last_bfs_entries += h_totals[l];
        clear_levels = std::min<uint32_t>(lcap, launched_levels + 2);
        if (mode == BfsMode::Sums) {
            distance_sums_kernel<<<(bg * GW + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(ws.reach.p, ws.dsum.p,
                                                                                bg * GW, ws.vals.p);
            g->stats.launches++;
        } else if (!rows_mode) {
            closeness_finalize_kernel<<<(bg * GW + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(
                ws.reach.p, ws.dsum.p, ws.src.p, bg * GW, wf_improved ? 1 : 0,
                (unsigned long long)n, ws.vals.p);
            g->stats.launches++;
        }
        RXC_CUDA(cudaEventRecord(g->ev[1], st));
        const double t0 = now_ms();
        if (rows_mode) {
            d2h_staged(pinned_for(g->device), out + (size_t)s0 * n, ws.rows.p, (size_t)cnt * n * 8, st);
        } else {
            RXC_CUDA(cudaMemcpyAsync(h_vals.data(), ws.vals.p, (size_t)cnt * 8 * vstride,
                                     cudaMemcpyDeviceToHost, st));
        }
        RXC_CUDA(cudaStreamSynchronize(st));
        RXC_CUDA(cudaGetLastError());
        if (!rows_mode) std::copy(h_vals.begin(), h_vals.begin() + cnt * vstride, out + s0 * vstride);
        g->stats.ms_d2h += now_ms() - t0;
        float f = 0.f;
        RXC_CUDA(cudaEventElapsedTime(&f, g->ev[0], g->ev[1]));
        g->stats.ms_forward += f;
        g->stats.batches++;
        g->stats.sources += cnt;
    }
}

// ------------------------------------------------------------------------------------------------
// distance-only BFS in the betweenness layout (bfs128_level_kernel): the shallow-graph regime
// ------------------------------------------------------------------------------------------------
// push: the adjacency the search expands along; pull: its transpose.  sources: distinct compact ids.
static void bfs128_run(rxc_graph *g, const DeviceCSR &push, const DeviceCSR &pull,
                       const std::vector<uint32_t> &sources, BfsMode mode, bool wf_improved,
                       double null_value, double *out) {
    const uint32_t n = g->host.n_live;
    if (sources.empty() || n == 0) return;
    const bool rows_mode = mode == BfsMode::Rows;
    const uint64_t total_groups = (sources.size() + BG - 1) / BG;
    const uint32_t lcap = n + 2;
    uint64_t ng = bc_ws_ensure(g, total_groups);
    if (rows_mode)  // keep one batch of rows under 4 GiB
        ng = std::min<uint64_t>(ng, std::max<uint64_t>(1, (4ull << 30) / ((uint64_t)n * BG * 8)));
    BcWorkspace &ws = g->bc_ws;
    cudaStream_t st = g->stream;
    DevBuf<unsigned long long> d_dsum;
    DevBuf<double> d_vals, d_rows;
    d_dsum.alloc((size_t)ng * BG);
    d_vals.alloc((size_t)ng * BG * 2);
    if (rows_mode) d_rows.alloc((size_t)ng * BG * n);

    BcParams p{};
    p.row = pull.row.p;
    p.col = pull.col.p;
    p.srow = push.row.p;
    p.scol = push.col.p;
    p.sc = ws.sc.p;
    p.sig32 = ws.sig32.p;
    p.overflow = ws.overflow.p;
    p.P0 = ws.P0.p;
    p.P1 = ws.P1.p;
    p.visited = ws.visited.p;
    p.qv = ws.qv.p;
    p.qm = ws.qm.p;
    p.qe = ws.qe.p;
    p.loff = ws.loff.p;
    p.lvlcnt = ws.lvlcnt.p;
    p.lvl_total = ws.lvl_total.p;
    p.reach = ws.reach.p;
    p.dsum = d_dsum.p;
    p.drows = rows_mode ? d_rows.p : nullptr;
    p.n_cols = n;
    p.stats = g->d_stats.p;
    p.src = ws.src.p;
    p.hubv = ws.hubv.p;
    p.hubm = ws.hubm.p;
    p.hubcnt = ws.hubcnt.p;
    p.hubcap = HUB_CAP;
    p.maybe = ws.maybe.p;
    p.mwords = (n + 31) / 32;
    p.sparse_max = (uint32_t)std::min<int64_t>(g->opt.bc_sparse_div > 0 ? n / g->opt.bc_sparse_div : 0, 0x7FFFFFFF);
    p.hub_degree = 0x7FFFFFFF;
    p.qcap = ws.qper * n;
    p.n = n;
    p.lcap = lcap;
    uint32_t &tagbase = ws.tagbase;
    uint32_t &clear_levels = ws.dirty_levels;

    std::vector<uint32_t> h_src((size_t)ng * BG), h_totals;
    const size_t vstride = mode == BfsMode::Sums ? 2 : 1;
    std::vector<double> h_vals((size_t)ng * BG * 2);
    const uint32_t vblocks = (n + BLOCK - 1) / BLOCK;

    for (uint64_t g0 = 0; g0 < total_groups; g0 += ng) {
        const uint32_t bg = (uint32_t)std::min<uint64_t>(ng, total_groups - g0);
        std::fill(h_src.begin(), h_src.end(), NO_SOURCE);
        const size_t s0 = (size_t)g0 * BG;
        const size_t cnt = std::min(sources.size() - s0, (size_t)bg * BG);
        std::copy(sources.begin() + s0, sources.begin() + s0 + cnt, h_src.begin());
        if ((uint64_t)tagbase + lcap + 16 > 0xFFFFFFF0ull) {
            RXC_CUDA(cudaMemsetAsync(ws.P0.p, 0, ws.P0.bytes(), st));
            RXC_CUDA(cudaMemsetAsync(ws.P1.p, 0, ws.P1.bytes(), st));
            tagbase = 0;
        }
        p.tagbase = tagbase;
        p.ngroups = bg;
        RXC_CUDA(cudaEventRecord(g->ev[0], st));
        RXC_CUDA(cudaMemcpyAsync(ws.src.p, h_src.data(), (size_t)bg * BG * 4, cudaMemcpyHostToDevice, st));
        RXC_CUDA(cudaMemset2DAsync(ws.lvlcnt.p, (size_t)lcap * 4, 0, (size_t)clear_levels * 4, ws.ng, st));
        RXC_CUDA(cudaMemsetAsync(ws.lvl_total.p, 0, (size_t)clear_levels * 4, st));
        RXC_CUDA(cudaMemsetAsync(ws.overflow.p, 0, 4, st));
        if (rows_mode) {
            const uint64_t len = (uint64_t)bg * BG * n;
            const uint32_t fb = (uint32_t)std::min<uint64_t>((len + BLOCK - 1) / BLOCK, (uint64_t)g->sms * 16);
            bfs_fill_f64_kernel<<<fb, BLOCK, 0, st>>>(d_rows.p, len, null_value);
            g->stats.launches++;
        }
        bc_init_kernel<<<dim3(vblocks, bg), BLOCK, 0, st>>>(p);
        bc_seed_kernel<true><<<bg, 32, 0, st>>>(p);
        bfs128_sources_kernel<<<(bg * BG + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(p, bg, rows_mode ? 1 : 0);
        g->stats.launches += 3;
        uint32_t launched = 0;
        const uint32_t L = run_levels(
            g, ws.lvl_total.p, n,
            [&](uint32_t level) {
                if (p.sparse_max) bc_mark_kernel<<<dim3(MARK_CTAS, bg), BLOCK, 0, st>>>(p, level);
                if (rows_mode) bfs128_level_kernel<true><<<dim3(vblocks, bg), BLOCK, 0, st>>>(p, level);
                else bfs128_level_kernel<false><<<dim3(vblocks, bg), BLOCK, 0, st>>>(p, level);
                launched++;
            },
            h_totals);
        g->stats.launches += (p.sparse_max ? 2 : 1) * launched;
        g->stats.levels += L;
        clear_levels = std::min<uint32_t>(lcap, launched + 2);
        uint32_t h_overflow = 0;
        RXC_CUDA(cudaMemcpyAsync(&h_overflow, ws.overflow.p, 4, cudaMemcpyDeviceToHost, st));
        if (mode == BfsMode::Sums) {
            distance_sums_kernel<<<(bg * BG + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(ws.reach.p, d_dsum.p, bg * BG, d_vals.p);
            g->stats.launches++;
        } else if (!rows_mode) {
            closeness_finalize_kernel<<<(bg * BG + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(
                ws.reach.p, d_dsum.p, ws.src.p, bg * BG, wf_improved ? 1 : 0, (unsigned long long)n, d_vals.p);
            g->stats.launches++;
        }
        RXC_CUDA(cudaEventRecord(g->ev[1], st));
        RXC_CUDA(cudaStreamSynchronize(st));
        tagbase += launched + 8;
        if (h_overflow & 2) {  // a level queue outgrew its typical size (deep graph forced into this regime)
            if (g->queue_wide) throw std::runtime_error("level queue overflow with worst-case queues");
            g->queue_wide = true;
            const std::vector<uint32_t> rest(sources.begin() + s0, sources.end());
            g->stats.levels -= L;
            bfs128_run(g, push, pull, rest, mode, wf_improved, null_value, out + s0 * (rows_mode ? (size_t)n : vstride));
            return;
        }
        const double t0 = now_ms();
        if (rows_mode) {
            d2h_staged(pinned_for(g->device), out + s0 * n, d_rows.p, cnt * (size_t)n * 8, st);
        } else {
            RXC_CUDA(cudaMemcpyAsync(h_vals.data(), d_vals.p, cnt * 8 * vstride, cudaMemcpyDeviceToHost, st));
            RXC_CUDA(cudaStreamSynchronize(st));
            std::copy(h_vals.begin(), h_vals.begin() + cnt * vstride, out + s0 * vstride);
        }
        RXC_CUDA(cudaGetLastError());
        g->stats.ms_d2h += now_ms() - t0;
        float f = 0.f;
        RXC_CUDA(cudaEventElapsedTime(&f, g->ev[0], g->ev[1]));
        g->stats.ms_forward += f;
        g->stats.batches++;
        g->stats.sources += cnt;
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-per-source BFS with the visited bitmap in shared memory (high-diameter graphs)
// ------------------------------------------------------------------------------------------------
// Shared-memory plan of the CTA-per-source kernel: the visited bitmap goes to shared memory when it
// fits next to the frontier queues, otherwise to a per-CTA slab in global memory.
static bool smem_bfs_fits(uint32_t n, uint32_t *qs_out, size_t *bytes_out, bool *global_bitmap = nullptr) {
    const uint32_t words = (n + 31) / 32;
    // Queue entries kept in shared memory (the rest spills to global memory).  Small graphs get a
    // quarter of n (at least 1024): a frontier of a lattice is a thin ring, and at n = 6451
    // (heavy-hex d = 51) full-size queues (52 KB) allowed 4 CTAs per SM where 14 KB allow 8.
    const uint32_t qs = std::min<uint32_t>(SB_QS, std::max<uint32_t>(1024, ((n / 4 + 31) / 32) * 32));
    size_t bytes = ((size_t)words + 2 * (size_t)qs) * 4;
    bool global_bm = false;
    if (bytes > 200 * 1024) {
        global_bm = true;
        bytes = 2 * (size_t)qs * 4;
    }
    if (qs_out) *qs_out = qs;
    if (bytes_out) *bytes_out = bytes;
    if (global_bitmap) *global_bitmap = global_bm;
    return true;
}

static void bfs_smem_run(rxc_graph *g, const DeviceCSR &csr, const uint32_t *sources, size_t k,
                         BfsMode mode, bool wf_improved, double null_value, double *out,
                         uint32_t *max_depth_out = nullptr) {
    const uint32_t n = g->host.n_live;
    if (k == 0 || n == 0) return;
    const bool rows_mode = mode == BfsMode::Rows;
    uint32_t qs = 0;
    size_t smem = 0;
    bool gbm = false;
    smem_bfs_fits(n, &qs, &smem, &gbm);
    // lattices (grid, heavy-hex: at most 4 arcs per vertex) get the ELL adjacency
    DeviceCSR &mcsr = const_cast<DeviceCSR &>(csr);
    if (mcsr.max_degree < 0) {
        DevBuf<uint32_t> d_max;
        d_max.alloc(1);
        RXC_CUDA(cudaMemsetAsync(d_max.p, 0, 4, g->stream));
        csr_max_degree_kernel<<<(n + BLOCK - 1) / BLOCK, BLOCK, 0, g->stream>>>(csr.row.p, n, d_max.p);
        uint32_t h_max = 0;
        RXC_CUDA(cudaMemcpyAsync(&h_max, d_max.p, 4, cudaMemcpyDeviceToHost, g->stream));
        RXC_CUDA(cudaStreamSynchronize(g->stream));
        mcsr.max_degree = h_max;
    }
    if (mcsr.max_degree <= 4 && g->opt.bfs_ell && !mcsr.ell.p) {
        mcsr.ell.alloc(n);
        csr_to_ell4_kernel<<<(n + BLOCK - 1) / BLOCK, BLOCK, 0, g->stream>>>(csr.row.p, csr.col.p, n, mcsr.ell.p);
    }
    const bool ell = mcsr.ell.p != nullptr && g->opt.bfs_ell;
    using KernelT = void (*)(SmemBfsParams);
    const KernelT table[8] = {bfs_smem_kernel<false, false, false>, bfs_smem_kernel<false, false, true>,
                              bfs_smem_kernel<false, true, false>,  bfs_smem_kernel<false, true, true>,
                              bfs_smem_kernel<true, false, false>,  bfs_smem_kernel<true, false, true>,
                              bfs_smem_kernel<true, true, false>,   bfs_smem_kernel<true, true, true>};
    const KernelT kernel = table[(rows_mode ? 4 : 0) + (gbm ? 2 : 0) + (ell ? 1 : 0)];
    static bool attr_done_dev[64] = {false};  // function attributes are per device
    bool &attr_done = attr_done_dev[g->device & 63];
    if (!attr_done) {
        for (KernelT k4 : table)
            RXC_CUDA(cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    // threads per source: a level of a small lattice has a few dozen frontier vertices, and more
    // sources per SM beat more threads per source (heavy-hex d = 51, n = 6451: 128 threads 0.88 ms,
    // 256 1.20, 512 2.00; grid 300 x 300: 512 threads 121 ms, 256 134, 128 178, 1024 245)
    const int auto_threads = n >= (1u << 18) ? 1024 : (n >= (1u << 16) ? 512 : (n >= (1u << 14) ? 256 : 128));
    const int threads = g->opt.bfs_threads > 0 ? (int)g->opt.bfs_threads : auto_threads;
    int per_sm = 1;
    const int sms = g->sms;
    RXC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    per_sm = std::max(per_sm, 1);
    cudaStream_t st = g->stream;

    // rows mode: keep one batch of rows under 8 GiB
    size_t batch = k;
    if (rows_mode) batch = std::max<size_t>(1, std::min<size_t>(k, (8ull << 30) / ((size_t)n * 8)));
    const uint32_t grid = (uint32_t)std::min<size_t>(batch, (size_t)sms * per_sm);
    DevBuf<uint32_t> d_src, d_spill, d_counter, d_bitmap;
    DevBuf<unsigned long long> d_reach, d_dsum;
    DevBuf<double> d_vals, d_rows;
    DevBuf<uint32_t> d_depth;
    if (max_depth_out) d_depth.alloc(batch);
    d_src.alloc(batch);
    d_spill.alloc((size_t)grid * 2 * n);
    if (gbm) d_bitmap.alloc((size_t)grid * ((n + 31) / 32));
    d_counter.alloc(1);
    d_reach.alloc(batch);
    d_dsum.alloc(batch);
    const size_t vstride = mode == BfsMode::Sums ? 2 : 1;
    if (rows_mode) d_rows.alloc(batch * n);
    else d_vals.alloc(batch * 2);

    SmemBfsParams p{};
    p.row = csr.row.p;
    p.col = csr.col.p;
    p.ell = ell ? mcsr.ell.p : nullptr;
    p.sources = d_src.p;
    p.qspill = d_spill.p;
    p.gbitmap = gbm ? d_bitmap.p : nullptr;
    p.next_source = d_counter.p;
    p.reach = d_reach.p;
    p.dsum = d_dsum.p;
    p.depth = max_depth_out ? d_depth.p : nullptr;
    p.stats = g->d_stats.p;
    p.rows = rows_mode ? d_rows.p : nullptr;
    p.n_cols = n;
    p.n = n;
    p.words = (n + 31) / 32;
    p.qs = qs;
    std::vector<double> h_vals(rows_mode ? 0 : batch * 2);
    for (size_t s0 = 0; s0 < k; s0 += batch) {
        const size_t cnt = std::min(batch, k - s0);
        p.nsrc = (uint32_t)cnt;
        RXC_CUDA(cudaEventRecord(g->ev[0], st));
        RXC_CUDA(cudaMemcpyAsync(d_src.p, sources + s0, cnt * 4, cudaMemcpyHostToDevice, st));
        RXC_CUDA(cudaMemsetAsync(d_counter.p, 0, 4, st));
        if (rows_mode) {
            const uint64_t len = (uint64_t)cnt * n;
            const uint32_t fb = (uint32_t)std::min<uint64_t>((len + BLOCK - 1) / BLOCK, (uint64_t)g->sms * 16);
            bfs_fill_f64_kernel<<<fb, BLOCK, 0, st>>>(d_rows.p, len, null_value);
            kernel<<<grid, threads, smem, st>>>(p);
            g->stats.launches += 2;
        } else {
            kernel<<<grid, threads, smem, st>>>(p);
            if (mode == BfsMode::Sums)
                distance_sums_kernel<<<((uint32_t)cnt + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(
                    d_reach.p, d_dsum.p, (uint32_t)cnt, d_vals.p);
            else
                closeness_finalize_flat_kernel<<<((uint32_t)cnt + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(
                    d_reach.p, d_dsum.p, (uint32_t)cnt, wf_improved ? 1 : 0, (unsigned long long)n,
                    d_vals.p);
            g->stats.launches += 2;
        }
        RXC_CUDA(cudaEventRecord(g->ev[1], st));
        const double t0 = now_ms();
        if (rows_mode)
            d2h_staged(pinned_for(g->device), out + s0 * n, d_rows.p, cnt * (size_t)n * 8, st);
        else
            RXC_CUDA(cudaMemcpyAsync(h_vals.data(), d_vals.p, cnt * 8 * vstride, cudaMemcpyDeviceToHost, st));
        RXC_CUDA(cudaStreamSynchronize(st));
        RXC_CUDA(cudaGetLastError());
        if (!rows_mode) std::copy(h_vals.begin(), h_vals.begin() + cnt * vstride, out + s0 * vstride);
        if (max_depth_out) {
            std::vector<uint32_t> h_depth(cnt);
            RXC_CUDA(cudaMemcpy(h_depth.data(), d_depth.p, cnt * 4, cudaMemcpyDeviceToHost));
            for (uint32_t d : h_depth) *max_depth_out = std::max(*max_depth_out, d);
        }
        g->stats.ms_d2h += now_ms() - t0;
        float f = 0.f;
        RXC_CUDA(cudaEventElapsedTime(&f, g->ev[0], g->ev[1]));
        g->stats.ms_forward += f;
        g->stats.batches++;
        g->stats.sources += cnt;
    }
}

// the transpose of one of the handle's adjacencies (what a bottom-up level pulls over)
static const DeviceCSR &transpose_of(rxc_graph *g, const DeviceCSR &csr) {
    if (!g->host.directed) return csr;  // `out` holds both directions
    if (&csr == &g->out) return g->in;
    if (&csr == &g->in) return g->out;
    return csr;  // the symmetrised structure is its own transpose
}

static void bfs_dispatch_distinct(rxc_graph *g, const DeviceCSR &csr, const std::vector<uint32_t> &sources,
                                  BfsMode mode, bool wf_improved, double null_value, double *out) {
    const uint32_t n = g->host.n_live;
    const size_t stride = mode == BfsMode::Rows ? (size_t)n : (mode == BfsMode::Sums ? 2 : 1);
    int64_t regime = g->opt.bfs_mode;  // 0 auto, 1 32-source words top-down, 2 CTA-per-source, 3 128-source bottom-up
    if (regime == 0) {
        const size_t k = sources.size(), np = std::min<size_t>(8, k);
        std::vector<uint32_t> pilot(np);
        for (size_t i = 0; i < np; i++) pilot[i] = sources[i * k / np];
        std::vector<double> scratch(np * stride);
        uint32_t depth = 0;
        bfs_smem_run(g, csr, pilot.data(), np, mode, wf_improved, null_value, scratch.data(), &depth);
        // the pilot's work is redone below with everything else: keep the counters exact
        RXC_CUDA(cudaMemsetAsync(g->d_stats.p, 0, g->d_stats.bytes(), g->stream));
        g->stats.sources -= np;
        // shallow graphs: the 128-source bottom-up kernel once there are enough sources to fill groups
        regime = depth > (uint32_t)g->opt.bfs_depth_switch ? 2 : (sources.size() >= 64 ? 3 : 1);
    }
    if (regime == 2)
        bfs_smem_run(g, csr, sources.data(), sources.size(), mode, wf_improved, null_value, out);
    else if (regime == 3)
        bfs128_run(g, csr, transpose_of(g, csr), sources, mode, wf_improved, null_value, out);
    else
        bfs_run(g, csr, sources, mode, wf_improved, null_value, out);
}

// Pick the regime.  Bit-parallel words pay off when the frontiers of different sources coincide
// (small-world graphs: a handful of levels); on lattices and other long, thin graphs every word
// carries one or two sources and a level launch per BFS level is pure overhead.  A pilot of up to
// eight sources spread over the list runs CTA-per-source (cheap: one CTA each) and reports its BFS
// depth; deep graphs stay in that regime, shallow ones go bit-parallel.  A source listed twice is
// computed once (the kernels seed one bit per vertex and group).
static void bfs_dispatch(rxc_graph *g, const DeviceCSR &csr, const std::vector<uint32_t> &sources,
                         BfsMode mode, bool wf_improved, double null_value, double *out) {
    const uint32_t n = g->host.n_live;
    if (sources.empty() || n == 0) return;
    std::vector<uint32_t> uniq = sources;
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    if (uniq.size() == sources.size()) {
        bfs_dispatch_distinct(g, csr, sources, mode, wf_improved, null_value, out);
        return;
    }
    const size_t stride = mode == BfsMode::Rows ? (size_t)n : (mode == BfsMode::Sums ? 2 : 1);
    std::vector<double> tmp(uniq.size() * stride);
    bfs_dispatch_distinct(g, csr, uniq, mode, wf_improved, null_value, tmp.data());
    for (size_t i = 0; i < sources.size(); i++) {
        const size_t j = std::lower_bound(uniq.begin(), uniq.end(), sources[i]) - uniq.begin();
        std::copy(tmp.begin() + j * stride, tmp.begin() + (j + 1) * stride, out + i * stride);
    }
}

// group closeness: one multi-source BFS over incoming arcs (Reversed(&graph), centrality.rs:1875-1899);
// returns the sum of distances of the reached vertices (members contribute 0)
static uint64_t group_closeness_dsum(rxc_graph *g, const std::vector<uint32_t> &members) {
    std::lock_guard<std::mutex> lock(g->mu);
    begin_call(g);
    const uint32_t n = g->host.n_live;
    uint64_t dsum = 0;
    if (n && !members.empty()) {
        cudaStream_t st = g->stream;
        const uint32_t lcap = n + 2;
        DevBuf<uint32_t> words, cnt, lvl_total, src, d_members;
        DevBuf<unsigned long long> sums;
        words.alloc((size_t)5 * n);  // visited, next0, next1, q0, q1
        cnt.alloc(3);
        lvl_total.alloc(lcap);
        src.alloc(GW);
        sums.alloc(2 * GW);
        d_members.alloc(members.size());
        std::vector<uint32_t> h_src(GW, NO_SOURCE);
        h_src[0] = members[0];
        RXC_CUDA(cudaMemcpyAsync(src.p, h_src.data(), GW * 4, cudaMemcpyHostToDevice, st));
        RXC_CUDA(cudaMemcpyAsync(d_members.p, members.data(), members.size() * 4, cudaMemcpyHostToDevice, st));
        RXC_CUDA(cudaMemsetAsync(lvl_total.p, 0, lvl_total.bytes(), st));
        BfsParams p{};
        p.row = g->pred().row.p;
        p.col = g->pred().col.p;
        p.visited = words.p;
        p.next0 = words.p + (size_t)n;
        p.next1 = words.p + (size_t)2 * n;
        p.q0 = words.p + (size_t)3 * n;
        p.q1 = words.p + (size_t)4 * n;
        p.cnt = cnt.p;
        p.lvl_total = lvl_total.p;
        p.reach = sums.p;
        p.dsum = sums.p + GW;
        p.stats = g->d_stats.p;
        p.src = src.p;
        p.n = n;
        p.n_cols = n;
        p.ngroups = 1;
        const uint32_t vblocks = (n + BLOCK - 1) / BLOCK;
        bfs_init_kernel<<<dim3(vblocks, 1), BLOCK, 0, st>>>(p);
        const uint32_t nmem = (uint32_t)members.size();
        bfs_seed_multi_kernel<<<(std::max<uint32_t>(nmem, GW) + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(p, d_members.p, nmem);
        std::vector<uint32_t> h_totals;
        uint32_t launched = 0;
        const uint32_t L = run_levels(
            g, lvl_total.p, n,
            [&](uint32_t level) {
                uint32_t est = level < h_totals.size() && h_totals[level] ? h_totals[level] : nmem;
                const uint32_t gx = std::max<uint32_t>(1, std::min((4 * est + BLOCK - 1) / BLOCK, vblocks));
                bfs_level_kernel<false><<<dim3(gx, 1), BLOCK, 0, st>>>(p, level);
                launched++;
            },
            h_totals);
        g->stats.launches += 2 + launched;
        g->stats.levels += L;
        g->stats.sources += 1;
        g->stats.batches += 1;
        unsigned long long h = 0;
        RXC_CUDA(cudaMemcpyAsync(&h, p.dsum, 8, cudaMemcpyDeviceToHost, st));
        end_call(g);
        RXC_CUDA(cudaStreamSynchronize(st));
        fetch_stats(g);
        dsum = h;
    } else {
        end_call(g);
    }
    return dsum;
}


// ------------------------------------------------------------------------------------------------
// weighted all-sources shortest path lengths (wsp_kernels.cuh)
// ------------------------------------------------------------------------------------------------
enum class WspMode { Closeness, Rows };

struct WspCosts {
    DevBuf<double> pull;  // cost of every arc of the pull adjacency
    bool has_inf = false; // some cost is +inf: "reached" and "finite distance" may differ
};

// edge_weight: e_bound doubles by ORIGINAL edge index (dead slots ignored).  reversed: the search
// follows incoming edges (Reversed(&graph), centrality.rs:1341), so a vertex pulls over its OUT-arcs.
static void wsp_costs(rxc_graph *g, const double *edge_weight, bool invert, bool reversed, WspCosts &c) {
    const DeviceCSR &pull = reversed ? g->succ() : g->pred();
    if (!pull.eid.p && pull.arcs) throw std::invalid_argument("snapshot has no edge ids");
    const size_t eb = g->host.e_bound;
    DevBuf<double> d_w;
    d_w.alloc(std::max<size_t>(eb, 1));
    if (eb) RXC_CUDA(cudaMemcpyAsync(d_w.p, edge_weight, eb * 8, cudaMemcpyHostToDevice, g->stream));
    c.pull.alloc(std::max<uint64_t>(pull.arcs, 1));
    if (pull.arcs) {
        const uint32_t blocks = (uint32_t)std::min<uint64_t>((pull.arcs + BLOCK - 1) / BLOCK, (uint64_t)g->sms * 16);
        wsp_arc_costs_kernel<<<blocks, BLOCK, 0, g->stream>>>(pull.eid.p, d_w.p, pull.arcs, invert ? 1 : 0, c.pull.p);
    }
    RXC_CUDA(cudaStreamSynchronize(g->stream));
}

// sources: compact ids, pairwise distinct.  Closeness: out[i] = closeness of sources[i] and, when
// pairs != nullptr, pairs[2i], pairs[2i+1] = (vertices at finite distance, sum of distances).
// Rows: out[i * n + v] = length of the shortest path sources[i] -> v, +inf where there is none.
static void wsp_run(rxc_graph *g, bool reversed, const double *d_cost, const std::vector<uint32_t> &sources,
                    WspMode mode, bool wf_improved, double *out, double *pairs) {
    const uint32_t n = g->host.n_live;
    if (sources.empty() || n == 0) return;
    static bool attr_done[64] = {false};
    if (!attr_done[g->device & 63]) {
        RXC_CUDA(cudaFuncSetAttribute(wsp_relax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WSP_SMEM_BYTES));
        attr_done[g->device & 63] = true;
    }
    const uint64_t total_groups = (sources.size() + BG - 1) / BG;
    const uint32_t lcap = n + 2;
    uint64_t ng = bc_ws_ensure(g, total_groups);
    if (mode == WspMode::Rows)  // keep one batch of output rows under 4 GiB
        ng = std::min<uint64_t>(ng, std::max<uint64_t>(1, (4ull << 30) / ((uint64_t)n * BG * 8)));
    BcWorkspace &ws = g->bc_ws;
    cudaStream_t st = g->stream;
    const DeviceCSR &pull = reversed ? g->succ() : g->pred();
    const DeviceCSR &push = reversed ? g->pred() : g->succ();

    WspParams p{};
    p.row = pull.row.p;
    p.col = pull.col.p;
    p.cost = d_cost;
    p.srow = push.row.p;
    p.scol = push.col.p;
    p.dist = ws.sc.p;
    p.C0 = ws.P0.p;
    p.C1 = ws.P1.p;
    p.maybe = ws.maybe.p;
    p.qv = ws.qv.p;
    p.lvlcnt = ws.lvlcnt.p;
    p.lvl_total = ws.lvl_total.p;
    p.stats = g->d_stats.p;
    p.src = ws.src.p;
    p.qstride = ws.qper * n;  // >= 2 n: one queue of at most n vertices per round parity
    p.n = n;
    p.lcap = lcap;
    p.mwords = (n + 31) / 32;

    const uint32_t vblocks = (n + BLOCK - 1) / BLOCK;
    const uint32_t ntiles = (n + WSP_TILE - 1) / WSP_TILE;
    DevBuf<double> part_sum, d_vals, d_rows;
    DevBuf<unsigned long long> part_cnt;
    if (mode == WspMode::Closeness) {
        part_sum.alloc((size_t)ng * ntiles * BG);
        part_cnt.alloc((size_t)ng * ntiles * BG);
        d_vals.alloc((size_t)ng * BG * 3);
    } else {
        d_rows.alloc((size_t)ng * BG * n);
    }
    std::vector<uint32_t> h_src((size_t)ng * BG), h_totals;
    std::vector<double> h_vals((size_t)ng * BG * 3);
    uint32_t &clear_levels = ws.dirty_levels;
    uint32_t &tagbase = ws.tagbase;

    for (uint64_t g0 = 0; g0 < total_groups; g0 += ng) {
        const uint32_t bg = (uint32_t)std::min<uint64_t>(ng, total_groups - g0);
        std::fill(h_src.begin(), h_src.end(), NO_SOURCE);
        const size_t s0 = (size_t)g0 * BG;
        const size_t cnt = std::min(sources.size() - s0, (size_t)bg * BG);
        std::copy(sources.begin() + s0, sources.begin() + s0 + cnt, h_src.begin());
        if ((uint64_t)tagbase + lcap + 16 > 0xFFFFFFF0ull) {
            RXC_CUDA(cudaMemsetAsync(ws.P0.p, 0, ws.P0.bytes(), st));
            RXC_CUDA(cudaMemsetAsync(ws.P1.p, 0, ws.P1.bytes(), st));
            tagbase = 0;
        }
        p.tagbase = tagbase;
        p.ngroups = bg;
        RXC_CUDA(cudaEventRecord(g->ev[0], st));
        RXC_CUDA(cudaMemcpyAsync(ws.src.p, h_src.data(), (size_t)bg * BG * 4, cudaMemcpyHostToDevice, st));
        RXC_CUDA(cudaMemset2DAsync(ws.lvlcnt.p, (size_t)lcap * 4, 0, (size_t)clear_levels * 4, ws.ng, st));
        RXC_CUDA(cudaMemsetAsync(ws.lvl_total.p, 0, (size_t)clear_levels * 4, st));
        wsp_fill_kernel<<<g->sms * 8, BLOCK, 0, st>>>(ws.sc.p, (uint64_t)bg * n * BG);
        wsp_seed_kernel<<<bg, 32, 0, st>>>(p);
        g->stats.launches += 2;
        uint32_t launched = 0;
        const uint32_t R = run_levels(
            g, ws.lvl_total.p, n + 1,
            [&](uint32_t round) {
                wsp_mark_kernel<<<dim3(MARK_CTAS * 4, bg), BLOCK, 0, st>>>(p, round);
                wsp_relax_kernel<<<dim3(vblocks, bg), BLOCK, WSP_SMEM_BYTES, st>>>(p, round);
                launched++;
            },
            h_totals);
        g->stats.launches += 2 * launched;
        g->stats.levels += R;
        clear_levels = std::min<uint32_t>(lcap, launched + 2);
        RXC_CUDA(cudaEventRecord(g->ev[1], st));
        if (mode == WspMode::Closeness) {
            wsp_sums_tile_kernel<<<dim3(ntiles, bg), BG, 0, st>>>(p, part_sum.p, part_cnt.p);
            wsp_sums_finish_kernel<<<bg, BG, 0, st>>>(p, part_sum.p, part_cnt.p, ntiles, wf_improved ? 1 : 0,
                                                       (unsigned long long)n, d_vals.p);
            g->stats.launches += 2;
            RXC_CUDA(cudaEventRecord(g->ev[2], st));
            const double t0 = now_ms();
            RXC_CUDA(cudaMemcpyAsync(h_vals.data(), d_vals.p, (size_t)bg * BG * 3 * 8, cudaMemcpyDeviceToHost, st));
            RXC_CUDA(cudaStreamSynchronize(st));
            RXC_CUDA(cudaGetLastError());
            std::copy(h_vals.begin(), h_vals.begin() + cnt, out + s0);
            if (pairs) std::copy(h_vals.begin() + (size_t)bg * BG, h_vals.begin() + (size_t)bg * BG + 2 * cnt, pairs + 2 * s0);
            g->stats.ms_d2h += now_ms() - t0;
        } else {
            wsp_rows_kernel<<<dim3((n + 31) / 32, BG / 32, bg), BLOCK, 0, st>>>(p, d_rows.p, n);
            g->stats.launches++;
            // te / v counters come from the tile pass; run it for the counters only when cheap
            RXC_CUDA(cudaEventRecord(g->ev[2], st));
            const double t0 = now_ms();
            d2h_staged(pinned_for(g->device), out + s0 * n, d_rows.p, cnt * (size_t)n * 8, st);
            RXC_CUDA(cudaGetLastError());
            g->stats.ms_d2h += now_ms() - t0;
        }
        float f = 0.f, b = 0.f;
        RXC_CUDA(cudaEventElapsedTime(&f, g->ev[0], g->ev[1]));
        RXC_CUDA(cudaEventElapsedTime(&b, g->ev[1], g->ev[2]));
        g->stats.ms_forward += f;
        g->stats.ms_backward += b;
        g->stats.batches++;
        g->stats.sources += cnt;
        tagbase += launched + 8;
    }
}

static bool has_inf_cost(const rxc_graph *g, const double *edge_weight, const uint8_t *alive, bool invert) {
    for (size_t e = 0; e < g->host.e_bound; e++) {
        if (alive && !alive[e]) continue;
        const double w = edge_weight[e];
        if (invert ? (w == 0.0) : std::isinf(w)) return true;
    }
    return false;
}

// closeness of the listed sources (compact ids, distinct) from weighted distances
static void wsp_closeness(rxc_graph *g, const double *edge_weight, bool invert, bool reversed,
                          const std::vector<uint32_t> &sources, bool wf_improved, double *out) {
    std::lock_guard<std::mutex> lock(g->mu);
    begin_call(g);
    if (!sources.empty() && g->host.n_live) {
        WspCosts c;
        wsp_costs(g, edge_weight, invert, reversed, c);
        std::vector<double> pairs(sources.size() * 2);
        wsp_run(g, reversed, c.pull.p, sources, WspMode::Closeness, wf_improved, out, pairs.data());
        if (has_inf_cost(g, edge_weight, nullptr, invert)) {
            // A vertex reached only through +inf-cost edges is "reachable" with an infinite distance
            // (dijkstra.rs:139-141 inserts the first score unconditionally): the sum is +inf and the
            // closeness 0.  Count the reachable vertices with the unweighted kernel.
            std::vector<double> rs(sources.size() * 2);
            bfs_dispatch(g, reversed ? g->pred() : g->succ(), sources, BfsMode::Sums, false, 0.0, rs.data());
            for (size_t i = 0; i < sources.size(); i++)
                if (rs[2 * i] > pairs[2 * i]) out[i] = 0.0;
        }
    }
    end_call(g);
    fetch_stats(g);
}

static const DeviceCSR &symmetric_csr(rxc_graph *g) {
    if (!g->host.directed) return g->out;
    if (!g->has_sym) {
        if (g->d_esrc.p) {  // device build: both directions of every edge, no edge ids needed
            const uint32_t n = g->host.n_live;
            const uint64_t m = g->host.n_edges;
            DevBuf<uint32_t> deg, cursor, tiles, small;
            deg.alloc((size_t)n + 1);
            cursor.alloc((size_t)n + 1);
            small.alloc(4);
            // degree = out-degree + in-degree, read off the two row arrays built at snapshot time
            csr_degree_sum_kernel<<<(n + BLOCK) / BLOCK, BLOCK, 0, g->stream>>>(g->out.row.p, g->in.row.p, n, deg.p);
            device_build_csr(g->sym, n, g->d_esrc.p, g->d_edst.p, nullptr, m, 2, deg.p, cursor, tiles,
                             small.p + 1, g->stream, g->sms);
            g->sym.arcs = 2 * (uint64_t)g->n_arcs;
        } else {
            HostCSR h = make_symmetric(g->host);
            g->sym.upload(h, g->stream);
        }
        RXC_CUDA(cudaStreamSynchronize(g->stream));
        g->has_sym = true;
    }
    return g->sym;
}

// ------------------------------------------------------------------------------------------------
// in-process multi-GPU: the handle carries one replica of the snapshot per extra device
// (rxc_set_devices).  Full calls shard the live sources round-robin over the replicas, one host
// thread per device; partial vectors are summed on the host (8 MB at config 3 -- the one-process-
// per-GPU mode uses ncclReduce instead, comm.cpp).  Disjoint outputs need no reduction.
// ------------------------------------------------------------------------------------------------
static std::vector<rxc_graph *> shards_of(rxc_graph *g) {
    std::vector<rxc_graph *> v{g};
    for (auto &r : g->replicas) v.push_back(r.get());
    return v;
}

template <class F>
static void run_on_shards(rxc_graph *g, F &&fn) {
    const std::vector<rxc_graph *> sh = shards_of(g);
    if (sh.size() == 1) {
        fn(sh[0], 0, 1);
        return;
    }
    std::vector<std::exception_ptr> errs(sh.size());
    std::vector<std::thread> threads;
    for (size_t k = 0; k < sh.size(); k++)
        threads.emplace_back([&, k] {
            try {
                RXC_CUDA(cudaSetDevice(sh[k]->device));
                fn(sh[k], k, sh.size());
            } catch (...) {
                errs[k] = std::current_exception();
            }
        });
    for (auto &t : threads) t.join();
    cudaSetDevice(g->device);
    // counters of the call: work adds up, device time is the slowest shard
    for (size_t k = 1; k < sh.size(); k++) {
        const rxc_stats &r = sh[k]->stats;
        rxc_stats &a = g->stats;
        a.te += r.te; a.te_dag += r.te_dag; a.v += r.v; a.sources += r.sources;
        a.levels += r.levels; a.launches += r.launches; a.batches += r.batches;
        a.ms_total = std::max(a.ms_total, r.ms_total);
        a.ms_forward = std::max(a.ms_forward, r.ms_forward);
        a.ms_backward = std::max(a.ms_backward, r.ms_backward);
        a.ms_d2h = std::max(a.ms_d2h, r.ms_d2h);
    }
    for (auto &e : errs)
        if (e) std::rethrow_exception(e);
}

static std::vector<uint32_t> shard_sources(uint32_t n_live, size_t k, size_t K) {
    std::vector<uint32_t> s;
    s.reserve(n_live / K + 1);
    for (uint32_t i = (uint32_t)k; i < n_live; i += (uint32_t)K) s.push_back(i);
    return s;
}

// all live sources, un-rescaled, into out (original indexing)
static void bc_full(rxc_graph *g, bool endpoints, bool edge_mode, double *out) {
    const size_t out_len = edge_mode ? g->host.e_bound : g->host.n_bound;
    const size_t K = 1 + g->replicas.size();
    if (K == 1) {
        bc_partial(g, all_sources(g), endpoints, edge_mode, out);
        return;
    }
    std::vector<std::vector<double>> parts(K);
    run_on_shards(g, [&](rxc_graph *r, size_t k, size_t nk) {
        parts[k].assign(std::max<size_t>(out_len, 1), 0.0);
        bc_partial(r, shard_sources(r->host.n_live, k, nk), endpoints, edge_mode, parts[k].data());
    });
    for (size_t i = 0; i < out_len; i++) {
        double sum = 0.0;
        for (size_t k = 0; k < K; k++) sum += parts[k][i];  // fixed order: deterministic
        out[i] = sum;
    }
}

}  // namespace rxc

// ================================================================================================
// C-ABI
// ================================================================================================
extern "C" {

const char *rxc_version(void) { return "rxcuda 0.1.0 (sm_100a)"; }
const char *rxc_last_error(void) { return g_err.c_str(); }

int rxc_device_count(void) {
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return cnt;
}

int rxc_set_device(int device) {
    return guarded([&] {
        require_device();
        RXC_CUDA(cudaSetDevice(device));
        g_device = device;
    });
}

int rxc_get_device(void) { return g_device; }

int rxc_set_devices(const int *devices, int n) {
    return guarded([&] {
        require_device();
        int cnt = 0;
        RXC_CUDA(cudaGetDeviceCount(&cnt));
        std::vector<int> v;
        for (int i = 0; i < n; i++) {
            if (!devices || devices[i] < 0 || devices[i] >= cnt)
                throw std::invalid_argument("device id out of range");
            if (std::find(v.begin(), v.end(), devices[i]) != v.end())
                throw std::invalid_argument("duplicate device id");
            v.push_back(devices[i]);
        }
        g_devices = v;
        if (!v.empty()) {
            RXC_CUDA(cudaSetDevice(v[0]));
            g_device = v[0];
        }
    });
}

// one snapshot on one device
static rxc_graph *snapshot_on(int device, uint32_t n_bound, const uint32_t *live_nodes, uint32_t n_live,
                              uint32_t e_bound, const uint32_t *edge_src, const uint32_t *edge_dst,
                              const uint32_t *edge_id, uint64_t n_edges, bool directed) {
    const double t0 = now_ms();
    std::unique_ptr<rxc_graph> g(new rxc_graph());
    g->device = device;
    g->opt = options_now();
    RXC_CUDA(cudaSetDevice(device));
    RXC_CUDA(cudaDeviceGetAttribute(&g->sms, cudaDevAttrMultiProcessorCount, device));
    pool_setup(device);
    RXC_CUDA(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
    for (auto &e : g->ev) RXC_CUDA(cudaEventCreate(&e));
    if (g->opt.device_csr) {
        device_snapshot(g.get(), n_bound, live_nodes, n_live, e_bound, edge_src, edge_dst, edge_id,
                        n_edges, directed);
    } else {
        g->host = make_snapshot(n_bound, live_nodes, n_live, e_bound, edge_src, edge_dst, edge_id,
                                n_edges, directed);
        g->out.upload(g->host.out, g->stream);
        if (g->host.directed) g->in.upload(g->host.in, g->stream);
        g->n_arcs = g->host.out.col.size();
    }
    g->d_stats.alloc(ST_STRIPES * ST_COUNT);
    RXC_CUDA(cudaMemsetAsync(g->d_stats.p, 0, g->d_stats.bytes(), g->stream));
    RXC_CUDA(cudaStreamSynchronize(g->stream));
    memset(&g->stats, 0, sizeof(g->stats));
    g->ms_snapshot = now_ms() - t0;
    g->stats.ms_snapshot = g->ms_snapshot;
    return g.release();
}

int rxc_graph_snapshot(uint32_t n_bound, const uint32_t *live_nodes, uint32_t n_live,
                       uint32_t e_bound, const uint32_t *edge_src, const uint32_t *edge_dst,
                       const uint32_t *edge_id, uint64_t n_edges, int directed,
                       rxc_graph **out_graph) {
    if (!out_graph) return RXC_ERR_INVALID;
    *out_graph = nullptr;
    rxc_graph *g = nullptr;
    int rc = guarded([&] {
        if ((n_live && !live_nodes) || (n_edges && (!edge_src || !edge_dst)))
            throw std::invalid_argument("null array");
        require_device();
        const double t0 = now_ms();
        // primary on the first device; with rxc_set_devices the CSR is replicated on the others
        const std::vector<int> devs = g_devices.empty() ? std::vector<int>{g_device} : g_devices;
        g = snapshot_on(devs[0], n_bound, live_nodes, n_live, e_bound, edge_src, edge_dst, edge_id,
                        n_edges, directed != 0);
        for (size_t d = 1; d < devs.size(); d++)
            g->replicas.emplace_back(snapshot_on(devs[d], n_bound, live_nodes, n_live, e_bound, edge_src,
                                                 edge_dst, edge_id, n_edges, directed != 0));
        RXC_CUDA(cudaSetDevice(g->device));
        g->ms_snapshot = now_ms() - t0;
        g->stats.ms_snapshot = g->ms_snapshot;
    });
    if (rc != RXC_OK) {
        delete g;
        return rc;
    }
    *out_graph = g;
    return RXC_OK;
}

// ---- content-keyed snapshot cache --------------------------------------------------
// The reference re-reads the graph on every call; a binding has no mutation counter to key on, so
// the key is a hash of what the caller hands over.  64-bit multiply-xorshift lanes over 8-byte words,
// two seeds (128 bits), chunks hashed by up to 8 host threads and folded in chunk order.
static uint64_t mix64(uint64_t x) {
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return x;
}

static void hash_chunk(const uint8_t *p, size_t len, uint64_t seed, uint64_t out[2]) {
    uint64_t a = seed ^ 0x9E3779B97F4A7C15ull, b = ~seed * 0xC2B2AE3D27D4EB4Full, c = seed + len, d = seed ^ len;
    size_t i = 0;
    for (; i + 32 <= len; i += 32) {  // four independent lanes: the multiplies pipeline
        uint64_t w[4];
        memcpy(w, p + i, 32);
        a = (a ^ w[0]) * 0xFF51AFD7ED558CCDull; a ^= a >> 29;
        b = (b ^ w[1]) * 0xC4CEB9FE1A85EC53ull; b ^= b >> 31;
        c = (c ^ w[2]) * 0x9FB21C651E98DF25ull; c ^= c >> 30;
        d = (d ^ w[3]) * 0xD6E8FEB86659FD93ull; d ^= d >> 28;
    }
    uint64_t tail[4] = {0, 0, 0, 0};
    memcpy(tail, p + i, len - i);
    a = mix64(a ^ tail[0]);
    b = mix64(b ^ tail[1]);
    c = mix64(c ^ tail[2]);
    d = mix64(d ^ tail[3]);
    out[0] = mix64(a + 3 * c) ^ mix64(b ^ (d << 1));
    out[1] = mix64(b + 5 * d) ^ mix64(a ^ (c >> 1));
}

static void hash_fold(uint64_t key[2], const uint64_t h[2]) {
    key[0] = mix64(key[0] ^ h[0]) + 0x9E3779B97F4A7C15ull;
    key[1] = mix64(key[1] + h[1]) ^ 0xC2B2AE3D27D4EB4Full;
}

static void hash_array(uint64_t key[2], const void *data, size_t bytes, uint64_t tagv) {
    constexpr size_t CHUNK = 4u << 20;
    const size_t nchunks = (bytes + CHUNK - 1) / CHUNK;
    std::vector<uint64_t> hs(2 * std::max<size_t>(nchunks, 1), 0);
    auto work = [&](size_t c0, size_t c1) {
        for (size_t c = c0; c < c1; c++)
            hash_chunk((const uint8_t *)data + c * CHUNK, std::min(CHUNK, bytes - c * CHUNK), tagv + c, &hs[2 * c]);
    };
    const size_t T = std::min<size_t>(8, nchunks);
    if (T <= 1) {
        work(0, nchunks);
    } else {
        std::vector<std::thread> th;
        for (size_t t = 0; t < T; t++) th.emplace_back(work, nchunks * t / T, nchunks * (t + 1) / T);
        for (auto &x : th) x.join();
    }
    const uint64_t head[2] = {tagv, (uint64_t)bytes};
    hash_fold(key, head);
    for (size_t c = 0; c < nchunks; c++) hash_fold(key, &hs[2 * c]);
}

static std::mutex g_cache_mu;             // guards g_cache and every handle's `refs`
static std::vector<rxc_graph *> g_cache;  // most recently used first
static size_t g_cache_cap = 2;            // a handle keeps its multi-GB scratch: few entries

static void graph_unref_locked(rxc_graph *g) {
    if (g && --g->refs == 0) delete g;
}

int rxc_graph_snapshot_cached(uint32_t n_bound, const uint32_t *live_nodes, uint32_t n_live,
                              uint32_t e_bound, const uint32_t *edge_src, const uint32_t *edge_dst,
                              const uint32_t *edge_id, uint64_t n_edges, int directed,
                              rxc_graph **out_graph, int *hit) {
    if (!out_graph) return RXC_ERR_INVALID;
    *out_graph = nullptr;
    if (hit) *hit = 0;
    if ((n_live && !live_nodes) || (n_edges && (!edge_src || !edge_dst))) {
        set_error("invalid argument: null array");
        return RXC_ERR_INVALID;
    }
    uint64_t key[2] = {0x243F6A8885A308D3ull, 0x13198A2E03707344ull};
    const uint64_t head[2] = {((uint64_t)n_bound << 32) | e_bound, (n_edges << 1) | (directed ? 1u : 0u)};
    hash_fold(key, head);
    uint64_t devs[2] = {(uint64_t)g_device, (uint64_t)g_devices.size()};
    for (int d : g_devices) devs[1] = mix64(devs[1] ^ (uint64_t)(d + 1));
    hash_fold(key, devs);
    hash_array(key, live_nodes, (size_t)n_live * 4, 1ull << 40);
    hash_array(key, edge_src, (size_t)n_edges * 4, 2ull << 40);
    hash_array(key, edge_dst, (size_t)n_edges * 4, 3ull << 40);
    if (edge_id) hash_array(key, edge_id, (size_t)n_edges * 4, 4ull << 40);
    if (key[0] == 0 && key[1] == 0) key[1] = 1;  // (0,0) means "not cached"
    {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        for (size_t i = 0; i < g_cache.size(); i++) {
            rxc_graph *c = g_cache[i];
            if (c->cache_key[0] == key[0] && c->cache_key[1] == key[1] && c->host.n_bound == n_bound &&
                c->host.n_live == n_live && c->host.e_bound == e_bound && c->host.n_edges == n_edges &&
                c->host.directed == (directed != 0)) {
                g_cache.erase(g_cache.begin() + i);
                g_cache.insert(g_cache.begin(), c);
                c->refs++;
                *out_graph = c;
                if (hit) *hit = 1;
                return RXC_OK;
            }
        }
    }
    rxc_graph *g = nullptr;
    const int rc = rxc_graph_snapshot(n_bound, live_nodes, n_live, e_bound, edge_src, edge_dst, edge_id, n_edges,
                                      directed, &g);
    if (rc != RXC_OK) return rc;
    std::lock_guard<std::mutex> lock(g_cache_mu);
    g->cache_key[0] = key[0];
    g->cache_key[1] = key[1];
    if (g_cache_cap > 0) {
        g->refs++;  // the cache's own reference
        g_cache.insert(g_cache.begin(), g);
        while (g_cache.size() > g_cache_cap) {
            rxc_graph *old = g_cache.back();
            g_cache.pop_back();
            graph_unref_locked(old);
        }
    }
    *out_graph = g;
    return RXC_OK;
}

int rxc_snapshot_cache_clear(void) {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    const int n = (int)g_cache.size();
    for (rxc_graph *c : g_cache) graph_unref_locked(c);
    g_cache.clear();
    slab_free_parked();
    return n;
}

void rxc_graph_free(rxc_graph *g) {
    if (!g) return;
    std::lock_guard<std::mutex> lock(g_cache_mu);
    graph_unref_locked(g);
}

int rxc_host_alloc(size_t bytes, void **out) {
    if (!out) return RXC_ERR_INVALID;
    *out = nullptr;
    return guarded([&] {
        require_device();
        void *p = nullptr;
        const cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            throw std::bad_alloc();
        }
        RXC_CUDA(e);
        *out = p;
    });
}

void rxc_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int rxc_graph_info(const rxc_graph *g, uint32_t *n_bound, uint32_t *n_live, uint32_t *e_bound,
                   uint64_t *n_arcs, int *directed) {
    if (!g) return RXC_ERR_INVALID;
    if (n_bound) *n_bound = g->host.n_bound;
    if (n_live) *n_live = g->host.n_live;
    if (e_bound) *e_bound = g->host.e_bound;
    if (n_arcs) *n_arcs = g->n_arcs;
    if (directed) *directed = g->host.directed ? 1 : 0;
    return RXC_OK;
}

void rxc_rescale(double *x, uint64_t len, uint64_t node_count, int normalized, int directed,
                 int include_endpoints) {
    // _rescale, rustworkx-core/src/centrality.rs:221-252: one multiply by a precomputed reciprocal
    bool do_scale = true;
    double scale = 1.0;
    if (normalized) {
        if (include_endpoints) {
            if (node_count < 2) do_scale = false;
            else scale = 1.0 / (double)(node_count * (node_count - 1));
        } else if (node_count <= 2) {
            do_scale = false;
        } else {
            scale = 1.0 / (double)((node_count - 1) * (node_count - 2));
        }
    } else if (!directed) {
        scale = 0.5;
    } else {
        do_scale = false;
    }
    if (do_scale)
        for (uint64_t i = 0; i < len; i++) x[i] = x[i] * scale;
}

int rxc_betweenness_partial(rxc_graph *g, const uint32_t *sources, uint64_t n_sources,
                            int include_endpoints, double *out) {
    if (!g || !out || (n_sources && !sources)) return RXC_ERR_INVALID;
    return guarded([&] { bc_partial(g, to_compact(g, sources, n_sources), include_endpoints != 0, false, out); });
}

int rxc_edge_betweenness_partial(rxc_graph *g, const uint32_t *sources, uint64_t n_sources,
                                 double *out) {
    if (!g || !out || (n_sources && !sources)) return RXC_ERR_INVALID;
    return guarded([&] { bc_partial(g, to_compact(g, sources, n_sources), false, true, out); });
}

int rxc_betweenness_source_order(rxc_graph *g, const uint32_t *sources, uint64_t n_sources,
                                 uint32_t *ordered) {
    if (!g || (n_sources && (!sources || !ordered))) return RXC_ERR_INVALID;
    return guarded([&] {
        std::lock_guard<std::mutex> lock(g->mu);
        begin_call(g);
        const std::vector<uint32_t> o = bc_order_sources(g, to_compact(g, sources, n_sources));
        end_call(g);
        for (uint64_t i = 0; i < n_sources; i++) ordered[i] = g->host.orig_of(o[i]);
    });
}

int rxc_betweenness_partial_reduce(rxc_graph *g, const uint32_t *sources, uint64_t n_sources,
                                   int include_endpoints, int root, double *out) {
    if (!g || (n_sources && !sources) || root < 0 || root >= comm_world()) return RXC_ERR_INVALID;
    if (comm_rank() == root && !out) return RXC_ERR_INVALID;
    return guarded([&] {
        bc_partial(g, to_compact(g, sources, n_sources), include_endpoints != 0, false, out, root);
    });
}

int rxc_edge_betweenness_partial_reduce(rxc_graph *g, const uint32_t *sources, uint64_t n_sources,
                                        int root, double *out) {
    if (!g || (n_sources && !sources) || root < 0 || root >= comm_world()) return RXC_ERR_INVALID;
    if (comm_rank() == root && !out) return RXC_ERR_INVALID;
    return guarded([&] { bc_partial(g, to_compact(g, sources, n_sources), false, true, out, root); });
}

int rxc_betweenness(rxc_graph *g, int include_endpoints, int normalized, double *out) {
    if (!g || !out) return RXC_ERR_INVALID;
    return guarded([&] {
        bc_full(g, include_endpoints != 0, false, out);
        rxc_rescale(out, g->host.n_bound, g->host.n_live, normalized, g->host.directed ? 1 : 0,
                    include_endpoints);
    });
}

int rxc_edge_betweenness(rxc_graph *g, int normalized, double *out) {
    if (!g || !out) return RXC_ERR_INVALID;
    return guarded([&] {
        bc_full(g, false, true, out);
        rxc_rescale(out, g->host.e_bound, g->host.n_live, normalized, g->host.directed ? 1 : 0, 1);
    });
}

int rxc_group_betweenness_partial(rxc_graph *g, const uint32_t *group, uint64_t n_group,
                                  const uint32_t *sources, uint64_t n_sources, double *out_sum) {
    if (!g || !out_sum || (n_group && !group) || (n_sources && !sources)) return RXC_ERR_INVALID;
    return guarded([&] {
        *out_sum = group_bc_partial(g, group_members(g, group, n_group), to_compact(g, sources, n_sources));
    });
}

int rxc_group_betweenness(rxc_graph *g, const uint32_t *group, uint64_t n_group, int normalized,
                          double *out) {
    if (!g || !out || (n_group && !group)) return RXC_ERR_INVALID;
    return guarded([&] {
        // group_betweenness_centrality, rustworkx-core/src/centrality.rs:1963-2101
        const std::vector<uint32_t> members = group_members(g, group, n_group);
        const uint64_t node_count = g->host.n_live, group_size = members.size();
        *out = 0.0;
        if (group_size == 0 || node_count <= 1) return;
        const size_t K = 1 + g->replicas.size();
        std::vector<double> parts(K, 0.0);
        run_on_shards(g, [&](rxc_graph *r, size_t k, size_t nk) {
            parts[k] = group_bc_partial(r, members, shard_sources(r->host.n_live, k, nk));
        });
        double gb = 0.0;
        for (double x : parts) gb += x;
        if (!g->host.directed) gb /= 2.0;
        if (normalized) {
            const uint64_t non_group = node_count - group_size;
            if (non_group > 1) {
                const double norm = g->host.directed ? (double)(non_group * (non_group - 1))
                                                     : (double)((non_group * (non_group - 1)) / 2);
                gb /= norm;
            }
        }
        *out = gb;
    });
}

int rxc_group_closeness(rxc_graph *g, const uint32_t *group, uint64_t n_group, double *out) {
    if (!g || !out || (n_group && !group)) return RXC_ERR_INVALID;
    return guarded([&] {
        // group_closeness_centrality, rustworkx-core/src/centrality.rs:1860-1916
        const std::vector<uint32_t> members = group_members(g, group, n_group);
        const uint64_t node_count = g->host.n_live, group_size = members.size();
        *out = 0.0;
        if (group_size >= node_count) return;
        const uint64_t dist_sum = group_closeness_dsum(g, members);
        if (dist_sum == 0) return;
        *out = (double)(node_count - group_size) / (double)dist_sum;
    });
}


int rxc_weighted_closeness_partial(rxc_graph *g, const double *edge_weight, int invert, int reversed,
                                   const uint32_t *sources, uint64_t n_sources, int wf_improved,
                                   double *out) {
    if (!g || !out || (n_sources && !sources) || (g->host.e_bound && !edge_weight)) return RXC_ERR_INVALID;
    return guarded([&] {
        const std::vector<uint32_t> src = to_compact(g, sources, n_sources);
        // duplicates: closeness is per source, so compute each distinct source once
        std::vector<uint32_t> uniq = src;
        std::sort(uniq.begin(), uniq.end());
        uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
        std::vector<double> vals(uniq.size());
        wsp_closeness(g, edge_weight, invert != 0, reversed != 0, uniq, wf_improved != 0, vals.data());
        for (size_t i = 0; i < src.size(); i++)
            out[i] = vals[std::lower_bound(uniq.begin(), uniq.end(), src[i]) - uniq.begin()];
    });
}

int rxc_newman_weighted_closeness(rxc_graph *g, const double *edge_weight, int wf_improved, double *out) {
    if (!g || !out || (g->host.e_bound && !edge_weight)) return RXC_ERR_INVALID;
    return guarded([&] {
        // newman_weighted_closeness_centrality, rustworkx-core/src/centrality.rs:1297-1357:
        // dijkstra over Reversed(&graph) with cost 1 / weight
        std::fill(out, out + g->host.n_bound, 0.0);
        run_on_shards(g, [&](rxc_graph *r, size_t k, size_t nk) {
            const std::vector<uint32_t> src = shard_sources(r->host.n_live, k, nk);
            std::vector<double> vals(src.size());
            wsp_closeness(r, edge_weight, true, true, src, wf_improved != 0, vals.data());
            for (size_t i = 0; i < src.size(); i++) out[r->host.orig_of(src[i])] = vals[i];
        });
    });
}

int rxc_all_pairs_dijkstra_lengths_rows(rxc_graph *g, const double *edge_weight, double unreachable_value,
                                        uint32_t row_begin, uint32_t row_end, double *out) {
    if (!g || (!out && row_end > row_begin) || (g->host.e_bound && !edge_weight)) return RXC_ERR_INVALID;
    return guarded([&] {
        if (row_begin > row_end || row_end > g->host.n_live)
            throw std::invalid_argument("row range outside [0, n_live]");
        // all_pairs_dijkstra_path_lengths, src/shortest_path/all_pairs_dijkstra.rs:36-105: dijkstra
        // along outgoing edges from every node
        run_on_shards(g, [&](rxc_graph *r, size_t k, size_t nk) {
            const uint64_t rows = row_end - row_begin;
            const uint32_t b = row_begin + (uint32_t)(rows * k / nk), e = row_begin + (uint32_t)(rows * (k + 1) / nk);
            if (b == e) return;
            std::lock_guard<std::mutex> lock(r->mu);
            begin_call(r);
            const uint32_t n = r->host.n_live;
            std::vector<uint32_t> src(e - b);
            for (uint32_t i = b; i < e; i++) src[i - b] = i;
            double *dst = out + (size_t)(b - row_begin) * n;
            WspCosts c;
            wsp_costs(r, edge_weight, false, false, c);
            wsp_run(r, false, c.pull.p, src, WspMode::Rows, false, dst, nullptr);
            const size_t len = (size_t)(e - b) * n;
            const double inf = std::numeric_limits<double>::infinity();
            if (has_inf_cost(r, edge_weight, nullptr, false)) {
                // reached through +inf-cost edges only: keep +inf; never reached: unreachable_value
                std::vector<double> hops(len);
                bfs_dispatch(r, r->succ(), src, BfsMode::Rows, false, -1.0, hops.data());
                for (size_t i = 0; i < len; i++)
                    if (hops[i] < 0.0) dst[i] = unreachable_value;
            } else if (!(unreachable_value == inf)) {
                for (size_t i = 0; i < len; i++)
                    if (dst[i] == inf) dst[i] = unreachable_value;
            }
            end_call(r);
            fetch_stats(r);
        });
    });
}

int rxc_closeness_partial(rxc_graph *g, const uint32_t *sources, uint64_t n_sources,
                          int wf_improved, double *out) {
    if (!g || !out || (n_sources && !sources)) return RXC_ERR_INVALID;
    return guarded([&] {
        std::lock_guard<std::mutex> lock(g->mu);
        begin_call(g);
        // closeness follows incoming edges: BFS over Reversed(&graph), centrality.rs:1208
        bfs_dispatch(g, g->pred(), to_compact(g, sources, n_sources), BfsMode::Closeness,
                     wf_improved != 0, 0.0, out);
        end_call(g);
        fetch_stats(g);
    });
}

int rxc_closeness(rxc_graph *g, int wf_improved, double *out) {
    if (!g || !out) return RXC_ERR_INVALID;
    return guarded([&] {
        std::fill(out, out + g->host.n_bound, 0.0);
        run_on_shards(g, [&](rxc_graph *r, size_t k, size_t nk) {
            std::lock_guard<std::mutex> lock(r->mu);
            begin_call(r);
            const std::vector<uint32_t> src = shard_sources(r->host.n_live, k, nk);
            std::vector<double> vals(src.size());
            // closeness follows incoming edges: BFS over Reversed(&graph), centrality.rs:1208
            bfs_dispatch(r, r->pred(), src, BfsMode::Closeness, wf_improved != 0, 0.0, vals.data());
            end_call(r);
            for (size_t i = 0; i < src.size(); i++) out[r->host.orig_of(src[i])] = vals[i];
            fetch_stats(r);
        });
    });
}

int rxc_distance_matrix_rows(rxc_graph *g, int as_undirected, double null_value,
                             uint32_t row_begin, uint32_t row_end, double *out) {
    if (!g || (!out && row_end > row_begin)) return RXC_ERR_INVALID;
    return guarded([&] {
        if (row_begin > row_end || row_end > g->host.n_live)
            throw std::invalid_argument("row range outside [0, n_live]");
        // every shard fills a contiguous block of rows of the caller's matrix
        run_on_shards(g, [&](rxc_graph *r, size_t k, size_t nk) {
            const uint64_t rows = row_end - row_begin;
            const uint32_t b = row_begin + (uint32_t)(rows * k / nk), e = row_begin + (uint32_t)(rows * (k + 1) / nk);
            std::lock_guard<std::mutex> lock(r->mu);
            begin_call(r);
            std::vector<uint32_t> src(e - b);
            for (uint32_t i = b; i < e; i++) src[i - b] = i;
            // graph_distance_matrix always passes as_undirected=true (src/shortest_path/mod.rs:1610)
            const DeviceCSR &csr = (as_undirected || !r->host.directed) ? symmetric_csr(r) : r->out;
            bfs_dispatch(r, csr, src, BfsMode::Rows, false, null_value,
                         out + (size_t)(b - row_begin) * r->host.n_live);
            end_call(r);
            fetch_stats(r);
        });
    });
}

int rxc_distance_sums(rxc_graph *g, int as_undirected, uint64_t *total_distance,
                      uint64_t *connected_pairs) {
    if (!g || !total_distance || !connected_pairs) return RXC_ERR_INVALID;
    return guarded([&] {
        const size_t K = 1 + g->replicas.size();
        std::vector<uint64_t> sums(K, 0), pairs(K, 0);
        run_on_shards(g, [&](rxc_graph *r, size_t k, size_t nk) {
            std::lock_guard<std::mutex> lock(r->mu);
            begin_call(r);
            const std::vector<uint32_t> src = shard_sources(r->host.n_live, k, nk);
            std::vector<double> vals(src.size() * 2);
            // outgoing neighbours; both directions for an undirected graph or as_undirected
            const DeviceCSR &csr = (as_undirected || !r->host.directed) ? symmetric_csr(r) : r->out;
            bfs_dispatch(r, csr, src, BfsMode::Sums, false, 0.0, vals.data());
            end_call(r);
            fetch_stats(r);
            for (size_t i = 0; i < src.size(); i++) {
                pairs[k] += (uint64_t)vals[2 * i] - 1;
                sums[k] += (uint64_t)vals[2 * i + 1];
            }
        });
        uint64_t sum = 0, prs = 0;
        for (size_t k = 0; k < K; k++) {
            sum += sums[k];
            prs += pairs[k];
        }
        *total_distance = sum;
        *connected_pairs = prs;
    });
}

int rxc_distance_matrix(rxc_graph *g, int as_undirected, double null_value, double *out) {
    if (!g) return RXC_ERR_INVALID;
    return rxc_distance_matrix_rows(g, as_undirected, null_value, 0, g->host.n_live, out);
}

int rxc_get_stats(const rxc_graph *g, rxc_stats *out) {
    if (!g || !out) return RXC_ERR_INVALID;
    *out = g->stats;
    return RXC_OK;
}

int rxc_set_option(const char *name, int64_t value) {
    if (!name) return RXC_ERR_INVALID;
    const std::string k(name);
    std::lock_guard<std::mutex> lock(g_opt_mu);
    if (k == "bc_groups") g_opt.bc_groups = value;
    else if (k == "bfs_groups") g_opt.bfs_groups = value;
    else if (k == "hub_degree" && value > 0) g_opt.hub_degree = value;
    else if (k == "bfs_mode" && value >= 0 && value <= 3) g_opt.bfs_mode = value;
    else if (k == "bc_order" && value >= 0 && value <= 2) g_opt.bc_order = value;
    else if (k == "bc_siblings" && value >= 1) g_opt.bc_siblings = value;
    else if (k == "bc_bwd_ctas" && value >= 0) g_opt.bc_bwd_ctas = value;
    else if (k == "bc_look" && value >= 0 && value <= 2) g_opt.bc_look = value;
    else if (k == "bfs_depth_switch" && value >= 0) g_opt.bfs_depth_switch = value;
    else if (k == "bfs_threads" && (value == 0 || value == 128 || value == 256 || value == 512 || value == 1024)) g_opt.bfs_threads = value;
    else if (k == "bc_sigma" && value >= 0 && value <= 2) g_opt.bc_sigma = value;
    else if (k == "bfs_ell" && (value == 0 || value == 1)) g_opt.bfs_ell = value;
    else if (k == "device_csr" && (value == 0 || value == 1)) g_opt.device_csr = value;
    else if (k == "bc_sparse_div" && value >= 0) g_opt.bc_sparse_div = value;
    else if (k == "mem_fraction_pct" && value > 0 && value <= 95) g_opt.mem_fraction_pct = value;
    else if (k == "snapshot_cache_entries" && value >= 0 && value <= 64) {
        std::lock_guard<std::mutex> cl(g_cache_mu);
        g_cache_cap = (size_t)value;
        while (g_cache.size() > g_cache_cap) {
            rxc_graph *old = g_cache.back();
            g_cache.pop_back();
            graph_unref_locked(old);
        }
    }
    else {
        set_error("unknown option: " + k);
        return RXC_ERR_INVALID;
    }
    return RXC_OK;
}

}  // extern "C"
