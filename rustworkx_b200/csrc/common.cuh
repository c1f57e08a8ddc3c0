// common.cuh -- shared device/host helpers for librxcuda (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace rxc {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int GW = 32;           // sources per group: one warp lane per source
constexpr int BLOCK = 256;       // threads per CTA for every kernel here
constexpr int WARPS = BLOCK / 32;
constexpr uint32_t NO_SOURCE = 0xFFFFFFFFu;

// Device-side work counters (one array per graph handle).
enum StatSlot { ST_TE = 0, ST_DAG = 1, ST_V = 2, ST_COUNT = 4 };
// The counters are striped: a launch of the dependency sweep has up to 2 million warps, and three
// atomics per warp on the SAME three addresses serialise in one L2 slice (and slow every other
// access that hashes to it).  A warp adds to the stripe its CTA index picks -- one 32-byte sector per
// stripe -- and the host sums the stripes (fetch_stats).
constexpr int ST_STRIPES = 256;

// RXC_PROBE builds (make variant VAR=probe DEFS=-DRXC_PROBE=1) count what the betweenness kernels ask of
// the memory system -- level-mask look-ups, gathered rows, sectors and 128-byte lines of those rows --
// and print the totals of every call to stderr; never timed, never shipped.
#ifndef RXC_PROBE
#define RXC_PROBE 0
#endif
#if RXC_PROBE
enum ProbeSlot { PR_LOOKUPS32 = 0, PR_ROWS32, PR_SECTORS32, PR_LINES32, PR_LOOKUPS64, PR_ROWS64, PR_SECTORS64, PR_LINES64, PR_COUNT };
__device__ unsigned long long g_probe[PR_COUNT];
#endif

void set_error(const std::string &msg);

struct CudaError {
    cudaError_t code;
    const char *file;
    int line;
};

#define RXC_CUDA(call)                                                      \
    do {                                                                    \
        cudaError_t _e = (call);                                            \
        if (_e != cudaSuccess) throw ::rxc::CudaError{_e, __FILE__, __LINE__}; \
    } while (0)

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ void stat_add(unsigned long long *stats, int slot, unsigned long long v) {
    const unsigned stripe = (blockIdx.x + blockIdx.y * 37u + (threadIdx.x >> 5) * 11u) & (ST_STRIPES - 1);
    atomicAdd(&stats[stripe * ST_COUNT + slot], v);
}

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}

}  // namespace rxc
