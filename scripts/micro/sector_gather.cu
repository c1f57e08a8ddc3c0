// Microbenchmark: what does the B200 memory system deliver for the access pattern of the betweenness
// row gathers?  Warps fetch, from random 1 KB rows (4 blocks of 256 bytes, 8 sectors of 32 bytes per
// block), `ns` sectors in each of `nb` blocks; the buffer is far larger than L2 (24 GB) or sits in it
// (48 MB).  Prints rows/s, block touches/s and GB/s of sector bytes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sector_gather sector_gather.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __host__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

__global__ void fill(uint4 *buf, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t h = mix((uint32_t)i);
        buf[i] = make_uint4(h, h * 3u, h * 5u, h * 7u);
    }
}

// lane l reads the 16-byte words l and l + 32 of a row (word w: block w / 16, sector (w / 2) % 8)
__global__ void __launch_bounds__(256) gather(const uint4 *buf, uint32_t nrows, int nb, int ns, int iters,
                                              unsigned long long *sink) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    bool on[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int w = lane + 32 * h, block = w / 16, sector = (w / 2) % 8;
        on[h] = block < nb && sector < ns;
    }
    uint32_t acc = 0;
    for (int it = 0; it < iters; it++) {
        uint4 v[4][2];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t r = mix(warp * 7919u + it * 4 + u) % nrows;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                v[u][h] = make_uint4(0, 0, 0, 0);
                if (on[h]) {
                    const uint4 *p = buf + (size_t)r * 64 + lane + 32 * h;
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(v[u][h].x), "=r"(v[u][h].y), "=r"(v[u][h].z), "=r"(v[u][h].w) : "l"(p));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc += v[u][0].x ^ v[u][0].w ^ v[u][1].x ^ v[u][1].w;
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

int main() {
    const size_t bytes = 24ull << 30;
    uint4 *buf;
    unsigned long long *sink;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&sink, 8);
    // (a cudaMemset leaves the pages in a cleared state that reads back without DRAM traffic)
    fill<<<148 * 16, 256>>>(buf, bytes / 16);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int ctas = 148 * 8 * 4, iters = 256;
    for (int small = 0; small < 2; small++) {
        const uint32_t nrows = small ? (48u << 20) / 1024 : (uint32_t)(bytes / 1024);
        for (int nb = 1; nb <= 4; nb *= 2) {
            for (int ns = 1; ns <= 8; ns *= 2) {
                float best = 1e30f;
                for (int rep = 0; rep < 3; rep++) {
                    cudaEventRecord(a);
                    gather<<<ctas, 256>>>(buf, nrows, nb, ns, iters, sink);
                    cudaEventRecord(b);
                    cudaEventSynchronize(b);
                    float ms;
                    cudaEventElapsedTime(&ms, a, b);
                    if (rep && ms < best) best = ms;
                }
                const double rows = (double)ctas * 8 * iters * 4;
                printf("%s  blocks/row %d  sectors/block %d: %6.2f Grows/s  %6.2f Gblocks/s  %7.1f GB/s  (%.2f ms)\n",
                       small ? "48 MB (L2)" : "24 GB     ", nb, ns, rows / 1e9 / (best / 1e3),
                       rows * nb / 1e9 / (best / 1e3), rows * nb * ns * 32 / 1e9 / (best / 1e3), best);
            }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
