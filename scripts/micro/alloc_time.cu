// How long does it take to get 30 GB of fresh device memory?  cudaMalloc vs the stream-ordered pool.
#include <cuda_runtime.h>
#include <stdio.h>
#include <chrono>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void run_malloc();
static void run_pool();
int main(int argc, char **argv) {
    cudaFree(0);
    if (argc > 1) { run_pool(); run_malloc(); } else { run_malloc(); run_pool(); }
}
static const size_t GB = 1ull << 30;
static size_t sizes[] = {16 * GB, 8 * GB, 4 * GB, 2 * GB};
static void run_malloc() {
    for (int pass = 0; pass < 2; pass++) {
        void *p[4];
        double t0 = now();
        for (int i = 0; i < 4; i++) cudaMalloc(&p[i], sizes[i]);
        cudaDeviceSynchronize();
        double t1 = now();
        for (int i = 0; i < 4; i++) cudaMemset(p[i], 0, sizes[i]);
        cudaDeviceSynchronize();
        double t2 = now();
        for (int i = 0; i < 4; i++) cudaFree(p[i]);
        double t3 = now();
        printf("cudaMalloc      30 GB: alloc %.1f ms, memset %.1f ms, free %.1f ms\n", t1 - t0, t2 - t1, t3 - t2);
    }
}
static void run_pool() {
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaMemPool_t pool;
    cudaDeviceGetDefaultMemPool(&pool, 0);
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    for (int pass = 0; pass < 2; pass++) {
        void *p[4];
        double t0 = now();
        for (int i = 0; i < 4; i++) cudaMallocAsync(&p[i], sizes[i], st);
        cudaStreamSynchronize(st);
        double t1 = now();
        for (int i = 0; i < 4; i++) cudaMemsetAsync(p[i], 0, sizes[i], st);
        cudaStreamSynchronize(st);
        double t2 = now();
        for (int i = 0; i < 4; i++) cudaFreeAsync(p[i], st);
        cudaStreamSynchronize(st);
        double t3 = now();
        printf("cudaMallocAsync 30 GB: alloc %.1f ms, memset %.1f ms, free %.1f ms (pass %d)\n", t1 - t0, t2 - t1, t3 - t2, pass);
    }
}
