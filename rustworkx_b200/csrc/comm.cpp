// comm.cpp -- the one collective of the path: a sum-reduce of per-rank partial centrality vectors.
//
// The reference is single-process (rayon threads share one Vec behind an RwLock,
// rustworkx-core/src/centrality.rs:107,280,317).  Here every rank (one process per GPU) owns a shard
// of the sources and a partial vector; partial vectors are additive, so a single ncclReduce over
// NVLink at the end replaces the lock.  NCCL is dlopen()ed so the library carries no link-time
// dependency; when the caller already loaded NCCL (e.g. through torch) that copy is reused.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <string.h>

#include <string>

#include "../../include/rxcuda.h"

namespace rxc {
void set_error(const std::string &msg);
int comm_reduce_device(double *d_inout, uint64_t len, int root, cudaStream_t st);
int comm_rendezvous_device(cudaStream_t st);
int comm_rank();
int comm_world();
}

namespace {

typedef struct {
    char internal[128];
} nccl_unique_id;
typedef void *nccl_comm_t;
typedef int nccl_result_t;
enum { NCCL_SUM = 0, NCCL_FLOAT64 = 8, NCCL_INT32 = 2 };

struct NcclApi {
    void *handle = nullptr;
    nccl_result_t (*GetUniqueId)(nccl_unique_id *) = nullptr;
    nccl_result_t (*CommInitRank)(nccl_comm_t *, int, nccl_unique_id, int) = nullptr;
    nccl_result_t (*CommDestroy)(nccl_comm_t) = nullptr;
    nccl_result_t (*Reduce)(const void *, void *, size_t, int, int, int, nccl_comm_t,
                            cudaStream_t) = nullptr;
    nccl_result_t (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t,
                               cudaStream_t) = nullptr;
    const char *(*GetErrorString)(nccl_result_t) = nullptr;
};

NcclApi g_nccl;
nccl_comm_t g_comm = nullptr;
cudaStream_t g_comm_stream = nullptr;
double *g_red_buf = nullptr;  // device staging of rxc_comm_reduce_sum_f64
uint64_t g_red_cap = 0;
int g_rank = 0, g_world = 1;

bool load_nccl() {
    if (g_nccl.handle) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        rxc::set_error(std::string("cannot dlopen libnccl.so.2: ") + dlerror());
        return false;
    }
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
    a.Reduce = (decltype(a.Reduce))dlsym(h, "ncclReduce");
    a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.Reduce || !a.AllReduce) {
        rxc::set_error("libnccl is missing required symbols");
        return false;
    }
    g_nccl = a;
    return true;
}

int nccl_fail(const char *what, nccl_result_t r) {
    rxc::set_error(std::string(what) + ": " +
                   (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error"));
    return RXC_ERR_NCCL;
}

int cuda_fail(const char *what, cudaError_t e) {
    rxc::set_error(std::string(what) + ": " + cudaGetErrorString(e));
    cudaGetLastError();
    return RXC_ERR_CUDA;
}

}  // namespace

namespace rxc {
// In-place sum-reduce of a DEVICE vector to `root`, queued on the caller's stream: the partial
// centrality vector never leaves HBM before it is combined (rxc_*_partial_reduce).  Without a
// communicator (single process) this is the identity.
int comm_reduce_device(double *d_inout, uint64_t len, int root, cudaStream_t st) {
    if (!g_comm || g_world == 1 || len == 0) return RXC_OK;
    nccl_result_t r = g_nccl.Reduce(d_inout, d_inout, len, NCCL_FLOAT64, NCCL_SUM, root, g_comm, st);
    if (r != 0) return nccl_fail("ncclReduce", r);
    return RXC_OK;
}
// Device-side rendezvous queued on the caller's stream (an all-reduce of one int): events around it
// time how long this rank waited for the slowest one, apart from the transfer of the real reduce.
int comm_rendezvous_device(cudaStream_t st) {
    if (!g_comm || g_world == 1) return RXC_OK;
    static int *d_flag[64] = {nullptr};  // one scratch word per device, kept for the process
    int dev = 0;
    cudaGetDevice(&dev);
    int *&d = d_flag[dev & 63];
    if (!d) {
        cudaError_t e = cudaMalloc((void **)&d, sizeof(int));
        if (e != cudaSuccess) return cuda_fail("cudaMalloc", e);
        cudaMemsetAsync(d, 0, sizeof(int), st);
    }
    nccl_result_t r = g_nccl.AllReduce(d, d, 1, NCCL_INT32, NCCL_SUM, g_comm, st);
    if (r != 0) return nccl_fail("ncclAllReduce", r);
    return RXC_OK;
}
int comm_rank() { return g_rank; }
int comm_world() { return g_comm ? g_world : 1; }
}  // namespace rxc

extern "C" {

int rxc_comm_unique_id(uint8_t id_out[128]) {
    if (!id_out) return RXC_ERR_INVALID;
    if (!load_nccl()) return RXC_ERR_NCCL;
    nccl_unique_id id;
    nccl_result_t r = g_nccl.GetUniqueId(&id);
    if (r != 0) return nccl_fail("ncclGetUniqueId", r);
    memcpy(id_out, id.internal, 128);
    return RXC_OK;
}

int rxc_comm_init(int rank, int world, const uint8_t id[128]) {
    if (!id || world < 1 || rank < 0 || rank >= world) return RXC_ERR_INVALID;
    if (!load_nccl()) return RXC_ERR_NCCL;
    if (g_comm) rxc_comm_destroy();
    cudaError_t e = cudaSetDevice(rxc_get_device());
    if (e != cudaSuccess) return cuda_fail("cudaSetDevice", e);
    e = cudaStreamCreateWithFlags(&g_comm_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return cuda_fail("cudaStreamCreate", e);
    nccl_unique_id uid;
    memcpy(uid.internal, id, 128);
    nccl_result_t r = g_nccl.CommInitRank(&g_comm, world, uid, rank);
    if (r != 0) {
        g_comm = nullptr;
        cudaStreamDestroy(g_comm_stream);
        g_comm_stream = nullptr;
        return nccl_fail("ncclCommInitRank", r);
    }
    g_rank = rank;
    g_world = world;
    return RXC_OK;
}

int rxc_comm_reduce_sum_f64(double *host_inout, uint64_t len, int root) {
    if (!g_comm) {
        rxc::set_error("rxc_comm_init has not been called");
        return RXC_ERR_NCCL;
    }
    if (len == 0) return RXC_OK;
    if (!host_inout) return RXC_ERR_INVALID;
    cudaError_t e = cudaSetDevice(rxc_get_device());
    if (e != cudaSuccess) return cuda_fail("cudaSetDevice", e);
    // staging buffer kept across calls: cudaMalloc / cudaFree synchronise the whole device
    if (g_red_cap < len) {
        if (g_red_buf) cudaFree(g_red_buf);
        g_red_buf = nullptr;
        g_red_cap = 0;
        e = cudaMalloc((void **)&g_red_buf, len * sizeof(double));
        if (e != cudaSuccess) return cuda_fail("cudaMalloc", e);
        g_red_cap = len;
    }
    double *d = g_red_buf;
    int rc = RXC_OK;
    e = cudaMemcpyAsync(d, host_inout, len * sizeof(double), cudaMemcpyHostToDevice, g_comm_stream);
    if (e != cudaSuccess) rc = cuda_fail("cudaMemcpyAsync", e);
    if (rc == RXC_OK) {
        nccl_result_t r = g_nccl.Reduce(d, d, len, NCCL_FLOAT64, NCCL_SUM, root, g_comm, g_comm_stream);
        if (r != 0) rc = nccl_fail("ncclReduce", r);
    }
    if (rc == RXC_OK && g_rank == root) {
        e = cudaMemcpyAsync(host_inout, d, len * sizeof(double), cudaMemcpyDeviceToHost, g_comm_stream);
        if (e != cudaSuccess) rc = cuda_fail("cudaMemcpyAsync", e);
    }
    e = cudaStreamSynchronize(g_comm_stream);
    if (e != cudaSuccess && rc == RXC_OK) rc = cuda_fail("cudaStreamSynchronize", e);
    return rc;
}

int rxc_comm_barrier(void) {
    if (!g_comm) {
        rxc::set_error("rxc_comm_init has not been called");
        return RXC_ERR_NCCL;
    }
    cudaError_t e = cudaSetDevice(rxc_get_device());
    if (e != cudaSuccess) return cuda_fail("cudaSetDevice", e);
    int *d = nullptr;
    e = cudaMalloc((void **)&d, sizeof(int));
    if (e != cudaSuccess) return cuda_fail("cudaMalloc", e);
    cudaMemsetAsync(d, 0, sizeof(int), g_comm_stream);
    int rc = RXC_OK;
    nccl_result_t r = g_nccl.AllReduce(d, d, 1, NCCL_INT32, NCCL_SUM, g_comm, g_comm_stream);
    if (r != 0) rc = nccl_fail("ncclAllReduce", r);
    e = cudaStreamSynchronize(g_comm_stream);
    if (e != cudaSuccess && rc == RXC_OK) rc = cuda_fail("cudaStreamSynchronize", e);
    cudaFree(d);
    return rc;
}

void rxc_comm_destroy(void) {
    if (g_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g_comm);
    g_comm = nullptr;
    if (g_red_buf) cudaFree(g_red_buf);
    g_red_buf = nullptr;
    g_red_cap = 0;
    if (g_comm_stream) cudaStreamDestroy(g_comm_stream);
    g_comm_stream = nullptr;
    g_rank = 0;
    g_world = 1;
}

}  // extern "C"
