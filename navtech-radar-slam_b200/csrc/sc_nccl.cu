// sc_nccl.cu -- the collective form of the sharded ScanContext search: north_star's "one NCCL all-reduce
// of the per-shard min distance over NVLink" (SURVEY 8e), for hosts that prefer a collective to the
// peer-memory exchange of sc_search_bin.cu / sc_kernels.cu (or cannot map peers by CUDA IPC).
//
//   every rank:  local search of its slice -> per query (dist, global idx, shift)
//   reduce 1:    ncclAllReduce(min) on the f64 distances            (nothing found = 1e7, the reference's sentinel)
//   reduce 2:    ncclAllReduce(min) on int64 (idx << 8 | shift), where only ranks whose distance EQUALS the
//                global minimum contribute, the others send INT64_MAX -- the lowest global index among equal
//                distances wins, exactly like the sequential strict-'<' scan (Scancontext.cpp:380-396)
// Both reductions carry Q elements for a batch of Q queries (16 B per query in total).
//
// libnccl is opened at run time (dlopen): the library has no link-time dependency on NCCL, and inside a
// PyTorch process the already-loaded NCCL is reused.  The communicator is created here from a 128-byte
// unique id the host program distributes (any channel), or an existing ncclComm_t is attached.
#include <dlfcn.h>

#include <algorithm>

#include "sc_common.cuh"

using namespace nrs_sc;

namespace {

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void *ncclComm_p;
typedef int (*fn_get_unique_id)(ncclUniqueId_t *);
typedef int (*fn_comm_init_rank)(ncclComm_p *, int, ncclUniqueId_t, int);
typedef int (*fn_comm_destroy)(ncclComm_p);
typedef int (*fn_all_reduce)(const void *, void *, size_t, int, int, ncclComm_p, cudaStream_t);
typedef const char *(*fn_error_string)(int);
constexpr int NCCL_INT64 = 4, NCCL_FLOAT64 = 8, NCCL_MIN = 3;      // nccl.h: ncclInt64, ncclFloat64, ncclMin

struct NcclApi {
    void *handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_error_string error_string = nullptr;
};

NcclApi *nccl_api()
{
    static std::mutex mu;
    static NcclApi api;
    static bool tried = false;
    std::lock_guard<std::mutex> lk(mu);
    if (!tried) {
        tried = true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (api.handle) {
            api.get_unique_id = (fn_get_unique_id)dlsym(api.handle, "ncclGetUniqueId");
            api.comm_init_rank = (fn_comm_init_rank)dlsym(api.handle, "ncclCommInitRank");
            api.comm_destroy = (fn_comm_destroy)dlsym(api.handle, "ncclCommDestroy");
            api.all_reduce = (fn_all_reduce)dlsym(api.handle, "ncclAllReduce");
            api.error_string = (fn_error_string)dlsym(api.handle, "ncclGetErrorString");
            if (!api.get_unique_id || !api.comm_init_rank || !api.comm_destroy || !api.all_reduce) api.handle = nullptr;
        }
    }
    return api.handle ? &api : nullptr;
}

#define NRS_NCCL(api, expr)                                                                              \
    do {                                                                                                 \
        int r__ = (expr);                                                                                \
        if (r__ != 0) {                                                                                  \
            nrs::set_error("%s -> NCCL error %d (%s)", #expr, r__, (api)->error_string ? (api)->error_string(r__) : "?"); \
            return NRS_ENCCL;                                                                            \
        }                                                                                                \
    } while (0)

__global__ void sc_nccl_pack_dist(const nrs_sc_result *__restrict__ local, int Q, double *__restrict__ dist)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < Q) {
        const double d = local[q].dist;
        dist[q] = (d < SENTINEL) ? d : SENTINEL;          // NaN never leaves the search; nothing found -> 1e7
    }
}
__global__ void sc_nccl_pack_key(const nrs_sc_result *__restrict__ local, const double *__restrict__ gmin, int Q,
                                 long long *__restrict__ key)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < Q) {
        const nrs_sc_result r = local[q];
        const bool mine = (r.dist < SENTINEL) && (r.dist == gmin[q]);
        key[q] = mine ? (((long long)r.idx << 8) | (long long)(r.shift & 0xff)) : 0x7fffffffffffffffll;
    }
}
__global__ void sc_nccl_unpack(const double *__restrict__ gmin, const long long *__restrict__ gkey, int Q,
                               nrs_sc_result *__restrict__ out)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < Q) {
        nrs_sc_result r;
        const long long k = gkey[q];
        if (gmin[q] < SENTINEL && k != 0x7fffffffffffffffll) { r.dist = gmin[q]; r.idx = (int32_t)(k >> 8); r.shift = (int32_t)(k & 0xff); }
        else { r.dist = SENTINEL; r.idx = 0; r.shift = 0; }
        out[q] = r;
    }
}

int ensure_buffers(nrs_sc_db *db)
{
    if (db->n_dist) return NRS_OK;
    NRS_CUDA(cudaMalloc(&db->n_dist, 2 * XCHG_QCAP * sizeof(double)));
    NRS_CUDA(cudaMalloc(&db->n_key, 2 * XCHG_QCAP * sizeof(long long)));
    return NRS_OK;
}

// queries on the device, results on the device, everything enqueued on the context stream
int batch_dev_nccl_locked(nrs_sc_db *db, const double *queries_dev, int Q, int first, int last, nrs_sc_result *results_dev)
{
    NcclApi *api = nccl_api();
    if (!api || !db->n_comm) { nrs::set_error("NCCL exchange not initialised (nrs_sc_nccl_init / nrs_sc_nccl_attach)"); return NRS_ENCCL; }
    nrs_ctx *ctx = db->ctx;
    cudaStream_t st = ctx->stream;
    for (int done = 0; done < Q; done += XCHG_QCAP) {
        const int q = std::min(XCHG_QCAP, Q - done);
        // local search: results carry GLOBAL keyframe indices (local slot * world + rank)
        int rc = nrs_sc::batch_dev_locked(db, queries_dev + (size_t)done * DESC, q, first, last, db->n_world, db->n_rank, db->d_results,
                                          false);
        if (rc) return rc;
        double *gmin = db->n_dist + XCHG_QCAP;
        long long *gkey = db->n_key + XCHG_QCAP;
        const int tb = 128, gb = (q + tb - 1) / tb;
        sc_nccl_pack_dist<<<gb, tb, 0, st>>>(db->d_results, q, db->n_dist);
        NRS_LAUNCHED();
        NRS_NCCL(api, api->all_reduce(db->n_dist, gmin, (size_t)q, NCCL_FLOAT64, NCCL_MIN, db->n_comm, st));
        sc_nccl_pack_key<<<gb, tb, 0, st>>>(db->d_results, gmin, q, db->n_key);
        NRS_LAUNCHED();
        NRS_NCCL(api, api->all_reduce(db->n_key, gkey, (size_t)q, NCCL_INT64, NCCL_MIN, db->n_comm, st));
        sc_nccl_unpack<<<gb, tb, 0, st>>>(gmin, gkey, q, results_dev + done);
        NRS_LAUNCHED();
    }
    return NRS_OK;
}

}  // namespace

extern "C" {

int nrs_nccl_available(void) { return nccl_api() ? 1 : 0; }

int nrs_nccl_get_unique_id(unsigned char *id_out)
{
    NRS_CHECK_ARG(id_out, "bad argument");
    NcclApi *api = nccl_api();
    if (!api) { nrs::set_error("libnccl.so.2 not found (dlopen)"); return NRS_ENCCL; }
    ncclUniqueId_t id;
    NRS_NCCL(api, api->get_unique_id(&id));
    memcpy(id_out, &id, sizeof(id));
    return NRS_OK;
}

int nrs_sc_nccl_init(nrs_sc_db *db, int world, int rank, const unsigned char *unique_id)
{
    NRS_CHECK_ARG(db && unique_id && world >= 1 && rank >= 0 && rank < world, "bad argument");
    NcclApi *api = nccl_api();
    if (!api) { nrs::set_error("libnccl.so.2 not found (dlopen)"); return NRS_ENCCL; }
    std::lock_guard<std::mutex> lk(db->mu);
    NRS_CHECK_ARG(db->n_comm == nullptr, "NCCL exchange already initialised for this database");
    NRS_CUDA(cudaSetDevice(db->ctx->device));
    ncclUniqueId_t id;
    memcpy(&id, unique_id, sizeof(id));
    ncclComm_p comm = nullptr;
    NRS_NCCL(api, api->comm_init_rank(&comm, world, id, rank));
    db->n_comm = comm;
    db->n_owned = true;
    db->n_world = world;
    db->n_rank = rank;
    return ensure_buffers(db);
}

int nrs_sc_nccl_attach(nrs_sc_db *db, void *nccl_comm, int world, int rank)
{
    NRS_CHECK_ARG(db && nccl_comm && world >= 1 && rank >= 0 && rank < world, "bad argument");
    if (!nccl_api()) { nrs::set_error("libnccl.so.2 not found (dlopen)"); return NRS_ENCCL; }
    std::lock_guard<std::mutex> lk(db->mu);
    NRS_CHECK_ARG(db->n_comm == nullptr, "NCCL exchange already initialised for this database");
    NRS_CUDA(cudaSetDevice(db->ctx->device));
    db->n_comm = nccl_comm;
    db->n_owned = false;
    db->n_world = world;
    db->n_rank = rank;
    return ensure_buffers(db);
}

int nrs_sc_query_batch_dev_nccl(nrs_sc_db *db, const double *queries_dev, int Q, int first, int last, nrs_sc_result *results_dev)
{
    NRS_CHECK_ARG(db && queries_dev && results_dev && Q >= 1, "bad argument");
    std::lock_guard<std::mutex> lk(db->mu);
    NRS_CHECK_ARG(first >= 0 && last <= db->size, "search range outside the database");
    std::lock_guard<std::mutex> lk2(db->ctx->mu);
    NRS_CUDA(cudaSetDevice(db->ctx->device));
    return batch_dev_nccl_locked(db, queries_dev, Q, first, last, results_dev);
}

// host queries in, host records out: the call a C++ host makes (no CUDA types in its code)
int nrs_sc_query_batch_sharded(nrs_sc_db *db, const double *queries_host, int Q, int first, int last, nrs_sc_result *results_host)
{
    NRS_CHECK_ARG(db && queries_host && results_host && Q >= 1, "bad argument");
    std::lock_guard<std::mutex> lk(db->mu);
    NRS_CHECK_ARG(db->x_ready || db->n_comm, "no exchange set up (nrs_sc_xchg_open or nrs_sc_nccl_init)");
    NRS_CHECK_ARG(first >= 0 && last <= db->size, "search range outside the database");
    nrs_ctx *ctx = db->ctx;
    std::lock_guard<std::mutex> lk2(ctx->mu);
    NRS_CUDA(cudaSetDevice(ctx->device));
    if (!db->d_results2) NRS_CUDA(cudaMalloc(&db->d_results2, XCHG_QCAP * sizeof(nrs_sc_result)));
    for (int done = 0; done < Q; done += XCHG_QCAP) {
        const int q = std::min(XCHG_QCAP, Q - done);
        NRS_CUDA(cudaMemcpyAsync(db->d_queries, queries_host + (size_t)done * DESC, (size_t)q * DESC * sizeof(double),
                                 cudaMemcpyHostToDevice, ctx->stream));
        int rc = db->x_ready ? nrs_sc::batch_dev_locked(db, db->d_queries, q, first, last, db->x_world, db->x_rank, db->d_results2, true)
                             : batch_dev_nccl_locked(db, db->d_queries, q, first, last, db->d_results2);
        if (rc) return rc;
        NRS_CUDA(cudaMemcpyAsync(results_host + done, db->d_results2, (size_t)q * sizeof(nrs_sc_result), cudaMemcpyDeviceToHost, ctx->stream));
        NRS_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return NRS_OK;
}

}  // extern "C"

// called by nrs_sc_db_destroy
void nrs_sc::nccl_release(nrs_sc_db *db)
{
    if (db->n_comm && db->n_owned) {
        NcclApi *api = nccl_api();
        if (api) api->comm_destroy(db->n_comm);
    }
    db->n_comm = nullptr;
    cudaFree(db->n_dist);
    cudaFree(db->n_key);
    cudaFree(db->d_results2);
    db->n_dist = nullptr; db->n_key = nullptr; db->d_results2 = nullptr;
}
