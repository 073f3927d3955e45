#include "comm.cuh"

#include <dlfcn.h>
#include <cstring>
#include <mutex>

namespace gpirt {

namespace {
// the few NCCL entry points we need, bound at run time (ABI-stable across NCCL 2.x)
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int (*fn_get_uid)(ncclUniqueId_t*);
typedef int (*fn_init_rank)(void** comm, int nranks, ncclUniqueId_t id, int rank);
typedef int (*fn_allreduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t s);
typedef int (*fn_allgather)(const void* send, void* recv, size_t sendcount, int dtype, void* comm, cudaStream_t s);
typedef int (*fn_destroy)(void* comm);
typedef const char* (*fn_errstr)(int);

struct Api {
    void* handle = nullptr;
    fn_get_uid get_uid = nullptr; fn_init_rank init_rank = nullptr; fn_allreduce allreduce = nullptr; fn_allgather allgather = nullptr;
    fn_destroy destroy = nullptr; fn_errstr errstr = nullptr;
    bool tried = false;
} api;

int load_api() {
    if (api.handle) return GPIRT_B200_OK;
    if (!api.tried) {
        api.tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) { api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
        if (api.handle) {
            api.get_uid = (fn_get_uid)dlsym(api.handle, "ncclGetUniqueId");
            api.init_rank = (fn_init_rank)dlsym(api.handle, "ncclCommInitRank");
            api.allreduce = (fn_allreduce)dlsym(api.handle, "ncclAllReduce");
            api.allgather = (fn_allgather)dlsym(api.handle, "ncclAllGather");
            api.destroy = (fn_destroy)dlsym(api.handle, "ncclCommDestroy");
            api.errstr = (fn_errstr)dlsym(api.handle, "ncclGetErrorString");
            if (!api.get_uid || !api.init_rank || !api.allreduce || !api.allgather || !api.destroy) { dlclose(api.handle); api.handle = nullptr; }
        }
    }
    if (!api.handle) {
        const char* why = dlerror();   // NULL when the library opened but a symbol was missing
        set_last_error("NCCL (libnccl.so.2) could not be loaded: %s", why ? why : "a required entry point is missing");
        return GPIRT_B200_ERR_NCCL;
    }
    return GPIRT_B200_OK;
}

int check(int rc, const char* what) {
    if (rc == 0) return GPIRT_B200_OK;
    set_last_error("%s failed: %s", what, api.errstr ? api.errstr(rc) : "nccl error");
    return GPIRT_B200_ERR_NCCL;
}
}  // namespace

int comm_unique_id(void* out128) {
    GP_TRY(load_api());
    ncclUniqueId_t id;
    GP_TRY(check(api.get_uid(&id), "ncclGetUniqueId"));
    std::memcpy(out128, &id, sizeof(id));
    return GPIRT_B200_OK;
}

// The communicator outlives the sampler: ncclCommInitRank costs 0.3-3 s, far more than a short MCMC call.  A call that
// passes a unique id creates (and caches) a communicator; a call with world_size > 1 and NO id re-uses the cached one
// (same rank / world).  All ranks must take the same decision, as with any collective set-up.
// `users` counts the live samplers that hold the cached pointer: it is never destroyed under them.
static struct { void* comm = nullptr; int rank = -1, world = 0, users = 0; } g_cached;
static std::mutex g_comm_mu;

int comm_init(Comm& c, int rank, int world, const void* unique_id128) {
    c.rank = rank; c.world = world;
    if (world <= 1) return GPIRT_B200_OK;
    GP_TRY(load_api());
    std::lock_guard<std::mutex> lock(g_comm_mu);
    if (!unique_id128) {
        if (g_cached.comm && g_cached.rank == rank && g_cached.world == world) {
            c.nccl_comm = g_cached.comm;
            g_cached.users += 1;
            return GPIRT_B200_OK;
        }
        set_last_error("world_size > 1 needs opts.nccl_unique_id (no cached communicator for rank %d of %d)", rank, world);
        return GPIRT_B200_ERR_ARG;
    }
    if (g_cached.comm && g_cached.users > 0) {
        set_last_error("a new nccl_unique_id was passed while %d live sampler(s) still use the cached communicator", g_cached.users);
        return GPIRT_B200_ERR_ARG;
    }
    if (g_cached.comm) { api.destroy(g_cached.comm); g_cached.comm = nullptr; }
    ncclUniqueId_t id;
    std::memcpy(&id, unique_id128, sizeof(id));
    GP_TRY(check(api.init_rank(&c.nccl_comm, world, id, rank), "ncclCommInitRank"));
    g_cached.comm = c.nccl_comm; g_cached.rank = rank; g_cached.world = world; g_cached.users = 1;
    return GPIRT_B200_OK;
}

int comm_allreduce_sum_f64(Comm& c, double* buf, size_t count, cudaStream_t stream) {
    if (c.world <= 1) return GPIRT_B200_OK;
    const int ncclFloat64 = 8, ncclSum = 0;
    return check(api.allreduce(buf, buf, count, ncclFloat64, ncclSum, c.nccl_comm, stream), "ncclAllReduce");
}

int comm_allgather_f64(Comm& c, double* buf, size_t count_per_rank, cudaStream_t stream) {
    if (c.world <= 1) return GPIRT_B200_OK;
    const int ncclFloat64 = 8;
    return check(api.allgather(buf + (size_t)c.rank * count_per_rank, buf, count_per_rank, ncclFloat64, c.nccl_comm, stream), "ncclAllGather");
}

void comm_destroy(Comm& c) {   // the cached communicator stays alive for the next call
    std::lock_guard<std::mutex> lock(g_comm_mu);
    if (c.nccl_comm && c.nccl_comm == g_cached.comm && g_cached.users > 0) g_cached.users -= 1;
    c.nccl_comm = nullptr;
}

int comm_shutdown() {
    std::lock_guard<std::mutex> lock(g_comm_mu);
    if (g_cached.users > 0) {
        set_last_error("%d live sampler(s) still use the NCCL communicator", g_cached.users);
        return GPIRT_B200_ERR_ARG;
    }
    if (g_cached.comm && api.destroy) api.destroy(g_cached.comm);
    g_cached.comm = nullptr; g_cached.rank = -1; g_cached.world = 0;
    return GPIRT_B200_OK;
}

}  // namespace gpirt
