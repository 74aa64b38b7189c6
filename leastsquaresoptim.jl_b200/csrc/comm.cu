// comm.cu — NCCL plumbing for the row-sharded dense paths.  One process per GPU; the
// 128-byte unique id is produced by rank 0 (lso_comm_unique_id) and distributed by the host program
// (torch.distributed broadcast in bench.py / tests, MPI or a file on the Julia side).  NCCL is loaded
// lazily with dlopen so that the library has no link-time dependency on it: single-GPU users never touch it.
#include "common.cuh"
#include <dlfcn.h>
#include <stdlib.h>

typedef struct { char internal[128]; } lso_nccl_uid;
typedef void* lso_nccl_comm_t;
typedef int (*fn_GetUniqueId)(lso_nccl_uid*);
typedef int (*fn_CommInitRank)(lso_nccl_comm_t*, int, lso_nccl_uid, int);
typedef int (*fn_CommDestroy)(lso_nccl_comm_t);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, lso_nccl_comm_t, cudaStream_t);
typedef int (*fn_AllGather)(const void*, void*, size_t, int, lso_nccl_comm_t, cudaStream_t);
typedef const char* (*fn_GetErrorString)(int);

static struct {
    void* handle = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_AllReduce AllReduce = nullptr;
    fn_AllGather AllGather = nullptr;
    fn_GetErrorString GetErrorString = nullptr;
} g_nccl;

#define LSO_NCCL_FLOAT64 8
#define LSO_NCCL_SUM 0

static int nccl_load(lso_ctx* ctx) {
    if (g_nccl.handle) return LSO_OK;
    const char* env = getenv("LSO_NCCL_LIB");
    const char* cands[] = {env, "libnccl.so.2", "libnccl.so", nullptr};
    void* h = nullptr;
    for (int i = 0; i < 4 && !h; ++i) {
        if (!cands[i]) { if (i == 0) continue; else break; }
        h = dlopen(cands[i], RTLD_NOW | RTLD_GLOBAL);
    }
    if (!h) return lso_set_error(ctx, LSO_ERR_NCCL, "cannot dlopen libnccl.so.2 (set LSO_NCCL_LIB): %s", dlerror());
    g_nccl.GetUniqueId = (fn_GetUniqueId)dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (fn_CommInitRank)dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (fn_CommDestroy)dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (fn_AllReduce)dlsym(h, "ncclAllReduce");
    g_nccl.AllGather = (fn_AllGather)dlsym(h, "ncclAllGather");
    g_nccl.GetErrorString = (fn_GetErrorString)dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.AllGather)
        return lso_set_error(ctx, LSO_ERR_NCCL, "libnccl is missing required symbols");
    g_nccl.handle = h;
    return LSO_OK;
}

#define LSO_CHECK_NCCL(ctx, expr)                                                                  \
    do {                                                                                           \
        int _r = (expr);                                                                           \
        if (_r != 0)                                                                               \
            return lso_set_error((ctx), LSO_ERR_NCCL, "%s failed: %s", #expr,                      \
                                 g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error"); \
    } while (0)

extern "C" int lso_comm_allgather(lso_ctx* ctx, const double* d_send, double* d_recv, int64_t count) {
    LSO_REQUIRE(ctx, ctx && ctx->nccl_comm, "no communicator (call lso_comm_init_rank)");
    LSO_ENTER(ctx);
    LSO_CHECK_NCCL(ctx, g_nccl.AllGather(d_send, d_recv, (size_t)count, LSO_NCCL_FLOAT64, ctx->nccl_comm, ctx->stream));
    return LSO_OK;
}

extern "C" int lso_comm_allgather_on(lso_ctx* ctx, const double* d_send, double* d_recv, int64_t count, cudaStream_t st) {
    LSO_REQUIRE(ctx, ctx && ctx->nccl_comm, "no communicator (call lso_comm_init_rank)");
    LSO_CHECK_NCCL(ctx, g_nccl.AllGather(d_send, d_recv, (size_t)count, LSO_NCCL_FLOAT64, ctx->nccl_comm, st));
    return LSO_OK;
}

extern "C" {

int lso_comm_unique_id(void* id128) {
    if (!id128) return lso_set_error(nullptr, LSO_ERR_ARG, "id128 is NULL");
    LSO_TRY(nccl_load(nullptr));
    LSO_CHECK_NCCL(nullptr, g_nccl.GetUniqueId((lso_nccl_uid*)id128));
    return LSO_OK;
}

int lso_comm_init_rank(lso_ctx* ctx, int nranks, int rank, const void* id128) {
    LSO_REQUIRE(ctx, ctx && id128, "NULL pointer");
    LSO_REQUIRE(ctx, nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
    LSO_REQUIRE(ctx, ctx->nccl_comm == nullptr, "communicator already initialised");
    LSO_TRY(nccl_load(ctx));
    LSO_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
    lso_nccl_uid uid;
    memcpy(&uid, id128, sizeof(uid));
    lso_nccl_comm_t comm = nullptr;
    LSO_CHECK_NCCL(ctx, g_nccl.CommInitRank(&comm, nranks, uid, rank));
    ctx->nccl_comm = comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    return LSO_OK;
}

int lso_comm_destroy(lso_ctx* ctx) {
    if (!ctx || !ctx->nccl_comm) return LSO_OK;
    cudaStreamSynchronize(ctx->stream);
    g_nccl.CommDestroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->nranks = 1;
    ctx->rank = 0;
    return LSO_OK;
}

int lso_comm_allreduce_sum(lso_ctx* ctx, double* d_buf, int64_t count) {
    LSO_REQUIRE(ctx, ctx && d_buf, "NULL pointer");
    if (ctx->nranks <= 1) return LSO_OK;
    LSO_REQUIRE(ctx, ctx->nccl_comm, "no communicator (call lso_comm_init_rank)");
    LSO_ENTER(ctx);
    LSO_CHECK_NCCL(ctx, g_nccl.AllReduce(d_buf, d_buf, (size_t)count, LSO_NCCL_FLOAT64, LSO_NCCL_SUM, ctx->nccl_comm,
                                         ctx->stream));
    return LSO_OK;
}

}  // extern "C"
