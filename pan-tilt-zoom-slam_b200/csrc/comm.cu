// comm.cu - NCCL plumbing for the multi-GPU bundle-adjustment path (one process per GPU).
//
// BA observation blocks are sharded by keyframe across ranks; after the fused pass every rank holds
// complete keyframe blocks for its own keyframes and PARTIAL landmark blocks, so the packed accumulator arena
// [cost | U | V | g_c | g_l] is summed with one ncclAllReduce(FP64, sum) on the context stream.
// libnccl.so.2 is resolved at run time with dlopen: single-GPU users need no NCCL, and a process that already loaded
// torch's bundled NCCL shares that copy (same SONAME).
#include <dlfcn.h>

#include "ba.h"

namespace {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
typedef int (*fn_get_unique_id)(NcclUniqueId*);
typedef int (*fn_comm_init_rank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_comm_destroy)(NcclComm);
typedef const char* (*fn_error_string)(int);

struct NcclApi {
    void* lib = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_error_string error_string = nullptr;
};

NcclApi g_api;

int load_nccl(ptzba_ctx* ctx) {
    if (g_api.lib) return PTZBA_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return ptzba_fail(ctx, PTZBA_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
    g_api.get_unique_id = (fn_get_unique_id)dlsym(lib, "ncclGetUniqueId");
    g_api.comm_init_rank = (fn_comm_init_rank)dlsym(lib, "ncclCommInitRank");
    g_api.all_reduce = (fn_all_reduce)dlsym(lib, "ncclAllReduce");
    g_api.comm_destroy = (fn_comm_destroy)dlsym(lib, "ncclCommDestroy");
    g_api.error_string = (fn_error_string)dlsym(lib, "ncclGetErrorString");
    if (!g_api.get_unique_id || !g_api.comm_init_rank || !g_api.all_reduce || !g_api.comm_destroy)
        return ptzba_fail(ctx, PTZBA_ERR_COMM, "libnccl.so.2 lacks a required symbol");
    g_api.lib = lib;
    return PTZBA_OK;
}

int nccl_fail(ptzba_ctx* ctx, const char* what, int code) {
    return ptzba_fail(ctx, PTZBA_ERR_COMM, "%s failed: %s", what, g_api.error_string ? g_api.error_string(code) : "?");
}

const int kNcclFloat64 = 8, kNcclSum = 0;   // ncclDataType_t / ncclRedOp_t values of nccl.h

}  // namespace

void ptzba_comm_release(ptzba_ctx* ctx) {
    if (ctx->nccl_comm && g_api.comm_destroy) g_api.comm_destroy((NcclComm)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
}

extern "C" int ptzba_comm_unique_id(ptzba_ctx* ctx, void* unique_id128) {
    if (!ctx) return PTZBA_ERR_ARG;
    ARG_CHECK(ctx, unique_id128);
    PROPAGATE(load_nccl(ctx));
    NcclUniqueId id;
    const int r = g_api.get_unique_id(&id);
    if (r != 0) return nccl_fail(ctx, "ncclGetUniqueId", r);
    memcpy(unique_id128, &id, sizeof(id));
    return PTZBA_OK;
}

extern "C" int ptzba_comm_init(ptzba_ctx* ctx, const void* unique_id128, int rank, int world_size) {
    if (!ctx) return PTZBA_ERR_ARG;
    ARG_CHECK(ctx, unique_id128 && world_size >= 1 && rank >= 0 && rank < world_size);
    PROPAGATE(load_nccl(ctx));
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    ptzba_comm_release(ctx);
    NcclUniqueId id;
    memcpy(&id, unique_id128, sizeof(id));
    NcclComm comm = nullptr;
    const int r = g_api.comm_init_rank(&comm, world_size, id, rank);
    if (r != 0) return nccl_fail(ctx, "ncclCommInitRank", r);
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->world = world_size;
    return PTZBA_OK;
}

extern "C" int ptzba_comm_allreduce_f64(ptzba_ctx* ctx, double* device_buf, int64_t count) {
    if (!ctx) return PTZBA_ERR_ARG;
    ARG_CHECK(ctx, device_buf && count >= 0);
    if (ctx->world <= 1 || count == 0) return PTZBA_OK;
    if (!ctx->nccl_comm) return ptzba_fail(ctx, PTZBA_ERR_STATE, "ptzba_comm_init has not been called");
    const int r = g_api.all_reduce(device_buf, device_buf, (size_t)count, kNcclFloat64, kNcclSum, (NcclComm)ctx->nccl_comm,
                                   ctx->stream);
    if (r != 0) return nccl_fail(ctx, "ncclAllReduce", r);
    return PTZBA_OK;
}

// sums the accumulators of the last fused pass ([cost | U | V | g_c | g_l]) over all ranks, in place, on the stream
extern "C" int ptzba_ba_allreduce(ptzba_ba* ba) {
    if (!ba) return PTZBA_ERR_ARG;
    return ptzba_comm_allreduce_f64(ba->ctx, ba->acc.base, (int64_t)ba->acc.count);
}
