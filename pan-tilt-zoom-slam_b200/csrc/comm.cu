// comm.cu - NCCL plumbing for the multi-GPU bundle-adjustment path (one process per GPU).
//
// BA observation blocks are sharded by keyframe across ranks; after the fused pass every rank holds
// complete keyframe blocks for its own keyframes and PARTIAL landmark blocks, so the packed accumulator arena
// [cost | U | V | g_c | g_l] is summed with one ncclAllReduce(FP64, sum) on the context stream.
// libnccl.so.2 is resolved at run time with dlopen: single-GPU users need no NCCL, and a process that already loaded
// torch's bundled NCCL shares that copy (same SONAME).
#include <dlfcn.h>

#include <vector>

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

namespace {

__global__ void k_touch_mask(int n_lm, const int32_t* __restrict__ lm_ptr, double* __restrict__ mask) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < n_lm) mask[l] = lm_ptr[l + 1] > lm_ptr[l] ? 1.0 : 0.0;
}

// shared[l] = 1 when more than one rank observes landmark l; compacted (in ascending id order, identical on every rank)
__global__ void k_shared_flags(int n_lm, const double* __restrict__ count, int32_t* __restrict__ flag) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < n_lm) flag[l] = count[l] > 1.5 ? 1 : 0;
}

// Only the shared landmarks THIS rank observes are packed (zeros otherwise) and unpacked: a landmark this rank never observes
// may lie outside the id range k_set_params clears, so writing the global sum into the arena would be re-added by the next
// pass's pack (stale blocks on every pass after the first with >= 3 ranks).
__global__ void k_pack_shared(int n_shared, const int32_t* __restrict__ ids, const int32_t* __restrict__ lm_ptr, const double* __restrict__ V,
                              const double* __restrict__ gl, const double* __restrict__ cost, double* __restrict__ buf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) buf[0] = cost[0];
    if (i >= n_shared) return;
    const int l = ids[i];
    double* b = buf + 1 + 5 * (size_t)i;
    if (lm_ptr[l + 1] > lm_ptr[l]) {
        b[0] = V[3 * (size_t)l]; b[1] = V[3 * (size_t)l + 1]; b[2] = V[3 * (size_t)l + 2];
        b[3] = gl[2 * (size_t)l]; b[4] = gl[2 * (size_t)l + 1];
    } else {
        b[0] = b[1] = b[2] = b[3] = b[4] = 0.0;
    }
}

__global__ void k_unpack_shared(int n_shared, const int32_t* __restrict__ ids, const int32_t* __restrict__ lm_ptr, const double* __restrict__ buf,
                                double* __restrict__ V, double* __restrict__ gl, double* __restrict__ cost) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) cost[0] = buf[0];
    if (i >= n_shared) return;
    const int l = ids[i];
    if (lm_ptr[l + 1] <= lm_ptr[l]) return;
    const double* b = buf + 1 + 5 * (size_t)i;
    V[3 * (size_t)l] = b[0]; V[3 * (size_t)l + 1] = b[1]; V[3 * (size_t)l + 2] = b[2];
    gl[2 * (size_t)l] = b[3]; gl[2 * (size_t)l + 1] = b[4];
}

}  // namespace

// Keyframe-sharded mode: finds the landmarks that are observed by more than one rank (one all-reduce of a per-landmark
// 0/1 mask, once per problem).  Afterwards ptzba_ba_allreduce exchanges ONLY the blocks of those shared landmarks (plus the
// cost): blocks of unshared landmarks are already complete on the only rank that sees them, and every keyframe's U / g_c is
// complete on the rank that owns the keyframe.  *n_shared_out (may be NULL) receives the number of shared landmarks.
extern "C" int ptzba_ba_setup_exchange(ptzba_ba* ba, int64_t* n_shared_out) {
    if (!ba) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = ba->ctx;
    cudaStream_t s = ctx->stream;
    const int M = ba->n_lm;
    ba->n_shared = 0;
    ba->exchange_ready = false;
    if (ctx->world > 1 && M > 0) {
        DevBuf<double> count;
        DevBuf<int32_t> flag;
        CU_CHECK(ctx, count.alloc(M));
        CU_CHECK(ctx, flag.alloc(M));
        k_touch_mask<<<div_up(M, 256), 256, 0, s>>>(M, ba->lm_ptr.p, count.p);
        KERNEL_POST(ctx);
        PROPAGATE(ptzba_comm_allreduce_f64(ctx, count.p, M));
        k_shared_flags<<<div_up(M, 256), 256, 0, s>>>(M, count.p, flag.p);
        KERNEL_POST(ctx);
        std::vector<int32_t> h(M), ids;
        CU_CHECK(ctx, cudaMemcpyAsync(h.data(), flag.p, (size_t)M * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CU_CHECK(ctx, cudaStreamSynchronize(s));
        for (int l = 0; l < M; ++l)
            if (h[l]) ids.push_back(l);
        ba->n_shared = (int64_t)ids.size();
        CU_CHECK(ctx, ba->shared_ids.alloc(ids.size() + 1));
        CU_CHECK(ctx, ba->shared_buf.alloc(1 + 5 * ids.size()));
        if (!ids.empty())
            CU_CHECK(ctx, cudaMemcpyAsync(ba->shared_ids.p, ids.data(), ids.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        CU_CHECK(ctx, cudaStreamSynchronize(s));
    }
    ba->exchange_ready = true;
    if (n_shared_out) *n_shared_out = ba->n_shared;
    return PTZBA_OK;
}

// Combines the accumulators of the last fused pass over all ranks, in place, on the stream.  Without
// ptzba_ba_setup_exchange: the whole packed arena [cost | U | V | g_c | g_l] is summed (every rank ends with every
// block).  With it: only the cost and the blocks of landmarks shared between ranks travel (see above).
extern "C" int ptzba_ba_allreduce(ptzba_ba* ba) {
    if (!ba) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = ba->ctx;
    if (!ba->exchange_ready) {
        ba->arena_foreign = ctx->world > 1;
        return ptzba_comm_allreduce_f64(ctx, ba->acc.base, (int64_t)ba->acc.count);
    }
    if (ctx->world <= 1) return PTZBA_OK;
    const int ns = (int)ba->n_shared;
    if (ns == 0) {
        // nothing is shared between the ranks: no block travels.  The cost stays a per-rank partial sum until somebody asks
        // for it (ptzba_ba_get_blocks, a collective call in this mode) - a pass over disjoint shards has no collective at all
        ba->cost_partial = true;
        return PTZBA_OK;
    }
    ba->cost_partial = false;
    cudaStream_t s = ctx->stream;
    k_pack_shared<<<div_up(ns > 0 ? ns : 1, 256), 256, 0, s>>>(ns, ba->shared_ids.p, ba->lm_ptr.p, ba->acc.V, ba->acc.gl, ba->acc.cost, ba->shared_buf.p);
    KERNEL_POST(ctx);
    PROPAGATE(ptzba_comm_allreduce_f64(ctx, ba->shared_buf.p, 1 + 5 * (int64_t)ns));
    k_unpack_shared<<<div_up(ns > 0 ? ns : 1, 256), 256, 0, s>>>(ns, ba->shared_ids.p, ba->lm_ptr.p, ba->shared_buf.p, ba->acc.V, ba->acc.gl, ba->acc.cost);
    KERNEL_POST(ctx);
    return PTZBA_OK;
}
