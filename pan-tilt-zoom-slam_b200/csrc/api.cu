// api.cu - context life cycle and error plumbing of libptzba.so (see include/ptzba.h).
#include <stdarg.h>

#include "common.h"

int ptzba_fail(ptzba_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

extern "C" int ptzba_version(void) { return PTZBA_VERSION; }

extern "C" int ptzba_create(int device, ptzba_ctx** out) {
    if (!out) return PTZBA_ERR_ARG;
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return PTZBA_ERR_CUDA;   // no CPU fallback
    if (device < 0 || device >= n_dev) return PTZBA_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return PTZBA_ERR_CUDA;
    ptzba_ctx* ctx = new ptzba_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return PTZBA_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PTZBA_ERR_CUDA; }
    ctx->stream = ctx->own_stream;
    if (cudaMallocHost((void**)&ctx->h_scalars, 256 * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_scalars, 256 * sizeof(double)) != cudaSuccess) {
        ptzba_destroy(ctx);
        return PTZBA_ERR_CUDA;
    }
    *out = ctx;
    return PTZBA_OK;
}

void ptzba_comm_release(ptzba_ctx* ctx);

extern "C" void ptzba_destroy(ptzba_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    ptzba_comm_release(ctx);
    if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
    if (ctx->d_scalars) cudaFree(ctx->d_scalars);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

extern "C" const char* ptzba_last_error(ptzba_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int ptzba_set_stream(ptzba_ctx* ctx, void* cuda_stream) {
    if (!ctx) return PTZBA_ERR_ARG;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return PTZBA_OK;
}

extern "C" int ptzba_synchronize(ptzba_ctx* ctx) {
    if (!ctx) return PTZBA_ERR_ARG;
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return PTZBA_OK;
}

extern "C" int64_t ptzba_launch_count(ptzba_ctx* ctx) { return ctx ? ctx->launches : -1; }

extern "C" int ptzba_profile_begin(ptzba_ctx* ctx) {
    if (!ctx) return PTZBA_ERR_ARG;
    for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
    ctx->prof_events.clear();
    ctx->profiling = true;
    return PTZBA_OK;
}

extern "C" int ptzba_profile_end(ptzba_ctx* ctx, int32_t* n_launches, double* total_ms) {
    if (!ctx) return PTZBA_ERR_ARG;
    ctx->profiling = false;
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    double tot = 0;
    for (size_t i = 0; i + 1 < ctx->prof_events.size(); i += 2) {
        float ms = 0;
        CU_CHECK(ctx, cudaEventElapsedTime(&ms, ctx->prof_events[i], ctx->prof_events[i + 1]));
        tot += ms;
    }
    if (n_launches) *n_launches = (int32_t)(ctx->prof_events.size() / 2);
    if (total_ms) *total_ms = tot;
    for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
    ctx->prof_events.clear();
    return PTZBA_OK;
}
