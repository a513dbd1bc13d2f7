// common.h - context object, error plumbing and small device-buffer helpers shared by all translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/ptzba.h"

struct ptzba_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;     // the stream all kernels are enqueued on
    int sm_count = 148;
    int64_t launches = 0;
    std::string err;
    // small pinned scratch for scalar read-backs
    double* h_scalars = nullptr;       // pinned, 256 doubles
    double* d_scalars = nullptr;       // device, 256 doubles
    // NCCL (loaded lazily with dlopen so that single-GPU use has no NCCL dependency)
    void* nccl_lib = nullptr;
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
    // optional per-launch event timing of the dominant kernel
    bool coop_configured = false;      // dynamic shared-memory opt-in of the cooperative Cholesky kernel done on this device
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;   // start/stop pairs
};

int ptzba_fail(ptzba_ctx* ctx, int code, const char* fmt, ...);

#define CU_CHECK(ctx, expr)                                                                        \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return ptzba_fail((ctx), PTZBA_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,  \
                              cudaGetErrorString(e__));                                            \
    } while (0)

#define ARG_CHECK(ctx, cond)                                                                       \
    do {                                                                                           \
        if (!(cond)) return ptzba_fail((ctx), PTZBA_ERR_ARG, "%s:%d argument check failed: %s",    \
                                       __FILE__, __LINE__, #cond);                                 \
    } while (0)

#define PROPAGATE(expr)                \
    do {                               \
        int s__ = (expr);              \
        if (s__ != PTZBA_OK) return s__; \
    } while (0)

// counts a kernel launch and checks the launch error
#define KERNEL_POST(ctx)                                                                           \
    do {                                                                                           \
        (ctx)->launches++;                                                                         \
        cudaError_t e__ = cudaGetLastError();                                                      \
        if (e__ != cudaSuccess)                                                                    \
            return ptzba_fail((ctx), PTZBA_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__,        \
                              __LINE__, cudaGetErrorString(e__));                                  \
    } while (0)

// RAII device buffer (freed with the owning object)
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
};

// Staging of a caller array (host or device) as a device pointer.
// mem == PTZBA_DEVICE: uses the caller pointer directly.  mem == PTZBA_HOST: copies into an owned device buffer.
template <typename T>
struct InArray {
    DevBuf<T> own;
    const T* d = nullptr;
    cudaError_t stage(int mem, const T* src, size_t count, cudaStream_t s) {
        if (mem == PTZBA_DEVICE || src == nullptr) {
            d = src;
            return cudaSuccess;
        }
        cudaError_t e = own.alloc(count);
        if (e != cudaSuccess) return e;
        d = own.p;
        if (count == 0) return cudaSuccess;
        return cudaMemcpyAsync(own.p, src, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
};

template <typename T>
struct OutArray {
    DevBuf<T> own;
    T* d = nullptr;
    T* host = nullptr;
    size_t count = 0;
    cudaError_t stage(int mem, T* dst, size_t cnt) {
        count = cnt;
        if (mem == PTZBA_DEVICE || dst == nullptr) {
            d = dst;
            host = nullptr;
            return cudaSuccess;
        }
        host = dst;
        cudaError_t e = own.alloc(cnt);
        d = own.p;
        return e;
    }
    cudaError_t finish(cudaStream_t s, size_t cnt_override = (size_t)-1) {
        if (!host) return cudaSuccess;
        size_t c = cnt_override == (size_t)-1 ? count : cnt_override;
        if (c == 0) return cudaSuccess;
        return cudaMemcpyAsync(host, d, c * sizeof(T), cudaMemcpyDeviceToHost, s);
    }
};

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
