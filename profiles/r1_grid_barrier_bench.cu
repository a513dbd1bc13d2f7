// micro-benchmark: cost of a device-wide barrier (own implementation vs cooperative groups) on 148 CTAs x 256 threads
#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&bar[0], 1u) == gridDim.x - 1) { atomicExch(&bar[0], 0u); __threadfence(); atomicAdd(&bar[1], 1u); }
        else { while (*reinterpret_cast<volatile unsigned*>(&bar[1]) == gen) {} }
        __threadfence();
    }
    ++gen;
    __syncthreads();
}
// release/acquire flavour: red.release + ld.acquire, no full fences
__device__ __forceinline__ void grid_barrier2(unsigned* bar, unsigned& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned old;
        asm volatile("atom.add.acq_rel.gpu.u32 %0, [%1], 1;" : "=r"(old) : "l"(bar) : "memory");
        if (old == gridDim.x - 1) {
            asm volatile("st.relaxed.gpu.u32 [%0], 0;" ::"l"(bar) : "memory");
            asm volatile("red.release.gpu.add.u32 [%0], 1;" ::"l"(bar + 1) : "memory");
        } else {
            unsigned v;
            do { asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(bar + 1) : "memory"); } while (v == gen);
        }
    }
    ++gen;
    __syncthreads();
}
__global__ void k_own(unsigned* bar, int iters, double* sink) {
    __shared__ unsigned sg; if (threadIdx.x == 0) sg = *(volatile unsigned*)&bar[1]; __syncthreads(); unsigned gen = sg;
    double a = threadIdx.x;
    for (int i = 0; i < iters; ++i) { a = a * 1.0000001 + 1.0; grid_barrier(bar, gen); }
    if (a == 12345.678) sink[0] = a;
}
__global__ void k_own2(unsigned* bar, int iters, double* sink) {
    __shared__ unsigned sg; if (threadIdx.x == 0) sg = *(volatile unsigned*)&bar[1]; __syncthreads(); unsigned gen = sg;
    double a = threadIdx.x;
    for (int i = 0; i < iters; ++i) { a = a * 1.0000001 + 1.0; grid_barrier2(bar, gen); }
    if (a == 12345.678) sink[0] = a;
}
__global__ void k_cg(int iters, double* sink) {
    cg::grid_group g = cg::this_grid();
    double a = threadIdx.x;
    for (int i = 0; i < iters; ++i) { a = a * 1.0000001 + 1.0; g.sync(); }
    if (a == 12345.678) sink[0] = a;
}
int main() {
    unsigned* bar; double* sink; cudaMalloc(&bar, 16); cudaMemset(bar, 0, 16); cudaMalloc(&sink, 8);
    int iters = 200; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    for (int grid : {148, 74, 32}) {
        for (int rep = 0; rep < 2; ++rep) {
            void* a1[] = {&bar, &iters, &sink};
            cudaEventRecord(e0); cudaLaunchCooperativeKernel((void*)k_own, dim3(grid), dim3(256), a1, 0, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1); printf("grid %3d own  : %.2f us/barrier (%s)\n", grid, ms * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
            cudaEventRecord(e0); cudaLaunchCooperativeKernel((void*)k_own2, dim3(grid), dim3(256), a1, 0, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1); printf("grid %3d own2 : %.2f us/barrier (%s)\n", grid, ms * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
            void* a2[] = {&iters, &sink};
            cudaEventRecord(e0); cudaLaunchCooperativeKernel((void*)k_cg, dim3(grid), dim3(256), a2, 0, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1); printf("grid %3d cg   : %.2f us/barrier (%s)\n", grid, ms * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
