// fp64_peak.cu - measured FP64 peaks of this GPU: the denominators of the dense-solve / EKF rooflines (BASELINE.md section 3 asks for a
// measured DFMA / DMMA figure because MEASURED_PEAKS.json has none).
//   DFMA: 8 independent FMA chains per thread, 1024 threads per CTA, 2 CTAs per SM   -> vector FP64 pipe
//   DMMA: mma.sync.aligned.m8n8k4.row.col.f64 (the only FP64 tensor shape PTX exposes), 4 independent accumulator tiles per warp
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak scripts/fp64_peak.cu ; run on the GPU box; prints one JSON line.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024) k_dfma(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(1024) k_dmma(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) { c[t][0] = threadIdx.x + t; c[t][1] = t; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int t = 0; t < 8; ++t) dmma(c[t][0], c[t][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int ctas = p.multiProcessorCount * 2, threads = 1024;
    double* out;
    cudaMalloc(&out, sizeof(double) * ctas * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best_f = 1e30f, best_m = 1e30f;
    const int it_f = 4096, it_m = 2048;
    for (int rep = 0; rep < 12; ++rep) {
        cudaEventRecord(e0);
        k_dfma<<<ctas, threads>>>(out, it_f, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best_f) best_f = ms;
        cudaEventRecord(e0);
        k_dmma<<<ctas, threads>>>(out, it_m, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best_m) best_m = ms;
    }
    const double flop_f = 2.0 * 64.0 * it_f * (double)ctas * threads;                 // 64 FMA per thread per iteration
    const double flop_m = 2.0 * 8 * 8 * 4 * 32.0 * it_m * (double)ctas * (threads / 32);   // 32 m8n8k4 MMAs per warp per iteration
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, \"dfma_ms\": %.3f, \"dmma_ms\": %.3f, "
           "\"how\": \"8 independent DFMA chains x 1024 threads x 2 CTAs/SM; 8 independent mma.sync.m8n8k4.f64 accumulator tiles per warp, 32 warps x 2 CTAs/SM; best of 10, CUDA events\"}\n",
           p.name, p.multiProcessorCount, flop_f / (best_f * 1e-3) / 1e12, flop_m / (best_m * 1e-3) / 1e12, best_f, best_m);
    return 0;
}
