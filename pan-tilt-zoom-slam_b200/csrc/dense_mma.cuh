// dense_mma.cuh - FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) tile product shared by the batched dense kernels of the EKF
// update (innovation-covariance Cholesky, Z = L^-1 [G | y], covariance down-date Z^T Z; ptz_slam.py:256-289).
//
// One CTA (256 threads = 8 warps) accumulates   acc += A[TM x K] * B[K x TN]   for one TM x TN output tile:
//   A(i, t) = A0[i + t * sAt]   (i contiguous in memory, i < mA else 0)
//   B(t, j) = B0[j + t * sBt]   (j contiguous in memory, j < nB else 0)        [B_TCONTIG: B0[t + j * sBt], t contiguous]
// K runs over [0, K) in slabs of 32 that are staged through shared memory: the next slab is requested into registers
// before the current one is multiplied (one shared-memory buffer, two __syncthreads per slab).  (Slabs of 16 with two
// shared-memory buffers were measured first: with few CTAs per SM the K loop is bound by the load latency per slab, so
// fewer, larger slabs win.)  Each warp owns (TM/8/WM) x (TN/8/WN) m8n8 accumulator tiles; per k-step of 4 it loads one A
// fragment per tile row and one B fragment per tile column (conflict-free: the slab row stride is 4 mod 16 doubles) and
// issues one DMMA per tile.  Measured DMMA peak on this B200: 37.1 TFLOP/s (profiles/r2_fp64_peak.json; DFMA: 34.1).
// FP64 has no tcgen05 kind: mma.sync (SASS DMMA.8x8x4) IS the FP64 tensor path of sm_100a.
#pragma once
#include <cuda_runtime.h>

namespace dmma {

constexpr int kThreads = 256;
constexpr int KS = 32;                      // K slab staged per step

__device__ __forceinline__ void mma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int TM, int TN>
struct Tile {
    static constexpr int MB = TM / 8, NBK = TN / 8;
    static constexpr int WM = (TM >= 64) ? 4 : 2, WN = 8 / WM;          // warp grid
    static constexpr int RM = MB / WM, RN = NBK / WN;                    // m8n8 tiles per warp
    static constexpr int SA = TM + 4, SB = TN + 4;                       // slab row strides (doubles): 4 mod 16
    static constexpr int kSmemDoubles = KS * (SA + SB);
    static constexpr int LA = KS * TM / kThreads, LB = KS * TN / kThreads;   // doubles per thread and slab
    static_assert(RM >= 1 && RN >= 1 && LA >= 1 && LB >= 1, "tile too small for 8 warps");

    // acc[rm][rn][2]: element (row g, cols 2*t4, 2*t4+1) of tile (rm, rn); g = lane / 4, t4 = lane % 4
    template <bool B_TCONTIG = false>
    __device__ static __forceinline__ void accumulate(const double* __restrict__ A0, size_t sAt, int mA, const double* __restrict__ B0,
                                                      size_t sBt, int nB, int K, double (&acc)[RM][RN][2], double* __restrict__ sm) {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int wm = warp % WM, wn = warp / WM;
        const int g = lane >> 2, t4 = lane & 3;
        double* As = sm;                          // [KS][SA]
        double* Bs = sm + KS * SA;                // [KS][SB]
        double ra[LA], rb[LB];
        auto gload = [&](int k0) {
#pragma unroll
            for (int q = 0; q < LA; ++q) {
                const int e = tid + kThreads * q, i = e % TM, t = e / TM;
                ra[q] = (i < mA && k0 + t < K) ? A0[(size_t)i + (size_t)(k0 + t) * sAt] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < LB; ++q) {
                const int e = tid + kThreads * q;
                if (B_TCONTIG) {
                    const int t = e % KS, j = e / KS;
                    rb[q] = (j < nB && k0 + t < K) ? B0[(size_t)(k0 + t) + (size_t)j * sBt] : 0.0;
                } else {
                    const int j = e % TN, t = e / TN;
                    rb[q] = (j < nB && k0 + t < K) ? B0[(size_t)j + (size_t)(k0 + t) * sBt] : 0.0;
                }
            }
        };
        auto sstore = [&]() {
#pragma unroll
            for (int q = 0; q < LA; ++q) {
                const int e = tid + kThreads * q, i = e % TM, t = e / TM;
                As[t * SA + i] = ra[q];
            }
#pragma unroll
            for (int q = 0; q < LB; ++q) {
                const int e = tid + kThreads * q;
                const int j = B_TCONTIG ? e / KS : e % TN, t = B_TCONTIG ? e % KS : e / TN;
                Bs[t * SB + j] = rb[q];
            }
        };
        if (K <= 0) return;
        gload(0);
        const int nslab = (K + KS - 1) / KS;
        for (int s = 0; s < nslab; ++s) {
            __syncthreads();                      // the previous slab has been consumed
            sstore();
            __syncthreads();
            if (s + 1 < nslab) gload((s + 1) * KS);
#pragma unroll
            for (int ks = 0; ks < KS / 4; ++ks) {
                double fa[RM], fb[RN];
#pragma unroll
                for (int rm = 0; rm < RM; ++rm) fa[rm] = As[(ks * 4 + t4) * SA + (wm * RM + rm) * 8 + g];
#pragma unroll
                for (int rn = 0; rn < RN; ++rn) fb[rn] = Bs[(ks * 4 + t4) * SB + (wn * RN + rn) * 8 + g];
#pragma unroll
                for (int rm = 0; rm < RM; ++rm)
#pragma unroll
                    for (int rn = 0; rn < RN; ++rn) mma884(acc[rm][rn][0], acc[rm][rn][1], fa[rm], fb[rn]);
            }
        }
        __syncthreads();                          // callers reuse the shared-memory buffer
    }

    // visits every accumulator element: f(row in tile, column in tile, value)
    template <typename F>
    __device__ static __forceinline__ void for_each(const double (&acc)[RM][RN][2], F f) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int wm = warp % WM, wn = warp / WM;
        const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
        for (int rm = 0; rm < RM; ++rm)
#pragma unroll
            for (int rn = 0; rn < RN; ++rn) {
                const int r = (wm * RM + rm) * 8 + g, c = (wn * RN + rn) * 8 + 2 * t4;
                f(r, c, acc[rm][rn][0]);
                f(r, c + 1, acc[rm][rn][1]);
            }
    }
};

}  // namespace dmma
