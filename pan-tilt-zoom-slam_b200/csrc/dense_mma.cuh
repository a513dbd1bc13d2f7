// dense_mma.cuh - FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) tile product shared by the batched dense kernels of the EKF
// update (innovation-covariance Cholesky / LU, Z = L^-1 [G | y], covariance down-date Z^T Z; ptz_slam.py:256-289).
//
// One CTA (256 threads = 8 warps) accumulates   acc += A[TM x K] * B[K x TN]   for one TM x TN output tile:
//   A(i, t) = A0[i + t * sAt]   (i contiguous in memory, i < mA else 0)
//   B(t, j) = B0[j + t * sBt]   (j contiguous in memory, j < nB else 0)        [B_TCONTIG: B0[t + j * sBt], t contiguous]
// K runs over [0, K) in slabs of 16 that travel through a FOUR-stage cp.async ring in shared memory (8-byte copies, zero
// fill outside the matrix): three slabs are in flight while one is multiplied, one __syncthreads per slab.  The batched
// kernels that use this core have short grids (a block row or block column of every matrix of the batch per launch), so the
// K loop of a CTA is a latency chain: register-staged single / double buffering left the FP64 tensor pipe idle between slabs
// (measured: ~50 us per launch at 64 sequences whatever the tile shape).
// Each warp owns (TM/8/WM) x (TN/8/WN) m8n8 accumulator tiles; per k-step of 4 it loads one A fragment per tile row and one B
// fragment per tile column (conflict-free: the slab row stride is 4 mod 16 doubles) and issues one DMMA per tile.
// Measured DMMA peak on this B200: 37.1 TFLOP/s (profiles/r2_fp64_peak.json; DFMA: 34.1).  FP64 has no tcgen05 kind:
// mma.sync (SASS DMMA.8x8x4) IS the FP64 tensor path of sm_100a.
#pragma once
#include <cuda_runtime.h>

namespace dmma {

constexpr int kThreads = 256;
constexpr int KS = 16;                      // K slab staged per step
constexpr int STAGES = 4;                   // cp.async ring depth

__device__ __forceinline__ void cp_async8(unsigned dst, const double* src, bool valid) {
    const unsigned sz = valid ? 8u : 0u;    // src-size 0: nothing is read, the 8 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int TM, int TN>
struct Tile {
    static constexpr int MB = TM / 8, NBK = TN / 8;
    static constexpr int WM = (TM >= 64) ? 4 : 2, WN = 8 / WM;          // warp grid
    static constexpr int RM = MB / WM, RN = NBK / WN;                    // m8n8 tiles per warp
    static constexpr int SA = TM + 4, SB = TN + 4;                       // slab row strides (doubles): 4 mod 16
    static constexpr int kSmemDoubles = STAGES * KS * (SA + SB);
    static constexpr int kSmemBytes = kSmemDoubles * 8;
    static constexpr int LA = KS * TM / kThreads, LB = KS * TN / kThreads;   // doubles per thread and slab
    static_assert(RM >= 1 && RN >= 1 && LA >= 1 && LB >= 1, "tile too small for 8 warps");

    // acc[rm][rn][2]: element (row g, cols 2*t4, 2*t4+1) of tile (rm, rn); g = lane / 4, t4 = lane % 4
    template <bool B_TCONTIG = false>
    __device__ static __forceinline__ void accumulate(const double* __restrict__ A0, size_t sAt, int mA, const double* __restrict__ B0,
                                                      size_t sBt, int nB, int K, double (&acc)[RM][RN][2], double* __restrict__ sm) {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int wm = warp % WM, wn = warp / WM;
        const int g = lane >> 2, t4 = lane & 3;
        double* As = sm;                          // [STAGES][KS][SA]
        double* Bs = sm + STAGES * KS * SA;       // [STAGES][KS][SB]
        const unsigned as_u = (unsigned)__cvta_generic_to_shared(As), bs_u = (unsigned)__cvta_generic_to_shared(Bs);
        auto issue = [&](int slab) {              // slab `slab` of K -> ring stage slab % STAGES
            const int k0 = slab * KS, st = slab % STAGES;
#pragma unroll
            for (int q = 0; q < LA; ++q) {
                const int e = tid + kThreads * q, i = e % TM, t = e / TM;
                const bool ok = i < mA && k0 + t < K;
                cp_async8(as_u + ((st * KS + t) * SA + i) * 8, ok ? A0 + (size_t)i + (size_t)(k0 + t) * sAt : A0, ok);
            }
#pragma unroll
            for (int q = 0; q < LB; ++q) {
                const int e = tid + kThreads * q;
                const int j = B_TCONTIG ? e / KS : e % TN, t = B_TCONTIG ? e % KS : e / TN;
                const bool ok = j < nB && k0 + t < K;
                const double* src = B_TCONTIG ? B0 + (size_t)(k0 + t) + (size_t)j * sBt : B0 + (size_t)j + (size_t)(k0 + t) * sBt;
                cp_async8(bs_u + ((st * KS + t) * SB + j) * 8, ok ? src : B0, ok);
            }
        };
        if (K <= 0) return;
        const int nslab = (K + KS - 1) / KS;
#pragma unroll
        for (int p = 0; p < STAGES - 1; ++p) {
            if (p < nslab) issue(p);
            cp_async_commit();
        }
        for (int s = 0; s < nslab; ++s) {
            cp_async_wait<STAGES - 2>();          // this thread's copies of slab s have landed ...
            __syncthreads();                      // ... everybody's have, and everybody is done with the stage refilled next
            if (s + STAGES - 1 < nslab) issue(s + STAGES - 1);
            cp_async_commit();
            const int st = s % STAGES;
#pragma unroll
            for (int ks = 0; ks < KS / 4; ++ks) {
                double fa[RM], fb[RN];
#pragma unroll
                for (int rm = 0; rm < RM; ++rm) fa[rm] = As[((st * KS) + ks * 4 + t4) * SA + (wm * RM + rm) * 8 + g];
#pragma unroll
                for (int rn = 0; rn < RN; ++rn) fb[rn] = Bs[((st * KS) + ks * 4 + t4) * SB + (wn * RN + rn) * 8 + g];
#pragma unroll
                for (int rm = 0; rm < RM; ++rm)
#pragma unroll
                    for (int rn = 0; rn < RN; ++rn) mma884(acc[rm][rn][0], acc[rm][rn][1], fa[rm], fb[rn]);
            }
        }
        cp_async_wait<0>();
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

// dynamic shared memory of a kernel built on a tile: above the 48 KB default the kernel has to opt in (once per process)
template <typename Kernel>
inline cudaError_t configure(Kernel k, int bytes) {
    return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

}  // namespace dmma
