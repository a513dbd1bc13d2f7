// dense.cu - hand-written dense FP64 factorisation kernels for the Schur-reduced camera system S (order 3(N-1)).
//
// The reference hands the whole problem to scipy's dense SVD (bundle_adjustment.py:200-202 -> _lsq/trf.py); here the
// reduced system is factorised with a blocked right-looking Cholesky (lower, column-major, in place):
//     for each 32-wide panel:  potf2 (one warp, registers + shuffles)  ->  trsm (row per thread, L_kk in smem)
//                              ->  syrk (64x64 tiles, FP64 FMA, lower tiles only)
// followed by blocked forward / backward substitutions for up to two right-hand sides at once.
// Bound by the FP64 pipe for large n (n^3/3 flops) and by launch latency for small n.
#include "common.h"
#include "dense.h"

namespace {

constexpr int NB = 32;        // panel width
constexpr int TS = 64;        // syrk tile

// ---- potf2: Cholesky of one NB x NB diagonal block by a single warp; lane = row -----------------------------------
__global__ void __launch_bounds__(32) k_potf2(double* __restrict__ A, int lda, int k, int nb, int* __restrict__ info) {
    const int lane = threadIdx.x;
    double row[NB];
    double* blk = A + (size_t)k + (size_t)k * lda;
#pragma unroll
    for (int j = 0; j < NB; ++j) row[j] = (lane < nb && j < nb && j <= lane) ? blk[lane + (size_t)j * lda] : 0.0;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        if (j < nb) {
            // pivot lives in lane j
            double piv = __shfl_sync(0xffffffffu, row[j], j);
            if (!(piv > 0.0)) { bad = true; piv = 1.0; }
            const double s = sqrt(piv);
            const double inv = 1.0 / s;
            if (lane == j) row[j] = s; else if (lane > j) row[j] *= inv;
            // trailing update: a_ik -= l_ij * l_kj for k in (j, i]
            const double lij = row[j];
#pragma unroll
            for (int c = j + 1; c < NB; ++c) {
                const double lcj = __shfl_sync(0xffffffffu, row[j], c);
                if (c < nb && lane >= c) row[c] = fma(-lij, lcj, row[c]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j)
        if (lane < nb && j < nb && j <= lane) blk[lane + (size_t)j * lda] = row[j];
    if (bad && lane == 0) atomicMax(info, k + 1);
}

// ---- trsm: rows below the diagonal block, X * L_kk^T = A_panel; one row per thread ------------------------------------
__global__ void __launch_bounds__(128) k_trsm_panel(double* __restrict__ A, int lda, int n, int k, int nb) {
    __shared__ double L[NB][NB + 1];
    for (int e = threadIdx.x; e < NB * NB; e += 128) {
        const int i = e % NB, j = e / NB;
        L[i][j] = (i < nb && j < nb && j <= i) ? A[(size_t)(k + i) + (size_t)(k + j) * lda] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    const int r = k + nb + blockIdx.x * 128 + threadIdx.x;
    if (r >= n) return;
    double x[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) x[j] = j < nb ? A[(size_t)r + (size_t)(k + j) * lda] : 0.0;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        double s = x[j];
#pragma unroll
        for (int t = 0; t < j; ++t) s = fma(-x[t], L[j][t], s);
        x[j] = s / L[j][j];
    }
#pragma unroll
    for (int j = 0; j < NB; ++j)
        if (j < nb) A[(size_t)r + (size_t)(k + j) * lda] = x[j];
}

// ---- syrk: C -= P P^T on the lower tiles of the trailing matrix; 256 threads, 4x4 outputs each -----------------------
__global__ void __launch_bounds__(256) k_syrk_lower(double* __restrict__ A, int lda, int n, int k, int nb) {
    // tile pair (ti >= tj) from the linear block index
    const int t0 = k + nb;
    const int m = n - t0;
    const int nt = (m + TS - 1) / TS;
    int b = blockIdx.x, ti = 0;
    // rows of the lower-triangular tile grid have 1,2,3,... tiles
    ti = (int)((sqrt(8.0 * b + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= b) ++ti;
    while (ti * (ti + 1) / 2 > b) --ti;
    const int tj = b - ti * (ti + 1) / 2;
    if (ti >= nt) return;
    __shared__ double Pi[NB][TS + 1];
    __shared__ double Pj[NB][TS + 1];
    const int tid = threadIdx.x;
    const int r0 = t0 + ti * TS, c0 = t0 + tj * TS;
    for (int e = tid; e < NB * TS; e += 256) {
        const int rr = e % TS, t = e / TS;
        const int gi = r0 + rr, gj = c0 + rr;
        Pi[t][rr] = (gi < n && t < nb) ? A[(size_t)gi + (size_t)(k + t) * lda] : 0.0;
        Pj[t][rr] = (gj < n && t < nb) ? A[(size_t)gj + (size_t)(k + t) * lda] : 0.0;
    }
    __syncthreads();
    const int tx = tid % 16, ty = tid / 16;      // rows tx + 16*a, cols ty + 16*b  (coalesced along rows)
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.0;
#pragma unroll 8
    for (int t = 0; t < NB; ++t) {
        double pi[4], pj[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) { pi[a] = Pi[t][tx + 16 * a]; pj[a] = Pj[t][ty + 16 * a]; }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][c] = fma(pi[a], pj[c], acc[a][c]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int gi = r0 + tx + 16 * a, gj = c0 + ty + 16 * c;
            if (gi < n && gj < n && gi >= gj) A[(size_t)gi + (size_t)gj * lda] -= acc[a][c];
        }
}

// ---- triangular solves with NRHS right-hand sides (columns of B, ldb), single CTA, blocked by 32 -----------------
// forward: L y = b ; backward: L^T x = y.  1024 threads.
template <int NRHS>
__global__ void __launch_bounds__(1024) k_trsv_lower(const double* __restrict__ L, int lda, int n, double* __restrict__ B,
                                                     int ldb, int backward) {
    __shared__ double xs[NRHS][NB];
    __shared__ double D[NB][NB + 1];
    const int tid = threadIdx.x;
    const int nblk = (n + NB - 1) / NB;
    for (int bi = 0; bi < nblk; ++bi) {
        const int blk = backward ? (nblk - 1 - bi) : bi;
        const int k = blk * NB;
        const int nb = (n - k) < NB ? (n - k) : NB;
        // diagonal block into smem
        for (int e = tid; e < NB * NB; e += 1024) {
            const int i = e % NB, j = e / NB;
            D[i][j] = (i < nb && j < nb && j <= i) ? L[(size_t)(k + i) + (size_t)(k + j) * lda] : 0.0;
        }
        __syncthreads();
        if (tid < NRHS) {       // one thread per RHS solves the small triangle
            double x[NB];
            for (int i = 0; i < nb; ++i) x[i] = B[(size_t)(k + i) + (size_t)tid * ldb];
            if (!backward) {
                for (int i = 0; i < nb; ++i) {
                    double s = x[i];
                    for (int t = 0; t < i; ++t) s = fma(-D[i][t], x[t], s);
                    x[i] = s / D[i][i];
                }
            } else {
                for (int i = nb - 1; i >= 0; --i) {
                    double s = x[i];
                    for (int t = i + 1; t < nb; ++t) s = fma(-D[t][i], x[t], s);
                    x[i] = s / D[i][i];
                }
            }
            for (int i = 0; i < nb; ++i) {
                B[(size_t)(k + i) + (size_t)tid * ldb] = x[i];
                xs[tid][i] = x[i];
            }
            for (int i = nb; i < NB; ++i) xs[tid][i] = 0.0;
        }
        __syncthreads();
        // update the remaining entries with the off-diagonal panel
        if (!backward) {
            // b[r] -= sum_j L[r, k+j] x[j]  for r >= k+nb   (thread per row, coalesced over r)
            for (int r = k + nb + tid; r < n; r += 1024) {
                double s[NRHS];
#pragma unroll
                for (int q = 0; q < NRHS; ++q) s[q] = 0.0;
                for (int j = 0; j < nb; ++j) {
                    const double l = L[(size_t)r + (size_t)(k + j) * lda];
#pragma unroll
                    for (int q = 0; q < NRHS; ++q) s[q] = fma(l, xs[q][j], s[q]);
                }
#pragma unroll
                for (int q = 0; q < NRHS; ++q) B[(size_t)r + (size_t)q * ldb] -= s[q];
            }
        } else {
            // b[c] -= sum_i L[k+i, c] x[i]  for c < k   (warp per column c, lanes over i -> coalesced over i)
            const int warp = tid >> 5, lane = tid & 31;
            for (int c = warp; c < k; c += 32) {
                const double l = lane < nb ? L[(size_t)(k + lane) + (size_t)c * lda] : 0.0;
#pragma unroll
                for (int q = 0; q < NRHS; ++q) {
                    double s = l * xs[q][lane];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                    if (lane == 0) B[(size_t)c + (size_t)q * ldb] -= s;
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace

int dense_potrf_lower(ptzba_ctx* ctx, double* A, int n, int lda, int* d_info) {
    cudaStream_t s = ctx->stream;
    CU_CHECK(ctx, cudaMemsetAsync(d_info, 0, sizeof(int), s));
    for (int k = 0; k < n; k += NB) {
        const int nb = (n - k) < NB ? (n - k) : NB;
        k_potf2<<<1, 32, 0, s>>>(A, lda, k, nb, d_info);
        KERNEL_POST(ctx);
        const int m = n - k - nb;
        if (m <= 0) break;
        k_trsm_panel<<<div_up(m, 128), 128, 0, s>>>(A, lda, n, k, nb);
        KERNEL_POST(ctx);
        const int nt = div_up(m, TS);
        k_syrk_lower<<<nt * (nt + 1) / 2, 256, 0, s>>>(A, lda, n, k, nb);
        KERNEL_POST(ctx);
    }
    return PTZBA_OK;
}

int dense_potrs_lower(ptzba_ctx* ctx, const double* L, int n, int lda, double* B, int ldb, int nrhs) {
    cudaStream_t s = ctx->stream;
    if (n <= 0) return PTZBA_OK;
    if (nrhs == 1) {
        k_trsv_lower<1><<<1, 1024, 0, s>>>(L, lda, n, B, ldb, 0);
        KERNEL_POST(ctx);
        k_trsv_lower<1><<<1, 1024, 0, s>>>(L, lda, n, B, ldb, 1);
        KERNEL_POST(ctx);
    } else if (nrhs == 2) {
        k_trsv_lower<2><<<1, 1024, 0, s>>>(L, lda, n, B, ldb, 0);
        KERNEL_POST(ctx);
        k_trsv_lower<2><<<1, 1024, 0, s>>>(L, lda, n, B, ldb, 1);
        KERNEL_POST(ctx);
    } else {
        return ptzba_fail(ctx, PTZBA_ERR_ARG, "dense_potrs_lower: nrhs must be 1 or 2");
    }
    return PTZBA_OK;
}
