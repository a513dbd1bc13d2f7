// dense.cu - hand-written BATCHED dense FP64 Cholesky kernels (the innovation covariances S = H P H^T + R of many independent
// EKF sequences, ptz_slam.py:256-259; the reference inverts S with numpy's LU).
//
// Blocked right-looking Cholesky (lower, column-major, in place), one launch pair per 32-wide panel for the whole batch:
//     k_potf2_trsm  every CTA factors the 32 x 32 diagonal block redundantly in shared memory, then solves its 128 panel rows
//     k_syrk_lower  64 x 64 tiles of the trailing matrix, FP64 FMA, lower tiles only
// and the multi-right-hand-side forward solve Z = L^-1 G on row-major G (k_fwd_diag_rows / k_fwd_update_rows).
// The single reduced camera system of bundle adjustment uses the cooperative one-kernel path of dense_coop.cu instead.
#include "common.h"
#include "dense.h"

namespace {

constexpr int NB = 32;        // panel width
constexpr int TS = 64;        // syrk tile

// ---- fused panel step: every CTA factors the NB x NB diagonal block redundantly in shared memory (128 threads, rsqrt on the
// critical path) and then solves its own 128 rows of the panel; CTA 0 also writes the factored diagonal block back.
// Replaces k_potf2 + k_trsm_panel (one launch less per panel and no single-warp kernel on the critical path).
__global__ void __launch_bounds__(128) k_potf2_trsm(double* __restrict__ A0, int lda, size_t stride, const int* __restrict__ n_arr,
                                                    int n_fixed, int k, int* __restrict__ info, int* __restrict__ info_b) {
    const int n = n_arr ? n_arr[blockIdx.y] : n_fixed;
    if (k >= n) return;
    const int nb = (n - k) < NB ? (n - k) : NB;
    double* A = A0 + stride * blockIdx.y;
    __shared__ double L[NB][NB + 1];
    const int tid = threadIdx.x;
    for (int e = tid; e < NB * NB; e += 128) {
        const int i = e % NB, j = e / NB;
        L[i][j] = (i < nb && j < nb && j <= i) ? A[(size_t)(k + i) + (size_t)(k + j) * lda] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    bool bad = false;
    for (int j = 0; j < NB; ++j) {
        double piv = L[j][j];
        if (!(piv > 0.0)) { bad = true; piv = 1.0; }
        const double inv = rsqrt(piv);
        __syncthreads();
        if (tid < NB) {
            if (tid == j) L[j][j] = piv * inv;
            else if (tid > j) L[tid][j] *= inv;
        }
        __syncthreads();
        // trailing update of the lower triangle: (i, c) with i >= c > j
        for (int e = tid; e < NB * NB; e += 128) {
            const int i = e % NB, c = e / NB;
            if (c > j && i >= c) L[i][c] = fma(-L[i][j], L[c][j], L[i][c]);
        }
        __syncthreads();
    }
    if (bad && tid == 0 && blockIdx.x == 0) {
        atomicMax(info, k + 1);
        if (info_b) info_b[blockIdx.y] = k + 1;
    }
    if (blockIdx.x == 0)
        for (int e = tid; e < NB * NB; e += 128) {
            const int i = e % NB, j = e / NB;
            if (i < nb && j < nb && j <= i) A[(size_t)(k + i) + (size_t)(k + j) * lda] = L[i][j];
        }
    const int r = k + nb + blockIdx.x * 128 + tid;
    if (r >= n || nb < NB) return;
    double x[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) x[j] = A[(size_t)r + (size_t)(k + j) * lda];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        double sacc = x[j];
#pragma unroll
        for (int t = 0; t < j; ++t) sacc = fma(-x[t], L[j][t], sacc);
        x[j] = sacc / L[j][j];
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) A[(size_t)r + (size_t)(k + j) * lda] = x[j];
}

// ---- syrk: C -= P P^T on the lower tiles of the trailing matrix; 256 threads, 4x4 outputs each -----------------------
__global__ void __launch_bounds__(256) k_syrk_lower(double* __restrict__ A0, int lda, size_t stride,
                                                    const int* __restrict__ n_arr, int n_fixed, int k) {
    const int n = n_arr ? n_arr[blockIdx.y] : n_fixed;
    if (k + NB >= n) return;
    const int nb = NB;
    double* A = A0 + stride * blockIdx.y;
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

}  // namespace

int dense_potrf_lower_batched(ptzba_ctx* ctx, double* A, int lda, size_t stride, const int* d_n_arr, int n_max, int batch,
                              int* d_info, int* d_info_per_batch) {
    cudaStream_t s = ctx->stream;
    CU_CHECK(ctx, cudaMemsetAsync(d_info, 0, sizeof(int), s));
    for (int k = 0; k < n_max; k += NB) {
        const int m = n_max - k - NB;
        k_potf2_trsm<<<dim3(m > 0 ? div_up(m, 128) : 1, batch), 128, 0, s>>>(A, lda, stride, d_n_arr, n_max, k, d_info, d_info_per_batch);
        KERNEL_POST(ctx);
        if (m <= 0) break;
        const int nt = div_up(m, TS);
        k_syrk_lower<<<dim3(nt * (nt + 1) / 2, batch), 256, 0, s>>>(A, lda, stride, d_n_arr, n_max, k);
        KERNEL_POST(ctx);
    }
    return PTZBA_OK;
}

namespace {

// block row k of  L Z = G  (L column-major lower, G row-major [n x ncols], ld = ldg): solve the NB x NB triangle for every
// column; one thread per column (coalesced along columns).
__global__ void __launch_bounds__(128) k_fwd_diag_rows(const double* __restrict__ L0, int lda, size_t strideL,
                                                       double* __restrict__ G0, int ldg, size_t strideG,
                                                       const int* __restrict__ n_arr, int extra_cols, int k) {
    const int n = n_arr[blockIdx.y];
    if (k >= n) return;
    const int nb = (n - k) < NB ? (n - k) : NB;
    const int ncols = n + extra_cols;
    const double* L = L0 + strideL * blockIdx.y;
    double* G = G0 + strideG * blockIdx.y;
    __shared__ double D[NB][NB + 1];
    for (int e = threadIdx.x; e < NB * NB; e += 128) {
        const int i = e % NB, j = e / NB;
        D[i][j] = (i < nb && j < nb && j <= i) ? L[(size_t)(k + i) + (size_t)(k + j) * lda] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= ncols) return;
    double x[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) x[i] = i < nb ? G[(size_t)(k + i) * ldg + c] : 0.0;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        double s = x[i];
#pragma unroll
        for (int t = 0; t < i; ++t) s = fma(-D[i][t], x[t], s);
        x[i] = s / D[i][i];
    }
#pragma unroll
    for (int i = 0; i < NB; ++i)
        if (i < nb) G[(size_t)(k + i) * ldg + c] = x[i];
}

// trailing rows: G[r, c] -= sum_t L[r, k+t] * G[k+t, c]   for r >= k+NB ; 64 x 64 tiles, 4x4 per thread
__global__ void __launch_bounds__(256) k_fwd_update_rows(const double* __restrict__ L0, int lda, size_t strideL,
                                                         double* __restrict__ G0, int ldg, size_t strideG,
                                                         const int* __restrict__ n_arr, int extra_cols, int k) {
    const int n = n_arr[blockIdx.z];
    const int r0 = k + NB + blockIdx.y * TS;
    if (r0 >= n) return;
    const int ncols = n + extra_cols;
    const int c0 = blockIdx.x * TS;
    if (c0 >= ncols) return;
    const double* L = L0 + strideL * blockIdx.z;
    double* G = G0 + strideG * blockIdx.z;
    __shared__ double Lp[NB][TS + 1];   // Lp[t][r]
    __shared__ double Zk[NB][TS + 1];   // Zk[t][c]
    const int tid = threadIdx.x;
    for (int e = tid; e < NB * TS; e += 256) {
        const int rr = e % TS, t = e / TS;
        Lp[t][rr] = (r0 + rr < n) ? L[(size_t)(r0 + rr) + (size_t)(k + t) * lda] : 0.0;
        Zk[t][rr] = (c0 + rr < ncols) ? G[(size_t)(k + t) * ldg + c0 + rr] : 0.0;
    }
    __syncthreads();
    const int tx = tid % 16, ty = tid / 16;   // cols tx + 16*b (coalesced), rows ty + 16*a
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 8
    for (int t = 0; t < NB; ++t) {
        double lr[4], zc[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) { lr[a] = Lp[t][ty + 16 * a]; zc[a] = Zk[t][tx + 16 * a]; }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fma(lr[a], zc[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int r = r0 + ty + 16 * a, c = c0 + tx + 16 * b;
            if (r < n && c < ncols) G[(size_t)r * ldg + c] -= acc[a][b];
        }
}

}  // namespace

// Z = L^-1 G in place for a batch; G row-major with n_b rows and n_b + extra_cols columns
int dense_fwd_solve_rows_batched(ptzba_ctx* ctx, const double* L, int lda, size_t strideL, double* G, int ldg, size_t strideG,
                                 const int* d_n_arr, int n_max, int extra_cols, int batch) {
    cudaStream_t s = ctx->stream;
    const int ncols_max = n_max + extra_cols;
    for (int k = 0; k < n_max; k += NB) {
        k_fwd_diag_rows<<<dim3(div_up(ncols_max, 128), batch), 128, 0, s>>>(L, lda, strideL, G, ldg, strideG, d_n_arr, extra_cols, k);
        KERNEL_POST(ctx);
        const int m = n_max - k - NB;
        if (m <= 0) break;
        k_fwd_update_rows<<<dim3(div_up(ncols_max, TS), div_up(m, TS), batch), 256, 0, s>>>(L, lda, strideL, G, ldg, strideG,
                                                                                            d_n_arr, extra_cols, k);
        KERNEL_POST(ctx);
    }
    return PTZBA_OK;
}
