// dense_lu.cu - batched dense FP64 LU factorisation with partial pivoting and the matching multi-RHS solves.
//
// Why LU and not Cholesky: the reference inverts the EKF innovation covariance with numpy.linalg.inv (LAPACK getrf/getri,
// slam_system/ptz_slam.py:259).  Because its covariance write-back drops cross terms (:281-289), P - and with it
// S = H P H^T + R - loses positive definiteness after a few frames (measured: eigenvalues down to -0.19 on seeded
// sequences), so a Cholesky would break down exactly where the reference keeps going.  Partial pivoting reproduces
// getrf's behaviour on those matrices.
//
// Layout: A column-major (lda), matrix b of the batch at A + b*stride, order n_arr[b] (device array).  Right-looking,
// panel width 32:  panel (one CTA per matrix: pivot search, swap, scale, rank-1 inside the panel) -> laswp on the
// other columns -> U12 = L11^-1 A12 -> A22 -= L21 U12 (64x64 tiles).  Right-hand sides are ROW-major [n x ncols].
#include "common.h"
#include "dense.h"

namespace {

constexpr int NB = 32;
constexpr int TS = 64;
constexpr int PT = 512;     // threads of the panel kernel

__global__ void __launch_bounds__(PT) k_lu_panel(double* __restrict__ A0, int lda, size_t stride, const int* __restrict__ n_arr,
                                                 int k0, int* __restrict__ ipiv0, int ld_ipiv, int* __restrict__ info) {
    const int b = blockIdx.x;
    const int n = n_arr[b];
    if (k0 >= n) return;
    const int nb = (n - k0) < NB ? (n - k0) : NB;
    double* A = A0 + stride * b;
    int* ipiv = ipiv0 + (size_t)ld_ipiv * b;
    __shared__ double s_val[PT / 32];
    __shared__ int s_idx[PT / 32];
    __shared__ int s_piv;
    __shared__ double s_row[NB];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int jj = 0; jj < nb; ++jj) {
        const int j = k0 + jj;
        // (1) pivot search in column j, rows j..n-1 (first index of the maximum magnitude, like idamax)
        double best = -1.0;
        int bi = j;
        const double* col = A + (size_t)j * lda;
        for (int r = j + tid; r < n; r += PT) {
            const double a = fabs(col[r]);
            if (a > best) { best = a; bi = r; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) { s_val[w] = best; s_idx[w] = bi; }
        __syncthreads();
        if (w == 0) {
            best = lane < PT / 32 ? s_val[lane] : -1.0;
            bi = lane < PT / 32 ? s_idx[lane] : 0x7fffffff;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (lane == 0) {
                s_piv = bi;
                ipiv[j] = bi;
                if (!(best > 0.0)) atomicMax(info, j + 1);
            }
        }
        __syncthreads();
        const int p = s_piv;
        // (2) swap rows j and p inside the panel, publish the pivot row
        if (tid < nb) {
            double* cj = A + (size_t)(k0 + tid) * lda;
            const double vj = cj[j], vp = cj[p];
            if (p != j) { cj[j] = vp; cj[p] = vj; }
            s_row[tid] = vp;
        }
        __syncthreads();
        const double piv = s_row[jj];
        const double inv = piv != 0.0 ? 1.0 / piv : 0.0;
        // (3) scale the column and rank-1 update of the remaining panel columns
        for (int r = j + 1 + tid; r < n; r += PT) {
            const double l = A[(size_t)j * lda + r] * inv;
            A[(size_t)j * lda + r] = l;
            for (int c = jj + 1; c < nb; ++c) A[(size_t)(k0 + c) * lda + r] = fma(-l, s_row[c], A[(size_t)(k0 + c) * lda + r]);
        }
        __syncthreads();
    }
}

// apply the panel's row interchanges to the columns outside the panel
__global__ void __launch_bounds__(256) k_lu_laswp(double* __restrict__ A0, int lda, size_t stride, const int* __restrict__ n_arr,
                                                  int k0, const int* __restrict__ ipiv0, int ld_ipiv) {
    const int b = blockIdx.y;
    const int n = n_arr[b];
    if (k0 >= n) return;
    const int nb = (n - k0) < NB ? (n - k0) : NB;
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= n || (c >= k0 && c < k0 + nb)) return;
    double* col = A0 + stride * b + (size_t)c * lda;
    const int* ipiv = ipiv0 + (size_t)ld_ipiv * b;
    for (int jj = 0; jj < nb; ++jj) {
        const int j = k0 + jj, p = ipiv[j];
        if (p != j) { const double t = col[j]; col[j] = col[p]; col[p] = t; }
    }
}

// U12 = L11^-1 A12 (unit lower), one column per thread
__global__ void __launch_bounds__(128) k_lu_trsm_u12(double* __restrict__ A0, int lda, size_t stride,
                                                     const int* __restrict__ n_arr, int k0) {
    const int b = blockIdx.y;
    const int n = n_arr[b];
    if (k0 + NB >= n) return;
    double* A = A0 + stride * b;
    __shared__ double L[NB][NB + 1];
    for (int e = threadIdx.x; e < NB * NB; e += 128) {
        const int i = e % NB, j = e / NB;
        L[i][j] = (j < i) ? A[(size_t)(k0 + i) + (size_t)(k0 + j) * lda] : 0.0;
    }
    __syncthreads();
    const int c = k0 + NB + blockIdx.x * 128 + threadIdx.x;
    if (c >= n) return;
    double* col = A + (size_t)c * lda + k0;
    double x[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) x[i] = col[i];
#pragma unroll
    for (int i = 1; i < NB; ++i) {
        double s = x[i];
#pragma unroll
        for (int t = 0; t < i; ++t) s = fma(-L[i][t], x[t], s);
        x[i] = s;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) col[i] = x[i];
}

// A22 -= L21 U12 ; 64x64 tiles, 4x4 per thread
__global__ void __launch_bounds__(256) k_lu_gemm(double* __restrict__ A0, int lda, size_t stride, const int* __restrict__ n_arr,
                                                 int k0) {
    const int b = blockIdx.z;
    const int n = n_arr[b];
    const int t0 = k0 + NB;
    const int r0 = t0 + blockIdx.x * TS, c0 = t0 + blockIdx.y * TS;
    if (r0 >= n || c0 >= n) return;
    double* A = A0 + stride * b;
    __shared__ double Lp[NB][TS + 1];   // Lp[t][r] = A[r0+r][k0+t]
    __shared__ double Uk[NB][TS + 1];   // Uk[t][c] = A[k0+t][c0+c]
    const int tid = threadIdx.x;
    for (int e = tid; e < NB * TS; e += 256) {
        const int rr = e % TS, t = e / TS;
        Lp[t][rr] = (r0 + rr < n) ? A[(size_t)(r0 + rr) + (size_t)(k0 + t) * lda] : 0.0;
    }
    for (int e = tid; e < NB * TS; e += 256) {
        const int t = e % NB, cc = e / NB;
        Uk[t][cc] = (c0 + cc < n) ? A[(size_t)(k0 + t) + (size_t)(c0 + cc) * lda] : 0.0;
    }
    __syncthreads();
    const int tx = tid % 16, ty = tid / 16;   // rows tx + 16a (coalesced), cols ty + 16c
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.0;
#pragma unroll 8
    for (int t = 0; t < NB; ++t) {
        double lr[4], uc[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) { lr[a] = Lp[t][tx + 16 * a]; uc[a] = Uk[t][ty + 16 * a]; }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][c] = fma(lr[a], uc[c], acc[a][c]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int r = r0 + tx + 16 * a, cc = c0 + ty + 16 * c;
            if (r < n && cc < n) A[(size_t)r + (size_t)cc * lda] -= acc[a][c];
        }
}

// perm[i] = source row of row i after all interchanges (one thread per matrix; n is a few thousand)
__global__ void k_lu_perm(const int* __restrict__ n_arr, const int* __restrict__ ipiv0, int ld_ipiv, int* __restrict__ perm0) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= gridDim.x * blockDim.x) return;
    const int n = n_arr[b];
    const int* ipiv = ipiv0 + (size_t)ld_ipiv * b;
    int* perm = perm0 + (size_t)ld_ipiv * b;
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int j = 0; j < n; ++j) {
        const int p = ipiv[j];
        if (p != j) { const int t = perm[j]; perm[j] = perm[p]; perm[p] = t; }
    }
}

// X[i][:] = B[perm[i]][:]   (row-major, ncols = n + extra_cols)
__global__ void __launch_bounds__(256) k_lu_gather_rows(const double* __restrict__ B0, double* __restrict__ X0, int ldg,
                                                        size_t strideG, const int* __restrict__ n_arr, int extra_cols,
                                                        const int* __restrict__ perm0, int ld_ipiv) {
    const int b = blockIdx.z;
    const int n = n_arr[b];
    const int i = blockIdx.y;
    if (i >= n) return;
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= n + extra_cols) return;
    const int src = perm0[(size_t)ld_ipiv * b + i];
    X0[strideG * b + (size_t)i * ldg + c] = B0[strideG * b + (size_t)src * ldg + c];
}

// block row k of the unit-lower forward solve: triangle for every column (one thread per column)
__global__ void __launch_bounds__(128) k_lu_fwd_diag(const double* __restrict__ A0, int lda, size_t stride, double* __restrict__ X0,
                                                     int ldg, size_t strideG, const int* __restrict__ n_arr, int extra_cols, int k) {
    const int b = blockIdx.y;
    const int n = n_arr[b];
    if (k >= n) return;
    const int nb = (n - k) < NB ? (n - k) : NB;
    const double* A = A0 + stride * b;
    double* X = X0 + strideG * b;
    __shared__ double D[NB][NB + 1];
    for (int e = threadIdx.x; e < NB * NB; e += 128) {
        const int i = e % NB, j = e / NB;
        D[i][j] = (i < nb && j < i) ? A[(size_t)(k + i) + (size_t)(k + j) * lda] : 0.0;
    }
    __syncthreads();
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= n + extra_cols) return;
    double x[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) x[i] = i < nb ? X[(size_t)(k + i) * ldg + c] : 0.0;
#pragma unroll
    for (int i = 1; i < NB; ++i) {
        double s = x[i];
#pragma unroll
        for (int t = 0; t < i; ++t) s = fma(-D[i][t], x[t], s);
        x[i] = s;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i)
        if (i < nb) X[(size_t)(k + i) * ldg + c] = x[i];
}

// block row k of the backward solve with U (upper, non-unit)
__global__ void __launch_bounds__(128) k_lu_bwd_diag(const double* __restrict__ A0, int lda, size_t stride, double* __restrict__ X0,
                                                     int ldg, size_t strideG, const int* __restrict__ n_arr, int extra_cols, int k) {
    const int b = blockIdx.y;
    const int n = n_arr[b];
    if (k >= n) return;
    const int nb = (n - k) < NB ? (n - k) : NB;
    const double* A = A0 + stride * b;
    double* X = X0 + strideG * b;
    __shared__ double D[NB][NB + 1];
    for (int e = threadIdx.x; e < NB * NB; e += 128) {
        const int i = e % NB, j = e / NB;
        D[i][j] = (i < nb && j < nb && j >= i) ? A[(size_t)(k + i) + (size_t)(k + j) * lda] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= n + extra_cols) return;
    double x[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) x[i] = i < nb ? X[(size_t)(k + i) * ldg + c] : 0.0;
#pragma unroll
    for (int i = NB - 1; i >= 0; --i) {
        double s = x[i];
#pragma unroll
        for (int t = i + 1; t < NB; ++t) s = fma(-D[i][t], x[t], s);
        x[i] = s / D[i][i];
    }
#pragma unroll
    for (int i = 0; i < NB; ++i)
        if (i < nb) X[(size_t)(k + i) * ldg + c] = x[i];
}

// X[r][c] -= sum_t M[r][k+t] X[k+t][c]  for the rows r in [lo, hi) : forward (r > panel, M = L) and backward (r < panel, M = U)
__global__ void __launch_bounds__(256) k_lu_update_rows(const double* __restrict__ A0, int lda, size_t stride,
                                                        double* __restrict__ X0, int ldg, size_t strideG,
                                                        const int* __restrict__ n_arr, int extra_cols, int k, int backward) {
    const int b = blockIdx.z;
    const int n = n_arr[b];
    if (k >= n) return;
    const int nb = (n - k) < NB ? (n - k) : NB;
    const int lo = backward ? 0 : k + NB, hi = backward ? k : n;
    const int r0 = lo + blockIdx.y * TS;
    if (r0 >= hi) return;
    const int ncols = n + extra_cols;
    const int c0 = blockIdx.x * TS;
    if (c0 >= ncols) return;
    const double* A = A0 + stride * b;
    double* X = X0 + strideG * b;
    __shared__ double Mp[NB][TS + 1];
    __shared__ double Xk[NB][TS + 1];
    const int tid = threadIdx.x;
    for (int e = tid; e < NB * TS; e += 256) {
        const int rr = e % TS, t = e / TS;
        Mp[t][rr] = (r0 + rr < hi && t < nb) ? A[(size_t)(r0 + rr) + (size_t)(k + t) * lda] : 0.0;
        Xk[t][rr] = (c0 + rr < ncols && t < nb) ? X[(size_t)(k + t) * ldg + c0 + rr] : 0.0;
    }
    __syncthreads();
    const int tx = tid % 16, ty = tid / 16;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.0;
#pragma unroll 8
    for (int t = 0; t < NB; ++t) {
        double mr[4], xc[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) { mr[a] = Mp[t][ty + 16 * a]; xc[a] = Xk[t][tx + 16 * a]; }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][c] = fma(mr[a], xc[c], acc[a][c]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int r = r0 + ty + 16 * a, cc = c0 + tx + 16 * c;
            if (r < hi && cc < ncols) X[(size_t)r * ldg + cc] -= acc[a][c];
        }
}

}  // namespace

int dense_getrf_batched(ptzba_ctx* ctx, double* A, int lda, size_t stride, const int* d_n_arr, int n_max, int batch, int* d_ipiv,
                        int ld_ipiv, int* d_info) {
    cudaStream_t s = ctx->stream;
    CU_CHECK(ctx, cudaMemsetAsync(d_info, 0, sizeof(int), s));
    for (int k0 = 0; k0 < n_max; k0 += NB) {
        k_lu_panel<<<batch, PT, 0, s>>>(A, lda, stride, d_n_arr, k0, d_ipiv, ld_ipiv, d_info);
        KERNEL_POST(ctx);
        k_lu_laswp<<<dim3(div_up(n_max, 256), batch), 256, 0, s>>>(A, lda, stride, d_n_arr, k0, d_ipiv, ld_ipiv);
        KERNEL_POST(ctx);
        const int m = n_max - k0 - NB;
        if (m <= 0) break;
        k_lu_trsm_u12<<<dim3(div_up(m, 128), batch), 128, 0, s>>>(A, lda, stride, d_n_arr, k0);
        KERNEL_POST(ctx);
        k_lu_gemm<<<dim3(div_up(m, TS), div_up(m, TS), batch), 256, 0, s>>>(A, lda, stride, d_n_arr, k0);
        KERNEL_POST(ctx);
    }
    return PTZBA_OK;
}

// X = A^-1 B for row-major B, X [n_b x (n_b + extra_cols)] (leading dimension ldg); B is left untouched.
int dense_getrs_rows_batched(ptzba_ctx* ctx, const double* LU, int lda, size_t stride, const int* d_ipiv, int* d_perm, int ld_ipiv,
                             const double* B, double* X, int ldg, size_t strideG, const int* d_n_arr, int n_max, int extra_cols,
                             int batch) {
    cudaStream_t s = ctx->stream;
    const int ncols_max = n_max + extra_cols;
    k_lu_perm<<<batch, 1, 0, s>>>(d_n_arr, d_ipiv, ld_ipiv, d_perm);
    KERNEL_POST(ctx);
    k_lu_gather_rows<<<dim3(div_up(ncols_max, 256), n_max, batch), 256, 0, s>>>(B, X, ldg, strideG, d_n_arr, extra_cols, d_perm, ld_ipiv);
    KERNEL_POST(ctx);
    for (int k = 0; k < n_max; k += NB) {
        k_lu_fwd_diag<<<dim3(div_up(ncols_max, 128), batch), 128, 0, s>>>(LU, lda, stride, X, ldg, strideG, d_n_arr, extra_cols, k);
        KERNEL_POST(ctx);
        const int m = n_max - k - NB;
        if (m <= 0) break;
        k_lu_update_rows<<<dim3(div_up(ncols_max, TS), div_up(m, TS), batch), 256, 0, s>>>(LU, lda, stride, X, ldg, strideG, d_n_arr,
                                                                                          extra_cols, k, 0);
        KERNEL_POST(ctx);
    }
    const int last = (n_max - 1) / NB * NB;
    for (int k = last; k >= 0; k -= NB) {
        k_lu_bwd_diag<<<dim3(div_up(ncols_max, 128), batch), 128, 0, s>>>(LU, lda, stride, X, ldg, strideG, d_n_arr, extra_cols, k);
        KERNEL_POST(ctx);
        if (k == 0) break;
        k_lu_update_rows<<<dim3(div_up(ncols_max, TS), div_up(k, TS), batch), 256, 0, s>>>(LU, lda, stride, X, ldg, strideG, d_n_arr,
                                                                                          extra_cols, k, 1);
        KERNEL_POST(ctx);
    }
    return PTZBA_OK;
}
