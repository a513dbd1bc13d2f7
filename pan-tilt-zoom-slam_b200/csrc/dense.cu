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
__global__ void __launch_bounds__(32) k_potf2(double* __restrict__ A0, int lda, size_t stride, const int* __restrict__ n_arr,
                                              int n_fixed, int k, int* __restrict__ info) {
    const int lane = threadIdx.x;
    const int n = n_arr ? n_arr[blockIdx.y] : n_fixed;
    if (k >= n) return;
    const int nb = (n - k) < NB ? (n - k) : NB;
    double* A = A0 + stride * blockIdx.y;
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
__global__ void __launch_bounds__(128) k_trsm_panel(double* __restrict__ A0, int lda, size_t stride,
                                                    const int* __restrict__ n_arr, int n_fixed, int k) {
    const int n = n_arr ? n_arr[blockIdx.y] : n_fixed;
    if (k + NB >= n) return;
    const int nb = NB;
    double* A = A0 + stride * blockIdx.y;
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

int dense_potrf_lower(ptzba_ctx* ctx, double* A, int n, int lda, int* d_info) {
    return dense_potrf_lower_batched(ctx, A, lda, 0, nullptr, n, 1, d_info, nullptr);
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

namespace {

// ---- inverses of the 32 x 32 diagonal blocks of L (one warp per block; lane j solves L x = e_j) ----------------------------
// Dinv block b is stored column-major: Dinv[b*1024 + i + 32*j] = (L_bb^-1)[i][j]
__global__ void __launch_bounds__(32) k_diag_inverse(const double* __restrict__ L, int lda, int n, double* __restrict__ Dinv) {
    __shared__ double T[NB][NB + 1];
    const int b = blockIdx.x, k = b * NB, lane = threadIdx.x;
    const int nb = (n - k) < NB ? (n - k) : NB;
    for (int e = lane; e < NB * NB; e += 32) {
        const int i = e % NB, j = e / NB;
        T[i][j] = (i < nb && j < nb && j <= i) ? L[(size_t)(k + i) + (size_t)(k + j) * lda] : (i == j ? 1.0 : 0.0);
    }
    __syncwarp();
    double x[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        double s = (i == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int t = 0; t < i; ++t) s = fma(-T[i][t], x[t], s);
        x[i] = s / T[i][i];
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) Dinv[(size_t)b * NB * NB + i + NB * lane] = x[i];
}

// ---- single right-hand side solves with the inverted diagonal blocks: every block step is a 32 x 32 mat-vec by one warp
// followed by a panel update spread over the whole CTA (no sequential triangle solve on the critical path) -----------------
__global__ void __launch_bounds__(1024) k_trsv_dinv(const double* __restrict__ L, int lda, int n, const double* __restrict__ Dinv,
                                                    double* __restrict__ bvec, int backward) {
    __shared__ double xs[NB];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nblk = (n + NB - 1) / NB;
    for (int bi = 0; bi < nblk; ++bi) {
        const int blk = backward ? (nblk - 1 - bi) : bi;
        const int k = blk * NB;
        const int nb = (n - k) < NB ? (n - k) : NB;
        {
            // x = Dinv * b_k (forward) or Dinv^T * b_k (backward): warp w forms entry w, one product per lane, all loads
            // independent (1024 threads = 32 x 32 entries of the block)
            const double* D = Dinv + (size_t)blk * NB * NB;
            const double bt = lane < nb ? bvec[k + lane] : 0.0;
            const double d = backward ? D[lane + NB * w] : D[w + NB * lane];
            double sacc = d * bt;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, off);
            if (lane == 0) xs[w] = w < nb ? sacc : 0.0;
        }
        __syncthreads();
        if (w == 0 && lane < nb) bvec[k + lane] = xs[lane];
        if (!backward) {
            for (int r = k + nb + tid; r < n; r += 1024) {
                double s = 0.0;
#pragma unroll 8
                for (int j = 0; j < NB; ++j) s = fma(j < nb ? L[(size_t)r + (size_t)(k + j) * lda] : 0.0, xs[j], s);
                bvec[r] -= s;
            }
        } else {
            for (int c = tid; c < k; c += 1024) {
                const double* col = L + (size_t)k + (size_t)c * lda;
                double s = 0.0;
#pragma unroll 8
                for (int i = 0; i < NB; ++i) s = fma(i < nb ? col[i] : 0.0, xs[i], s);
                bvec[c] -= s;
            }
        }
        __syncthreads();
    }
}

}  // namespace

int dense_diag_inverse(ptzba_ctx* ctx, const double* L, int n, int lda, double* Dinv) {
    if (n <= 0) return PTZBA_OK;
    k_diag_inverse<<<div_up(n, NB), 32, 0, ctx->stream>>>(L, lda, n, Dinv);
    KERNEL_POST(ctx);
    return PTZBA_OK;
}

int dense_potrs_dinv(ptzba_ctx* ctx, const double* L, int n, int lda, const double* Dinv, double* b) {
    if (n <= 0) return PTZBA_OK;
    k_trsv_dinv<<<1, 1024, 0, ctx->stream>>>(L, lda, n, Dinv, b, 0);
    KERNEL_POST(ctx);
    k_trsv_dinv<<<1, 1024, 0, ctx->stream>>>(L, lda, n, Dinv, b, 1);
    KERNEL_POST(ctx);
    return PTZBA_OK;
}
