// dense_lu.cu - batched dense FP64 LU factorisation with partial pivoting and the matching multi-RHS solves.
//
// Why LU and not Cholesky: the reference inverts the EKF innovation covariance with numpy.linalg.inv (LAPACK getrf/getri,
// slam_system/ptz_slam.py:259).  Because its covariance write-back drops cross terms (:281-289), P - and with it
// S = H P H^T + R - loses positive definiteness after a few frames (measured: eigenvalues down to -0.19 on seeded
// sequences), so a Cholesky would break down exactly where the reference keeps going.  Partial pivoting reproduces
// getrf's behaviour on those matrices.
//
// Layout: A column-major (lda), matrix b of the batch at A + b*stride, order n_arr[b] (device array).  Right-looking,
// panel width 32:  panel (one CTA per matrix; the panel is factored in sub-panels of W columns that live in SHARED memory
// for their whole pivot search / swap / scale / rank-1 sequence) -> laswp on the other columns -> U12 = L11^-1 A12 ->
// A22 -= L21 U12 (64x64 tiles on the FP64 tensor cores).  Right-hand sides are ROW-major [n x ncols]; the two triangular
// solves are left-looking (one DMMA product over the whole K range per block row, dense_mma.cuh).
#include "common.h"
#include "dense.h"
#include "dense_mma.cuh"

namespace {

constexpr int NB = 32;
constexpr int TS = 64;
constexpr int PT = 512;     // threads of the panel kernel

// One CTA per matrix factors the 32-column panel k0 in sub-panels of W columns.  A sub-panel (rows j0 .. n-1, W columns) is
// copied to shared memory once, goes through its W pivot searches / swaps / scalings / rank-1 updates there, is written back,
// and is then applied to the panel columns right of it (W x w triangular solve + rank-W update through L2).  The round-1
// kernel did every rank-1 update of the whole 32-column panel through global memory: 265 us per panel at n = 1150.
// Shared memory: W * (n - k0) doubles (+ small); W is chosen by the host so that it fits.
template <int W>
__global__ void __launch_bounds__(PT) k_lu_panel(double* __restrict__ A0, int lda, size_t stride, const int* __restrict__ n_arr,
                                                 int k0, int* __restrict__ ipiv0, int ld_ipiv, int* __restrict__ info) {
    extern __shared__ __align__(16) double Sp[];     // [W][mp]: column c of the sub-panel at Sp + c * mp
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
    __shared__ double s_U[W][NB];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int q0 = 0; q0 < nb; q0 += W) {
        const int j0 = k0 + q0;                       // first column / row of the sub-panel
        const int wq = (nb - q0) < W ? (nb - q0) : W;
        const int mp = n - j0;                        // its rows j0 .. n-1
        for (int c = 0; c < wq; ++c)
            for (int r = tid; r < mp; r += PT) Sp[(size_t)c * mp + r] = A[(size_t)(j0 + c) * lda + j0 + r];
        __syncthreads();
        for (int jj = 0; jj < wq; ++jj) {
            const int j = j0 + jj;
            // (1) pivot search in column jj of the sub-panel, local rows jj .. mp-1 (first index of the maximum magnitude, like idamax)
            double best = -1.0;
            int bi = jj;
            const double* col = Sp + (size_t)jj * mp;
            for (int r = jj + tid; r < mp; r += PT) {
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
                    ipiv[j] = j0 + bi;
                    if (!(best > 0.0)) atomicMax(info, j + 1);
                }
            }
            __syncthreads();
            const int pl = s_piv;                      // local pivot row
            // (2) swap rows jj and pl: inside the sub-panel (shared memory) and in the other columns of the panel (global memory);
            // publish the pivot row of the sub-panel
            if (tid < nb) {
                const int c = tid - q0;                // column relative to the sub-panel
                if (c >= 0 && c < wq) {
                    double* cc = Sp + (size_t)c * mp;
                    const double vj = cc[jj], vp = cc[pl];
                    if (pl != jj) { cc[jj] = vp; cc[pl] = vj; }
                    s_row[c] = vp;
                } else if (pl != jj) {
                    double* cg = A + (size_t)(k0 + tid) * lda + j0;
                    const double vj = cg[jj], vp = cg[pl];
                    cg[jj] = vp; cg[pl] = vj;
                }
            }
            __syncthreads();
            const double piv = s_row[jj];
            const double inv = piv != 0.0 ? 1.0 / piv : 0.0;
            // (3) scale the column and rank-1 update of the remaining sub-panel columns, all in shared memory
            for (int r = jj + 1 + tid; r < mp; r += PT) {
                const double l = Sp[(size_t)jj * mp + r] * inv;
                Sp[(size_t)jj * mp + r] = l;
#pragma unroll
                for (int c = 0; c < W; ++c)
                    if (c > jj && c < wq) Sp[(size_t)c * mp + r] = fma(-l, s_row[c], Sp[(size_t)c * mp + r]);
            }
            __syncthreads();
        }
        // (4) write the factored sub-panel back
        for (int c = 0; c < wq; ++c)
            for (int r = tid; r < mp; r += PT) A[(size_t)(j0 + c) * lda + j0 + r] = Sp[(size_t)c * mp + r];
        // (5) apply it to the panel columns right of it: U = L11^-1 A12 (unit lower W x W), then A22 -= L21 U
        const int wr = nb - q0 - wq;                   // columns to the right, inside the panel
        if (wr > 0) {
            __syncthreads();                           // the row swaps of those columns (global memory) are visible
            if (tid < wr) {
                double* cg = A + (size_t)(j0 + wq + tid) * lda + j0;
                double x[W];
#pragma unroll
                for (int i = 0; i < W; ++i) x[i] = i < wq ? cg[i] : 0.0;
#pragma unroll
                for (int i = 1; i < W; ++i) {
                    double sacc = x[i];
#pragma unroll
                    for (int t = 0; t < i; ++t) sacc = fma(-Sp[(size_t)t * mp + i], x[t], sacc);
                    x[i] = sacc;
                }
#pragma unroll
                for (int i = 0; i < W; ++i) {
                    if (i < wq) cg[i] = x[i];
                    s_U[i][tid] = x[i];
                }
            }
            __syncthreads();
            for (int r = wq + tid; r < mp; r += PT) {
                double l[W];
#pragma unroll
                for (int t = 0; t < W; ++t) l[t] = t < wq ? Sp[(size_t)t * mp + r] : 0.0;
                for (int c = 0; c < wr; ++c) {
                    double* dst = A + (size_t)(j0 + wq + c) * lda + j0 + r;
                    double v = *dst;
#pragma unroll
                    for (int t = 0; t < W; ++t) v = fma(-l[t], s_U[t][c], v);
                    *dst = v;
                }
            }
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

// A22 -= L21 U12 ; 64x64 tiles on the FP64 tensor cores (K = 32)
__global__ void __launch_bounds__(dmma::kThreads) k_lu_gemm(double* __restrict__ A0, int lda, size_t stride, const int* __restrict__ n_arr,
                                                           int k0) {
    using T = dmma::Tile<64, 64>;
    extern __shared__ __align__(16) double sm[];
    const int b = blockIdx.z;
    const int n = n_arr[b];
    const int t0 = k0 + NB;
    const int r0 = t0 + blockIdx.x * TS, c0 = t0 + blockIdx.y * TS;
    if (r0 >= n || c0 >= n) return;
    double* A = A0 + stride * b;
    double acc[T::RM][T::RN][2];
#pragma unroll
    for (int a = 0; a < T::RM; ++a)
#pragma unroll
        for (int c = 0; c < T::RN; ++c) acc[a][c][0] = acc[a][c][1] = 0.0;
    // L21(i, t) = A[r0 + i, k0 + t] (i contiguous); U12(t, j) = A[k0 + t, c0 + j] (t contiguous)
    T::accumulate<true>(A + r0 + (size_t)k0 * lda, (size_t)lda, n - r0, A + k0 + (size_t)c0 * lda, (size_t)lda, n - c0, NB, acc, sm);
    T::for_each(acc, [&](int i, int j, double v) {
        if (r0 + i < n && c0 + j < n) A[(size_t)(r0 + i) + (size_t)(c0 + j) * lda] -= v;
    });
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

// block row k (NBS = 64 rows) of X, columns c0 .. c0+63, LEFT-looking:  X[k+i, c] -= sum_{t in [t_lo, t_hi)} M[k+i, t] * X[t, c]
// forward solve with the unit-lower factor: t in [0, k); backward solve with U: t in [k + 64, n).  One DMMA product over the
// whole K range: every entry of X is updated once (the right-looking rank-32 updates of round 1 re-streamed X per panel); the
// 64 x 64 triangle of the block row follows in the same launch (one thread per column, the column in shared memory).
constexpr int NBS = 64;
__global__ void __launch_bounds__(dmma::kThreads) k_lu_rows_ll(const double* __restrict__ A0, int lda, size_t stride,
                                                              double* __restrict__ X0, int ldg, size_t strideG,
                                                              const int* __restrict__ n_arr, int extra_cols, int k, int backward) {
    using T = dmma::Tile<NBS, 64>;
    extern __shared__ __align__(16) double sm[];
    const int b = blockIdx.y;
    const int n = n_arr[b];
    if (k >= n) return;
    const int ncols = n + extra_cols;
    const int c0 = blockIdx.x * 64;
    if (c0 >= ncols) return;
    const int t_lo = backward ? k + NBS : 0, t_hi = backward ? n : k;
    const double* A = A0 + stride * b;
    double* X = X0 + strideG * b;
    double acc[T::RM][T::RN][2];
#pragma unroll
    for (int a = 0; a < T::RM; ++a)
#pragma unroll
        for (int c = 0; c < T::RN; ++c) acc[a][c][0] = acc[a][c][1] = 0.0;
    T::accumulate(A + k + (size_t)t_lo * lda, (size_t)lda, n - k, X + (size_t)t_lo * ldg + c0, (size_t)ldg, ncols - c0, t_hi - t_lo, acc, sm);
    T::for_each(acc, [&](int i, int j, double v) {
        if (k + i < n && c0 + j < ncols) X[(size_t)(k + i) * ldg + c0 + j] -= v;
    });
    const int nb = (n - k) < NBS ? (n - k) : NBS;
    double (*D)[NBS + 1] = reinterpret_cast<double (*)[NBS + 1]>(sm);
    double (*xs)[65] = reinterpret_cast<double (*)[65]>(sm + NBS * (NBS + 1));
    __syncthreads();                                   // the updated rows are visible to the CTA; the ring buffer is free
    for (int e = threadIdx.x; e < NBS * NBS; e += dmma::kThreads) {
        const int i = e % NBS, j = e / NBS;
        const bool in = i < nb && j < nb && (backward ? j >= i : j < i);
        D[i][j] = in ? A[(size_t)(k + i) + (size_t)(k + j) * lda] : ((backward && i == j) ? 1.0 : 0.0);
    }
    __syncthreads();
    const int c = c0 + threadIdx.x;
    if (threadIdx.x >= 64 || c >= ncols) return;
    const int tc = threadIdx.x;
    for (int i = 0; i < nb; ++i) xs[i][tc] = X[(size_t)(k + i) * ldg + c];
    if (!backward) {
        for (int i = 1; i < nb; ++i) {
            double s0 = xs[i][tc], s1 = 0.0, s2 = 0.0, s3 = 0.0;      // four independent chains
            int tt = 0;
            for (; tt + 3 < i; tt += 4) {
                s0 = fma(-D[i][tt], xs[tt][tc], s0); s1 = fma(-D[i][tt + 1], xs[tt + 1][tc], s1);
                s2 = fma(-D[i][tt + 2], xs[tt + 2][tc], s2); s3 = fma(-D[i][tt + 3], xs[tt + 3][tc], s3);
            }
            for (; tt < i; ++tt) s0 = fma(-D[i][tt], xs[tt][tc], s0);
            const double r = (s0 + s1) + (s2 + s3);
            xs[i][tc] = r;
            X[(size_t)(k + i) * ldg + c] = r;
        }
    } else {
        for (int i = nb - 1; i >= 0; --i) {
            double s0 = xs[i][tc], s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int tt = i + 1;
            for (; tt + 3 < nb; tt += 4) {
                s0 = fma(-D[i][tt], xs[tt][tc], s0); s1 = fma(-D[i][tt + 1], xs[tt + 1][tc], s1);
                s2 = fma(-D[i][tt + 2], xs[tt + 2][tc], s2); s3 = fma(-D[i][tt + 3], xs[tt + 3][tc], s3);
            }
            for (; tt < nb; ++tt) s0 = fma(-D[i][tt], xs[tt][tc], s0);
            const double r = ((s0 + s1) + (s2 + s3)) / D[i][i];
            xs[i][tc] = r;
            X[(size_t)(k + i) * ldg + c] = r;
        }
    }
}

}  // namespace

int dense_getrf_batched(ptzba_ctx* ctx, double* A, int lda, size_t stride, const int* d_n_arr, int n_max, int batch, int* d_ipiv,
                        int ld_ipiv, int* d_info) {
    cudaStream_t s = ctx->stream;
    static bool cfg = false;
    if (!cfg) { CU_CHECK(ctx, dmma::configure(k_lu_gemm, dmma::Tile<64, 64>::kSmemBytes)); cfg = true; }
    CU_CHECK(ctx, cudaMemsetAsync(d_info, 0, sizeof(int), s));
    for (int k0 = 0; k0 < n_max; k0 += NB) {
        // widest sub-panel whose rows fit in shared memory
        const size_t rows = (size_t)(n_max - k0);
        // a single CTA's column steps are a latency chain: with more matrices than SMs two panel CTAs share an SM (narrower
        // sub-panels), with few matrices the widest sub-panel that fits wins
        const size_t lim = (batch > ctx->sm_count ? 100 : 200) * 1024;
        if (rows * 8 * sizeof(double) <= lim) {
            static bool cfg8 = false;
            if (!cfg8) { CU_CHECK(ctx, cudaFuncSetAttribute(k_lu_panel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); cfg8 = true; }
            k_lu_panel<8><<<batch, PT, rows * 8 * sizeof(double), s>>>(A, lda, stride, d_n_arr, k0, d_ipiv, ld_ipiv, d_info);
        } else if (rows * 4 * sizeof(double) <= lim) {
            static bool cfg4 = false;
            if (!cfg4) { CU_CHECK(ctx, cudaFuncSetAttribute(k_lu_panel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); cfg4 = true; }
            k_lu_panel<4><<<batch, PT, rows * 4 * sizeof(double), s>>>(A, lda, stride, d_n_arr, k0, d_ipiv, ld_ipiv, d_info);
        } else if (rows * 2 * sizeof(double) <= lim) {
            static bool cfg2 = false;
            if (!cfg2) { CU_CHECK(ctx, cudaFuncSetAttribute(k_lu_panel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); cfg2 = true; }
            k_lu_panel<2><<<batch, PT, rows * 2 * sizeof(double), s>>>(A, lda, stride, d_n_arr, k0, d_ipiv, ld_ipiv, d_info);
        } else {
            if (rows * sizeof(double) > 200 * 1024)
                return ptzba_fail(ctx, PTZBA_ERR_ARG, "LU panel of %zu rows exceeds the shared-memory sub-panel", rows);
            static bool cfg1 = false;
            if (!cfg1) { CU_CHECK(ctx, cudaFuncSetAttribute(k_lu_panel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); cfg1 = true; }
            k_lu_panel<1><<<batch, PT, rows * sizeof(double), s>>>(A, lda, stride, d_n_arr, k0, d_ipiv, ld_ipiv, d_info);
        }
        KERNEL_POST(ctx);
        k_lu_laswp<<<dim3(div_up(n_max, 256), batch), 256, 0, s>>>(A, lda, stride, d_n_arr, k0, d_ipiv, ld_ipiv);
        KERNEL_POST(ctx);
        const int m = n_max - k0 - NB;
        if (m <= 0) break;
        k_lu_trsm_u12<<<dim3(div_up(m, 128), batch), 128, 0, s>>>(A, lda, stride, d_n_arr, k0);
        KERNEL_POST(ctx);
        k_lu_gemm<<<dim3(div_up(m, TS), div_up(m, TS), batch), dmma::kThreads, dmma::Tile<64, 64>::kSmemBytes, s>>>(A, lda, stride, d_n_arr, k0);
        KERNEL_POST(ctx);
    }
    return PTZBA_OK;
}

// X = A^-1 B for row-major B, X [n_b x (n_b + extra_cols)] (leading dimension ldg); B is left untouched.
int dense_getrs_rows_batched(ptzba_ctx* ctx, const double* LU, int lda, size_t stride, const int* d_ipiv, int* d_perm, int ld_ipiv,
                             const double* B, double* X, int ldg, size_t strideG, const int* d_n_arr, int n_max, int extra_cols,
                             int batch) {
    cudaStream_t s = ctx->stream;
    static bool cfg = false;
    if (!cfg) { CU_CHECK(ctx, dmma::configure(k_lu_rows_ll, dmma::Tile<NBS, 64>::kSmemBytes)); cfg = true; }
    const int ncols_max = n_max + extra_cols;
    k_lu_perm<<<batch, 1, 0, s>>>(d_n_arr, d_ipiv, ld_ipiv, d_perm);
    KERNEL_POST(ctx);
    k_lu_gather_rows<<<dim3(div_up(ncols_max, 256), n_max, batch), 256, 0, s>>>(B, X, ldg, strideG, d_n_arr, extra_cols, d_perm, ld_ipiv);
    KERNEL_POST(ctx);
    for (int k = 0; k < n_max; k += NBS) {
        k_lu_rows_ll<<<dim3(div_up(ncols_max, 64), batch), dmma::kThreads, dmma::Tile<NBS, 64>::kSmemBytes, s>>>(LU, lda, stride, X, ldg, strideG, d_n_arr, extra_cols, k, 0);
        KERNEL_POST(ctx);
    }
    const int last = (n_max - 1) / NBS * NBS;
    for (int k = last; k >= 0; k -= NBS) {
        // (sequences shorter than n_max skip the block rows beyond their order; within a sequence the rows right of block k are final)
        k_lu_rows_ll<<<dim3(div_up(ncols_max, 64), batch), dmma::kThreads, dmma::Tile<NBS, 64>::kSmemBytes, s>>>(LU, lda, stride, X, ldg, strideG, d_n_arr, extra_cols, k, 1);
        KERNEL_POST(ctx);
    }
    return PTZBA_OK;
}
