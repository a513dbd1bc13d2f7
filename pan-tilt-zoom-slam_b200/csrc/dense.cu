// dense.cu - hand-written BATCHED dense FP64 Cholesky kernels (the innovation covariances S = H P H^T + R of many independent
// EKF sequences, ptz_slam.py:256-259; the reference inverts S with numpy's LU).
//
// Blocked LEFT-looking Cholesky (lower, column-major, in place), one launch pair per 32-wide panel for the whole batch:
//     k_chol_ll_update  64 x 32 tiles of block column k: C -= L[:, 0:k] L[k:k+32, 0:k]^T  - one DMMA product over the whole K
//                       range, so every entry of the matrix is read and written ONCE (the right-looking rank-32 updates of
//                       round 1 re-streamed the trailing matrix for every panel: ~1.3 FMA per byte)
//     k_potf2_trsm      every CTA factors the 32 x 32 diagonal block redundantly in shared memory, then solves its 128 panel rows
// and the multi-right-hand-side forward solve Z = L^-1 G on row-major G, left-looking as well (k_fwd_ll_update: 32 x 64 tiles of
// block row k, K = k; k_fwd_diag_rows).  The tile products run on the FP64 tensor cores (dense_mma.cuh).
// The single reduced camera system of bundle adjustment uses the cooperative one-kernel path of dense_coop.cu instead.
#include "common.h"
#include "dense.h"
#include "dense_mma.cuh"

namespace {

constexpr int NB = 32;        // panel width

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

// ---- left-looking update of block column k (32 columns): rows r0 .. r0+63 of it, C -= L[r0.., 0:k] * L[k..k+31, 0:k]^T ----------
__global__ void __launch_bounds__(dmma::kThreads) k_chol_ll_update(double* __restrict__ A0, int lda, size_t stride,
                                                                  const int* __restrict__ n_arr, int n_fixed, int k) {
    using T = dmma::Tile<64, 32>;
    extern __shared__ __align__(16) double sm[];
    const int n = n_arr ? n_arr[blockIdx.y] : n_fixed;
    const int r0 = k + blockIdx.x * 64;
    if (k >= n || r0 >= n) return;
    double* A = A0 + stride * blockIdx.y;
    double acc[T::RM][T::RN][2];
#pragma unroll
    for (int a = 0; a < T::RM; ++a)
#pragma unroll
        for (int b = 0; b < T::RN; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    T::accumulate(A + r0, (size_t)lda, n - r0, A + k, (size_t)lda, n - k, k, acc, sm);
    T::for_each(acc, [&](int i, int j, double v) {
        const int gi = r0 + i, gj = k + j;
        if (gi < n && gj < n && gi >= gj) A[(size_t)gi + (size_t)gj * lda] -= v;
    });
}

}  // namespace

int dense_potrf_lower_batched(ptzba_ctx* ctx, double* A, int lda, size_t stride, const int* d_n_arr, int n_max, int batch,
                              int* d_info, int* d_info_per_batch) {
    cudaStream_t s = ctx->stream;
    static bool cfg = false;
    if (!cfg) { CU_CHECK(ctx, dmma::configure(k_chol_ll_update, dmma::Tile<64, 32>::kSmemBytes)); cfg = true; }
    CU_CHECK(ctx, cudaMemsetAsync(d_info, 0, sizeof(int), s));
    for (int k = 0; k < n_max; k += NB) {
        if (k > 0) {
            k_chol_ll_update<<<dim3(div_up(n_max - k, 64), batch), dmma::kThreads, dmma::Tile<64, 32>::kSmemBytes, s>>>(A, lda, stride, d_n_arr, n_max, k);
            KERNEL_POST(ctx);
        }
        const int m = n_max - k - NB;
        k_potf2_trsm<<<dim3(m > 0 ? div_up(m, 128) : 1, batch), 128, 0, s>>>(A, lda, stride, d_n_arr, n_max, k, d_info, d_info_per_batch);
        KERNEL_POST(ctx);
    }
    return PTZBA_OK;
}

namespace {

// block row k (NBS = 64 rows) of  L Z = G  (L column-major lower, G row-major [n x ncols], ld = ldg), columns c0 .. c0+63:
//   G[k+i, c] -= sum_{t<k} L[k+i, t] * Z[t, c]   (left-looking, K = k; FP64 tensor cores), then the 64 x 64 triangle for the same
// columns (one thread per column, the column in shared memory), so that a block row is ONE launch.  64-row blocks: a 64 x 64 tile
// needs (64 + 64) doubles from L2 per 64 x 64 FMAs and K step - the 32-row version was L2-bandwidth bound at half the DMMA peak.
constexpr int NBS = 64;
__global__ void __launch_bounds__(dmma::kThreads) k_fwd_ll_update(const double* __restrict__ L0, int lda, size_t strideL,
                                                                 double* __restrict__ G0, int ldg, size_t strideG,
                                                                 const int* __restrict__ n_arr, int extra_cols, int k) {
    using T = dmma::Tile<NBS, 64>;
    extern __shared__ __align__(16) double sm[];
    const int n = n_arr[blockIdx.y];
    if (k >= n) return;
    const int ncols = n + extra_cols;
    const int c0 = blockIdx.x * 64;
    if (c0 >= ncols) return;
    const double* L = L0 + strideL * blockIdx.y;
    double* G = G0 + strideG * blockIdx.y;
    double acc[T::RM][T::RN][2];
#pragma unroll
    for (int a = 0; a < T::RM; ++a)
#pragma unroll
        for (int b = 0; b < T::RN; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    T::accumulate(L + k, (size_t)lda, n - k, G + c0, (size_t)ldg, ncols - c0, k, acc, sm);
    T::for_each(acc, [&](int i, int j, double v) {
        if (k + i < n && c0 + j < ncols) G[(size_t)(k + i) * ldg + c0 + j] -= v;
    });
    const int nb = (n - k) < NBS ? (n - k) : NBS;
    double (*D)[NBS + 1] = reinterpret_cast<double (*)[NBS + 1]>(sm);
    double (*xs)[65] = reinterpret_cast<double (*)[65]>(sm + NBS * (NBS + 1));
    __syncthreads();                                   // the updated rows are visible to the CTA; the ring buffer is free
    for (int e = threadIdx.x; e < NBS * NBS; e += dmma::kThreads) {
        const int i = e % NBS, j = e / NBS;
        D[i][j] = (i < nb && j < nb && j <= i) ? L[(size_t)(k + i) + (size_t)(k + j) * lda] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    const int c = c0 + threadIdx.x;
    if (threadIdx.x >= 64 || c >= ncols) return;
    const int tc = threadIdx.x;
    for (int i = 0; i < nb; ++i) xs[i][tc] = G[(size_t)(k + i) * ldg + c];
    for (int i = 0; i < nb; ++i) {
        double s0 = xs[i][tc], s1 = 0.0, s2 = 0.0, s3 = 0.0;      // four independent chains
        int tt = 0;
        for (; tt + 3 < i; tt += 4) {
            s0 = fma(-D[i][tt], xs[tt][tc], s0); s1 = fma(-D[i][tt + 1], xs[tt + 1][tc], s1);
            s2 = fma(-D[i][tt + 2], xs[tt + 2][tc], s2); s3 = fma(-D[i][tt + 3], xs[tt + 3][tc], s3);
        }
        for (; tt < i; ++tt) s0 = fma(-D[i][tt], xs[tt][tc], s0);
        const double r = ((s0 + s1) + (s2 + s3)) / D[i][i];
        xs[i][tc] = r;
        G[(size_t)(k + i) * ldg + c] = r;
    }
}

}  // namespace

// Z = L^-1 G in place for a batch; G row-major with n_b rows and n_b + extra_cols columns
int dense_fwd_solve_rows_batched(ptzba_ctx* ctx, const double* L, int lda, size_t strideL, double* G, int ldg, size_t strideG,
                                 const int* d_n_arr, int n_max, int extra_cols, int batch) {
    cudaStream_t s = ctx->stream;
    static bool cfg = false;
    if (!cfg) { CU_CHECK(ctx, dmma::configure(k_fwd_ll_update, dmma::Tile<NBS, 64>::kSmemBytes)); cfg = true; }
    const int ncols_max = n_max + extra_cols;
    for (int k = 0; k < n_max; k += NBS) {
        k_fwd_ll_update<<<dim3(div_up(ncols_max, 64), batch), dmma::kThreads, dmma::Tile<NBS, 64>::kSmemBytes, s>>>(L, lda, strideL, G, ldg, strideG, d_n_arr, extra_cols, k);
        KERNEL_POST(ctx);
    }
    return PTZBA_OK;
}
