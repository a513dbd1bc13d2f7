// dense.h - dense FP64 factorisation entry points (dense.cu); all pointers are device pointers, work is enqueued on
// ctx->stream.  Matrices are column-major, lower triangle referenced.
#pragma once
#include "common.h"

// batched in-place Cholesky A_b = L_b L_b^T: matrix b lives at A + b*stride and has order d_n_arr[b] (device array; nullptr = n_max for all)
// d_info_per_batch (may be nullptr): entry b is set non-zero when matrix b is not positive definite (must be pre-zeroed)
int dense_potrf_lower_batched(ptzba_ctx* ctx, double* A, int lda, size_t stride, const int* d_n_arr, int n_max, int batch,
                              int* d_info, int* d_info_per_batch);
// Z = L^-1 G in place; G is ROW-major with d_n_arr[b] rows and d_n_arr[b] + extra_cols columns (leading dimension ldg)
int dense_fwd_solve_rows_batched(ptzba_ctx* ctx, const double* L, int lda, size_t strideL, double* G, int ldg, size_t strideG,
                                 const int* d_n_arr, int n_max, int extra_cols, int batch);

// ---- LU with partial pivoting (dense_lu.cu); same batching convention ---------------------------------------------
// d_ipiv / d_perm: [batch x ld_ipiv] ints.  *d_info = 0, or 1 + the first column whose pivot was exactly zero.
int dense_getrf_batched(ptzba_ctx* ctx, double* A, int lda, size_t stride, const int* d_n_arr, int n_max, int batch, int* d_ipiv,
                        int ld_ipiv, int* d_info);
// X = A^-1 B; B, X row-major [n_b x (n_b + extra_cols)], leading dimension ldg; B is left untouched
int dense_getrs_rows_batched(ptzba_ctx* ctx, const double* LU, int lda, size_t stride, const int* d_ipiv, int* d_perm, int ld_ipiv,
                             const double* B, double* X, int ldg, size_t strideG, const int* d_n_arr, int n_max, int extra_cols,
                             int batch);

// ---- single-matrix, latency-optimised path (dense_coop.cu): one persistent cooperative kernel factors A = L L^T and
// inverts the 128 x 128 diagonal blocks of L into Dinv_store (dense_coop_dinv_doubles(n) doubles).  *d_info (device int) = 0, or 1 + the first row of the panel with a non-positive pivot.
size_t dense_coop_dinv_doubles(int n);
int dense_potrf_coop(ptzba_ctx* ctx, double* A, int n, int lda, int* d_info, double* Dinv_store);
int dense_potrs_coop(ptzba_ctx* ctx, const double* L, int n, int lda, const double* Dinv_store, double* b);
