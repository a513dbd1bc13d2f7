// dense.h - dense FP64 factorisation entry points (dense.cu); all pointers are device pointers, work is enqueued on
// ctx->stream.  Matrices are column-major, lower triangle referenced.
#pragma once
#include "common.h"

// in-place Cholesky A = L L^T.  *d_info (device int) = 0 on success, else 1 + first row of the failing panel.
int dense_potrf_lower(ptzba_ctx* ctx, double* A, int n, int lda, int* d_info);
// solves L L^T X = B in place for nrhs (1 or 2) right-hand sides stored as columns of B (ldb)
int dense_potrs_lower(ptzba_ctx* ctx, const double* L, int n, int lda, double* B, int ldb, int nrhs);
