// dense_coop.cu - latency-optimised FP64 Cholesky / solves for ONE matrix of moderate order (the Schur-reduced camera
// system of keyframe bundle adjustment: order 3(N-1) = 765 at 256 keyframes, 3069 at 1024).
//
// The blocked factorisation of dense.cu launches two kernels per 32-wide panel; at n = 765 that is 48 dependent launches
// whose critical path (single-CTA panel factorisation, launch gaps) costs ~0.94 ms for 1.5e8 flops.  Here the whole
// factorisation is ONE persistent cooperative kernel (one CTA per SM, device-wide barriers between the phases of a panel):
//     per panel k:   every CTA factors the 32 x 32 diagonal block redundantly (one warp, registers + shuffles) and inverts
//                    it, so the panel solve X = A_panel L_kk^-T is a small mat-mul with independent dot products;
//                    barrier; lower 64 x 64 tiles of the trailing matrix are updated, tiles round-robin over the CTAs;
//                    barrier.
//     afterwards:    inverses of the 128 x 128 diagonal blocks of L by two levels of block doubling
//                        inv [[A 0] [C B]] = [[A^-1 0] [-B^-1 C A^-1  B^-1]]
//                    so that a triangular solve is 6 block steps (n = 765) instead of 24.
// Loads of matrix entries written by other CTAs inside the kernel bypass L1 (__ldcg): L1 is not coherent across SMs.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.h"
#include "dense.h"

namespace {

constexpr int NB = 32;         // panel width
constexpr int TS = 64;         // trailing-update tile
constexpr int IB = 128;        // order of the inverted diagonal blocks used by the solves
constexpr int kCoopWorkers = 256;                  // threads of the panel solve / trailing update
constexpr int kCoopHelpers = 128;                  // look-ahead group: updates and factors the NEXT diagonal block during the trailing update
constexpr int kCoopCta = kCoopWorkers + kCoopHelpers;

constexpr int kTri = NB * (NB + 1) / 2;            // entries of a 32 x 32 lower triangle

// column-major list of the lower-triangle positions (i << 8 | c): the entries right of column j are the suffix that starts at
// tri_offset(j + 1), so the trailing update of a Cholesky column is spread evenly over the lanes of one warp
__device__ __forceinline__ int tri_offset(int c) { return NB * c - c * (c - 1) / 2; }
__device__ __forceinline__ void warp_build_tri(unsigned short* tab, int lane) {
    for (int c = 0; c < NB; ++c)
        if (lane >= c) tab[tri_offset(c) + lane - c] = (unsigned short)((lane << 8) | c);
    __syncwarp();
}

// look-ahead group: Cholesky of the 32 x 32 block held in Ls (lower, identity padded); 1 / L_jj into sInv.
// ONE warp, lane = row, right-looking, warp-synchronous (no named barrier): per column the pivot is a broadcast read, every lane
// scales its entry of the column, then updates its own row right of it - (31 - j) independent read-modify-writes whose partner
// L(c, j) is a broadcast read.  The round-1 version (128 threads, entries in registers, two bar.sync per column) needed 610
// cycles per column (19.5 k per panel, the critical path of the whole factorisation at n = 765).
__device__ __forceinline__ bool warp_potf2(double (*Ls)[NB + 1], double* sInv, double (*col)[NB], int lane) {
    // lane = row, the row lives in registers (fully unrolled: 496 FMAs, ~1.6 k instructions); only the scaled column travels
    // through shared memory (double buffered: one __syncwarp per column).  Entries right of the diagonal hold garbage that is
    // never read.
    double a[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) a[c] = Ls[lane][c];
    bool bad = false;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        double piv = __shfl_sync(0xffffffffu, a[j], j);
        if (!(piv > 0.0)) { bad = true; piv = 1.0; }
        const double inv = rsqrt(piv);
        const double lij = (lane == j) ? piv * inv : a[j] * inv;
        a[j] = lij;
        col[j & 1][lane] = lij;
        if (lane == 0) sInv[j] = inv;
        __syncwarp();
#pragma unroll
        for (int c = j + 1; c < NB; ++c) a[c] = fma(-lij, col[j & 1][c], a[c]);
    }
#pragma unroll
    for (int c = 0; c < NB; ++c)
        if (c <= lane) Ls[lane][c] = a[c];
    __syncwarp();
    return bad;
}

__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(kCoopCta)
k_potrf_coop(double* __restrict__ A, int n, int lda, int* __restrict__ info, double* __restrict__ Dinv32,
             double* __restrict__ Dinv, long long* __restrict__ dbg) {
    // dbg (PTZBA_TRACE): SM clock cycles of CTA 0 per phase {first block, panel solve, barrier, trailing update | look-ahead, barrier, block inverses}
    long long tprev = dbg ? clock64() : 0;
#define COOP_TICK(slot)                                                         \
    do {                                                                        \
        if (dbg && blockIdx.x == 0 && threadIdx.x == 0) {                       \
            const long long tn = clock64();                                     \
            dbg[slot] += tn - tprev;                                            \
            tprev = tn;                                                         \
        }                                                                       \
    } while (0)
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    extern __shared__ __align__(16) double sm[];
    // shared-memory map (doubles): Ls[32][33] | Li[32][33] | Pi[32][65] | Pj[32][65] | Ws[32][33]  (the doubling levels reuse all of it)
    double (*Ls)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm);
    double (*Li)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm + NB * (NB + 1));
    double (*Pi)[TS + 1] = reinterpret_cast<double (*)[TS + 1]>(sm + 2 * NB * (NB + 1));
    double (*Pj)[TS + 1] = reinterpret_cast<double (*)[TS + 1]>(sm + 2 * NB * (NB + 1) + NB * (TS + 1));
    double (*Ws)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(sm + 2 * NB * (NB + 1) + 2 * NB * (TS + 1));
    __shared__ double sInv[NB];
    __shared__ unsigned short sTri[kTri];
    __shared__ double sCol[2][NB];
    const int tid = threadIdx.x, lane = tid & 31;
    const bool helper = tid >= kCoopWorkers;
    const int G = gridDim.x, cta = blockIdx.x;
    const int nblk = (n + NB - 1) / NB;
    // ---- first diagonal block (every CTA, redundantly) ----
    for (int e = tid; e < NB * NB; e += kCoopCta) {
        const int i = e % NB, j = e / NB;
        Ls[i][j] = (i < n && j < n && j <= i) ? A[(size_t)i + (size_t)j * lda] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    if (tid >= kCoopWorkers && tid < kCoopWorkers + 32) {
        warp_build_tri(sTri, lane);
        if (warp_potf2(Ls, sInv, sCol, lane) && lane == 0 && cta == 0) atomicMax(info, 1);
    }
    grid.sync();        // every CTA has read the first block before CTA 0 overwrites it with its factor (also a CTA barrier)
    COOP_TICK(0);
    for (int kb = 0; kb < nblk; ++kb) {
        const int k = kb * NB;
        const int nb = (n - k) < NB ? (n - k) : NB;
        // Ls holds the factor of diagonal block kb, sInv the reciprocals of its diagonal
        if (cta == 0) {
            for (int e = tid; e < NB * NB; e += kCoopCta) {
                const int i = e % NB, j = e / NB;
                if (i < nb && j < nb && j <= i) A[(size_t)(k + i) + (size_t)(k + j) * lda] = Ls[i][j];
            }
        }
        const int m0 = k + NB;                      // first row below the panel (only full panels have rows below)
        if (m0 >= n) break;
        // ---- panel rows below the block: X = A_panel * L_kk^-T; 64 rows per CTA staged in shared memory, 4 threads per row ----
        if (tid < 64) {
            // row r of the panel: x L_kk^T = a by right-looking substitution in registers (32 loads in flight, 528 FMAs)
            for (int rb = cta * 64; m0 + rb < n; rb += G * 64) {
                const int r = m0 + rb + tid;
                if (r < n) {
                    double a[NB];
#pragma unroll
                    for (int j = 0; j < NB; ++j) a[j] = __ldcg(A + (size_t)r + (size_t)(k + j) * lda);
#pragma unroll
                    for (int t = 0; t < NB; ++t) {
                        a[t] *= sInv[t];
#pragma unroll
                        for (int j = t + 1; j < NB; ++j) a[j] = fma(-a[t], Ls[j][t], a[j]);
                    }
#pragma unroll
                    for (int j = 0; j < NB; ++j) A[(size_t)r + (size_t)(k + j) * lda] = a[j];
                }
            }
        }
        COOP_TICK(1);
        grid.sync();
        COOP_TICK(2);
        if (helper) {
            // ---- look-ahead: the next diagonal block, updated with this panel's rows, factored and inverted ----
            const int k1 = m0;
            long long th = dbg ? clock64() : 0;
#define HELP_TICK(slot)                                                         \
    do {                                                                        \
        if (dbg && blockIdx.x == 0 && threadIdx.x == kCoopWorkers) {             \
            const long long tn = clock64();                                     \
            dbg[slot] += tn - th;                                               \
            th = tn;                                                            \
        }                                                                       \
    } while (0)
            const int hid = tid - kCoopWorkers;
            asm volatile("bar.sync 3, 384;" ::: "memory");      // the workers have loaded and updated the next diagonal block (Ls)
            HELP_TICK(7);
            if (hid < 32 && warp_potf2(Ls, sInv, sCol, lane) && hid == 0 && cta == 0) atomicMax(info, k1 + 1);
            if (dbg && blockIdx.x == 0 && threadIdx.x == kCoopWorkers) dbg[16 + kb] = clock64() - th;
            HELP_TICK(8);
#undef HELP_TICK
        } else {
            // ---- trailing update: lower 64 x 64 tiles, C -= P_i P_j^T.  Software pipelined over this CTA's tiles: the panel rows
            // of the NEXT tile are requested (registers) before the current tile is multiplied, the tile's old values are
            // requested at its start and only consumed at its end, so no L2 round trip is exposed between tiles ----
            {
                // ---- next diagonal block: load it and this panel's rows of it, apply the panel's update, hand it to the look-ahead
                // group (which factors it while the tiles below are updated); every CTA keeps its own copy ----
                const int k1 = m0, nb1 = (n - k1) < NB ? (n - k1) : NB;
                const int row = tid & 31, j0 = (tid >> 5) * 4;
                double w[4], d[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = j0 + jj;
                    w[jj] = row < nb1 ? __ldcg(A + (size_t)(k1 + row) + (size_t)(k + j) * lda) : 0.0;                    // L(k1 + row, k + j)
                    d[jj] = (row < nb1 && j < nb1 && j <= row) ? __ldcg(A + (size_t)(k1 + row) + (size_t)(k1 + j) * lda) : (row == j ? 1.0 : 0.0);
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) { Ws[row][j0 + jj] = w[jj]; Ls[row][j0 + jj] = d[jj]; }
                worker_sync();
                for (int e = tid; e < kTri; e += kCoopWorkers) {    // Ls(i, c) -= sum_t W(i, t) W(c, t)
                    const int code = sTri[e], i = code >> 8, c = code & 255;
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                    for (int t = 0; t < NB; t += 4) {
                        s0 = fma(Ws[i][t], Ws[c][t], s0);
                        s1 = fma(Ws[i][t + 1], Ws[c][t + 1], s1);
                        s2 = fma(Ws[i][t + 2], Ws[c][t + 2], s2);
                        s3 = fma(Ws[i][t + 3], Ws[c][t + 3], s3);
                    }
                    Ls[i][c] -= (s0 + s1) + (s2 + s3);
                }
                asm volatile("bar.arrive 3, 384;" ::: "memory");
            }
            const int m = n - m0;
            const int nt = (m + TS - 1) / TS;
            const int ntiles = nt * (nt + 1) / 2;
            const int tx = tid % 16, ty = tid / 16;
            auto tile_rc = [&](int b, int& r0, int& c0) {
                int ti = (int)((sqrtf(8.0f * (float)b + 1.0f) - 1.0f) * 0.5f);
                while ((ti + 1) * (ti + 2) / 2 <= b) ++ti;
                while (ti * (ti + 1) / 2 > b) --ti;
                const int tj = b - ti * (ti + 1) / 2;
                r0 = m0 + ti * TS; c0 = m0 + tj * TS;
            };
            double npi[8], npj[8];                                       // this thread's share of the next tile's panel rows
            auto load_panels = [&](int b) {
                int r0, c0;
                tile_rc(b, r0, c0);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = tid + kCoopWorkers * u, rr = e % TS, t = e / TS;
                    const int gi = r0 + rr, gj = c0 + rr;
                    npi[u] = gi < n ? __ldcg(A + (size_t)gi + (size_t)(k + t) * lda) : 0.0;
                    npj[u] = gj < n ? __ldcg(A + (size_t)gj + (size_t)(k + t) * lda) : 0.0;
                }
            };
            if (cta < ntiles) load_panels(cta);
            for (int b = cta; b < ntiles; b += G) {
                int r0, c0;
                tile_rc(b, r0, c0);
                double cold[4][4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int gi = r0 + tx + 16 * a, gj = c0 + ty + 16 * c;
                        cold[a][c] = (gi < n && gj < n && gi >= gj) ? __ldcg(A + (size_t)gi + (size_t)gj * lda) : 0.0;
                    }
                worker_sync();                                           // previous tile's products are done with Pi / Pj
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = tid + kCoopWorkers * u, rr = e % TS, t = e / TS;
                    Pi[t][rr] = npi[u];
                    Pj[t][rr] = npj[u];
                }
                worker_sync();
                if (b + G < ntiles) load_panels(b + G);
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
                        // the next diagonal block is private to the look-ahead group (every CTA updates its own copy)
                        if (gi < n && gj < n && gi >= gj && !(gi < m0 + NB && gj < m0 + NB)) A[(size_t)gi + (size_t)gj * lda] = cold[a][c] - acc[a][c];
                    }
            }
        }
        COOP_TICK(3);
        grid.sync();
        COOP_TICK(4);
    }
    grid.sync();
    // ---- inverses of the 32 x 32 diagonal blocks of L (one block per CTA): right-looking forward substitution on X = I,
    // all threads of the CTA on the 1024 entries, row t final after step t ----
    for (int b = cta; b < nblk; b += G) {
        const int k = b * NB, nb = (n - k) < NB ? (n - k) : NB;
        __syncthreads();
        for (int e = tid; e < NB * NB; e += kCoopCta) {
            const int i = e % NB, j = e / NB;
            Ls[i][j] = (i < nb && j < nb && j <= i) ? __ldcg(A + (size_t)(k + i) + (size_t)(k + j) * lda) : (i == j ? 1.0 : 0.0);
            Li[i][j] = (i == j) ? 1.0 : 0.0;
        }
        __syncthreads();
        if (tid < NB) sInv[tid] = 1.0 / Ls[tid][tid];
        __syncthreads();
        for (int t = 0; t < NB; ++t) {
            if (tid <= t) Li[t][tid] *= sInv[t];
            __syncthreads();
            for (int e = tid; e < NB * NB; e += kCoopCta) {
                const int i = e % NB, c = e / NB;
                if (i > t && c <= t) Li[i][c] = fma(-Ls[i][t], Li[t][c], Li[i][c]);
            }
            __syncthreads();
        }
        for (int e = tid; e < NB * NB; e += kCoopCta) Dinv32[(size_t)b * NB * NB + e] = Li[e % NB][e / NB];
    }
    grid.sync();
    COOP_TICK(9);
    // ---- inverses of the 128 x 128 diagonal blocks: level 0 (64-blocks from 32-blocks), level 1 (128 from 64).  One job =
    // one 16-column slice of one merge  inv [[A 0] [C B]] = [[A^-1 0] [-B^-1 C A^-1  B^-1]] : 24 jobs per level at n = 765.
    // shared memory (doubles): Bi[h*h] | C[h*h] | Ai slice [h*16] | T slice [h*16]; Dinv block layout: column-major IB x IB ----
    const int nib = (n + IB - 1) / IB;
    constexpr int SL = 16;
    for (int level = 0; level < 2; ++level) {
        const int h = NB << level;                 // order of the inverted halves: 32, then 64
        const int per = IB / (2 * h);              // merges per 128-block: 2, then 1
        const int nsl = h / SL;                    // column slices per merge
        double* sBi = sm; double* sC = sm + h * h; double* sAi = sm + 2 * h * h; double* sT = sAi + h * SL;
        for (int job = cta; job < nib * per * nsl; job += G) {
            const int sl = job % nsl, mq = job / nsl, ib = mq / per, q = mq - ib * per;
            const int o = ib * IB + q * 2 * h;     // first row/col of this merge inside the matrix
            const int lo = q * 2 * h, j0 = sl * SL; // offset inside the 128-block, first column of the slice
            double* D = Dinv + (size_t)ib * IB * IB;
            __syncthreads();
            for (int e = tid; e < h * h; e += kCoopCta) {
                const int i = e % h, j = e / h;
                if (level == 0) {
                    const int bb = (o >> 5) + 1;                 // 32-block index (may lie beyond the matrix: identity)
                    sBi[e] = bb < nblk ? __ldcg(Dinv32 + (size_t)bb * NB * NB + e) : (i == j ? 1.0 : 0.0);
                } else {
                    sBi[e] = __ldcg(D + (size_t)(lo + h + i) + (size_t)(lo + h + j) * IB);
                }
                const int gi = o + h + i, gj = o + j;
                sC[e] = (gi < n && gj < n) ? __ldcg(A + (size_t)gi + (size_t)gj * lda) : 0.0;
            }
            for (int e = tid; e < h * SL; e += kCoopCta) {
                const int i = e % h, j = j0 + e / h;
                if (level == 0) {
                    const int ba = o >> 5;
                    sAi[e] = ba < nblk ? __ldcg(Dinv32 + (size_t)ba * NB * NB + i + NB * j) : (i == j ? 1.0 : 0.0);
                } else {
                    sAi[e] = __ldcg(D + (size_t)(lo + i) + (size_t)(lo + j) * IB);
                }
            }
            __syncthreads();
            for (int e = tid; e < h * SL; e += kCoopCta) {       // T = C * Ai (Ai lower triangular: t >= j)
                const int i = e % h, jj = e / h, j = j0 + jj;
                double s0 = 0.0;
                for (int t = j; t < h; ++t) s0 = fma(sC[i + h * t], sAi[t + h * jj], s0);
                sT[e] = s0;
            }
            __syncthreads();
            for (int e = tid; e < h * SL; e += kCoopCta) {       // Out = -Bi * T (Bi lower triangular: t <= i)
                const int i = e % h, jj = e / h, j = j0 + jj;
                double s0 = 0.0;
                for (int t = 0; t <= i; ++t) s0 = fma(sBi[i + h * t], sT[t + h * jj], s0);
                D[(size_t)(lo + h + i) + (size_t)(lo + j) * IB] = -s0;
                D[(size_t)(lo + i) + (size_t)(lo + h + j) * IB] = 0.0;
                if (level == 0) {
                    D[(size_t)(lo + i) + (size_t)(lo + j) * IB] = sAi[e];
                    D[(size_t)(lo + h + i) + (size_t)(lo + h + j) * IB] = sBi[i + h * j];
                }
            }
        }
        grid.sync();
    }
    COOP_TICK(5);
#undef COOP_TICK
}

// ---- L L^T x = b with the inverted 128 x 128 diagonal blocks, cooperative multi-CTA kernel: the block mat-vec with the
// inverted diagonal block is done by every CTA redundantly (128 x 128, from L2), the update of the remaining right-hand side
// is spread over the CTAs (rows for the forward pass, columns for the backward pass), one device-wide barrier per block step.
// (A single-CTA version reads all of L through one SM at ~50 GB/s: 104 us at n = 765, 1.5 ms at n = 3069.)
__global__ void __launch_bounds__(256) k_potrs_coop(const double* __restrict__ L, int lda, int n, const double* __restrict__ Dinv,
                                                    double* __restrict__ bvec) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ double xs[IB];
    __shared__ double bs[IB];
    __shared__ double part[2][IB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const int nib = (n + IB - 1) / IB;
    for (int pass = 0; pass < 2; ++pass) {
        const bool backward = pass == 1;
        for (int bi = 0; bi < nib; ++bi) {
            const int blk = backward ? (nib - 1 - bi) : bi;
            const int k = blk * IB;
            const int nb = (n - k) < IB ? (n - k) : IB;
            const double* D = Dinv + (size_t)blk * IB * IB;
            if (tid < IB) bs[tid] = tid < nb ? __ldcg(bvec + k + tid) : 0.0;
            __syncthreads();
            if (!backward) {
                // x_r = sum_{c <= r} D[r][c] b_c: thread (r, half) takes every second column, rows contiguous across threads
                const int r = tid & (IB - 1), q = tid >> 7;
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int h = 0; h < 4; ++h) {                   // 4 batches of 16 independent loads (columns q + 2m)
                    double dv[16];
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        const int c = q + 2 * (16 * h + m);
                        dv[m] = c <= r ? D[r + (size_t)c * IB] : 0.0;
                    }
#pragma unroll
                    for (int m = 0; m < 16; m += 2) {
                        s0 = fma(dv[m], bs[q + 2 * (16 * h + m)], s0);
                        s1 = fma(dv[m + 1], bs[q + 2 * (16 * h + m + 1)], s1);
                    }
                }
                part[q][r] = s0 + s1;
                __syncthreads();
                if (tid < IB) xs[tid] = tid < nb ? part[0][tid] + part[1][tid] : 0.0;
            } else {
                // x_r = sum_{c >= r} D[c][r] b_c: warp per row, lanes along the contiguous column r of D
#pragma unroll
                for (int h = 0; h < 2; ++h) {                   // rows warp + 8j: two batches of 8 rows, 32 loads in flight
                    double dv[8][4];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int r = warp + 8 * (8 * h + j);
#pragma unroll
                        for (int m = 0; m < 4; ++m) dv[j][m] = (lane + 32 * m >= r) ? D[lane + 32 * m + (size_t)r * IB] : 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int r = warp + 8 * (8 * h + j);
                        double s = 0.0;
#pragma unroll
                        for (int m = 0; m < 4; ++m) s = fma(dv[j][m], bs[lane + 32 * m], s);
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                        if (lane == 0) xs[r] = r < nb ? s : 0.0;
                    }
                }
            }
            __syncthreads();
            if (cta == 0 && tid < nb) bvec[k + tid] = xs[tid];
            if (!backward) {
                // rows below the block, 64 rows per CTA round-robin; thread (row, quarter of the 128 columns), 32 loads in flight
                const int rr = tid & 63, q = tid >> 6;
                for (int r0 = k + nb + cta * 64; r0 < n; r0 += G * 64) {
                    const int r = r0 + rr;
                    double s = 0.0;
                    if (r < n) {
                        const double* row = L + (size_t)r + (size_t)(k + 32 * q) * lda;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            double v[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = (32 * q + 16 * h + j < nb) ? row[(size_t)(16 * h + j) * lda] : 0.0;
#pragma unroll
                            for (int j = 0; j < 16; ++j) s = fma(v[j], xs[32 * q + 16 * h + j], s);
                        }
                    }
                    __syncthreads();
                    reinterpret_cast<double*>(part)[q * 64 + rr] = s;
                    __syncthreads();
                    if (q == 0 && r < n) {
                        const double* pp = reinterpret_cast<const double*>(part);
                        bvec[r] = __ldcg(bvec + r) - ((pp[rr] + pp[64 + rr]) + (pp[128 + rr] + pp[192 + rr]));
                    }
                }
            } else {
                // columns left of the block, one warp per column round-robin over all warps of the grid; lanes along i
                for (int c = cta * 8 + warp; c < k; c += G * 8) {
                    const double* col = L + (size_t)k + (size_t)c * lda;
                    double v[4];
#pragma unroll
                    for (int m = 0; m < 4; ++m) v[m] = (lane + 32 * m < nb) ? col[lane + 32 * m] : 0.0;
                    double s = 0.0;
#pragma unroll
                    for (int m = 0; m < 4; ++m) s = fma(v[m], xs[lane + 32 * m], s);
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                    if (lane == 0) bvec[c] = __ldcg(bvec + c) - s;
                }
            }
            grid.sync();
        }
    }
}

}  // namespace

size_t dense_coop_dinv_doubles(int n) { return (size_t)((n + IB - 1) / IB) * IB * IB + (size_t)((n + NB - 1) / NB) * NB * NB + 16; }

int dense_potrf_coop(ptzba_ctx* ctx, double* A, int n, int lda, int* d_info, double* Dinv_store) {
    if (n <= 0) return PTZBA_OK;
    const size_t smem = (size_t)4 * TS * TS * sizeof(double);      // 128 KB: four 64 x 64 blocks of the doubling level
    if (!ctx->coop_configured) {            // per context: function attributes belong to the device the context runs on
        CU_CHECK(ctx, cudaFuncSetAttribute(k_potrf_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->coop_configured = true;
    }
    const int nib = (n + IB - 1) / IB;
    double* Dinv = Dinv_store;
    double* Dinv32 = Dinv_store + (size_t)nib * IB * IB;
    int grid = ctx->sm_count;
    long long* dbg = nullptr;
    static const bool trace = getenv("PTZBA_TRACE") != nullptr;
    if (trace) {
        CU_CHECK(ctx, cudaMalloc((void**)&dbg, 64 * sizeof(long long)));
        CU_CHECK(ctx, cudaMemsetAsync(dbg, 0, 64 * sizeof(long long), ctx->stream));
    }
    void* args[] = {&A, &n, &lda, &d_info, &Dinv32, &Dinv, &dbg};
    CU_CHECK(ctx, cudaLaunchCooperativeKernel((const void*)k_potrf_coop, dim3(grid), dim3(kCoopCta), args, smem, ctx->stream));
    ctx->launches++;
    if (trace) {
        long long h[64];
        CU_CHECK(ctx, cudaMemcpyAsync(h, dbg, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        fprintf(stderr, "[ptzba trace] potrf_coop n=%d cycles: first_block=%lld panel=%lld bar=%lld update|lookahead=%lld bar=%lld inv32=%lld blockinv=%lld | helper: wait_for_block=%lld potf2=%lld\n", n,
                h[0], h[1], h[2], h[3], h[4], h[9], h[5], h[7], h[8]);
        fprintf(stderr, "[ptzba trace] potf2 per panel:");
        for (int i = 0; i < 24; ++i) fprintf(stderr, " %lld", h[16 + i]);
        fprintf(stderr, "\n");
        cudaFree(dbg);
    }
    return PTZBA_OK;
}

int dense_potrs_coop(ptzba_ctx* ctx, const double* L, int n, int lda, const double* Dinv_store, double* b) {
    if (n <= 0) return PTZBA_OK;
    // enough CTAs that every block step's update is one round (64 rows or 8 columns per CTA), at most one per SM
    int grid = (n + 63) / 64;
    if (grid > ctx->sm_count) grid = ctx->sm_count;
    if (grid < 1) grid = 1;
    void* args[] = {&L, &lda, &n, &Dinv_store, &b};
    CU_CHECK(ctx, cudaLaunchCooperativeKernel((const void*)k_potrs_coop, dim3(grid), dim3(256), args, 0, ctx->stream));
    ctx->launches++;
    return PTZBA_OK;
}

// Solves A x = b for a symmetric positive definite A (host arrays, column- or row-major alike, only the lower triangle is read)
// with the cooperative Cholesky + block-inverse solve that bundle adjustment uses for its reduced camera system.
// *info = 0, or 1 + the first row of the panel where the factorisation met a non-positive pivot.
extern "C" int ptzba_dense_solve_spd(ptzba_ctx* ctx, int n, const double* A, const double* b, double* x, int* info) {
    if (!ctx) return PTZBA_ERR_ARG;
    ARG_CHECK(ctx, n >= 1 && A && b && x && info);
    cudaStream_t s = ctx->stream;
    DevBuf<double> dA, db, dinv;
    DevBuf<int> dinfo;
    CU_CHECK(ctx, dA.alloc((size_t)n * n));
    CU_CHECK(ctx, db.alloc(n));
    CU_CHECK(ctx, dinv.alloc(dense_coop_dinv_doubles(n)));
    CU_CHECK(ctx, dinfo.alloc(1));
    CU_CHECK(ctx, cudaMemcpyAsync(dA.p, A, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, s));
    CU_CHECK(ctx, cudaMemcpyAsync(db.p, b, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
    CU_CHECK(ctx, cudaMemsetAsync(dinfo.p, 0, sizeof(int), s));
    PROPAGATE(dense_potrf_coop(ctx, dA.p, n, n, dinfo.p, dinv.p));
    PROPAGATE(dense_potrs_coop(ctx, dA.p, n, n, dinv.p, db.p));
    CU_CHECK(ctx, cudaMemcpyAsync(x, db.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
    CU_CHECK(ctx, cudaMemcpyAsync(info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    return PTZBA_OK;
}
