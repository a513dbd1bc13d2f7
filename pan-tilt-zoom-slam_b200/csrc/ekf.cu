// ekf.cu - batched EKF predict / update for ray-landmark PTZ tracking.
//
// Reference behaviour (slam_system/ptz_slam.py):
//   predict   :418-426   ptz += velocity ; P[0:3,0:3] += 5*diag(angle_var, angle_var, f_var)
//   update    :210-289   project all rays with the predicted pose, keep strict in-image ones, intersect with the observed
//                        ids (util.get_overlap_index :75-96); y = z - h ; gather P_s ; H from compute_h_jacobian (:73-138) ;
//                        S = H P_s H^T + observe_var I ; K = P_s H^T S^-1 ; pose/velocity/rays += K y ;
//                        P+ = (I - K H) P_s written back ONLY on the pose block and the theta-theta / phi-phi entries (:281-289).
//
// Device algorithm per sequence (n matched rays, s = 3 + 2n, n2 = 2n), P treated as symmetric:
//   G = H P_s (n2 x s, columns de-interleaved [pose | theta_1..n | phi_1..n], plus y as an extra column)
//   S = G H^T + R          -> LU with partial pivoting (dense_lu.cu, batched).  S is symmetric but NOT positive definite in
//                             general: the reference's write-back drops cross terms, P turns indefinite after a few frames and
//                             numpy.linalg.inv (getrf) keeps going where a Cholesky would break down.
//   X = S^-1 [G | y]       -> delta = G^T x_y (x_y = last column), P+ = P_s - G^T X on the blocks the reference writes back
// Many independent sequences are processed as one batch (grid z / y = sequence), in waves that share the G / S workspace.
// FP64-pipe bound: ~ (n2^3/3 + n2^2 s + n^2 n2) flops per sequence-frame.
#include <algorithm>
#include <vector>

#include "common.h"
#include "dense.h"
#include "dense_mma.cuh"
#include "ptz_jac.cuh"
#include "ptz_math.cuh"

struct ptzba_ekf_batch {
    ptzba_ctx* ctx = nullptr;
    ptzba_ekf_params prm;
    // n_ray = CAPACITY (stride) of the per-sequence ray / covariance arrays; sequence b currently uses its first n_act[b] rays
    // (ptzba_ekf_batch_add_rays / _remove_rays change that count on the device; the capacity grows on demand)
    int n_seq = 0, n_ray = 0, max_obs = 0, s_tot = 0;
    bool has_disp = false;
    std::vector<int32_t> h_n_act;
    DevBuf<int32_t> n_act;
    // sequences whose innovation covariance was indefinite once go straight to the pivoted-LU route afterwards (P stays indefinite)
    std::vector<char> h_use_lu;
    DevBuf<int32_t> use_lu;
    DevBuf<double> rays, P, ptz, vel, disp;
    // observations of the current step
    DevBuf<double> obs_xy;
    DevBuf<int32_t> obs_idx, obs_cnt;
    // per-step results
    DevBuf<int32_t> n_mat, n2, m_ray, flags;
    DevBuf<double> Jc, Jr, y;
    // wave workspace
    int wave = 0, ldg = 0, lds = 0;
    long n_lu_total = 0;               // waves that needed the LU route (diagnostic)
    DevBuf<double> G, X, S, dpart;
    DevBuf<int32_t> ipiv, perm;
    DevBuf<int32_t> chol_fail, nm_chol, n2_chol, nm_lu, n2_lu;   // per-sequence path selection (Cholesky first, LU on breakdown)
    // pinned read-back block: [n_seq matched | n_seq chol_fail | 8 flags]
    int32_t* h_back = nullptr;
    // ray bookkeeping scratch (remove_rays): index map and one covariance-sized buffer
    DevBuf<int32_t> idx_map;
    DevBuf<double> P_scratch;
    ~ptzba_ekf_batch() { if (h_back) cudaFreeHost(h_back); }
};

namespace {

constexpr int kT = 256;

__global__ void k_ekf_predict(int n_seq, double* __restrict__ ptz, const double* __restrict__ vel, double* __restrict__ P,
                              size_t strideP, int s_tot, double angle_var, double f_var) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_seq) return;
    for (int e = 0; e < 3; ++e) ptz[3 * b + e] += vel[3 * b + e];            // ptz_slam.py:418-419
    double* p = P + strideP * b;
    p[0] += 5.0 * angle_var;                                                  // ptz_slam.py:425-426
    p[(size_t)s_tot + 1] += 5.0 * angle_var;
    p[2 * (size_t)s_tot + 2] += 5.0 * f_var;
}

__global__ void k_ekf_predict_cov(int n_seq, double* __restrict__ P, size_t strideP, int s_tot, double angle_var, double f_var) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_seq) return;
    double* p = P + strideP * b;
    p[0] += 5.0 * angle_var;                                                  // ptz_slam.py:425-426
    p[(size_t)s_tot + 1] += 5.0 * angle_var;
    p[2 * (size_t)s_tot + 2] += 5.0 * f_var;
}

// one CTA per sequence: in-image test of the observed rays, ordered compaction, innovation and Jacobian blocks
__global__ void __launch_bounds__(kT) k_ekf_match(int n_ray, const int32_t* __restrict__ n_act, int max_obs, const double* __restrict__ ptz_all,
                                                  const double* __restrict__ rays_all, const double* __restrict__ disp,
                                                  ptzba_ekf_params prm, const double* __restrict__ obs_xy_all,
                                                  const int32_t* __restrict__ obs_idx_all, const int32_t* __restrict__ obs_cnt,
                                                  int32_t* __restrict__ n_mat, int32_t* __restrict__ n2,
                                                  int32_t* __restrict__ m_ray_all, double* __restrict__ y_all,
                                                  double* __restrict__ Jc_all, double* __restrict__ Jr_all,
                                                  int* __restrict__ flags) {
    const int b = blockIdx.x;
    __shared__ CamFull cams[7];
    __shared__ int warp_cnt[kT / 32];
    __shared__ int carry;
    const double* ptz = ptz_all + 3 * (size_t)b;
    if (threadIdx.x < 7) cams[threadIdx.x] = h_cam_variant(threadIdx.x, ptz[0], ptz[1], ptz[2], prm.u, prm.v, disp);
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const double* rays = rays_all + 2 * (size_t)n_ray * b;
    const double* oxy = obs_xy_all + 2 * (size_t)max_obs * b;
    const int32_t* oidx = obs_idx_all + (size_t)max_obs * b;
    int32_t* m_ray = m_ray_all + (size_t)max_obs * b;
    double* y = y_all + 2 * (size_t)max_obs * b;
    double* Jc = Jc_all + 6 * (size_t)max_obs * b;
    double* Jr = Jr_all + 4 * (size_t)max_obs * b;
    int cnt = obs_cnt[b];
    if (cnt > max_obs) { cnt = max_obs; if (threadIdx.x == 0) atomicOr(flags, 1); }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < cnt; base += kT) {
        const int k = base + threadIdx.x;
        bool keep = false;
        int r = -1;
        double px = 0, py = 0, th = 0, ph = 0;
        if (k < cnt) {
            r = oidx[k];
            if (r < 0 || r >= n_act[b]) {
                atomicOr(flags, 2);
            } else {
                th = rays[2 * (size_t)r];
                ph = rays[2 * (size_t)r + 1];
                double q;
                project_full(cams[0], th, ph, px, py, q);
                keep = (0.0 < px) && (px < prm.width) && (0.0 < py) && (py < prm.height);   // ptz_camera.py:226
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[w] = __popc(m);
        __syncthreads();
        int off = carry;
        for (int q = 0; q < w; ++q) off += warp_cnt[q];
        if (keep) {
            const int j = off + __popc(m & ((1u << lane) - 1u));
            m_ray[j] = r;
            y[2 * j] = oxy[2 * (size_t)k] - px;                                             // ptz_slam.py:229
            y[2 * j + 1] = oxy[2 * (size_t)k + 1] - py;
            double jc[6], jr[4];
            h_blocks_eval(cams, disp, th, ph, prm.jac_mode, jc, jr);
#pragma unroll
            for (int e = 0; e < 6; ++e) Jc[6 * (size_t)j + e] = jc[e];
#pragma unroll
            for (int e = 0; e < 4; ++e) Jr[4 * (size_t)j + e] = jr[e];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int s = 0;
            for (int q = 0; q < kT / 32; ++q) s += warp_cnt[q];
            carry += s;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { n_mat[b] = carry; n2[b] = 2 * carry; }
}

// column c of the de-interleaved sub-state -> row/column index in the full interleaved covariance
__device__ __forceinline__ int full_index(int c, int n, const int32_t* __restrict__ m_ray) {
    if (c < 3) return c;
    if (c < 3 + n) return 3 + 2 * m_ray[c - 3];
    return 4 + 2 * m_ray[c - 3 - n];
}

// G[2j+a][c] = sum_b Jc[j][a][b] P[b][pc] + Jr[j][a][0] P[3+2r_j][pc] + Jr[j][a][1] P[4+2r_j][pc] ; column s holds y
__global__ void __launch_bounds__(128) k_ekf_G(int b0, int max_obs, int s_tot, const double* __restrict__ P_all, size_t strideP,
                                               const int32_t* __restrict__ n_mat, const int32_t* __restrict__ m_ray_all,
                                               const double* __restrict__ Jc_all, const double* __restrict__ Jr_all,
                                               const double* __restrict__ y_all, double* __restrict__ G_all, int ldg,
                                               size_t strideG) {
    const int wb = blockIdx.z, b = b0 + wb;
    const int n = n_mat[b];
    const int j = blockIdx.y;
    if (j >= n) return;
    const int s = 3 + 2 * n;
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c > s) return;
    const int32_t* m_ray = m_ray_all + (size_t)max_obs * b;
    double* G = G_all + strideG * wb;
    if (c == s) {
        G[(size_t)(2 * j) * ldg + c] = y_all[2 * (size_t)max_obs * b + 2 * j];
        G[(size_t)(2 * j + 1) * ldg + c] = y_all[2 * (size_t)max_obs * b + 2 * j + 1];
        return;
    }
    const double* P = P_all + strideP * b;
    const int pc = full_index(c, n, m_ray);
    const int rj = 3 + 2 * m_ray[j];
    const double p0 = P[pc], p1 = P[(size_t)s_tot + pc], p2 = P[2 * (size_t)s_tot + pc];
    const double pt = P[(size_t)rj * s_tot + pc], pp = P[(size_t)(rj + 1) * s_tot + pc];
    const double* jc = Jc_all + 6 * ((size_t)max_obs * b + j);
    const double* jr = Jr_all + 4 * ((size_t)max_obs * b + j);
    G[(size_t)(2 * j) * ldg + c] = jc[0] * p0 + jc[1] * p1 + jc[2] * p2 + jr[0] * pt + jr[1] * pp;
    G[(size_t)(2 * j + 1) * ldg + c] = jc[3] * p0 + jc[4] * p1 + jc[5] * p2 + jr[2] * pt + jr[3] * pp;
}

// S = G H^T + observe_var I.  Thread (i, k) computes S[i][2k], S[i][2k+1] and stores them at (row 2k+a, column i) of the
// column-major array the Cholesky reads (S is symmetric), i.e. contiguous along k.
__global__ void __launch_bounds__(128) k_ekf_S(int b0, int max_obs, const int32_t* __restrict__ n_mat,
                                               const double* __restrict__ Jc_all, const double* __restrict__ Jr_all,
                                               const double* __restrict__ G_all, int ldg, size_t strideG,
                                               double* __restrict__ S_all, int lds, size_t strideS, double observe_var) {
    const int wb = blockIdx.z, b = b0 + wb;
    const int n = n_mat[b];
    const int i = blockIdx.y;
    if (i >= 2 * n) return;
    const int k = blockIdx.x * 128 + threadIdx.x;
    if (k >= n) return;
    const double* G = G_all + strideG * wb + (size_t)i * ldg;
    const double* jc = Jc_all + 6 * ((size_t)max_obs * b + k);
    const double* jr = Jr_all + 4 * ((size_t)max_obs * b + k);
    const double g0 = G[0], g1 = G[1], g2 = G[2], gt = G[3 + k], gp = G[3 + n + k];
    double s0 = g0 * jc[0] + g1 * jc[1] + g2 * jc[2] + gt * jr[0] + gp * jr[1];
    double s1 = g0 * jc[3] + g1 * jc[4] + g2 * jc[5] + gt * jr[2] + gp * jr[3];
    if (i == 2 * k) s0 += observe_var;                                                      // ptz_slam.py:256-257
    if (i == 2 * k + 1) s1 += observe_var;
    double* S = S_all + strideS * wb + (size_t)i * lds;
    S[2 * k] = s0;
    S[2 * k + 1] = s1;
}

// delta = G^T x_y (x_y = last column of X = S^-1 [G | y]) in two deterministic stages: partial sums over chunks of 64 rows
// (grid: column tiles x row chunks x sequences; the round-1 kernel ran one serial 2n-long dot product per column), then the
// chunks are added in order and applied to pose / velocity / rays (ptz_slam.py:262-277)
constexpr int kDeltaRows = 64;
__global__ void __launch_bounds__(kT) k_ekf_delta_partial(int b0, const int32_t* __restrict__ n_mat, const double* __restrict__ G_all,
                                                          const double* __restrict__ X_all, int ldg, size_t strideG,
                                                          double* __restrict__ part_all, int n_chunk_max) {
    const int wb = blockIdx.z, b = b0 + wb;
    const int n = n_mat[b];
    const int s = 3 + 2 * n;
    const int c = blockIdx.x * kT + threadIdx.x;
    const int i0 = blockIdx.y * kDeltaRows;
    if (c >= s || i0 >= 2 * n) return;
    const double* G = G_all + strideG * wb;
    const double* X = X_all + strideG * wb;
    const int i1 = min(i0 + kDeltaRows, 2 * n);
    double a0 = 0.0, a1 = 0.0;
    int i = i0;
    for (; i + 1 < i1; i += 2) {
        a0 = fma(G[(size_t)i * ldg + c], X[(size_t)i * ldg + s], a0);
        a1 = fma(G[(size_t)(i + 1) * ldg + c], X[(size_t)(i + 1) * ldg + s], a1);
    }
    if (i < i1) a0 = fma(G[(size_t)i * ldg + c], X[(size_t)i * ldg + s], a0);
    part_all[((size_t)wb * n_chunk_max + blockIdx.y) * ldg + c] = a0 + a1;
}

__global__ void __launch_bounds__(kT) k_ekf_delta_apply(int b0, int max_obs, int n_ray, const int32_t* __restrict__ n_mat,
                                                        const int32_t* __restrict__ m_ray_all, const double* __restrict__ part_all,
                                                        int ldg, int n_chunk_max, double* __restrict__ ptz, double* __restrict__ vel,
                                                        double* __restrict__ rays_all) {
    const int wb = blockIdx.y, b = b0 + wb;
    const int n = n_mat[b];
    const int s = 3 + 2 * n;
    const int c = blockIdx.x * kT + threadIdx.x;
    if (c >= s || n == 0) return;
    const int n_chunk = (2 * n + kDeltaRows - 1) / kDeltaRows;
    double acc = 0.0;
    for (int q = 0; q < n_chunk; ++q) acc += part_all[((size_t)wb * n_chunk_max + q) * ldg + c];
    if (c < 3) {
        ptz[3 * (size_t)b + c] += acc;
        vel[3 * (size_t)b + c] = acc;
    } else {
        const int32_t* m_ray = m_ray_all + (size_t)max_obs * b;
        double* rays = rays_all + 2 * (size_t)n_ray * b;
        if (c < 3 + n) rays[2 * (size_t)m_ray[c - 3]] += acc;
        else rays[2 * (size_t)m_ray[c - 3 - n] + 1] += acc;
    }
}

// pose block: P[a][b] -= sum_i G[i][a] X[i][b]   (one warp per sequence)
__global__ void __launch_bounds__(32) k_ekf_pp_pose(int b0, const int32_t* __restrict__ n_mat, const double* __restrict__ G_all,
                                                    const double* __restrict__ X_all, int ldg, size_t strideG,
                                                    double* __restrict__ P_all, size_t strideP, int s_tot) {
    const int wb = blockIdx.x, b = b0 + wb;
    const int n2 = 2 * n_mat[b];
    const double* G = G_all + strideG * wb;
    const double* X = X_all + strideG * wb;
    double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < n2; i += 32) {
        const double g0 = G[(size_t)i * ldg], g1 = G[(size_t)i * ldg + 1], g2 = G[(size_t)i * ldg + 2];
        const double x0 = X[(size_t)i * ldg], x1 = X[(size_t)i * ldg + 1], x2 = X[(size_t)i * ldg + 2];
        a[0] = fma(g0, x0, a[0]); a[1] = fma(g0, x1, a[1]); a[2] = fma(g0, x2, a[2]);
        a[3] = fma(g1, x0, a[3]); a[4] = fma(g1, x1, a[4]); a[5] = fma(g1, x2, a[5]);
        a[6] = fma(g2, x0, a[6]); a[7] = fma(g2, x1, a[7]); a[8] = fma(g2, x2, a[8]);
    }
#pragma unroll
    for (int e = 0; e < 9; ++e)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a[e] += __shfl_xor_sync(0xffffffffu, a[e], off);
    if (threadIdx.x < 9 && n2 > 0) {
        double* P = P_all + strideP * b;
        double val = a[0];
#pragma unroll
        for (int e = 1; e < 9; ++e) if (threadIdx.x == e) val = a[e];
        P[(size_t)(threadIdx.x / 3) * s_tot + threadIdx.x % 3] -= val;
    }
}

// theta-theta (blockIdx.y = 0) and phi-phi (1) blocks: P[q_j][q_k] -= sum_i G[i][off+j] X[i][off+k], 64 x 64 tiles on the FP64
// tensor cores (dense_mma.cuh), K = all n2 rows.  G^T S^-1 G is symmetric, so only the lower tile pairs are formed and mirrored
// on write.
__global__ void __launch_bounds__(dmma::kThreads) k_ekf_pp_blocks(int b0, int max_obs, const int32_t* __restrict__ n_mat,
                                                                 const int32_t* __restrict__ m_ray_all, const double* __restrict__ G_all,
                                                                 const double* __restrict__ X_all, int ldg, size_t strideG,
                                                                 double* __restrict__ P_all, size_t strideP, int s_tot) {
    using T = dmma::Tile<64, 64>;
    extern __shared__ __align__(16) double sm[];
    const int wb = blockIdx.z, b = b0 + wb;
    const int n = n_mat[b];
    const int nt = (n + 63) / 64;
    int blk = blockIdx.x;
    int ti = (int)((sqrt(8.0 * blk + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= blk) ++ti;
    while (ti * (ti + 1) / 2 > blk) --ti;
    const int tj = blk - ti * (ti + 1) / 2;
    if (ti >= nt) return;
    const int e = blockIdx.y;                       // 0 theta, 1 phi
    const int off = 3 + e * n;
    const double* G = G_all + strideG * wb;
    const double* X = X_all + strideG * wb;
    const int j0 = ti * 64, k0 = tj * 64;
    double acc[T::RM][T::RN][2];
#pragma unroll
    for (int a = 0; a < T::RM; ++a)
#pragma unroll
        for (int c = 0; c < T::RN; ++c) acc[a][c][0] = acc[a][c][1] = 0.0;
    T::accumulate(G + off + j0, (size_t)ldg, n - j0, X + off + k0, (size_t)ldg, n - k0, 2 * n, acc, sm);
    const int32_t* m_ray = m_ray_all + (size_t)max_obs * b;
    double* P = P_all + strideP * b;
    T::for_each(acc, [&](int jj, int kk, double v) {
        const int j = j0 + jj, k = k0 + kk;
        if (j < n && k < n) {
            const size_t qj = 3 + e + 2 * (size_t)m_ray[j], qk = 3 + e + 2 * (size_t)m_ray[k];
            if (ti != tj) {
                P[qj * s_tot + qk] -= v;
                P[qk * s_tot + qj] -= v;
            } else if (j >= k) {
                P[qj * s_tot + qk] -= v;
                if (j != k) P[qk * s_tot + qj] -= v;
            }
        }
    });
}


// path selection BEFORE the factorisation: sequences flagged use_lu skip the Cholesky attempt (n = 0 on that path, which
// makes every batched kernel skip them)
__global__ void k_ekf_route(int n_seq, const int32_t* __restrict__ use_lu, const int32_t* __restrict__ n_mat,
                            int32_t* __restrict__ nm_chol, int32_t* __restrict__ n2_chol, int32_t* __restrict__ nm_lu,
                            int32_t* __restrict__ n2_lu) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_seq) return;
    const int n = n_mat[b];
    const bool f = use_lu[b] != 0;
    nm_chol[b] = f ? 0 : n; n2_chol[b] = f ? 0 : 2 * n;
    nm_lu[b] = f ? n : 0;   n2_lu[b] = f ? 2 * n : 0;
}

// path selection AFTER the Cholesky attempt: sequences whose S was positive definite finish on the Cholesky path, the others
// are redone with pivoted LU and keep that route in later frames
__global__ void k_ekf_split(int n_seq, const int32_t* __restrict__ fail, const int32_t* __restrict__ n_mat,
                            int32_t* __restrict__ nm_chol, int32_t* __restrict__ n2_chol, int32_t* __restrict__ nm_lu,
                            int32_t* __restrict__ n2_lu, int32_t* __restrict__ use_lu) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_seq) return;
    if (fail[b] == 0) return;
    const int n = n_mat[b];
    nm_chol[b] = 0; n2_chol[b] = 0;
    nm_lu[b] = n;   n2_lu[b] = 2 * n;
    use_lu[b] = 1;
}

// X = G for the rows / columns in use (row-major, n2 rows, n2 + extra columns)
__global__ void __launch_bounds__(256) k_copy_rows(const double* __restrict__ G0, double* __restrict__ X0, int ldg, size_t strideG,
                                                   const int32_t* __restrict__ n2_arr, int extra_cols) {
    const int b = blockIdx.z;
    const int n = n2_arr[b];
    const int i = blockIdx.y;
    if (i >= n) return;
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= n + extra_cols) return;
    X0[strideG * b + (size_t)i * ldg + c] = G0[strideG * b + (size_t)i * ldg + c];
}

__global__ void k_ekf_init_cov(int n_seq, int s_tot, double* __restrict__ P, size_t strideP, double angle_var, double f_var) {
    // state_cov = angle_var * I ; [2][2] = f_var   (ptz_slam.py:199-200); P is pre-zeroed
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_seq || i >= s_tot) return;
    P[strideP * b + (size_t)i * s_tot + i] = (i == 2) ? f_var : angle_var;
}

// ---- ray bookkeeping on the device (PtzSlam.remove_rays / add_rays, ptz_slam.py:291-388) ---------------------------------------
// compacted copy: dst[i][j] = src[map[i]][map[j]] for i, j < s_new (both with leading dimension ld)
__global__ void k_ekf_compact_cov(int s_new, int ld, const int32_t* __restrict__ map, const double* __restrict__ src, double* __restrict__ dst) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= s_new) return;
    dst[(size_t)i * ld + j] = src[(size_t)map[i] * ld + map[j]];
}
__global__ void k_ekf_compact_rays(int n_new, const int32_t* __restrict__ keep, const double* __restrict__ src, double* __restrict__ dst) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_new) return;
    dst[2 * r] = src[2 * (size_t)keep[r]];
    dst[2 * r + 1] = src[2 * (size_t)keep[r] + 1];
}
// rows / columns [s_old, s_new) of P: zero, angle_var on the diagonal (ptz_slam.py:378-384)
__global__ void k_ekf_grow_cov(int s_old, int s_new, int ld, double* __restrict__ P, double angle_var) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= s_new || (i < s_old && j < s_old)) return;
    P[(size_t)i * ld + j] = (i == j) ? angle_var : 0.0;
}

int ekf_alloc_work(ptzba_ekf_batch* B) {
    ptzba_ctx* ctx = B->ctx;
    const int n_seq = B->n_seq, max_obs = B->max_obs;
    B->ldg = (3 + 2 * max_obs + 1 + 3) / 4 * 4;
    B->lds = 2 * max_obs > 0 ? 2 * max_obs : 1;
    // wave size: share at most ~24 GB of G/S workspace
    const size_t per_seq = (2 * (size_t)B->ldg * 2 * (size_t)max_obs + (size_t)B->lds * B->lds) * sizeof(double);
    size_t wave = per_seq ? (size_t)24e9 / per_seq : (size_t)n_seq;
    if (wave < 1) wave = 1;
    B->wave = (int)std::min<size_t>(wave, (size_t)n_seq);
    CU_CHECK(ctx, B->obs_xy.alloc((size_t)n_seq * max_obs * 2)); CU_CHECK(ctx, B->obs_idx.alloc((size_t)n_seq * max_obs));
    CU_CHECK(ctx, B->m_ray.alloc((size_t)n_seq * max_obs));
    CU_CHECK(ctx, B->Jc.alloc((size_t)n_seq * max_obs * 6)); CU_CHECK(ctx, B->Jr.alloc((size_t)n_seq * max_obs * 4));
    CU_CHECK(ctx, B->y.alloc((size_t)n_seq * max_obs * 2));
    CU_CHECK(ctx, B->G.alloc((size_t)B->wave * B->ldg * 2 * (size_t)max_obs)); CU_CHECK(ctx, B->X.alloc((size_t)B->wave * B->ldg * 2 * (size_t)max_obs));
    CU_CHECK(ctx, B->S.alloc((size_t)B->wave * B->lds * B->lds));
    CU_CHECK(ctx, B->dpart.alloc((size_t)B->wave * (div_up(2 * max_obs, kDeltaRows) + 1) * B->ldg));
    CU_CHECK(ctx, B->ipiv.alloc((size_t)B->wave * B->lds)); CU_CHECK(ctx, B->perm.alloc((size_t)B->wave * B->lds));
    return PTZBA_OK;
}

// grows the per-sequence capacity (rays, covariance with its leading dimension) and / or the per-step observation capacity
int ekf_reserve(ptzba_ekf_batch* B, int cap_ray, int max_obs) {
    ptzba_ctx* ctx = B->ctx;
    cudaStream_t s = ctx->stream;
    if (cap_ray > B->n_ray) {
        const int s_old = B->s_tot, s_new = 3 + 2 * cap_ray;
        DevBuf<double> rays2, P2;
        CU_CHECK(ctx, rays2.alloc((size_t)B->n_seq * cap_ray * 2));
        CU_CHECK(ctx, P2.alloc((size_t)B->n_seq * s_new * s_new));
        CU_CHECK(ctx, cudaMemsetAsync(P2.p, 0, (size_t)B->n_seq * s_new * s_new * sizeof(double), s));
        for (int b = 0; b < B->n_seq; ++b) {
            const int na = B->h_n_act[b], sa = 3 + 2 * na;
            if (na) CU_CHECK(ctx, cudaMemcpyAsync(rays2.p + 2 * (size_t)cap_ray * b, B->rays.p + 2 * (size_t)B->n_ray * b, (size_t)na * 2 * sizeof(double),
                                                  cudaMemcpyDeviceToDevice, s));
            CU_CHECK(ctx, cudaMemcpy2DAsync(P2.p + (size_t)s_new * s_new * b, (size_t)s_new * sizeof(double), B->P.p + (size_t)s_old * s_old * b,
                                            (size_t)s_old * sizeof(double), (size_t)sa * sizeof(double), sa, cudaMemcpyDeviceToDevice, s));
        }
        CU_CHECK(ctx, cudaStreamSynchronize(s));
        std::swap(B->rays.p, rays2.p); std::swap(B->rays.n, rays2.n);
        std::swap(B->P.p, P2.p); std::swap(B->P.n, P2.n);
        B->n_ray = cap_ray; B->s_tot = s_new;
        B->P_scratch.release();
    }
    if (max_obs > B->n_ray) max_obs = B->n_ray;
    if (max_obs > B->max_obs) {
        B->max_obs = max_obs;
        PROPAGATE(ekf_alloc_work(B));
    }
    return PTZBA_OK;
}

// One predict + update (or update only) for every sequence.  h_obs_count (host copy of the per-sequence observation counts, or
// nullptr) bounds the launch grids; the matched counts themselves stay on the device, so the step has ONE host synchronisation
// per wave (verdicts: argument flags, Cholesky breakdown per sequence, matched counts) - the round-1 version had three.
int ekf_step(ptzba_ekf_batch* B, bool do_predict, int32_t* out_matched, const int32_t* h_obs_count) {
    ptzba_ctx* ctx = B->ctx;
    cudaStream_t s = ctx->stream;
    const int n_seq = B->n_seq, max_obs = B->max_obs, s_tot = B->s_tot;
    const size_t strideP = (size_t)s_tot * s_tot;
    static bool cfg = false;
    if (!cfg) { CU_CHECK(ctx, dmma::configure(k_ekf_pp_blocks, dmma::Tile<64, 64>::kSmemBytes)); cfg = true; }
    if (do_predict) {
        k_ekf_predict<<<div_up(n_seq, 128), 128, 0, s>>>(n_seq, B->ptz.p, B->vel.p, B->P.p, strideP, s_tot, B->prm.angle_var,
                                                        B->prm.f_var);
        KERNEL_POST(ctx);
    }
    CU_CHECK(ctx, cudaMemsetAsync(B->flags.p, 0, 8 * sizeof(int), s));
    k_ekf_match<<<n_seq, kT, 0, s>>>(B->n_ray, B->n_act.p, max_obs, B->ptz.p, B->rays.p, B->has_disp ? B->disp.p : nullptr, B->prm,
                                     B->obs_xy.p, B->obs_idx.p, B->obs_cnt.p, B->n_mat.p, B->n2.p, B->m_ray.p, B->y.p, B->Jc.p,
                                     B->Jr.p, B->flags.p);
    KERNEL_POST(ctx);
    int32_t* h_n = B->h_back;
    int32_t* h_fail = B->h_back + n_seq;
    int32_t* h_flags = B->h_back + 2 * n_seq;
    const size_t strideG = (size_t)B->ldg * (2 * (size_t)max_obs), strideS = (size_t)B->lds * B->lds;
    const int n_chunk_max = div_up(2 * max_obs, kDeltaRows) + 1;
    for (int b0 = 0; b0 < n_seq; b0 += B->wave) {
        const int wb = std::min(B->wave, n_seq - b0);
        int n_max = 0;                                   // upper bound of the matched rays of any sequence of the wave
        for (int q = 0; q < wb; ++q) n_max = std::max(n_max, h_obs_count ? std::min((int)h_obs_count[b0 + q], max_obs) : max_obs);
        bool any_chol = false, any_lu = false;
        for (int q = 0; q < wb; ++q) { any_chol = any_chol || !B->h_use_lu[b0 + q]; any_lu = any_lu || B->h_use_lu[b0 + q]; }
        if (n_max > 0) {
            const int s_max = 3 + 2 * n_max;
            k_ekf_route<<<div_up(wb, 128), 128, 0, s>>>(wb, B->use_lu.p + b0, B->n_mat.p + b0, B->nm_chol.p + b0, B->n2_chol.p + b0,
                                                       B->nm_lu.p + b0, B->n2_lu.p + b0);
            KERNEL_POST(ctx);
            k_ekf_G<<<dim3(div_up(s_max + 1, 128), n_max, wb), 128, 0, s>>>(b0, max_obs, s_tot, B->P.p, strideP, B->n_mat.p, B->m_ray.p,
                                                                           B->Jc.p, B->Jr.p, B->y.p, B->G.p, B->ldg, strideG);
            KERNEL_POST(ctx);
            CU_CHECK(ctx, cudaMemsetAsync(B->chol_fail.p + b0, 0, (size_t)wb * sizeof(int32_t), s));
        }
        if (n_max > 0 && any_chol) {
            const int s_max = 3 + 2 * n_max;
            k_ekf_S<<<dim3(div_up(n_max, 128), 2 * n_max, wb), 128, 0, s>>>(b0, max_obs, B->nm_chol.p, B->Jc.p, B->Jr.p, B->G.p, B->ldg,
                                                                           strideG, B->S.p, B->lds, strideS, B->prm.observe_var);
            KERNEL_POST(ctx);
            // ---- Cholesky first: S is positive definite in the first frames of a sequence; pivoted LU where it breaks down ----
            PROPAGATE(dense_potrf_lower_batched(ctx, B->S.p, B->lds, strideS, B->n2_chol.p + b0, 2 * n_max, wb, B->flags.p + 3,
                                                B->chol_fail.p + b0));
            k_ekf_split<<<div_up(wb, 128), 128, 0, s>>>(wb, B->chol_fail.p + b0, B->n_mat.p + b0, B->nm_chol.p + b0, B->n2_chol.p + b0,
                                                       B->nm_lu.p + b0, B->n2_lu.p + b0, B->use_lu.p + b0);
            KERNEL_POST(ctx);
            // Z = L^-1 [G | y] in X ; delta = Z^T z_y ; P+ = P - Z^T Z   (same kernels with G := X := Z); sequences whose Cholesky
            // broke down have n = 0 on this path and are skipped by every kernel
            k_copy_rows<<<dim3(div_up(s_max + 1, 256), 2 * n_max, wb), 256, 0, s>>>(B->G.p, B->X.p, B->ldg, strideG, B->n2_chol.p + b0, 4);
            KERNEL_POST(ctx);
            PROPAGATE(dense_fwd_solve_rows_batched(ctx, B->S.p, B->lds, strideS, B->X.p, B->ldg, strideG, B->n2_chol.p + b0,
                                                   2 * n_max, 4, wb));
            k_ekf_delta_partial<<<dim3(div_up(s_max, kT), div_up(2 * n_max, kDeltaRows), wb), kT, 0, s>>>(b0, B->nm_chol.p, B->X.p, B->X.p, B->ldg,
                                                                                                         strideG, B->dpart.p, n_chunk_max);
            KERNEL_POST(ctx);
            k_ekf_delta_apply<<<dim3(div_up(s_max, kT), wb), kT, 0, s>>>(b0, max_obs, B->n_ray, B->nm_chol.p, B->m_ray.p, B->dpart.p, B->ldg,
                                                                        n_chunk_max, B->ptz.p, B->vel.p, B->rays.p);
            KERNEL_POST(ctx);
            k_ekf_pp_pose<<<wb, 32, 0, s>>>(b0, B->nm_chol.p, B->X.p, B->X.p, B->ldg, strideG, B->P.p, strideP, s_tot);
            KERNEL_POST(ctx);
            const int ntc = div_up(n_max, 64);
            k_ekf_pp_blocks<<<dim3(ntc * (ntc + 1) / 2, 2, wb), dmma::kThreads, dmma::Tile<64, 64>::kSmemBytes, s>>>(b0, max_obs, B->nm_chol.p, B->m_ray.p, B->X.p, B->X.p, B->ldg,
                                                                                       strideG, B->P.p, strideP, s_tot);
            KERNEL_POST(ctx);
            // the Cholesky verdicts decide whether this wave needs the LU route at all: one extra synchronisation, only while
            // some sequence of the wave is still on the Cholesky route
            CU_CHECK(ctx, cudaMemcpyAsync(h_fail + b0, B->chol_fail.p + b0, (size_t)wb * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            CU_CHECK(ctx, cudaStreamSynchronize(s));
            for (int q = 0; q < wb; ++q)
                if (h_fail[b0 + q]) { B->h_use_lu[b0 + q] = 1; any_lu = true; }
        }
        if (n_max > 0 && any_lu) {
            // indefinite S (the reference's write-back made P indefinite): the getrf route; n_max bounds the order of every sequence
            const int n_max_lu = n_max, sm = 3 + 2 * n_max_lu;
            k_ekf_S<<<dim3(div_up(n_max_lu, 128), 2 * n_max_lu, wb), 128, 0, s>>>(b0, max_obs, B->nm_lu.p, B->Jc.p, B->Jr.p, B->G.p, B->ldg,
                                                                                 strideG, B->S.p, B->lds, strideS, B->prm.observe_var);
            KERNEL_POST(ctx);
            PROPAGATE(dense_getrf_batched(ctx, B->S.p, B->lds, strideS, B->n2_lu.p + b0, 2 * n_max_lu, wb, B->ipiv.p, B->lds, B->flags.p + 2));
            PROPAGATE(dense_getrs_rows_batched(ctx, B->S.p, B->lds, strideS, B->ipiv.p, B->perm.p, B->lds, B->G.p, B->X.p, B->ldg, strideG,
                                               B->n2_lu.p + b0, 2 * n_max_lu, 4, wb));
            k_ekf_delta_partial<<<dim3(div_up(sm, kT), div_up(2 * n_max_lu, kDeltaRows), wb), kT, 0, s>>>(b0, B->nm_lu.p, B->G.p, B->X.p, B->ldg,
                                                                                                         strideG, B->dpart.p, n_chunk_max);
            KERNEL_POST(ctx);
            k_ekf_delta_apply<<<dim3(div_up(sm, kT), wb), kT, 0, s>>>(b0, max_obs, B->n_ray, B->nm_lu.p, B->m_ray.p, B->dpart.p, B->ldg,
                                                                     n_chunk_max, B->ptz.p, B->vel.p, B->rays.p);
            KERNEL_POST(ctx);
            k_ekf_pp_pose<<<wb, 32, 0, s>>>(b0, B->nm_lu.p, B->G.p, B->X.p, B->ldg, strideG, B->P.p, strideP, s_tot);
            KERNEL_POST(ctx);
            const int ntl = div_up(n_max_lu, 64);
            k_ekf_pp_blocks<<<dim3(ntl * (ntl + 1) / 2, 2, wb), dmma::kThreads, dmma::Tile<64, 64>::kSmemBytes, s>>>(b0, max_obs, B->nm_lu.p, B->m_ray.p, B->G.p, B->X.p, B->ldg,
                                                                                       strideG, B->P.p, strideP, s_tot);
            KERNEL_POST(ctx);
            B->n_lu_total += 1;
        }
        // ---- the wave's closing host synchronisation: matched counts, argument flags, LU verdict ----
        CU_CHECK(ctx, cudaMemcpyAsync(h_n + b0, B->n_mat.p + b0, (size_t)wb * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CU_CHECK(ctx, cudaMemcpyAsync(h_flags, B->flags.p, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CU_CHECK(ctx, cudaStreamSynchronize(s));
        if (h_flags[0] & 2) return ptzba_fail(ctx, PTZBA_ERR_ARG, "observed ray index out of range");
        if (h_flags[0] & 1) return ptzba_fail(ctx, PTZBA_ERR_ARG, "observation count exceeds max_obs");
        if (h_flags[2] != 0)
            return ptzba_fail(ctx, PTZBA_ERR_NUMERIC, "innovation covariance is singular (zero pivot in column %d)", h_flags[2] - 1);
    }
    if (out_matched) memcpy(out_matched, h_n, (size_t)n_seq * sizeof(int32_t));
    return PTZBA_OK;
}

}  // namespace

extern "C" int ptzba_ekf_batch_create(ptzba_ctx* ctx, const ptzba_ekf_params* prm, int n_seq, int n_ray, int max_obs,
                                      const double* rays0, const double* ptz0, ptzba_ekf_batch** out) {
    if (!ctx) return PTZBA_ERR_ARG;
    ARG_CHECK(ctx, prm && out && n_seq > 0 && n_ray >= 0 && max_obs >= 0 && ptz0 && (n_ray == 0 || rays0));
    ARG_CHECK(ctx, prm->jac_mode == PTZBA_JAC_ANALYTIC || prm->jac_mode == PTZBA_JAC_CENTRAL_FD);
    *out = nullptr;
    cudaStream_t s = ctx->stream;
    ptzba_ekf_batch* B = new ptzba_ekf_batch();
    B->ctx = ctx; B->prm = *prm; B->n_seq = n_seq; B->n_ray = n_ray;
    if (max_obs > n_ray) max_obs = n_ray;
    B->max_obs = max_obs; B->s_tot = 3 + 2 * n_ray;
    B->h_n_act.assign(n_seq, n_ray);
    B->h_use_lu.assign(n_seq, 0);
    for (int e = 0; e < 6; ++e) B->has_disp = B->has_disp || prm->disp[e] != 0.0;
    const size_t strideP = (size_t)B->s_tot * B->s_tot;
    auto fail = [&](int code) { delete B; return code; };
#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return fail(ptzba_fail(ctx, PTZBA_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,   \
                                   cudaGetErrorString(e__)));                                          \
    } while (0)
    CU_TRY(B->rays.alloc((size_t)n_seq * n_ray * 2)); CU_TRY(B->P.alloc((size_t)n_seq * strideP));
    CU_TRY(B->ptz.alloc((size_t)n_seq * 3)); CU_TRY(B->vel.alloc((size_t)n_seq * 3)); CU_TRY(B->disp.alloc(6));
    CU_TRY(B->obs_cnt.alloc(n_seq)); CU_TRY(B->n_mat.alloc(n_seq)); CU_TRY(B->n2.alloc(n_seq)); CU_TRY(B->n_act.alloc(n_seq));
    CU_TRY(B->flags.alloc(8));
    CU_TRY(B->use_lu.alloc(n_seq));
    CU_TRY(cudaMemsetAsync(B->use_lu.p, 0, (size_t)n_seq * sizeof(int32_t), s));
    CU_TRY(B->chol_fail.alloc(n_seq)); CU_TRY(B->nm_chol.alloc(n_seq)); CU_TRY(B->n2_chol.alloc(n_seq));
    CU_TRY(B->nm_lu.alloc(n_seq)); CU_TRY(B->n2_lu.alloc(n_seq));
    CU_TRY(cudaMallocHost((void**)&B->h_back, (2 * (size_t)n_seq + 8) * sizeof(int32_t)));
    { const int st = ekf_alloc_work(B); if (st != PTZBA_OK) return fail(st); }
    if (n_ray) CU_TRY(cudaMemcpyAsync(B->rays.p, rays0, (size_t)n_seq * n_ray * 2 * sizeof(double), cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(B->ptz.p, ptz0, (size_t)n_seq * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(B->disp.p, prm->disp, 6 * sizeof(double), cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(B->n_act.p, B->h_n_act.data(), (size_t)n_seq * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemsetAsync(B->vel.p, 0, (size_t)n_seq * 3 * sizeof(double), s));
    CU_TRY(cudaMemsetAsync(B->P.p, 0, (size_t)n_seq * strideP * sizeof(double), s));
    k_ekf_init_cov<<<dim3(div_up(B->s_tot, 256), n_seq), 256, 0, s>>>(n_seq, B->s_tot, B->P.p, strideP, prm->angle_var, prm->f_var);
    ctx->launches++;
    CU_TRY(cudaStreamSynchronize(s));
    CU_TRY(cudaGetLastError());
#undef CU_TRY
    *out = B;
    return PTZBA_OK;
}

extern "C" void ptzba_ekf_batch_destroy(ptzba_ekf_batch* B) {
    if (!B) return;
    cudaStreamSynchronize(B->ctx->stream);
    delete B;
}

static int ekf_load_obs(ptzba_ekf_batch* B, int mem, const double* obs_xy, const int32_t* obs_index, const int32_t* obs_count) {
    ptzba_ctx* ctx = B->ctx;
    cudaStream_t s = ctx->stream;
    const cudaMemcpyKind kind = mem == PTZBA_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    const size_t n = (size_t)B->n_seq * B->max_obs;
    if (n) {
        CU_CHECK(ctx, cudaMemcpyAsync(B->obs_xy.p, obs_xy, n * 2 * sizeof(double), kind, s));
        CU_CHECK(ctx, cudaMemcpyAsync(B->obs_idx.p, obs_index, n * sizeof(int32_t), kind, s));
    }
    CU_CHECK(ctx, cudaMemcpyAsync(B->obs_cnt.p, obs_count, (size_t)B->n_seq * sizeof(int32_t), kind, s));
    return PTZBA_OK;
}

extern "C" int ptzba_ekf_batch_step(ptzba_ekf_batch* B, int mem, const double* obs_xy, const int32_t* obs_index,
                                    const int32_t* obs_count, int32_t* out_matched) {
    if (!B) return PTZBA_ERR_ARG;
    ARG_CHECK(B->ctx, obs_count && (B->max_obs == 0 || (obs_xy && obs_index)));
    PROPAGATE(ekf_load_obs(B, mem, obs_xy, obs_index, obs_count));
    return ekf_step(B, true, out_matched, mem == PTZBA_HOST ? obs_count : nullptr);
}

extern "C" int ptzba_ekf_batch_update_only(ptzba_ekf_batch* B, int mem, const double* obs_xy, const int32_t* obs_index,
                                           const int32_t* obs_count, int32_t* out_matched) {
    if (!B) return PTZBA_ERR_ARG;
    ARG_CHECK(B->ctx, obs_count && (B->max_obs == 0 || (obs_xy && obs_index)));
    PROPAGATE(ekf_load_obs(B, mem, obs_xy, obs_index, obs_count));
    return ekf_step(B, false, out_matched, mem == PTZBA_HOST ? obs_count : nullptr);
}

extern "C" int ptzba_ekf_batch_get(ptzba_ekf_batch* B, double* ptz, double* velocity, double* rays) {
    if (!B) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = B->ctx;
    cudaStream_t s = ctx->stream;
    if (ptz) CU_CHECK(ctx, cudaMemcpyAsync(ptz, B->ptz.p, (size_t)B->n_seq * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (velocity) CU_CHECK(ctx, cudaMemcpyAsync(velocity, B->vel.p, (size_t)B->n_seq * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (rays && B->n_ray)
        CU_CHECK(ctx, cudaMemcpyAsync(rays, B->rays.p, (size_t)B->n_seq * B->n_ray * 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    return PTZBA_OK;
}

extern "C" int ptzba_ekf_batch_set(ptzba_ekf_batch* B, int seq, const double* ptz3, const double* velocity3, const double* rays,
                                   const double* state_cov) {
    if (!B) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = B->ctx;
    ARG_CHECK(ctx, seq >= 0 && seq < B->n_seq);
    cudaStream_t s = ctx->stream;
    const size_t strideP = (size_t)B->s_tot * B->s_tot;
    const int na = B->h_n_act[seq], sa = 3 + 2 * na;
    if (ptz3) CU_CHECK(ctx, cudaMemcpyAsync(B->ptz.p + 3 * (size_t)seq, ptz3, 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    if (velocity3) CU_CHECK(ctx, cudaMemcpyAsync(B->vel.p + 3 * (size_t)seq, velocity3, 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    if (rays && na)
        CU_CHECK(ctx, cudaMemcpyAsync(B->rays.p + 2 * (size_t)B->n_ray * seq, rays, (size_t)na * 2 * sizeof(double), cudaMemcpyHostToDevice, s));
    // state_cov: dense (3 + 2 n_active)^2 row-major, placed with the batch's leading dimension
    if (state_cov) {
        CU_CHECK(ctx, cudaMemcpy2DAsync(B->P.p + strideP * seq, (size_t)B->s_tot * sizeof(double), state_cov, (size_t)sa * sizeof(double),
                                        (size_t)sa * sizeof(double), sa, cudaMemcpyHostToDevice, s));
        B->h_use_lu[seq] = 0;                       // a new covariance gets a new Cholesky attempt
        CU_CHECK(ctx, cudaMemsetAsync(B->use_lu.p + seq, 0, sizeof(int32_t), s));
    }
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    return PTZBA_OK;
}

extern "C" int ptzba_ekf_batch_get_cov(ptzba_ekf_batch* B, int seq, double* state_cov) {
    if (!B) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = B->ctx;
    ARG_CHECK(ctx, seq >= 0 && seq < B->n_seq && state_cov);
    const size_t strideP = (size_t)B->s_tot * B->s_tot;
    const int sa = 3 + 2 * B->h_n_act[seq];
    CU_CHECK(ctx, cudaMemcpy2DAsync(state_cov, (size_t)sa * sizeof(double), B->P.p + strideP * seq, (size_t)B->s_tot * sizeof(double),
                                    (size_t)sa * sizeof(double), sa, cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return PTZBA_OK;
}

// ---- ray bookkeeping between frames, on the resident state (ptz_slam.py:291-388) ---------------------------
extern "C" int ptzba_ekf_batch_n_rays(ptzba_ekf_batch* B, int seq, int32_t* n_active, int32_t* capacity) {
    if (!B) return PTZBA_ERR_ARG;
    ARG_CHECK(B->ctx, seq >= 0 && seq < B->n_seq);
    if (n_active) *n_active = B->h_n_act[seq];
    if (capacity) *capacity = B->n_ray;
    return PTZBA_OK;
}

extern "C" int ptzba_ekf_batch_get_rays(ptzba_ekf_batch* B, int seq, double* rays) {
    if (!B) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = B->ctx;
    ARG_CHECK(ctx, seq >= 0 && seq < B->n_seq && (rays || B->h_n_act[seq] == 0));
    if (B->h_n_act[seq])
        CU_CHECK(ctx, cudaMemcpyAsync(rays, B->rays.p + 2 * (size_t)B->n_ray * seq, (size_t)B->h_n_act[seq] * 2 * sizeof(double),
                                      cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return PTZBA_OK;
}

// PtzSlam.remove_rays (ptz_slam.py:291-315): the rays delete_index[0..n_del) of sequence `seq` leave the state together with their
// covariance rows / columns; the remaining rays keep their order
extern "C" int ptzba_ekf_batch_remove_rays(ptzba_ekf_batch* B, int seq, int n_del, const int32_t* delete_index) {
    if (!B) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = B->ctx;
    ARG_CHECK(ctx, seq >= 0 && seq < B->n_seq && n_del >= 0 && (n_del == 0 || delete_index));
    if (n_del == 0) return PTZBA_OK;
    cudaStream_t s = ctx->stream;
    const int na = B->h_n_act[seq];
    std::vector<char> del(na, 0);
    for (int i = 0; i < n_del; ++i) {
        ARG_CHECK(ctx, delete_index[i] >= 0 && delete_index[i] < na);
        del[delete_index[i]] = 1;
    }
    std::vector<int32_t> keep, map;
    keep.reserve(na);
    for (int r = 0; r < na; ++r)
        if (!del[r]) keep.push_back(r);
    const int n_new = (int)keep.size(), s_new = 3 + 2 * n_new;
    map.resize(s_new + n_new);
    map[0] = 0; map[1] = 1; map[2] = 2;
    for (int r = 0; r < n_new; ++r) { map[3 + 2 * r] = 3 + 2 * keep[r]; map[4 + 2 * r] = 4 + 2 * keep[r]; map[s_new + r] = keep[r]; }
    const size_t strideP = (size_t)B->s_tot * B->s_tot;
    CU_CHECK(ctx, B->idx_map.alloc(3 + 3 * (size_t)B->n_ray));
    CU_CHECK(ctx, B->P_scratch.alloc(strideP + 2 * (size_t)B->n_ray));
    CU_CHECK(ctx, cudaMemcpyAsync(B->idx_map.p, map.data(), map.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    double* P = B->P.p + strideP * seq;
    double* rays = B->rays.p + 2 * (size_t)B->n_ray * seq;
    k_ekf_compact_cov<<<dim3(div_up(s_new, 256), s_new), 256, 0, s>>>(s_new, B->s_tot, B->idx_map.p, P, B->P_scratch.p);
    KERNEL_POST(ctx);
    if (n_new) {
        k_ekf_compact_rays<<<div_up(n_new, 256), 256, 0, s>>>(n_new, B->idx_map.p + s_new, rays, B->P_scratch.p + strideP);
        KERNEL_POST(ctx);
        CU_CHECK(ctx, cudaMemcpyAsync(rays, B->P_scratch.p + strideP, (size_t)n_new * 2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    }
    CU_CHECK(ctx, cudaMemcpy2DAsync(P, (size_t)B->s_tot * sizeof(double), B->P_scratch.p, (size_t)B->s_tot * sizeof(double),
                                    (size_t)s_new * sizeof(double), s_new, cudaMemcpyDeviceToDevice, s));
    B->h_n_act[seq] = n_new;
    CU_CHECK(ctx, cudaMemcpyAsync(B->n_act.p + seq, &B->h_n_act[seq], sizeof(int32_t), cudaMemcpyHostToDevice, s));
    CU_CHECK(ctx, cudaStreamSynchronize(s));      // the host vectors go out of scope
    return PTZBA_OK;
}

// PtzSlam.add_rays (ptz_slam.py:365-384): k new rays are appended to sequence `seq` with variance angle_var and no correlation
extern "C" int ptzba_ekf_batch_add_rays(ptzba_ekf_batch* B, int seq, int k, const double* new_rays) {
    if (!B) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = B->ctx;
    ARG_CHECK(ctx, seq >= 0 && seq < B->n_seq && k >= 0 && (k == 0 || new_rays));
    if (k == 0) return PTZBA_OK;
    cudaStream_t s = ctx->stream;
    const int na = B->h_n_act[seq], n_new = na + k;
    if (n_new > B->n_ray) PROPAGATE(ekf_reserve(B, n_new + n_new / 2 + 16, B->max_obs));
    const size_t strideP = (size_t)B->s_tot * B->s_tot;
    CU_CHECK(ctx, cudaMemcpyAsync(B->rays.p + 2 * ((size_t)B->n_ray * seq + na), new_rays, (size_t)k * 2 * sizeof(double), cudaMemcpyHostToDevice, s));
    const int s_old = 3 + 2 * na, s_new = 3 + 2 * n_new;
    k_ekf_grow_cov<<<dim3(div_up(s_new, 256), s_new), 256, 0, s>>>(s_old, s_new, B->s_tot, B->P.p + strideP * seq, B->prm.angle_var);
    KERNEL_POST(ctx);
    B->h_n_act[seq] = n_new;
    CU_CHECK(ctx, cudaMemcpyAsync(B->n_act.p + seq, &B->h_n_act[seq], sizeof(int32_t), cudaMemcpyHostToDevice, s));
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    return PTZBA_OK;
}

extern "C" int ptzba_ekf_batch_predict_cov(ptzba_ekf_batch* B) {
    if (!B) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = B->ctx;
    k_ekf_predict_cov<<<div_up(B->n_seq, 128), 128, 0, ctx->stream>>>(B->n_seq, B->P.p, (size_t)B->s_tot * B->s_tot, B->s_tot, B->prm.angle_var,
                                                                     B->prm.f_var);
    KERNEL_POST(ctx);
    return PTZBA_OK;
}

extern "C" int ptzba_ekf_batch_max_obs(ptzba_ekf_batch* B, int32_t* max_obs) {
    if (!B || !max_obs) return PTZBA_ERR_ARG;
    *max_obs = B->max_obs;
    return PTZBA_OK;
}

// which factorisation route every sequence is on: route[b] = 0 Cholesky, 1 pivoted LU (its innovation covariance was indefinite once)
extern "C" int ptzba_ekf_batch_route(ptzba_ekf_batch* B, int32_t* route) {
    if (!B || !route) return PTZBA_ERR_ARG;
    for (int b = 0; b < B->n_seq; ++b) route[b] = B->h_use_lu[b];
    return PTZBA_OK;
}

// capacity for rays per sequence and observations per step (both only grow)
extern "C" int ptzba_ekf_batch_reserve(ptzba_ekf_batch* B, int ray_capacity, int max_obs) {
    if (!B) return PTZBA_ERR_ARG;
    ARG_CHECK(B->ctx, ray_capacity >= 0 && max_obs >= 0);
    return ekf_reserve(B, ray_capacity, max_obs);
}

// single sequence, HOST buffers, in place (the reference's PtzSlam.ekf_update contract)
extern "C" int ptzba_ekf_update(ptzba_ctx* ctx, const ptzba_ekf_params* prm, int n_total, double* rays, double* state_cov,
                                double* ptz3, double* velocity3, int m, const double* observed_xy, const int32_t* observed_index,
                                int32_t* out_n_matched) {
    if (!ctx) return PTZBA_ERR_ARG;
    ARG_CHECK(ctx, prm && n_total >= 0 && state_cov && ptz3 && velocity3 && m >= 0 && (n_total == 0 || rays));
    ARG_CHECK(ctx, m == 0 || (observed_xy && observed_index));
    ARG_CHECK(ctx, m <= n_total);
    ptzba_ekf_batch* B = nullptr;
    PROPAGATE(ptzba_ekf_batch_create(ctx, prm, 1, n_total, m, rays, ptz3, &B));
    int st = ptzba_ekf_batch_set(B, 0, nullptr, nullptr, nullptr, state_cov);
    int32_t cnt = m, matched = 0;
    if (st == PTZBA_OK) st = ptzba_ekf_batch_update_only(B, PTZBA_HOST, observed_xy, observed_index, &cnt, &matched);
    if (st == PTZBA_OK) st = ptzba_ekf_batch_get(B, ptz3, velocity3, rays);
    if (st == PTZBA_OK) st = ptzba_ekf_batch_get_cov(B, 0, state_cov);
    if (out_n_matched) *out_n_matched = matched;
    ptzba_ekf_batch_destroy(B);
    return st;
}
