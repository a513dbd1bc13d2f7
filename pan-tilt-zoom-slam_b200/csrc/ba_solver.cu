// ba_solver.cu - Schur-complement reduction and the trust-region (Levenberg-Marquardt / More) driver of keyframe BA.
//
// Reference call being replaced: scipy.optimize.least_squares(_compute_residual, x0, x_scale='jac', ftol=1e-4,
// method='trf')  (slam_system/bundle_adjustment.py:200-202).  scipy's trf_no_bounds (_lsq/trf.py) works on a dense
// forward-difference Jacobian and an SVD; the same algorithm is run here on the normal equations
//     (J^T J + alpha D^2) delta = -J^T r,   D = diag(scale_inv)  (x_scale='jac', _lsq/common.py:compute_jac_scale)
// whose block structure [U W; W^T V] is eliminated landmark-first:
//     S = (U + alpha D_c^2) - sum_l W_l (V_l + alpha D_l^2)^-1 W_l^T        (order 3(N-1), dense, Cholesky in dense.cu)
// The secular equation ||D delta(alpha)|| = Delta of solve_lsq_trust_region (_lsq/common.py) is solved with the same
// safeguarded Newton iteration, every evaluation being one Schur formation + factorisation + two solves.
// W blocks (3x2 per observation) are never stored: each pass recomputes them from the keyframe / landmark trig tables.
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <utility>
#include <vector>

#include <cub/cub.cuh>

#include "ba.h"
#include "dense.h"

namespace {

constexpr int kThreads = 256;
constexpr double K1 = PTZ_DEG2RAD;
constexpr double K2 = PTZ_DEG2RAD * PTZ_DEG2RAD;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// W = J_c^T J_l (3x2, per degree / pixel units), row-major w[r*2+s]; r in (pan,tilt,f), s in (theta,phi)
__device__ __forceinline__ void obs_W(const CamTrig& c, const LmTrig& l, double* w) {
    double x, y;
    ObsGeom g;
    project_fast_jac(c, l, 0.0, 0.0, x, y, g);
    const double aa = fma(g.xa, g.xa, g.ya * g.ya), ap = fma(g.xa, g.xp, g.ya * g.yp);
    w[0] = -K2 * aa;                               w[1] = -K2 * ap;
    w[2] = K2 * fma(g.xt, g.xa, g.yt * g.ya);      w[3] = K2 * fma(g.xt, g.xp, g.yt * g.yp);
    w[4] = K1 * fma(g.px, g.xa, g.py * g.ya);      w[5] = K1 * fma(g.px, g.xp, g.py * g.yp);
}

// ---- vector kernels over the full layout F = [3N camera slots | 2M landmark slots] ---------------------------------
// D = max(D_old, sqrt(diag(J^T J)))  (first call: zeros -> 1)     _lsq/common.py:compute_jac_scale
__global__ void k_scale_update(int n_pose, int n_lm, const double* __restrict__ U, const double* __restrict__ V,
                               double* __restrict__ D, int first) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nF = 3 * n_pose + 2 * n_lm;
    if (i >= nF) return;
    double diag;
    if (i < 3 * n_pose) {
        const int c = i / 3, e = i - 3 * c;
        diag = U[6 * (size_t)c + (e == 0 ? 0 : e == 1 ? 3 : 5)];
    } else {
        const int j = i - 3 * n_pose, l = j >> 1;
        diag = V[3 * (size_t)l + ((j & 1) ? 2 : 0)];
    }
    double s = sqrt(diag);
    if (first) { if (s == 0.0) s = 1.0; } else { s = fmax(s, D[i]); }
    D[i] = s;
}

// out[i] = -g[i]  (camera 0 slots forced to 0)
__global__ void k_neg(int nF, const double* __restrict__ g, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nF) out[i] = (i < 3) ? 0.0 : -g[i];
}
// w = D^2 * delta
__global__ void k_d2_mul(int nF, const double* __restrict__ D, const double* __restrict__ d, double* __restrict__ w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nF) w[i] = (i < 3) ? 0.0 : D[i] * D[i] * d[i];
}
// w = g / D^2 (camera 0 slots zero)
__global__ void k_div_d2(int nF, const double* __restrict__ D, const double* __restrict__ g, double* __restrict__ w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nF) w[i] = (i < 3) ? 0.0 : g[i] / (D[i] * D[i]);
}

// xt = x + t * delta on the packed layout (x = F + 3)
__global__ void k_axpy_packed(int n, const double* __restrict__ x, const double* __restrict__ delta_full, double t,
                              double* __restrict__ xt, double* __restrict__ delta_scaled_full) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double d = t * delta_full[i + 3];
        delta_scaled_full[i + 3] = d;
        xt[i] = x[i] + d;
    }
}
// sums: out[0] += sum (D a)^2 ; out[1] += sum g a ; out[2] += sum a^2 ; out[3] += sum b c ; out[4] += sum x^2 (packed)
// out[5] = max(out[5], max |g|)
// DETERMINISTIC (fixed summation order, no floating-point atomics): in the partitioned multi-GPU solve every rank runs this
// kernel on bit-identical replicated vectors and branches on the results (trust-region radius, secular iteration, termination
// tests); a last-bit difference between ranks would let them issue different sequences of collectives.  Blocks write their
// partial sums to `part`, the block that takes the last ticket adds them up in a fixed order.
__global__ void __launch_bounds__(256)
k_sums(int nF, const double* __restrict__ D, const double* __restrict__ g, const double* __restrict__ a,
       const double* __restrict__ b, const double* __restrict__ c, const double* __restrict__ xfull,
       double* __restrict__ part, unsigned* __restrict__ ticket, double* __restrict__ out) {
    __shared__ double sw[8][6];
    __shared__ bool last;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, mx = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nF; i += gridDim.x * blockDim.x) {
        if (i < 3) continue;
        const double ai = a ? a[i] : 0.0;
        if (a) { const double da = D[i] * ai; s0 = fma(da, da, s0); s2 = fma(ai, ai, s2); }
        if (g) { s1 = fma(g[i], ai, s1); mx = fmax(mx, fabs(g[i])); }
        if (b) s3 = fma(b[i], c[i], s3);
        if (xfull) s4 = fma(xfull[i], xfull[i], s4);
    }
    auto reduce6 = [&](double& v0, double& v1, double& v2, double& v3, double& v4, double& vm) {
        v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3); v4 = warp_sum(v4);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) vm = fmax(vm, __shfl_xor_sync(0xffffffffu, vm, off));
    };
    reduce6(s0, s1, s2, s3, s4, mx);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sw[warp][0] = s0; sw[warp][1] = s1; sw[warp][2] = s2; sw[warp][3] = s3; sw[warp][4] = s4; sw[warp][5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[6] = {0, 0, 0, 0, 0, 0};
        for (int w = 0; w < 8; ++w) {
#pragma unroll
            for (int j = 0; j < 5; ++j) t[j] += sw[w][j];
            t[5] = fmax(t[5], sw[w][5]);
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) part[6 * (size_t)blockIdx.x + j] = t[j];
        __threadfence();
        last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;       // wraps to 0: ready for the next launch
    }
    __syncthreads();
    if (!last || warp != 0) return;
    __threadfence();
    const volatile double* vp = part;
    s0 = s1 = s2 = s3 = s4 = mx = 0.0;
    for (int blk = lane; blk < (int)gridDim.x; blk += 32) {
        s0 += vp[6 * blk]; s1 += vp[6 * blk + 1]; s2 += vp[6 * blk + 2]; s3 += vp[6 * blk + 3]; s4 += vp[6 * blk + 4];
        mx = fmax(mx, vp[6 * blk + 5]);
    }
    reduce6(s0, s1, s2, s3, s4, mx);
    if (lane == 0) {
        out[0] += s0; out[1] += s1; out[2] += s2; out[3] += s3; out[4] += s4;
        out[5] = fmax(out[5], mx);
    }
}

// ---- landmark block inverse ---------------------------------------------------------------------------------------------
// Vinv = (V + alpha diag(D_l^2))^-1, packed (00,01,11).  A landmark without observations (V = 0, alpha = 0) gets a zero
// inverse: it decouples and does not move.  flag[0] counts singular blocks of observed landmarks.
__global__ void k_vinv(int n_lm, const double* __restrict__ V, const double* __restrict__ Dl, double alpha,
                       const int32_t* __restrict__ lm_ptr, double* __restrict__ Vinv, int* __restrict__ flag) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lm) return;
    const double d0 = Dl[2 * l], d1 = Dl[2 * l + 1];
    const double a = V[3 * (size_t)l] + alpha * d0 * d0, b = V[3 * (size_t)l + 1], c = V[3 * (size_t)l + 2] + alpha * d1 * d1;
    const double det = a * c - b * b;
    const bool observed = lm_ptr[l + 1] > lm_ptr[l];
    if (det > 1e-14 * a * c && det > 0.0) {
        const double id = 1.0 / det;
        Vinv[3 * (size_t)l] = c * id; Vinv[3 * (size_t)l + 1] = -b * id; Vinv[3 * (size_t)l + 2] = a * id;
    } else {
        Vinv[3 * (size_t)l] = Vinv[3 * (size_t)l + 1] = Vinv[3 * (size_t)l + 2] = 0.0;
        if (observed) atomicAdd(flag, 1);
    }
}

// ---- reduced system: diagonal blocks --------------------------------------------------------------------------------------
// S (n x n column-major, n = 3(N-1)) is zero-filled by a memset; this writes U_c + alpha D_c^2 on the diagonal blocks.
__global__ void k_schur_diag(int n_pose, const double* __restrict__ U, const double* __restrict__ Dc, double alpha,
                             double* __restrict__ S, int ld) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (c >= n_pose) return;
    const double* u = U + 6 * (size_t)c;
    const double d0 = Dc[3 * c], d1 = Dc[3 * c + 1], d2 = Dc[3 * c + 2];
    const int o = 3 * (c - 1);
    double m[3][3] = {{u[0] + alpha * d0 * d0, u[1], u[2]}, {u[1], u[3] + alpha * d1 * d1, u[4]}, {u[2], u[4], u[5] + alpha * d2 * d2}};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) S[(size_t)(o + r) + (size_t)(o + s) * ld] = m[r][s];
}

// ---- reduced system: - W Vinv W^T over all observation pairs of each landmark; one warp per landmark -------------------
// smem per warp: dmax * (12 doubles + 1 int): W_i and Y_i = W_i Vinv of every observation of the landmark.
// The d(d+1)/2 unordered pairs x 9 block entries are spread over the lanes ENTRY by ENTRY with the row index fastest, so
// the three lanes of one block column hit one 32-byte sector of the column-major S: the L2 atomic unit sees 3 sectors per
// pair instead of 9 (the FP64 RED rate is per sector).  Only the lower triangle of S is written.
__global__ void __launch_bounds__(kThreads)
k_schur_pairs(int lm_lo, int lm_hi, const int32_t* __restrict__ lm_ptr, const int32_t* __restrict__ s_cam,
              const CamTrig* __restrict__ cam_trig, const LmTrig* __restrict__ lm_trig, const double* __restrict__ Vinv,
              int dmax, double* __restrict__ S, int ld) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warps = kThreads / 32;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* Wsm = reinterpret_cast<double*>(smem_raw) + (size_t)wid * dmax * 12;
    int* Csm = reinterpret_cast<int*>(reinterpret_cast<double*>(smem_raw) + (size_t)warps * dmax * 12) + (size_t)wid * dmax;
    const int total_warps = gridDim.x * warps;
    const int q3 = lane / 9, e9 = lane - 9 * q3, sc = e9 / 3, r = e9 - 3 * sc;      // row index fastest: 3 lanes share a 32-byte sector of S
    for (int l = lm_lo + blockIdx.x * warps + wid; l < lm_hi; l += total_warps) {
        const int b = lm_ptr[l], d = lm_ptr[l + 1] - b;
        if (d == 0) continue;
        const double v00 = Vinv[3 * (size_t)l], v01 = Vinv[3 * (size_t)l + 1], v11 = Vinv[3 * (size_t)l + 2];
        const LmTrig lt = lm_trig[l];
        for (int i = lane; i < d; i += 32) {
            const int cam = s_cam[b + i];
            Csm[i] = cam;
            if (cam > 0) {
                double w[6];
                obs_W(cam_trig[cam], lt, w);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    Wsm[i * 12 + 2 * r] = w[2 * r];
                    Wsm[i * 12 + 2 * r + 1] = w[2 * r + 1];
                    Wsm[i * 12 + 6 + 2 * r] = fma(w[2 * r], v00, w[2 * r + 1] * v01);       // Y = W Vinv
                    Wsm[i * 12 + 6 + 2 * r + 1] = fma(w[2 * r], v01, w[2 * r + 1] * v11);
                }
            }
        }
        __syncwarp();
        // 27 lanes = 3 pairs x 9 block entries per step; a lane's entry (r, sc) never changes, and its pair index advances by 3,
        // so (i, j) is updated incrementally (no division / square root per entry: the kernel is instruction-issue bound)
        const int npairs = d * (d + 1) / 2;
        if (lane < 27) {
            int i = (q3 == 0) ? 0 : 1, j = (q3 == 2) ? 1 : 0;          // pair q3 of the row-major list (0,0) (1,0) (1,1) (2,0) ...
            for (int p = q3; p < npairs; p += 3) {
                const int ci = Csm[i], cj = Csm[j];
                if (ci > 0 && cj > 0 && !(i == j && r < sc)) {      // (a diagonal block only needs its lower triangle)
                    const double* yi = Wsm + i * 12 + 6;
                    const double* wj = Wsm + j * 12;
                    // block (max, min) of S receives Y_i W_j^T, transposed when the pair is listed the other way round
                    const bool sw = ci < cj;
                    const int a = sw ? sc : r, bb = sw ? r : sc;
                    double val = fma(yi[2 * a], wj[2 * bb], yi[2 * a + 1] * wj[2 * bb + 1]);
                    if (ci == cj && i != j) val += fma(yi[2 * sc], wj[2 * r], yi[2 * sc + 1] * wj[2 * r + 1]);   // same keyframe twice: B + B^T
                    const int cr = sw ? cj : ci, cc = sw ? ci : cj;
                    atomicAdd(S + (size_t)(3 * (cr - 1) + r) + (size_t)(3 * (cc - 1) + sc) * ld, -val);
                }
                j += 3;
                while (j > i) { j -= i + 1; ++i; }
            }
        }
        __syncwarp();
    }
}

// ---- default Schur formation: keyframe-pair-major pair list (k_schur_pairs is the fallback for > 65535 keyframes or > 2^31 pairs) ----
// k_schur_pairs issues one FP64 RED per block entry per observation pair (189 M at 256 x 100k x 2M) and sits at the L2
// atomic rate (0.97 ms at cfg3; this kernel: 0.46 ms).  The sparsity pattern is static, so the pairs are listed once per
// problem, KEYFRAME-PAIR MAJOR:
//   entry = (key = max_cam << 16 | min_cam, val = landmark | duplicate flag << 31), radix-sorted by key.
// The runtime kernel then streams the list like the keyframe-major fused pass: a warp keeps the 3x3 block of the current
// keyframe pair in registers across its whole run and commits it with 9 REDs when the key changes - REDs drop from 9 per
// entry to 9 per (warp, run), and what is left is FP64 arithmetic (two W evaluations + 30 FMA per entry) on 8 B / entry of
// streamed data.  W depends only on (keyframe, landmark), so a landmark observed m times from one keyframe contributes m^2
// times: the position pairs i != j of the same keyframe carry the duplicate flag (weight 2 = B + B^T of k_schur_pairs).
__global__ void k_pl_count(int lm_lo, int lm_hi, const int32_t* __restrict__ lm_ptr, const int32_t* __restrict__ s_cam,
                           long long* __restrict__ counts) {
    const int l = lm_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= lm_hi) return;
    long long m = 0;
    for (int k = lm_ptr[l]; k < lm_ptr[l + 1]; ++k) m += s_cam[k] > 0;
    counts[l - lm_lo] = m * (m + 1) / 2;
}

__global__ void k_pl_emit(int lm_lo, int lm_hi, const int32_t* __restrict__ lm_ptr, const int32_t* __restrict__ s_cam,
                          const long long* __restrict__ offs, uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
    const int l = lm_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= lm_hi) return;
    long long o = offs[l - lm_lo];
    const int b = lm_ptr[l], e = lm_ptr[l + 1];
    for (int i = b; i < e; ++i) {
        const int ci = s_cam[i];
        if (ci <= 0) continue;
        for (int j = b; j <= i; ++j) {
            const int cj = s_cam[j];
            if (cj <= 0) continue;
            const uint32_t hi = (uint32_t)(ci > cj ? ci : cj), lo = (uint32_t)(ci > cj ? cj : ci);
            key[o] = hi << 16 | lo;
            val[o] = (uint32_t)l | ((i != j && ci == cj) ? 0x80000000u : 0u);
            ++o;
        }
    }
}

__global__ void __launch_bounds__(kThreads, 2)
k_schur_pairlist(long long n_ent, long long chunk, const uint32_t* __restrict__ key, const uint32_t* __restrict__ val,
                 const CamTrig* __restrict__ cam_trig, const LmTrig* __restrict__ lm_trig, const double* __restrict__ Vinv,
                 double* __restrict__ S, int ld) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const long long begin = w * chunk;
    long long end = begin + chunk;
    if (end > n_ent) end = n_ent;
    if (begin >= end) return;
    const uint32_t kNone = 0xffffffffu;
    double a[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) a[q] = 0.0;
    uint32_t wkey = kNone;
    CamTrig wr = {0, 1, 0, 1, 1}, wc = {0, 1, 0, 1, 1};
    auto commit = [&](uint32_t k, const double* b) {              // S(block max,min) -= b, lower triangle only on the diagonal
        const int cr = (int)(k >> 16), cc = (int)(k & 0xffffu);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int sc = 0; sc < 3; ++sc)
                if (!(cr == cc && r < sc)) atomicAdd(S + (size_t)(3 * (cr - 1) + r) + (size_t)(3 * (cc - 1) + sc) * ld, -b[r * 3 + sc]);
    };
    auto flush = [&]() {
        if (wkey != kNone) {
#pragma unroll
            for (int q = 0; q < 9; ++q) a[q] = warp_sum(a[q]);
            if (lane == 0) commit(wkey, a);
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) a[q] = 0.0;
    };
    uint32_t nk = begin + lane < end ? key[begin + lane] : kNone, nv = begin + lane < end ? val[begin + lane] : 0u;
    for (long long base = begin; base < end; base += 32) {
        const uint32_t k = nk, v = nv;
        const bool act = base + lane < end;
        const long long en = base + 32 + lane;                     // next step's entry: in flight during this step's arithmetic
        nk = en < end ? key[en] : kNone;
        nv = en < end ? val[en] : 0u;
        const uint32_t k_lo = __shfl_sync(0xffffffffu, k, 0);                 // lane 0 is always active
        const bool uniform = __ballot_sync(0xffffffffu, k == k_lo || !act) == 0xffffffffu;
        if (uniform) {
            if (k_lo != wkey) {
                flush();
                wkey = k_lo;
                wr = cam_trig[k_lo >> 16];
                wc = cam_trig[k_lo & 0xffffu];
            }
        } else {
            flush();
            wkey = kNone;
        }
        if (act) {
            const int l = (int)(v & 0x7fffffffu);
            const double wt = (v >> 31) ? 2.0 : 1.0;
            const LmTrig lt = lm_trig[l];
            const double v00 = wt * Vinv[3 * (size_t)l], v01 = wt * Vinv[3 * (size_t)l + 1], v11 = wt * Vinv[3 * (size_t)l + 2];
            const int cr = (int)(k >> 16), cc = (int)(k & 0xffffu);
            double Wr[6], Wc[6];
            obs_W(uniform ? wr : cam_trig[cr], lt, Wr);
            if (cr == cc) {
#pragma unroll
                for (int q = 0; q < 6; ++q) Wc[q] = Wr[q];
            } else {
                obs_W(uniform ? wc : cam_trig[cc], lt, Wc);
            }
            double b[9];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double y0 = fma(Wr[2 * r], v00, Wr[2 * r + 1] * v01), y1 = fma(Wr[2 * r], v01, Wr[2 * r + 1] * v11);
#pragma unroll
                for (int sc = 0; sc < 3; ++sc) b[r * 3 + sc] = fma(y0, Wc[2 * sc], y1 * Wc[2 * sc + 1]);
            }
            if (uniform) {
#pragma unroll
                for (int q = 0; q < 9; ++q) a[q] += b[q];
            } else {
                commit(k, b);           // the warp straddles a key boundary (once per run): commit per lane
            }
        }
    }
    flush();
}

// ---- reduced right-hand side: out_c[cam] -= W_i Vinv rhs_l  (keyframe accumulators privatised in shared memory) --------
__global__ void __launch_bounds__(kThreads)
k_reduce_rhs(int64_t lo, int64_t hi, int64_t chunk, const int32_t* __restrict__ s_cam, const int32_t* __restrict__ s_lm,
             const CamTrig* __restrict__ cam_trig, const LmTrig* __restrict__ lm_trig, const double* __restrict__ Vinv,
             const double* __restrict__ rhs_l, int n_pose, int use_smem, double* __restrict__ out_c) {
    extern __shared__ __align__(16) double sacc[];     // [n_pose*3] when use_smem
    if (use_smem) {
        for (int i = threadIdx.x; i < n_pose * 3; i += kThreads) sacc[i] = 0.0;
        __syncthreads();
    }
    const int64_t begin = lo + (int64_t)blockIdx.x * chunk;
    int64_t end = begin + chunk;
    if (end > hi) end = hi;
    for (int64_t k = begin + threadIdx.x; k < end; k += kThreads) {
        const int cam = s_cam[k];
        if (cam == 0) continue;
        const int l = s_lm[k];
        const double v00 = Vinv[3 * (size_t)l], v01 = Vinv[3 * (size_t)l + 1], v11 = Vinv[3 * (size_t)l + 2];
        const double r0 = rhs_l[2 * (size_t)l], r1 = rhs_l[2 * (size_t)l + 1];
        const double t0 = fma(v00, r0, v01 * r1), t1 = fma(v01, r0, v11 * r1);
        double w[6];
        obs_W(cam_trig[cam], lm_trig[l], w);
        double* dst = (use_smem ? sacc : out_c) + 3 * (size_t)cam;
        atomicAdd(dst + 0, -fma(w[0], t0, w[1] * t1));
        atomicAdd(dst + 1, -fma(w[2], t0, w[3] * t1));
        atomicAdd(dst + 2, -fma(w[4], t0, w[5] * t1));
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_pose * 3; i += kThreads) {
            const double val = sacc[i];
            if (val != 0.0) atomicAdd(out_c + i, val);
        }
    }
}

// ---- back-substitution: tmp_l[l] += W_i^T y_c[cam_i]  (segmented warp reduction, heads commit atomically) --------------
__global__ void __launch_bounds__(kThreads)
k_backsub_accum(int64_t lo, int64_t hi, int64_t chunk, const int32_t* __restrict__ s_cam, const int32_t* __restrict__ s_lm,
                const CamTrig* __restrict__ cam_trig, const LmTrig* __restrict__ lm_trig, const double* __restrict__ y_c,
                double* __restrict__ tmp_l) {
    const int lane = threadIdx.x & 31;
    const int64_t begin = lo + (int64_t)blockIdx.x * chunk;
    int64_t end = begin + chunk;
    if (end > hi) end = hi;
    for (int64_t base = begin; base < end; base += kThreads) {
        const int64_t k = base + threadIdx.x;
        const bool act = k < end;
        int l = -1;
        double a0 = 0, a1 = 0;
        if (act) {
            l = s_lm[k];
            const int cam = s_cam[k];
            if (cam != 0) {
                double w[6];
                obs_W(cam_trig[cam], lm_trig[l], w);
                const double y0 = y_c[3 * (size_t)cam], y1 = y_c[3 * (size_t)cam + 1], y2 = y_c[3 * (size_t)cam + 2];
                a0 = fma(w[0], y0, fma(w[2], y1, w[4] * y2));
                a1 = fma(w[1], y0, fma(w[3], y1, w[5] * y2));
            }
        }
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int ol = __shfl_down_sync(0xffffffffu, l, off);
            const double t0 = __shfl_down_sync(0xffffffffu, a0, off), t1 = __shfl_down_sync(0xffffffffu, a1, off);
            if (lane + off < 32 && ol == l) { a0 += t0; a1 += t1; }
        }
        const int prev = __shfl_up_sync(0xffffffffu, l, 1);
        if (act && (lane == 0 || prev != l)) {
            atomicAdd(tmp_l + 2 * (size_t)l, a0);
            atomicAdd(tmp_l + 2 * (size_t)l + 1, a1);
        }
    }
}

// y_l = Vinv (rhs_l - tmp_l)
// landmarks outside [lm_lo, lm_hi) belong to other ranks: zero here, filled in by the all-reduce that follows
__global__ void k_backsub_final(int n_lm, int lm_lo, int lm_hi, const double* __restrict__ Vinv, const double* __restrict__ rhs_l,
                                const double* __restrict__ tmp_l, double* __restrict__ y_l) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lm) return;
    if (l < lm_lo || l >= lm_hi) { y_l[2 * (size_t)l] = 0.0; y_l[2 * (size_t)l + 1] = 0.0; return; }
    const double r0 = rhs_l[2 * (size_t)l] - tmp_l[2 * (size_t)l], r1 = rhs_l[2 * (size_t)l + 1] - tmp_l[2 * (size_t)l + 1];
    const double v00 = Vinv[3 * (size_t)l], v01 = Vinv[3 * (size_t)l + 1], v11 = Vinv[3 * (size_t)l + 2];
    y_l[2 * (size_t)l] = fma(v00, r0, v01 * r1);
    y_l[2 * (size_t)l + 1] = fma(v01, r0, v11 * r1);
}

// ---- ||J delta||^2 ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_jvp_sumsq(int64_t lo, int64_t hi, const int32_t* __restrict__ s_cam, const int32_t* __restrict__ s_lm,
            const CamTrig* __restrict__ cam_trig, const LmTrig* __restrict__ lm_trig, const double* __restrict__ d_c,
            const double* __restrict__ d_l, double* __restrict__ out) {
    __shared__ double sw[kThreads / 32];
    double acc = 0.0;
    for (int64_t k = lo + (int64_t)blockIdx.x * kThreads + threadIdx.x; k < hi; k += (int64_t)gridDim.x * kThreads) {
        const int cam = s_cam[k], l = s_lm[k];
        double x, y;
        ObsGeom g;
        project_fast_jac(cam_trig[cam], lm_trig[l], 0.0, 0.0, x, y, g);
        const double dth = K1 * d_l[2 * (size_t)l], dph = K1 * d_l[2 * (size_t)l + 1];
        double dp = 0, dt = 0, df = 0;
        if (cam != 0) { dp = K1 * d_c[3 * (size_t)cam]; dt = K1 * d_c[3 * (size_t)cam + 1]; df = d_c[3 * (size_t)cam + 2]; }
        const double da = dth - dp;
        const double jx = fma(g.xa, da, fma(g.xt, dt, fma(g.px, df, g.xp * dph)));
        const double jy = fma(g.ya, da, fma(g.yt, dt, fma(g.py, df, g.yp * dph)));
        acc = fma(jx, jx, fma(jy, jy, acc));
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int w = 0; w < kThreads / 32; ++w) s += sw[w];
        atomicAdd(out, s);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
struct PhaseTrace {      // PTZBA_TRACE=1: CUDA-event time of each solver phase, printed to stderr (debug aid)
    bool on = false;
    cudaStream_t s = nullptr;
    std::vector<std::pair<const char*, cudaEvent_t>> ev;
    void begin(cudaStream_t st) { on = getenv("PTZBA_TRACE") != nullptr; s = st; mark("start"); }
    void mark(const char* name) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        ev.push_back({name, e});
    }
    void report(const char* title) {
        if (!on) return;
        cudaStreamSynchronize(s);
        fprintf(stderr, "[ptzba trace] %s:", title);
        for (size_t i = 1; i < ev.size(); ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[i - 1].second, ev[i].second);
            fprintf(stderr, " %s=%.1fus", ev[i].first, ms * 1e3f);
        }
        fprintf(stderr, "\n");
        for (auto& p : ev) cudaEventDestroy(p.second);
        ev.clear();
    }
};

struct Solver {
    PhaseTrace tr;
    ptzba_ba* ba;
    ptzba_ctx* ctx;
    cudaStream_t s;
    int N, M, n, nF;
    int64_t chunk = 0;
    int obs_grid = 0;
    int pair_warps_smem = 0, pair_grid = 0;
    bool use_pairlist = false;   // keyframe-pair-major entry list (k_schur_pairlist) unless the problem is too large for its keys
    int n_factor = 0;
    int* h_flags = nullptr;      // pinned host copy of the factorisation verdict (slot 200 of ctx->h_scalars): truly asynchronous
    double *g, *D, *Dc, *Dl;
    DevBuf<int>& flags;      // persistent scratch owned by the problem (no allocation per solve)
    DevBuf<double>&tmp_l, &w_full, &dinv;
    explicit Solver(ptzba_ba* b) : ba(b), ctx(b->ctx), s(b->ctx->stream), flags(b->sol_flags), tmp_l(b->sol_tmp_l), w_full(b->sol_w), dinv(b->sol_dinv) {}

    int init() {
        N = ba->n_pose; M = ba->n_lm; n = 3 * (N - 1); nF = 3 * N + 2 * M;
        h_flags = reinterpret_cast<int*>(ctx->h_scalars + 200);
        CU_CHECK(ctx, ba->x_cur.alloc(nF)); CU_CHECK(ctx, ba->x_trial.alloc(nF));
        CU_CHECK(ctx, ba->scale_inv.alloc(nF));
        CU_CHECK(ctx, ba->Sred.alloc((size_t)(n > 0 ? n : 1) * (n > 0 ? n : 1)));
        CU_CHECK(ctx, ba->rhs_c.alloc(nF)); CU_CHECK(ctx, ba->sol_c.alloc(nF)); CU_CHECK(ctx, ba->sol2_c.alloc(nF));
        CU_CHECK(ctx, ba->rhs_l.alloc(nF));   // reduced rhs scratch (camera part only)
        CU_CHECK(ctx, ba->Vinv.alloc((size_t)M * 3));
        CU_CHECK(ctx, flags.alloc(4)); CU_CHECK(ctx, tmp_l.alloc((size_t)2 * M)); CU_CHECK(ctx, w_full.alloc(nF));
        CU_CHECK(ctx, dinv.alloc(dense_coop_dinv_doubles(n > 0 ? n : 1)));
        if (!ba->sums_ticket.p) {
            CU_CHECK(ctx, ba->sums_part.alloc((size_t)6 * ctx->sm_count));
            CU_CHECK(ctx, ba->sums_ticket.alloc(1));
            CU_CHECK(ctx, cudaMemsetAsync(ba->sums_ticket.p, 0, sizeof(unsigned), s));
        }

        g = ba->acc.gc;           // gc | gl contiguous = full layout
        D = ba->scale_inv.p; Dc = D; Dl = D + 3 * N;
        obs_grid = ctx->sm_count * 4;
        const int64_t n_part = ba->lmo_hi - ba->lmo_lo;          // this rank's observations (landmark-major slice)
        chunk = (n_part + obs_grid - 1) / obs_grid;
        chunk = (chunk + kThreads - 1) / kThreads * kThreads;
        if (chunk < kThreads) chunk = kThreads;
        obs_grid = (int)((n_part + chunk - 1) / chunk);
        if (obs_grid < 1) obs_grid = 1;
        const size_t per_warp = (size_t)(ba->max_degree > 0 ? ba->max_degree : 1) * (12 * sizeof(double) + sizeof(int));
        const size_t need = per_warp * (kThreads / 32) + 16;
        if (need > 200 * 1024)
            return ptzba_fail(ctx, PTZBA_ERR_ARG, "a landmark has %d observations: exceeds the shared-memory tile of the Schur kernel",
                              ba->max_degree);
        pair_warps_smem = (int)need;
        CU_CHECK(ctx, cudaFuncSetAttribute(k_schur_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, pair_warps_smem));
        if ((size_t)N * 3 * sizeof(double) > 40 * 1024)
            CU_CHECK(ctx, cudaFuncSetAttribute(k_reduce_rhs, cudaFuncAttributeMaxDynamicSharedMemorySize, N * 3 * (int)sizeof(double)));
        int per_sm = 1;
        CU_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_schur_pairs, kThreads, pair_warps_smem));
        if (per_sm < 1) per_sm = 1;
        pair_grid = ctx->sm_count * per_sm;
        const int max_grid = div_up(ba->lm_hi - ba->lm_lo, kThreads / 32);
        if (pair_grid > max_grid) pair_grid = max_grid > 0 ? max_grid : 1;
        use_pairlist = ba->schur_mode != PTZBA_SCHUR_PER_LANDMARK && N <= 65535;
        if (use_pairlist) PROPAGATE(build_pair_list());
        return PTZBA_OK;
    }

    // keyframe-pair-major list of all (observation, observation) pairs of every landmark; built once per problem
    int build_pair_list() {
        const int nl = ba->lm_hi - ba->lm_lo;
        if (ba->pl_ready || nl <= 0 || ba->n_obs == 0) return PTZBA_OK;
        DevBuf<long long> counts, offs;
        DevBuf<uint32_t> key_in, val_in;
        DevBuf<unsigned char> tmp;
        CU_CHECK(ctx, counts.alloc((size_t)nl + 1));
        CU_CHECK(ctx, offs.alloc((size_t)nl + 1));
        CU_CHECK(ctx, cudaMemsetAsync(counts.p, 0, ((size_t)nl + 1) * sizeof(long long), s));
        k_pl_count<<<div_up(nl, 256), 256, 0, s>>>(ba->lm_lo, ba->lm_hi, ba->lm_ptr.p, ba->s_cam.p, counts.p);
        KERNEL_POST(ctx);
        size_t bytes = 0;
        CU_CHECK(ctx, cub::DeviceScan::ExclusiveSum(nullptr, bytes, counts.p, offs.p, nl + 1, s));
        CU_CHECK(ctx, tmp.alloc(bytes));
        CU_CHECK(ctx, cub::DeviceScan::ExclusiveSum(tmp.p, bytes, counts.p, offs.p, nl + 1, s));
        long long total = 0;
        CU_CHECK(ctx, cudaMemcpyAsync(&total, offs.p + nl, sizeof(long long), cudaMemcpyDeviceToHost, s));
        CU_CHECK(ctx, cudaStreamSynchronize(s));
        if (total <= 0 || total > 0x7fffffffLL) {      // cub's 32-bit item count; larger lists keep the per-landmark kernel
            use_pairlist = false;
            return PTZBA_OK;
        }
        CU_CHECK(ctx, key_in.alloc((size_t)total)); CU_CHECK(ctx, val_in.alloc((size_t)total));
        CU_CHECK(ctx, ba->pl_key.alloc((size_t)total)); CU_CHECK(ctx, ba->pl_val.alloc((size_t)total));
        k_pl_emit<<<div_up(nl, 256), 256, 0, s>>>(ba->lm_lo, ba->lm_hi, ba->lm_ptr.p, ba->s_cam.p, offs.p, key_in.p, val_in.p);
        KERNEL_POST(ctx);
        bytes = 0;
        CU_CHECK(ctx, cub::DeviceRadixSort::SortPairs(nullptr, bytes, key_in.p, ba->pl_key.p, val_in.p, ba->pl_val.p, (int)total, 0, 32, s));
        CU_CHECK(ctx, tmp.alloc(bytes));
        CU_CHECK(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, bytes, key_in.p, ba->pl_key.p, val_in.p, ba->pl_val.p, (int)total, 0, 32, s));
        CU_CHECK(ctx, cudaStreamSynchronize(s));       // the temporaries go out of scope
        ba->pl_n = total;
        ba->pl_ready = true;
        return PTZBA_OK;
    }

    // forms and factors S(alpha); *ok = false when the factorisation broke down or an observed landmark block is singular
    int factor(double alpha, bool* ok) {
        tr.begin(s);
        CU_CHECK(ctx, cudaMemsetAsync(flags.p, 0, 4 * sizeof(int), s));
        k_vinv<<<div_up(M, 256), 256, 0, s>>>(M, ba->acc.V, Dl, alpha, ba->lm_ptr.p, ba->Vinv.p, flags.p);
        KERNEL_POST(ctx);
        if (n > 0) {
            CU_CHECK(ctx, cudaMemsetAsync(ba->Sred.p, 0, (size_t)n * n * sizeof(double), s));
            if (ba->part_rank == 0) {                       // the diagonal blocks enter the all-reduced sum once
                k_schur_diag<<<div_up(N, 128), 128, 0, s>>>(N, ba->acc.U, Dc, alpha, ba->Sred.p, n);
                KERNEL_POST(ctx);
            }
            tr.mark("vinv+memset+diag");
            if (ba->lm_hi > ba->lm_lo && ba->n_obs > 0) {
                if (use_pairlist && ba->pl_ready) {
                    const long long per_warp = 1024;              // entries per warp: 32 steps, one or two keyframe pairs
                    const long long n_warps = (ba->pl_n + per_warp - 1) / per_warp;
                    k_schur_pairlist<<<(unsigned)((n_warps + kThreads / 32 - 1) / (kThreads / 32)), kThreads, 0, s>>>(
                        ba->pl_n, per_warp, ba->pl_key.p, ba->pl_val.p, ba->cam_trig.p, ba->lm_trig.p, ba->Vinv.p, ba->Sred.p, n);
                } else {
                    k_schur_pairs<<<pair_grid, kThreads, pair_warps_smem, s>>>(ba->lm_lo, ba->lm_hi, ba->lm_ptr.p, ba->s_cam.p, ba->cam_trig.p,
                                                                              ba->lm_trig.p, ba->Vinv.p, ba->max_degree,
                                                                              ba->Sred.p, n);
                }
                KERNEL_POST(ctx);
            }
            if (ba->part_world > 1) PROPAGATE(ptzba_comm_allreduce_f64(ctx, ba->Sred.p, (int64_t)n * n));
            tr.mark("schur_pairs");
            PROPAGATE(dense_potrf_coop(ctx, ba->Sred.p, n, n, flags.p + 1, dinv.p));
            tr.mark("potrf+block_inverses");
        }
        tr.report("factor");
        ++n_factor;
        // the verdict (singular landmark block / non-positive pivot) is read together with the step's scalars in step_at():
        // one host synchronisation per step instead of two; a solve with a broken factor only produces numbers that are dropped
        CU_CHECK(ctx, cudaMemcpyAsync(h_flags, flags.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
        *ok = true;
        return PTZBA_OK;
    }

    // y = (A + alpha D^2)^-1 rhs using the current factorisation; rhs / y in the full layout
    int solve(const double* rhs, double* y) {
        tr.begin(s);
        const double* rhs_c = rhs;
        const double* rhs_l = rhs + 3 * N;
        double* red = ba->rhs_l.p;     // reduced camera rhs (full camera layout, slot 0 unused)
        if (ba->part_rank == 0) CU_CHECK(ctx, cudaMemcpyAsync(red, rhs_c, (size_t)3 * N * sizeof(double), cudaMemcpyDeviceToDevice, s));
        else CU_CHECK(ctx, cudaMemsetAsync(red, 0, (size_t)3 * N * sizeof(double), s));
        if (ba->lmo_hi > ba->lmo_lo && n > 0) {
            const int use_smem = (size_t)N * 3 * sizeof(double) <= 200 * 1024;
            k_reduce_rhs<<<obs_grid, kThreads, use_smem ? N * 3 * sizeof(double) : 0, s>>>(
                ba->lmo_lo, ba->lmo_hi, chunk, ba->s_cam.p, ba->s_lm.p, ba->cam_trig.p, ba->lm_trig.p, ba->Vinv.p, rhs_l, N, use_smem, red);
            KERNEL_POST(ctx);
        }
        if (ba->part_world > 1) PROPAGATE(ptzba_comm_allreduce_f64(ctx, red, (int64_t)3 * N));
        tr.mark("reduce_rhs");
        if (n > 0) PROPAGATE(dense_potrs_coop(ctx, ba->Sred.p, n, n, dinv.p, red + 3));
        tr.mark("potrs");
        CU_CHECK(ctx, cudaMemcpyAsync(y, red, (size_t)3 * N * sizeof(double), cudaMemcpyDeviceToDevice, s));
        CU_CHECK(ctx, cudaMemsetAsync(y, 0, 3 * sizeof(double), s));
        CU_CHECK(ctx, cudaMemsetAsync(tmp_l.p, 0, (size_t)2 * M * sizeof(double), s));
        if (ba->lmo_hi > ba->lmo_lo && n > 0) {
            k_backsub_accum<<<obs_grid, kThreads, 0, s>>>(ba->lmo_lo, ba->lmo_hi, chunk, ba->s_cam.p, ba->s_lm.p, ba->cam_trig.p,
                                                         ba->lm_trig.p, y, tmp_l.p);
            KERNEL_POST(ctx);
        }
        k_backsub_final<<<div_up(M, 256), 256, 0, s>>>(M, ba->lm_lo, ba->lm_hi, ba->Vinv.p, rhs_l, tmp_l.p, y + 3 * N);
        KERNEL_POST(ctx);
        if (ba->part_world > 1 && M > 0) PROPAGATE(ptzba_comm_allreduce_f64(ctx, y + 3 * N, (int64_t)2 * M));
        tr.mark("backsub");
        tr.report("solve");
        return PTZBA_OK;
    }

    int read_scalars(double* h, int cnt) {
        CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_scalars, ba->scal.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, s));
        CU_CHECK(ctx, cudaStreamSynchronize(s));
        for (int i = 0; i < cnt; ++i) h[i] = ctx->h_scalars[i];
        return PTZBA_OK;
    }

    // delta = -(A + alpha D^2)^-1 g  -> ba->sol_c ; returns ||D delta||, and (optionally) phi' pieces
    int step_at(double alpha, bool want_derivative, bool* ok, double* p_norm, double* wq) {
        PROPAGATE(factor(alpha, ok));
        double* rhs = ba->rhs_c.p;
        k_neg<<<div_up(nF, 256), 256, 0, s>>>(nF, g, rhs);
        KERNEL_POST(ctx);
        PROPAGATE(solve(rhs, ba->sol_c.p));
        CU_CHECK(ctx, cudaMemsetAsync(ba->scal.p, 0, 8 * sizeof(double), s));
        const double *b = nullptr, *c = nullptr;
        if (want_derivative) {
            k_d2_mul<<<div_up(nF, 256), 256, 0, s>>>(nF, D, ba->sol_c.p, w_full.p);
            KERNEL_POST(ctx);
            PROPAGATE(solve(w_full.p, ba->sol2_c.p));
            b = w_full.p; c = ba->sol2_c.p;
        }
        k_sums<<<ctx->sm_count, 256, 0, s>>>(nF, D, nullptr, ba->sol_c.p, b, c, nullptr, ba->sums_part.p, ba->sums_ticket.p, ba->scal.p);
        KERNEL_POST(ctx);
        double h[4];
        PROPAGATE(read_scalars(h, 4));                  // synchronises: h_flags of factor() has arrived as well
        *ok = (h_flags[0] == 0 && h_flags[1] == 0);
        *p_norm = std::sqrt(h[0]);
        if (wq) *wq = h[3];
        if (!std::isfinite(*p_norm)) *ok = false;
        return PTZBA_OK;
    }
};

}  // namespace

// ===========================================================================================================================
extern "C" int ptzba_ba_solve(ptzba_ba* ba, int mem, double* x, const double* reference_pose3,
                              const ptzba_ba_options* opt, ptzba_ba_report* report) {
    if (!ba) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = ba->ctx;
    ARG_CHECK(ctx, x && reference_pose3 && opt);
    const auto t_start = std::chrono::steady_clock::now();
    Solver S(ba);
    PROPAGATE(S.init());
    cudaStream_t s = S.s;
    const int N = S.N, M = S.M, nF = S.nF;
    const int nx = 3 * (N - 1) + 2 * M;
    const double ftol = opt->ftol, xtol = opt->xtol, gtol = opt->gtol;
    int max_nfev = opt->max_nfev > 0 ? opt->max_nfev : 100 * (nx > 0 ? nx : 1);

    CU_CHECK(ctx, ba->ref_stage.alloc(4));
    CU_CHECK(ctx, cudaMemcpyAsync(ba->ref_stage.p, reference_pose3, 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    struct { const double* d; } d_ref = {ba->ref_stage.p};
    double* xc = ba->x_cur.p + 3;      // packed view of the full layout
    double* xt = ba->x_trial.p + 3;
    CU_CHECK(ctx, cudaMemsetAsync(ba->x_cur.p, 0, 3 * sizeof(double), s));
    CU_CHECK(ctx, cudaMemsetAsync(ba->x_trial.p, 0, 3 * sizeof(double), s));
    CU_CHECK(ctx, cudaMemcpyAsync(xc, x, (size_t)nx * sizeof(double),
                                  mem == PTZBA_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));

    auto eval_normal = [&](const double* xp, double* cost) -> int {
        PROPAGATE(ba_set_params(ba, xp, d_ref.d, true));
        PROPAGATE(ba_fused_pass(ba, nullptr));
        CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_scalars, ba->acc.cost, sizeof(double), cudaMemcpyDeviceToHost, s));
        CU_CHECK(ctx, cudaStreamSynchronize(s));
        *cost = 0.5 * ctx->h_scalars[0];
        return PTZBA_OK;
    };

    double cost = 0;
    PROPAGATE(eval_normal(xc, &cost));
    int nfev = 1, njev = 1;
    const double cost0 = cost;
    k_scale_update<<<div_up(nF, 256), 256, 0, s>>>(N, M, ba->acc.U, ba->acc.V, S.D, 1);
    KERNEL_POST(ctx);
    // Delta = ||x0 * scale_inv||  ; g_norm, ||g_h|| etc. come from k_sums
    double h[8];
    CU_CHECK(ctx, cudaMemsetAsync(ba->scal.p, 0, 8 * sizeof(double), s));
    k_sums<<<ctx->sm_count, 256, 0, s>>>(nF, S.D, nullptr, ba->x_cur.p, nullptr, nullptr, nullptr, ba->sums_part.p, ba->sums_ticket.p, ba->scal.p);
    KERNEL_POST(ctx);
    PROPAGATE(S.read_scalars(h, 1));
    double Delta = std::sqrt(h[0]);
    if (Delta == 0) Delta = 1.0;

    double alpha = 0.0;
    int status = -1, nit = 0;
    double g_norm = 0, step_norm = -1, actual_reduction = -1;
    if (opt->verbose) printf("%12s %12s %16s %16s %14s %14s\n", "Iteration", "Total nfev", "Cost", "Cost reduction", "Step norm", "Optimality");

    while (true) {
        // g_norm (inf-norm of the gradient), ||x|| and ||g_h|| = ||g / D|| (sum (D * (g / D^2))^2) in ONE read-back:
        // the two reductions write to different slots of the scalar block
        CU_CHECK(ctx, cudaMemsetAsync(ba->scal.p, 0, 16 * sizeof(double), s));
        k_sums<<<ctx->sm_count, 256, 0, s>>>(nF, S.D, S.g, S.g, nullptr, nullptr, ba->x_cur.p, ba->sums_part.p, ba->sums_ticket.p, ba->scal.p);
        KERNEL_POST(ctx);
        k_div_d2<<<div_up(nF, 256), 256, 0, s>>>(nF, S.D, S.g, S.w_full.p);
        KERNEL_POST(ctx);
        k_sums<<<ctx->sm_count, 256, 0, s>>>(nF, S.D, nullptr, S.w_full.p, nullptr, nullptr, nullptr, ba->sums_part.p, ba->sums_ticket.p, ba->scal.p + 8);
        KERNEL_POST(ctx);
        double h16[16];
        PROPAGATE(S.read_scalars(h16, 9));
        g_norm = h16[5];
        const double x_norm = std::sqrt(h16[4]);
        const double gh_norm = std::sqrt(h16[8]);
        if (g_norm < gtol) status = 1;
        if (opt->verbose) {
            if (nit == 0) printf("%12d %12d %16.4e %16s %14s %14.2e\n", nit, nfev, cost, "", "", g_norm);
            else printf("%12d %12d %16.4e %16.2e %14.2e %14.2e\n", nit, nfev, cost, actual_reduction, step_norm, g_norm);
        }
        if (status != -1 || nfev >= max_nfev) break;

        actual_reduction = -1;
        double cost_new = cost;
        // Gauss-Newton attempt state is independent of Delta: cache it across the inner loop
        bool gn_done = false, full_rank = false;
        double gn_norm = 0, gn_wq = 0;
        while (actual_reduction <= 0 && nfev < max_nfev) {
            // ---- solve_lsq_trust_region ----------------------------------------------------------------------------
            double p_norm = 0, wq = 0;
            bool ok = false, have_step = false;
            if (!gn_done) {
                PROPAGATE(S.step_at(0.0, true, &ok, &gn_norm, &gn_wq));
                full_rank = ok;
                gn_done = true;
                if (full_rank && gn_norm <= Delta) { have_step = true; p_norm = gn_norm; alpha = 0.0; }
            } else if (full_rank && gn_norm <= Delta) {
                PROPAGATE(S.step_at(0.0, false, &ok, &p_norm, nullptr));
                have_step = true; alpha = 0.0;
            }
            double rescale = 1.0;
            if (!have_step) {
                double alpha_upper = gh_norm / Delta;
                double alpha_lower = 0.0;
                if (full_rank) {
                    const double phi = gn_norm - Delta, phi_prime = -gn_wq / gn_norm;
                    alpha_lower = -phi / phi_prime;
                }
                if ((!full_rank && alpha == 0.0)) alpha = std::fmax(0.001 * alpha_upper, std::sqrt(alpha_lower * alpha_upper));
                for (int it = 0; it < 10; ++it) {
                    if (alpha < alpha_lower || alpha > alpha_upper)
                        alpha = std::fmax(0.001 * alpha_upper, std::sqrt(alpha_lower * alpha_upper));
                    PROPAGATE(S.step_at(alpha, true, &ok, &p_norm, &wq));
                    if (!ok) { alpha_lower = alpha; alpha = alpha * 10 + 1e-12 * alpha_upper; continue; }
                    const double phi = p_norm - Delta, phi_prime = -wq / p_norm;
                    if (phi < 0) alpha_upper = alpha;
                    const double ratio = phi / phi_prime;
                    alpha_lower = std::fmax(alpha_lower, alpha - ratio);
                    alpha -= (phi + Delta) * ratio / Delta;
                    if (std::fabs(phi) < 0.01 * Delta) break;
                }
                PROPAGATE(S.step_at(alpha, false, &ok, &p_norm, nullptr));
                if (!ok) return ptzba_fail(ctx, PTZBA_ERR_NUMERIC, "reduced camera system is not positive definite at alpha=%g", alpha);
                rescale = Delta / p_norm;      // p *= Delta / norm(p)
            }
            const double step_h_norm = p_norm * rescale;
            // ---- trial point, predicted and actual reduction ---------------------------------------------------------
            k_axpy_packed<<<div_up(nx, 256), 256, 0, s>>>(nx, xc, ba->sol_c.p, rescale, xt, ba->sol2_c.p);
            KERNEL_POST(ctx);
            CU_CHECK(ctx, cudaMemsetAsync(ba->sol2_c.p, 0, 3 * sizeof(double), s));
            CU_CHECK(ctx, cudaMemsetAsync(ba->scal.p, 0, 8 * sizeof(double), s));
            k_sums<<<ctx->sm_count, 256, 0, s>>>(nF, S.D, S.g, ba->sol2_c.p, nullptr, nullptr, nullptr, ba->sums_part.p, ba->sums_ticket.p, ba->scal.p);
            KERNEL_POST(ctx);
            if (ba->lmo_hi > ba->lmo_lo) {
                k_jvp_sumsq<<<S.obs_grid, kThreads, 0, s>>>(ba->lmo_lo, ba->lmo_hi, ba->s_cam.p, ba->s_lm.p, ba->cam_trig.p, ba->lm_trig.p,
                                                           ba->sol2_c.p, ba->sol2_c.p + 3 * N, ba->scal.p + 6);
                KERNEL_POST(ctx);
            }
            // residual at the trial point (trig tables switch to x_trial; restored by the next fused pass)
            PROPAGATE(ba_set_params(ba, xt, d_ref.d));
            PROPAGATE(ba_residual_pass(ba, nullptr, ba->scal.p + 7, /*reduce=*/false));
            // ||J delta||^2 and the trial sum of squares are partial sums over this rank's slice: ONE all-reduce for both
            if (ba->part_world > 1) PROPAGATE(ptzba_comm_allreduce_f64(ctx, ba->scal.p + 6, 2));
            PROPAGATE(S.read_scalars(h, 8));
            ++nfev;
            const double g_dot_step = h[1], step_sq = h[2], Js_sq = h[6];
            cost_new = 0.5 * h[7];
            const double predicted_reduction = -(0.5 * Js_sq + g_dot_step);
            if (!std::isfinite(cost_new)) {
                Delta = 0.25 * step_h_norm;
                PROPAGATE(ba_set_params(ba, xc, d_ref.d));
                continue;
            }
            actual_reduction = cost - cost_new;
            double ratio;
            if (predicted_reduction > 0) ratio = actual_reduction / predicted_reduction;
            else if (predicted_reduction == 0 && actual_reduction == 0) ratio = 1;
            else ratio = 0;
            double Delta_new = Delta;
            if (ratio < 0.25) Delta_new = 0.25 * step_h_norm;
            else if (ratio > 0.75 && step_h_norm > 0.95 * Delta) Delta_new = 2.0 * Delta;
            step_norm = std::sqrt(step_sq);
            const bool ftol_ok = actual_reduction < ftol * cost && ratio > 0.25;
            const bool xtol_ok = step_norm < xtol * (xtol + x_norm);
            if (ftol_ok && xtol_ok) status = 4; else if (ftol_ok) status = 2; else if (xtol_ok) status = 3;
            if (status != -1) break;
            alpha *= Delta / Delta_new;
            Delta = Delta_new;
            if (actual_reduction <= 0) PROPAGATE(ba_set_params(ba, xc, d_ref.d));   // stay at x: restore the tables
        }
        if (actual_reduction > 0) {
            std::swap(ba->x_cur.p, ba->x_trial.p);
            xc = ba->x_cur.p + 3; xt = ba->x_trial.p + 3;
            PROPAGATE(eval_normal(xc, &cost));      // J, g at the accepted point (cost equals cost_new)
            ++njev;
            k_scale_update<<<div_up(nF, 256), 256, 0, s>>>(N, M, ba->acc.U, ba->acc.V, S.D, 0);
            KERNEL_POST(ctx);
        } else {
            step_norm = 0;
            actual_reduction = 0;
            PROPAGATE(ba_set_params(ba, xc, d_ref.d));
        }
        ++nit;
    }
    if (status == -1) status = 0;
    CU_CHECK(ctx, cudaMemcpyAsync(x, xc, (size_t)nx * sizeof(double),
                                  mem == PTZBA_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, s));
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    if (report) {
        report->cost0 = cost0; report->cost = cost; report->optimality = g_norm; report->status = status;
        report->nfev = nfev; report->njev = njev; report->nit = nit; report->n_factor = S.n_factor;
        report->ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
    }
    return PTZBA_OK;
}

// One LM iteration's device work at fixed damping (benchmark unit): fused pass + Schur + Cholesky + back-substitution +
// predicted reduction + trial residual pass.  x is not modified.
extern "C" int ptzba_ba_lm_iteration(ptzba_ba* ba, int mem, const double* x, const double* reference_pose3, double alpha,
                                     double* out_pred_reduction, double* out_cost_trial) {
    if (!ba) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = ba->ctx;
    ARG_CHECK(ctx, x && reference_pose3);
    Solver S(ba);
    PROPAGATE(S.init());
    cudaStream_t s = S.s;
    const int N = S.N, M = S.M, nF = S.nF;
    const int nx = 3 * (N - 1) + 2 * M;
    CU_CHECK(ctx, ba->ref_stage.alloc(4));
    CU_CHECK(ctx, cudaMemcpyAsync(ba->ref_stage.p, reference_pose3, 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    struct { const double* d; } d_ref = {ba->ref_stage.p};
    double* xc = ba->x_cur.p + 3;
    double* xt = ba->x_trial.p + 3;
    CU_CHECK(ctx, cudaMemsetAsync(ba->x_cur.p, 0, 3 * sizeof(double), s));
    CU_CHECK(ctx, cudaMemcpyAsync(xc, x, (size_t)nx * sizeof(double),
                                  mem == PTZBA_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
    PROPAGATE(ba_set_params(ba, xc, d_ref.d, true));
    PROPAGATE(ba_fused_pass(ba, nullptr));
    k_scale_update<<<div_up(nF, 256), 256, 0, s>>>(N, M, ba->acc.U, ba->acc.V, S.D, 1);
    KERNEL_POST(ctx);
    bool ok = false;
    double p_norm = 0;
    PROPAGATE(S.step_at(alpha, false, &ok, &p_norm, nullptr));
    if (!ok) return ptzba_fail(ctx, PTZBA_ERR_NUMERIC, "reduced camera system is not positive definite at alpha=%g", alpha);
    k_axpy_packed<<<div_up(nx, 256), 256, 0, s>>>(nx, xc, ba->sol_c.p, 1.0, xt, ba->sol2_c.p);
    KERNEL_POST(ctx);
    CU_CHECK(ctx, cudaMemsetAsync(ba->sol2_c.p, 0, 3 * sizeof(double), s));
    CU_CHECK(ctx, cudaMemsetAsync(ba->scal.p, 0, 8 * sizeof(double), s));
    k_sums<<<ctx->sm_count, 256, 0, s>>>(nF, S.D, S.g, ba->sol2_c.p, nullptr, nullptr, nullptr, ba->sums_part.p, ba->sums_ticket.p, ba->scal.p);
    KERNEL_POST(ctx);
    if (ba->lmo_hi > ba->lmo_lo) {
        k_jvp_sumsq<<<S.obs_grid, kThreads, 0, s>>>(ba->lmo_lo, ba->lmo_hi, ba->s_cam.p, ba->s_lm.p, ba->cam_trig.p, ba->lm_trig.p,
                                                   ba->sol2_c.p, ba->sol2_c.p + 3 * N, ba->scal.p + 6);
        KERNEL_POST(ctx);
    }
    PROPAGATE(ba_set_params(ba, xt, d_ref.d));
    PROPAGATE(ba_residual_pass(ba, nullptr, ba->scal.p + 7, /*reduce=*/false));
    if (ba->part_world > 1) PROPAGATE(ptzba_comm_allreduce_f64(ctx, ba->scal.p + 6, 2));
    double h[8];
    PROPAGATE(S.read_scalars(h, 8));
    if (out_pred_reduction) *out_pred_reduction = -(0.5 * h[6] + h[1]);
    if (out_cost_trial) *out_cost_trial = 0.5 * h[7];
    return PTZBA_OK;
}
