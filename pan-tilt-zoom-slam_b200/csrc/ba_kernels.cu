// ba_kernels.cu - per-observation passes of keyframe bundle adjustment.
//
// Reference behaviour (slam_system/bundle_adjustment.py):
//   _compute_residual :25-106  r = [proj(cam_i, l) - kp_i] per observation, camera 0 = fixed reference pose (:57-59)
//   the Jacobian scipy forms by forward differences (:200-202) is replaced by the analytic 2x3 / 2x2 blocks of
//   SURVEY.md Appendix A, and J^T J / J^T r are assembled directly into per-keyframe 3x3 (U, g_c) and per-landmark
//   2x2 (V, g_l) blocks without ever materialising J.
//
// Data layout in HBM (per observation, landmark-major sorted, SoA): cam_idx int32 | lm_idx int32 | obs_x f64 | obs_y f64
// = 24 B read; residual 16 B written (caller order).  Per-keyframe trig (sp,cp,st,ct,f) and per-landmark trig
// (sin th, cos th, T, S) tables are rebuilt once per parameter update so that no transcendental is evaluated per
// observation.  Algorithmic traffic of the fused pass: 40 B/obs + 56 B/landmark + 96 B/keyframe.
#include <cub/cub.cuh>
#include <stdlib.h>

#include "ba.h"

namespace {

constexpr int kFusedThreads = 256;

// ---------------------------------------------------------------------------------------------------------------
// problem set-up kernels
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_check_ranges(int64_t n_obs, const int32_t* __restrict__ cam, const int32_t* __restrict__ lm,
                               int n_pose, int n_lm, int* __restrict__ bad) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_obs; k += (int64_t)gridDim.x * blockDim.x) {
        if (cam[k] < 0 || cam[k] >= n_pose || lm[k] < 0 || lm[k] >= n_lm) atomicOr(bad, 1);
        if (k > 0 && lm[k] < lm[k - 1]) atomicOr(bad, 2);   // not landmark-major
    }
}

__global__ void k_iota(int64_t n, int32_t* __restrict__ a) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) a[k] = (int32_t)k;
}

__global__ void k_gather_obs(int64_t n, const int32_t* __restrict__ perm, const int32_t* __restrict__ cam,
                             const double* __restrict__ obs_xy, int32_t* __restrict__ s_cam,
                             double* __restrict__ s_ox, double* __restrict__ s_oy) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t o = perm ? perm[k] : (int32_t)k;
        s_cam[k] = cam[o];
        const double2 p = __ldg(reinterpret_cast<const double2*>(obs_xy) + o);
        s_ox[k] = p.x;
        s_oy[k] = p.y;
    }
}

// keyframe-major gather: position k takes landmark-major position perm[k]; c_orig = caller position of the residual
__global__ void k_gather_cm(int64_t n, const int32_t* __restrict__ perm, const int32_t* __restrict__ s_lm,
                            const double* __restrict__ s_ox, const double* __restrict__ s_oy, const int32_t* __restrict__ orig,
                            int32_t* __restrict__ c_lm, double* __restrict__ c_ox, double* __restrict__ c_oy,
                            int32_t* __restrict__ c_orig) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t p = perm[k];
        c_lm[k] = s_lm[p];
        c_ox[k] = s_ox[p];
        c_oy[k] = s_oy[p];
        c_orig[k] = orig ? orig[p] : p;
    }
}

// CSR offsets: lm_ptr[l] = first sorted position with lm >= l (binary search per landmark), lm_ptr[M] = n_obs
__global__ void k_lm_ptr(int n_lm, int64_t n_obs, const int32_t* __restrict__ s_lm, int32_t* __restrict__ lm_ptr,
                         int* __restrict__ max_degree) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l > n_lm) return;
    int64_t lo = 0, hi = n_obs;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (s_lm[mid] < l) lo = mid + 1; else hi = mid;
    }
    lm_ptr[l] = (int32_t)lo;
    if (l < n_lm) {
        int64_t lo2 = lo, hi2 = n_obs;
        while (lo2 < hi2) {
            const int64_t mid = (lo2 + hi2) >> 1;
            if (s_lm[mid] <= l) lo2 = mid + 1; else hi2 = mid;
        }
        atomicMax(max_degree, (int)(lo2 - lo));
    }
}

// x = [pose_1..pose_{N-1}, landmarks]; pose_0 = reference pose (bundle_adjustment.py:57-59)
__global__ void k_set_params(int n_pose, int n_lm, const double* __restrict__ x, const double* __restrict__ ref3,
                             double* __restrict__ poses, double* __restrict__ rays, CamTrig* __restrict__ cam_trig,
                             LmTrig* __restrict__ lm_trig) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pose) {
        const double* src = (i == 0) ? ref3 : (x + 3 * (size_t)(i - 1));
        const double p = src[0], t = src[1], f = src[2];
        poses[3 * i] = p; poses[3 * i + 1] = t; poses[3 * i + 2] = f;
        cam_trig[i] = make_cam_trig(p, t, f);
    }
    if (i < n_lm) {
        // x + 3(N-1) is only 8-byte aligned when N is even: scalar loads
        const double* lp = x + 3 * (size_t)(n_pose - 1) + 2 * (size_t)i;
        const double th = lp[0], ph = lp[1];
        reinterpret_cast<double2*>(rays)[i] = make_double2(th, ph);
        lm_trig[i] = make_lm_trig(th, ph);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// fused residual + Jacobian + normal-equation assembly
// ---------------------------------------------------------------------------------------------------------------
// L2 prefetch of a contiguous byte range (no registers, no shared memory): the demand loads that follow hit L2
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ double shfl_down_d(double v, int off) { return __shfl_down_sync(0xffffffffu, v, off); }

// sum of v over the run of equal keys that starts at this lane (keys are non-decreasing inside the warp).
// Keys are sorted, so once no lane has an equal key at distance `off` none has one further away: the remaining steps
// (10 SHFL each) are skipped - runs are ~5 lanes long in the quad kernels, i.e. 3 of 5 steps are usually enough.
__device__ __forceinline__ void seg_reduce5(int key, int lane, double& a, double& b, double& c, double& d, double& e) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int ok = __shfl_down_sync(0xffffffffu, key, off);
        const bool eq = (lane + off < 32) && (ok == key) && (key >= 0);
        if (!__any_sync(0xffffffffu, eq)) break;
        const double ta = shfl_down_d(a, off), tb = shfl_down_d(b, off), tc = shfl_down_d(c, off),
                     td = shfl_down_d(d, off), te = shfl_down_d(e, off);
        if (eq) { a += ta; b += tb; c += tc; d += td; e += te; }
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// CAM_SMEM: per-CTA shared-memory copies of the keyframe trig table and of the 9 keyframe accumulators.
template <bool CAM_SMEM, bool DO_CAM>
__global__ void __launch_bounds__(kFusedThreads)
k_ba_fused(int64_t n_obs, int64_t chunk, const int32_t* __restrict__ s_cam, const int32_t* __restrict__ s_lm,
           const double* __restrict__ s_ox, const double* __restrict__ s_oy, const int32_t* __restrict__ orig,
           const CamTrig* __restrict__ cam_trig, const LmTrig* __restrict__ lm_trig, int n_pose, double u, double v,
           double* __restrict__ resid, double* __restrict__ gU, double* __restrict__ gGc, double* __restrict__ gV,
           double* __restrict__ gGl, double* __restrict__ gCost) {
    extern __shared__ __align__(16) double smem[];
    double* sTrig = smem;                               // [n_pose*5]
    double* sAcc = smem + (size_t)n_pose * 5;           // [n_pose*9]
    __shared__ double sWarp[kFusedThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31;
    if (CAM_SMEM) {
        for (int i = tid; i < n_pose * 5; i += kFusedThreads) sTrig[i] = reinterpret_cast<const double*>(cam_trig)[i];
        if (DO_CAM) for (int i = tid; i < n_pose * 9; i += kFusedThreads) sAcc[i] = 0.0;
        __syncthreads();
    }
    const int64_t begin = (int64_t)blockIdx.x * chunk;
    int64_t end = begin + chunk;
    if (end > n_obs) end = n_obs;
    double cost = 0.0;
    for (int64_t base = begin; base < end; base += kFusedThreads) {
        const int64_t k = base + tid;
        const bool act = k < end;
        int lm = -1, cam = 0;
        double rx = 0, ry = 0;
        ObsGeom g = {0, 0, 0, 0, 0, 0, 0, 0};
        if (act) {
            lm = s_lm[k];
            cam = s_cam[k];
            const double ox = s_ox[k], oy = s_oy[k];
            CamTrig c;
            if (CAM_SMEM) {
                const double* t = sTrig + 5 * cam;
                c.sp = t[0]; c.cp = t[1]; c.st = t[2]; c.ct = t[3]; c.f = t[4];
            } else {
                c = cam_trig[cam];
            }
            const LmTrig l = lm_trig[lm];
            double x, y;
            project_fast_jac(c, l, u, v, x, y, g);
            rx = x - ox;
            ry = y - oy;
            if (resid) {
                const int64_t o = orig ? (int64_t)orig[k] : k;
                reinterpret_cast<double2*>(resid)[o] = make_double2(rx, ry);
            }
            cost = fma(rx, rx, fma(ry, ry, cost));
        }
        // landmark blocks (radian units; scaled to degrees when flushed)
        double vtt = fma(g.xa, g.xa, g.ya * g.ya);
        double vtp = fma(g.xa, g.xp, g.ya * g.yp);
        double vpp = fma(g.xp, g.xp, g.yp * g.yp);
        double glt = fma(g.xa, rx, g.ya * ry);
        double glp = fma(g.xp, rx, g.yp * ry);
        // keyframe blocks: pan column = -alpha column; keyframe 0 is fixed (no block)
        if (DO_CAM && act && cam != 0) {
            const double upt = -fma(g.xa, g.xt, g.ya * g.yt);
            const double upf = -fma(g.xa, g.px, g.ya * g.py);
            const double utt = fma(g.xt, g.xt, g.yt * g.yt);
            const double utf = fma(g.xt, g.px, g.yt * g.py);
            const double uff = fma(g.px, g.px, g.py * g.py);
            const double gct = fma(g.xt, rx, g.yt * ry);
            const double gcf = fma(g.px, rx, g.py * ry);
            double* a = (CAM_SMEM ? sAcc : gU) + (size_t)cam * (CAM_SMEM ? 9 : 6);
            atomicAdd(a + 0, vtt);
            atomicAdd(a + 1, upt);
            atomicAdd(a + 2, upf);
            atomicAdd(a + 3, utt);
            atomicAdd(a + 4, utf);
            atomicAdd(a + 5, uff);
            double* b = CAM_SMEM ? (a + 6) : (gGc + (size_t)cam * 3);
            atomicAdd(b + 0, -glt);
            atomicAdd(b + 1, gct);
            atomicAdd(b + 2, gcf);
        }
        seg_reduce5(lm, lane, vtt, vtp, vpp, glt, glp);
        const int prev = __shfl_up_sync(0xffffffffu, lm, 1);
        if (act && (lane == 0 || prev != lm)) {
            const double k1 = PTZ_DEG2RAD, k2 = PTZ_DEG2RAD * PTZ_DEG2RAD;
            atomicAdd(gV + 3 * (size_t)lm + 0, vtt * k2);
            atomicAdd(gV + 3 * (size_t)lm + 1, vtp * k2);
            atomicAdd(gV + 3 * (size_t)lm + 2, vpp * k2);
            atomicAdd(gGl + 2 * (size_t)lm + 0, glt * k1);
            atomicAdd(gGl + 2 * (size_t)lm + 1, glp * k1);
        }
    }
    // cost: warp -> CTA -> one atomic
    cost = warp_sum(cost);
    if (lane == 0) sWarp[tid >> 5] = cost;
    __syncthreads();
    if (tid == 0) {
        double s = 0;
        for (int w = 0; w < kFusedThreads / 32; ++w) s += sWarp[w];
        atomicAdd(gCost, s);
    }
    if (CAM_SMEM && DO_CAM) {
        // flush keyframe accumulators (converted to per-degree units): entries (pp,pt,pf,tt,tf,ff | gp,gt,gf)
        const double k1 = PTZ_DEG2RAD, k2 = PTZ_DEG2RAD * PTZ_DEG2RAD;
        for (int i = tid; i < n_pose * 9; i += kFusedThreads) {
            const double val = sAcc[i];
            if (val == 0.0) continue;
            const int c = i / 9, e = i - 9 * c;
            const double sc = (e == 0 || e == 1 || e == 3) ? k2 : (e == 2 || e == 4 || e == 6 || e == 7) ? k1 : 1.0;
            if (e < 6) atomicAdd(gU + 6 * (size_t)c + e, val * sc);
            else atomicAdd(gGc + 3 * (size_t)c + (e - 6), val * sc);
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// keyframe-major fused pass.  Observations are sorted by keyframe (landmark ascending inside a keyframe), so a warp's
// 32 observations almost always belong to ONE keyframe: its 9 accumulators live in registers across the whole chunk
// and are committed with one warp reduction when the keyframe changes (no shared-memory atomics).  The per-landmark
// blocks are committed with FP64 RED atomics that resolve in L2 (the 4 MB of landmark blocks stay L2 resident).
// PACK = 1 regroups lanes with shuffles so that the 3 (V) / 2 (g_l) values of one landmark travel in one 32-byte sector.
// ---------------------------------------------------------------------------------------------------------------
template <int PACK, int MINB>
__global__ void __launch_bounds__(kFusedThreads, MINB)
k_ba_fused_cm(int64_t n_obs, int64_t chunk, const int32_t* __restrict__ c_cam, const int32_t* __restrict__ c_lm,
              const double* __restrict__ c_ox, const double* __restrict__ c_oy, const int32_t* __restrict__ c_orig,
              const CamTrig* __restrict__ cam_trig, const LmTrig* __restrict__ lm_trig, double u, double v,
              double* __restrict__ resid, double* __restrict__ gU, double* __restrict__ gGc, double* __restrict__ gV,
              double* __restrict__ gGl, double* __restrict__ gCost) {
    __shared__ double sWarp[kFusedThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31;
    const double k1 = PTZ_DEG2RAD;
    const int64_t begin = (int64_t)blockIdx.x * chunk;
    int64_t end = begin + chunk;
    if (end > n_obs) end = n_obs;
    double cost = 0.0;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0, a8 = 0;   // pp,pt,pf,tt,tf,ff | gp,gt,gf
    int wcam = -1;                                                                    // warp-uniform current keyframe
    // commit the register accumulators of keyframe wcam: warp reduction, lane 0 issues 9 REDs (per-degree units)
    auto flush = [&]() {
        if (wcam > 0) {
            a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3); a4 = warp_sum(a4);
            a5 = warp_sum(a5); a6 = warp_sum(a6); a7 = warp_sum(a7); a8 = warp_sum(a8);
            if (lane == 0) {
                double* U = gU + 6 * (size_t)wcam;
                double* G = gGc + 3 * (size_t)wcam;
                atomicAdd(U + 0, a0); atomicAdd(U + 1, a1 * k1); atomicAdd(U + 2, a2);
                atomicAdd(U + 3, a3 * k1 * k1); atomicAdd(U + 4, a4 * k1); atomicAdd(U + 5, a5);
                atomicAdd(G + 0, a6); atomicAdd(G + 1, a7 * k1); atomicAdd(G + 2, a8);
            }
        }
        a0 = a1 = a2 = a3 = a4 = a5 = a6 = a7 = a8 = 0.0;
    };
    for (int64_t base = begin; base < end; base += kFusedThreads) {
        const int64_t k = base + tid;
        const bool act = k < end;
        int lm = -1, cam = -1;
        double rx = 0, ry = 0, kxa = 0, kya = 0, kxp = 0, kyp = 0, xt = 0, yt = 0, px = 0, py = 0;
        if (act) {
            cam = c_cam[k];
            lm = c_lm[k];
            const double ox = c_ox[k], oy = c_oy[k];
            const CamTrig c = cam_trig[cam];
            const LmTrig l = lm_trig[lm];
            double x, y;
            ObsGeom g;
            project_fast_jac(c, l, u, v, x, y, g);
            rx = x - ox;
            ry = y - oy;
            if (PACK != 2) {
                if (resid) reinterpret_cast<double2*>(resid)[c_orig[k]] = make_double2(rx, ry);
                cost = fma(rx, rx, fma(ry, ry, cost));
            }
            kxa = k1 * g.xa; kya = k1 * g.ya; kxp = k1 * g.xp; kyp = k1 * g.yp;
            xt = g.xt; yt = g.yt; px = g.px; py = g.py;
        }
        // landmark blocks, final (per-degree) units
        const double vtt = fma(kxa, kxa, kya * kya);
        const double vtp = fma(kxa, kxp, kya * kyp);
        const double vpp = fma(kxp, kxp, kyp * kyp);
        const double glt = fma(kxa, rx, kya * ry);
        const double glp = fma(kxp, rx, kyp * ry);
        if (PACK == 0) {
            if (act) {
                atomicAdd(gV + 3 * (size_t)lm + 0, vtt);
                atomicAdd(gV + 3 * (size_t)lm + 1, vtp);
                atomicAdd(gV + 3 * (size_t)lm + 2, vpp);
                atomicAdd(gGl + 2 * (size_t)lm + 0, glt);
                atomicAdd(gGl + 2 * (size_t)lm + 1, glp);
            }
        } else if (PACK == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {            // V: 4-lane groups, 8 observations per instruction
                const int src = 8 * j + (lane >> 2), e = lane & 3;
                const int lmj = __shfl_sync(0xffffffffu, lm, src);
                const double t0 = __shfl_sync(0xffffffffu, vtt, src), t1 = __shfl_sync(0xffffffffu, vtp, src),
                             t2 = __shfl_sync(0xffffffffu, vpp, src);
                if (e < 3 && lmj >= 0) atomicAdd(gV + 3 * (size_t)lmj + e, e == 0 ? t0 : e == 1 ? t1 : t2);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {            // g_l: 2-lane groups, 16 observations per instruction
                const int src = 16 * j + (lane >> 1), e = lane & 1;
                const int lmj = __shfl_sync(0xffffffffu, lm, src);
                const double t0 = __shfl_sync(0xffffffffu, glt, src), t1 = __shfl_sync(0xffffffffu, glp, src);
                if (lmj >= 0) atomicAdd(gGl + 2 * (size_t)lmj + e, e ? t1 : t0);
            }
        }
        // keyframe blocks
        const int cam_lo = __shfl_sync(0xffffffffu, cam, 0);
        const unsigned same = __ballot_sync(0xffffffffu, cam == cam_lo || !act);
        if (same == 0xffffffffu) {
            if (cam_lo != wcam) { flush(); wcam = cam_lo; }
            if (act && cam > 0) {
                a0 += vtt;
                a1 -= fma(kxa, xt, kya * yt);
                a2 -= fma(kxa, px, kya * py);
                a3 = fma(xt, xt, fma(yt, yt, a3));
                a4 = fma(xt, px, fma(yt, py, a4));
                a5 = fma(px, px, fma(py, py, a5));
                a6 -= glt;
                a7 = fma(xt, rx, fma(yt, ry, a7));
                a8 = fma(px, rx, fma(py, ry, a8));
            }
        } else {
            // the warp straddles a keyframe boundary (rare): commit what is in registers, then this iteration per lane
            flush();
            wcam = -1;
            if (act && cam > 0) {
                double* U = gU + 6 * (size_t)cam;
                double* G = gGc + 3 * (size_t)cam;
                atomicAdd(U + 0, vtt);
                atomicAdd(U + 1, -k1 * fma(kxa, xt, kya * yt));
                atomicAdd(U + 2, -fma(kxa, px, kya * py));
                atomicAdd(U + 3, k1 * k1 * fma(xt, xt, yt * yt));
                atomicAdd(U + 4, k1 * fma(xt, px, yt * py));
                atomicAdd(U + 5, fma(px, px, py * py));
                atomicAdd(G + 0, -glt);
                atomicAdd(G + 1, k1 * fma(xt, rx, yt * ry));
                atomicAdd(G + 2, fma(px, rx, py * ry));
            }
        }
    }
    flush();
    if (PACK == 2) return;
    cost = warp_sum(cost);
    if (lane == 0) sWarp[tid >> 5] = cost;
    __syncthreads();
    if (tid == 0) {
        double s = 0;
        for (int w = 0; w < kFusedThreads / 32; ++w) s += sWarp[w];
        atomicAdd(gCost, s);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Two coherent passes (fused_variant 6): every sum is formed where its operands are adjacent, nothing is scattered.
//   k_ba_lm_pass   landmark-major: residual (written), cost, per-landmark V / g_l by warp-segmented reduction
//   k_ba_cam_pass  keyframe-major: per-keyframe U / g_c in registers across the whole chunk
// Both are latency-bound streaming kernels, so each thread keeps the NEXT observation's loads in flight (register
// double buffering) while it computes the current one.
// ---------------------------------------------------------------------------------------------------------------
struct ObsLoad { int cam, lm; double ox, oy; };

__global__ void __launch_bounds__(kFusedThreads, 4)
k_ba_lm_pass(int64_t n_obs, int64_t chunk, const int32_t* __restrict__ s_cam, const int32_t* __restrict__ s_lm,
             const double* __restrict__ s_ox, const double* __restrict__ s_oy, const int32_t* __restrict__ orig,
             const CamTrig* __restrict__ cam_trig, const LmTrig* __restrict__ lm_trig, int n_pose, double u, double v,
             double* __restrict__ resid, double* __restrict__ gV, double* __restrict__ gGl, double* __restrict__ gCost) {
    extern __shared__ __align__(16) double smem[];      // keyframe trig, SoA: [5][n_pose]  (bank = keyframe id)
    __shared__ double sWarp[kFusedThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < n_pose * 5; i += kFusedThreads) {
        const int c = i / 5, e = i - 5 * c;
        smem[(size_t)e * n_pose + c] = reinterpret_cast<const double*>(cam_trig)[i];
    }
    __syncthreads();
    const int64_t begin = (int64_t)blockIdx.x * chunk;
    int64_t end = begin + chunk;
    if (end > n_obs) end = n_obs;
    double cost = 0.0;
    const double k1 = PTZ_DEG2RAD;
    // prologue: first observation of this thread
    int64_t k = begin + tid;
    ObsLoad cur = {0, -1, 0.0, 0.0};
    LmTrig curT = {0, 1, 0, 1};
    if (k < end) {
        cur.cam = s_cam[k]; cur.lm = s_lm[k]; cur.ox = s_ox[k]; cur.oy = s_oy[k];
        curT = lm_trig[cur.lm];
    }
    for (int64_t base = begin; base < end; base += kFusedThreads) {
        const bool act = k < end;
        // issue the next iteration's loads before touching the current data
        const int64_t kn = k + kFusedThreads;
        ObsLoad nxt = {0, -1, 0.0, 0.0};
        LmTrig nxtT = {0, 1, 0, 1};
        if (kn < end) {
            nxt.cam = s_cam[kn]; nxt.lm = s_lm[kn]; nxt.ox = s_ox[kn]; nxt.oy = s_oy[kn];
            nxtT = lm_trig[nxt.lm];
        }
        double rx = 0, ry = 0, kxa = 0, kya = 0, kxp = 0, kyp = 0;
        const int lm = act ? cur.lm : -1;
        if (act) {
            CamTrig c;
            c.sp = smem[cur.cam]; c.cp = smem[n_pose + cur.cam]; c.st = smem[2 * n_pose + cur.cam];
            c.ct = smem[3 * n_pose + cur.cam]; c.f = smem[4 * n_pose + cur.cam];
            double x, y;
            ObsGeom g;
            project_fast_jac(c, curT, u, v, x, y, g);
            rx = x - cur.ox;
            ry = y - cur.oy;
            if (resid) {
                const int64_t o = orig ? (int64_t)orig[k] : k;
                reinterpret_cast<double2*>(resid)[o] = make_double2(rx, ry);
            }
            cost = fma(rx, rx, fma(ry, ry, cost));
            kxa = k1 * g.xa; kya = k1 * g.ya; kxp = k1 * g.xp; kyp = k1 * g.yp;
        }
        double vtt = fma(kxa, kxa, kya * kya);
        double vtp = fma(kxa, kxp, kya * kyp);
        double vpp = fma(kxp, kxp, kyp * kyp);
        double glt = fma(kxa, rx, kya * ry);
        double glp = fma(kxp, rx, kyp * ry);
        seg_reduce5(lm, lane, vtt, vtp, vpp, glt, glp);
        const int prev = __shfl_up_sync(0xffffffffu, lm, 1);
        if (act && (lane == 0 || prev != lm)) {
            atomicAdd(gV + 3 * (size_t)lm + 0, vtt);
            atomicAdd(gV + 3 * (size_t)lm + 1, vtp);
            atomicAdd(gV + 3 * (size_t)lm + 2, vpp);
            atomicAdd(gGl + 2 * (size_t)lm + 0, glt);
            atomicAdd(gGl + 2 * (size_t)lm + 1, glp);
        }
        cur = nxt; curT = nxtT; k = kn;
    }
    cost = warp_sum(cost);
    if (lane == 0) sWarp[tid >> 5] = cost;
    __syncthreads();
    if (tid == 0) {
        double s = 0;
        for (int w = 0; w < kFusedThreads / 32; ++w) s += sWarp[w];
        atomicAdd(gCost, s);
    }
}

template <int MINB>
__global__ void __launch_bounds__(kFusedThreads, MINB)
k_ba_cam_pass(int64_t n_obs, int64_t chunk, const int32_t* __restrict__ c_cam, const int32_t* __restrict__ c_lm,
              const double* __restrict__ c_ox, const double* __restrict__ c_oy, const CamTrig* __restrict__ cam_trig,
              const LmTrig* __restrict__ lm_trig, double u, double v, double* __restrict__ gU, double* __restrict__ gGc,
              int pf) {
    const int tid = threadIdx.x, lane = tid & 31;
    const double k1 = PTZ_DEG2RAD;
    const int64_t begin = (int64_t)blockIdx.x * chunk;
    int64_t end = begin + chunk;
    if (end > n_obs) end = n_obs;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0, a8 = 0;   // radian units, scaled on commit
    int wcam = -1;
    CamTrig wc = {0, 1, 0, 1, 1};
    auto flush = [&]() {
        if (wcam > 0) {
            a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3); a4 = warp_sum(a4);
            a5 = warp_sum(a5); a6 = warp_sum(a6); a7 = warp_sum(a7); a8 = warp_sum(a8);
            if (lane == 0) {
                double* U = gU + 6 * (size_t)wcam;
                double* G = gGc + 3 * (size_t)wcam;
                const double k2 = k1 * k1;
                atomicAdd(U + 0, a0 * k2); atomicAdd(U + 1, a1 * k2); atomicAdd(U + 2, a2 * k1);
                atomicAdd(U + 3, a3 * k2); atomicAdd(U + 4, a4 * k1); atomicAdd(U + 5, a5);
                atomicAdd(G + 0, a6 * k1); atomicAdd(G + 1, a7 * k1); atomicAdd(G + 2, a8);
            }
        }
        a0 = a1 = a2 = a3 = a4 = a5 = a6 = a7 = a8 = 0.0;
    };
    int64_t k = begin + tid;
    int cam = -1, lm = 0;
    double ox = 0, oy = 0;
    LmTrig lt = {0, 1, 0, 1};
    if (k < end) { cam = c_cam[k]; lm = c_lm[k]; ox = c_ox[k]; oy = c_oy[k]; lt = lm_trig[lm]; }
    for (int64_t base = begin; base < end; base += kFusedThreads) {
        if (pf > 0 && tid < 4 && ((base - begin) & (4 * kFusedThreads - 1)) == 0) {
            const int64_t pb = base + (int64_t)pf * 4 * kFusedThreads;
            if (pb + 4 * kFusedThreads <= end) {
                if (tid == 0) l2_prefetch(c_cam + pb, kFusedThreads * 4 * 4);
                else if (tid == 1) l2_prefetch(c_lm + pb, kFusedThreads * 4 * 4);
                else if (tid == 2) l2_prefetch(c_ox + pb, kFusedThreads * 4 * 8);
                else l2_prefetch(c_oy + pb, kFusedThreads * 4 * 8);
            }
        }
        const bool act = k < end;
        const int64_t kn = k + kFusedThreads;
        int ncam = -1, nlm = 0;
        double nox = 0, noy = 0;
        LmTrig nlt = {0, 1, 0, 1};
        if (kn < end) { ncam = c_cam[kn]; nlm = c_lm[kn]; nox = c_ox[kn]; noy = c_oy[kn]; nlt = lm_trig[nlm]; }
        const int cam_lo = __shfl_sync(0xffffffffu, cam, 0);
        const unsigned same = __ballot_sync(0xffffffffu, cam == cam_lo || !act);
        const bool uniform = same == 0xffffffffu;
        if (uniform) {
            if (cam_lo != wcam) { flush(); wcam = cam_lo; if (wcam >= 0) wc = cam_trig[wcam]; }
        } else {
            flush();
            wcam = -1;
        }
        if (act && cam > 0) {
            const CamTrig c = uniform ? wc : cam_trig[cam];
            double x, y;
            ObsGeom g;
            project_fast_jac(c, lt, u, v, x, y, g);
            const double rx = x - ox, ry = y - oy;
            const double upp = fma(g.xa, g.xa, g.ya * g.ya);
            const double upt = -fma(g.xa, g.xt, g.ya * g.yt);
            const double upf = -fma(g.xa, g.px, g.ya * g.py);
            const double utt = fma(g.xt, g.xt, g.yt * g.yt);
            const double utf = fma(g.xt, g.px, g.yt * g.py);
            const double uff = fma(g.px, g.px, g.py * g.py);
            const double gp = -fma(g.xa, rx, g.ya * ry);
            const double gt = fma(g.xt, rx, g.yt * ry);
            const double gf = fma(g.px, rx, g.py * ry);
            if (uniform) {
                a0 += upp; a1 += upt; a2 += upf; a3 += utt; a4 += utf; a5 += uff; a6 += gp; a7 += gt; a8 += gf;
            } else {   // warp straddles a keyframe boundary (once per keyframe): commit per lane
                double* U = gU + 6 * (size_t)cam;
                double* G = gGc + 3 * (size_t)cam;
                const double k2 = k1 * k1;
                atomicAdd(U + 0, upp * k2); atomicAdd(U + 1, upt * k2); atomicAdd(U + 2, upf * k1);
                atomicAdd(U + 3, utt * k2); atomicAdd(U + 4, utf * k1); atomicAdd(U + 5, uff);
                atomicAdd(G + 0, gp * k1); atomicAdd(G + 1, gt * k1); atomicAdd(G + 2, gf);
            }
        }
        cam = ncam; lm = nlm; ox = nox; oy = noy; lt = nlt; k = kn;
    }
    flush();
}

// ---------------------------------------------------------------------------------------------------------------
// fused_variant 7: the two coherent passes with FOUR consecutive observations per thread.
//   * every thread issues 128-/256-bit vector loads (int4 indices, 4 x f64 pixels): 4x fewer load instructions, four
//     independent gathers in flight per thread, and residuals leave as two 256-bit stores;
//   * landmark-major pass: a run that ends inside a thread is committed directly, only the thread's LAST run enters the
//     warp-segmented reduction, so the shuffle traffic per observation drops 4x;
//   * keyframe-major pass: unchanged idea (register accumulators, one warp reduction per keyframe change).
// Chunks are multiples of 1024 observations so that every thread's quad is 32-byte aligned.
// ---------------------------------------------------------------------------------------------------------------
struct __align__(32) D4 { double a, b, c, d; };
constexpr int kQuad = 4;

__device__ __forceinline__ void commit_lm(double* __restrict__ gV, double* __restrict__ gGl, int lm, double vtt, double vtp,
                                          double vpp, double glt, double glp) {
    atomicAdd(gV + 3 * (size_t)lm + 0, vtt);
    atomicAdd(gV + 3 * (size_t)lm + 1, vtp);
    atomicAdd(gV + 3 * (size_t)lm + 2, vpp);
    atomicAdd(gGl + 2 * (size_t)lm + 0, glt);
    atomicAdd(gGl + 2 * (size_t)lm + 1, glp);
}

template <int MINB>
__global__ void __launch_bounds__(kFusedThreads, MINB)
k_ba_lm_pass4(int64_t n_obs, int64_t chunk, const int32_t* __restrict__ s_cam, const int32_t* __restrict__ s_lm,
              const double* __restrict__ s_ox, const double* __restrict__ s_oy, const int32_t* __restrict__ orig,
              const CamTrig* __restrict__ cam_trig, const LmTrig* __restrict__ lm_trig, int n_pose, double u, double v,
              double* __restrict__ resid, double* __restrict__ gV, double* __restrict__ gGl, double* __restrict__ gCost,
              int pf) {
    extern __shared__ __align__(16) double smem[];      // keyframe trig, 48 B per keyframe: {sp,cp} {st,ct} {f,-}: two LDS.128 + one LDS.64
    __shared__ double sWarp[kFusedThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < n_pose * 5; i += kFusedThreads) {
        const int c = i / 5, e = i - 5 * c;
        smem[(size_t)c * 6 + e] = reinterpret_cast<const double*>(cam_trig)[i];
    }
    __syncthreads();
    const int64_t begin = (int64_t)blockIdx.x * chunk;
    int64_t end = begin + chunk;
    if (end > n_obs) end = n_obs;
    const double k1 = PTZ_DEG2RAD;
    double cost = 0.0;
    for (int64_t base = begin; base < end; base += kFusedThreads * kQuad) {
        if (pf > 0 && tid < 4) {
            // one lane per stream asks L2 for the CTA's 1024 observations `pf` iterations ahead
            const int64_t pb = base + (int64_t)pf * kFusedThreads * kQuad;
            if (pb + kFusedThreads * kQuad <= end) {
                if (tid == 0) l2_prefetch(s_cam + pb, kFusedThreads * kQuad * 4);
                else if (tid == 1) l2_prefetch(s_lm + pb, kFusedThreads * kQuad * 4);
                else if (tid == 2) l2_prefetch(s_ox + pb, kFusedThreads * kQuad * 8);
                else l2_prefetch(s_oy + pb, kFusedThreads * kQuad * 8);
            }
        }
        const int64_t k0 = base + (int64_t)tid * kQuad;
        int cam[kQuad], lm[kQuad];
        double ox[kQuad], oy[kQuad];
        if (k0 + kQuad <= end) {
            const int4 c4 = __ldg(reinterpret_cast<const int4*>(s_cam + k0));
            const int4 l4 = __ldg(reinterpret_cast<const int4*>(s_lm + k0));
            const D4 x4 = *reinterpret_cast<const D4*>(s_ox + k0);
            const D4 y4 = *reinterpret_cast<const D4*>(s_oy + k0);
            cam[0] = c4.x; cam[1] = c4.y; cam[2] = c4.z; cam[3] = c4.w;
            lm[0] = l4.x; lm[1] = l4.y; lm[2] = l4.z; lm[3] = l4.w;
            ox[0] = x4.a; ox[1] = x4.b; ox[2] = x4.c; ox[3] = x4.d;
            oy[0] = y4.a; oy[1] = y4.b; oy[2] = y4.c; oy[3] = y4.d;
        } else {
#pragma unroll
            for (int i = 0; i < kQuad; ++i) {
                const bool in = k0 + i < end;
                cam[i] = in ? s_cam[k0 + i] : 0;
                lm[i] = in ? s_lm[k0 + i] : -1;
                ox[i] = in ? s_ox[k0 + i] : 0.0;
                oy[i] = in ? s_oy[k0 + i] : 0.0;
            }
        }
        double rx[kQuad], ry[kQuad];
        int cur = -1;
        LmTrig lt = {0, 1, 0, 1};
        double vtt = 0, vtp = 0, vpp = 0, glt = 0, glp = 0;
#pragma unroll
        for (int i = 0; i < kQuad; ++i) {
            rx[i] = 0.0; ry[i] = 0.0;
            if (lm[i] < 0) continue;
            if (lm[i] != cur) {
                if (cur >= 0) commit_lm(gV, gGl, cur, vtt, vtp, vpp, glt, glp);   // run ended inside this thread
                cur = lm[i];
                lt = lm_trig[cur];
                vtt = vtp = vpp = glt = glp = 0.0;
            }
            CamTrig c;
            {
                const double2* t = reinterpret_cast<const double2*>(smem + (size_t)cam[i] * 6);
                const double2 pa = t[0], ti = t[1];
                c.sp = pa.x; c.cp = pa.y; c.st = ti.x; c.ct = ti.y; c.f = smem[(size_t)cam[i] * 6 + 4];
            }
            double x, y;
            ObsGeom g;
            project_fast_jac(c, lt, u, v, x, y, g);
            rx[i] = x - ox[i];
            ry[i] = y - oy[i];
            cost = fma(rx[i], rx[i], fma(ry[i], ry[i], cost));
            const double kxa = k1 * g.xa, kya = k1 * g.ya, kxp = k1 * g.xp, kyp = k1 * g.yp;
            vtt = fma(kxa, kxa, fma(kya, kya, vtt));
            vtp = fma(kxa, kxp, fma(kya, kyp, vtp));
            vpp = fma(kxp, kxp, fma(kyp, kyp, vpp));
            glt = fma(kxa, rx[i], fma(kya, ry[i], glt));
            glp = fma(kxp, rx[i], fma(kyp, ry[i], glp));
        }
        if (resid) {
            if (!orig && k0 + kQuad <= end) {
                D4* dst = reinterpret_cast<D4*>(resid + 2 * k0);
                dst[0] = D4{rx[0], ry[0], rx[1], ry[1]};
                dst[1] = D4{rx[2], ry[2], rx[3], ry[3]};
            } else {
#pragma unroll
                for (int i = 0; i < kQuad; ++i)
                    if (lm[i] >= 0) {
                        const int64_t o = orig ? (int64_t)orig[k0 + i] : k0 + i;
                        reinterpret_cast<double2*>(resid)[o] = make_double2(rx[i], ry[i]);
                    }
            }
        }
        // the thread's last run joins the warp-segmented reduction (keys non-decreasing across lanes, -1 = none)
        seg_reduce5(cur, lane, vtt, vtp, vpp, glt, glp);
        const int prev = __shfl_up_sync(0xffffffffu, cur, 1);
        if (cur >= 0 && (lane == 0 || prev != cur)) commit_lm(gV, gGl, cur, vtt, vtp, vpp, glt, glp);
    }
    cost = warp_sum(cost);
    if (lane == 0) sWarp[tid >> 5] = cost;
    __syncthreads();
    if (tid == 0) {
        double s = 0;
        for (int w = 0; w < kFusedThreads / 32; ++w) s += sWarp[w];
        atomicAdd(gCost, s);
    }
}

template <int MINB>
__global__ void __launch_bounds__(kFusedThreads, MINB)
k_ba_cam_pass4(int64_t n_obs, int64_t chunk, const int32_t* __restrict__ c_cam, const int32_t* __restrict__ c_lm,
               const double* __restrict__ c_ox, const double* __restrict__ c_oy, const CamTrig* __restrict__ cam_trig,
               const LmTrig* __restrict__ lm_trig, double u, double v, double* __restrict__ gU, double* __restrict__ gGc) {
    const int tid = threadIdx.x, lane = tid & 31;
    const double k1 = PTZ_DEG2RAD, k2 = PTZ_DEG2RAD * PTZ_DEG2RAD;
    const int64_t begin = (int64_t)blockIdx.x * chunk;
    int64_t end = begin + chunk;
    if (end > n_obs) end = n_obs;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0, a8 = 0;   // radian units, scaled on commit
    int wcam = -1;
    CamTrig wc = {0, 1, 0, 1, 1};
    auto flush = [&]() {
        if (wcam > 0) {
            a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3); a4 = warp_sum(a4);
            a5 = warp_sum(a5); a6 = warp_sum(a6); a7 = warp_sum(a7); a8 = warp_sum(a8);
            if (lane == 0) {
                double* U = gU + 6 * (size_t)wcam;
                double* G = gGc + 3 * (size_t)wcam;
                atomicAdd(U + 0, a0 * k2); atomicAdd(U + 1, a1 * k2); atomicAdd(U + 2, a2 * k1);
                atomicAdd(U + 3, a3 * k2); atomicAdd(U + 4, a4 * k1); atomicAdd(U + 5, a5);
                atomicAdd(G + 0, a6 * k1); atomicAdd(G + 1, a7 * k1); atomicAdd(G + 2, a8);
            }
        }
        a0 = a1 = a2 = a3 = a4 = a5 = a6 = a7 = a8 = 0.0;
    };
    for (int64_t base = begin; base < end; base += kFusedThreads * kQuad) {
        const int64_t k0 = base + (int64_t)tid * kQuad;
        int cam[kQuad], lm[kQuad];
        double ox[kQuad], oy[kQuad];
        if (k0 + kQuad <= end) {
            const int4 c4 = __ldg(reinterpret_cast<const int4*>(c_cam + k0));
            const int4 l4 = __ldg(reinterpret_cast<const int4*>(c_lm + k0));
            const D4 x4 = *reinterpret_cast<const D4*>(c_ox + k0);
            const D4 y4 = *reinterpret_cast<const D4*>(c_oy + k0);
            cam[0] = c4.x; cam[1] = c4.y; cam[2] = c4.z; cam[3] = c4.w;
            lm[0] = l4.x; lm[1] = l4.y; lm[2] = l4.z; lm[3] = l4.w;
            ox[0] = x4.a; ox[1] = x4.b; ox[2] = x4.c; ox[3] = x4.d;
            oy[0] = y4.a; oy[1] = y4.b; oy[2] = y4.c; oy[3] = y4.d;
        } else {
#pragma unroll
            for (int i = 0; i < kQuad; ++i) {
                const bool in = k0 + i < end;
                cam[i] = in ? c_cam[k0 + i] : -1;
                lm[i] = in ? c_lm[k0 + i] : 0;
                ox[i] = in ? c_ox[k0 + i] : 0.0;
                oy[i] = in ? c_oy[k0 + i] : 0.0;
            }
        }
        LmTrig lt[kQuad];
#pragma unroll
        for (int i = 0; i < kQuad; ++i) lt[i] = lm_trig[cam[i] >= 0 ? lm[i] : 0];     // four independent gathers in flight
        // is the whole warp (128 observations) inside one keyframe?  (-1 = past the end, ignored)
        const int first = __shfl_sync(0xffffffffu, cam[0], 0);
        bool mine = true;
#pragma unroll
        for (int i = 0; i < kQuad; ++i) mine = mine && (cam[i] == first || cam[i] < 0);
        const bool uniform = __all_sync(0xffffffffu, mine) && first >= 0;
        if (uniform) {
            if (first != wcam) { flush(); wcam = first; wc = cam_trig[wcam]; }
        } else {
            flush();
            wcam = -1;
        }
#pragma unroll
        for (int i = 0; i < kQuad; ++i) {
            if (cam[i] <= 0) continue;                      // past the end, or the fixed reference keyframe
            const CamTrig c = uniform ? wc : cam_trig[cam[i]];
            double x, y;
            ObsGeom g;
            project_fast_jac(c, lt[i], u, v, x, y, g);
            const double rx = x - ox[i], ry = y - oy[i];
            const double upp = fma(g.xa, g.xa, g.ya * g.ya);
            const double upt = -fma(g.xa, g.xt, g.ya * g.yt);
            const double upf = -fma(g.xa, g.px, g.ya * g.py);
            const double utt = fma(g.xt, g.xt, g.yt * g.yt);
            const double utf = fma(g.xt, g.px, g.yt * g.py);
            const double uff = fma(g.px, g.px, g.py * g.py);
            const double gp = -fma(g.xa, rx, g.ya * ry);
            const double gt = fma(g.xt, rx, g.yt * ry);
            const double gf = fma(g.px, rx, g.py * ry);
            if (uniform) {
                a0 += upp; a1 += upt; a2 += upf; a3 += utt; a4 += utf; a5 += uff; a6 += gp; a7 += gt; a8 += gf;
            } else {
                double* U = gU + 6 * (size_t)cam[i];
                double* G = gGc + 3 * (size_t)cam[i];
                atomicAdd(U + 0, upp * k2); atomicAdd(U + 1, upt * k2); atomicAdd(U + 2, upf * k1);
                atomicAdd(U + 3, utt * k2); atomicAdd(U + 4, utf * k1); atomicAdd(U + 5, uff);
                atomicAdd(G + 0, gp * k1); atomicAdd(G + 1, gt * k1); atomicAdd(G + 2, gf);
            }
        }
    }
    flush();
}

// ---------------------------------------------------------------------------------------------------------------
// fused_variant 13: ONE launch, two CTA roles running side by side on every SM.
//   landmark role  (k_ba_lm_pass4's work): residual, cost, V / g_l, quads of consecutive observations per thread
//   keyframe role  (k_ba_cam_pass's work): U / g_c in registers, one observation per thread
// The two roles touch disjoint accumulators, so they need no ordering; launched together they overlap each other's
// memory latency and share the FP64 / LSU pipes, and the second launch (and its tail) disappears.
// Both roles are software pipelined TWO iterations deep: index loads run two iterations ahead, so that the dependent
// landmark-trig gather (address = a loaded index) can itself be issued a full iteration before its use.
// CAMREP = 8 replicates the keyframe trig table 8x in shared memory, one copy per 16-byte bank group, so that the
// per-observation keyframe lookups (random rows) are bank-conflict free: lane l reads copy (l & 7).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDualThreads = 256;

template <int CAMREP>
__device__ __forceinline__ void dual_fill_cam(double* __restrict__ sCam, const CamTrig* __restrict__ cam_trig, int n_pose) {
    const int tid = threadIdx.x;
    if (CAMREP == 1) {
        for (int i = tid; i < n_pose * 5; i += kDualThreads) {
            const int c = i / 5, e = i - 5 * c;
            sCam[(size_t)c * 6 + e] = reinterpret_cast<const double*>(cam_trig)[i];
        }
    } else {
        // part j in {0: (sp,cp), 1: (st,ct), 2: (f,0)}; entry (row c, copy g) at ((j * n_pose + c) * 8 + g) * 16 bytes
        double2* s2 = reinterpret_cast<double2*>(sCam);
        for (int i = tid; i < n_pose * 3 * 8; i += kDualThreads) {
            const int cj = i >> 3, c = cj % n_pose, j = cj / n_pose;
            const double* t = reinterpret_cast<const double*>(cam_trig) + 5 * (size_t)c;
            s2[i] = (j == 0) ? make_double2(t[0], t[1]) : (j == 1) ? make_double2(t[2], t[3]) : make_double2(t[4], 0.0);
        }
    }
}

template <int CAMREP>
__device__ __forceinline__ CamTrig dual_cam(const double* __restrict__ sCam, int cam, int n_pose, int lane) {
    CamTrig c;
    if (CAMREP == 1) {
        const double2* t = reinterpret_cast<const double2*>(sCam + (size_t)cam * 6);
        const double2 pa = t[0], ti = t[1];
        c.sp = pa.x; c.cp = pa.y; c.st = ti.x; c.ct = ti.y; c.f = sCam[(size_t)cam * 6 + 4];
    } else {
        const double2* s2 = reinterpret_cast<const double2*>(sCam) + (lane & 7);
        const double2 pa = s2[(size_t)cam * 8], ti = s2[((size_t)n_pose + cam) * 8], ff = s2[((size_t)2 * n_pose + cam) * 8];
        c.sp = pa.x; c.cp = pa.y; c.st = ti.x; c.ct = ti.y; c.f = ff.x;
    }
    return c;
}

template <int CAMREP>
__device__ __forceinline__ void dual_lm_role(const double* __restrict__ sCam, double* __restrict__ sWarp, int64_t begin,
                                             int64_t end, const int32_t* __restrict__ s_cam, const int32_t* __restrict__ s_lm,
                                             const double* __restrict__ s_ox, const double* __restrict__ s_oy,
                                             const int32_t* __restrict__ orig, const LmTrig* __restrict__ lm_trig, int n_pose,
                                             double u, double v, double* __restrict__ resid, double* __restrict__ gV,
                                             double* __restrict__ gGl, double* __restrict__ gCost) {
    const int tid = threadIdx.x, lane = tid & 31;
    constexpr int64_t kStep = (int64_t)kDualThreads * kQuad;
    const double k1 = PTZ_DEG2RAD;
    // masked quad loads: arrays are padded by 4 entries, positions >= end read as "no observation" (lm = -1)
    auto ld_lm4 = [&](int64_t k) {
        int4 r = make_int4(-1, -1, -1, -1);
        if (k < end) {
            r = __ldg(reinterpret_cast<const int4*>(s_lm + k));
            if (k + 1 >= end) r.y = -1;
            if (k + 2 >= end) r.z = -1;
            if (k + 3 >= end) r.w = -1;
        }
        return r;
    };
    int4 camC = make_int4(0, 0, 0, 0), camN = camC, lmC, lmN, lmN2;
    D4 oxC = {0, 0, 0, 0}, oyC = oxC, oxN = oxC, oyN = oxC;
    const LmTrig unitT = {0, 1, 0, 1};
    LmTrig tAC = unitT, tBC = unitT, tAN = unitT, tBN = unitT;
    int64_t k0 = begin + (int64_t)tid * kQuad;
    // prologue: quad 0 complete, index quad 1, (index quad 2 and the rest of quad 1 are issued at the top of iteration 0)
    lmC = ld_lm4(k0);
    lmN = ld_lm4(k0 + kStep);
    if (k0 < end) {
        camC = __ldg(reinterpret_cast<const int4*>(s_cam + k0));
        oxC = *reinterpret_cast<const D4*>(s_ox + k0);
        oyC = *reinterpret_cast<const D4*>(s_oy + k0);
    }
    if (lmC.x >= 0) tAC = lm_trig[lmC.x];
    if (lmC.w >= 0 && lmC.w != lmC.x) tBC = lm_trig[lmC.w];
    double cost = 0.0;
    for (int64_t base = begin; base < end; base += kStep, k0 += kStep) {
        // ---- issue everything the NEXT iterations need before touching the current quad ----
        lmN2 = ld_lm4(k0 + 2 * kStep);
        if (k0 + kStep < end) {
            camN = __ldg(reinterpret_cast<const int4*>(s_cam + k0 + kStep));
            oxN = *reinterpret_cast<const D4*>(s_ox + k0 + kStep);
            oyN = *reinterpret_cast<const D4*>(s_oy + k0 + kStep);
        }
        if (lmN.x >= 0) tAN = lm_trig[lmN.x];
        if (lmN.w >= 0 && lmN.w != lmN.x) tBN = lm_trig[lmN.w];
        // ---- current quad ----
        const int cam[kQuad] = {camC.x, camC.y, camC.z, camC.w};
        const int lm[kQuad] = {lmC.x, lmC.y, lmC.z, lmC.w};
        const double ox[kQuad] = {oxC.a, oxC.b, oxC.c, oxC.d};
        const double oy[kQuad] = {oyC.a, oyC.b, oyC.c, oyC.d};
        double rx[kQuad], ry[kQuad];
        int cur = -1;
        LmTrig lt = unitT;
        double vtt = 0, vtp = 0, vpp = 0, glt = 0, glp = 0;
#pragma unroll
        for (int i = 0; i < kQuad; ++i) {
            rx[i] = 0.0; ry[i] = 0.0;
            if (lm[i] < 0) continue;
            if (lm[i] != cur) {
                if (cur >= 0) commit_lm(gV, gGl, cur, vtt, vtp, vpp, glt, glp);   // run ended inside this thread
                cur = lm[i];
                lt = (cur == lm[0]) ? tAC : (cur == lm[3]) ? tBC : lm_trig[cur];
                vtt = vtp = vpp = glt = glp = 0.0;
            }
            const CamTrig c = dual_cam<CAMREP>(sCam, cam[i], n_pose, lane);
            double x, y;
            ObsGeom g;
            project_fast_jac(c, lt, u, v, x, y, g);
            rx[i] = x - ox[i];
            ry[i] = y - oy[i];
            cost = fma(rx[i], rx[i], fma(ry[i], ry[i], cost));
            const double kxa = k1 * g.xa, kya = k1 * g.ya, kxp = k1 * g.xp, kyp = k1 * g.yp;
            vtt = fma(kxa, kxa, fma(kya, kya, vtt));
            vtp = fma(kxa, kxp, fma(kya, kyp, vtp));
            vpp = fma(kxp, kxp, fma(kyp, kyp, vpp));
            glt = fma(kxa, rx[i], fma(kya, ry[i], glt));
            glp = fma(kxp, rx[i], fma(kyp, ry[i], glp));
        }
        if (resid) {
            if (!orig && k0 + kQuad <= end) {
                D4* dst = reinterpret_cast<D4*>(resid + 2 * k0);
                dst[0] = D4{rx[0], ry[0], rx[1], ry[1]};
                dst[1] = D4{rx[2], ry[2], rx[3], ry[3]};
            } else {
#pragma unroll
                for (int i = 0; i < kQuad; ++i)
                    if (lm[i] >= 0) {
                        const int64_t o = orig ? (int64_t)orig[k0 + i] : k0 + i;
                        reinterpret_cast<double2*>(resid)[o] = make_double2(rx[i], ry[i]);
                    }
            }
        }
        // the thread's last run joins the warp-segmented reduction (keys non-decreasing across lanes, -1 = none)
        seg_reduce5(cur, lane, vtt, vtp, vpp, glt, glp);
        const int prev = __shfl_up_sync(0xffffffffu, cur, 1);
        if (cur >= 0 && (lane == 0 || prev != cur)) commit_lm(gV, gGl, cur, vtt, vtp, vpp, glt, glp);
        // ---- rotate the pipeline registers ----
        camC = camN; oxC = oxN; oyC = oyN; lmC = lmN; tAC = tAN; tBC = tBN; lmN = lmN2;
    }
    cost = warp_sum(cost);
    if (lane == 0) sWarp[tid >> 5] = cost;
    __syncthreads();
    if (tid == 0) {
        double s = 0;
        for (int w = 0; w < kDualThreads / 32; ++w) s += sWarp[w];
        atomicAdd(gCost, s);
    }
}

__device__ __forceinline__ void dual_cam_role(int64_t begin, int64_t end, const int32_t* __restrict__ c_cam,
                                              const int32_t* __restrict__ c_lm, const double* __restrict__ c_ox,
                                              const double* __restrict__ c_oy, const CamTrig* __restrict__ cam_trig,
                                              const LmTrig* __restrict__ lm_trig, double u, double v, double* __restrict__ gU,
                                              double* __restrict__ gGc) {
    const int tid = threadIdx.x, lane = tid & 31;
    const double k1 = PTZ_DEG2RAD, k2 = PTZ_DEG2RAD * PTZ_DEG2RAD;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0, a8 = 0;   // radian units, scaled on commit
    int wcam = -1;
    CamTrig wc = {0, 1, 0, 1, 1};
    auto flush = [&]() {
        if (wcam > 0) {
            a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3); a4 = warp_sum(a4);
            a5 = warp_sum(a5); a6 = warp_sum(a6); a7 = warp_sum(a7); a8 = warp_sum(a8);
            if (lane == 0) {
                double* U = gU + 6 * (size_t)wcam;
                double* G = gGc + 3 * (size_t)wcam;
                atomicAdd(U + 0, a0 * k2); atomicAdd(U + 1, a1 * k2); atomicAdd(U + 2, a2 * k1);
                atomicAdd(U + 3, a3 * k2); atomicAdd(U + 4, a4 * k1); atomicAdd(U + 5, a5);
                atomicAdd(G + 0, a6 * k1); atomicAdd(G + 1, a7 * k1); atomicAdd(G + 2, a8);
            }
        }
        a0 = a1 = a2 = a3 = a4 = a5 = a6 = a7 = a8 = 0.0;
    };
    const LmTrig unitT = {0, 1, 0, 1};
    int64_t k = begin + tid;
    int cam = -1, camN = -1, lmN = -1, lmN2 = -1;
    double ox = 0, oy = 0, oxN = 0, oyN = 0;
    LmTrig lt = unitT, ltN = unitT;
    if (k < end) { cam = c_cam[k]; ox = c_ox[k]; oy = c_oy[k]; lt = lm_trig[c_lm[k]]; }
    if (k + kDualThreads < end) lmN = c_lm[k + kDualThreads];
    for (int64_t base = begin; base < end; base += kDualThreads, k += kDualThreads) {
        // ---- issue the next iterations' loads ----
        lmN2 = (k + 2 * kDualThreads < end) ? c_lm[k + 2 * kDualThreads] : -1;
        camN = -1;
        if (k + kDualThreads < end) { camN = c_cam[k + kDualThreads]; oxN = c_ox[k + kDualThreads]; oyN = c_oy[k + kDualThreads]; }
        if (lmN >= 0) ltN = lm_trig[lmN];
        // ---- current observation ----
        const bool act = cam >= 0;
        const int cam_lo = __shfl_sync(0xffffffffu, cam, 0);
        const unsigned same = __ballot_sync(0xffffffffu, cam == cam_lo || !act);
        const bool uniform = same == 0xffffffffu;
        if (uniform) {
            if (cam_lo != wcam) { flush(); wcam = cam_lo; if (wcam >= 0) wc = cam_trig[wcam]; }
        } else {
            flush();
            wcam = -1;
        }
        if (act && cam > 0) {
            const CamTrig c = uniform ? wc : cam_trig[cam];
            double x, y;
            ObsGeom g;
            project_fast_jac(c, lt, u, v, x, y, g);
            const double rx = x - ox, ry = y - oy;
            const double upp = fma(g.xa, g.xa, g.ya * g.ya);
            const double upt = -fma(g.xa, g.xt, g.ya * g.yt);
            const double upf = -fma(g.xa, g.px, g.ya * g.py);
            const double utt = fma(g.xt, g.xt, g.yt * g.yt);
            const double utf = fma(g.xt, g.px, g.yt * g.py);
            const double uff = fma(g.px, g.px, g.py * g.py);
            const double gp = -fma(g.xa, rx, g.ya * ry);
            const double gt = fma(g.xt, rx, g.yt * ry);
            const double gf = fma(g.px, rx, g.py * ry);
            if (uniform) {
                a0 += upp; a1 += upt; a2 += upf; a3 += utt; a4 += utf; a5 += uff; a6 += gp; a7 += gt; a8 += gf;
            } else {   // warp straddles a keyframe boundary (once per keyframe): commit per lane
                double* U = gU + 6 * (size_t)cam;
                double* G = gGc + 3 * (size_t)cam;
                atomicAdd(U + 0, upp * k2); atomicAdd(U + 1, upt * k2); atomicAdd(U + 2, upf * k1);
                atomicAdd(U + 3, utt * k2); atomicAdd(U + 4, utf * k1); atomicAdd(U + 5, uff);
                atomicAdd(G + 0, gp * k1); atomicAdd(G + 1, gt * k1); atomicAdd(G + 2, gf);
            }
        }
        cam = camN; ox = oxN; oy = oyN; lt = ltN; lmN = lmN2;
    }
    flush();
}

// CTA b takes the landmark role when the Bresenham line b * n_lm_ctas / gridDim steps, so the two roles alternate in
// dispatch order and every SM hosts both.
template <int CAMREP>
__global__ void __launch_bounds__(kDualThreads, 2)
k_ba_dual(int64_t n_obs, int n_lm_ctas, int only, int64_t chunk_lm, int64_t chunk_cam, const int32_t* __restrict__ s_cam,
          const int32_t* __restrict__ s_lm, const double* __restrict__ s_ox, const double* __restrict__ s_oy,
          const int32_t* __restrict__ orig, const int32_t* __restrict__ c_cam, const int32_t* __restrict__ c_lm,
          const double* __restrict__ c_ox, const double* __restrict__ c_oy, const CamTrig* __restrict__ cam_trig,
          const LmTrig* __restrict__ lm_trig, int n_pose, double u, double v, double* __restrict__ resid,
          double* __restrict__ gU, double* __restrict__ gGc, double* __restrict__ gV, double* __restrict__ gGl,
          double* __restrict__ gCost) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double sWarp[kDualThreads / 32];
    const int64_t b = blockIdx.x, G = gridDim.x;
    const int lm_before = (int)(b * n_lm_ctas / G);
    const bool is_lm = (int)((b + 1) * n_lm_ctas / G) > lm_before;
    if ((is_lm && only == 2) || (!is_lm && only == 1)) return;     // role isolation for profiling (PTZBA_DUAL_ONLY)
    if (is_lm) {
        const int64_t begin = (int64_t)lm_before * chunk_lm;
        int64_t end = begin + chunk_lm;
        if (end > n_obs) end = n_obs;
        if (begin >= end) return;
        dual_fill_cam<CAMREP>(smem, cam_trig, n_pose);
        __syncthreads();
        dual_lm_role<CAMREP>(smem, sWarp, begin, end, s_cam, s_lm, s_ox, s_oy, orig, lm_trig, n_pose, u, v, resid, gV, gGl, gCost);
    } else {
        const int64_t begin = (b - lm_before) * chunk_cam;
        int64_t end = begin + chunk_cam;
        if (end > n_obs) end = n_obs;
        if (begin >= end) return;
        dual_cam_role(begin, end, c_cam, c_lm, c_ox, c_oy, cam_trig, lm_trig, u, v, gU, gGc);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// fused_variant 14: ONE pass over the landmark-major list; the geometry of an observation is evaluated exactly once.
//
// The landmark side (V, g_l) is reduced where it is adjacent, as in k_ba_lm_pass4.  The keyframe side (U, g_c) is
// transposed through shared memory with a STATIC schedule: the observation structure never changes between passes, so
// ptzba_ba_create sorts every tile of kOneTile consecutive observations by keyframe once (k_build_tiles) and stores
//   kslot[obs]           uint16  position of the observation in its tile's keyframe-sorted order
//   kptr[tile][N + 1]    uint16  start of every keyframe's run in that order
// Phase A (thread = 4 consecutive observations): projection, residual, analytic blocks, landmark sums; the six numbers the
//   keyframe side needs (d/d alpha, N/z, residual) are written to the staging slot kslot[obs].
// Phase B (thread = keyframe): walks its run of staging slots, forms the 9 unique entries of J_c^T J_c / J_c^T r in
//   registers and adds them to the CTA's private keyframe table (row owned by this thread: no atomics anywhere).
// The table is flushed with one RED per entry when the CTA has finished its chunk.
// Traffic: 24 B + 2 B (kslot) read and 16 B written per observation, (N + 1) * 2 B per tile.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kOneThreads = 384;
constexpr int kOneTile = kOneThreads * kQuad;      // 1536 observations
constexpr int kOneWarpTrig = 16;                   // landmark ids are contiguous inside a warp's 128 observations

// one CTA per tile: counting sort of the tile's observations by keyframe (deterministic: stable inside a keyframe)
__global__ void __launch_bounds__(256)
k_build_tiles(int64_t n_obs, int64_t chunk, int tiles_per_cta, int n_pose, const int32_t* __restrict__ s_cam,
              uint16_t* __restrict__ kslot, uint16_t* __restrict__ kptr) {
    extern __shared__ int sb[];
    int* sKey = sb;                       // [kOneTile]
    int* sCnt = sb + kOneTile;            // [n_pose + 1]
    const int tile = blockIdx.x, cta = tile / tiles_per_cta, t = tile - cta * tiles_per_cta;
    const int64_t cbeg = (int64_t)cta * chunk;
    int64_t cend = cbeg + chunk;
    if (cend > n_obs) cend = n_obs;
    const int64_t tbeg = cbeg + (int64_t)t * kOneTile;
    int64_t cnt64 = cend - tbeg;
    if (cnt64 > kOneTile) cnt64 = kOneTile;
    const int cnt = cnt64 < 0 ? 0 : (int)cnt64;
    for (int i = threadIdx.x; i <= n_pose; i += blockDim.x) sCnt[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const int c = s_cam[tbeg + i];
        sKey[i] = c;
        atomicAdd(&sCnt[c + 1], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int c = 0; c <= n_pose; ++c) { run += sCnt[c]; sCnt[c] = run; }      // sCnt[c] = start of keyframe c
    }
    __syncthreads();
    uint16_t* kp = kptr + (size_t)tile * (n_pose + 1);
    for (int c = threadIdx.x; c <= n_pose; c += blockDim.x) kp[c] = (uint16_t)sCnt[c];
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const int c = sKey[i];
        int rank = 0;
        for (int j = 0; j < i; ++j) rank += (sKey[j] == c);
        kslot[tbeg + i] = (uint16_t)(sCnt[c] + rank);
    }
}

struct OneSmem {            // carve-up of the dynamic shared memory of k_ba_onepass
    double2* p0;            // [kOneTile] (d x / d alpha, d y / d alpha)
    double2* p1;            // [kOneTile] (Nx / z, Ny / z)
    double2* p2;            // [kOneTile] (r_x, r_y)
    double* cam;            // [n_pose * 6] keyframe trig rows {sp,cp,st,ct,f,-}
    double* tab;            // [n_pose * 9] keyframe accumulators, radian units
    uint16_t* kptr;         // [n_pose + 1]
    LmTrig* wtrig;          // [warps][kOneWarpTrig] per-warp cache of the landmark trig rows of the warp's 128 observations
};

__device__ __forceinline__ OneSmem one_smem(unsigned char* base, int n_pose) {
    OneSmem s;
    s.p0 = reinterpret_cast<double2*>(base);
    s.p1 = s.p0 + kOneTile;
    s.p2 = s.p1 + kOneTile;
    s.cam = reinterpret_cast<double*>(s.p2 + kOneTile);
    s.tab = s.cam + (size_t)n_pose * 6;
    s.wtrig = reinterpret_cast<LmTrig*>(s.cam + ((size_t)n_pose * 15 + 3) / 4 * 4);      // 32-byte aligned
    s.kptr = reinterpret_cast<uint16_t*>(s.wtrig + (kOneThreads / 32) * kOneWarpTrig);
    return s;
}

static size_t one_smem_bytes(int n_pose) {
    return (size_t)kOneTile * 48 + ((size_t)n_pose * 15 + 3) / 4 * 4 * sizeof(double) + (kOneThreads / 32) * kOneWarpTrig * sizeof(LmTrig) +
           ((size_t)n_pose + 1 + 7) / 8 * 8 * sizeof(uint16_t);
}

__global__ void __launch_bounds__(kOneThreads, 2)
k_ba_onepass(int64_t n_obs, int64_t chunk, int tiles_per_cta, int debug, const int32_t* __restrict__ s_cam,
             const int32_t* __restrict__ s_lm, const double* __restrict__ s_ox, const double* __restrict__ s_oy,
             const int32_t* __restrict__ orig, const uint16_t* __restrict__ kslot, const uint16_t* __restrict__ kptr,
             const CamTrig* __restrict__ cam_trig, const LmTrig* __restrict__ lm_trig, int n_pose, int n_lm, double u,
             double v, double* __restrict__ resid, double* __restrict__ gU, double* __restrict__ gGc,
             double* __restrict__ gV, double* __restrict__ gGl, double* __restrict__ gCost) {
    extern __shared__ __align__(32) unsigned char dyn[];
    __shared__ double sWarp[kOneThreads / 32];
    const OneSmem sm = one_smem(dyn, n_pose);
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t cbeg = (int64_t)blockIdx.x * chunk;
    int64_t cend = cbeg + chunk;
    if (cend > n_obs) cend = n_obs;
    if (cbeg >= cend) return;
    for (int i = tid; i < n_pose * 5; i += kOneThreads) {
        const int c = i / 5, e = i - 5 * c;
        sm.cam[(size_t)c * 6 + e] = reinterpret_cast<const double*>(cam_trig)[i];
    }
    for (int i = tid; i < n_pose * 9; i += kOneThreads) sm.tab[i] = 0.0;
    const double k1 = PTZ_DEG2RAD, k2 = PTZ_DEG2RAD * PTZ_DEG2RAD;
    double cost = 0.0;
    int tile = blockIdx.x * tiles_per_cta;
    // software pipeline across the tile loop: the quad of tile i+1 is requested right after phase A of tile i (its
    // registers are dead during phase B), the two landmark trig rows it needs right after phase B
    int4 cq = make_int4(0, 0, 0, 0), lq = make_int4(-1, -1, -1, -1);
    D4 xq = {0, 0, 0, 0}, yq = {0, 0, 0, 0};
    ushort4 sq = make_ushort4(0, 0, 0, 0);
    const LmTrig unitT = {0, 1, 0, 1};
    LmTrig* wTrig = sm.wtrig + (tid >> 5) * kOneWarpTrig;
    LmTrig tpre = unitT;        // row (wfirst + lane) of the trig table, lanes < kOneWarpTrig
    int wfirst = -1;            // first landmark id of the warp's quads in the coming tile
    // arrays are padded by >= 4 entries: a quad that starts inside the chunk is always loadable, its tail is masked
#define ONE_LOAD_QUAD(K0)                                                                  \
    do {                                                                                   \
        const int64_t kk = (K0);                                                           \
        lq = make_int4(-1, -1, -1, -1);                                                    \
        if (kk < cend) {                                                                   \
            cq = __ldg(reinterpret_cast<const int4*>(s_cam + kk));                         \
            lq = __ldg(reinterpret_cast<const int4*>(s_lm + kk));                          \
            xq = *reinterpret_cast<const D4*>(s_ox + kk);                                  \
            yq = *reinterpret_cast<const D4*>(s_oy + kk);                                  \
            sq = __ldg(reinterpret_cast<const ushort4*>(kslot + kk));                      \
            if (kk + 1 >= cend) lq.y = -1;                                                 \
            if (kk + 2 >= cend) lq.z = -1;                                                 \
            if (kk + 3 >= cend) lq.w = -1;                                                 \
        }                                                                                  \
    } while (0)
#define ONE_LOAD_TRIGS()                                                                   \
    do {                                                                                   \
        wfirst = __shfl_sync(0xffffffffu, lq.x, 0);                                        \
        if (lane < kOneWarpTrig && wfirst >= 0 && wfirst + lane < n_lm) tpre = lm_trig[wfirst + lane];   \
    } while (0)
    ONE_LOAD_QUAD(cbeg + (int64_t)tid * kQuad);
    ONE_LOAD_TRIGS();
    for (int64_t tbeg = cbeg; tbeg < cend; tbeg += kOneTile, ++tile) {
        // ---- phase A: this thread's quad ----
        const uint16_t* kp = kptr + (size_t)tile * (n_pose + 1);
        for (int i = tid; i <= n_pose; i += kOneThreads) sm.kptr[i] = kp[i];
        const int64_t k0 = tbeg + (int64_t)tid * kQuad;
        const int cam[kQuad] = {cq.x, cq.y, cq.z, cq.w}, lm[kQuad] = {lq.x, lq.y, lq.z, lq.w};
        const int slot[kQuad] = {sq.x, sq.y, sq.z, sq.w};
        const double ox[kQuad] = {xq.a, xq.b, xq.c, xq.d}, oy[kQuad] = {yq.a, yq.b, yq.c, yq.d};
        double rx[kQuad], ry[kQuad];
        int cur = -1;
        LmTrig lt = unitT;
        double vtt = 0, vtp = 0, vpp = 0, glt = 0, glp = 0;
        if (lane < kOneWarpTrig) wTrig[lane] = tpre;
        const int wf = wfirst;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < kQuad; ++i) {
            rx[i] = 0.0; ry[i] = 0.0;
            if (lm[i] < 0) continue;
            if (lm[i] != cur) {
                if (cur >= 0) commit_lm(gV, gGl, cur, vtt, vtp, vpp, glt, glp);   // run ended inside this thread
                cur = lm[i];
                lt = (cur - wf < kOneWarpTrig) ? wTrig[cur - wf] : lm_trig[cur];
                vtt = vtp = vpp = glt = glp = 0.0;
            }
            CamTrig c;
            {
                const double2* t = reinterpret_cast<const double2*>(sm.cam + (size_t)cam[i] * 6);
                const double2 pa = t[0], ti = t[1];
                c.sp = pa.x; c.cp = pa.y; c.st = ti.x; c.ct = ti.y; c.f = sm.cam[(size_t)cam[i] * 6 + 4];
            }
            double x, y;
            ObsGeom g;
            project_fast_jac(c, lt, u, v, x, y, g);
            rx[i] = x - ox[i];
            ry[i] = y - oy[i];
            if (!(debug & 2)) {
                sm.p0[slot[i]] = make_double2(g.xa, g.ya);
                sm.p1[slot[i]] = make_double2(g.px, g.py);
                sm.p2[slot[i]] = make_double2(rx[i], ry[i]);
            }
            cost = fma(rx[i], rx[i], fma(ry[i], ry[i], cost));
            const double kxa = k1 * g.xa, kya = k1 * g.ya, kxp = k1 * g.xp, kyp = k1 * g.yp;
            vtt = fma(kxa, kxa, fma(kya, kya, vtt));
            vtp = fma(kxa, kxp, fma(kya, kyp, vtp));
            vpp = fma(kxp, kxp, fma(kyp, kyp, vpp));
            glt = fma(kxa, rx[i], fma(kya, ry[i], glt));
            glp = fma(kxp, rx[i], fma(kyp, ry[i], glp));
        }
        if (resid) {
            if (!orig && k0 + kQuad <= cend) {
                D4* dst = reinterpret_cast<D4*>(resid + 2 * k0);
                dst[0] = D4{rx[0], ry[0], rx[1], ry[1]};
                dst[1] = D4{rx[2], ry[2], rx[3], ry[3]};
            } else {
#pragma unroll
                for (int i = 0; i < kQuad; ++i)
                    if (lm[i] >= 0) {
                        const int64_t o = orig ? (int64_t)orig[k0 + i] : k0 + i;
                        reinterpret_cast<double2*>(resid)[o] = make_double2(rx[i], ry[i]);
                    }
            }
        }
        seg_reduce5(cur, lane, vtt, vtp, vpp, glt, glp);
        const int prev = __shfl_up_sync(0xffffffffu, cur, 1);
        if (cur >= 0 && (lane == 0 || prev != cur)) commit_lm(gV, gGl, cur, vtt, vtp, vpp, glt, glp);
        ONE_LOAD_QUAD(k0 + kOneTile);      // next tile's quad: in flight during phase B
        __syncthreads();
        // ---- phase B: this thread's keyframes (keyframe 0 is the fixed reference pose: no block) ----
        for (int key = tid; key < n_pose && !(debug & 1); key += kOneThreads) {
            const int s0 = sm.kptr[key], s1 = sm.kptr[key + 1];
            if (key == 0 || s0 == s1) continue;
            const double f = sm.cam[(size_t)key * 6 + 4];
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0, a8 = 0;
            for (int s = s0; s < s1; ++s) {
                const double2 da = sm.p0[s], pz = sm.p1[s], r = sm.p2[s];
                const double xt = f * pz.x * pz.y;                 // f Nx Ny / z^2
                const double yt = fma(f * pz.y, pz.y, f);          // f (1 + Ny^2 / z^2)
                a0 = fma(da.x, da.x, fma(da.y, da.y, a0));
                a1 = fma(da.x, xt, fma(da.y, yt, a1));
                a2 = fma(da.x, pz.x, fma(da.y, pz.y, a2));
                a3 = fma(xt, xt, fma(yt, yt, a3));
                a4 = fma(xt, pz.x, fma(yt, pz.y, a4));
                a5 = fma(pz.x, pz.x, fma(pz.y, pz.y, a5));
                a6 = fma(da.x, r.x, fma(da.y, r.y, a6));
                a7 = fma(xt, r.x, fma(yt, r.y, a7));
                a8 = fma(pz.x, r.x, fma(pz.y, r.y, a8));
            }
            double* t = sm.tab + (size_t)key * 9;
            t[0] += a0; t[1] -= a1; t[2] -= a2; t[3] += a3; t[4] += a4; t[5] += a5; t[6] -= a6; t[7] += a7; t[8] += a8;
        }
        ONE_LOAD_TRIGS();
        __syncthreads();
    }
#undef ONE_LOAD_QUAD
#undef ONE_LOAD_TRIGS
    cost = warp_sum(cost);
    if (lane == 0) sWarp[tid >> 5] = cost;
    __syncthreads();
    if (tid == 0) {
        double s = 0;
        for (int w = 0; w < kOneThreads / 32; ++w) s += sWarp[w];
        atomicAdd(gCost, s);
    }
    // flush the keyframe table (converted to per-degree units): entries (pp,pt,pf,tt,tf,ff | gp,gt,gf)
    for (int i = tid; i < n_pose * 9; i += kOneThreads) {
        const double val = sm.tab[i];
        if (val == 0.0) continue;
        const int c = i / 9, e = i - 9 * c;
        const double sc = (e == 0 || e == 1 || e == 3) ? k2 : (e == 2 || e == 4 || e == 6 || e == 7) ? k1 : 1.0;
        if (e < 6) atomicAdd(gU + 6 * (size_t)c + e, val * sc);
        else atomicAdd(gGc + 3 * (size_t)c + (e - 6), val * sc);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// fused_variant 9: the two coherent passes fed by a TMA (cp.async.bulk) ring buffer.
// One elected thread per CTA streams 1024-observation tiles of the four SoA arrays (and, in the landmark-major pass, the
// tile's contiguous slice of the landmark trig table) into shared memory, kStages tiles ahead, completing on an mbarrier;
// all warps consume the current tile from shared memory (conflict-free 128-bit LDS).  HBM latency is therefore hidden by
// the copy engine instead of by occupancy, and the streaming loads cost no LSU instruction slots.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTile = kFusedThreads * kQuad;     // 1024 observations
constexpr int kStages = 2;
constexpr int kTileTrig = 128;                   // landmark trig entries staged per tile (more -> direct global loads)

struct __align__(128) TileBuf {
    int cam[kTile];
    int lm[kTile];
    double ox[kTile];
    double oy[kTile];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// producer side: issue the copies of global tile `t` into ring slot `slot`
__device__ __forceinline__ void issue_tile(TileBuf* buf, uint64_t* bar, int64_t t, int64_t n_obs, const int32_t* cam,
                                           const int32_t* lm, const double* ox, const double* oy, LmTrig* trig_dst,
                                           const LmTrig* lm_trig, const int2* tile_lm, int* trig_info) {
    const int64_t k0 = t * kTile;
    int64_t cnt = n_obs - k0;
    if (cnt > kTile) cnt = kTile;
    const uint32_t c4 = (uint32_t)((cnt + 3) / 4 * 4);          // arrays are padded by 4 entries
    uint32_t bytes = c4 * 4u * 2u + c4 * 8u * 2u;
    int n_trig = 0, lo = 0;
    if (trig_dst) {
        const int2 r = tile_lm[t];                              // (first landmark id, number of landmark ids) of the tile
        lo = r.x;
        if (r.y <= kTileTrig) { n_trig = r.y; bytes += (uint32_t)n_trig * (uint32_t)sizeof(LmTrig); }
        trig_info[0] = lo;
        trig_info[1] = n_trig;
    }
    mbar_expect_tx(bar, bytes);
    tma_load(buf->cam, cam + k0, c4 * 4u, bar);
    tma_load(buf->lm, lm + k0, c4 * 4u, bar);
    tma_load(buf->ox, ox + k0, c4 * 8u, bar);
    tma_load(buf->oy, oy + k0, c4 * 8u, bar);
    if (n_trig > 0) tma_load(trig_dst, lm_trig + lo, (uint32_t)n_trig * (uint32_t)sizeof(LmTrig), bar);
}

__global__ void k_tile_lm_ranges(int64_t n_obs, int64_t n_tiles, const int32_t* __restrict__ s_lm, int2* __restrict__ tile_lm) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const int64_t k0 = t * kTile;
    int64_t k1 = k0 + kTile;
    if (k1 > n_obs) k1 = n_obs;
    const int lo = s_lm[k0], hi = s_lm[k1 - 1];
    tile_lm[t] = make_int2(lo, hi - lo + 1);
}

__global__ void __launch_bounds__(kFusedThreads, 3)
k_ba_lm_pass_tma(int64_t n_obs, int64_t n_tiles, int tiles_per_cta, const int32_t* __restrict__ s_cam,
                 const int32_t* __restrict__ s_lm, const double* __restrict__ s_ox, const double* __restrict__ s_oy,
                 const int32_t* __restrict__ orig, const CamTrig* __restrict__ cam_trig, const LmTrig* __restrict__ lm_trig,
                 const int2* __restrict__ tile_lm, int n_pose, double u, double v, double* __restrict__ resid,
                 double* __restrict__ gV, double* __restrict__ gGl, double* __restrict__ gCost) {
    extern __shared__ __align__(128) unsigned char dyn[];
    TileBuf* tiles = reinterpret_cast<TileBuf*>(dyn);
    LmTrig* trigs = reinterpret_cast<LmTrig*>(dyn + sizeof(TileBuf) * kStages);
    double* sCam = reinterpret_cast<double*>(dyn + sizeof(TileBuf) * kStages + sizeof(LmTrig) * kTileTrig * kStages);   // SoA [5][n_pose]
    __shared__ uint64_t full[kStages];
    __shared__ int trig_info[kStages][2];
    __shared__ double sWarp[kFusedThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t t_begin = (int64_t)blockIdx.x * tiles_per_cta;
    int64_t t_end = t_begin + tiles_per_cta;
    if (t_end > n_tiles) t_end = n_tiles;
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < n_pose * 5; i += kFusedThreads) {
        const int c = i / 5, e = i - 5 * c;
        sCam[(size_t)c * 6 + e] = reinterpret_cast<const double*>(cam_trig)[i];
    }
    __syncthreads();
    if (tid == 0)
        for (int i = 0; i < kStages - 1; ++i)
            if (t_begin + i < t_end)
                issue_tile(&tiles[i], &full[i], t_begin + i, n_obs, s_cam, s_lm, s_ox, s_oy, trigs + (size_t)i * kTileTrig,
                           lm_trig, tile_lm, trig_info[i]);
    const double k1 = PTZ_DEG2RAD;
    double cost = 0.0;
    for (int64_t t = t_begin; t < t_end; ++t) {
        const int it = (int)(t - t_begin);
        const int slot = it % kStages;
        if (tid == 0) {
            const int64_t tn = t + kStages - 1;
            if (tn < t_end) {
                const int ns = (it + kStages - 1) % kStages;
                issue_tile(&tiles[ns], &full[ns], tn, n_obs, s_cam, s_lm, s_ox, s_oy, trigs + (size_t)ns * kTileTrig, lm_trig,
                           tile_lm, trig_info[ns]);
            }
        }
        mbar_wait(&full[slot], (uint32_t)((it / kStages) & 1));
        const TileBuf& tb = tiles[slot];
        const LmTrig* sTrig = trigs + (size_t)slot * kTileTrig;
        const int tlo = trig_info[slot][0], tn_trig = trig_info[slot][1];
        const int64_t k0 = t * kTile + (int64_t)tid * kQuad;
        const int q = tid * kQuad;
        const int4 c4 = *reinterpret_cast<const int4*>(tb.cam + q);
        const int4 l4 = *reinterpret_cast<const int4*>(tb.lm + q);
        const double2 xa = *reinterpret_cast<const double2*>(tb.ox + q), xb = *reinterpret_cast<const double2*>(tb.ox + q + 2);
        const double2 ya = *reinterpret_cast<const double2*>(tb.oy + q), yb = *reinterpret_cast<const double2*>(tb.oy + q + 2);
        int cam[kQuad] = {c4.x, c4.y, c4.z, c4.w};
        int lm[kQuad] = {l4.x, l4.y, l4.z, l4.w};
        const double ox[kQuad] = {xa.x, xa.y, xb.x, xb.y};
        const double oy[kQuad] = {ya.x, ya.y, yb.x, yb.y};
#pragma unroll
        for (int i = 0; i < kQuad; ++i)
            if (k0 + i >= n_obs) lm[i] = -1;
        double rx[kQuad], ry[kQuad];
        int cur = -1;
        LmTrig lt = {0, 1, 0, 1};
        double vtt = 0, vtp = 0, vpp = 0, glt = 0, glp = 0;
#pragma unroll
        for (int i = 0; i < kQuad; ++i) {
            rx[i] = 0.0; ry[i] = 0.0;
            if (lm[i] < 0) continue;
            if (lm[i] != cur) {
                if (cur >= 0) commit_lm(gV, gGl, cur, vtt, vtp, vpp, glt, glp);
                cur = lm[i];
                lt = (tn_trig > 0) ? sTrig[cur - tlo] : lm_trig[cur];
                vtt = vtp = vpp = glt = glp = 0.0;
            }
            CamTrig c;
            {
                const double2* tq = reinterpret_cast<const double2*>(sCam + (size_t)cam[i] * 6);
                const double2 pa = tq[0], ti = tq[1];
                c.sp = pa.x; c.cp = pa.y; c.st = ti.x; c.ct = ti.y; c.f = sCam[(size_t)cam[i] * 6 + 4];
            }
            double x, y;
            ObsGeom g;
            project_fast_jac(c, lt, u, v, x, y, g);
            rx[i] = x - ox[i];
            ry[i] = y - oy[i];
            cost = fma(rx[i], rx[i], fma(ry[i], ry[i], cost));
            const double kxa = k1 * g.xa, kya = k1 * g.ya, kxp = k1 * g.xp, kyp = k1 * g.yp;
            vtt = fma(kxa, kxa, fma(kya, kya, vtt));
            vtp = fma(kxa, kxp, fma(kya, kyp, vtp));
            vpp = fma(kxp, kxp, fma(kyp, kyp, vpp));
            glt = fma(kxa, rx[i], fma(kya, ry[i], glt));
            glp = fma(kxp, rx[i], fma(kyp, ry[i], glp));
        }
        if (resid) {
            if (!orig && k0 + kQuad <= n_obs) {
                D4* dst = reinterpret_cast<D4*>(resid + 2 * k0);
                dst[0] = D4{rx[0], ry[0], rx[1], ry[1]};
                dst[1] = D4{rx[2], ry[2], rx[3], ry[3]};
            } else {
#pragma unroll
                for (int i = 0; i < kQuad; ++i)
                    if (lm[i] >= 0) {
                        const int64_t o = orig ? (int64_t)orig[k0 + i] : k0 + i;
                        reinterpret_cast<double2*>(resid)[o] = make_double2(rx[i], ry[i]);
                    }
            }
        }
        seg_reduce5(cur, lane, vtt, vtp, vpp, glt, glp);
        const int prev = __shfl_up_sync(0xffffffffu, cur, 1);
        if (cur >= 0 && (lane == 0 || prev != cur)) commit_lm(gV, gGl, cur, vtt, vtp, vpp, glt, glp);
        __syncthreads();      // every warp is done with this slot before the producer refills it
    }
    cost = warp_sum(cost);
    if (lane == 0) sWarp[tid >> 5] = cost;
    __syncthreads();
    if (tid == 0) {
        double s = 0;
        for (int w = 0; w < kFusedThreads / 32; ++w) s += sWarp[w];
        atomicAdd(gCost, s);
    }
}

__global__ void __launch_bounds__(kFusedThreads, 3)
k_ba_cam_pass_tma(int64_t n_obs, int64_t n_tiles, int tiles_per_cta, const int32_t* __restrict__ c_cam,
                  const int32_t* __restrict__ c_lm, const double* __restrict__ c_ox, const double* __restrict__ c_oy,
                  const CamTrig* __restrict__ cam_trig, const LmTrig* __restrict__ lm_trig, double u, double v,
                  double* __restrict__ gU, double* __restrict__ gGc) {
    extern __shared__ __align__(128) unsigned char dyn[];
    TileBuf* tiles = reinterpret_cast<TileBuf*>(dyn);
    __shared__ uint64_t full[kStages];
    const int tid = threadIdx.x, lane = tid & 31;
    const double k1 = PTZ_DEG2RAD, k2 = PTZ_DEG2RAD * PTZ_DEG2RAD;
    const int64_t t_begin = (int64_t)blockIdx.x * tiles_per_cta;
    int64_t t_end = t_begin + tiles_per_cta;
    if (t_end > n_tiles) t_end = n_tiles;
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int i = 0; i < kStages - 1; ++i)
            if (t_begin + i < t_end)
                issue_tile(&tiles[i], &full[i], t_begin + i, n_obs, c_cam, c_lm, c_ox, c_oy, nullptr, nullptr, nullptr, nullptr);
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0, a8 = 0;
    int wcam = -1;
    CamTrig wc = {0, 1, 0, 1, 1};
    auto flush = [&]() {
        if (wcam > 0) {
            a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3); a4 = warp_sum(a4);
            a5 = warp_sum(a5); a6 = warp_sum(a6); a7 = warp_sum(a7); a8 = warp_sum(a8);
            if (lane == 0) {
                double* U = gU + 6 * (size_t)wcam;
                double* G = gGc + 3 * (size_t)wcam;
                atomicAdd(U + 0, a0 * k2); atomicAdd(U + 1, a1 * k2); atomicAdd(U + 2, a2 * k1);
                atomicAdd(U + 3, a3 * k2); atomicAdd(U + 4, a4 * k1); atomicAdd(U + 5, a5);
                atomicAdd(G + 0, a6 * k1); atomicAdd(G + 1, a7 * k1); atomicAdd(G + 2, a8);
            }
        }
        a0 = a1 = a2 = a3 = a4 = a5 = a6 = a7 = a8 = 0.0;
    };
    for (int64_t t = t_begin; t < t_end; ++t) {
        const int it = (int)(t - t_begin);
        const int slot = it % kStages;
        if (tid == 0) {
            const int64_t tn = t + kStages - 1;
            if (tn < t_end) {
                const int ns = (it + kStages - 1) % kStages;
                issue_tile(&tiles[ns], &full[ns], tn, n_obs, c_cam, c_lm, c_ox, c_oy, nullptr, nullptr, nullptr, nullptr);
            }
        }
        mbar_wait(&full[slot], (uint32_t)((it / kStages) & 1));
        const TileBuf& tb = tiles[slot];
        const int64_t k0 = t * kTile + (int64_t)tid * kQuad;
        const int q = tid * kQuad;
        const int4 c4 = *reinterpret_cast<const int4*>(tb.cam + q);
        const int4 l4 = *reinterpret_cast<const int4*>(tb.lm + q);
        int cam[kQuad] = {c4.x, c4.y, c4.z, c4.w};
        const int lm[kQuad] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
        for (int i = 0; i < kQuad; ++i)
            if (k0 + i >= n_obs) cam[i] = -1;
        LmTrig lt[kQuad];
#pragma unroll
        for (int i = 0; i < kQuad; ++i) lt[i] = lm_trig[cam[i] >= 0 ? lm[i] : 0];     // four independent gathers in flight
        const double2 xa = *reinterpret_cast<const double2*>(tb.ox + q), xb = *reinterpret_cast<const double2*>(tb.ox + q + 2);
        const double2 ya = *reinterpret_cast<const double2*>(tb.oy + q), yb = *reinterpret_cast<const double2*>(tb.oy + q + 2);
        const double ox[kQuad] = {xa.x, xa.y, xb.x, xb.y};
        const double oy[kQuad] = {ya.x, ya.y, yb.x, yb.y};
        const int first = __shfl_sync(0xffffffffu, cam[0], 0);
        bool mine = true;
#pragma unroll
        for (int i = 0; i < kQuad; ++i) mine = mine && (cam[i] == first || cam[i] < 0);
        const bool uniform = __all_sync(0xffffffffu, mine) && first >= 0;
        if (uniform) {
            if (first != wcam) { flush(); wcam = first; wc = cam_trig[wcam]; }
        } else {
            flush();
            wcam = -1;
        }
#pragma unroll
        for (int i = 0; i < kQuad; ++i) {
            if (cam[i] <= 0) continue;
            const CamTrig c = uniform ? wc : cam_trig[cam[i]];
            double x, y;
            ObsGeom g;
            project_fast_jac(c, lt[i], u, v, x, y, g);
            const double rx = x - ox[i], ry = y - oy[i];
            const double upp = fma(g.xa, g.xa, g.ya * g.ya);
            const double upt = -fma(g.xa, g.xt, g.ya * g.yt);
            const double upf = -fma(g.xa, g.px, g.ya * g.py);
            const double utt = fma(g.xt, g.xt, g.yt * g.yt);
            const double utf = fma(g.xt, g.px, g.yt * g.py);
            const double uff = fma(g.px, g.px, g.py * g.py);
            const double gp = -fma(g.xa, rx, g.ya * ry);
            const double gt = fma(g.xt, rx, g.yt * ry);
            const double gf = fma(g.px, rx, g.py * ry);
            if (uniform) {
                a0 += upp; a1 += upt; a2 += upf; a3 += utt; a4 += utf; a5 += uff; a6 += gp; a7 += gt; a8 += gf;
            } else {
                double* U = gU + 6 * (size_t)cam[i];
                double* G = gGc + 3 * (size_t)cam[i];
                atomicAdd(U + 0, upp * k2); atomicAdd(U + 1, upt * k2); atomicAdd(U + 2, upf * k1);
                atomicAdd(U + 3, utt * k2); atomicAdd(U + 4, utf * k1); atomicAdd(U + 5, uff);
                atomicAdd(G + 0, gp * k1); atomicAdd(G + 1, gt * k1); atomicAdd(G + 2, gf);
            }
        }
        __syncthreads();
    }
    flush();
}

// global-accumulator variant leaves radian units in U/gc; this converts them in place
__global__ void k_scale_cam_blocks(int n_pose, double* __restrict__ U, double* __restrict__ gc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pose * 9) return;
    const int c = i / 9, e = i - 9 * c;
    const double k1 = PTZ_DEG2RAD, k2 = PTZ_DEG2RAD * PTZ_DEG2RAD;
    const double sc = (e == 0 || e == 1 || e == 3) ? k2 : (e == 2 || e == 4 || e == 6 || e == 7) ? k1 : 1.0;
    if (e < 6) U[6 * (size_t)c + e] *= sc; else gc[3 * (size_t)c + (e - 6)] *= sc;
}

// residual-only pass (trial points of the trust-region loop, and _compute_residual itself)
__global__ void __launch_bounds__(kFusedThreads)
k_ba_residual(int64_t n_obs, const int32_t* __restrict__ s_cam, const int32_t* __restrict__ s_lm,
              const double* __restrict__ s_ox, const double* __restrict__ s_oy, const int32_t* __restrict__ orig,
              const CamTrig* __restrict__ cam_trig, const LmTrig* __restrict__ lm_trig, double u, double v,
              double* __restrict__ resid, double* __restrict__ gSumSq) {
    __shared__ double sWarp[kFusedThreads / 32];
    double cost = 0.0;
    for (int64_t k = (int64_t)blockIdx.x * kFusedThreads + threadIdx.x; k < n_obs; k += (int64_t)gridDim.x * kFusedThreads) {
        const CamTrig c = cam_trig[s_cam[k]];
        const LmTrig l = lm_trig[s_lm[k]];
        double x, y;
        project_fast(c, l, u, v, x, y);
        const double rx = x - s_ox[k], ry = y - s_oy[k];
        if (resid) {
            const int64_t o = orig ? (int64_t)orig[k] : k;
            reinterpret_cast<double2*>(resid)[o] = make_double2(rx, ry);
        }
        cost = fma(rx, rx, fma(ry, ry, cost));
    }
    cost = warp_sum(cost);
    if ((threadIdx.x & 31) == 0) sWarp[threadIdx.x >> 5] = cost;
    __syncthreads();
    if (threadIdx.x == 0 && gSumSq) {
        double s = 0;
        for (int w = 0; w < kFusedThreads / 32; ++w) s += sWarp[w];
        atomicAdd(gSumSq, s);
    }
}

int stream_grid(ptzba_ctx* ctx, int64_t n, int threads, int per_sm) {
    int64_t g = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)ctx->sm_count * per_sm;
    if (g > cap) g = cap;
    return g < 1 ? 1 : (int)g;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// device-level passes
// ---------------------------------------------------------------------------------------------------------------
int ba_set_params(ptzba_ba* ba, const double* d_x, const double* d_ref_pose3) {
    ptzba_ctx* ctx = ba->ctx;
    const int n = ba->n_pose > ba->n_lm ? ba->n_pose : ba->n_lm;
    k_set_params<<<div_up(n, 256), 256, 0, ctx->stream>>>(ba->n_pose, ba->n_lm, d_x, d_ref_pose3, ba->poses.p,
                                                          ba->rays.p, ba->cam_trig.p, ba->lm_trig.p);
    KERNEL_POST(ctx);
    return PTZBA_OK;
}

int ba_fused_pass(ptzba_ba* ba, double* d_resid) {
    ptzba_ctx* ctx = ba->ctx;
    cudaStream_t s = ctx->stream;
    CU_CHECK(ctx, cudaMemsetAsync(ba->acc.base, 0, ba->acc.count * sizeof(double), s));
    if (ba->n_obs == 0) return PTZBA_OK;
    const int32_t* orig = ba->identity_perm ? nullptr : ba->orig.p;
    // contiguous chunk per CTA, multiple of the CTA width so that warps stay aligned to 32 observations
    int64_t chunk = (ba->n_obs + ba->fused_grid - 1) / ba->fused_grid;
    chunk = (chunk + kFusedThreads - 1) / kFusedThreads * kFusedThreads;
    const int grid = (int)((ba->n_obs + chunk - 1) / chunk);
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->profiling) {
        CU_CHECK(ctx, cudaEventCreate(&ev0));
        CU_CHECK(ctx, cudaEventCreate(&ev1));
        ctx->prof_events.push_back(ev0);
        ctx->prof_events.push_back(ev1);
        CU_CHECK(ctx, cudaEventRecord(ev0, s));
    }
    if (ba->fused_variant >= 1) {
#define LAUNCH_CM(P, B)                                                                                                    \
    k_ba_fused_cm<P, B><<<grid, kFusedThreads, 0, s>>>(ba->n_obs, chunk, ba->c_cam.p, ba->c_lm.p, ba->c_ox.p, ba->c_oy.p,     \
                                                       ba->c_orig.p, ba->cam_trig.p, ba->lm_trig.p, ba->u, ba->v, d_resid,  \
                                                       ba->acc.U, ba->acc.gc, ba->acc.V, ba->acc.gl, ba->acc.cost)
        switch (ba->fused_variant) {
            case 1: LAUNCH_CM(0, 2); break;
            case 2: LAUNCH_CM(1, 2); break;
            case 9: {
                const int64_t n_tiles = (ba->n_obs + kTile - 1) / kTile;
                const size_t smA = sizeof(TileBuf) * kStages + sizeof(LmTrig) * kTileTrig * kStages + (size_t)ba->n_pose * 6 * sizeof(double);
                const size_t smB = sizeof(TileBuf) * kStages;
                int tpcA = (int)((n_tiles + ba->grid_tma_lm - 1) / ba->grid_tma_lm);
                int tpcB = (int)((n_tiles + ba->grid_tma_cam - 1) / ba->grid_tma_cam);
                const int gridA = (int)((n_tiles + tpcA - 1) / tpcA), gridB = (int)((n_tiles + tpcB - 1) / tpcB);
                k_ba_lm_pass_tma<<<gridA, kFusedThreads, smA, s>>>(ba->n_obs, n_tiles, tpcA, ba->s_cam.p, ba->s_lm.p, ba->s_ox.p,
                                                                 ba->s_oy.p, orig, ba->cam_trig.p, ba->lm_trig.p, ba->tile_lm.p,
                                                                 ba->n_pose, ba->u, ba->v, d_resid, ba->acc.V, ba->acc.gl, ba->acc.cost);
                ctx->launches++;
                k_ba_cam_pass_tma<<<gridB, kFusedThreads, smB, s>>>(ba->n_obs, n_tiles, tpcB, ba->c_cam.p, ba->c_lm.p, ba->c_ox.p,
                                                                  ba->c_oy.p, ba->cam_trig.p, ba->lm_trig.p, ba->u, ba->v, ba->acc.U,
                                                                  ba->acc.gc);
                break;
            }
            case 14: {
                k_ba_onepass<<<ba->one_grid, kOneThreads, one_smem_bytes(ba->n_pose), s>>>(
                    ba->n_obs, ba->one_chunk, ba->one_tiles_per_cta, ba->one_debug, ba->s_cam.p, ba->s_lm.p, ba->s_ox.p, ba->s_oy.p, orig,
                    ba->kslot.p, ba->kptr.p, ba->cam_trig.p, ba->lm_trig.p, ba->n_pose, ba->n_lm, ba->u, ba->v, d_resid, ba->acc.U,
                    ba->acc.gc, ba->acc.V, ba->acc.gl, ba->acc.cost);
                break;
            }
            case 13: {
                const int n_lm_ctas = ba->dual_lm_ctas, n_cam_ctas = ba->dual_grid - ba->dual_lm_ctas;
                const int64_t q = (int64_t)kDualThreads * kQuad;
                int64_t chunkA = (ba->n_obs + n_lm_ctas - 1) / n_lm_ctas;
                chunkA = (chunkA + q - 1) / q * q;
                int64_t chunkB = (ba->n_obs + n_cam_ctas - 1) / n_cam_ctas;
                chunkB = (chunkB + kDualThreads - 1) / kDualThreads * kDualThreads;
                const size_t sm = (size_t)ba->n_pose * (ba->dual_camrep == 8 ? 384 : 48);
#define LAUNCH_DUAL(REP)                                                                                                    \
    k_ba_dual<REP><<<ba->dual_grid, kDualThreads, sm, s>>>(ba->n_obs, n_lm_ctas, ba->dual_only, chunkA, chunkB, ba->s_cam.p, ba->s_lm.p,      \
                                                           ba->s_ox.p, ba->s_oy.p, orig, ba->c_cam.p, ba->c_lm.p, ba->c_ox.p, \
                                                           ba->c_oy.p, ba->cam_trig.p, ba->lm_trig.p, ba->n_pose, ba->u,     \
                                                           ba->v, d_resid, ba->acc.U, ba->acc.gc, ba->acc.V, ba->acc.gl,     \
                                                           ba->acc.cost)
                if (ba->dual_camrep == 8) LAUNCH_DUAL(8); else LAUNCH_DUAL(1);
#undef LAUNCH_DUAL
                break;
            }
            case 8: {
                // best measured combination so far: quad landmark-major pass + prefetching keyframe-major pass
                const int64_t q = (int64_t)kFusedThreads * kQuad;
                int64_t chunkA = (ba->n_obs + ba->grid_lm_pass4 - 1) / ba->grid_lm_pass4;
                chunkA = (chunkA + q - 1) / q * q;
                const int gridA = (int)((ba->n_obs + chunkA - 1) / chunkA);
                k_ba_lm_pass4<3><<<gridA, kFusedThreads, (size_t)ba->n_pose * 6 * sizeof(double), s>>>(
                    ba->n_obs, chunkA, ba->s_cam.p, ba->s_lm.p, ba->s_ox.p, ba->s_oy.p, orig, ba->cam_trig.p, ba->lm_trig.p,
                    ba->n_pose, ba->u, ba->v, d_resid, ba->acc.V, ba->acc.gl, ba->acc.cost, ba->l2_pf);
                ctx->launches++;
                int64_t chunkB = (ba->n_obs + ba->grid_cam_pass - 1) / ba->grid_cam_pass;
                chunkB = (chunkB + kFusedThreads - 1) / kFusedThreads * kFusedThreads;
                const int gridB = (int)((ba->n_obs + chunkB - 1) / chunkB);
                k_ba_cam_pass<3><<<gridB, kFusedThreads, 0, s>>>(ba->n_obs, chunkB, ba->c_cam.p, ba->c_lm.p, ba->c_ox.p, ba->c_oy.p,
                                                             ba->cam_trig.p, ba->lm_trig.p, ba->u, ba->v, ba->acc.U, ba->acc.gc,
                                                             ba->l2_pf);
                break;
            }
            case 7: {
                const int64_t q = (int64_t)kFusedThreads * kQuad;
                int64_t chunkA = (ba->n_obs + ba->grid_lm_pass4 - 1) / ba->grid_lm_pass4;
                chunkA = (chunkA + q - 1) / q * q;
                const int gridA = (int)((ba->n_obs + chunkA - 1) / chunkA);
                k_ba_lm_pass4<3><<<gridA, kFusedThreads, (size_t)ba->n_pose * 6 * sizeof(double), s>>>(
                    ba->n_obs, chunkA, ba->s_cam.p, ba->s_lm.p, ba->s_ox.p, ba->s_oy.p, orig, ba->cam_trig.p, ba->lm_trig.p,
                    ba->n_pose, ba->u, ba->v, d_resid, ba->acc.V, ba->acc.gl, ba->acc.cost, ba->l2_pf);
                ctx->launches++;
                int64_t chunkB = (ba->n_obs + ba->grid_cam_pass4 - 1) / ba->grid_cam_pass4;
                chunkB = (chunkB + q - 1) / q * q;
                const int gridB = (int)((ba->n_obs + chunkB - 1) / chunkB);
                k_ba_cam_pass4<3><<<gridB, kFusedThreads, 0, s>>>(ba->n_obs, chunkB, ba->c_cam.p, ba->c_lm.p, ba->c_ox.p, ba->c_oy.p,
                                                              ba->cam_trig.p, ba->lm_trig.p, ba->u, ba->v, ba->acc.U, ba->acc.gc);
                break;
            }
            case 6: {
                int64_t chunkA = (ba->n_obs + ba->grid_lm_pass - 1) / ba->grid_lm_pass;
                chunkA = (chunkA + kFusedThreads - 1) / kFusedThreads * kFusedThreads;
                const int gridA = (int)((ba->n_obs + chunkA - 1) / chunkA);
                k_ba_lm_pass<<<gridA, kFusedThreads, (size_t)ba->n_pose * 5 * sizeof(double), s>>>(
                    ba->n_obs, chunkA, ba->s_cam.p, ba->s_lm.p, ba->s_ox.p, ba->s_oy.p, orig, ba->cam_trig.p, ba->lm_trig.p,
                    ba->n_pose, ba->u, ba->v, d_resid, ba->acc.V, ba->acc.gl, ba->acc.cost);
                ctx->launches++;
                int64_t chunkB = (ba->n_obs + ba->grid_cam_pass - 1) / ba->grid_cam_pass;
                chunkB = (chunkB + kFusedThreads - 1) / kFusedThreads * kFusedThreads;
                const int gridB = (int)((ba->n_obs + chunkB - 1) / chunkB);
                k_ba_cam_pass<3><<<gridB, kFusedThreads, 0, s>>>(ba->n_obs, chunkB, ba->c_cam.p, ba->c_lm.p, ba->c_ox.p, ba->c_oy.p,
                                                             ba->cam_trig.p, ba->lm_trig.p, ba->u, ba->v, ba->acc.U, ba->acc.gc,
                                                             ba->l2_pf);
                break;
            }
            default: {
                // two coherent passes: landmark-major (r, V, g_l, cost) then keyframe-major (U, g_c); no scattered sums
                int64_t chunkA = (ba->n_obs + ba->fused_grid_lm - 1) / ba->fused_grid_lm;
                chunkA = (chunkA + kFusedThreads - 1) / kFusedThreads * kFusedThreads;
                const int gridA = (int)((ba->n_obs + chunkA - 1) / chunkA);
                k_ba_fused<true, false><<<gridA, kFusedThreads, ba->fused_smem, s>>>(
                    ba->n_obs, chunkA, ba->s_cam.p, ba->s_lm.p, ba->s_ox.p, ba->s_oy.p, orig, ba->cam_trig.p, ba->lm_trig.p,
                    ba->n_pose, ba->u, ba->v, d_resid, ba->acc.U, ba->acc.gc, ba->acc.V, ba->acc.gl, ba->acc.cost);
                ctx->launches++;
                LAUNCH_CM(2, 3);
                break;
            }
        }
#undef LAUNCH_CM
        KERNEL_POST(ctx);
        if (ev1) CU_CHECK(ctx, cudaEventRecord(ev1, s));
    } else if (ba->fused_cam_smem) {
        k_ba_fused<true, true><<<grid, kFusedThreads, ba->fused_smem, s>>>(
            ba->n_obs, chunk, ba->s_cam.p, ba->s_lm.p, ba->s_ox.p, ba->s_oy.p, orig, ba->cam_trig.p, ba->lm_trig.p,
            ba->n_pose, ba->u, ba->v, d_resid, ba->acc.U, ba->acc.gc, ba->acc.V, ba->acc.gl, ba->acc.cost);
        KERNEL_POST(ctx);
        if (ev1) CU_CHECK(ctx, cudaEventRecord(ev1, s));
    } else {
        k_ba_fused<false, true><<<grid, kFusedThreads, 0, s>>>(
            ba->n_obs, chunk, ba->s_cam.p, ba->s_lm.p, ba->s_ox.p, ba->s_oy.p, orig, ba->cam_trig.p, ba->lm_trig.p,
            ba->n_pose, ba->u, ba->v, d_resid, ba->acc.U, ba->acc.gc, ba->acc.V, ba->acc.gl, ba->acc.cost);
        KERNEL_POST(ctx);
        if (ev1) CU_CHECK(ctx, cudaEventRecord(ev1, s));
        k_scale_cam_blocks<<<div_up(ba->n_pose * 9, 256), 256, 0, s>>>(ba->n_pose, ba->acc.U, ba->acc.gc);
        KERNEL_POST(ctx);
    }
    return PTZBA_OK;
}

int ba_residual_pass(ptzba_ba* ba, double* d_resid, double* d_sumsq) {
    ptzba_ctx* ctx = ba->ctx;
    cudaStream_t s = ctx->stream;
    if (d_sumsq) CU_CHECK(ctx, cudaMemsetAsync(d_sumsq, 0, sizeof(double), s));
    if (ba->n_obs == 0) return PTZBA_OK;
    const int32_t* orig = ba->identity_perm ? nullptr : ba->orig.p;
    k_ba_residual<<<stream_grid(ctx, ba->n_obs, kFusedThreads, 8), kFusedThreads, 0, s>>>(
        ba->n_obs, ba->s_cam.p, ba->s_lm.p, ba->s_ox.p, ba->s_oy.p, orig, ba->cam_trig.p, ba->lm_trig.p, ba->u, ba->v,
        d_resid, d_sumsq);
    KERNEL_POST(ctx);
    return PTZBA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------------------------
extern "C" int ptzba_ba_create(ptzba_ctx* ctx, int mem, int n_pose, int n_landmark, int64_t n_obs,
                               const int32_t* cam_idx, const int32_t* lm_idx, const double* obs_xy, double u, double v,
                               ptzba_ba** out) {
    if (!ctx) return PTZBA_ERR_ARG;
    ARG_CHECK(ctx, out && n_pose >= 1 && n_landmark >= 0 && n_obs >= 0 && n_obs < (int64_t)2147483000);
    ARG_CHECK(ctx, n_obs == 0 || (cam_idx && lm_idx && obs_xy));
    *out = nullptr;
    cudaStream_t s = ctx->stream;
    ptzba_ba* ba = new ptzba_ba();
    ba->ctx = ctx; ba->n_pose = n_pose; ba->n_lm = n_landmark; ba->n_obs = n_obs; ba->u = u; ba->v = v;
    auto fail = [&](int code) { delete ba; return code; };
#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return fail(ptzba_fail(ctx, PTZBA_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,   \
                                   cudaGetErrorString(e__)));                                          \
    } while (0)
    InArray<int32_t> in_cam, in_lm;
    InArray<double> in_xy;
    CU_TRY(in_cam.stage(mem, cam_idx, (size_t)n_obs, s));
    CU_TRY(in_lm.stage(mem, lm_idx, (size_t)n_obs, s));
    CU_TRY(in_xy.stage(mem, obs_xy, (size_t)n_obs * 2, s));
    CU_TRY(ba->s_cam.alloc(n_obs + 4)); CU_TRY(ba->s_lm.alloc(n_obs + 4));
    CU_TRY(ba->s_ox.alloc(n_obs + 4)); CU_TRY(ba->s_oy.alloc(n_obs + 4));
    CU_TRY(ba->lm_ptr.alloc((size_t)n_landmark + 1));
    CU_TRY(ba->poses.alloc((size_t)n_pose * 3)); CU_TRY(ba->rays.alloc((size_t)n_landmark * 2));
    CU_TRY(ba->cam_trig.alloc(n_pose)); CU_TRY(ba->lm_trig.alloc(n_landmark));
    ba->acc.count = 1 + (size_t)n_pose * 9 + (size_t)n_landmark * 5;
    CU_TRY(ba->accum_store.alloc(ba->acc.count));
    ba->acc.base = ba->accum_store.p;
    ba->acc.cost = ba->acc.base;
    // [cost | U | V | gc | gl]: gc and gl are adjacent so that (gc, gl) is the gradient in the solver's full layout
    ba->acc.U = ba->acc.base + 1;
    ba->acc.V = ba->acc.U + (size_t)n_pose * 6;
    ba->acc.gc = ba->acc.V + (size_t)n_landmark * 3;
    ba->acc.gl = ba->acc.gc + (size_t)n_pose * 3;
    CU_TRY(ba->scal.alloc(64));

    DevBuf<int> flags;
    CU_TRY(flags.alloc(2));
    CU_TRY(cudaMemsetAsync(flags.p, 0, 2 * sizeof(int), s));
    int h_flags[2] = {0, 0};
    if (n_obs > 0) {
        k_check_ranges<<<stream_grid(ctx, n_obs, 256, 8), 256, 0, s>>>(n_obs, in_cam.d, in_lm.d, n_pose, n_landmark, flags.p);
        ctx->launches++;
        CU_TRY(cudaMemcpyAsync(h_flags, flags.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        if (h_flags[0] & 1) return fail(ptzba_fail(ctx, PTZBA_ERR_ARG, "observation index out of range"));
        ba->identity_perm = !(h_flags[0] & 2);
        if (ba->identity_perm) {
            CU_TRY(cudaMemcpyAsync(ba->s_lm.p, in_lm.d, (size_t)n_obs * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
            k_gather_obs<<<stream_grid(ctx, n_obs, 256, 8), 256, 0, s>>>(n_obs, nullptr, in_cam.d, in_xy.d, ba->s_cam.p,
                                                                         ba->s_ox.p, ba->s_oy.p);
            ctx->launches++;
        } else {
            // stable LSD radix sort by landmark id keeps the caller's order inside a landmark
            DevBuf<int32_t> iota;
            DevBuf<unsigned char> tmp;
            CU_TRY(ba->orig.alloc(n_obs));
            CU_TRY(iota.alloc(n_obs));
            k_iota<<<stream_grid(ctx, n_obs, 256, 8), 256, 0, s>>>(n_obs, iota.p);
            ctx->launches++;
            size_t bytes = 0;
            int end_bit = 1;
            while ((1ll << end_bit) < (long long)n_landmark && end_bit < 31) ++end_bit;
            CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, in_lm.d, ba->s_lm.p, iota.p, ba->orig.p, (int)n_obs, 0,
                                                   end_bit, s));
            CU_TRY(tmp.alloc(bytes));
            CU_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, in_lm.d, ba->s_lm.p, iota.p, ba->orig.p, (int)n_obs, 0,
                                                   end_bit, s));
            k_gather_obs<<<stream_grid(ctx, n_obs, 256, 8), 256, 0, s>>>(n_obs, ba->orig.p, in_cam.d, in_xy.d, ba->s_cam.p,
                                                                         ba->s_ox.p, ba->s_oy.p);
            ctx->launches++;
            CU_TRY(cudaStreamSynchronize(s));
        }
    }
    CU_TRY(cudaMemsetAsync(flags.p + 1, 0, sizeof(int), s));
    k_lm_ptr<<<div_up(n_landmark + 1, 256), 256, 0, s>>>(n_landmark, n_obs, ba->s_lm.p, ba->lm_ptr.p, flags.p + 1);
    ctx->launches++;
    CU_TRY(cudaMemcpyAsync(h_flags + 1, flags.p + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    CU_TRY(cudaGetLastError());
    ba->max_degree = h_flags[1];

    // keyframe-major copy: stable sort of the landmark-major list by keyframe id
    {
        const char* env = getenv("PTZBA_FUSED_VARIANT");
        if (env) ba->fused_variant = atoi(env);
        if (const char* e = getenv("PTZBA_L2_PREFETCH")) ba->l2_pf = atoi(e);
    }
    if (n_obs > 0) {
        DevBuf<int32_t> iota, perm2;
        DevBuf<unsigned char> tmp;
        CU_TRY(ba->c_cam.alloc(n_obs + 4)); CU_TRY(ba->c_lm.alloc(n_obs + 4)); CU_TRY(ba->c_orig.alloc(n_obs + 4));
        CU_TRY(ba->c_ox.alloc(n_obs + 4)); CU_TRY(ba->c_oy.alloc(n_obs + 4));
        CU_TRY(iota.alloc(n_obs)); CU_TRY(perm2.alloc(n_obs));
        k_iota<<<stream_grid(ctx, n_obs, 256, 8), 256, 0, s>>>(n_obs, iota.p);
        ctx->launches++;
        size_t bytes = 0;
        int end_bit = 1;
        while ((1ll << end_bit) < (long long)n_pose && end_bit < 31) ++end_bit;
        CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, ba->s_cam.p, ba->c_cam.p, iota.p, perm2.p, (int)n_obs, 0, end_bit, s));
        CU_TRY(tmp.alloc(bytes));
        CU_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, ba->s_cam.p, ba->c_cam.p, iota.p, perm2.p, (int)n_obs, 0, end_bit, s));
        k_gather_cm<<<stream_grid(ctx, n_obs, 256, 8), 256, 0, s>>>(n_obs, perm2.p, ba->s_lm.p, ba->s_ox.p, ba->s_oy.p,
                                                                    ba->identity_perm ? nullptr : ba->orig.p, ba->c_lm.p,
                                                                    ba->c_ox.p, ba->c_oy.p, ba->c_orig.p);
        ctx->launches++;
        CU_TRY(cudaStreamSynchronize(s));
    }

    // launch geometry of the fused pass: one wave of resident CTAs; keyframe tables in shared memory when they fit
    const size_t smem = (size_t)n_pose * 14 * sizeof(double);
    int per_sm = 0;
    ba->fused_cam_smem = smem <= 200 * 1024;
    if (ba->fused_cam_smem) {
        CU_TRY(cudaFuncSetAttribute(k_ba_fused<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ba_fused<true, true>, kFusedThreads, smem));
        ba->fused_smem = (int)smem;
    } else {
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ba_fused<false, true>, kFusedThreads, 0));
        ba->fused_smem = 0;
    }
    if (ba->fused_variant >= 1) {
        {
            int pa = 1, pb = 1;
            const size_t sm5 = (size_t)n_pose * 6 * sizeof(double);
            if (sm5 > 40 * 1024) CU_TRY(cudaFuncSetAttribute(k_ba_lm_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm5));
            CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pa, k_ba_lm_pass, kFusedThreads, sm5));
            CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pb, k_ba_cam_pass<3>, kFusedThreads, 0));
            int pa4 = 1, pb4 = 1;
            if (sm5 > 40 * 1024) CU_TRY(cudaFuncSetAttribute(k_ba_lm_pass4<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm5));
            CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pa4, k_ba_lm_pass4<3>, kFusedThreads, sm5));
            CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pb4, k_ba_cam_pass4<3>, kFusedThreads, 0));
            ba->grid_lm_pass4 = ctx->sm_count * (pa4 < 1 ? 1 : pa4);
            ba->grid_cam_pass4 = ctx->sm_count * (pb4 < 1 ? 1 : pb4);
            {
                const size_t smA = sizeof(TileBuf) * kStages + sizeof(LmTrig) * kTileTrig * kStages + (size_t)n_pose * 6 * sizeof(double);
                const size_t smB = sizeof(TileBuf) * kStages;
                int qa = 1, qb = 1;
                if (smA <= 220 * 1024) {
                    CU_TRY(cudaFuncSetAttribute(k_ba_lm_pass_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA));
                    CU_TRY(cudaFuncSetAttribute(k_ba_cam_pass_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smB));
                    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&qa, k_ba_lm_pass_tma, kFusedThreads, smA));
                    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&qb, k_ba_cam_pass_tma, kFusedThreads, smB));
                } else if (ba->fused_variant == 9) {
                    ba->fused_variant = 8;
                }
                ba->grid_tma_lm = ctx->sm_count * (qa < 1 ? 1 : qa);
                ba->grid_tma_cam = ctx->sm_count * (qb < 1 ? 1 : qb);
                const int64_t n_tiles = (n_obs + kTile - 1) / kTile;
                CU_TRY(ba->tile_lm.alloc((size_t)n_tiles + 1));
                if (n_tiles > 0) {
                    k_tile_lm_ranges<<<div_up(n_tiles, 256), 256, 0, s>>>(n_obs, n_tiles, ba->s_lm.p, ba->tile_lm.p);
                    ctx->launches++;
                }
            }
            ba->grid_lm_pass = ctx->sm_count * (pa < 1 ? 1 : pa);
            ba->grid_cam_pass = ctx->sm_count * (pb < 1 ? 1 : pb);
        }
        {
            // dual-role kernel: 8x replicated keyframe trig table while it fits beside a second resident CTA
            ba->dual_camrep = ((size_t)n_pose * 384 <= 100 * 1024) ? 8 : 1;
            if (const char* e = getenv("PTZBA_DUAL_REP")) ba->dual_camrep = atoi(e) == 8 ? 8 : 1;
            const size_t smd = (size_t)n_pose * (ba->dual_camrep == 8 ? 384 : 48);
            int pd = 1;
            if (ba->dual_camrep == 8) {
                CU_TRY(cudaFuncSetAttribute(k_ba_dual<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smd));
                CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pd, k_ba_dual<8>, kDualThreads, smd));
            } else if (smd <= 220 * 1024) {
                CU_TRY(cudaFuncSetAttribute(k_ba_dual<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smd));
                CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pd, k_ba_dual<1>, kDualThreads, smd));
            } else if (ba->fused_variant == 13) {
                ba->fused_variant = 8;
            }
            if (pd < 2) pd = 2;                      // at least one CTA of each role per SM (queued if not co-resident)
            ba->dual_grid = ctx->sm_count * pd;
            int pct = 55;                            // share of the CTAs that take the landmark role
            if (const char* e = getenv("PTZBA_DUAL_SPLIT")) pct = atoi(e);
            if (pct < 5) pct = 5;
            if (pct > 95) pct = 95;
            ba->dual_lm_ctas = ba->dual_grid * pct / 100;
            if (const char* e = getenv("PTZBA_DUAL_ONLY")) ba->dual_only = atoi(e);
        }
        if (n_obs > 0) {
            // single-pass kernel: static per-tile keyframe sort (falls back to the two-pass kernels when the keyframe
            // tables do not fit in shared memory or a keyframe id does not fit the 16-bit run offsets)
            const size_t smo = one_smem_bytes(n_pose);
            if (smo <= 220 * 1024) {
                int po = 1;
                CU_TRY(cudaFuncSetAttribute(k_ba_onepass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smo));
                CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&po, k_ba_onepass, kOneThreads, smo));
                if (po < 1) po = 1;
                int grid = ctx->sm_count * po;
                int64_t chunk = (n_obs + grid - 1) / grid;
                chunk = (chunk + kQuad - 1) / kQuad * kQuad;
                grid = (int)((n_obs + chunk - 1) / chunk);
                ba->one_grid = grid;
                ba->one_chunk = chunk;
                ba->one_tiles_per_cta = (int)((chunk + kOneTile - 1) / kOneTile);
                if (const char* e = getenv("PTZBA_ONE_DEBUG")) ba->one_debug = atoi(e);
                const size_t n_tiles = (size_t)grid * ba->one_tiles_per_cta;
                CU_TRY(ba->kslot.alloc((size_t)n_obs + 8));
                CU_TRY(ba->kptr.alloc(n_tiles * ((size_t)n_pose + 1)));
                const size_t smb = ((size_t)kOneTile + n_pose + 1) * sizeof(int);
                if (smb > 40 * 1024) CU_TRY(cudaFuncSetAttribute(k_build_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
                k_build_tiles<<<(unsigned)n_tiles, 256, smb, s>>>(n_obs, chunk, ba->one_tiles_per_cta, n_pose, ba->s_cam.p,
                                                                  ba->kslot.p, ba->kptr.p);
                ctx->launches++;
                CU_TRY(cudaGetLastError());
            } else if (ba->fused_variant == 14) {
                ba->fused_variant = 8;
            }
        }
        int per_sm_lm = 1;
        CU_TRY(cudaFuncSetAttribute(k_ba_fused<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_lm, k_ba_fused<true, false>, kFusedThreads, smem));
        ba->fused_grid_lm = ctx->sm_count * (per_sm_lm < 1 ? 1 : per_sm_lm);
        switch (ba->fused_variant) {
            case 1: CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ba_fused_cm<0, 2>, kFusedThreads, 0)); break;
            case 2: CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ba_fused_cm<1, 2>, kFusedThreads, 0)); break;
            default: CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ba_fused_cm<2, 3>, kFusedThreads, 0)); break;
        }
    }
    if (per_sm < 1) per_sm = 1;
    ba->fused_grid = ctx->sm_count * per_sm;
#undef CU_TRY
    *out = ba;
    return PTZBA_OK;
}

extern "C" void ptzba_ba_destroy(ptzba_ba* ba) {
    if (!ba) return;
    cudaStreamSynchronize(ba->ctx->stream);
    delete ba;
}

// Persistent staging (no allocation on the hot path): x and the reference pose are copied into ba->x_stage /
// ba->ref_stage; host residuals are produced into ba->resid and copied out.
static int ba_stage_inputs(ptzba_ba* ba, int mem, const double* x, const double* reference_pose3, const double** d_x) {
    ptzba_ctx* ctx = ba->ctx;
    cudaStream_t s = ctx->stream;
    const size_t nx = 3 * (size_t)(ba->n_pose - 1) + 2 * (size_t)ba->n_lm;
    CU_CHECK(ctx, ba->ref_stage.alloc(4));
    CU_CHECK(ctx, cudaMemcpyAsync(ba->ref_stage.p, reference_pose3, 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    if (mem == PTZBA_DEVICE) {
        *d_x = x;
    } else {
        CU_CHECK(ctx, ba->x_stage.alloc(nx + 2));
        if (nx) CU_CHECK(ctx, cudaMemcpyAsync(ba->x_stage.p, x, nx * sizeof(double), cudaMemcpyHostToDevice, s));
        *d_x = ba->x_stage.p;
    }
    return PTZBA_OK;
}

extern "C" int ptzba_ba_residual(ptzba_ba* ba, int mem, const double* x, const double* reference_pose3,
                                 double* residual) {
    if (!ba) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = ba->ctx;
    ARG_CHECK(ctx, x && reference_pose3 && residual);
    cudaStream_t s = ctx->stream;
    const double* d_x = nullptr;
    PROPAGATE(ba_stage_inputs(ba, mem, x, reference_pose3, &d_x));
    double* d_r = residual;
    if (mem == PTZBA_HOST) {
        CU_CHECK(ctx, ba->resid.alloc(2 * (size_t)ba->n_obs));
        d_r = ba->resid.p;
    }
    PROPAGATE(ba_set_params(ba, d_x, ba->ref_stage.p));
    PROPAGATE(ba_residual_pass(ba, d_r, nullptr));
    if (mem == PTZBA_HOST) {
        if (ba->n_obs) CU_CHECK(ctx, cudaMemcpyAsync(residual, d_r, 2 * (size_t)ba->n_obs * sizeof(double), cudaMemcpyDeviceToHost, s));
        CU_CHECK(ctx, cudaStreamSynchronize(s));
    }
    return PTZBA_OK;
}

extern "C" int ptzba_ba_normal_equations(ptzba_ba* ba, int mem, const double* x, const double* reference_pose3,
                                         double* residual, double* U, double* gc, double* V, double* gl, double* cost) {
    if (!ba) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = ba->ctx;
    ARG_CHECK(ctx, x && reference_pose3);
    cudaStream_t s = ctx->stream;
    const double* d_x = nullptr;
    PROPAGATE(ba_stage_inputs(ba, mem, x, reference_pose3, &d_x));
    double* d_r = residual;
    if (mem == PTZBA_HOST && residual) {
        CU_CHECK(ctx, ba->resid.alloc(2 * (size_t)ba->n_obs));
        d_r = ba->resid.p;
    }
    PROPAGATE(ba_set_params(ba, d_x, ba->ref_stage.p));
    PROPAGATE(ba_fused_pass(ba, d_r));
    const cudaMemcpyKind kind = mem == PTZBA_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (U) CU_CHECK(ctx, cudaMemcpyAsync(U, ba->acc.U, (size_t)ba->n_pose * 6 * sizeof(double), kind, s));
    if (gc) CU_CHECK(ctx, cudaMemcpyAsync(gc, ba->acc.gc, (size_t)ba->n_pose * 3 * sizeof(double), kind, s));
    if (V && ba->n_lm) CU_CHECK(ctx, cudaMemcpyAsync(V, ba->acc.V, (size_t)ba->n_lm * 3 * sizeof(double), kind, s));
    if (gl && ba->n_lm) CU_CHECK(ctx, cudaMemcpyAsync(gl, ba->acc.gl, (size_t)ba->n_lm * 2 * sizeof(double), kind, s));
    if (mem == PTZBA_HOST && residual && ba->n_obs)
        CU_CHECK(ctx, cudaMemcpyAsync(residual, d_r, 2 * (size_t)ba->n_obs * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (cost) {
        CU_CHECK(ctx, cudaMemcpyAsync(ctx->h_scalars, ba->acc.cost, sizeof(double), cudaMemcpyDeviceToHost, s));
        CU_CHECK(ctx, cudaStreamSynchronize(s));
        *cost = 0.5 * ctx->h_scalars[0];
    } else if (mem == PTZBA_HOST) {
        CU_CHECK(ctx, cudaStreamSynchronize(s));
    }
    return PTZBA_OK;
}

extern "C" int ptzba_ba_get_blocks(ptzba_ba* ba, double* U, double* gc, double* V, double* gl, double* cost) {
    if (!ba) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = ba->ctx;
    cudaStream_t s = ctx->stream;
    if (U) CU_CHECK(ctx, cudaMemcpyAsync(U, ba->acc.U, (size_t)ba->n_pose * 6 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (gc) CU_CHECK(ctx, cudaMemcpyAsync(gc, ba->acc.gc, (size_t)ba->n_pose * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (V && ba->n_lm) CU_CHECK(ctx, cudaMemcpyAsync(V, ba->acc.V, (size_t)ba->n_lm * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (gl && ba->n_lm) CU_CHECK(ctx, cudaMemcpyAsync(gl, ba->acc.gl, (size_t)ba->n_lm * 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    double sumsq = 0;
    if (cost) CU_CHECK(ctx, cudaMemcpyAsync(&sumsq, ba->acc.cost, sizeof(double), cudaMemcpyDeviceToHost, s));
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    if (cost) *cost = 0.5 * sumsq;
    return PTZBA_OK;
}
