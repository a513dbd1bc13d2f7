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
#ifndef PTZBA_LM_MINB
#define PTZBA_LM_MINB 3
#endif
#ifndef PTZBA_CAM_MINB
#define PTZBA_CAM_MINB 3
#endif
constexpr int kLmMinB = PTZBA_LM_MINB, kCamMinB = PTZBA_CAM_MINB;   // resident CTAs per SM the two passes are compiled for

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

// keyframe-major gather: position k takes landmark-major position perm[k]
__global__ void k_gather_cm(int64_t n, const int32_t* __restrict__ perm, const int32_t* __restrict__ s_lm,
                            const double* __restrict__ s_ox, const double* __restrict__ s_oy, int32_t* __restrict__ c_lm,
                            double* __restrict__ c_ox, double* __restrict__ c_oy) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t p = perm[k];
        c_lm[k] = s_lm[p];
        c_ox[k] = s_ox[p];
        c_oy[k] = s_oy[p];
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

// x = [pose_1..pose_{N-1}, landmarks]; pose_0 = reference pose (bundle_adjustment.py:57-59).
// ZERO: also clears the accumulator arena (this kernel already visits every keyframe and landmark), which saves the
// separate memset launch in front of the fused pass.
template <bool ZERO>
__global__ void k_set_params(int n_pose, int cam_lo, int cam_hi, int lm_lo, int lm_hi, const double* __restrict__ x,
                             const double* __restrict__ ref3, CamTrig* __restrict__ cam_trig, LmTrig* __restrict__ lm_trig,
                             double* __restrict__ aCost, double* __restrict__ aU, double* __restrict__ aGc,
                             double* __restrict__ aV, double* __restrict__ aGl) {
    // only the keyframes [cam_lo, cam_hi) and landmarks [lm_lo, lm_hi) that this problem's observations touch are visited
    // (everything for a whole problem; one pan sector's share in the keyframe-sharded multi-GPU mode)
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (ZERO && t == 0) aCost[0] = 0.0;
    const int c = cam_lo + t;
    if (c < cam_hi) {
        const double* src = (c == 0) ? ref3 : (x + 3 * (size_t)(c - 1));
        cam_trig[c] = make_cam_trig(src[0], src[1], src[2]);
        if (ZERO) {
#pragma unroll
            for (int e = 0; e < 6; ++e) aU[6 * (size_t)c + e] = 0.0;
#pragma unroll
            for (int e = 0; e < 3; ++e) aGc[3 * (size_t)c + e] = 0.0;
        }
    }
    const int l = lm_lo + t;
    if (l < lm_hi) {
        // x + 3(N-1) is only 8-byte aligned when N is even: scalar loads
        const double* lp = x + 3 * (size_t)(n_pose - 1) + 2 * (size_t)l;
        lm_trig[l] = make_lm_trig(lp[0], lp[1]);
        if (ZERO) {
            aV[3 * (size_t)l] = 0.0; aV[3 * (size_t)l + 1] = 0.0; aV[3 * (size_t)l + 2] = 0.0;
            aGl[2 * (size_t)l] = 0.0; aGl[2 * (size_t)l + 1] = 0.0;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// fused residual + Jacobian + normal-equation assembly
// ---------------------------------------------------------------------------------------------------------------
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

// ---------------------------------------------------------------------------------------------------------------
// The fused pass = two coherent passes; every sum is formed where its operands are adjacent, nothing is scattered.
//   k_ba_lm_pass4  landmark-major: residual (written), cost, per-landmark V / g_l (thread-serial over 4 consecutive
//                  observations, then a warp-segmented reduction of each thread's last run)
//   k_ba_cam_pass  keyframe-major: per-keyframe U / g_c live in registers across the whole chunk
// Why not one pass: FP64 has no native shared-memory atomic add and L2 REDs cost ~3.9 ns per sector chip-wide, so
// whichever side is scattered costs more than re-streaming 24 B/observation and re-evaluating the geometry.  The
// alternatives that were built, measured and removed (one pass with shared-memory CAS atomics, keyframe-major with
// L2 REDs, TMA/mbarrier tile rings, one launch with two CTA roles, one pass with a static per-tile keyframe sort) are
// tabulated in DESIGN.md section 5 with their ncu counters.
// ---------------------------------------------------------------------------------------------------------------
template <int MINB, bool LOOKAHEAD2 = false>
__global__ void __launch_bounds__(kFusedThreads, MINB)
k_ba_cam_pass(int64_t lo, int64_t hi, int64_t chunk, const int32_t* __restrict__ c_cam, const int32_t* __restrict__ c_lm,
              const double* __restrict__ c_ox, const double* __restrict__ c_oy, const CamTrig* __restrict__ cam_trig,
              const LmTrig* __restrict__ lm_trig, double u, double v, double* __restrict__ gU, double* __restrict__ gGc) {
    const int tid = threadIdx.x, lane = tid & 31;
    const double k1 = PTZ_DEG2RAD;
    const int64_t begin = lo + (int64_t)blockIdx.x * chunk;
    int64_t end = begin + chunk;
    if (end > hi) end = hi;
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
    // LOOKAHEAD2 (opt-in experiment PTZBA_CAM_LOOKAHEAD, not yet run on hardware): the landmark index is requested TWO steps
    // ahead, so that the dependent trig gather of the next step can issue at once - 29 % of this kernel's stall samples sit on
    // that gather's address waiting for an index requested one step earlier (profiles/r1_ncu_fused_source_stalls.txt)
    int lm_ahead = 0;
    if (LOOKAHEAD2 && k + kFusedThreads < end) lm_ahead = c_lm[k + kFusedThreads];
    for (int64_t base = begin; base < end; base += kFusedThreads) {
        const bool act = k < end;
        const int64_t kn = k + kFusedThreads;
        int ncam = -1, nlm = 0;
        double nox = 0, noy = 0;
        LmTrig nlt = {0, 1, 0, 1};
        if (LOOKAHEAD2) {
            if (kn < end) { nlm = lm_ahead; nlt = lm_trig[nlm]; ncam = c_cam[kn]; nox = c_ox[kn]; noy = c_oy[kn]; }
            if (kn + kFusedThreads < end) lm_ahead = c_lm[kn + kFusedThreads];
        } else if (kn < end) { ncam = c_cam[kn]; nlm = c_lm[kn]; nox = c_ox[kn]; noy = c_oy[kn]; nlt = lm_trig[nlm]; }
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

// 4 consecutive observations per thread: 128-/256-bit loads, quads 32-byte aligned (arrays are padded by 4 entries)
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

template <int MINB, bool CAM_SMEM, bool IDX_AHEAD = false>
__global__ void __launch_bounds__(kFusedThreads, MINB)
k_ba_lm_pass4(int64_t lo, int64_t hi, int64_t chunk, const int32_t* __restrict__ s_cam, const int32_t* __restrict__ s_lm,
              const double* __restrict__ s_ox, const double* __restrict__ s_oy, const int32_t* __restrict__ orig,
              const CamTrig* __restrict__ cam_trig, const LmTrig* __restrict__ lm_trig, int n_pose, double u, double v,
              double* __restrict__ resid, double* __restrict__ gV, double* __restrict__ gGl, double* __restrict__ gCost) {
    extern __shared__ __align__(16) double smem[];      // keyframe trig, 48 B per keyframe: {sp,cp} {st,ct} {f,-}: two LDS.128 + one LDS.64
    __shared__ double sWarp[kFusedThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31;
    if (CAM_SMEM) {
        for (int i = tid; i < n_pose * 5; i += kFusedThreads) {
            const int c = i / 5, e = i - 5 * c;
            smem[(size_t)c * 6 + e] = reinterpret_cast<const double*>(cam_trig)[i];
        }
        __syncthreads();
    }
    // this CTA's observations: [max(begin, lo), end); quads stay aligned to multiples of 4 of the array index
    const int64_t begin = (lo & ~(int64_t)3) + (int64_t)blockIdx.x * chunk;
    int64_t end = begin + chunk;
    if (end > hi) end = hi;
    const double k1 = PTZ_DEG2RAD;
    double cost = 0.0;
    // IDX_AHEAD (opt-in experiment PTZBA_LM_IDX_AHEAD, not yet run on hardware): the two index vectors of the NEXT iteration are
    // requested one iteration early (8 registers), so that an iteration starts with its indices present and its trig gather /
    // keyframe rows can go out at once; the observation vectors are still requested at the top and are only needed after the
    // projection.  19 % of this kernel's stall samples sit on the first use of the streamed indices.
    int4 c4n = make_int4(0, 0, 0, 0), l4n = make_int4(0, 0, 0, 0);
    if (IDX_AHEAD) {
        const int64_t kf = begin + (int64_t)tid * kQuad;
        if (kf >= lo && kf + kQuad <= end) {
            c4n = __ldg(reinterpret_cast<const int4*>(s_cam + kf));
            l4n = __ldg(reinterpret_cast<const int4*>(s_lm + kf));
        }
    }
    for (int64_t base = begin; base < end; base += kFusedThreads * kQuad) {
        const int64_t k0 = base + (int64_t)tid * kQuad;
        int cam[kQuad], lm[kQuad];
        double ox[kQuad], oy[kQuad];
        if (k0 >= lo && k0 + kQuad <= end) {
            int4 c4, l4;
            if (IDX_AHEAD) {
                c4 = c4n; l4 = l4n;
                const int64_t kn = k0 + (int64_t)kFusedThreads * kQuad;
                if (kn + kQuad <= end) {
                    c4n = __ldg(reinterpret_cast<const int4*>(s_cam + kn));
                    l4n = __ldg(reinterpret_cast<const int4*>(s_lm + kn));
                }
            } else {
                c4 = __ldg(reinterpret_cast<const int4*>(s_cam + k0));
                l4 = __ldg(reinterpret_cast<const int4*>(s_lm + k0));
            }
            const D4 x4 = *reinterpret_cast<const D4*>(s_ox + k0);
            const D4 y4 = *reinterpret_cast<const D4*>(s_oy + k0);
            cam[0] = c4.x; cam[1] = c4.y; cam[2] = c4.z; cam[3] = c4.w;
            lm[0] = l4.x; lm[1] = l4.y; lm[2] = l4.z; lm[3] = l4.w;
            ox[0] = x4.a; ox[1] = x4.b; ox[2] = x4.c; ox[3] = x4.d;
            oy[0] = y4.a; oy[1] = y4.b; oy[2] = y4.c; oy[3] = y4.d;
        } else {
#pragma unroll
            for (int i = 0; i < kQuad; ++i) {
                const bool in = k0 + i >= lo && k0 + i < end;
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
            if (CAM_SMEM) {
                const double2* t = reinterpret_cast<const double2*>(smem + (size_t)cam[i] * 6);
                const double2 pa = t[0], ti = t[1];
                c.sp = pa.x; c.cp = pa.y; c.st = ti.x; c.ct = ti.y; c.f = smem[(size_t)cam[i] * 6 + 4];
            } else {
                c = cam_trig[cam[i]];
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
            if (!orig && k0 >= lo && k0 + kQuad <= end) {
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

// ---------------------------------------------------------------------------------------------------------------
// EXPERIMENT, opt-in (PTZBA_FUSED_RING=1), NOT the default; parity-green on small problems on a B200, not yet timed: the same two passes with the four
// observation streams fed by the copy engine (cp.async.bulk + mbarrier) into PER-WARP two-slot shared-memory rings.
// Motivation: the default kernels are bound by exposed load latency of a few resident warps; register
// prefetch costs occupancy, and the earlier CTA-wide TMA ring (one __syncthreads per 1024-observation tile) put all warps
// in lock-step.  Here every warp owns its ring and its mbarriers, so warps stay free-running, the streaming loads cost no
// registers and no LSU issue slots, and a slot is refilled (by the warp's lane 0) as soon as the warp has copied it to
// registers - one to two groups of 128 observations ahead of the arithmetic.  Arithmetic, reduction and commit order are
// those of k_ba_lm_pass4 / k_ba_cam_pass.  Requires lo % 4 == 0 (TMA source alignment); the host falls back otherwise.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kGroup = 32 * kQuad;               // observations per warp step
constexpr int kRing = 2;                         // slots per warp
constexpr int kWarps = kFusedThreads / 32;
struct __align__(16) WarpSlot { int cam[kGroup]; int lm[kGroup]; double ox[kGroup]; double oy[kGroup]; };   // 3 KB

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
        "RING_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra RING_DONE;\n"
        "bra RING_WAIT;\n"
        "RING_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// one lane: start the copies of observations [k0, min(k0 + kGroup, end)) into `slot` (arrays are padded by 4 entries)
__device__ __forceinline__ void ring_issue(WarpSlot* slot, uint64_t* bar, int64_t k0, int64_t end, const int32_t* cam,
                                           const int32_t* lm, const double* ox, const double* oy) {
    int64_t cnt = end - k0;
    if (cnt > kGroup) cnt = kGroup;
    const uint32_t c4 = (uint32_t)((cnt + 3) & ~(int64_t)3);
    mbar_expect_tx(bar, c4 * 24u);
    tma_load(slot->cam, cam + k0, c4 * 4u, bar);
    tma_load(slot->lm, lm + k0, c4 * 4u, bar);
    tma_load(slot->ox, ox + k0, c4 * 8u, bar);
    tma_load(slot->oy, oy + k0, c4 * 8u, bar);
}

template <int MINB, bool DUAL_GATHER>
__global__ void __launch_bounds__(kFusedThreads, MINB)
k_ba_lm_pass_ring(int64_t lo, int64_t hi, int64_t chunk, const int32_t* __restrict__ s_cam, const int32_t* __restrict__ s_lm,
                  const double* __restrict__ s_ox, const double* __restrict__ s_oy, const int32_t* __restrict__ orig,
                  const CamTrig* __restrict__ cam_trig, const LmTrig* __restrict__ lm_trig, int n_pose, double u, double v,
                  double* __restrict__ resid, double* __restrict__ gV, double* __restrict__ gGl, double* __restrict__ gCost) {
    extern __shared__ __align__(16) unsigned char dyn[];      // cp.async.bulk needs 16-byte aligned destinations
    WarpSlot* slots = reinterpret_cast<WarpSlot*>(dyn);                                              // [kWarps][kRing]
    double* smem = reinterpret_cast<double*>(dyn + sizeof(WarpSlot) * kWarps * kRing);               // keyframe trig, 48 B each
    __shared__ uint64_t full[kWarps][kRing];
    __shared__ double sWarp[kWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t begin = lo + (int64_t)blockIdx.x * chunk;
    int64_t end = begin + chunk;
    if (end > hi) end = hi;
    // group j of this warp = observations [begin + (warp + kWarps j) kGroup, + kGroup): the CTA still streams its chunk front to back
    const int64_t n_groups = end > begin ? (end - begin + kGroup - 1) / kGroup : 0;
    const int n_mine = n_groups > warp ? (int)((n_groups - warp + kWarps - 1) / kWarps) : 0;
    WarpSlot* my = slots + warp * kRing;
    if (lane == 0) {
        for (int r = 0; r < kRing; ++r) mbar_init(&full[warp][r], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int j = 0; j < kRing && j < n_mine; ++j)
            ring_issue(&my[j], &full[warp][j], begin + ((int64_t)warp + (int64_t)kWarps * j) * kGroup, end, s_cam, s_lm, s_ox, s_oy);
    }
    for (int i = tid; i < n_pose * 5; i += kFusedThreads) {          // overlaps with the first copies
        const int c = i / 5, e = i - 5 * c;
        smem[(size_t)c * 6 + e] = reinterpret_cast<const double*>(cam_trig)[i];
    }
    __syncthreads();
    const double k1 = PTZ_DEG2RAD;
    double cost = 0.0;
#pragma unroll 1
    for (int j = 0; j < n_mine; ++j) {
        const int r = j % kRing;
        mbar_wait(&full[warp][r], (uint32_t)((j / kRing) & 1));
        const WarpSlot& ws = my[r];
        const int q = lane * kQuad;
        const int4 c4 = *reinterpret_cast<const int4*>(ws.cam + q);
        const int4 l4 = *reinterpret_cast<const int4*>(ws.lm + q);
        const double2 xa = *reinterpret_cast<const double2*>(ws.ox + q), xb = *reinterpret_cast<const double2*>(ws.ox + q + 2);
        const double2 ya = *reinterpret_cast<const double2*>(ws.oy + q), yb = *reinterpret_cast<const double2*>(ws.oy + q + 2);
        __syncwarp();                                                // every lane has its copy: the slot may be refilled
        const int64_t g0 = begin + ((int64_t)warp + (int64_t)kWarps * j) * kGroup;
        if (lane == 0 && j + kRing < n_mine)
            ring_issue(&my[r], &full[warp][r], g0 + (int64_t)kWarps * kRing * kGroup, end, s_cam, s_lm, s_ox, s_oy);
        const int64_t k0 = g0 + q;
        const int cam[kQuad] = {c4.x, c4.y, c4.z, c4.w};
        int lm[kQuad] = {l4.x, l4.y, l4.z, l4.w};
        const double ox[kQuad] = {xa.x, xa.y, xb.x, xb.y};
        const double oy[kQuad] = {ya.x, ya.y, yb.x, yb.y};
#pragma unroll
        for (int i = 0; i < kQuad; ++i)
            if (k0 + i >= end) lm[i] = -1;
        // the quad's landmark trig rows are gathered up front, both at once (a quad spans at most two landmarks unless a
        // landmark has fewer than 3 observations): the profile of k_ba_lm_pass4 shows 13 % of all stall samples on the first
        // use of a gather issued inside the serial loop below
        // (DUAL_GATHER costs 8 more live registers: 60 bytes of spills at the 80-register cap - to be decided by measurement)
        const int lmA = lm[0] >= 0 ? lm[0] : 0, lmB = lm[kQuad - 1] >= 0 ? lm[kQuad - 1] : lmA;
        LmTrig ltA = {0, 1, 0, 1}, ltB = {0, 1, 0, 1};
        if (DUAL_GATHER) { ltA = lm_trig[lmA]; ltB = lm_trig[lmB]; }
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
                if (DUAL_GATHER) lt = cur == lmA ? ltA : cur == lmB ? ltB : lm_trig[cur];
                else lt = lm_trig[cur];
                vtt = vtp = vpp = glt = glp = 0.0;
            }
            const double2* t = reinterpret_cast<const double2*>(smem + (size_t)cam[i] * 6);
            const double2 pa = t[0], ti = t[1];
            CamTrig c;
            c.sp = pa.x; c.cp = pa.y; c.st = ti.x; c.ct = ti.y; c.f = smem[(size_t)cam[i] * 6 + 4];
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
        seg_reduce5(cur, lane, vtt, vtp, vpp, glt, glp);
        const int prev = __shfl_up_sync(0xffffffffu, cur, 1);
        if (cur >= 0 && (lane == 0 || prev != cur)) commit_lm(gV, gGl, cur, vtt, vtp, vpp, glt, glp);
    }
    cost = warp_sum(cost);
    if (lane == 0) sWarp[warp] = cost;
    __syncthreads();
    if (tid == 0) {
        double sum = 0;
        for (int w = 0; w < kWarps; ++w) sum += sWarp[w];
        atomicAdd(gCost, sum);
    }
}

template <int MINB>
__global__ void __launch_bounds__(kFusedThreads, MINB)
k_ba_cam_pass_ring(int64_t lo, int64_t hi, int64_t chunk, const int32_t* __restrict__ c_cam, const int32_t* __restrict__ c_lm,
                   const double* __restrict__ c_ox, const double* __restrict__ c_oy, const CamTrig* __restrict__ cam_trig,
                   const LmTrig* __restrict__ lm_trig, double u, double v, double* __restrict__ gU, double* __restrict__ gGc) {
    extern __shared__ __align__(16) unsigned char dyn[];      // cp.async.bulk needs 16-byte aligned destinations
    WarpSlot* slots = reinterpret_cast<WarpSlot*>(dyn);
    __shared__ uint64_t full[kWarps][kRing];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double k1 = PTZ_DEG2RAD;
    const int64_t begin = lo + (int64_t)blockIdx.x * chunk;
    int64_t end = begin + chunk;
    if (end > hi) end = hi;
    const int64_t n_groups = end > begin ? (end - begin + kGroup - 1) / kGroup : 0;
    const int n_mine = n_groups > warp ? (int)((n_groups - warp + kWarps - 1) / kWarps) : 0;
    WarpSlot* my = slots + warp * kRing;
    if (lane == 0) {
        for (int r = 0; r < kRing; ++r) mbar_init(&full[warp][r], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int j = 0; j < kRing && j < n_mine; ++j)
            ring_issue(&my[j], &full[warp][j], begin + ((int64_t)warp + (int64_t)kWarps * j) * kGroup, end, c_cam, c_lm, c_ox, c_oy);
    }
    __syncwarp();
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
    // one step = 32 consecutive observations of a group (4 steps per group).  The streams are read from the warp's slot when
    // they are needed (shared-memory latency); only the dependent landmark-trig gather of step t + 1 is started before step t
    // is evaluated.
    const int n_steps = n_mine * (kGroup / 32);
    LmTrig lt = {0, 1, 0, 1};
#pragma unroll 1
    for (int t = -1; t < n_steps; ++t) {
        LmTrig nlt = {0, 1, 0, 1};
        const int tn = t + 1;
        if (tn < n_steps) {
            const int jn = tn >> 2, subn = tn & 3, rn = jn % kRing;
            if (subn == 0) mbar_wait(&full[warp][rn], (uint32_t)((jn / kRing) & 1));
            const int en = subn * 32 + lane;
            if (begin + ((int64_t)warp + (int64_t)kWarps * jn) * kGroup + en < end) nlt = lm_trig[my[rn].lm[en]];
        }
        if (t >= 0) {
            const int j = t >> 2, sub = t & 3, r = j % kRing;
            const int e = sub * 32 + lane;
            const int64_t g0 = begin + ((int64_t)warp + (int64_t)kWarps * j) * kGroup;
            const bool act = g0 + e < end;
            const int cam = act ? my[r].cam[e] : -1;
            const double ox = my[r].ox[e], oy = my[r].oy[e];
            if (sub == 3) {                                          // last read of this slot: refill it
                __syncwarp();
                if (lane == 0 && j + kRing < n_mine)
                    ring_issue(&my[r], &full[warp][r], g0 + (int64_t)kWarps * kRing * kGroup, end, c_cam, c_lm, c_ox, c_oy);
            }
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
                } else {
                    double* U = gU + 6 * (size_t)cam;
                    double* G = gGc + 3 * (size_t)cam;
                    const double k2 = k1 * k1;
                    atomicAdd(U + 0, upp * k2); atomicAdd(U + 1, upt * k2); atomicAdd(U + 2, upf * k1);
                    atomicAdd(U + 3, utt * k2); atomicAdd(U + 4, utf * k1); atomicAdd(U + 5, uff);
                    atomicAdd(G + 0, gp * k1); atomicAdd(G + 1, gt * k1); atomicAdd(G + 2, gf);
                }
            }
        }
        lt = nlt;
    }
    flush();
}

// residual-only pass (trial points of the trust-region loop, and _compute_residual itself)
__global__ void __launch_bounds__(kFusedThreads)
k_ba_residual(int64_t lo, int64_t hi, const int32_t* __restrict__ s_cam, const int32_t* __restrict__ s_lm,
              const double* __restrict__ s_ox, const double* __restrict__ s_oy, const int32_t* __restrict__ orig,
              const CamTrig* __restrict__ cam_trig, const LmTrig* __restrict__ lm_trig, double u, double v,
              double* __restrict__ resid, double* __restrict__ gSumSq) {
    __shared__ double sWarp[kFusedThreads / 32];
    double cost = 0.0;
    for (int64_t k = lo + (int64_t)blockIdx.x * kFusedThreads + threadIdx.x; k < hi; k += (int64_t)gridDim.x * kFusedThreads) {
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
int ba_set_params(ptzba_ba* ba, const double* d_x, const double* d_ref_pose3, bool zero_acc) {
    ptzba_ctx* ctx = ba->ctx;
    const int nc = ba->touch_cam_hi - ba->touch_cam_lo, nl = ba->touch_lm_hi - ba->touch_lm_lo;
    const int n = nc > nl ? nc : nl;
    if (n > 0) {
        if (zero_acc)
            k_set_params<true><<<div_up(n, 256), 256, 0, ctx->stream>>>(ba->n_pose, ba->touch_cam_lo, ba->touch_cam_hi, ba->touch_lm_lo,
                                                                        ba->touch_lm_hi, d_x, d_ref_pose3, ba->cam_trig.p, ba->lm_trig.p,
                                                                        ba->acc.cost, ba->acc.U, ba->acc.gc, ba->acc.V, ba->acc.gl);
        else
            k_set_params<false><<<div_up(n, 256), 256, 0, ctx->stream>>>(ba->n_pose, ba->touch_cam_lo, ba->touch_cam_hi, ba->touch_lm_lo,
                                                                         ba->touch_lm_hi, d_x, d_ref_pose3, ba->cam_trig.p, ba->lm_trig.p,
                                                                         nullptr, nullptr, nullptr, nullptr, nullptr);
        KERNEL_POST(ctx);
    }
    ba->acc_zeroed = zero_acc && n > 0;
    return PTZBA_OK;
}

// -> ba->acc.  The arena must be zero: ba_set_params(..., true) just before, else it is cleared here.
int ba_fused_pass(ptzba_ba* ba, double* d_resid) {
    ptzba_ctx* ctx = ba->ctx;
    cudaStream_t s = ctx->stream;
    // (a whole-arena all-reduce also fills blocks this rank's observations never touch: clear everything then)
    if (!ba->acc_zeroed || ba->arena_foreign) CU_CHECK(ctx, cudaMemsetAsync(ba->acc.base, 0, ba->acc.count * sizeof(double), s));
    ba->acc_zeroed = false;
    ba->arena_foreign = false;
    if (ba->n_obs == 0) return PTZBA_OK;
    const int32_t* orig = ba->identity_perm ? nullptr : ba->orig.p;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->profiling) {
        CU_CHECK(ctx, cudaEventCreate(&ev0));
        CU_CHECK(ctx, cudaEventCreate(&ev1));
        ctx->prof_events.push_back(ev0);
        ctx->prof_events.push_back(ev1);
        CU_CHECK(ctx, cudaEventRecord(ev0, s));
    }
    CU_CHECK(ctx, cudaEventRecord(ctx->ev_fork, s));      // everything before (k_set_params, the arena clear) precedes both passes
    // contiguous chunk per CTA: a multiple of 128 observations (every warp iteration starts on a 32-byte boundary of
    // all four streams), sized so that all resident CTAs of the single wave get the same amount of work.  With a
    // partition (ptzba_ba_set_partition) only this rank's slices of the two sorted lists are visited.
    const size_t smA = ba->cam_smem ? (size_t)ba->n_pose * 6 * sizeof(double) : 0;
    const int64_t nA = ba->lmo_hi - (ba->lmo_lo & ~(int64_t)3);
    // opt-in experiment: per-warp copy-engine rings (see k_ba_lm_pass_ring); needs 16-byte aligned slice starts
    static const char* ring_env = getenv("PTZBA_FUSED_RING");          // "1": rings; "2": rings + up-front dual landmark gather
    static const bool ring = ring_env != nullptr;
    static const bool ring_dual = ring && ring_env[0] == '2';
    const size_t smRing = sizeof(WarpSlot) * kWarps * kRing;
    const bool ringA = ring && ba->cam_smem && ba->lmo_lo % 4 == 0 && smRing + smA <= 200 * 1024;
    const bool ringB = ring && ba->cmo_lo % 4 == 0;
    if (ring && ba->grid_lm_ring == 0) {
        int pa = 1, pb = 1;
        if (ringA) {
            CU_CHECK(ctx, cudaFuncSetAttribute(k_ba_lm_pass_ring<kLmMinB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smRing + smA)));
            CU_CHECK(ctx, cudaFuncSetAttribute(k_ba_lm_pass_ring<kLmMinB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smRing + smA)));
            CU_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pa, k_ba_lm_pass_ring<kLmMinB, false>, kFusedThreads, smRing + smA));
        }
        CU_CHECK(ctx, cudaFuncSetAttribute(k_ba_cam_pass_ring<kCamMinB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smRing));
        CU_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pb, k_ba_cam_pass_ring<kCamMinB>, kFusedThreads, smRing));
        ba->grid_lm_ring = ctx->sm_count * (pa < 1 ? 1 : pa);
        ba->grid_cam_ring = ctx->sm_count * (pb < 1 ? 1 : pb);
    }
    if (nA > 0 && ringA) {
        int64_t chunkA = (nA + ba->grid_lm_ring - 1) / ba->grid_lm_ring;
        chunkA = (chunkA + kGroup - 1) / kGroup * kGroup;
        const int gridA = (int)((nA + chunkA - 1) / chunkA);
        if (ring_dual)
            k_ba_lm_pass_ring<kLmMinB, true><<<gridA, kFusedThreads, smRing + smA, s>>>(ba->lmo_lo, ba->lmo_hi, chunkA, ba->s_cam.p, ba->s_lm.p,
                                                                                      ba->s_ox.p, ba->s_oy.p, orig, ba->cam_trig.p, ba->lm_trig.p,
                                                                                      ba->n_pose, ba->u, ba->v, d_resid, ba->acc.V, ba->acc.gl,
                                                                                      ba->acc.cost);
        else
            k_ba_lm_pass_ring<kLmMinB, false><<<gridA, kFusedThreads, smRing + smA, s>>>(ba->lmo_lo, ba->lmo_hi, chunkA, ba->s_cam.p, ba->s_lm.p,
                                                                                       ba->s_ox.p, ba->s_oy.p, orig, ba->cam_trig.p, ba->lm_trig.p,
                                                                                       ba->n_pose, ba->u, ba->v, d_resid, ba->acc.V, ba->acc.gl,
                                                                                       ba->acc.cost);
        KERNEL_POST(ctx);
    } else if (nA > 0) {
        int64_t chunkA = (nA + ba->grid_lm_pass - 1) / ba->grid_lm_pass;
        chunkA = (chunkA + 127) / 128 * 128;
        const int gridA = (int)((nA + chunkA - 1) / chunkA);
        static const bool idx_ahead = getenv("PTZBA_LM_IDX_AHEAD") != nullptr;
        if (ba->cam_smem && idx_ahead) {
            if (smA > 40 * 1024)
                CU_CHECK(ctx, cudaFuncSetAttribute(k_ba_lm_pass4<kLmMinB, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA));
            k_ba_lm_pass4<kLmMinB, true, true><<<gridA, kFusedThreads, smA, s>>>(ba->lmo_lo, ba->lmo_hi, chunkA, ba->s_cam.p, ba->s_lm.p,
                                                                                ba->s_ox.p, ba->s_oy.p, orig, ba->cam_trig.p, ba->lm_trig.p,
                                                                                ba->n_pose, ba->u, ba->v, d_resid, ba->acc.V, ba->acc.gl,
                                                                                ba->acc.cost);
        } else if (ba->cam_smem)
            k_ba_lm_pass4<kLmMinB, true><<<gridA, kFusedThreads, smA, s>>>(ba->lmo_lo, ba->lmo_hi, chunkA, ba->s_cam.p, ba->s_lm.p, ba->s_ox.p,
                                                                    ba->s_oy.p, orig, ba->cam_trig.p, ba->lm_trig.p, ba->n_pose, ba->u, ba->v,
                                                                    d_resid, ba->acc.V, ba->acc.gl, ba->acc.cost);
        else
            k_ba_lm_pass4<kLmMinB, false><<<gridA, kFusedThreads, 0, s>>>(ba->lmo_lo, ba->lmo_hi, chunkA, ba->s_cam.p, ba->s_lm.p, ba->s_ox.p,
                                                                   ba->s_oy.p, orig, ba->cam_trig.p, ba->lm_trig.p, ba->n_pose, ba->u, ba->v,
                                                                   d_resid, ba->acc.V, ba->acc.gl, ba->acc.cost);
        KERNEL_POST(ctx);
    }
    // the keyframe-major pass touches disjoint accumulators: it runs on the side stream, concurrently with the landmark pass
    // (each kernel is a single wave; the second one fills the SMs the first one's finishing CTAs leave idle)
    const int64_t nB = ba->cmo_hi - ba->cmo_lo;
    static const bool concurrent = getenv("PTZBA_SERIAL_PASSES") == nullptr;
    static const bool cam_lookahead = getenv("PTZBA_CAM_LOOKAHEAD") != nullptr;
    if (nB > 0) {
        int64_t chunkB = (nB + ba->grid_cam_pass - 1) / ba->grid_cam_pass;
        chunkB = (chunkB + kFusedThreads - 1) / kFusedThreads * kFusedThreads;
        const int gridB = (int)((nB + chunkB - 1) / chunkB);
        cudaStream_t sb = concurrent ? ctx->side_stream : s;
        if (concurrent) CU_CHECK(ctx, cudaStreamWaitEvent(sb, ctx->ev_fork, 0));
        if (ringB) {
            int64_t chunkR = (nB + ba->grid_cam_ring - 1) / ba->grid_cam_ring;
            chunkR = (chunkR + kGroup - 1) / kGroup * kGroup;
            const int gridR = (int)((nB + chunkR - 1) / chunkR);
            k_ba_cam_pass_ring<kCamMinB><<<gridR, kFusedThreads, smRing, sb>>>(ba->cmo_lo, ba->cmo_hi, chunkR, ba->c_cam.p, ba->c_lm.p, ba->c_ox.p,
                                                                             ba->c_oy.p, ba->cam_trig.p, ba->lm_trig.p, ba->u, ba->v,
                                                                             ba->acc.U, ba->acc.gc);
        } else if (cam_lookahead) {
            k_ba_cam_pass<kCamMinB, true><<<gridB, kFusedThreads, 0, sb>>>(ba->cmo_lo, ba->cmo_hi, chunkB, ba->c_cam.p, ba->c_lm.p, ba->c_ox.p,
                                                                          ba->c_oy.p, ba->cam_trig.p, ba->lm_trig.p, ba->u, ba->v, ba->acc.U,
                                                                          ba->acc.gc);
        } else {
            k_ba_cam_pass<kCamMinB><<<gridB, kFusedThreads, 0, sb>>>(ba->cmo_lo, ba->cmo_hi, chunkB, ba->c_cam.p, ba->c_lm.p, ba->c_ox.p, ba->c_oy.p,
                                                                    ba->cam_trig.p, ba->lm_trig.p, ba->u, ba->v, ba->acc.U, ba->acc.gc);
        }
        KERNEL_POST(ctx);
        if (concurrent) {
            CU_CHECK(ctx, cudaEventRecord(ctx->ev_join, sb));
            CU_CHECK(ctx, cudaStreamWaitEvent(s, ctx->ev_join, 0));
        }
    }
    if (ev1) CU_CHECK(ctx, cudaEventRecord(ev1, s));
    if (ba->part_world > 1) PROPAGATE(ptzba_comm_allreduce_f64(ctx, ba->acc.base, (int64_t)ba->acc.count));
    return PTZBA_OK;
}

int ba_residual_pass(ptzba_ba* ba, double* d_resid, double* d_sumsq) {
    ptzba_ctx* ctx = ba->ctx;
    cudaStream_t s = ctx->stream;
    if (d_sumsq) CU_CHECK(ctx, cudaMemsetAsync(d_sumsq, 0, sizeof(double), s));
    if (ba->n_obs == 0) return PTZBA_OK;
    const int32_t* orig = ba->identity_perm ? nullptr : ba->orig.p;
    if (ba->lmo_hi > ba->lmo_lo) {
        k_ba_residual<<<stream_grid(ctx, ba->lmo_hi - ba->lmo_lo, kFusedThreads, 8), kFusedThreads, 0, s>>>(
            ba->lmo_lo, ba->lmo_hi, ba->s_cam.p, ba->s_lm.p, ba->s_ox.p, ba->s_oy.p, orig, ba->cam_trig.p, ba->lm_trig.p, ba->u, ba->v,
            d_resid, d_sumsq);
        KERNEL_POST(ctx);
    }
    if (ba->part_world > 1 && d_sumsq) PROPAGATE(ptzba_comm_allreduce_f64(ctx, d_sumsq, 1));
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
    ba->lm_hi = n_landmark; ba->lmo_hi = n_obs; ba->cmo_hi = n_obs;
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
    CU_TRY(ba->cam_trig.alloc(n_pose)); CU_TRY(ba->lm_trig.alloc(n_landmark));
    ba->acc.count = 1 + (size_t)n_pose * 9 + (size_t)n_landmark * 5;
    CU_TRY(ba->accum_store.alloc(ba->acc.count));
    CU_TRY(cudaMemsetAsync(ba->accum_store.p, 0, ba->acc.count * sizeof(double), s));      // untouched blocks stay zero for ever
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
    if (n_obs > 0) {
        DevBuf<int32_t> iota, perm2;
        DevBuf<unsigned char> tmp;
        CU_TRY(ba->c_cam.alloc(n_obs + 4)); CU_TRY(ba->c_lm.alloc(n_obs + 4));
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
        k_gather_cm<<<stream_grid(ctx, n_obs, 256, 8), 256, 0, s>>>(n_obs, perm2.p, ba->s_lm.p, ba->s_ox.p, ba->s_oy.p, ba->c_lm.p,
                                                                    ba->c_ox.p, ba->c_oy.p);
        ctx->launches++;
        CU_TRY(cudaStreamSynchronize(s));
    }

    // keyframes / landmarks touched by this problem's observations (first and last entry of the two sorted lists)
    ba->touch_cam_lo = 0; ba->touch_cam_hi = n_pose; ba->touch_lm_lo = 0; ba->touch_lm_hi = n_landmark;
    if (n_obs > 0) {
        int32_t e[4];
        CU_TRY(cudaMemcpyAsync(e + 0, ba->c_cam.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(e + 1, ba->c_cam.p + (n_obs - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(e + 2, ba->s_lm.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(e + 3, ba->s_lm.p + (n_obs - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        ba->touch_cam_lo = e[0]; ba->touch_cam_hi = e[1] + 1; ba->touch_lm_lo = e[2]; ba->touch_lm_hi = e[3] + 1;
    }
    // launch geometry of the fused pass: one wave of resident CTAs; keyframe trig table in shared memory when it fits
    {
        const size_t smA = (size_t)n_pose * 6 * sizeof(double);
        ba->cam_smem = smA <= 200 * 1024;
        int pa = 1, pb = 1;
        if (ba->cam_smem) {
            if (smA > 40 * 1024)
                CU_TRY(cudaFuncSetAttribute(k_ba_lm_pass4<kLmMinB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA));
            CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pa, k_ba_lm_pass4<kLmMinB, true>, kFusedThreads, smA));
        } else {
            CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pa, k_ba_lm_pass4<kLmMinB, false>, kFusedThreads, 0));
        }
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pb, k_ba_cam_pass<kCamMinB>, kFusedThreads, 0));
        ba->grid_lm_pass = ctx->sm_count * (pa < 1 ? 1 : pa);
        ba->grid_cam_pass = ctx->sm_count * (pb < 1 ? 1 : pb);
    }
#undef CU_TRY
    *out = ba;
    return PTZBA_OK;
}

// Replicated data, partitioned work: every rank holds the whole observation list and visits only its slice of it in
// the per-observation kernels - landmarks [lm_lo, lm_hi) of the landmark-major list (all observations of a landmark stay
// on one rank, so landmark blocks, Schur pair products and back-substituted landmark steps are complete locally) and
// positions [cm_lo, cm_hi) of the keyframe-major list.  Partial sums are combined with ncclAllReduce on the context
// stream inside the passes and the solver; ptzba_comm_init must have been called with the same rank / world size.
extern "C" int ptzba_ba_set_partition(ptzba_ba* ba, int rank, int world_size, int lm_lo, int lm_hi, int64_t cm_lo, int64_t cm_hi) {
    if (!ba) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = ba->ctx;
    ARG_CHECK(ctx, world_size >= 1 && rank >= 0 && rank < world_size);
    ARG_CHECK(ctx, lm_lo >= 0 && lm_lo <= lm_hi && lm_hi <= ba->n_lm && cm_lo >= 0 && cm_lo <= cm_hi && cm_hi <= ba->n_obs);
    if (world_size > 1 && (!ctx->nccl_comm || ctx->world != world_size || ctx->rank != rank))
        return ptzba_fail(ctx, PTZBA_ERR_STATE, "ptzba_comm_init(rank=%d, world=%d) must precede ptzba_ba_set_partition", rank, world_size);
    int32_t ends[2] = {0, 0};
    CU_CHECK(ctx, cudaMemcpyAsync(&ends[0], ba->lm_ptr.p + lm_lo, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaMemcpyAsync(&ends[1], ba->lm_ptr.p + lm_hi, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    ba->part_rank = rank; ba->part_world = world_size;
    ba->lm_lo = lm_lo; ba->lm_hi = lm_hi;
    ba->lmo_lo = ends[0]; ba->lmo_hi = ends[1];
    ba->cmo_lo = cm_lo; ba->cmo_hi = cm_hi;
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
    PROPAGATE(ba_set_params(ba, d_x, ba->ref_stage.p, true));
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
    if (cost && ba->cost_partial) {          // exchange mode without shared landmarks: the cost was left as a per-rank partial sum
        PROPAGATE(ptzba_comm_allreduce_f64(ctx, ba->acc.cost, 1));
        ba->cost_partial = false;
    }
    if (cost) CU_CHECK(ctx, cudaMemcpyAsync(&sumsq, ba->acc.cost, sizeof(double), cudaMemcpyDeviceToHost, s));
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    if (cost) *cost = 0.5 * sumsq;
    return PTZBA_OK;
}
