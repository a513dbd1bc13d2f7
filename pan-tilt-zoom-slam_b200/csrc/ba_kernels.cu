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
#define PTZBA_LM_MINB 2
#endif
constexpr int kLmMinB = PTZBA_LM_MINB;   // resident CTAs per SM the fused pass is compiled for (3 = 80 registers spills: measured slower)

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

// keyframe-major scatter into the padded layout: sorted position k (keyframe cam[k], landmark-major position perm[k]) goes to
// pad[cam] + (k - ptr[cam])
__global__ void k_scatter_cm(int64_t n, const int32_t* __restrict__ perm, const int32_t* __restrict__ cam, const int32_t* __restrict__ ptr,
                             const int32_t* __restrict__ pad, const int32_t* __restrict__ s_lm, const double* __restrict__ s_ox,
                             const double* __restrict__ s_oy, int32_t* __restrict__ c_lm, double* __restrict__ c_ox,
                             double* __restrict__ c_oy) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t p = perm[k], c = cam[k];
        const int64_t d = (int64_t)pad[c] + (k - ptr[c]);
        c_lm[d] = s_lm[p];
        c_ox[d] = s_ox[p];
        c_oy[d] = s_oy[p];
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
        if (max_degree) atomicMax(max_degree, (int)(lo2 - lo));
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
// The fused pass = ONE launch (k_ba_fused) whose CTAs take one of two roles; every sum is formed where its operands are
// adjacent, nothing is scattered.
//   lm_role   landmark-major: residual (written), cost, per-landmark V / g_l
//   cam_role  keyframe-major: per-keyframe U / g_c live in registers across a whole run of the keyframe
// Both roles give every thread FOUR consecutive observations (one 128-bit index load and one 256-bit load per
// stream), evaluate the four projections as independent FP64 chains without a data-dependent branch between them
// (ILP 4: the round-1 kernels were dependency-latency bound with one serial, branchy chain per thread), and keep a
// rolling half-iteration pipeline of landmark-trig gathers in registers: the rows an iteration uses first were
// requested during the previous iteration, the rest at its top, index vectors two iterations ahead.
// Why two roles and not one pass: FP64 has no native shared-memory atomic add and L2 REDs cost ~5 ns per entry
// chip-wide, so whichever side is scattered costs more than re-streaming 20 B/observation and re-evaluating the
// geometry.  The alternatives that were built, measured and removed (one pass with shared-memory CAS atomics,
// keyframe-major with L2 REDs, CTA-wide and per-warp TMA/mbarrier rings, two launches on two streams, cp.async-staged
// gathers, L1 prefetch, one pass with a static per-tile keyframe sort) are tabulated in DESIGN.md section 5.
// ---------------------------------------------------------------------------------------------------------------
struct __align__(32) D4 { double a, b, c, d; };

constexpr int kQuad = 4;
constexpr int kCamIter = 32 * kQuad;       // observations one warp consumes per iteration of the keyframe-major pass

// Keyframe-major pass.  The keyframe-major copy is stored PADDED: every keyframe's run starts at a multiple of 128
// entries (padding entries carry landmark id -1), so a warp iteration (32 lanes x 4 observations) never mixes two
// keyframes, needs no per-observation keyframe id (iter_cam[it], one int per 128 observations, names it) and loads
// with aligned 128-/256-bit accesses: 20 B per observation are streamed.  A CTA walks a contiguous range of
// iterations, its warps interleaved; a warp keeps the nine sums of its current keyframe in registers and commits them
// (warp reduction + 9 REDs) when the keyframe changes - about once per CTA.
// (Measured and removed: staging the gathered trig rows of the next iteration in shared memory with cp.async - 27.5 us instead of
// 19.5 us for this role alone, MIO-throttled by the per-lane 16-byte copies; requesting them into L1 one iteration ahead with
// prefetch.global.L1 - 48.6 instead of 42.2 us for the whole pass - or with dummy loads - 42.1 us, no gain.)
__device__ __forceinline__ void
cam_role(int cta, int it_lo, int it_hi, int iters_per_cta, const int32_t* __restrict__ iter_cam, const int32_t* __restrict__ c_lm,
         const double* __restrict__ c_ox, const double* __restrict__ c_oy, const CamTrig* __restrict__ cam_trig,
         const LmTrig* __restrict__ lm_trig, double u, double v, double* __restrict__ gU, double* __restrict__ gGc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarpsPerCta = kFusedThreads / 32;
    const int cta_begin = it_lo + cta * iters_per_cta;
    int cta_end = cta_begin + iters_per_cta;
    if (cta_end > it_hi) cta_end = it_hi;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0, a8 = 0;   // radian units, signs and scales applied on commit
    int wcam = -1;
    CamTrig wc = {0, 1, 0, 1, 1};
    auto flush = [&]() {
        if (wcam > 0) {
            a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3); a4 = warp_sum(a4);
            a5 = warp_sum(a5); a6 = warp_sum(a6); a7 = warp_sum(a7); a8 = warp_sum(a8);
            if (lane == 0) {
                double* U = gU + 6 * (size_t)wcam;
                double* G = gGc + 3 * (size_t)wcam;
                const double k1 = PTZ_DEG2RAD, k2 = k1 * k1;
                // d/d pan = -d/d alpha: the pan row / column changes sign
                atomicAdd(U + 0, a0 * k2); atomicAdd(U + 1, -a1 * k2); atomicAdd(U + 2, -a2 * k1);
                atomicAdd(U + 3, a3 * k2); atomicAdd(U + 4, a4 * k1); atomicAdd(U + 5, a5);
                atomicAdd(G + 0, -a6 * k1); atomicAdd(G + 1, a7 * k1); atomicAdd(G + 2, a8);
            }
        }
        a0 = a1 = a2 = a3 = a4 = a5 = a6 = a7 = a8 = 0.0;
    };
    // Rolling software pipeline at half-iteration granularity, registers only: while the first two observations of an iteration
    // are evaluated the trig rows of its last two are in flight, and while those are evaluated the rows of the NEXT iteration's
    // first two are - the gather (an L2 round trip, the top stall of the round-2 profile: 15 % of all samples on its first use)
    // always has two observations' worth of arithmetic to hide behind, at the register cost of the four rows held before.
    // Landmark ids are loaded two iterations ahead, observed pixels one.
    int it = cta_begin + warp;
    if (it >= cta_end) return;
    auto ld_ids = [&](int i) { return __ldg(reinterpret_cast<const int4*>(c_lm + (size_t)i * kCamIter + (size_t)lane * kQuad)); };
    auto row = [&](int l) { return lm_trig[l < 0 ? 0 : l]; };
    int4 idA = ld_ids(it), idB = make_int4(-1, -1, -1, -1);
    D4 oxA, oyA, oxB = {0, 0, 0, 0}, oyB = {0, 0, 0, 0};
    int camA = __ldg(iter_cam + it), camB = -1;
    {
        const size_t o = (size_t)it * kCamIter + (size_t)lane * kQuad;
        oxA = *reinterpret_cast<const D4*>(c_ox + o);
        oyA = *reinterpret_cast<const D4*>(c_oy + o);
    }
    if (it + kWarpsPerCta < cta_end) idB = ld_ids(it + kWarpsPerCta);
    LmTrig p0 = row(idA.x), p1 = row(idA.y);                  // rows of the first two observations of the current iteration
    auto accumulate = [&](const LmTrig& lt, double ox, double oy, bool valid) {
        double x, y;
        ObsGeom g;
        project_fast_jac(wc, lt, u, v, x, y, g);
        const double rx = x - ox, ry = y - oy;
        if (valid) {                                          // predicated FMAs, no branch
            a0 = fma(g.xa, g.xa, fma(g.ya, g.ya, a0));
            a1 = fma(g.xa, g.xt, fma(g.ya, g.yt, a1));
            a2 = fma(g.xa, g.px, fma(g.ya, g.py, a2));
            a3 = fma(g.xt, g.xt, fma(g.yt, g.yt, a3));
            a4 = fma(g.xt, g.px, fma(g.yt, g.py, a4));
            a5 = fma(g.px, g.px, fma(g.py, g.py, a5));
            a6 = fma(g.xa, rx, fma(g.ya, ry, a6));
            a7 = fma(g.xt, rx, fma(g.yt, ry, a7));
            a8 = fma(g.px, rx, fma(g.py, ry, a8));
        }
    };
#pragma unroll 1
    for (; it < cta_end; it += kWarpsPerCta) {
        const int itn = it + kWarpsPerCta;
        const LmTrig q0 = row(idA.z), q1 = row(idA.w);        // rows of this iteration's last two observations: in flight during the first two
        int4 idC = make_int4(-1, -1, -1, -1);
        if (itn < cta_end) {                                  // next iteration's pixels, the ids of the one after
            const size_t o = (size_t)itn * kCamIter + (size_t)lane * kQuad;
            oxB = *reinterpret_cast<const D4*>(c_ox + o);
            oyB = *reinterpret_cast<const D4*>(c_oy + o);
            camB = __ldg(iter_cam + itn);
            if (itn + kWarpsPerCta < cta_end) idC = ld_ids(itn + kWarpsPerCta);
        }
        if (camA != wcam) { flush(); wcam = camA; wc = cam_trig[camA]; }              // warp-uniform
        const bool on = camA > 0;                                                     // keyframe 0 is the fixed reference pose
        if (on) {
            accumulate(p0, oxA.a, oyA.a, idA.x >= 0);
            accumulate(p1, oxA.b, oyA.b, idA.y >= 0);
        }
        p0 = row(idB.x); p1 = row(idB.y);                     // rows of the next iteration's first two: in flight during the last two
        if (on) {
            accumulate(q0, oxA.c, oyA.c, idA.z >= 0);
            accumulate(q1, oxA.d, oyA.d, idA.w >= 0);
        }
        idA = idB; idB = idC; oxA = oxB; oyA = oyB; camA = camB;
    }
    flush();
}

// commit of one landmark's sums (radian units -> per degree)
__device__ __forceinline__ void commit_lm(double* __restrict__ gV, double* __restrict__ gGl, int lm, double vtt, double vtp,
                                          double vpp, double glt, double glp) {
    const double k1 = PTZ_DEG2RAD, k2 = k1 * k1;
    atomicAdd(gV + 3 * (size_t)lm + 0, vtt * k2);
    atomicAdd(gV + 3 * (size_t)lm + 1, vtp * k2);
    atomicAdd(gV + 3 * (size_t)lm + 2, vpp * k2);
    atomicAdd(gGl + 2 * (size_t)lm + 0, glt * k1);
    atomicAdd(gGl + 2 * (size_t)lm + 1, glp * k1);
}

// 4 consecutive landmark ids k0 .. k0+3 of the landmark-major list restricted to [lo, end): -1 = no observation in this slot
__device__ __forceinline__ int4 load_lm_ids(int64_t k0, int64_t lo, int64_t end, const int32_t* __restrict__ s_lm) {
    if (k0 >= lo && k0 + kQuad <= end) return __ldg(reinterpret_cast<const int4*>(s_lm + k0));
    int4 r;
    r.x = (k0 >= lo && k0 < end) ? s_lm[k0] : -1;
    r.y = (k0 + 1 >= lo && k0 + 1 < end) ? s_lm[k0 + 1] : -1;
    r.z = (k0 + 2 >= lo && k0 + 2 < end) ? s_lm[k0 + 2] : -1;
    r.w = (k0 + 3 >= lo && k0 + 3 < end) ? s_lm[k0 + 3] : -1;
    return r;
}

struct QuadObs {
    int cam[kQuad];
    double ox[kQuad], oy[kQuad];
};

__device__ __forceinline__ void load_quad_obs(QuadObs& q, int64_t k0, int64_t lo, int64_t end, const int32_t* __restrict__ s_cam,
                                              const double* __restrict__ s_ox, const double* __restrict__ s_oy) {
    if (k0 >= lo && k0 + kQuad <= end) {
        const int4 c4 = __ldg(reinterpret_cast<const int4*>(s_cam + k0));
        const D4 x4 = *reinterpret_cast<const D4*>(s_ox + k0);
        const D4 y4 = *reinterpret_cast<const D4*>(s_oy + k0);
        q.cam[0] = c4.x; q.cam[1] = c4.y; q.cam[2] = c4.z; q.cam[3] = c4.w;
        q.ox[0] = x4.a; q.ox[1] = x4.b; q.ox[2] = x4.c; q.ox[3] = x4.d;
        q.oy[0] = y4.a; q.oy[1] = y4.b; q.oy[2] = y4.c; q.oy[3] = y4.d;
    } else {
#pragma unroll
        for (int i = 0; i < kQuad; ++i) {
            const bool in = k0 + i >= lo && k0 + i < end;
            q.cam[i] = in ? s_cam[k0 + i] : 0;
            q.ox[i] = in ? s_ox[k0 + i] : 0.0;
            q.oy[i] = in ? s_oy[k0 + i] : 0.0;
        }
    }
}

// Landmark-major pass.  A quad of 4 consecutive observations spans the tail of one landmark's run and / or the head of
// the next (landmark ids are non-decreasing): the sums of the observations that share the LAST id of the quad join the
// warp-segmented reduction (key = that id); the sums of those that share the FIRST id, when it differs, are committed by
// the thread itself; an id strictly between the two (a landmark with at most two observations) is committed per
// observation.  All three cases are predicated adds on the same straight-line code.
// Software pipeline per thread, registers only: landmark ids two iterations ahead, keyframe ids / observed pixels one
// iteration ahead; the trig rows of the quad's first and last landmark (all a quad needs unless a third landmark sits
// inside it) roll half an iteration ahead of their use (see the loop).  (Measured and removed: both rows a whole iteration
// ahead in registers - 28.0 instead of 25.5 us, the 16 extra registers spill; - and staged in shared memory by cp.async -
// 49 us, MIO-throttled.)
template <bool CAM_SMEM>
__device__ __forceinline__ void
lm_role(int cta, int64_t lo, int64_t hi, int64_t chunk, const int32_t* __restrict__ s_cam, const int32_t* __restrict__ s_lm,
        const double* __restrict__ s_ox, const double* __restrict__ s_oy, const int32_t* __restrict__ orig,
        const CamTrig* __restrict__ cam_trig, const LmTrig* __restrict__ lm_trig, int n_pose, double u, double v,
        double* __restrict__ resid, double* __restrict__ gV, double* __restrict__ gGl, double* __restrict__ gCost,
        double* __restrict__ smem, double* __restrict__ sWarp) {
    // smem: keyframe trig, 48 B per keyframe: {sp,cp} {st,ct} {f,-}: two LDS.128 + one LDS.64
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // this CTA's observations: [max(begin, lo), end); quads stay aligned to multiples of 4 of the array index
    const int64_t begin = (lo & ~(int64_t)3) + (int64_t)cta * chunk;
    int64_t end = begin + chunk;
    if (end > hi) end = hi;
    constexpr int64_t kStep = (int64_t)kFusedThreads * kQuad;
    int64_t k0 = begin + (int64_t)tid * kQuad;
    int4 idA = load_lm_ids(k0, lo, end, s_lm), idB = load_lm_ids(k0 + kStep, lo, end, s_lm);
    QuadObs qA, qB;
    load_quad_obs(qA, k0, lo, end, s_cam, s_ox, s_oy);
    qB = qA;
    auto row = [&](int l) { return lm_trig[l < 0 ? 0 : l]; };
    LmTrig ltF = row(idA.x);                    // trig row of the quad's first landmark
    if (CAM_SMEM) {
        for (int i = tid; i < n_pose * 5; i += kFusedThreads) {
            const int c = i / 5, e = i - 5 * c;
            smem[(size_t)c * 6 + e] = reinterpret_cast<const double*>(cam_trig)[i];
        }
        __syncthreads();
    }
    double cost = 0.0;
#pragma unroll 1
    for (int64_t base = begin; base < end; base += kStep, k0 += kStep) {
        // rolling pipeline, registers only: the row of the quad's LAST landmark travels while observation 0 (which always
        // belongs to the first landmark) is evaluated; the first row of the NEXT quad travels during observations 2 and 3
        const LmTrig ltL = row(idA.w);
        if (base + kStep < end) load_quad_obs(qB, k0 + kStep, lo, end, s_cam, s_ox, s_oy);
        const int4 idC = load_lm_ids(k0 + 2 * kStep, lo, end, s_lm);
        const int lm[kQuad] = {idA.x, idA.y, idA.z, idA.w};
        const int lm_first = lm[0], lm_last = lm[kQuad - 1];
        const bool direct = resid && !orig && k0 >= lo && k0 + kQuad <= end;
        double t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;      // observations with the quad's last id
        double h0 = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0;      // observations with the quad's first id (when it differs)
        auto one = [&](int i, double& rx, double& ry) {
            const bool is_last = lm[i] == lm_last, is_first = lm[i] == lm_first;
            LmTrig lt = (i == 0) ? ltF : (i == kQuad - 1 ? ltL : (is_last ? ltL : ltF));
            if (i != 0 && i != kQuad - 1 && !is_last && !is_first) lt = row(lm[i]);   // a third landmark inside the quad: rare
            CamTrig c;
            if (CAM_SMEM) {
                const double2* t = reinterpret_cast<const double2*>(smem + (size_t)qA.cam[i] * 6);
                const double2 pa = t[0], ti = t[1];
                c.sp = pa.x; c.cp = pa.y; c.st = ti.x; c.ct = ti.y; c.f = smem[(size_t)qA.cam[i] * 6 + 4];
            } else {
                c = cam_trig[qA.cam[i]];
            }
            double x, y;
            ObsGeom g;
            project_fast_jac(c, lt, u, v, x, y, g);
            const bool valid = lm[i] >= 0;
            rx = valid ? x - qA.ox[i] : 0.0;
            ry = valid ? y - qA.oy[i] : 0.0;
            cost = fma(rx, rx, fma(ry, ry, cost));
            const double vtt = fma(g.xa, g.xa, g.ya * g.ya), vtp = fma(g.xa, g.xp, g.ya * g.yp), vpp = fma(g.xp, g.xp, g.yp * g.yp);
            const double glt = fma(g.xa, rx, g.ya * ry), glp = fma(g.xp, rx, g.yp * ry);
            if (is_last) { t0 += vtt; t1 += vtp; t2 += vpp; t3 += glt; t4 += glp; }
            else if (is_first) { h0 += vtt; h1 += vtp; h2 += vpp; h3 += glt; h4 += glp; }
            else if (valid) commit_lm(gV, gGl, lm[i], vtt, vtp, vpp, glt, glp);        // only i = 1, 2 can get here
            if (resid && !direct && valid) {
                const int64_t o = orig ? (int64_t)orig[k0 + i] : k0 + i;
                reinterpret_cast<double2*>(resid)[o] = make_double2(rx, ry);
            }
        };
        double ra, rb, rc, rd;
        one(0, ra, rb);
        one(1, rc, rd);
        if (direct) *reinterpret_cast<D4*>(resid + 2 * k0) = D4{ra, rb, rc, rd};
        const LmTrig nF = row(idB.x);               // first row of the next quad: in flight during observations 2 and 3
        one(2, ra, rb);
        one(3, rc, rd);
        if (direct) *reinterpret_cast<D4*>(resid + 2 * k0 + 4) = D4{ra, rb, rc, rd};
        if (lm_first != lm_last && lm_first >= 0) commit_lm(gV, gGl, lm_first, h0, h1, h2, h3, h4);   // run ended inside this thread
        // the thread's last run joins the warp-segmented reduction (keys non-decreasing across lanes, -1 = none)
        seg_reduce5(lm_last, lane, t0, t1, t2, t3, t4);
        const int prev = __shfl_up_sync(0xffffffffu, lm_last, 1);
        if (lm_last >= 0 && (lane == 0 || prev != lm_last)) commit_lm(gV, gGl, lm_last, t0, t1, t2, t3, t4);
        idA = idB; idB = idC; qA = qB; ltF = nF;
    }
    cost = warp_sum(cost);
    if (lane == 0) sWarp[warp] = cost;
    __syncthreads();
    if (tid == 0) {
        double s = 0;
        for (int w = 0; w < kFusedThreads / 32; ++w) s += sWarp[w];
        atomicAdd(gCost, s);
    }
}

struct FusedArgs {
    // landmark-major role
    int64_t lo, hi, chunk;
    const int32_t *s_cam, *s_lm, *orig;
    const double *s_ox, *s_oy;
    double *resid, *gV, *gGl, *gCost;
    // keyframe-major role
    int it_lo, it_hi, iters_per_cta;
    const int32_t *iter_cam, *c_lm;
    const double *c_ox, *c_oy;
    double *gU, *gGc;
    // shared
    const CamTrig* cam_trig;
    const LmTrig* lm_trig;
    int n_pose;
    double u, v;
    int n_lm_cta, n_cam_cta;      // CTAs per role
};

// ONE launch for both passes: CTA b takes the landmark-major role iff floor((b+1) G_lm / G) > floor(b G_lm / G) (the two roles
// are interleaved evenly in launch order, so that every SM hosts CTAs of both kinds from the first to the last cycle: the
// landmark-major role is bound by the LSU / shared-memory pipe, the keyframe-major role by L2 gathers and the FP64 pipe).  The
// two-launch form of round 1 (two streams, fork / join events) was measured against it and removed: 43.2 vs 42.1 us.
template <int MINB, bool CAM_SMEM>
__global__ void __launch_bounds__(kFusedThreads, MINB) k_ba_fused(const FusedArgs a) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double sWarp[kFusedThreads / 32];
    const int G = a.n_lm_cta + a.n_cam_cta, b = blockIdx.x;
    const int before = (int)(((int64_t)b * a.n_lm_cta) / G), after = (int)(((int64_t)(b + 1) * a.n_lm_cta) / G);
    if (after > before)
        lm_role<CAM_SMEM>(before, a.lo, a.hi, a.chunk, a.s_cam, a.s_lm, a.s_ox, a.s_oy, a.orig, a.cam_trig, a.lm_trig, a.n_pose, a.u, a.v,
                          a.resid, a.gV, a.gGl, a.gCost, smem, sWarp);
    else
        cam_role(b - before, a.it_lo, a.it_hi, a.iters_per_cta, a.iter_cam, a.c_lm, a.c_ox, a.c_oy, a.cam_trig, a.lm_trig, a.u, a.v, a.gU,
                 a.gGc);
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
    // work split.  Landmark-major role: contiguous chunks that are multiples of 1024 observations (every thread of every CTA
    // iteration owns an aligned quad); keyframe-major role: contiguous ranges of 128-entry iterations.  With a partition
    // (ptzba_ba_set_partition) only this rank's slices of the two sorted lists are visited.
    const size_t smA = ba->cam_smem ? (size_t)ba->n_pose * 6 * sizeof(double) : 0;
    const int64_t nA = ba->lmo_hi - (ba->lmo_lo & ~(int64_t)3);
    const int nB = ba->cit_hi - ba->cit_lo;                 // iterations of 128 padded keyframe-major entries
    FusedArgs a;
    a.lo = ba->lmo_lo; a.hi = ba->lmo_hi; a.s_cam = ba->s_cam.p; a.s_lm = ba->s_lm.p; a.orig = orig; a.s_ox = ba->s_ox.p; a.s_oy = ba->s_oy.p;
    a.resid = d_resid; a.gV = ba->acc.V; a.gGl = ba->acc.gl; a.gCost = ba->acc.cost;
    a.it_lo = ba->cit_lo; a.it_hi = ba->cit_hi; a.iter_cam = ba->iter_cam.p; a.c_lm = ba->c_lm.p; a.c_ox = ba->c_ox.p; a.c_oy = ba->c_oy.p;
    a.gU = ba->acc.U; a.gGc = ba->acc.gc;
    a.cam_trig = ba->cam_trig.p; a.lm_trig = ba->lm_trig.p; a.n_pose = ba->n_pose; a.u = ba->u; a.v = ba->v;
    const int G = ba->grid_fused;                           // one wave of resident CTAs
    const double wA = nA > 0 ? (double)nA * ba->fused_lm_share : 0.0, wB = (double)nB * kCamIter * (100.0 - ba->fused_lm_share);
    int g_lm = (nA > 0) ? (int)(G * wA / (wA + wB) + 0.5) : 0;
    if (nA > 0 && g_lm < 1) g_lm = 1;
    if (nB > 0 && g_lm > G - 1) g_lm = G - 1;
    if (nB <= 0) g_lm = nA > 0 ? G : 0;
    const int g_cam = nB > 0 ? G - g_lm : 0;
    a.n_lm_cta = g_lm; a.n_cam_cta = g_cam;
    a.chunk = 1024; a.iters_per_cta = 1;
    if (g_lm > 0) {
        a.chunk = ((nA + g_lm - 1) / g_lm + 1023) / 1024 * 1024;
        a.n_lm_cta = (int)((nA + a.chunk - 1) / a.chunk);
    }
    if (g_cam > 0) {
        a.iters_per_cta = (nB + g_cam - 1) / g_cam;
        a.n_cam_cta = (nB + a.iters_per_cta - 1) / a.iters_per_cta;
    }
    if (a.n_lm_cta + a.n_cam_cta > 0) {
        if (ba->cam_smem) k_ba_fused<kLmMinB, true><<<a.n_lm_cta + a.n_cam_cta, kFusedThreads, smA, s>>>(a);
        else k_ba_fused<kLmMinB, false><<<a.n_lm_cta + a.n_cam_cta, kFusedThreads, 0, s>>>(a);
        KERNEL_POST(ctx);
    }
    if (ev1) CU_CHECK(ctx, cudaEventRecord(ev1, s));
    if (ba->part_world > 1) PROPAGATE(ptzba_comm_allreduce_f64(ctx, ba->acc.base, (int64_t)ba->acc.count));
    return PTZBA_OK;
}

int ba_residual_pass(ptzba_ba* ba, double* d_resid, double* d_sumsq, bool reduce) {
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
    if (reduce && ba->part_world > 1 && d_sumsq) PROPAGATE(ptzba_comm_allreduce_f64(ctx, d_sumsq, 1));
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
    ba->lm_hi = n_landmark; ba->lmo_hi = n_obs;
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

    // keyframe-major copy: stable sort of the landmark-major list by keyframe id, then every keyframe's run is moved to a
    // start that is a multiple of 128 entries (padding entries: landmark id -1); iter_cam names the keyframe of every
    // group of 128 entries.  touch_* = keyframes / landmarks touched by this problem's observations.
    ba->touch_cam_lo = 0; ba->touch_cam_hi = n_pose; ba->touch_lm_lo = 0; ba->touch_lm_hi = n_landmark;
    ba->n_cam_iter = 0;
    if (n_obs > 0) {
        DevBuf<int32_t> iota, perm2, cam_sorted, d_ptr, d_pad;
        DevBuf<unsigned char> tmp;
        CU_TRY(iota.alloc(n_obs)); CU_TRY(perm2.alloc(n_obs)); CU_TRY(cam_sorted.alloc(n_obs));
        k_iota<<<stream_grid(ctx, n_obs, 256, 8), 256, 0, s>>>(n_obs, iota.p);
        ctx->launches++;
        size_t bytes = 0;
        int end_bit = 1;
        while ((1ll << end_bit) < (long long)n_pose && end_bit < 31) ++end_bit;
        CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, ba->s_cam.p, cam_sorted.p, iota.p, perm2.p, (int)n_obs, 0, end_bit, s));
        CU_TRY(tmp.alloc(bytes));
        CU_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, ba->s_cam.p, cam_sorted.p, iota.p, perm2.p, (int)n_obs, 0, end_bit, s));
        // run offsets of the keyframes in the sorted list (binary search per keyframe), padded offsets on the host
        CU_TRY(d_ptr.alloc((size_t)n_pose + 1)); CU_TRY(d_pad.alloc((size_t)n_pose + 1));
        k_lm_ptr<<<div_up(n_pose + 1, 256), 256, 0, s>>>(n_pose, n_obs, cam_sorted.p, d_ptr.p, nullptr);
        ctx->launches++;
        std::vector<int32_t> h_ptr((size_t)n_pose + 1), h_pad((size_t)n_pose + 1);
        CU_TRY(cudaMemcpyAsync(h_ptr.data(), d_ptr.p, ((size_t)n_pose + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));
        int64_t padded = 0;
        for (int c = 0; c < n_pose; ++c) {
            h_pad[c] = (int32_t)padded;
            padded += ((int64_t)(h_ptr[c + 1] - h_ptr[c]) + kCamIter - 1) / kCamIter * kCamIter;
            if (padded > (int64_t)2147483000) return fail(ptzba_fail(ctx, PTZBA_ERR_ARG, "padded keyframe-major list exceeds 2^31 entries"));
        }
        h_pad[n_pose] = (int32_t)padded;
        ba->n_cam_iter = (int)(padded / kCamIter);
        std::vector<int32_t> h_iter((size_t)ba->n_cam_iter);
        int first_cam = -1, last_cam = -1;
        for (int c = 0; c < n_pose; ++c) {
            if (h_ptr[c + 1] > h_ptr[c]) { if (first_cam < 0) first_cam = c; last_cam = c; }
            for (int it = h_pad[c] / kCamIter; it < h_pad[c + 1] / kCamIter; ++it) h_iter[it] = c;
        }
        CU_TRY(ba->c_lm.alloc(padded + 4)); CU_TRY(ba->c_ox.alloc(padded + 4)); CU_TRY(ba->c_oy.alloc(padded + 4));
        CU_TRY(ba->iter_cam.alloc((size_t)ba->n_cam_iter + 1));
        CU_TRY(cudaMemsetAsync(ba->c_lm.p, 0xff, (size_t)(padded + 4) * sizeof(int32_t), s));
        CU_TRY(cudaMemsetAsync(ba->c_ox.p, 0, (size_t)(padded + 4) * sizeof(double), s));
        CU_TRY(cudaMemsetAsync(ba->c_oy.p, 0, (size_t)(padded + 4) * sizeof(double), s));
        CU_TRY(cudaMemcpyAsync(d_pad.p, h_pad.data(), ((size_t)n_pose + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        CU_TRY(cudaMemcpyAsync(ba->iter_cam.p, h_iter.data(), (size_t)ba->n_cam_iter * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        k_scatter_cm<<<stream_grid(ctx, n_obs, 256, 8), 256, 0, s>>>(n_obs, perm2.p, cam_sorted.p, d_ptr.p, d_pad.p, ba->s_lm.p, ba->s_ox.p,
                                                                     ba->s_oy.p, ba->c_lm.p, ba->c_ox.p, ba->c_oy.p);
        ctx->launches++;
        int32_t e[2];
        CU_TRY(cudaMemcpyAsync(e + 0, ba->s_lm.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaMemcpyAsync(e + 1, ba->s_lm.p + (n_obs - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CU_TRY(cudaStreamSynchronize(s));     // the host vectors and temporaries go out of scope
        ba->touch_cam_lo = first_cam; ba->touch_cam_hi = last_cam + 1; ba->touch_lm_lo = e[0]; ba->touch_lm_hi = e[1] + 1;
    }
    ba->cit_lo = 0; ba->cit_hi = ba->n_cam_iter;
    // launch geometry of the fused pass: one wave of resident CTAs; keyframe trig table in shared memory when it fits
    {
        const size_t smTab = (size_t)n_pose * 6 * sizeof(double);
        ba->cam_smem = smTab <= 100 * 1024;            // two CTAs per SM
        const size_t smA = ba->cam_smem ? smTab : 0;
        int pa = 1;
        if (ba->cam_smem) {
            CU_TRY(cudaFuncSetAttribute(k_ba_fused<kLmMinB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA));
            CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pa, k_ba_fused<kLmMinB, true>, kFusedThreads, smA));
        } else {
            CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pa, k_ba_fused<kLmMinB, false>, kFusedThreads, 0));
        }
        ba->grid_fused = ctx->sm_count * (pa < 1 ? 1 : pa);
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
    ba->pl_ready = false;            // the pair list covers the rank's landmark slice: rebuilt at the next solve
    ba->lm_lo = lm_lo; ba->lm_hi = lm_hi;
    ba->lmo_lo = ends[0]; ba->lmo_hi = ends[1];
    // the keyframe-major slice, given in observation positions, becomes a range of 128-entry iterations of the padded list
    ba->cit_lo = ba->n_obs > 0 ? (int)((cm_lo * (int64_t)ba->n_cam_iter) / ba->n_obs) : 0;
    ba->cit_hi = ba->n_obs > 0 ? (int)((cm_hi * (int64_t)ba->n_cam_iter) / ba->n_obs) : 0;
    return PTZBA_OK;
}

extern "C" int ptzba_ba_set_option(ptzba_ba* ba, int option, int value) {
    if (!ba) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = ba->ctx;
    switch (option) {
        case PTZBA_OPT_SCHUR_MODE:
            ARG_CHECK(ctx, value == PTZBA_SCHUR_AUTO || value == PTZBA_SCHUR_PER_LANDMARK || value == PTZBA_SCHUR_PAIR_LIST);
            ba->schur_mode = value;
            return PTZBA_OK;
        case PTZBA_OPT_FUSED_LM_SHARE:
            ARG_CHECK(ctx, value >= 1 && value <= 99);
            ba->fused_lm_share = value;
            return PTZBA_OK;
        default:
            return ptzba_fail(ctx, PTZBA_ERR_ARG, "unknown option %d", option);
    }
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
    // with a work partition only this rank's landmark slice is written: the other entries read as zero
    if (ba->part_world > 1 && ba->n_obs) CU_CHECK(ctx, cudaMemsetAsync(d_r, 0, 2 * (size_t)ba->n_obs * sizeof(double), s));
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
    if (ba->part_world > 1 && d_r && ba->n_obs) CU_CHECK(ctx, cudaMemsetAsync(d_r, 0, 2 * (size_t)ba->n_obs * sizeof(double), s));
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

// ---- pipelined host-buffer pass --------------------------------------------------------------------------------------------------
// begin: x (host, ideally pinned) travels on the problem's own copy stream, the fused pass runs on the context stream once it has
// arrived, and the requested block ranges travel back on the copy stream once the pass is done; the call returns at once.
// wait: blocks until the outputs of the last begin have landed and stores the cost.  With several problems in rotation the copies
// of one problem overlap the kernels of the next (the e2e figure of bench.py); a second begin on the SAME problem first waits,
// on the device, for the previous download of its accumulators.
extern "C" int ptzba_ba_normal_equations_begin(ptzba_ba* ba, const double* x, const double* reference_pose3, int kf_lo, int kf_hi,
                                               int lm_lo, int lm_hi, double* U, double* gc, double* V, double* gl, double* cost) {
    if (!ba) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = ba->ctx;
    ARG_CHECK(ctx, x && reference_pose3 && kf_lo >= 0 && kf_lo <= kf_hi && kf_hi <= ba->n_pose && lm_lo >= 0 && lm_lo <= lm_hi && lm_hi <= ba->n_lm);
    if (ba->async_pending) return ptzba_fail(ctx, PTZBA_ERR_STATE, "ptzba_ba_wait must follow ptzba_ba_normal_equations_begin");
    if (!ba->cp_stream) {
        CU_CHECK(ctx, cudaStreamCreateWithFlags(&ba->cp_stream, cudaStreamNonBlocking));
        CU_CHECK(ctx, cudaEventCreateWithFlags(&ba->ev_in, cudaEventDisableTiming));
        CU_CHECK(ctx, cudaEventCreateWithFlags(&ba->ev_pass, cudaEventDisableTiming));
        CU_CHECK(ctx, cudaEventCreateWithFlags(&ba->ev_out, cudaEventDisableTiming));
        CU_CHECK(ctx, cudaMallocHost((void**)&ba->h_cost, 4 * sizeof(double)));
        CU_CHECK(ctx, cudaEventRecord(ba->ev_out, ba->cp_stream));
    }
    cudaStream_t s = ctx->stream, c = ba->cp_stream;
    const size_t nx = 3 * (size_t)(ba->n_pose - 1) + 2 * (size_t)ba->n_lm;
    CU_CHECK(ctx, ba->ref_stage.alloc(4));
    CU_CHECK(ctx, ba->x_stage.alloc(nx + 2));
    // (the staging buffers are free: the previous pass on this problem was waited for)
    ba->h_cost[1] = reference_pose3[0]; ba->h_cost[2] = reference_pose3[1]; ba->h_cost[3] = reference_pose3[2];
    CU_CHECK(ctx, cudaMemcpyAsync(ba->ref_stage.p, ba->h_cost + 1, 3 * sizeof(double), cudaMemcpyHostToDevice, c));
    if (nx) CU_CHECK(ctx, cudaMemcpyAsync(ba->x_stage.p, x, nx * sizeof(double), cudaMemcpyHostToDevice, c));
    CU_CHECK(ctx, cudaEventRecord(ba->ev_in, c));
    CU_CHECK(ctx, cudaStreamWaitEvent(s, ba->ev_in, 0));
    CU_CHECK(ctx, cudaStreamWaitEvent(s, ba->ev_out, 0));          // the accumulators are no longer being downloaded
    PROPAGATE(ba_set_params(ba, ba->x_stage.p, ba->ref_stage.p, true));
    PROPAGATE(ba_fused_pass(ba, nullptr));
    // keyframe-sharded mode with the compact exchange set up: the blocks of the shared landmarks are summed before they travel
    if (ba->exchange_ready && ctx->world > 1) PROPAGATE(ptzba_ba_allreduce(ba));
    CU_CHECK(ctx, cudaEventRecord(ba->ev_pass, s));
    CU_CHECK(ctx, cudaStreamWaitEvent(c, ba->ev_pass, 0));
    const int nk = kf_hi - kf_lo, nl = lm_hi - lm_lo;
    if (U && nk) CU_CHECK(ctx, cudaMemcpyAsync(U, ba->acc.U + 6 * (size_t)kf_lo, (size_t)nk * 6 * sizeof(double), cudaMemcpyDeviceToHost, c));
    if (gc && nk) CU_CHECK(ctx, cudaMemcpyAsync(gc, ba->acc.gc + 3 * (size_t)kf_lo, (size_t)nk * 3 * sizeof(double), cudaMemcpyDeviceToHost, c));
    if (V && nl) CU_CHECK(ctx, cudaMemcpyAsync(V, ba->acc.V + 3 * (size_t)lm_lo, (size_t)nl * 3 * sizeof(double), cudaMemcpyDeviceToHost, c));
    if (gl && nl) CU_CHECK(ctx, cudaMemcpyAsync(gl, ba->acc.gl + 2 * (size_t)lm_lo, (size_t)nl * 2 * sizeof(double), cudaMemcpyDeviceToHost, c));
    CU_CHECK(ctx, cudaMemcpyAsync(ba->h_cost, ba->acc.cost, sizeof(double), cudaMemcpyDeviceToHost, c));
    CU_CHECK(ctx, cudaEventRecord(ba->ev_out, c));
    ba->async_pending = true;
    ba->async_cost_dst = cost;
    return PTZBA_OK;
}

extern "C" int ptzba_ba_wait(ptzba_ba* ba) {
    if (!ba) return PTZBA_ERR_ARG;
    ptzba_ctx* ctx = ba->ctx;
    if (!ba->async_pending) return PTZBA_OK;
    CU_CHECK(ctx, cudaEventSynchronize(ba->ev_out));
    if (ba->async_cost_dst) *ba->async_cost_dst = 0.5 * ba->h_cost[0];
    ba->async_pending = false;
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
