// pose_refine.cu - batched 3-parameter pose estimation on fixed ray <-> pixel matches.
//
// Reference behaviour:
//   slam_system/relocalization.py:22-40, :186-187   least_squares(_compute_residual, pose, x_scale='jac', ftol=1e-4, method='trf')
//                                                    on the matches of ONE pose (the rays are fixed landmarks)
//   slam_system/rf_map/util/ptz_pose_estimation.cpp:95-239 (preemptiveRANSACOneToMany): hypotheses from two-point minimal
//       samples; per round a random sample of B matches is projected under every hypothesis, the outliers (pixel distance >
//       threshold) are counted, the better half survives and every survivor is re-optimised on its inliers (optimizePTZ,
//       Levenberg-Marquardt) - until one hypothesis is left.  The reference walks the hypotheses one by one on the CPU.
// Here one CTA owns one hypothesis; a launch scores or refines ALL hypotheses of a round at once:
//   k_pose_score   outlier count (+ mean error) of every hypothesis on the sampled matches
//   k_pose_refine  Levenberg-Marquardt on (pan, tilt, f) over the hypothesis' inliers: residual, analytic 2x3 Jacobian
//                  (shared with the EKF: ptz_jac.cuh), J^T J / J^T r reduced over the CTA, 3x3 solve
//                  with Marquardt damping in one thread, one pass over the matches per iteration.
#include "common.h"
#include "ptz_jac.cuh"
#include "ptz_math.cuh"

namespace {

constexpr int kPT = 128;

__device__ __forceinline__ double block_sum(double v, double* sw) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kPT / 32; ++w) s += sw[w];
    return s;
}

__global__ void __launch_bounds__(kPT) k_pose_score(int n_sel, const int32_t* __restrict__ sel, const double* __restrict__ ptz_all, double u,
                                                   double v, const double* __restrict__ rays, const double* __restrict__ points,
                                                   double threshold, int32_t* __restrict__ out_outliers, double* __restrict__ out_mean_err) {
    __shared__ double sw[kPT / 32];
    __shared__ CamFull cam;
    const int h = blockIdx.x;
    if (threadIdx.x == 0) cam = make_cam(ptz_all[3 * h], ptz_all[3 * h + 1], ptz_all[3 * h + 2], u, v, nullptr);
    __syncthreads();
    double outl = 0.0, err = 0.0;
    for (int k = threadIdx.x; k < n_sel; k += kPT) {
        const int i = sel ? sel[k] : k;
        double x, y, q;
        project_full(cam, rays[2 * (size_t)i], rays[2 * (size_t)i + 1], x, y, q);
        const double dx = x - points[2 * (size_t)i], dy = y - points[2 * (size_t)i + 1];
        const double d = sqrt(dx * dx + dy * dy);
        err += d;
        if (!(d <= threshold)) outl += 1.0;        // ptz_pose_estimation.cpp:187: min_dis > threshold -> loss += 1 (NaN counts as an outlier)
    }
    outl = block_sum(outl, sw);
    err = block_sum(err, sw);
    if (threadIdx.x == 0) {
        out_outliers[h] = (int32_t)(outl + 0.5);
        if (out_mean_err) out_mean_err[h] = n_sel > 0 ? err / n_sel : 0.0;
    }
}

struct Normal10 { double c, a00, a01, a02, a11, a12, a22, g0, g1, g2; };

// cost = 0.5 sum r^2, A = J^T J (upper), g = J^T r over the used matches at pose p
__device__ __forceinline__ Normal10 pose_normal(const double* p, double u, double v, int n_sel, const int32_t* __restrict__ sel,
                                                const double* __restrict__ rays, const double* __restrict__ points,
                                                const unsigned* __restrict__ used, double* sw, CamFull* s_cam) {
    __syncthreads();
    if (threadIdx.x == 0) *s_cam = make_cam(p[0], p[1], p[2], u, v, nullptr);
    __syncthreads();
    const CamFull cam = *s_cam;
    Normal10 t = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = threadIdx.x; k < n_sel; k += kPT) {
        if (!((used[k >> 5] >> (k & 31)) & 1u)) continue;
        const int i = sel ? sel[k] : k;
        const double th = rays[2 * (size_t)i], ph = rays[2 * (size_t)i + 1];
        double x, y, q, jc[6], jr[4];
        project_full(cam, th, ph, x, y, q);
        jac_analytic_full(cam, nullptr, th, ph, jc, jr);
        const double rx = x - points[2 * (size_t)i], ry = y - points[2 * (size_t)i + 1];
        t.c += 0.5 * (rx * rx + ry * ry);
        t.a00 += jc[0] * jc[0] + jc[3] * jc[3]; t.a01 += jc[0] * jc[1] + jc[3] * jc[4]; t.a02 += jc[0] * jc[2] + jc[3] * jc[5];
        t.a11 += jc[1] * jc[1] + jc[4] * jc[4]; t.a12 += jc[1] * jc[2] + jc[4] * jc[5]; t.a22 += jc[2] * jc[2] + jc[5] * jc[5];
        t.g0 += jc[0] * rx + jc[3] * ry; t.g1 += jc[1] * rx + jc[4] * ry; t.g2 += jc[2] * rx + jc[5] * ry;
    }
    t.c = block_sum(t.c, sw);
    t.a00 = block_sum(t.a00, sw); t.a01 = block_sum(t.a01, sw); t.a02 = block_sum(t.a02, sw);
    t.a11 = block_sum(t.a11, sw); t.a12 = block_sum(t.a12, sw); t.a22 = block_sum(t.a22, sw);
    t.g0 = block_sum(t.g0, sw); t.g1 = block_sum(t.g1, sw); t.g2 = block_sum(t.g2, sw);
    return t;
}

// solves (A + lam diag(A)) d = -g for the symmetric 3x3 A (Cramer); returns false when the damped matrix is singular
__device__ __forceinline__ bool solve3(const Normal10& t, double lam, double* d) {
    const double a = t.a00 * (1 + lam), b = t.a01, c = t.a02, e = t.a11 * (1 + lam), f = t.a12, i = t.a22 * (1 + lam);
    const double c00 = e * i - f * f, c01 = c * f - b * i, c02 = b * f - c * e;
    const double det = a * c00 + b * c01 + c * c02;
    if (!(fabs(det) > 0.0) || !isfinite(det)) return false;
    const double c11 = a * i - c * c, c12 = b * c - a * f, c22 = a * e - b * b;
    const double id = 1.0 / det;
    d[0] = -(c00 * t.g0 + c01 * t.g1 + c02 * t.g2) * id;
    d[1] = -(c01 * t.g0 + c11 * t.g1 + c12 * t.g2) * id;
    d[2] = -(c02 * t.g0 + c12 * t.g1 + c22 * t.g2) * id;
    return true;
}

__global__ void __launch_bounds__(kPT) k_pose_refine(int n_sel, const int32_t* __restrict__ sel, double* __restrict__ ptz_all, double u, double v,
                                                    const double* __restrict__ rays, const double* __restrict__ points, double threshold,
                                                    int min_used, int max_iter, double ftol, double* __restrict__ out_cost,
                                                    int32_t* __restrict__ out_n_used, int32_t* __restrict__ out_iters) {
    extern __shared__ unsigned used[];          // bit k: selected match k is an inlier of the incoming pose
    __shared__ double sw[kPT / 32];
    __shared__ CamFull s_cam;
    __shared__ double s_p[3], s_d[3];
    __shared__ int s_flag;
    const int h = blockIdx.x;
    const int n_words = (n_sel + 31) / 32;
    for (int w = threadIdx.x; w < n_words; w += kPT) used[w] = 0u;
    if (threadIdx.x == 0) {
        s_p[0] = ptz_all[3 * h]; s_p[1] = ptz_all[3 * h + 1]; s_p[2] = ptz_all[3 * h + 2];
        s_cam = make_cam(s_p[0], s_p[1], s_p[2], u, v, nullptr);
    }
    __syncthreads();
    // inliers of the incoming pose on the selected matches (threshold <= 0: every selected match is used)
    double cnt = 0.0;
    for (int k = threadIdx.x; k < n_sel; k += kPT) {
        bool in = true;
        if (threshold > 0.0) {
            const int i = sel ? sel[k] : k;
            double x, y, q;
            project_full(s_cam, rays[2 * (size_t)i], rays[2 * (size_t)i + 1], x, y, q);
            const double dx = x - points[2 * (size_t)i], dy = y - points[2 * (size_t)i + 1];
            in = sqrt(dx * dx + dy * dy) <= threshold;
        }
        if (in) { atomicOr(&used[k >> 5], 1u << (k & 31)); cnt += 1.0; }
    }
    cnt = block_sum(cnt, sw);
    const int n_used = (int)(cnt + 0.5);
    if (threadIdx.x == 0 && out_n_used) out_n_used[h] = n_used;
    int iters = 0;
    double p[3] = {s_p[0], s_p[1], s_p[2]};
    Normal10 cur = pose_normal(p, u, v, n_sel, sel, rays, points, used, sw, &s_cam);
    if (n_used > min_used) {                     // ptz_pose_estimation.cpp:204: refine only with more than 4 inliers
        double lam = 1e-3;
        while (iters < max_iter) {
            if (threadIdx.x == 0) s_flag = solve3(cur, lam, s_d) ? 1 : 0;
            __syncthreads();
            if (!s_flag) { lam *= 10.0; if (lam > 1e12) break; continue; }
            double trial[3] = {p[0] + s_d[0], p[1] + s_d[1], p[2] + s_d[2]};
            const Normal10 nxt = pose_normal(trial, u, v, n_sel, sel, rays, points, used, sw, &s_cam);
            ++iters;
            if (isfinite(nxt.c) && nxt.c < cur.c) {
                const double red = cur.c - nxt.c;
                const bool done = red <= ftol * cur.c;
                p[0] = trial[0]; p[1] = trial[1]; p[2] = trial[2];
                cur = nxt;
                lam = fmax(lam * 0.1, 1e-15);
                if (done) break;
            } else {
                lam *= 10.0;
                if (lam > 1e12) break;
            }
        }
    }
    if (threadIdx.x == 0) {
        ptz_all[3 * h] = p[0]; ptz_all[3 * h + 1] = p[1]; ptz_all[3 * h + 2] = p[2];
        if (out_cost) out_cost[h] = cur.c;
        if (out_iters) out_iters[h] = iters;
    }
}

}  // namespace

extern "C" int ptzba_pose_score(ptzba_ctx* ctx, int mem, int n_hyp, const double* ptz, double u, double v, int n, const double* rays,
                                const double* points, int n_sel, const int32_t* sel, double threshold, int32_t* out_outliers,
                                double* out_mean_err) {
    if (!ctx) return PTZBA_ERR_ARG;
    ARG_CHECK(ctx, n_hyp >= 0 && n >= 0 && n_sel >= 0 && out_outliers && (n_hyp == 0 || ptz) && (n == 0 || (rays && points)));
    if (n_hyp == 0) return PTZBA_OK;
    cudaStream_t s = ctx->stream;
    if (!sel) n_sel = n;
    InArray<double> a_ptz, a_rays, a_pts;
    InArray<int32_t> a_sel;
    OutArray<int32_t> o_out;
    OutArray<double> o_err;
    CU_CHECK(ctx, a_ptz.stage(mem, ptz, (size_t)n_hyp * 3, s));
    CU_CHECK(ctx, a_rays.stage(mem, rays, (size_t)n * 2, s));
    CU_CHECK(ctx, a_pts.stage(mem, points, (size_t)n * 2, s));
    CU_CHECK(ctx, a_sel.stage(mem, sel, (size_t)n_sel, s));
    CU_CHECK(ctx, o_out.stage(mem, out_outliers, (size_t)n_hyp));
    CU_CHECK(ctx, o_err.stage(mem, out_mean_err, (size_t)n_hyp));
    k_pose_score<<<n_hyp, kPT, 0, s>>>(n_sel, a_sel.d, a_ptz.d, u, v, a_rays.d, a_pts.d, threshold, o_out.d, o_err.d);
    KERNEL_POST(ctx);
    CU_CHECK(ctx, o_out.finish(s));
    CU_CHECK(ctx, o_err.finish(s));
    if (mem == PTZBA_HOST) CU_CHECK(ctx, cudaStreamSynchronize(s));
    return PTZBA_OK;
}

extern "C" int ptzba_pose_refine(ptzba_ctx* ctx, int mem, int n_hyp, double* ptz, double u, double v, int n, const double* rays,
                                 const double* points, int n_sel, const int32_t* sel, double threshold, int min_used, int max_iter,
                                 double ftol, double* out_cost, int32_t* out_n_used, int32_t* out_iters) {
    if (!ctx) return PTZBA_ERR_ARG;
    ARG_CHECK(ctx, n_hyp >= 0 && n >= 0 && n_sel >= 0 && max_iter >= 0 && (n_hyp == 0 || ptz) && (n == 0 || (rays && points)));
    if (n_hyp == 0) return PTZBA_OK;
    cudaStream_t s = ctx->stream;
    if (!sel) n_sel = n;
    const size_t smem = ((size_t)(n_sel + 31) / 32 + 1) * sizeof(unsigned);
    if (smem > 200 * 1024) return ptzba_fail(ctx, PTZBA_ERR_ARG, "%d selected matches exceed the inlier mask in shared memory", n_sel);
    DevBuf<double> d_ptz;
    double* dp = ptz;
    if (mem == PTZBA_HOST) {
        CU_CHECK(ctx, d_ptz.alloc((size_t)n_hyp * 3));
        CU_CHECK(ctx, cudaMemcpyAsync(d_ptz.p, ptz, (size_t)n_hyp * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
        dp = d_ptz.p;
    }
    InArray<double> a_rays, a_pts;
    InArray<int32_t> a_sel;
    OutArray<double> o_cost;
    OutArray<int32_t> o_used, o_it;
    CU_CHECK(ctx, a_rays.stage(mem, rays, (size_t)n * 2, s));
    CU_CHECK(ctx, a_pts.stage(mem, points, (size_t)n * 2, s));
    CU_CHECK(ctx, a_sel.stage(mem, sel, (size_t)n_sel, s));
    CU_CHECK(ctx, o_cost.stage(mem, out_cost, (size_t)n_hyp));
    CU_CHECK(ctx, o_used.stage(mem, out_n_used, (size_t)n_hyp));
    CU_CHECK(ctx, o_it.stage(mem, out_iters, (size_t)n_hyp));
    if (smem > 40 * 1024) CU_CHECK(ctx, cudaFuncSetAttribute(k_pose_refine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_pose_refine<<<n_hyp, kPT, smem, s>>>(n_sel, a_sel.d, dp, u, v, a_rays.d, a_pts.d, threshold, min_used, max_iter, ftol, o_cost.d, o_used.d,
                                         o_it.d);
    KERNEL_POST(ctx);
    if (mem == PTZBA_HOST) CU_CHECK(ctx, cudaMemcpyAsync(ptz, dp, (size_t)n_hyp * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CU_CHECK(ctx, o_cost.finish(s));
    CU_CHECK(ctx, o_used.finish(s));
    CU_CHECK(ctx, o_it.finish(s));
    if (mem == PTZBA_HOST) CU_CHECK(ctx, cudaStreamSynchronize(s));
    return PTZBA_OK;
}
