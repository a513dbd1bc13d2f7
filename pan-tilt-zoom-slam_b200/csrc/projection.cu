// projection.cu - batched ray projection / back-projection / measurement-Jacobian kernels and their C-ABI entries.
//
// Reference behaviour reproduced (slam_system/):
//   PTZCamera.project_ray ptz_camera.py:191-210, project_rays :212-234 (strict in-image filter + ordered compaction),
//   PTZCamera.back_project_to_ray(s) :287-325, TransFunction.from_ray_to_image / from_image_to_ray
//   transformation.py:99-175, PtzSlam.compute_h_jacobian ptz_slam.py:73-138.
//
// Layout: rays / points / pixels are [n,2] row-major doubles, i.e. one 16-byte double2 per element, so every thread
// issues one 128-bit load and one 128-bit store; camera parameters are evaluated once per CTA into shared memory.
// These kernels are HBM-bound: 16 B in + 16 B out per (camera, ray) pair.
#include "common.h"
#include "ptz_math.cuh"
#include "ptz_jac.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ double2 ld2(const double* p, int64_t i) {
    return __ldg(reinterpret_cast<const double2*>(p) + i);
}
__device__ __forceinline__ void st2(double* p, int64_t i, double a, double b) {
    reinterpret_cast<double2*>(p)[i] = make_double2(a, b);
}

// grid: (ray tiles, camera groups).  out[(cam * n_ray + ray) * 2]
// A thread keeps ITS ray's direction d = (tan th, -tan ph sqrt(tan^2 th + 1), 1) in registers - the two tangents and the square
// root are the expensive part of project_full - and walks the cameras of its group, whose trig lives in shared memory: per
// (camera, ray) pair that leaves 13 FP64 operations, one division and one 16-byte store, i.e. the kernel is bound by the
// pixel stream it writes.  (Round 1 evaluated tan / tan / sqrt per pair: FP64-bound at 4096 cameras x 2000 rays.)
// The arithmetic per pair is project_full's, operation by operation, so results are bit-identical to the per-pair form.
constexpr int kCamGroup = 32;
__global__ void __launch_bounds__(kThreads) k_project_grid(int n_cam, int n_ray, const double* __restrict__ ptz, double u,
                                                           double v, const double* __restrict__ disp,
                                                           const double* __restrict__ rays,
                                                           double* __restrict__ out) {
    __shared__ CamFull cams[kCamGroup];
    const int c0 = blockIdx.y * kCamGroup;
    const int nc = min(kCamGroup, n_cam - c0);
    if (threadIdx.x < nc) {
        const int c = c0 + threadIdx.x;
        cams[threadIdx.x] = make_cam(ptz[3 * c], ptz[3 * c + 1], ptz[3 * c + 2], u, v, disp);
    }
    __syncthreads();
    for (int r = blockIdx.x * kThreads + threadIdx.x; r < n_ray; r += gridDim.x * kThreads) {
        const double2 ray = ld2(rays, r);
        const double tx = tan(ray.x * PTZ_DEG2RAD);
        const double tp = tan(ray.y * PTZ_DEG2RAD);
        const double r0 = tx;
        const double r1 = -tp * sqrt(tx * tx + 1.0);
#pragma unroll 4
        for (int k = 0; k < nc; ++k) {
            const CamFull& c = cams[k];
            const double a0 = c.cp * r0 - c.sp;
            const double a2 = c.sp * r0 + c.cp;
            const double q0 = a0 + c.d0;
            const double q1 = c.ct * r1 + c.st * a2 + c.d1;
            const double q2 = -c.st * r1 + c.ct * a2 + c.d2;
            const double iz = 1.0 / q2;
            st2(out, (int64_t)(c0 + k) * n_ray + r, c.f * q0 * iz + c.u, c.f * q1 * iz + c.v);
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_project_pairs(int64_t n_pair, const double* __restrict__ ptz, double u,
                                                            double v, const double* __restrict__ rays,
                                                            const int32_t* __restrict__ cam_idx,
                                                            const int32_t* __restrict__ ray_idx,
                                                            double* __restrict__ out) {
    for (int64_t k = (int64_t)blockIdx.x * kThreads + threadIdx.x; k < n_pair; k += (int64_t)gridDim.x * kThreads) {
        const int c = cam_idx[k];
        const CamFull cc = make_cam(ptz[3 * c], ptz[3 * c + 1], ptz[3 * c + 2], u, v, nullptr);
        const double2 ray = ld2(rays, ray_idx[k]);
        double x, y, q2;
        project_full(cc, ray.x, ray.y, x, y, q2);
        st2(out, k, x, y);
    }
}

// ---- project_rays with the strict in-image filter: flags + per-block counts, scan, ordered scatter ---------------
__global__ void __launch_bounds__(kThreads) k_filter_count(int n_ray, const double* __restrict__ ptz3, double u,
                                                           double v, const double* __restrict__ disp,
                                                           const double* __restrict__ rays, double height, double width,
                                                           double* __restrict__ xy_tmp, int32_t* __restrict__ block_count) {
    __shared__ CamFull cam;
    __shared__ int warp_cnt[kThreads / 32];
    if (threadIdx.x == 0) cam = make_cam(ptz3[0], ptz3[1], ptz3[2], u, v, disp);
    __syncthreads();
    const int r = blockIdx.x * kThreads + threadIdx.x;
    bool keep = false;
    if (r < n_ray) {
        const double2 ray = ld2(rays, r);
        double x, y, q2;
        project_full(cam, ray.x, ray.y, x, y, q2);
        st2(xy_tmp, r, x, y);
        keep = (0.0 < x) && (x < width) && (0.0 < y) && (y < height);   // ptz_camera.py:226
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < kThreads / 32; ++w) s += warp_cnt[w];
        block_count[blockIdx.x] = s;
    }
}

// single block exclusive scan of the block counts (n_blocks is small: n_ray / 256)
__global__ void k_scan_blocks(int n_blocks, int32_t* __restrict__ block_count, int32_t* __restrict__ total) {
    __shared__ int carry;
    __shared__ int tmp[1024];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int val = i < n_blocks ? block_count[i] : 0;
        tmp[threadIdx.x] = val;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            int t = threadIdx.x >= off ? tmp[threadIdx.x - off] : 0;
            __syncthreads();
            tmp[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n_blocks) block_count[i] = carry + tmp[threadIdx.x] - val;
        __syncthreads();
        if (threadIdx.x == 1023) carry += tmp[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(kThreads) k_filter_scatter(int n_ray, const double* __restrict__ xy_tmp, double height,
                                                             double width, const int32_t* __restrict__ block_offset,
                                                             double* __restrict__ out_xy, int32_t* __restrict__ out_index) {
    __shared__ int warp_off[kThreads / 32];
    const int r = blockIdx.x * kThreads + threadIdx.x;
    bool keep = false;
    double2 p = make_double2(0, 0);
    if (r < n_ray) {
        p = ld2(xy_tmp, r);
        keep = (0.0 < p.x) && (p.x < width) && (0.0 < p.y) && (p.y < height);
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_off[w] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int k = 0; k < kThreads / 32; ++k) {
            int c = warp_off[k];
            warp_off[k] = s;
            s += c;
        }
    }
    __syncthreads();
    if (keep) {
        const int pos = block_offset[blockIdx.x] + warp_off[w] + __popc(m & ((1u << lane) - 1u));
        st2(out_xy, pos, p.x, p.y);
        out_index[pos] = r;
    }
}

__global__ void __launch_bounds__(kThreads) k_backproject(int64_t n, const double* __restrict__ ptz, double u, double v,
                                                          const double* __restrict__ disp,
                                                          const double* __restrict__ points,
                                                          const int32_t* __restrict__ cam_idx,
                                                          double* __restrict__ out_rays) {
    __shared__ CamFull cam0;
    if (cam_idx == nullptr) {
        if (threadIdx.x == 0) cam0 = make_cam(ptz[0], ptz[1], ptz[2], u, v, disp);
        __syncthreads();
    }
    for (int64_t k = (int64_t)blockIdx.x * kThreads + threadIdx.x; k < n; k += (int64_t)gridDim.x * kThreads) {
        CamFull c;
        if (cam_idx) {
            const int ci = cam_idx[k];
            c = make_cam(ptz[3 * ci], ptz[3 * ci + 1], ptz[3 * ci + 2], u, v, disp);
        } else {
            c = cam0;
        }
        const double2 p = ld2(points, k);
        double th, ph;
        backproject_full(c, p.x, p.y, th, ph);
        st2(out_rays, k, th, ph);
    }
}

// ---- measurement Jacobian blocks (device functions in ptz_jac.cuh) -----------------------------------------------
// mode 1 reproduces ptz_slam.py:95-136 operation by operation: 10 projections per ray, central differences.
__global__ void __launch_bounds__(kThreads) k_h_blocks(int n_ray, const double* __restrict__ ptz3, double u, double v,
                                                       const double* __restrict__ disp, const double* __restrict__ rays,
                                                       int mode, double* __restrict__ jc_out, double* __restrict__ jr_out) {
    __shared__ CamFull cams[7];   // base, pan-/+, tilt-/+, f-/+
    if (threadIdx.x < 7) cams[threadIdx.x] = h_cam_variant(threadIdx.x, ptz3[0], ptz3[1], ptz3[2], u, v, disp);
    __syncthreads();
    for (int r = blockIdx.x * kThreads + threadIdx.x; r < n_ray; r += gridDim.x * kThreads) {
        const double2 ray = ld2(rays, r);
        double jc[6], jr[4];
        h_blocks_eval(cams, disp, ray.x, ray.y, mode, jc, jr);
#pragma unroll
        for (int i = 0; i < 6; ++i) jc_out[(int64_t)r * 6 + i] = jc[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) jr_out[(int64_t)r * 4 + i] = jr[i];
    }
}

// scatter blocks into the dense H the reference returns (H pre-zeroed): rows 2i,2i+1; cols 0..2 and 3+2i,4+2i
__global__ void k_h_dense_fill(int n_ray, const double* __restrict__ jc, const double* __restrict__ jr,
                               double* __restrict__ H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ray) return;
    const int64_t ld = 3 + 2 * (int64_t)n_ray;
    double* r0 = H + (2 * (int64_t)i) * ld;
    double* r1 = r0 + ld;
    r0[0] = jc[6 * i + 0]; r0[1] = jc[6 * i + 1]; r0[2] = jc[6 * i + 2];
    r1[0] = jc[6 * i + 3]; r1[1] = jc[6 * i + 4]; r1[2] = jc[6 * i + 5];
    r0[3 + 2 * i] = jr[4 * i + 0]; r0[4 + 2 * i] = jr[4 * i + 1];
    r1[3 + 2 * i] = jr[4 * i + 2]; r1[4 + 2 * i] = jr[4 * i + 3];
}

int grid_for(ptzba_ctx* ctx, int64_t n) {
    int64_t g = (n + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)ctx->sm_count * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

extern "C" int ptzba_project(ptzba_ctx* ctx, int mem, int n_cam, const double* ptz, double u, double v,
                             const double* disp, int n_ray, const double* rays, double* out_xy) {
    if (!ctx) return PTZBA_ERR_ARG;
    ARG_CHECK(ctx, n_cam >= 0 && n_ray >= 0 && (n_cam == 0 || ptz) && (n_ray == 0 || rays));
    ARG_CHECK(ctx, out_xy || n_cam == 0 || n_ray == 0);
    if (n_cam == 0 || n_ray == 0) return PTZBA_OK;
    cudaStream_t s = ctx->stream;
    InArray<double> d_ptz, d_rays, d_disp;
    OutArray<double> d_out;
    CU_CHECK(ctx, d_ptz.stage(mem, ptz, (size_t)n_cam * 3, s));
    CU_CHECK(ctx, d_rays.stage(mem, rays, (size_t)n_ray * 2, s));
    CU_CHECK(ctx, d_disp.stage(PTZBA_HOST, disp, 6, s));
    CU_CHECK(ctx, d_out.stage(mem, out_xy, (size_t)n_cam * n_ray * 2));
    const int n_groups = div_up(n_cam, kCamGroup);
    // enough ray tiles that the grid fills the machine a few times over even with one camera group
    int tiles = div_up(n_ray, kThreads);
    const int want = div_up((int64_t)ctx->sm_count * 8, n_groups);
    if (tiles > want) tiles = want < 1 ? 1 : want;
    dim3 grid(tiles, n_groups);
    if (n_groups > 65535) return ptzba_fail(ctx, PTZBA_ERR_ARG, "n_cam > %d: use ptzba_project_pairs", 65535 * kCamGroup);
    k_project_grid<<<grid, kThreads, 0, s>>>(n_cam, n_ray, d_ptz.d, u, v, d_disp.d, d_rays.d, d_out.d);
    KERNEL_POST(ctx);
    CU_CHECK(ctx, d_out.finish(s));
    if (mem == PTZBA_HOST) CU_CHECK(ctx, cudaStreamSynchronize(s));
    return PTZBA_OK;
}

extern "C" int ptzba_project_rays_filtered(ptzba_ctx* ctx, int mem, const double* ptz3, double u, double v,
                                           const double* disp, int n_ray, const double* rays, double height,
                                           double width, double* out_xy, int32_t* out_index, int32_t* out_count) {
    if (!ctx) return PTZBA_ERR_ARG;
    ARG_CHECK(ctx, ptz3 && n_ray >= 0 && out_count && (n_ray == 0 || (rays && out_xy && out_index)));
    *out_count = 0;
    if (n_ray == 0) return PTZBA_OK;
    cudaStream_t s = ctx->stream;
    InArray<double> d_ptz, d_rays, d_disp;
    OutArray<double> d_xy;
    OutArray<int32_t> d_idx;
    DevBuf<double> tmp;
    DevBuf<int32_t> blocks;
    const int nb = div_up(n_ray, kThreads);
    CU_CHECK(ctx, d_ptz.stage(PTZBA_HOST, ptz3, 3, s));
    CU_CHECK(ctx, d_rays.stage(mem, rays, (size_t)n_ray * 2, s));
    CU_CHECK(ctx, d_disp.stage(PTZBA_HOST, disp, 6, s));
    CU_CHECK(ctx, d_xy.stage(mem, out_xy, (size_t)n_ray * 2));
    CU_CHECK(ctx, d_idx.stage(mem, out_index, (size_t)n_ray));
    CU_CHECK(ctx, tmp.alloc((size_t)n_ray * 2));
    CU_CHECK(ctx, blocks.alloc((size_t)nb + 1));
    k_filter_count<<<nb, kThreads, 0, s>>>(n_ray, d_ptz.d, u, v, d_disp.d, d_rays.d, height, width, tmp.p, blocks.p);
    KERNEL_POST(ctx);
    k_scan_blocks<<<1, 1024, 0, s>>>(nb, blocks.p, blocks.p + nb);
    KERNEL_POST(ctx);
    k_filter_scatter<<<nb, kThreads, 0, s>>>(n_ray, tmp.p, height, width, blocks.p, d_xy.d, d_idx.d);
    KERNEL_POST(ctx);
    int32_t total = 0;
    CU_CHECK(ctx, cudaMemcpyAsync(&total, blocks.p + nb, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    *out_count = total;
    CU_CHECK(ctx, d_xy.finish(s, (size_t)total * 2));
    CU_CHECK(ctx, d_idx.finish(s, (size_t)total));
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    return PTZBA_OK;
}

extern "C" int ptzba_project_pairs(ptzba_ctx* ctx, int mem, int n_cam, const double* ptz, double u, double v,
                                   int n_ray, const double* rays, int64_t n_pair, const int32_t* cam_idx,
                                   const int32_t* ray_idx, double* out_xy) {
    if (!ctx) return PTZBA_ERR_ARG;
    ARG_CHECK(ctx, n_cam > 0 && n_ray > 0 && n_pair >= 0 && ptz && rays);
    if (n_pair == 0) return PTZBA_OK;
    ARG_CHECK(ctx, cam_idx && ray_idx && out_xy);
    cudaStream_t s = ctx->stream;
    InArray<double> d_ptz, d_rays;
    InArray<int32_t> d_c, d_r;
    OutArray<double> d_out;
    CU_CHECK(ctx, d_ptz.stage(mem, ptz, (size_t)n_cam * 3, s));
    CU_CHECK(ctx, d_rays.stage(mem, rays, (size_t)n_ray * 2, s));
    CU_CHECK(ctx, d_c.stage(mem, cam_idx, (size_t)n_pair, s));
    CU_CHECK(ctx, d_r.stage(mem, ray_idx, (size_t)n_pair, s));
    CU_CHECK(ctx, d_out.stage(mem, out_xy, (size_t)n_pair * 2));
    k_project_pairs<<<grid_for(ctx, n_pair), kThreads, 0, s>>>(n_pair, d_ptz.d, u, v, d_rays.d, d_c.d, d_r.d, d_out.d);
    KERNEL_POST(ctx);
    CU_CHECK(ctx, d_out.finish(s));
    if (mem == PTZBA_HOST) CU_CHECK(ctx, cudaStreamSynchronize(s));
    return PTZBA_OK;
}

extern "C" int ptzba_backproject(ptzba_ctx* ctx, int mem, int n_cam, const double* ptz, double u, double v,
                                 const double* disp, int64_t n, const double* points, const int32_t* cam_idx,
                                 double* out_rays) {
    if (!ctx) return PTZBA_ERR_ARG;
    ARG_CHECK(ctx, n_cam > 0 && ptz && n >= 0);
    if (n == 0) return PTZBA_OK;
    ARG_CHECK(ctx, points && out_rays);
    cudaStream_t s = ctx->stream;
    InArray<double> d_ptz, d_pts, d_disp;
    InArray<int32_t> d_c;
    OutArray<double> d_out;
    CU_CHECK(ctx, d_ptz.stage(mem, ptz, (size_t)n_cam * 3, s));
    CU_CHECK(ctx, d_pts.stage(mem, points, (size_t)n * 2, s));
    CU_CHECK(ctx, d_disp.stage(PTZBA_HOST, disp, 6, s));
    CU_CHECK(ctx, d_c.stage(mem, cam_idx, (size_t)n, s));
    CU_CHECK(ctx, d_out.stage(mem, out_rays, (size_t)n * 2));
    k_backproject<<<grid_for(ctx, n), kThreads, 0, s>>>(n, d_ptz.d, u, v, d_disp.d, d_pts.d, d_c.d, d_out.d);
    KERNEL_POST(ctx);
    CU_CHECK(ctx, d_out.finish(s));
    if (mem == PTZBA_HOST) CU_CHECK(ctx, cudaStreamSynchronize(s));
    return PTZBA_OK;
}

// device-side helper shared with ekf.cu: all pointers are device pointers
int ptzba_h_blocks_device(ptzba_ctx* ctx, const double* d_ptz3, double u, double v, const double* d_disp, int n_ray,
                          const double* d_rays, int mode, double* d_jc, double* d_jr) {
    if (n_ray == 0) return PTZBA_OK;
    k_h_blocks<<<grid_for(ctx, n_ray), kThreads, 0, ctx->stream>>>(n_ray, d_ptz3, u, v, d_disp, d_rays, mode, d_jc, d_jr);
    KERNEL_POST(ctx);
    return PTZBA_OK;
}

extern "C" int ptzba_h_jacobian_blocks(ptzba_ctx* ctx, int mem, const double* ptz3, double u, double v,
                                       const double* disp, int n_ray, const double* rays, int mode, double* jc,
                                       double* jr) {
    if (!ctx) return PTZBA_ERR_ARG;
    ARG_CHECK(ctx, ptz3 && n_ray >= 0 && (mode == PTZBA_JAC_ANALYTIC || mode == PTZBA_JAC_CENTRAL_FD));
    if (n_ray == 0) return PTZBA_OK;
    ARG_CHECK(ctx, rays && jc && jr);
    cudaStream_t s = ctx->stream;
    InArray<double> d_ptz, d_rays, d_disp;
    OutArray<double> d_jc, d_jr;
    CU_CHECK(ctx, d_ptz.stage(PTZBA_HOST, ptz3, 3, s));
    CU_CHECK(ctx, d_rays.stage(mem, rays, (size_t)n_ray * 2, s));
    CU_CHECK(ctx, d_disp.stage(PTZBA_HOST, disp, 6, s));
    CU_CHECK(ctx, d_jc.stage(mem, jc, (size_t)n_ray * 6));
    CU_CHECK(ctx, d_jr.stage(mem, jr, (size_t)n_ray * 4));
    PROPAGATE(ptzba_h_blocks_device(ctx, d_ptz.d, u, v, d_disp.d, n_ray, d_rays.d, mode, d_jc.d, d_jr.d));
    CU_CHECK(ctx, d_jc.finish(s));
    CU_CHECK(ctx, d_jr.finish(s));
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    return PTZBA_OK;
}

extern "C" int ptzba_h_jacobian_dense(ptzba_ctx* ctx, int mem, const double* ptz3, double u, double v,
                                      const double* disp, int n_ray, const double* rays, int mode, double* H) {
    if (!ctx) return PTZBA_ERR_ARG;
    ARG_CHECK(ctx, ptz3 && n_ray >= 0 && (mode == PTZBA_JAC_ANALYTIC || mode == PTZBA_JAC_CENTRAL_FD));
    if (n_ray == 0) return PTZBA_OK;
    ARG_CHECK(ctx, rays && H);
    cudaStream_t s = ctx->stream;
    InArray<double> d_ptz, d_rays, d_disp;
    OutArray<double> d_H;
    DevBuf<double> jc, jr;
    const size_t hn = (size_t)(2 * n_ray) * (3 + 2 * (size_t)n_ray);
    CU_CHECK(ctx, d_ptz.stage(PTZBA_HOST, ptz3, 3, s));
    CU_CHECK(ctx, d_rays.stage(mem, rays, (size_t)n_ray * 2, s));
    CU_CHECK(ctx, d_disp.stage(PTZBA_HOST, disp, 6, s));
    CU_CHECK(ctx, d_H.stage(mem, H, hn));
    CU_CHECK(ctx, jc.alloc((size_t)n_ray * 6));
    CU_CHECK(ctx, jr.alloc((size_t)n_ray * 4));
    PROPAGATE(ptzba_h_blocks_device(ctx, d_ptz.d, u, v, d_disp.d, n_ray, d_rays.d, mode, jc.p, jr.p));
    CU_CHECK(ctx, cudaMemsetAsync(d_H.d, 0, hn * sizeof(double), s));
    k_h_dense_fill<<<div_up(n_ray, 128), 128, 0, s>>>(n_ray, jc.p, jr.p, d_H.d);
    KERNEL_POST(ctx);
    CU_CHECK(ctx, d_H.finish(s));
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    return PTZBA_OK;
}
