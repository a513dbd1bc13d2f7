/*
 * ptzba.h - C-ABI of libptzba.so, the B200 (sm_100a) implementation of the Pan-tilt-zoom-SLAM hot path.
 *
 * Boundary convention: the reference's only FFI is ctypes -> librf_map_python
 * (slam_system/rf_map/python_package/rf_map_wrapper.py:14-62, rf_map.hpp:47-64): opaque handle from *_new,
 * caller-owned numeric buffers passed as raw double pointers, synchronous calls.  This header keeps that shape
 * (extern "C", plain pointers and sizes, no C++/torch types) with ONE intentional tightening: every call returns
 * an int status instead of void+printf (rf_map.cpp:108-117), because there is no CPU fallback to hide a failure.
 *
 * Units: angles in DEGREES, focal length / pixels in PIXELS, all reals are IEEE double, indices are int32.
 * Every array pointer of a call lives in the memory space named by its `mem` argument:
 *   PTZBA_HOST   - pageable or pinned host memory; the library stages H2D/D2H itself and synchronises
 *   PTZBA_DEVICE - device memory of the context's GPU; work is enqueued on the context stream (ptzba_set_stream)
 *                  and is NOT synchronised unless the call returns scalars to the host.
 * Scalars and small fixed-size structs (ptz[3], disp[6], options, reports) are always host memory.
 */
#ifndef PTZBA_H
#define PTZBA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTZBA_VERSION 100

/* status codes */
#define PTZBA_OK 0
#define PTZBA_ERR_ARG 1        /* bad argument (null pointer, negative size, index out of range) */
#define PTZBA_ERR_CUDA 2       /* CUDA runtime error; see ptzba_last_error */
#define PTZBA_ERR_NUMERIC 3    /* factorisation broke down (matrix not positive definite) */
#define PTZBA_ERR_COMM 4       /* NCCL error */
#define PTZBA_ERR_STATE 5      /* call not valid in the object's current state */

#define PTZBA_HOST 0
#define PTZBA_DEVICE 1

/* Jacobian evaluation mode */
#define PTZBA_JAC_ANALYTIC 0   /* closed-form derivatives */
#define PTZBA_JAC_CENTRAL_FD 1 /* the reference's central differences, delta = 0.001 deg / 0.1 px (ptz_slam.py:87-88) */

typedef struct ptzba_ctx ptzba_ctx;          /* one per process/GPU  (cf. RFMap_new, rf_map.hpp:47) */
typedef struct ptzba_ba ptzba_ba;            /* a bundle-adjustment problem resident on the GPU */
typedef struct ptzba_ekf_batch ptzba_ekf_batch; /* many independent EKF sequences resident on the GPU */

/* ---- context ------------------------------------------------------------------------------------------- */
int ptzba_version(void);
int ptzba_create(int device, ptzba_ctx** out);
void ptzba_destroy(ptzba_ctx* ctx);
const char* ptzba_last_error(ptzba_ctx* ctx);               /* valid until the next call on ctx */
int ptzba_set_stream(ptzba_ctx* ctx, void* cuda_stream);    /* cudaStream_t; NULL = the context's own stream */
int ptzba_synchronize(ptzba_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
int64_t ptzba_launch_count(ptzba_ctx* ctx);
/* per-launch CUDA-event timing of the dominant kernel (the fused BA pass): between begin and end every launch of
 * that kernel is bracketed by two events on the launching stream; end synchronises and returns count and total. */
int ptzba_profile_begin(ptzba_ctx* ctx);
int ptzba_profile_end(ptzba_ctx* ctx, int32_t* n_launches, double* total_ms);

/* dense SPD solve A x = b (host arrays, order n, only the lower triangle of the symmetric A is read) with the cooperative
 * Cholesky + block-inverse kernels that bundle adjustment uses for its reduced camera system (replaces the dense SVD scipy runs
 * inside least_squares, bundle_adjustment.py:200-202).  *info = 0, or 1 + the first row of the panel with a non-positive pivot. */
int ptzba_dense_solve_spd(ptzba_ctx* ctx, int n, const double* A, const double* b, double* x, int* info);

/* ---- A1/A2: projection  (PTZCamera.project_ray ptz_camera.py:191-210, project_rays :212-234,
 *                           TransFunction.from_ray_to_image transformation.py:99-135) ------------------------- */
/* every camera x every ray.  ptz[n_cam*3] (pan,tilt,f), disp[6] or NULL (lambda_1..6, ptz_camera.py:106-115),
 * rays[n_ray*2] (theta,phi), out_xy[n_cam*n_ray*2]. */
int ptzba_project(ptzba_ctx* ctx, int mem, int n_cam, const double* ptz, double u, double v, const double* disp,
                  int n_ray, const double* rays, double* out_xy);
/* one camera, strict in-image filter 0<x<width, 0<y<height and order-preserving compaction (project_rays with
 * height/width given): out_xy[n_ray*2] and out_index[n_ray] receive *out_count kept points. */
int ptzba_project_rays_filtered(ptzba_ctx* ctx, int mem, const double* ptz3, double u, double v, const double* disp,
                                int n_ray, const double* rays, double height, double width,
                                double* out_xy, int32_t* out_index, int32_t* out_count);
/* explicit (camera, ray) pair list: out_xy[n_pair*2]. */
int ptzba_project_pairs(ptzba_ctx* ctx, int mem, int n_cam, const double* ptz, double u, double v,
                        int n_ray, const double* rays, int64_t n_pair, const int32_t* cam_idx,
                        const int32_t* ray_idx, double* out_xy);

/* ---- A3: back-projection  (PTZCamera.back_project_to_ray(s) ptz_camera.py:287-325,
 *                             TransFunction.from_image_to_ray transformation.py:137-175) ---------------------- */
/* points[n*2] -> rays[n*2]; cam_idx == NULL: all points use ptz[0..2]; else point k uses camera cam_idx[k]. */
int ptzba_backproject(ptzba_ctx* ctx, int mem, int n_cam, const double* ptz, double u, double v, const double* disp,
                      int64_t n, const double* points, const int32_t* cam_idx, double* out_rays);

/* ---- A4: measurement Jacobian  (PtzSlam.compute_h_jacobian ptz_slam.py:73-138) ---------------------------- */
/* blocks: jc[n*6] row-major 2x3 (cols pan,tilt,f), jr[n*4] row-major 2x2 (cols theta,phi). */
int ptzba_h_jacobian_blocks(ptzba_ctx* ctx, int mem, const double* ptz3, double u, double v, const double* disp,
                            int n_ray, const double* rays, int mode, double* jc, double* jr);
/* dense H[2n x (3+2n)] row-major exactly as the reference returns it (zeros included). */
int ptzba_h_jacobian_dense(ptzba_ctx* ctx, int mem, const double* ptz3, double u, double v, const double* disp,
                           int n_ray, const double* rays, int mode, double* H);

/* ---- N2: pose estimation on fixed ray <-> pixel matches, batched over hypotheses (relocalization.py:22-40, :186-187;
 *            rf_map/util/ptz_pose_estimation.cpp:95-239 preemptive RANSAC: score the hypotheses on a sample of matches, keep the
 *            better half, re-optimise the survivors on their inliers).  rays[n*2] (theta, phi), points[n*2] pixels, sel[n_sel]
 *            indices of the sampled matches (NULL = all n), ptz[n_hyp*3].  One CTA per hypothesis. -------------------------------- */
/* outliers[h] = number of selected matches whose pixel distance under hypothesis h exceeds `threshold`; mean_err may be NULL */
int ptzba_pose_score(ptzba_ctx* ctx, int mem, int n_hyp, const double* ptz, double u, double v, int n, const double* rays,
                     const double* points, int n_sel, const int32_t* sel, double threshold, int32_t* out_outliers, double* out_mean_err);
/* Levenberg-Marquardt on (pan, tilt, f) of every hypothesis over the selected matches that are inliers of its incoming pose
 * (threshold <= 0: all selected matches), analytic Jacobian; a hypothesis with at most min_used inliers is left unchanged
 * (the reference uses 4).  Stops when an accepted step reduces the cost by less than ftol * cost, or after max_iter cost
 * evaluations.  ptz in/out; out_cost = 0.5 sum r^2 over the used matches, out_n_used, out_iters: [n_hyp], any may be NULL. */
int ptzba_pose_refine(ptzba_ctx* ctx, int mem, int n_hyp, double* ptz, double u, double v, int n, const double* rays,
                      const double* points, int n_sel, const int32_t* sel, double threshold, int min_used, int max_iter, double ftol,
                      double* out_cost, int32_t* out_n_used, int32_t* out_iters);

/* ---- A5: EKF update  (PtzSlam.ekf_update ptz_slam.py:210-289, predict lines :418-426) ---------------------- */
typedef struct ptzba_ekf_params {
    double u, v;            /* principal point */
    double disp[6];         /* displacement lambdas (zeros = none) */
    double observe_var;     /* 0.1   ptz_slam.py:68  */
    double angle_var;       /* 0.001 ptz_slam.py:70  */
    double f_var;           /* 1     ptz_slam.py:71  */
    double height, width;   /* image size for the in-image filter */
    int jac_mode;           /* PTZBA_JAC_* */
} ptzba_ekf_params;

/* single sequence, HOST buffers, in place like the reference:
 *   rays[n_total*2], state_cov[(3+2 n_total)^2] row-major dense, ptz3 (in: predicted pose, out: updated),
 *   velocity3 (out), observed_xy[m*2], observed_index[m] ascending global ray ids.
 *   *out_n_matched receives the size of observed ∩ in-image (ptz_slam.py:228-233).  */
int ptzba_ekf_update(ptzba_ctx* ctx, const ptzba_ekf_params* prm, int n_total, double* rays, double* state_cov,
                     double* ptz3, double* velocity3, int m, const double* observed_xy,
                     const int32_t* observed_index, int32_t* out_n_matched);

/* many independent sequences resident on the GPU (config 4); every sequence has n_ray rays */
int ptzba_ekf_batch_create(ptzba_ctx* ctx, const ptzba_ekf_params* prm, int n_seq, int n_ray, int max_obs,
                           const double* rays0 /*[n_seq*n_ray*2] host*/, const double* ptz0 /*[n_seq*3] host*/,
                           ptzba_ekf_batch** out);
void ptzba_ekf_batch_destroy(ptzba_ekf_batch* b);
/* one predict+update for every sequence.  obs_xy[n_seq*max_obs*2], obs_index[n_seq*max_obs], obs_count[n_seq]
 * live in `mem`.  out_matched[n_seq] (host, may be NULL). */
int ptzba_ekf_batch_step(ptzba_ekf_batch* b, int mem, const double* obs_xy, const int32_t* obs_index,
                         const int32_t* obs_count, int32_t* out_matched);
/* update only (no predict): the contract of PtzSlam.ekf_update, where the caller has already predicted the pose */
int ptzba_ekf_batch_update_only(ptzba_ekf_batch* b, int mem, const double* obs_xy, const int32_t* obs_index,
                                const int32_t* obs_count, int32_t* out_matched);
/* overwrite the state of one sequence from host buffers (any may be NULL): ptz3, velocity3, rays[n_ray*2],
 * state_cov[(3+2 n_ray)^2] row-major */
int ptzba_ekf_batch_set(ptzba_ekf_batch* b, int seq, const double* ptz3, const double* velocity3, const double* rays,
                        const double* state_cov);
/* copy state back to host: ptz[n_seq*3], velocity[n_seq*3], rays[n_seq*n_ray*2] (any may be NULL) */
int ptzba_ekf_batch_get(ptzba_ekf_batch* b, double* ptz, double* velocity, double* rays);
/* dense covariance of one sequence, (3+2 n_active)^2 row-major, host */
int ptzba_ekf_batch_get_cov(ptzba_ekf_batch* b, int seq, double* state_cov);
/* ---- N4: ray bookkeeping between frames on the RESIDENT state (PtzSlam.remove_rays / add_rays, ptz_slam.py:291-388).
 * n_ray of ptzba_ekf_batch_create is the initial ray count of every sequence AND the initial capacity (stride) of the per-sequence
 * arrays; afterwards every sequence has its own active count, the first n_active rays of its slot.  ptzba_ekf_batch_set /
 * _get_cov / _get_rays address the active part of one sequence (dense, no padding); ptzba_ekf_batch_get keeps returning rays
 * with the capacity as stride. */
int ptzba_ekf_batch_n_rays(ptzba_ekf_batch* b, int seq, int32_t* n_active, int32_t* capacity);
int ptzba_ekf_batch_get_rays(ptzba_ekf_batch* b, int seq, double* rays /*[n_active*2] host*/);
/* the rays delete_index[0..n_del) (host, any order) leave sequence `seq` with their covariance rows / columns; order of the others kept */
int ptzba_ekf_batch_remove_rays(ptzba_ekf_batch* b, int seq, int n_del, const int32_t* delete_index);
/* k rays (host [k*2]) are appended to sequence `seq`: variance angle_var, no correlation; the capacity grows when needed */
int ptzba_ekf_batch_add_rays(ptzba_ekf_batch* b, int seq, int k, const double* new_rays);
/* covariance part of the predict step alone: P[0:3,0:3] += 5 diag(angle_var, angle_var, f_var) for every sequence (ptz_slam.py:425-426) */
int ptzba_ekf_batch_predict_cov(ptzba_ekf_batch* b);
/* current observation capacity per step: obs_xy / obs_index of the step calls are [n_seq * max_obs * 2] / [n_seq * max_obs] */
int ptzba_ekf_batch_max_obs(ptzba_ekf_batch* b, int32_t* max_obs);
/* factorisation route of every sequence (host int32[n_seq]): 0 = Cholesky of the innovation covariance, 1 = pivoted LU (S was
 * indefinite once - the reference's covariance write-back, ptz_slam.py:281-289, makes P indefinite - and stays on that route) */
int ptzba_ekf_batch_route(ptzba_ekf_batch* b, int32_t* route);
/* grow the ray capacity per sequence and / or the observation capacity per step (max_obs) ahead of time */
int ptzba_ekf_batch_reserve(ptzba_ekf_batch* b, int ray_capacity, int max_obs);

/* ---- A6/A7: bundle adjustment  (bundle_adjustment._compute_residual bundle_adjustment.py:25-106,
 *             steps 2-3 :167-208, scipy least_squares(method='trf', x_scale='jac') call :200-202;
 *             flat per-observation index arrays as in cvx_pgl::bundleAdjustment pgl_ptz_camera.h:112-120) -------- */
/* n_pose keyframes (keyframe 0 = fixed reference pose), n_landmark rays, n_obs observations in CALLER order:
 * cam_idx[n_obs], lm_idx[n_obs], obs_xy[n_obs*2].  The library keeps a landmark-major sorted copy on the GPU;
 * residuals are always returned in caller order.  Arrays live in `mem`. */
int ptzba_ba_create(ptzba_ctx* ctx, int mem, int n_pose, int n_landmark, int64_t n_obs, const int32_t* cam_idx,
                    const int32_t* lm_idx, const double* obs_xy, double u, double v, ptzba_ba** out);
void ptzba_ba_destroy(ptzba_ba* ba);
/* x = [pose_1 .. pose_{N-1} (pan,tilt,f), landmark_0 .. landmark_{M-1} (theta,phi)]  (bundle_adjustment.py:178-197),
 * reference_pose3 = pose_0.  residual[2*n_obs] = proj - obs, (x,y) interleaved, caller order. */
int ptzba_ba_residual(ptzba_ba* ba, int mem, const double* x, const double* reference_pose3, double* residual);
/* fused residual + analytic Jacobian + J^T J / J^T r assembly.  Outputs (any may be NULL except cost):
 *   residual[2*n_obs]; U[n_pose*6] packed upper (pp,pt,pf,tt,tf,ff); gc[n_pose*3]; V[n_landmark*3] (tt,tp,pp);
 *   gl[n_landmark*2]; cost (host double) = 0.5 * sum r^2.  Keyframe 0's U/gc are zero (fixed pose). */
int ptzba_ba_normal_equations(ptzba_ba* ba, int mem, const double* x, const double* reference_pose3,
                              double* residual, double* U, double* gc, double* V, double* gl, double* cost);

/* pipelined form of the host-buffer pass (no residual vector): `begin` enqueues H2D of x on the problem's own copy stream, the fused
 * pass on the context stream and D2H of the block ranges U/gc[kf_lo, kf_hi) and V/gl[lm_lo, lm_hi) (packed as above, any pointer may
 * be NULL) on the copy stream, and returns at once; `wait` blocks until they have landed and stores the cost.  Host buffers should
 * be pinned.  With several problems in rotation the copies of one overlap the kernels of the next. */
int ptzba_ba_normal_equations_begin(ptzba_ba* ba, const double* x, const double* reference_pose3, int kf_lo, int kf_hi, int lm_lo,
                                    int lm_hi, double* U, double* gc, double* V, double* gl, double* cost);
int ptzba_ba_wait(ptzba_ba* ba);

typedef struct ptzba_ba_options {
    double ftol, xtol, gtol;   /* scipy least_squares tolerances; the reference passes ftol=1e-4 (others default 1e-8) */
    int max_nfev;              /* <=0: 100 * n_params like scipy */
    int verbose;               /* 1: print one line per iteration to stdout (scipy verbose=2 equivalent) */
} ptzba_ba_options;

typedef struct ptzba_ba_report {
    double cost0, cost;        /* 0.5*||r||^2 at x0 / at the solution */
    double optimality;         /* ||J^T r||_inf */
    int status;                /* scipy codes: 0 max_nfev, 1 gtol, 2 ftol, 3 xtol, 4 ftol&xtol */
    int nfev, njev, nit;
    int n_factor;              /* Schur factorisations performed */
    double ms_total;           /* wall time of the solve (host clock) */
} ptzba_ba_report;

/* trust-region (Levenberg-Marquardt / More) solve with scipy-TRF semantics on the Schur-reduced normal equations.
 * x (in/out) lives in `mem`. */
int ptzba_ba_solve(ptzba_ba* ba, int mem, double* x, const double* reference_pose3, const ptzba_ba_options* opt,
                   ptzba_ba_report* report);
/* one LM iteration's worth of device work at fixed damping, for benchmarking (fused pass + Schur + Cholesky +
 * back-substitution + trial residual pass); x is not modified. */
int ptzba_ba_lm_iteration(ptzba_ba* ba, int mem, const double* x, const double* reference_pose3, double alpha,
                          double* out_pred_reduction, double* out_cost_trial);

/* tuning knobs of one problem.  PTZBA_OPT_SCHUR_MODE: how the reduced camera system S = U - sum_l W V^-1 W^T is formed -
 * AUTO (the keyframe-pair-major pair list, built once per problem; falls back to PER_LANDMARK above 65535 keyframes or
 * 2^31 observation pairs), PER_LANDMARK (one warp per landmark, one FP64 RED per block entry per observation pair; no
 * extra memory), PAIR_LIST (same as AUTO). */
#define PTZBA_OPT_SCHUR_MODE 1
/* PTZBA_OPT_FUSED_LM_SHARE: the fused pass is ONE launch with two interleaved CTA roles (landmark-major: residual, V, g_l;
 * keyframe-major: U, g_c); this is the per cent (1..99, default 57) of the CTAs' work budget given to the landmark-major role. */
#define PTZBA_OPT_FUSED_LM_SHARE 3
#define PTZBA_SCHUR_AUTO 0
#define PTZBA_SCHUR_PER_LANDMARK 1
#define PTZBA_SCHUR_PAIR_LIST 2
int ptzba_ba_set_option(ptzba_ba* ba, int option, int value);

/* ---- N3: match graph -> observation list (image_process.py:612-650 landmark ids; bundle_adjustment.py:67-98 residual order).
 * Keypoints of all images are numbered globally ("nodes": node_img[n_node] = image of the node, node_xy[n_node*2] = its pixel);
 * matches are edges (edge_a[k] in image i, edge_b[k] in image j > i) in the reference's visiting order.  out_label[n_node] receives
 * the landmark id of every keypoint (-1: unmatched) exactly as the reference's sequential propagation assigns it, *out_n_landmark
 * their number; out_cam / out_lm [2*n_edge] and out_xy [4*n_edge] (all three or none) the flat observation list that
 * ptzba_ba_create consumes.  Arrays live in `mem`; the two counts are host memory. */
int ptzba_match_graph_to_observations(ptzba_ctx* ctx, int mem, int n_node, const int32_t* node_img, const double* node_xy, int n_edge,
                                      const int32_t* edge_a, const int32_t* edge_b, int32_t* out_label, int32_t* out_n_landmark,
                                      int32_t* out_cam, int32_t* out_lm, double* out_xy);

/* multi-GPU: observations are sharded by keyframe across ranks (each rank creates its ptzba_ba from its shard with
 * the GLOBAL n_pose/n_landmark); landmark blocks and the reduced camera system are summed with ncclAllReduce.
 * unique_id: the 128-byte ncclUniqueId produced by rank 0 (ptzba_comm_unique_id) and distributed by the host
 * plumbing (torch.distributed broadcast). */
int ptzba_comm_unique_id(ptzba_ctx* ctx, void* unique_id128);
int ptzba_comm_init(ptzba_ctx* ctx, const void* unique_id128, int rank, int world_size);
int ptzba_comm_allreduce_f64(ptzba_ctx* ctx, double* device_buf, int64_t count);
/* combines the accumulators of the last fused pass over all ranks (in place, on the stream).  Default: the whole packed
 * arena [cost | U | V | g_c | g_l] is summed and every rank ends with every block.  After ptzba_ba_setup_exchange (one
 * all-reduce of a per-landmark touch mask, once per problem) only the cost and the blocks of landmarks observed by MORE THAN
 * ONE rank travel: blocks of the other landmarks are already complete on the only rank that observes them and each keyframe's
 * U / g_c is complete on the rank that owns the keyframe (the result is distributed, not replicated).  When no landmark is
 * shared the pass needs no collective at all: the cost is then summed lazily by ptzba_ba_get_blocks (collective in that case). */
int ptzba_ba_allreduce(ptzba_ba* ba);
int ptzba_ba_setup_exchange(ptzba_ba* ba, int64_t* n_shared_out);
/* Second multi-GPU mode - replicated data, partitioned work (the distributed SOLVE): every rank creates its ptzba_ba from
 * the WHOLE observation list and then restricts the per-observation kernels to its slice: landmarks [lm_lo, lm_hi) of the
 * landmark-major list (all observations of a landmark stay on one rank, so landmark blocks, Schur pair products and
 * back-substituted landmark steps are complete locally) and positions [cm_lo, cm_hi) of the keyframe-major list
 * (pan-tilt-zoom-slam_b200/dist.py:solve_partition computes balanced slices).  ptzba_ba_normal_equations, ptzba_ba_solve
 * and ptzba_ba_lm_iteration then all-reduce the partial sums themselves (packed blocks, the reduced camera system, the
 * reduced right-hand side, landmark steps, two scalars per trial point) with ncclAllReduce on the context stream, and every
 * rank ends with the same solution.  The residual VECTOR is only written for the rank's own observations.
 * ptzba_comm_init(rank, world_size) must have been called first.  Do not combine with ptzba_ba_allreduce. */
int ptzba_ba_set_partition(ptzba_ba* ba, int rank, int world_size, int lm_lo, int lm_hi, int64_t cm_lo, int64_t cm_hi);
/* copies the accumulators of the last fused pass to host buffers (any may be NULL): U[n_pose*6], gc[n_pose*3],
 * V[n_landmark*3], gl[n_landmark*2], cost */
int ptzba_ba_get_blocks(ptzba_ba* ba, double* U, double* gc, double* V, double* gl, double* cost);

#ifdef __cplusplus
}
#endif
#endif /* PTZBA_H */
