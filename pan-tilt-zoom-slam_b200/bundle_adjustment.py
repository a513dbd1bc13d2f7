"""
Keyframe bundle adjustment drop-in (reference: slam_system/bundle_adjustment.py).

  _compute_residual(x, n_pose, n_landmark, n_residual, keypoints, src_pt_index, dst_pt_index, landmark_index,
                    u, v, reference_pose, verbose=False)                          bundle_adjustment.py:25-106
  bundle_adjustment(images, image_indices, feature_method, initial_ptzs, center, rotation, u, v, save_path,
                    verbose=False) -> (optimized_landmarks[M,2], keyframes)        bundle_adjustment.py:109-251

The image-free core (reference steps 2-3, :167-208) is `bundle_adjustment_core`; the GPU problem object is
`BAProblem` (flat per-observation arrays, as cvx_pgl::bundleAdjustment pgl_ptz_camera.h:112-120 takes them).
Residuals, analytic Jacobian blocks, J^T J / J^T r assembly, the Schur complement, the Cholesky solve and the
trust-region loop all run in libptzba on the GPU; nothing here computes on the CPU.
"""
import ctypes
import math

import numpy as np

from . import _lib
from .synth import flatten_match_graph


class BAProblem:
    """A BA problem resident on the GPU: n_pose keyframes (pose 0 fixed), n_landmark rays, flat observations."""

    def __init__(self, n_pose, n_landmark, cam_idx, lm_idx, obs_xy, u, v, ctx=None, mem=_lib.HOST):
        self.ctx = ctx or _lib.get_context()
        self.n_pose, self.n_landmark = int(n_pose), int(n_landmark)
        self.u, self.v = float(u), float(v)
        if mem == _lib.HOST:
            cam_idx, lm_idx, obs_xy = _lib.i32(cam_idx), _lib.i32(lm_idx), _lib.f64(obs_xy).reshape(-1, 2)
            self.n_obs = int(cam_idx.shape[0])
            assert lm_idx.shape[0] == self.n_obs and obs_xy.shape[0] == self.n_obs
            pc, pl, po = _lib.ptr(cam_idx), _lib.ptr(lm_idx), _lib.ptr(obs_xy)
        else:   # device pointers (ints) + explicit count
            pc, pl, po, self.n_obs = cam_idx
        h = ctypes.c_void_p()
        self.ctx.check(self.ctx.lib.ptzba_ba_create(self.ctx.handle, mem, self.n_pose, self.n_landmark, self.n_obs,
                                                    pc, pl, po, self.u, self.v, ctypes.byref(h)))
        self.handle = h

    @property
    def n_params(self):
        return 3 * (self.n_pose - 1) + 2 * self.n_landmark

    def set_partition(self, rank, world_size, lm_range, cm_range):
        """Distributed solve (replicated data, partitioned work): this rank visits landmarks lm_range = (lo, hi) of the
        landmark-major list and positions cm_range = (lo, hi) of the keyframe-major list; see dist.solve_partition."""
        self.ctx.check(self.ctx.lib.ptzba_ba_set_partition(self.handle, int(rank), int(world_size), int(lm_range[0]),
                                                           int(lm_range[1]), int(cm_range[0]), int(cm_range[1])))

    def set_option(self, option, value):
        """Tuning knobs (include/ptzba.h: PTZBA_OPT_*), e.g. set_option(_lib.OPT_SCHUR_MODE, _lib.SCHUR_PER_LANDMARK)."""
        self.ctx.check(self.ctx.lib.ptzba_ba_set_option(self.handle, int(option), int(value)))

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.ptzba_ba_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def residual(self, x, reference_pose):
        """r[2*n_obs] = proj - obs in caller order (x,y interleaved)."""
        x = _lib.f64(x)
        assert x.shape[0] == self.n_params
        ref = _lib.f64(reference_pose)
        r = np.empty(2 * self.n_obs, np.float64)
        self.ctx.check(self.ctx.lib.ptzba_ba_residual(self.handle, _lib.HOST, _lib.ptr(x), _lib.ptr(ref), _lib.ptr(r)))
        return r

    def normal_equations(self, x, reference_pose, want_residual=True):
        """Fused pass -> dict(residual, U[N,3,3], gc[N,3], V[M,2,2], gl[M,2], cost)."""
        x = _lib.f64(x)
        assert x.shape[0] == self.n_params
        ref = _lib.f64(reference_pose)
        N, M = self.n_pose, self.n_landmark
        r = np.empty(2 * self.n_obs, np.float64) if want_residual else None
        Up, gc = np.empty((N, 6)), np.empty((N, 3))
        Vp, gl = np.empty((M, 3)), np.empty((M, 2))
        cost = ctypes.c_double(0.0)
        self.ctx.check(self.ctx.lib.ptzba_ba_normal_equations(self.handle, _lib.HOST, _lib.ptr(x), _lib.ptr(ref),
                                                              _lib.ptr(r), _lib.ptr(Up), _lib.ptr(gc), _lib.ptr(Vp),
                                                              _lib.ptr(gl), ctypes.byref(cost)))
        U = np.empty((N, 3, 3))
        U[:, 0, 0], U[:, 0, 1], U[:, 0, 2] = Up[:, 0], Up[:, 1], Up[:, 2]
        U[:, 1, 0], U[:, 1, 1], U[:, 1, 2] = Up[:, 1], Up[:, 3], Up[:, 4]
        U[:, 2, 0], U[:, 2, 1], U[:, 2, 2] = Up[:, 2], Up[:, 4], Up[:, 5]
        V = np.empty((M, 2, 2))
        V[:, 0, 0], V[:, 0, 1], V[:, 1, 0], V[:, 1, 1] = Vp[:, 0], Vp[:, 1], Vp[:, 1], Vp[:, 2]
        return dict(residual=r, U=U, gc=gc, V=V, gl=gl, cost=cost.value)

    def normal_equations_device(self, x_dev_ptr, reference_pose, resid_dev_ptr=None):
        """Fused pass with device-resident x (and optional device residual buffer); asynchronous on the context
        stream, results stay in the problem's accumulators on the GPU (no host copy, no synchronisation)."""
        ref = _lib.f64(reference_pose)
        self.ctx.check(self.ctx.lib.ptzba_ba_normal_equations(self.handle, _lib.DEVICE, _lib.ptr(int(x_dev_ptr)),
                                                              _lib.ptr(ref), _lib.ptr(resid_dev_ptr and int(resid_dev_ptr)),
                                                              None, None, None, None, None))

    def normal_equations_into(self, x, reference_pose, r, Up, gc, Vp, gl):
        """Fused pass through HOST buffers the caller owns (e.g. pinned): x in; r, packed U/gc/V/gl out; returns cost."""
        ref = _lib.f64(reference_pose)
        cost = ctypes.c_double(0.0)
        self.ctx.check(self.ctx.lib.ptzba_ba_normal_equations(self.handle, _lib.HOST, _lib.ptr(x), _lib.ptr(ref),
                                                              _lib.ptr(r), _lib.ptr(Up), _lib.ptr(gc), _lib.ptr(Vp),
                                                              _lib.ptr(gl), ctypes.byref(cost)))
        return cost.value

    def normal_equations_begin(self, x, reference_pose, kf_range, lm_range, Up, gc, Vp, gl, cost_out):
        """Pipelined host-buffer pass: enqueue H2D(x) -> fused pass -> D2H of U/gc[kf_range) and V/gl[lm_range) into the caller's
        (pinned) buffers and return at once; `wait()` blocks until they have landed.  cost_out: ctypes.c_double."""
        ref = _lib.f64(reference_pose)
        self.ctx.check(self.ctx.lib.ptzba_ba_normal_equations_begin(
            self.handle, _lib.ptr(x), _lib.ptr(ref), int(kf_range[0]), int(kf_range[1]), int(lm_range[0]), int(lm_range[1]),
            _lib.ptr(Up), _lib.ptr(gc), _lib.ptr(Vp), _lib.ptr(gl), ctypes.cast(ctypes.byref(cost_out), ctypes.c_void_p)))

    def wait(self):
        self.ctx.check(self.ctx.lib.ptzba_ba_wait(self.handle))

    def lm_iteration(self, x, reference_pose, alpha=0.0):
        """One LM iteration's device work at fixed damping (benchmark unit); returns (predicted_reduction, trial_cost)."""
        x = _lib.f64(x)
        ref = _lib.f64(reference_pose)
        pred, trial = ctypes.c_double(0.0), ctypes.c_double(0.0)
        self.ctx.check(self.ctx.lib.ptzba_ba_lm_iteration(self.handle, _lib.HOST, _lib.ptr(x), _lib.ptr(ref), float(alpha),
                                                          ctypes.byref(pred), ctypes.byref(trial)))
        return pred.value, trial.value

    def solve(self, x0, reference_pose, ftol=1e-8, xtol=1e-8, gtol=1e-8, max_nfev=0, verbose=0):
        """Trust-region solve with scipy-TRF semantics (least_squares(method='trf', x_scale='jac'))."""
        x = _lib.f64(x0).copy()
        assert x.shape[0] == self.n_params
        ref = _lib.f64(reference_pose)
        opt = _lib.BaOptions(ftol, xtol, gtol, int(max_nfev), int(verbose))
        rep = _lib.BaReport()
        self.ctx.check(self.ctx.lib.ptzba_ba_solve(self.handle, _lib.HOST, _lib.ptr(x), _lib.ptr(ref),
                                                   ctypes.byref(opt), ctypes.byref(rep)))
        return x, {k: getattr(rep, k) for k, _ in _lib.BaReport._fields_}


_problem_cache = {}


def _problem_for(keypoints, src_pt_index, dst_pt_index, landmark_index, n_pose, n_landmark, u, v):
    """scipy calls _compute_residual(x, *args) many times with the same match graph: keep it on the GPU."""
    key = (id(keypoints), id(src_pt_index), id(dst_pt_index), id(landmark_index), n_pose, n_landmark, u, v)
    hit = _problem_cache.get(key)
    if hit is not None and hit[0] is keypoints:
        return hit[1]
    cam, lm, xy = flatten_match_graph(keypoints, src_pt_index, dst_pt_index, landmark_index)
    prob = BAProblem(n_pose, n_landmark, cam, lm, xy, u, v)
    _problem_cache.clear()
    _problem_cache[key] = (keypoints, prob)
    return prob


def _compute_residual(x, n_pose, n_landmark, n_residual, keypoints, src_pt_index, dst_pt_index, landmark_index, u, v,
                      reference_pose, verbose=False):
    """bundle_adjustment.py:25-106: residual vector of N-1 free poses and M landmarks (proj - obs, pair-major)."""
    x = np.asarray(x, dtype=np.float64)
    assert x.shape[0] == (n_pose - 1) * 3 + n_landmark * 2
    N = len(keypoints)
    assert n_pose == N
    assert len(src_pt_index) == N and len(dst_pt_index) == N and len(landmark_index) == N
    reference_pose = np.asarray(reference_pose, dtype=np.float64)
    assert reference_pose.shape[0] == 3
    for i in range(N):
        assert len(src_pt_index[i]) == N and len(dst_pt_index[i]) == N and len(landmark_index[i]) == N
    prob = _problem_for(keypoints, src_pt_index, dst_pt_index, landmark_index, n_pose, n_landmark, u, v)
    residual = prob.residual(x, reference_pose)
    assert residual.shape[0] == n_residual
    if verbose:
        err = np.sqrt(residual[0::2] ** 2 + residual[1::2] ** 2).sum()
        print("reprojection error is %f" % (err / (n_residual / 2)))
    return residual


def initial_landmarks(points, src_pt_index, dst_pt_index, landmark_index, n_landmark, initial_ptzs, u, v):
    """bundle_adjustment.py:183-194: every landmark is initialised by back-projecting its source observation with
    the source keyframe's pose; later matches overwrite earlier ones ("last write wins")."""
    from .transformation import TransFunction
    N = len(points)
    last_cam = np.full(n_landmark, -1, np.int64)
    last_xy = np.zeros((n_landmark, 2))
    for i in range(N):
        for j in range(N):
            s, l = src_pt_index[i][j], landmark_index[i][j]
            m = min(len(s), len(dst_pt_index[i][j]), len(l))
            if m == 0:
                continue
            l = np.asarray(l[:m], dtype=np.int64)
            last_cam[l] = i                       # numpy keeps the last assignment for repeated ids, like the loop
            last_xy[l] = np.asarray(points[i])[np.asarray(s[:m], dtype=np.int64), :2]
    rays = np.zeros((n_landmark, 2))
    seen = np.nonzero(last_cam >= 0)[0]
    if len(seen):
        rays[seen] = TransFunction.from_image_to_rays_batch(u, v, np.asarray(initial_ptzs, dtype=np.float64),
                                                            last_xy[seen], last_cam[seen])
    return rays


def bundle_adjustment_core(points, src_pt_index, dst_pt_index, landmark_index, n_landmark, initial_ptzs, u, v,
                           ftol=1e-4, xtol=1e-8, gtol=1e-8, verbose=False):
    """Reference steps 2-3 (bundle_adjustment.py:167-208) without images.
    Returns (all_poses[N,3], optimized_landmarks[M,2], report)."""
    initial_ptzs = np.asarray(initial_ptzs, dtype=np.float64)
    N = len(points)
    assert initial_ptzs.shape == (N, 3)
    n_residual = sum(len(src_pt_index[i][j]) * 4 for i in range(N) for j in range(N))
    if verbose:
        print('residual number is %d.' % n_residual)
    rays0 = initial_landmarks(points, src_pt_index, dst_pt_index, landmark_index, n_landmark, initial_ptzs, u, v)
    x0 = np.concatenate([initial_ptzs[1:].ravel(), rays0.ravel()])
    cam, lm, xy = flatten_match_graph(points, src_pt_index, dst_pt_index, landmark_index)
    prob = BAProblem(N, n_landmark, cam, lm, xy, u, v)
    try:
        x, rep = prob.solve(x0, initial_ptzs[0], ftol=ftol, xtol=xtol, gtol=gtol, verbose=2 if verbose else 0)
    finally:
        prob.close()
    all_poses = np.concatenate([initial_ptzs[0], x[:3 * (N - 1)]]).reshape(N, 3)
    return all_poses, x[3 * (N - 1):].reshape(-1, 2), rep


def bundle_adjustment(images, image_indices, feature_method, initial_ptzs, center, rotation, u, v, save_path,
                      verbose=False, build_matching_graph=None):
    """bundle_adjustment.py:109-251.  Step 1 (SIFT/ORB matching, image_process.build_matching_graph) is the vision
    front-end and is out of scope of this library: pass it as `build_matching_graph`
    (a callable with the reference's signature (images, image_match_mask, feature_method, verbose) -> the 7-tuple
    of bundle_adjustment.py:147-150).  Steps 2-5 follow the reference and return KeyFrame objects (:214-248)."""
    N = len(images)
    initial_ptzs = np.asarray(initial_ptzs, dtype=np.float64)
    center, rotation = np.asarray(center), np.asarray(rotation)
    assert N >= 1
    assert len(image_indices) == N
    assert initial_ptzs.shape[0] == N and initial_ptzs.shape[1] == 3
    assert center.shape[0] == 3 and rotation.shape[0] == 3 and rotation.shape[1] == 3
    assert feature_method == 'sift' or feature_method == 'orb' or feature_method == 'latch'
    if build_matching_graph is None:
        raise NotImplementedError("image matching (image_process.build_matching_graph) is outside this library; "
                                  "pass build_matching_graph=... or call bundle_adjustment_core with a match graph")
    # step 1: pair mask by pan overlap > 5 degrees (bundle_adjustment.py:135-144, util.overlap_pan_angle :49-72)
    mask = [[0] * N for _ in range(N)]
    for i in range(N):
        for j in range(N):
            if overlap_pan_angle(initial_ptzs[i][2], initial_ptzs[i][0], initial_ptzs[j][2], initial_ptzs[j][0], 1280) > 5:
                mask[i][j] = 1
    keypoints, descriptors, points, src_pt_index, dst_pt_index, landmark_index, n_landmark = \
        build_matching_graph(images, mask, feature_method, verbose)
    all_poses, optimized_landmarks, _ = bundle_adjustment_core(points, src_pt_index, dst_pt_index, landmark_index,
                                                               n_landmark, initial_ptzs, u, v, verbose=verbose)
    from .key_frame import KeyFrame
    keyframes = []
    for i in range(N):
        # (local, global) pairs with image i as source, then as destination, de-duplicated through ONE set built from the
        # whole list: the iteration order of a set depends on how it was filled, and it is the order of landmark_index /
        # feature_pts in the keyframe (bundle_adjustment.py:222-241)
        pairs = []
        for j in range(N):
            pairs.extend(zip(src_pt_index[i][j], landmark_index[i][j]))
        for j in range(N):
            pairs.extend(zip(dst_pt_index[j][i], landmark_index[j][i]))
        pairs = set((int(a), int(b)) for a, b in pairs)
        local_index = [p[0] for p in pairs]
        global_index = [p[1] for p in pairs]
        key_frame = KeyFrame(images[i], image_indices[i], center, rotation, u, v, all_poses[i, 0], all_poses[i, 1], all_poses[i, 2])
        key_frame.feature_pts = [keypoints[i][j] for j in local_index]                      # bundle_adjustment.py:243-245
        key_frame.feature_des = None if descriptors is None else np.asarray(descriptors[i]).take(local_index, axis=0)
        key_frame.landmark_index = np.array(global_index, dtype=np.int32)
        keyframes.append(key_frame)
    return optimized_landmarks, keyframes


def overlap_pan_angle(fl_1, pan_1, fl_2, pan_2, im_width):
    """util.py:49-72: overlapped pan angle (degrees) of two views."""
    w = im_width / 2
    d1 = math.atan(w / fl_1) * 180.0 / math.pi
    d2 = math.atan(w / fl_2) * 180.0 / math.pi
    return max(0, min(pan_1 + d1, pan_2 + d2) - max(pan_1 - d1, pan_2 - d2))
