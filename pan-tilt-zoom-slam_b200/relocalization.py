"""
Pose refinement on ray <-> pixel matches for relocalisation.

Reference: slam_system/relocalization.py:22-40 (_compute_residual), :96-189 (relocalization_camera) and
nearest_neighbor.py:18-101 (NNBasedMap).  The landmarks (rays) are fixed and only (pan, tilt, f) of ONE camera moves, so
the work per iteration is one projection of n rays and their 2x3 pose Jacobians - the same device functions the EKF and
bundle adjustment use (ptzba_project, ptzba_h_jacobian_blocks); the 3x3 trust-region algebra stays in scipy's
least_squares with the reference's settings (method='trf', x_scale='jac', ftol=1e-4).

    _compute_residual(pose, rays, points, u, v)      relocalization.py:22-40 / nearest_neighbor.py:65-85
    pose_jacobian(pose, rays, points, u, v)          analytic [2n,3] (the reference lets scipy difference the residual)
    refine_pose(pose, rays, points, u, v)            relocalization.py:186-187 / nearest_neighbor.py:98-99
    score_hypotheses(ptzs, rays, points, u, v, thr)  inlier count of H candidate poses in one launch
                                                     (rf_map/util/ptz_pose_estimation.cpp:95-239 scores hypotheses one by one)
    refine_poses_batched(ptzs, rays, points, ...)    Levenberg-Marquardt on (pan, tilt, f) of H poses at once on the GPU
    ptz_from_two_points(...)                         the two-point minimal solver that seeds the hypotheses
                                                     (rf_map/util/eigen_geometry_util.cpp:126-167)
    preemptive_ransac(rays, points, u, v, ...)       ptz_pose_estimation.cpp:95-239 with scoring and refinement of ALL hypotheses of
                                                     a round in one launch each
    relocalization_camera(map, img, pose, ...)       relocalization.py:96-189, OpenCV front-end injected
    NNBasedMap                                       nearest_neighbor.py:18-101, exact nearest neighbour instead of FLANN
"""
import numpy as np

from . import _lib
from .scene_map import Map
from .ptz_camera import PTZCamera
from .transformation import TransFunction


def _compute_residual(pose, rays, points, u, v):
    """relocalization.py:22-40: [2n] reprojection residual (projection - point) of the fixed rays under `pose`."""
    pose = _lib.f64(pose).reshape(1, 3)
    rays = _lib.f64(rays).reshape(-1, 2)
    if len(rays) == 0:
        return np.zeros(0)
    xy = TransFunction.from_rays_to_image_batch(u, v, pose, rays)[0]
    return (xy - np.asarray(points, dtype=np.float64).reshape(-1, 2)).reshape(-1)


def pose_jacobian(pose, rays, points, u, v):
    """d residual / d (pan, tilt, f), [2n,3], angles per degree (closed form)."""
    ctx = _lib.get_context()
    ptz = _lib.f64(pose).reshape(3)
    rays = _lib.f64(rays).reshape(-1, 2)
    n = rays.shape[0]
    jc = np.empty((n, 2, 3), np.float64)
    jr = np.empty((n, 2, 2), np.float64)
    ctx.check(ctx.lib.ptzba_h_jacobian_blocks(ctx.handle, _lib.HOST, _lib.ptr(ptz), float(u), float(v), None, n,
                                              _lib.ptr(rays), _lib.JAC_ANALYTIC, _lib.ptr(jc), _lib.ptr(jr)))
    return jc.reshape(2 * n, 3)


def refine_pose(pose, rays, points, u, v, ftol=1e-4, verbose=0):
    """relocalization.py:186-187: least_squares(_compute_residual, pose, x_scale='jac', ftol=1e-4, method='trf').
    Returns scipy's OptimizeResult (the reference returns .x)."""
    from scipy.optimize import least_squares
    return least_squares(_compute_residual, np.asarray(pose, dtype=np.float64), jac=pose_jacobian, verbose=verbose,
                         x_scale='jac', ftol=ftol, method='trf', args=(rays, points, u, v))


def score_hypotheses(ptzs, rays, points, u, v, threshold=2.0, sel=None):
    """Reprojection of the (selected) rays under H candidate poses in one device call (one CTA per hypothesis); returns
    (inlier_count[H], mean_error[H]) with inlier = pixel distance <= threshold (the preemptive-RANSAC score of
    ptz_pose_estimation.cpp:163-195, which counts the complement as its loss)."""
    ctx = _lib.get_context()
    ptzs = _lib.f64(ptzs).reshape(-1, 3)
    rays = _lib.f64(rays).reshape(-1, 2)
    points = _lib.f64(points).reshape(-1, 2)
    n, H = rays.shape[0], ptzs.shape[0]
    sel_a = None if sel is None else _lib.i32(sel)
    n_sel = n if sel_a is None else int(sel_a.shape[0])
    outl = np.zeros(H, np.int32)
    err = np.zeros(H, np.float64)
    ctx.check(ctx.lib.ptzba_pose_score(ctx.handle, _lib.HOST, H, _lib.ptr(ptzs), float(u), float(v), n, _lib.ptr(rays), _lib.ptr(points),
                                       n_sel, _lib.ptr(sel_a), float(threshold), _lib.ptr(outl), _lib.ptr(err)))
    return n_sel - outl, err


def refine_poses_batched(ptzs, rays, points, u, v, sel=None, threshold=0.0, min_used=4, max_iter=30, ftol=1e-10):
    """Levenberg-Marquardt on (pan, tilt, f) of H poses at once (one CTA per pose, analytic Jacobian, csrc/pose_refine.cu): the
    batched device form of refine_pose / optimizePTZ.  threshold > 0 restricts every pose to the selected matches that are inliers
    of its incoming value.  Returns (ptzs[H,3], cost[H] = 0.5 sum r^2 over the used matches, n_used[H], iterations[H])."""
    ctx = _lib.get_context()
    out = _lib.f64(ptzs).reshape(-1, 3).copy()
    rays = _lib.f64(rays).reshape(-1, 2)
    points = _lib.f64(points).reshape(-1, 2)
    n, H = rays.shape[0], out.shape[0]
    sel_a = None if sel is None else _lib.i32(sel)
    n_sel = n if sel_a is None else int(sel_a.shape[0])
    cost = np.zeros(H); used = np.zeros(H, np.int32); iters = np.zeros(H, np.int32)
    ctx.check(ctx.lib.ptzba_pose_refine(ctx.handle, _lib.HOST, H, _lib.ptr(out), float(u), float(v), n, _lib.ptr(rays), _lib.ptr(points),
                                        n_sel, _lib.ptr(sel_a), float(threshold), int(min_used), int(max_iter), float(ftol),
                                        _lib.ptr(cost), _lib.ptr(used), _lib.ptr(iters)))
    return out, cost, used, iters


def ptz_from_two_points(pan_tilt1, pan_tilt2, point1, point2, pp):
    """Two-point minimal solver (rf_map/util/eigen_geometry_util.cpp:126-167), vectorised over H samples: the focal length from the
    angle between the two rays against the pixel vectors (cos = (p1.p2 + f^2) / (|(p1,f)| |(p2,f)|), a quadratic in f^2), then pan
    and tilt as the mean of what either match implies.  Arrays [H,2]; returns (ptz[H,3], valid[H])."""
    pt1 = np.asarray(pan_tilt1, dtype=np.float64).reshape(-1, 2)
    pt2 = np.asarray(pan_tilt2, dtype=np.float64).reshape(-1, 2)
    p1 = np.asarray(point1, dtype=np.float64).reshape(-1, 2) - np.asarray(pp, dtype=np.float64)
    p2 = np.asarray(point2, dtype=np.float64).reshape(-1, 2) - np.asarray(pp, dtype=np.float64)
    a, b, c = (p1 * p1).sum(1), (p2 * p2).sum(1), (p1 * p2).sum(1)
    # z axis rotated by the pan / tilt differences: its z component is the cosine of the angle between the two rays
    dp, dt = np.radians(pt2[:, 0] - pt1[:, 0]), np.radians(pt2[:, 1] - pt1[:, 1])
    d = np.cos(dp) * np.cos(dt)
    d2 = d * d
    # d^2 (a + F)(b + F) = (c + F)^2,  F = f^2
    qa, qb, qc = d2 - 1.0, d2 * (a + b) - 2.0 * c, d2 * a * b - c * c
    disc = qb * qb - 4.0 * qa * qc
    valid = (disc >= 0.0) & (np.abs(qa) > 1e-14)
    sq = np.sqrt(np.where(valid, disc, 0.0))
    with np.errstate(divide="ignore", invalid="ignore"):
        r1, r2 = (-qb + sq) / (2.0 * qa), (-qb - sq) / (2.0 * qa)
    # the root that is positive and consistent with the sign of the cosine: (c + F) has the sign of d
    ok1 = (r1 > 0) & ((c + r1) * d > 0)
    ok2 = (r2 > 0) & ((c + r2) * d > 0)
    F = np.where(ok1, r1, np.where(ok2, r2, np.nan))
    valid &= np.isfinite(F)
    f = np.sqrt(np.where(valid, F, 1.0))
    pan = 0.5 * ((pt1[:, 0] - np.degrees(np.arctan2(p1[:, 0], f))) + (pt2[:, 0] - np.degrees(np.arctan2(p2[:, 0], f))))
    tilt = 0.5 * ((pt1[:, 1] + np.degrees(np.arctan2(p1[:, 1], f))) + (pt2[:, 1] + np.degrees(np.arctan2(p2[:, 1], f))))
    return np.stack([pan, tilt, f], 1), valid


def preemptive_ransac(rays, points, u, v, init_ptz, threshold=2.0, sample_number=32, n_iteration=1024, n_keep=512, seed=0,
                      max_iter=20, ftol=1e-8):
    """Preemptive RANSAC over (pan, tilt, f) hypotheses (ptz_pose_estimation.cpp:95-239): hypotheses from two-point samples
    (plus the initial pose); per round a random sample of `sample_number` matches is scored under every hypothesis, the better half
    survives and is re-optimised on its inliers; until one is left.  Scoring and refinement of a round are ONE launch each over all
    hypotheses.  Returns (ptz[3] or None when there are at most 12 matches, like the reference)."""
    rays = _lib.f64(rays).reshape(-1, 2)
    points = _lib.f64(points).reshape(-1, 2)
    n = rays.shape[0]
    if n <= 12:
        return None
    rng = np.random.default_rng(seed)
    k1 = rng.integers(0, n, n_iteration)
    k2 = rng.integers(0, n, n_iteration)
    keep = k1 != k2
    cand, valid = ptz_from_two_points(rays[k1[keep]], rays[k2[keep]], points[k1[keep]], points[k2[keep]], (u, v))
    hyp = np.vstack([np.asarray(init_ptz, dtype=np.float64).reshape(1, 3), cand[valid]])[:n_keep + 1]
    while len(hyp) > 1:
        sel = rng.integers(0, n, sample_number).astype(np.int32)
        inl, _ = score_hypotheses(hyp, rays, points, u, v, threshold, sel=sel)
        order = np.argsort(-inl, kind="stable")                    # loss = outliers: fewest first, ties keep their order
        hyp = hyp[order[:max(len(hyp) // 2, 1)]]
        hyp, _, _, _ = refine_poses_batched(hyp, rays, points, u, v, sel=sel, threshold=threshold, min_used=4, max_iter=max_iter, ftol=ftol)
    return hyp[0]


def select_nearest_keyframe(match_counts):
    """relocalization.py:128-168: the keyframe with strictly the most matches wins (first one on ties); -1 when no
    keyframe has any.  `match_counts[i]` is len(index1) for keyframe i, or None / 0 when it has no keypoints or matches."""
    nearest, best = -1, 0
    for i, c in enumerate(match_counts):
        if c is not None and c > best:
            nearest, best = i, c
    return nearest


def relocalization_camera(map, img, pose, detect=None, match=None, verbose=0):
    """relocalization.py:96-189 with the OpenCV front-end injected:

        detect(img, n_features)                  -> (kp[n,2], des[n,d])       detect_compute_sift_array + bounding-box mask
        match(kp1, des1, kp2, des2)              -> (pt1, index1, pt2, index2)   match_sift_features(..., pts_array=True)

    Finds the keyframe sharing most matches with `img`, turns that keyframe's matched keypoints into rays with the
    keyframe pose and refines `pose` on the ray <-> pixel matches.  Returns the pose unchanged when nothing matches."""
    if detect is None or match is None:
        raise NotImplementedError("feature detection / matching (OpenCV) is outside this library: pass detect= and match=")
    kp, des = detect(img, 300)
    counts = []
    for keyframe in map.keyframe_list:
        keyframe_kp, keyframe_des = detect(keyframe.img, 300)
        if len(keyframe_kp) == 0:
            counts.append(None)
            continue
        _, index1, _, _ = match(keyframe_kp, keyframe_des, kp, des)
        counts.append(None if index1 is None else len(index1))
    nearest = select_nearest_keyframe(counts)
    if nearest == -1:
        if verbose:
            print("No matching keyframe!")
        return np.asarray(pose)
    keyframe = map.keyframe_list[nearest]
    # _recompute_matching_ray (:43-93): denser detection, match image -> keyframe, keyframe pixels -> rays
    kp, des = detect(img, 1000)
    keyframe_kp, keyframe_des = detect(keyframe.img, 1000)
    pt1, _, pt2, _ = match(kp, des, keyframe_kp, keyframe_des)
    rays = TransFunction.from_image_to_rays_batch(keyframe.u, keyframe.v, [keyframe.pan, keyframe.tilt, keyframe.f], pt2)
    return refine_pose(pose, rays, pt1, keyframe.u, keyframe.v, verbose=verbose).x


class NNBasedMap(Map):
    """nearest_neighbor.py:18-101: map of (ray, descriptor) pairs; a lost frame is matched by descriptor and its pose
    refined on the matched rays.  The reference's FLANN kd-tree (approximate) is replaced by exact nearest neighbour."""
    MAX_DIST = 2000             # nearest_neighbor.py:38 (FLANN reports squared L2 distances)

    def __init__(self):
        super(NNBasedMap, self).__init__('sift')
        self.global_des = np.ndarray([0, 128], dtype=np.float32)

    def find_nearest(self, des):
        """:33-43 -> (matched_keypoint_index, matched_ray_index)."""
        des = np.asarray(des, dtype=np.float64)
        ref = np.asarray(self.global_des, dtype=np.float64)
        if len(des) == 0 or len(ref) == 0:
            return [], []
        d2 = (des ** 2).sum(1)[:, None] - 2.0 * des @ ref.T + (ref ** 2).sum(1)[None]
        nearest = d2.argmin(axis=1)
        keep = np.nonzero(d2[np.arange(len(des)), nearest] < self.MAX_DIST)[0]
        return keep.tolist(), nearest[keep].tolist()

    def add_keyframe_without_ba(self, keyframe, verbose=False):
        """:45-55: every keypoint of the keyframe becomes a ray of the map (back-projection on the GPU)."""
        super(NNBasedMap, self).add_keyframe_without_ba(keyframe, verbose)
        camera = PTZCamera((keyframe.u, keyframe.v), keyframe.center, keyframe.base_rotation)
        camera.set_ptz((keyframe.pan, keyframe.tilt, keyframe.f))
        rays = camera.back_project_to_rays(keyframe.feature_pts)
        self.global_ray = np.vstack([self.global_ray, rays])
        des = np.asarray(keyframe.feature_des)
        self.global_des = np.vstack([self.global_des.reshape(-1, des.shape[1]) if len(self.global_des) else
                                     np.zeros((0, des.shape[1]), des.dtype), des])

    def add_keyframes(self, keyframe_list):
        for keyframe in keyframe_list:
            self.add_keyframe_without_ba(keyframe)

    compute_residual = staticmethod(_compute_residual)          # :65-85

    def relocalize(self, keyframe, init_ptz=None, verbose=0):
        """:88-99: refine (pan, tilt, f) of `keyframe` on its descriptor matches into the map; returns the pose [3].
        `init_ptz` is accepted for PtzSlam.relocalize, which hands every rf_map the start pose (ptz_slam.py:491); the start
        pose used is the keyframe's own, as in the reference."""
        keypoint_index, ray_index = self.find_nearest(keyframe.feature_des)
        pose = np.array([keyframe.pan, keyframe.tilt, keyframe.f], dtype=np.float64)
        if len(ray_index) == 0:
            return pose
        rays = self.global_ray[ray_index]
        points = np.asarray(keyframe.feature_pts)[keypoint_index]
        return refine_pose(pose, rays, points, keyframe.u, keyframe.v, verbose=verbose).x
