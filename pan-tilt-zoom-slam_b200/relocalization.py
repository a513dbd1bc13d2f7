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


def score_hypotheses(ptzs, rays, points, u, v, threshold=2.0):
    """Reprojection of all rays under H candidate poses in one device call; returns (inlier_count[H], mean_error[H])
    with inlier = pixel distance < threshold (the preemptive-RANSAC score of ptz_pose_estimation.cpp)."""
    ptzs = _lib.f64(ptzs).reshape(-1, 3)
    points = np.asarray(points, dtype=np.float64).reshape(-1, 2)
    xy = TransFunction.from_rays_to_image_batch(u, v, ptzs, rays)
    dist = np.sqrt(((xy - points[None]) ** 2).sum(axis=2))
    return (dist < threshold).sum(axis=1), dist.mean(axis=1) if dist.shape[1] else np.zeros(len(ptzs))


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

    def relocalize(self, keyframe, verbose=0):
        """:88-99: refine (pan, tilt, f) of `keyframe` on its descriptor matches into the map; returns the pose [3]."""
        keypoint_index, ray_index = self.find_nearest(keyframe.feature_des)
        pose = np.array([keyframe.pan, keyframe.tilt, keyframe.f], dtype=np.float64)
        if len(ray_index) == 0:
            return pose
        rays = self.global_ray[ray_index]
        points = np.asarray(keyframe.feature_pts)[keypoint_index]
        return refine_pose(pose, rays, points, keyframe.u, keyframe.v, verbose=verbose).x
