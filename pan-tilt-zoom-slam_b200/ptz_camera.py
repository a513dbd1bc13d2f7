"""
PTZCamera drop-in (reference: slam_system/ptz_camera.py:16-325).

Same constructor, attributes and method contracts as the reference class; the ray <-> pixel maps run in the
hand-written CUDA kernels of csrc/projection.cu through the C-ABI.  Angles in degrees, pixels in pixels, float64.
Batched extensions (project_rays_multi, back_project_points) are additive; no reference signature changes.
Out of scope here: 3-D point projection and the homography->PTZ helper.
"""
import ctypes

import numpy as np

from . import _lib


def _rodrigues(rvec):
    """Rotation vector -> matrix (the reference calls cv.Rodrigues for a (3,) base rotation, ptz_camera.py:40-42)."""
    rvec = np.asarray(rvec, dtype=np.float64)
    th = np.linalg.norm(rvec)
    if th < 1e-15:
        return np.eye(3)
    k = rvec / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


class PTZCamera:
    def __init__(self, principal_point, camera_center, base_rotation, displacement=None):
        if displacement is not None:
            assert len(displacement) == 6
        self.principal_point = principal_point
        self.camera_center = camera_center
        base_rotation = np.asarray(base_rotation)
        assert base_rotation.shape == (3, 3) or base_rotation.shape == (3,)
        self.base_rotation = base_rotation if base_rotation.shape == (3, 3) else _rodrigues(base_rotation)
        self.pan = 0.0
        self.tilt = 0.0
        self.focal_length = 2000
        self.displacement = np.zeros(6)
        if displacement is not None:
            self.displacement = displacement
        self.projection_matrix = np.zeros((3, 4))

    # -- small host-side accessors kept for API compatibility (ptz_camera.py:55-152) --------------------------
    def compute_camera_matrix(self):
        return np.array([[self.focal_length, 0, self.principal_point[0]],
                         [0, self.focal_length, self.principal_point[1]],
                         [0, 0, 1]])

    def compute_dispalcement(self):
        fl, wt = self.focal_length, self.displacement
        return np.array([wt[0] + wt[3] * fl, wt[1] + wt[4] * fl, wt[2] + wt[5] * fl])

    def compute_pan_matrix(self):
        """ptz_camera.py:83-93."""
        c, s = np.cos(np.radians(self.pan)), np.sin(np.radians(self.pan))
        return np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]])

    def compute_tilt_matrix(self):
        """ptz_camera.py:95-104."""
        c, s = np.cos(np.radians(self.tilt)), np.sin(np.radians(self.tilt))
        return np.array([[1, 0, 0], [0, c, s], [0, -s, c]])

    def compute_rotation_matrix(self):
        """ptz_camera.py:65-81: R_tilt R_pan R_base."""
        return self.compute_tilt_matrix() @ self.compute_pan_matrix() @ self.base_rotation

    def recompute_matrix(self):
        """ptz_camera.py:117-141: P = K [I | d] [R 0; 0 1] [I -C; 0 1] (3 x 4), d = displacement of the projection centre."""
        pose = np.eye(4)
        pose[0:3, 0:3] = self.compute_rotation_matrix()
        pose[0:3, 3] = -pose[0:3, 0:3] @ np.asarray(self.camera_center, dtype=np.float64)
        disp = np.eye(3, 4)
        disp[:, 3] = self.compute_dispalcement()
        self.projection_matrix = self.compute_camera_matrix() @ disp @ pose
        return self.projection_matrix

    # -- world points (the court model); general-camera helpers off the ray path, host algebra as in the reference ------
    def project_3d_point(self, p):
        """ptz_camera.py:154-165: world point -> (x, y)."""
        pts, _ = self.project_3d_points(np.asarray(p, dtype=np.float64).reshape(1, 3))
        return float(pts[0, 0]), float(pts[0, 1])

    def project_3d_points(self, ps, height=0, width=0):
        """ptz_camera.py:167-189: (points, index) with the same strict in-image filter as project_rays."""
        ps = np.asarray(ps, dtype=np.float64).reshape(-1, 3)
        uvw = np.hstack([ps, np.ones((len(ps), 1))]) @ self.recompute_matrix().T
        assert np.all(uvw[:, 2] != 0.0)
        pts = uvw[:, 0:2] / uvw[:, 2:3]
        if height != 0 and width != 0:
            keep = (0 < pts[:, 0]) & (pts[:, 0] < width) & (0 < pts[:, 1]) & (pts[:, 1] < height)
            return pts[keep], np.nonzero(keep)[0].astype(np.float64)
        return pts, np.ndarray([0])

    def back_project_to_3d_point(self, x, y):
        """ptz_camera.py:236-273: the point of the ground plane z = 0 seen at pixel (x, y) (displacement ignored, as there)."""
        return self.back_project_to_3d_points(np.array([[x, y]], dtype=np.float64))[0]

    def back_project_to_3d_points(self, keypoints):
        """ptz_camera.py:275-285."""
        kp = np.asarray(keypoints, dtype=np.float64).reshape(-1, 2)
        inv_mat = np.linalg.inv(self.compute_camera_matrix() @ self.compute_rotation_matrix())
        center = np.asarray(self.camera_center, dtype=np.float64)
        rays = np.hstack([kp, np.ones((len(kp), 1))]) @ inv_mat.T
        coe = (0.0 - center[2]) / rays[:, 2]
        return rays * coe[:, None] + center

    def get_ptz(self):
        return np.array([self.pan, self.tilt, self.focal_length])

    def set_ptz(self, ptz):
        self.pan, self.tilt, self.focal_length = ptz

    # -- device calls ------------------------------------------------------------------------------------------
    def _ptz_buf(self):
        return _lib.f64([self.pan, self.tilt, self.focal_length])

    def _disp_buf(self):
        d = _lib.f64(self.displacement)
        return d if np.any(d != 0.0) else None

    def project_ray(self, ray):
        """ptz_camera.py:191-210: one ray [theta, phi] -> (x, y)."""
        pts, _ = self.project_rays(np.asarray(ray, dtype=np.float64).reshape(1, 2))
        return float(pts[0, 0]), float(pts[0, 1])

    def project_rays(self, rays, height=0, width=0):
        """ptz_camera.py:212-234: (points[m,2], index[m]); strict in-image filter when height/width are given,
        otherwise all points and an empty index.  index is float64-typed like the reference's."""
        ctx = _lib.get_context()
        rays = _lib.f64(rays).reshape(-1, 2)
        n = rays.shape[0]
        ptz, disp = self._ptz_buf(), self._disp_buf()
        u, v = float(self.principal_point[0]), float(self.principal_point[1])
        if height != 0 and width != 0:
            out = np.empty((n, 2), np.float64)
            idx = np.empty(n, np.int32)
            cnt = ctypes.c_int32(0)
            ctx.check(ctx.lib.ptzba_project_rays_filtered(ctx.handle, _lib.HOST, _lib.ptr(ptz), u, v, _lib.ptr(disp), n,
                                                          _lib.ptr(rays), float(height), float(width), _lib.ptr(out),
                                                          _lib.ptr(idx), ctypes.byref(cnt)))
            m = cnt.value
            return out[:m].copy(), idx[:m].astype(np.float64)
        out = np.empty((n, 2), np.float64)
        ctx.check(ctx.lib.ptzba_project(ctx.handle, _lib.HOST, 1, _lib.ptr(ptz), u, v, _lib.ptr(disp), n,
                                        _lib.ptr(rays), _lib.ptr(out)))
        return out, np.ndarray([0])

    def back_project_to_ray(self, x, y):
        """ptz_camera.py:287-312: pixel -> (theta, phi) in degrees."""
        r = self.back_project_to_rays(np.array([[x, y]], dtype=np.float64))
        return float(r[0, 0]), float(r[0, 1])

    def back_project_to_rays(self, points):
        """ptz_camera.py:314-325: points[n,2] -> rays[n,2]."""
        ctx = _lib.get_context()
        points = _lib.f64(points).reshape(-1, 2)
        n = points.shape[0]
        out = np.empty((n, 2), np.float64)
        ptz, disp = self._ptz_buf(), self._disp_buf()
        ctx.check(ctx.lib.ptzba_backproject(ctx.handle, _lib.HOST, 1, _lib.ptr(ptz), float(self.principal_point[0]),
                                            float(self.principal_point[1]), _lib.ptr(disp), n, _lib.ptr(points), None,
                                            _lib.ptr(out)))
        return out


def project_rays_multi(ptzs, rays, u, v, displacement=None):
    """Additive batched form: every camera ptzs[c] x every ray -> xy[c, n, 2] in one launch."""
    ctx = _lib.get_context()
    ptzs = _lib.f64(ptzs).reshape(-1, 3)
    rays = _lib.f64(rays).reshape(-1, 2)
    out = np.empty((ptzs.shape[0], rays.shape[0], 2), np.float64)
    disp = None if displacement is None else _lib.f64(displacement)
    ctx.check(ctx.lib.ptzba_project(ctx.handle, _lib.HOST, ptzs.shape[0], _lib.ptr(ptzs), float(u), float(v),
                                    _lib.ptr(disp), rays.shape[0], _lib.ptr(rays), _lib.ptr(out)))
    return out
