"""
TransFunction drop-in for the two hot-path members (reference: slam_system/transformation.py:99-175).
from_ray_to_image / from_image_to_ray are the displacement-free twins of PTZCamera.project_ray /
back_project_to_ray; both run on the GPU through libptzba.  Array-valued *_batch forms are additive.
"""
import numpy as np

from . import _lib


class TransFunction:
    @staticmethod
    def from_ray_to_image(u, v, f, c_p, c_t, p, t):
        """transformation.py:99-135 -> (x, y)."""
        xy = TransFunction.from_rays_to_image_batch(u, v, np.array([[c_p, c_t, f]], dtype=np.float64),
                                                    np.array([[p, t]], dtype=np.float64))
        return float(xy[0, 0, 0]), float(xy[0, 0, 1])

    @staticmethod
    def from_image_to_ray(u, v, f, c_p, c_t, x, y):
        """transformation.py:137-175 -> (theta, phi) in degrees."""
        r = TransFunction.from_image_to_rays_batch(u, v, np.array([c_p, c_t, f], dtype=np.float64),
                                                   np.array([[x, y]], dtype=np.float64))
        return float(r[0, 0]), float(r[0, 1])

    @staticmethod
    def from_rays_to_image_batch(u, v, ptzs, rays):
        """ptzs[c,3] (pan,tilt,f) x rays[n,2] -> xy[c,n,2]."""
        ctx = _lib.get_context()
        ptzs = _lib.f64(ptzs).reshape(-1, 3)
        rays = _lib.f64(rays).reshape(-1, 2)
        out = np.empty((ptzs.shape[0], rays.shape[0], 2), np.float64)
        ctx.check(ctx.lib.ptzba_project(ctx.handle, _lib.HOST, ptzs.shape[0], _lib.ptr(ptzs), float(u), float(v), None,
                                        rays.shape[0], _lib.ptr(rays), _lib.ptr(out)))
        return out

    @staticmethod
    def from_image_to_rays_batch(u, v, ptz, points, cam_idx=None):
        """points[n,2] -> rays[n,2]; ptz is [3] or, with cam_idx[n], a [c,3] table."""
        ctx = _lib.get_context()
        ptz = _lib.f64(ptz).reshape(-1, 3)
        points = _lib.f64(points).reshape(-1, 2)
        out = np.empty_like(points)
        ci = None if cam_idx is None else _lib.i32(cam_idx)
        ctx.check(ctx.lib.ptzba_backproject(ctx.handle, _lib.HOST, ptz.shape[0], _lib.ptr(ptz), float(u), float(v), None,
                                            points.shape[0], _lib.ptr(points), _lib.ptr(ci), _lib.ptr(out)))
        return out

    # -- world-point helpers ("general slam", transformation.py:23-97, 178-249): a handful of 3-vector products per call,
    #    host algebra exactly as in the reference (they feed the court-model overlays, not the ray hot path) ---------------
    @staticmethod
    def _pan_tilt_rotation(p, t):
        cp, sp = np.cos(np.radians(p)), np.sin(np.radians(p))
        ct, st = np.cos(np.radians(t)), np.sin(np.radians(t))
        return np.array([[1, 0, 0], [0, ct, st], [0, -st, ct]]) @ np.array([[cp, 0, -sp], [0, 1, 0], [sp, 0, cp]])

    @staticmethod
    def from_3dpoint_to_image(u, v, f, p, t, c, base_r, pos):
        """transformation.py:23-55: world point -> (x, y) through K R_tilt R_pan R_base (pos - c)."""
        k = np.array([[f, 0, u], [0, f, v], [0, 0, 1]])
        q = k @ (TransFunction._pan_tilt_rotation(p, t) @ np.asarray(base_r) @ (np.asarray(pos) - np.asarray(c)))
        return q[0] / q[2], q[1] / q[2]

    @staticmethod
    def from_image_to_3dpoint(u, v, f, p, t, c, base_r, point2d):
        """transformation.py:58-97: the ground-plane (z = 0) point seen at a pixel."""
        k = np.array([[f, 0, u], [0, f, v], [0, 0, 1]])
        inv_mat = np.linalg.inv(k @ TransFunction._pan_tilt_rotation(p, t) @ np.asarray(base_r))
        d = inv_mat @ np.array([point2d[0], point2d[1], 1.0])
        return d * ((0.0 - c[2]) / d[2]) + np.asarray(c)

    @staticmethod
    def from_3dpoint_to_ray(proj_center, pos, base_r):
        """transformation.py:178-191: world point -> (theta, phi) in degrees."""
        x, y, z = np.asarray(base_r) @ (np.asarray(pos) - np.asarray(proj_center))
        return float(np.degrees(np.arctan(x / z))), float(np.degrees(np.arctan(-y / np.sqrt(x * x + z * z))))

    @staticmethod
    def from_ray_to_relative_3dpoint(t, p):
        """transformation.py:194-205: ray (theta, phi) -> direction [tan theta, -tan phi sqrt(tan^2 theta + 1), 1]."""
        tt = np.tan(np.radians(t))
        return np.array([tt, -np.tan(np.radians(p)) * np.sqrt(tt * tt + 1), 1])

    @staticmethod
    def from_relative_3dpoint_to_image(u, v, f, p, t, pos):
        """transformation.py:208-236: camera-frame direction -> (x, y) through K R_tilt R_pan."""
        q = np.array([[f, 0, u], [0, f, v], [0, 0, 1]]) @ (TransFunction._pan_tilt_rotation(p, t) @ np.asarray(pos))
        return q[0] / q[2], q[1] / q[2]

    @staticmethod
    def from_3dpoint_to_relative_3dpoint(c, base_r, pos):
        """transformation.py:239-249: world point -> camera-frame direction with unit depth."""
        q = np.asarray(base_r) @ (np.asarray(pos) - np.asarray(c))
        return q / q[2]
