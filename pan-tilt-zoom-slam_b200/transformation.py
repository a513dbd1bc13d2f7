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
