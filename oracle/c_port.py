"""
ORACLE - TEST INFRASTRUCTURE ONLY.  ctypes loader of oracle/libptz_oracle.so (the plain-C CPU restatement in
oracle/ptz_oracle_c.c).  Used by tests/ and by bench.py's cpu_baseline / `--impl reference` legs only.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libptz_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        lib = ctypes.CDLL(_SO)
        P, I, L, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
        lib.oracle_ba_residual.restype = None
        lib.oracle_ba_residual.argtypes = [L, P, P, P, P, P, D, D, P, I]
        lib.oracle_ba_fused.restype = D
        lib.oracle_ba_fused.argtypes = [L, P, P, P, I, I, P, P, D, D, P, P, P, P, P, I]
        lib.oracle_max_threads.restype = I
        _lib = lib
    return _lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def max_threads():
    return int(load().oracle_max_threads())


def ba_fused(poses, rays, cam_idx, lm_idx, obs_xy, u, v, want_residual=True, n_threads=0):
    lib = load()
    poses = np.ascontiguousarray(poses, np.float64); rays = np.ascontiguousarray(rays, np.float64)
    cam_idx = np.ascontiguousarray(cam_idx, np.int32); lm_idx = np.ascontiguousarray(lm_idx, np.int32)
    obs_xy = np.ascontiguousarray(obs_xy, np.float64)
    N, M, n = len(poses), len(rays), len(cam_idx)
    r = np.empty((n, 2)) if want_residual else None
    U = np.empty((N, 6)); gc = np.empty((N, 3)); V = np.empty((M, 3)); gl = np.empty((M, 2))
    cost = lib.oracle_ba_fused(n, _p(cam_idx), _p(lm_idx), _p(obs_xy), N, M, _p(poses), _p(rays), float(u), float(v),
                               _p(r), _p(U), _p(gc), _p(V), _p(gl), int(n_threads))
    return r, U, gc, V, gl, cost


def ba_residual(poses, rays, cam_idx, lm_idx, obs_xy, u, v, n_threads=0):
    lib = load()
    poses = np.ascontiguousarray(poses, np.float64); rays = np.ascontiguousarray(rays, np.float64)
    cam_idx = np.ascontiguousarray(cam_idx, np.int32); lm_idx = np.ascontiguousarray(lm_idx, np.int32)
    obs_xy = np.ascontiguousarray(obs_xy, np.float64)
    r = np.empty((len(cam_idx), 2))
    lib.oracle_ba_residual(len(cam_idx), _p(cam_idx), _p(lm_idx), _p(obs_xy), _p(poses), _p(rays), float(u), float(v), _p(r),
                           int(n_threads))
    return r
