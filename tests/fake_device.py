"""A test double for the slice of the C-ABI (include/ptzba.h) that PtzSlam, PTZCamera and the relocaliser call, served by the CPU
oracle.  It exists so that the HOST side of the product's device path - what is current where (`_dev_valid`, `_rays_stale`,
`_cov_stale`), when the resident filter state is uploaded, re-created, compacted and grown, how observation buffers are padded -
can be exercised on a machine without a GPU by the same drivers the GPU tests use (tests/test_zx_cfg1_end_to_end.py).

TEST INFRASTRUCTURE ONLY: it lives under tests/, nothing in the package can reach it, and it is not a fallback - a test installs it
explicitly with `install(monkeypatch)`, which replaces `_lib.get_context` for that test.  The functions mirror the argument
conventions of csrc/ekf.cu / projection.cu as the Python binding uses them (raw pointers from `_lib.ptr`, `ctypes.byref` outputs)."""
import ctypes

import numpy as np

from oracle import ptz_oracle as O
from ptz_slam_b200 import _lib


def _view(p, shape, dtype=np.float64):
    """numpy view of caller memory behind a `_lib.ptr` value (None stays None)."""
    if p is None:
        return None
    addr = p.value if hasattr(p, "value") else int(p)
    if addr is None:
        return None
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype)
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class _Batch:
    def __init__(self, prm, rays, ptz, max_obs):
        self.u, self.v = prm.u, prm.v
        self.disp = np.array(list(prm.disp)) if any(prm.disp) else None
        self.observe_var, self.angle_var, self.f_var = prm.observe_var, prm.angle_var, prm.f_var
        self.height, self.width, self.jac_mode = prm.height, prm.width, prm.jac_mode
        self.rays = np.array(rays, dtype=np.float64).reshape(-1, 2)
        n = len(self.rays)
        self.cov = self.angle_var * np.eye(3 + 2 * n)
        self.cov[2, 2] = self.f_var
        self.ptz, self.velocity = np.array(ptz, dtype=np.float64), np.zeros(3)
        self.max_obs = int(max_obs)


class FakeLib:
    """One sequence per batch is all PtzSlam uses; calls return 0 (PTZBA_OK)."""

    def __init__(self):
        self.batches, self.next_handle, self.calls = {}, 1, {}

    def __getattr__(self, name):            # anything not modelled here must not be reached by these tests
        raise AttributeError("fake device: %s is not modelled" % name)

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    def _b(self, h):
        return self.batches[h.value if hasattr(h, "value") else int(h)]

    # ---- projection.cu ---------------------------------------------------------------------------------------------------------
    def ptzba_project(self, ctx, mem, n_cam, ptz, u, v, disp, n, rays, out):
        self._count("project")
        assert mem == _lib.HOST
        ptz, rays, out = _view(ptz, (n_cam, 3)), _view(rays, (n, 2)), _view(out, (n_cam, n, 2))
        d = _view(disp, (6,))
        for c in range(n_cam):
            x, y, _ = O.project_rays_vec(ptz[c, 0], ptz[c, 1], ptz[c, 2], u, v, rays, d)
            out[c, :, 0], out[c, :, 1] = x, y
        return 0

    def ptzba_project_rays_filtered(self, ctx, mem, ptz3, u, v, disp, n, rays, height, width, out_xy, out_idx, count):
        self._count("project_rays_filtered")
        assert mem == _lib.HOST
        ptz, rays = _view(ptz3, (3,)), _view(rays, (n, 2))
        pts, idx = O.project_rays(ptz[0], ptz[1], ptz[2], u, v, rays, height, width, _view(disp, (6,)))
        m = len(idx)
        _view(out_xy, (n, 2))[:m] = pts
        _view(out_idx, (n,), np.int32)[:m] = np.asarray(idx).astype(np.int32)
        count._obj.value = m
        return 0

    def ptzba_backproject(self, ctx, mem, n_cam, ptz, u, v, disp, n, points, cam_idx, out):
        self._count("backproject")
        assert mem == _lib.HOST
        ptz, points, out = _view(ptz, (n_cam, 3)), _view(points, (n, 2)), _view(out, (n, 2))
        ci = _view(cam_idx, (n,), np.int32)
        d = _view(disp, (6,))
        if ci is None:
            out[:] = O.back_project_to_rays_vec(ptz[0, 0], ptz[0, 1], ptz[0, 2], u, v, points, d)
        else:
            for c in np.unique(ci):
                sel = ci == c
                out[sel] = O.back_project_to_rays_vec(ptz[c, 0], ptz[c, 1], ptz[c, 2], u, v, points[sel], d)
        return 0

    def ptzba_h_jacobian_blocks(self, ctx, mem, ptz3, u, v, disp, n, rays, mode, jc, jr):
        self._count("h_jacobian_blocks")
        assert mem == _lib.HOST and mode == _lib.JAC_ANALYTIC and disp is None
        ptz, rays = _view(ptz3, (3,)), _view(rays, (n, 2))
        Jc, Jr = O.jacobian_blocks_analytic(ptz[0], ptz[1], ptz[2], rays[:, 0], rays[:, 1])
        _view(jc, (n, 2, 3))[:] = Jc
        _view(jr, (n, 2, 2))[:] = Jr
        return 0

    # ---- ekf.cu: the resident one-sequence batch ------------------------------------------------------------------------------------
    def ptzba_ekf_batch_create(self, ctx, prm, n_seq, n_ray, max_obs, rays, ptz0, out_handle):
        self._count("create")
        assert n_seq == 1 and 1 <= max_obs <= max(n_ray, 1)
        h = self.next_handle
        self.next_handle += 1
        self.batches[h] = _Batch(prm._obj, _view(rays, (n_ray, 2)), _view(ptz0, (3,)), max_obs)
        out_handle._obj.value = h
        return 0

    def ptzba_ekf_batch_destroy(self, h):
        self._count("destroy")
        self.batches.pop(h.value if hasattr(h, "value") else int(h), None)

    def ptzba_ekf_batch_set(self, h, seq, ptz, velocity, rays, cov):
        self._count("set")
        b = self._b(h)
        assert seq == 0
        n = len(b.rays)
        if ptz is not None:
            b.ptz = _view(ptz, (3,)).copy()
        if velocity is not None:
            b.velocity = _view(velocity, (3,)).copy()
        if rays is not None:
            b.rays = _view(rays, (n, 2)).copy()
        if cov is not None:
            b.cov = _view(cov, (3 + 2 * n, 3 + 2 * n)).copy()
        return 0

    def ptzba_ekf_batch_get(self, h, ptz, velocity, rays):
        b = self._b(h)
        if ptz is not None:
            _view(ptz, (3,))[:] = b.ptz
        if velocity is not None:
            _view(velocity, (3,))[:] = b.velocity
        if rays is not None:
            _view(rays, (len(b.rays), 2))[:] = b.rays
        return 0

    def ptzba_ekf_batch_get_cov(self, h, seq, out):
        self._count("get_cov")
        b = self._b(h)
        s = 3 + 2 * len(b.rays)
        _view(out, (s, s))[:] = b.cov
        return 0

    def ptzba_ekf_batch_get_rays(self, h, seq, out):
        self._count("get_rays")
        b = self._b(h)
        _view(out, (len(b.rays), 2))[:] = b.rays
        return 0

    def ptzba_ekf_batch_n_rays(self, h, seq, n_active, capacity):
        n_active._obj.value = len(self._b(h).rays)
        return 0

    def ptzba_ekf_batch_max_obs(self, h, out):
        out._obj.value = self._b(h).max_obs
        return 0

    def ptzba_ekf_batch_reserve(self, h, ray_capacity, max_obs):
        b = self._b(h)
        b.max_obs = max(b.max_obs, int(max_obs))
        return 0

    def ptzba_ekf_batch_predict_cov(self, h):
        self._count("predict_cov")
        b = self._b(h)
        b.cov[0:3, 0:3] += 5 * np.diag([b.angle_var, b.angle_var, b.f_var])
        return 0

    def ptzba_ekf_batch_update_only(self, h, mem, obs_xy, obs_index, obs_count, matched_out):
        self._count("update_only")
        b = self._b(h)
        assert mem == _lib.HOST and b.jac_mode == _lib.JAC_CENTRAL_FD
        m = int(_view(obs_count, (1,), np.int32)[0])
        assert m <= b.max_obs, "the caller must reserve room for its observations"
        rows = max(m, b.max_obs)             # the binding pads the observation arrays to max_obs
        obs = _view(obs_xy, (rows, 2))[:m].copy()
        idx = _view(obs_index, (rows,), np.int32)[:m].astype(np.int64)
        s = O.EkfState(b.rays, b.ptz, b.u, b.v, b.disp, b.angle_var, b.f_var, b.observe_var)
        s.state_cov, s.velocity = b.cov, b.velocity
        matched = O.ekf_update(s, obs, idx, b.height, b.width)
        b.rays, b.cov, b.ptz, b.velocity = s.rays, s.state_cov, s.ptz, s.velocity
        _view(matched_out, (1,), np.int32)[0] = len(matched)
        return 0

    def ptzba_ekf_batch_remove_rays(self, h, seq, n_del, delete_index):
        self._count("remove_rays")
        b = self._b(h)
        d = _view(delete_index, (n_del,), np.int32).astype(np.int64)
        assert len(np.unique(d)) == len(d) and (np.diff(d) > 0).all(), "ascending, unique ids"
        b.rays = np.delete(b.rays, d, axis=0)
        p = np.stack([2 * d + 3, 2 * d + 4], 1).reshape(-1)
        b.cov = np.delete(np.delete(b.cov, p, axis=0), p, axis=1)
        return 0

    def ptzba_ekf_batch_add_rays(self, h, seq, k, new_rays):
        self._count("add_rays")
        b = self._b(h)
        b.rays = np.vstack([b.rays, _view(new_rays, (k, 2))])
        s_old = b.cov.shape[0]
        cov = np.zeros((s_old + 2 * k, s_old + 2 * k))
        cov[:s_old, :s_old] = b.cov
        d = np.arange(s_old, s_old + 2 * k)
        cov[d, d] = b.angle_var
        b.cov = cov
        return 0


    # ---- ba_kernels.cu / ba_solver.cu: the calls of bundle_adjustment_core (host buffers) -----------------------------------------
    def ptzba_ba_create(self, ctx, mem, n_pose, n_landmark, n_obs, cam_idx, lm_idx, obs_xy, u, v, out_handle):
        self._count("ba_create")
        assert mem == _lib.HOST
        h = self.next_handle
        self.next_handle += 1
        self.batches[h] = dict(N=n_pose, M=n_landmark, cam=_view(cam_idx, (n_obs,), np.int32).copy(), lm=_view(lm_idx, (n_obs,), np.int32).copy(),
                               xy=_view(obs_xy, (n_obs, 2)).copy(), u=u, v=v)
        out_handle._obj.value = h
        return 0

    def ptzba_ba_destroy(self, h):
        self.batches.pop(h.value if hasattr(h, "value") else int(h), None)

    def ptzba_ba_solve(self, h, mem, x, reference_pose3, opt, rep):
        """least_squares(method='trf', x_scale='jac') semantics with the exact Jacobian: oracle.trf_solve."""
        self._count("ba_solve")
        p = self._b(h)
        assert mem == _lib.HOST
        N, M = p["N"], p["M"]
        xv = _view(x, (3 * (N - 1) + 2 * M,))
        ref = _view(reference_pose3, (3,)).copy()

        def fun(z):
            poses, rays = O.ba_unpack(z, N, ref)
            return np.ravel(O.ba_residual_flat(poses, rays, p["cam"], p["lm"], p["xy"], p["u"], p["v"]))

        def jac(z):
            poses, rays = O.ba_unpack(z, N, ref)
            return O.ba_jacobian_sparse(poses, rays, p["cam"], p["lm"]).toarray()

        o, r = opt._obj, rep._obj
        out = O.trf_solve(fun, jac, xv.copy(), ftol=o.ftol, xtol=o.xtol, gtol=o.gtol, max_nfev=o.max_nfev or None)
        xv[:] = out["x"]
        r.cost, r.status, r.nfev, r.njev, r.nit = out["cost"], out["status"], out["nfev"], out["njev"], out["nit"]
        return 0


class FakeContext:
    def __init__(self):
        self.lib, self.handle, self.device = FakeLib(), None, 0

    @staticmethod
    def check(status):
        assert status == 0


def install(monkeypatch):
    """Route `_lib.get_context()` to a fresh fake device for the duration of one test; returns the context (its `.lib.calls`
    counts what the product asked the device to do)."""
    ctx = FakeContext()
    monkeypatch.setattr(_lib, "get_context", lambda device=None: ctx)
    return ctx
