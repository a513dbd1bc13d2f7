"""GPU parity: projection / back-projection / measurement Jacobian through the C-ABI vs reference goldens + oracle."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import ptz_oracle as O
import ptz_slam_b200  # noqa: F401
from ptz_slam_b200 import synth, _lib
from ptz_slam_b200.ptz_camera import PTZCamera, project_rays_multi
from ptz_slam_b200.transformation import TransFunction

pytestmark = pytest.mark.gpu
H, W = synth.IMAGE_H, synth.IMAGE_W
CC = np.array([13.0099, -14.8109, 6.1790])
# tolerance of BASELINE.json: relative 1e-9 (plus 1e-9 px absolute for values that pass through zero)
RTOL, ATOL = 1e-9, 1e-9


def cam_for(ptz, uv, disp=None):
    c = PTZCamera((uv[0], uv[1]), CC, np.eye(3), disp)
    c.set_ptz(ptz)
    return c


def test_project_ray_golden():
    d = load_golden("projection.npz")
    for tag, disp in (("nodisp", None), ("disp", d["disp"])):
        g = d["project_ray_" + tag]
        xy = project_rays_multi(d["ptzs"], d["rays"], d["uv"][0], d["uv"][1], disp)
        np.testing.assert_allclose(xy, g, rtol=RTOL, atol=ATOL)
        cam = cam_for(d["ptzs"][1], d["uv"], disp)
        np.testing.assert_allclose(cam.project_ray(d["rays"][7]), g[1, 7], rtol=RTOL, atol=ATOL)


def test_from_ray_to_image_golden():
    d = load_golden("projection.npz")
    xy = TransFunction.from_rays_to_image_batch(d["uv"][0], d["uv"][1], d["ptzs"], d["rays"])
    # rays behind the camera (z <= 0) are outside the model's domain: from_ray_to_image mirrors y there
    g = d["from_ray_to_image"]
    same = np.abs(g - d["project_ray_nodisp"]).max(axis=2) < 1e-6
    assert same.sum() > 200
    np.testing.assert_allclose(xy[same], g[same], rtol=RTOL, atol=ATOL)
    p = d["ptzs"][0]
    np.testing.assert_allclose(TransFunction.from_ray_to_image(d["uv"][0], d["uv"][1], p[2], p[0], p[1], 20.0, -10.0),
                               O.from_ray_to_image(d["uv"][0], d["uv"][1], p[2], p[0], p[1], 20.0, -10.0), rtol=RTOL)


def test_project_rays_filter_contract():
    d = load_golden("projection.npz")
    for c, ptz in enumerate(d["ptzs"]):
        cam = cam_for(ptz, d["uv"])
        pts, idx = cam.project_rays(d["prs_rays_%d" % c], H, W)
        assert idx.dtype == np.float64
        np.testing.assert_array_equal(idx, d["prs_index_%d" % c])
        np.testing.assert_allclose(pts, d["prs_points_%d" % c], rtol=RTOL, atol=ATOL)
        pts2, idx2 = cam.project_rays(d["prs_rays_%d" % c][:17])
        assert len(idx2) == 0 and pts2.shape == (17, 2)
        np.testing.assert_allclose(pts2, d["prs_all_points_%d" % c], rtol=RTOL, atol=ATOL)
    # empty input
    pts, idx = cam.project_rays(np.zeros((0, 2)), H, W)
    assert pts.shape == (0, 2) and len(idx) == 0


def test_project_rays_filter_large_ordered():
    rng = np.random.default_rng(3)
    rays = np.stack([rng.uniform(30, 90, 100003), rng.uniform(-20, 5, 100003)], 1)
    cam = cam_for([60.0, -8.0, 3000.0], (640.0, 360.0))
    pts, idx = cam.project_rays(rays, H, W)
    po, io = O.project_rays(60.0, -8.0, 3000.0, 640.0, 360.0, rays, H, W)
    np.testing.assert_array_equal(idx, io)
    np.testing.assert_allclose(pts, po, rtol=RTOL, atol=ATOL)
    assert np.all(np.diff(idx) > 0)


def test_back_projection_golden():
    d = load_golden("backprojection.npz")
    for tag, disp in (("nodisp", None), ("disp", d["disp"])):
        for c, ptz in enumerate(d["ptzs"]):
            cam = cam_for(ptz, d["uv"], disp)
            r = cam.back_project_to_rays(d["points"])
            np.testing.assert_allclose(r, d["back_project_" + tag][c], rtol=RTOL, atol=1e-10)
    for c, ptz in enumerate(d["ptzs"]):
        r = TransFunction.from_image_to_rays_batch(d["uv"][0], d["uv"][1], ptz, d["points"])
        np.testing.assert_allclose(r, d["from_image_to_ray"][c], rtol=RTOL, atol=1e-10)
    ptz = d["ptzs"][0]
    np.testing.assert_allclose(cam_for(ptz, d["uv"]).back_project_to_ray(*d["points"][5]), d["back_project_nodisp"][0, 5],
                               rtol=RTOL, atol=1e-10)


def test_round_trip_full_size():
    """Size-independent property at config-4 scale: back_project(project(ray)) == ray for 4096 x 2000 pairs."""
    rng = np.random.default_rng(4)
    ptzs = np.stack([rng.uniform(45, 75, 4096), rng.uniform(-10, -6, 4096), rng.uniform(1900, 4200, 4096)], 1)
    rays = np.stack([rng.uniform(50, 70, 2000), rng.uniform(-11, -5, 2000)], 1)
    xy = project_rays_multi(ptzs, rays, 640.0, 360.0)
    cam_idx = np.repeat(np.arange(4096, dtype=np.int32), 2000)
    back = TransFunction.from_image_to_rays_batch(640.0, 360.0, ptzs, xy.reshape(-1, 2), cam_idx).reshape(4096, 2000, 2)
    np.testing.assert_allclose(back, np.broadcast_to(rays, back.shape), rtol=0, atol=1e-9)


def test_h_jacobian_golden():
    d = load_golden("h_jacobian.npz")
    ctx = _lib.get_context()
    for tag, disp in (("nodisp", None), ("disp", d["disp"])):
        ptz, rays, Hg = _lib.f64(d["ptz_" + tag]), _lib.f64(d["rays_" + tag]), d["H_" + tag]
        n = len(rays)
        dd = None if disp is None else _lib.f64(disp)
        for mode, rtol, atol in ((_lib.JAC_CENTRAL_FD, 1e-9, 2e-9), (_lib.JAC_ANALYTIC, 2e-9, 2e-9)):
            Hd = np.empty((2 * n, 3 + 2 * n))
            ctx.check(ctx.lib.ptzba_h_jacobian_dense(ctx.handle, _lib.HOST, _lib.ptr(ptz), d["uv"][0], d["uv"][1],
                                                     _lib.ptr(dd), n, _lib.ptr(rays), mode, _lib.ptr(Hd)))
            np.testing.assert_allclose(Hd, Hg, rtol=rtol, atol=atol)
            assert np.count_nonzero(Hd) <= 10 * n
