"""GPU parity: EKF update (single sequence drop-in and batched) vs the reference golden run and the oracle."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import ptz_oracle as O
import ptz_slam_b200  # noqa: F401
from ptz_slam_b200 import synth, _lib
from ptz_slam_b200.ptz_camera import PTZCamera
from ptz_slam_b200.ptz_slam import PtzSlam, BatchedEkfTracker

pytestmark = pytest.mark.gpu
H, W = synth.IMAGE_H, synth.IMAGE_W
CC = np.array([13.0099, -14.8109, 6.1790])


def _cam(ptz, uv, disp=None):
    c = PTZCamera((uv[0], uv[1]), CC, np.eye(3), disp)
    c.set_ptz(ptz)
    return c


def test_compute_h_jacobian_dropin():
    d = load_golden("h_jacobian.npz")
    for tag, disp in (("nodisp", None), ("disp", d["disp"])):
        slam = PtzSlam()
        slam.cameras = [_cam(d["ptz_" + tag], d["uv"], disp)]
        Hm = slam.compute_h_jacobian(*d["ptz_" + tag], d["rays_" + tag])
        np.testing.assert_allclose(Hm, d["H_" + tag], rtol=1e-9, atol=2e-9)


def test_ekf_update_six_frames_golden():
    """The reference's PtzSlam.ekf_update + predict lines over 6 frames (tests/golden/make_golden.py:gen_ekf)."""
    d = load_golden("ekf.npz")
    slam = PtzSlam()
    slam.init_rays(d["rays0"], _cam(d["ptz0"], d["uv"]))
    for k in range(1, int(d["n_frames"]) + 1):
        slam.predict()
        slam.ekf_update(d["obs_xy_%d" % k], d["obs_idx_%d" % k], H, W)
        np.testing.assert_allclose(slam.current_camera.get_ptz(), d["ptz_%d" % k], rtol=1e-10, atol=1e-9)
        np.testing.assert_allclose(slam.velocity, d["vel_%d" % k], rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(slam.rays, d["rays_%d" % k], rtol=1e-10, atol=1e-9)
        np.testing.assert_allclose(slam.state_cov, d["cov_%d" % k], rtol=1e-7, atol=1e-12)
    # write-back quirk (ptz_slam.py:281-289): pose<->ray and theta<->phi covariances are never written
    assert np.all(slam.state_cov[0:3, 3:] == 0) and np.all(slam.state_cov[3::2, 4::2] == 0)


def test_ekf_analytic_mode_matches_fd_mode():
    d = load_golden("ekf.npz")
    res = []
    for mode in (_lib.JAC_CENTRAL_FD, _lib.JAC_ANALYTIC):
        slam = PtzSlam()
        slam.jacobian_mode = mode
        slam.init_rays(d["rays0"], _cam(d["ptz0"], d["uv"]))
        for k in range(1, 4):
            slam.predict()
            slam.ekf_update(d["obs_xy_%d" % k], d["obs_idx_%d" % k], H, W)
        res.append((slam.current_camera.get_ptz(), slam.rays.copy()))
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-9, atol=1e-8)
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=1e-9, atol=1e-8)


def test_ekf_edge_cases():
    """No observation in view -> nothing changes; observations outside the image are ignored like the reference."""
    seq = synth.make_ekf_sequence(30, 2, seed=5)
    slam = PtzSlam()
    slam.init_rays(seq.rays0, _cam(seq.ptz_gt[0], (synth.PP_U, synth.PP_V)))
    slam.predict()
    P0, r0 = slam.state_cov.copy(), slam.rays.copy()
    n = slam.ekf_update(np.zeros((0, 2)), np.zeros(0, np.int64), H, W)
    assert n == 0 and np.array_equal(slam.state_cov, P0) and np.array_equal(slam.rays, r0)
    assert np.all(slam.velocity == 0)
    with pytest.raises(_lib.PtzbaError):
        slam.ekf_update(np.zeros((1, 2)), np.array([9999]), H, W)


def _run_oracle(seq, n_frames):
    s = O.EkfState(seq.rays0, seq.ptz_gt[0], synth.PP_U, synth.PP_V)
    for k in range(1, n_frames):
        O.ekf_predict(s)
        O.ekf_update(s, seq.obs_xy[k], seq.obs_idx[k], H, W)
    return s


@pytest.mark.parametrize("n_rays,n_frames", [(200, 5), (600, 3)])
def test_batched_tracker_vs_oracle(n_rays, n_frames):
    """Independent sequences with different seeds in one batch: each must equal its own oracle run."""
    seqs = [synth.make_ekf_sequence(n_rays, n_frames, seed=2000 + i) for i in range(4)]
    max_obs = max(len(i) for s in seqs for i in s.obs_idx)
    trk = BatchedEkfTracker(np.stack([s.rays0 for s in seqs]), np.stack([s.ptz_gt[0] for s in seqs]), synth.PP_U, synth.PP_V,
                            max_obs, H, W, jacobian_mode=_lib.JAC_CENTRAL_FD)
    for k in range(1, n_frames):
        xy, ix, cnt = trk.pack_observations([s.obs_xy[k] for s in seqs], [s.obs_idx[k] for s in seqs])
        matched = trk.step(xy, ix, cnt)
        assert np.all(matched > 0)
    ptz, vel, rays = trk.get_state()
    for b, s in enumerate(seqs):
        o = _run_oracle(s, n_frames)
        np.testing.assert_allclose(ptz[b], o.ptz, rtol=1e-9, atol=1e-8)
        np.testing.assert_allclose(vel[b], o.velocity, rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(rays[b], o.rays, rtol=1e-9, atol=1e-8)
        P = trk.get_cov(b)
        np.testing.assert_allclose(P, o.state_cov, rtol=1e-6, atol=1e-11)
        # (accuracy against ground truth is NOT asserted: the reference filter itself drifts once its write-back has made
        #  the covariance indefinite - parity with the reference algorithm is the contract here)
    trk.close()


class _Img:
    """stand-in for an image: add_rays only reads its shape"""
    shape = (H, W, 3)


def _detector_for(points):
    return lambda img, n: (np.asarray(points, dtype=np.float64), None)


def test_resident_state_ray_bookkeeping_matches_host_path():
    """N4 on the device: after an update the filter state lives on the GPU; remove_rays / add_rays / predict then run there
    (ptzba_ekf_batch_remove_rays / _add_rays / _predict_cov).  A twin instance is forced onto the host numpy path (the reference's
    own bookkeeping, golden-tested in test_ray_bookkeeping.py) before every such call: both must agree bit for bit, also when the
    capacity has to grow."""
    seq = synth.make_ekf_sequence(160, 8, seed=77)
    rng = np.random.default_rng(5)
    slams = [PtzSlam(), PtzSlam()]
    for s_ in slams:
        s_.des = None
        s_.init_rays(seq.rays0, _cam(seq.ptz_gt[0], (synth.PP_U, synth.PP_V)))
    n_global = len(seq.rays0)
    alive = np.arange(n_global)                     # global id of every ray currently in the state
    for k in range(1, 8):
        pos = {g: i for i, g in enumerate(alive)}
        sel = [j for j, g in enumerate(seq.obs_idx[k]) if g in pos]
        obs_xy = seq.obs_xy[k][sel]
        obs_ix = np.array([pos[g] for g in seq.obs_idx[k][sel]], np.int64)
        drop = np.sort(rng.choice(len(alive), size=7, replace=False))
        new_pts = np.stack([rng.uniform(30, W - 30, 40), rng.uniform(30, H - 30, 40)], 1)
        for which, s_ in enumerate(slams):
            if which == 1:
                s_.invalidate_device()
            s_.predict()
            n = s_.ekf_update(obs_xy, obs_ix, H, W)
            assert n > 10
            if which == 1:
                s_.invalidate_device()
            else:
                assert s_._dev_valid
            s_.remove_rays(drop)
            if which == 1:
                s_.invalidate_device()
            kp, kp_idx = s_.add_rays(_Img, None, _detector_for(new_pts))
            if which == 0:
                assert s_._dev_valid and s_._rays_stale          # nothing was downloaded for the bookkeeping itself except rays
        alive = np.concatenate([np.delete(alive, drop), -np.arange(1, len(slams[0].rays) - (len(alive) - 7) + 1) - 1000 * k])
        # (not bit for bit: the resident sequence stays on the pivoted-LU route once its S was indefinite, the re-uploaded twin
        # gets a fresh Cholesky attempt every frame - same mathematics, different rounding)
        np.testing.assert_allclose(slams[0].rays, slams[1].rays, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(slams[0].state_cov, slams[1].state_cov, rtol=1e-7, atol=1e-12)
        np.testing.assert_allclose(slams[0].current_camera.get_ptz(), slams[1].current_camera.get_ptz(), rtol=1e-9, atol=1e-9)
        assert slams[0].state_cov.shape == slams[1].state_cov.shape
    assert len(slams[0].rays) > n_global                                   # the device capacity had to grow


def test_cfg2_fifty_frames_vs_oracle():
    """Config 2 (soccer cloud, 3 000 rays, every visible ray observed as in localization_and_mapping.py): 50 consecutive frames of
    predict + update on the RESIDENT state against the oracle of the reference (ptz_slam.py:210-289, :418-426).

    The reference's write-back (:281-289) makes P indefinite from the second frame on, so almost every frame takes the pivoted-LU
    route.  It also makes the recursion itself unstable on long synthetic runs: two runs of the ORACLE whose initial rays differ by
    1e-13 degrees agree to 1e-10 px for ~20 frames and drift apart afterwards (1e-7 px at frame 25, 1e-5 at 35; measured, see
    DESIGN.md section 2).  Parity is therefore asserted in two ways:
      * free running for the first 20 frames: 1e-6 rad on angles and 1e-3 px on the focal length at every frame (BASELINE.json's
        parameter tolerance), rays within 1e-6 rad;
      * all 50 frames one step at a time from the oracle's own state (the device state is re-seeded from the oracle before each
        step): pose within 1e-6 rad / 1e-3 px, rays within 1e-6 rad, and the covariance blocks the reference writes back to 1e-6
        relative (of the largest entry)."""
    n_rays, n_frames = 3000, 50
    seq = synth.make_ekf_sequence(n_rays, n_frames + 1, seed=1002, keep_prob=1.0)
    tol_deg = np.degrees(1e-6)
    cam0 = _cam(seq.ptz_gt[0], (synth.PP_U, synth.PP_V))
    free = PtzSlam()
    free.init_rays(seq.rays0, cam0)
    s = O.EkfState(seq.rays0, seq.ptz_gt[0], synth.PP_U, synth.PP_V)
    step = PtzSlam()
    step.init_rays(seq.rays0, _cam(seq.ptz_gt[0], (synth.PP_U, synth.PP_V)))
    worst_free, worst_step = np.zeros(3), np.zeros(3)
    for k in range(1, n_frames + 1):
        # one step from the oracle's state
        step.rays = s.rays.copy()
        step.state_cov = s.state_cov.copy()
        step.velocity = s.velocity.copy()
        step.cameras[-1].set_ptz(s.ptz.copy())
        step.predict()
        n_step = step.ekf_update(seq.obs_xy[k], seq.obs_idx[k], H, W)
        if k <= 20:
            free.predict()
            n_free = free.ekf_update(seq.obs_xy[k], seq.obs_idx[k], H, W)
        O.ekf_predict(s)
        matched = O.ekf_update(s, seq.obs_xy[k], seq.obs_idx[k], H, W)
        assert n_step == len(matched)
        d = np.abs(step.current_camera.get_ptz() - s.ptz)
        worst_step = np.maximum(worst_step, d)
        assert d[0] < tol_deg and d[1] < tol_deg and d[2] < 1e-3, ("one step", k, d)
        assert np.abs(step.rays - s.rays).max() < tol_deg, ("one step rays", k)
        if k % 10 == 0:
            scale = np.abs(s.state_cov).max()
            assert np.abs(step.state_cov - s.state_cov).max() <= 1e-6 * scale, ("one step cov", k)
        if k <= 20:
            assert n_free == len(matched)
            d = np.abs(free.current_camera.get_ptz() - s.ptz)
            worst_free = np.maximum(worst_free, d)
            assert d[0] < tol_deg and d[1] < tol_deg and d[2] < 1e-3, ("free running", k, d)
            if k == 20:
                assert np.abs(free.rays - s.rays).max() < tol_deg
    print("cfg2: worst |d pan|, |d tilt| (deg), |d f| (px): free running 20 frames", worst_free, "; one step, 50 frames", worst_step)
