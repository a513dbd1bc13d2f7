"""GPU parity at the shapes of BASELINE.json's configs 2, 4 and 5 (config 1 = golden-file tests, config 3 = test_gpu_ba)."""
import numpy as np
import pytest

from oracle import ptz_oracle as O
import ptz_slam_b200  # noqa: F401
from ptz_slam_b200 import synth, _lib
from ptz_slam_b200 import bundle_adjustment as BA
from ptz_slam_b200.ptz_camera import PTZCamera
from ptz_slam_b200.ptz_slam import PtzSlam, BatchedEkfTracker

pytestmark = pytest.mark.gpu
H, W = synth.IMAGE_H, synth.IMAGE_W


def test_cfg2_single_sequence_3k_rays():
    """Config 2 shape: one EKF sequence over 3000 rays (first frames), PtzSlam drop-in vs the oracle of the reference."""
    seq = synth.make_ekf_sequence(3000, 4, seed=1002)
    cam = PTZCamera((synth.PP_U, synth.PP_V), np.zeros(3), np.eye(3))
    cam.set_ptz(seq.ptz_gt[0])
    slam = PtzSlam()
    slam.init_rays(seq.rays0, cam)
    s = O.EkfState(seq.rays0, seq.ptz_gt[0], synth.PP_U, synth.PP_V)
    for k in range(1, 4):
        slam.predict()
        n = slam.ekf_update(seq.obs_xy[k], seq.obs_idx[k], H, W)
        O.ekf_predict(s)
        matched = O.ekf_update(s, seq.obs_xy[k], seq.obs_idx[k], H, W)
        assert n == len(matched) and n > 100
        np.testing.assert_allclose(slam.current_camera.get_ptz(), s.ptz, rtol=1e-9, atol=1e-8)
        np.testing.assert_allclose(slam.rays, s.rays, rtol=1e-9, atol=1e-8)
    np.testing.assert_allclose(slam.state_cov, s.state_cov, rtol=1e-6, atol=1e-11)
    assert np.all(slam.state_cov[0:3, 3:] == 0)


def test_cfg4_batched_sequences_2k_rays():
    """Config 4 shape at a bounded batch: independent sequences x 2000 rays, every sequence equals its own oracle run."""
    n_seq, n_frames = 6, 3
    seqs = [synth.make_ekf_sequence(2000, n_frames, seed=2000 + i) for i in range(n_seq)]
    max_obs = max(len(i) for q in seqs for i in q.obs_idx)
    trk = BatchedEkfTracker(np.stack([q.rays0 for q in seqs]), np.stack([q.ptz_gt[0] for q in seqs]), synth.PP_U, synth.PP_V,
                            max_obs, H, W, jacobian_mode=_lib.JAC_CENTRAL_FD)
    for k in range(1, n_frames):
        trk.step(*trk.pack_observations([q.obs_xy[k] for q in seqs], [q.obs_idx[k] for q in seqs]))
    ptz, vel, rays = trk.get_state()
    for b in (0, n_seq - 1):
        s = O.EkfState(seqs[b].rays0, seqs[b].ptz_gt[0], synth.PP_U, synth.PP_V)
        for k in range(1, n_frames):
            O.ekf_predict(s)
            O.ekf_update(s, seqs[b].obs_xy[k], seqs[b].obs_idx[k], H, W)
        np.testing.assert_allclose(ptz[b], s.ptz, rtol=1e-9, atol=1e-8)
        np.testing.assert_allclose(rays[b], s.rays, rtol=1e-9, atol=1e-8)
        np.testing.assert_allclose(trk.get_cov(b), s.state_cov, rtol=1e-6, atol=1e-11)
    # sequences are independent: a batch of one gives the same answer as the same sequence inside the batch
    solo = BatchedEkfTracker(seqs[2].rays0[None], seqs[2].ptz_gt[0][None], synth.PP_U, synth.PP_V, max_obs, H, W,
                             jacobian_mode=_lib.JAC_CENTRAL_FD)
    for k in range(1, n_frames):
        solo.step(*solo.pack_observations([seqs[2].obs_xy[k]], [seqs[2].obs_idx[k]]))
    p1, _, r1 = solo.get_state()
    np.testing.assert_allclose(p1[0], ptz[2], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(r1[0], rays[2], rtol=1e-12, atol=1e-12)
    trk.close(); solo.close()


def test_cfg5_full_size_properties():
    """Config 5 (1024 kf x 1M rays x 20M obs, 40 degree pan range): size-independent properties of the fused pass."""
    fb = synth.make_flat_ba(1024, 1000000, 20000000, seed=1005, pan_sweep=40.0)
    prob = BA.BAProblem(fb.n_pose, fb.n_landmark, fb.cam_idx, fb.lm_idx, fb.obs_xy, synth.PP_U, synth.PP_V)
    x = fb.x0()
    out = prob.normal_equations(x, fb.ptz_init[0])
    r = out["residual"]
    assert abs(out["cost"] - 0.5 * np.dot(r, r)) <= 1e-11 * out["cost"]
    np.testing.assert_array_equal(prob.residual(x, fb.ptz_init[0]), r)
    # subset parity against the oracle (first 100k observations = first landmarks) incl. their landmark blocks
    poses, rays = O.ba_unpack(x, fb.n_pose, fb.ptz_init[0])
    n_sub = int(np.searchsorted(fb.lm_idx, 5000))
    ro, Uo, gco, Vo, glo, _ = O.ba_normal_equations(poses, rays[:5000], fb.cam_idx[:n_sub], fb.lm_idx[:n_sub], fb.obs_xy[:n_sub],
                                                   synth.PP_U, synth.PP_V)
    np.testing.assert_allclose(r[:2 * n_sub], ro.ravel(), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(out["V"][:5000], Vo, rtol=1e-9, atol=1e-12 * np.abs(Vo).max())
    np.testing.assert_allclose(out["gl"][:5000], glo, rtol=1e-8, atol=1e-11 * np.abs(glo).max())
    # pan column = -theta column: the pan-pan trace over free keyframes equals the theta-theta trace of their observations
    free = fb.cam_idx != 0
    tt = np.zeros(fb.n_landmark)
    # V_tt summed over all landmarks = sum over ALL observations; subtract keyframe 0's part via the oracle on its observations
    k0 = np.nonzero(~free)[0]
    Jc0, Jr0 = O.jacobian_blocks_analytic(poses[0, 0], poses[0, 1], poses[0, 2], rays[fb.lm_idx[k0], 0], rays[fb.lm_idx[k0], 1])
    lhs = out["U"][:, 0, 0].sum()
    rhs = out["V"][:, 0, 0].sum() - np.sum(Jr0[:, :, 0] ** 2)
    assert abs(lhs - rhs) <= 1e-9 * rhs
    assert np.all(out["U"][0] == 0)
    # every block and every residual of the full-size problem against the C port of the reference's pass
    from test_gpu_ba import check_against_c_port
    check_against_c_port(out, fb, x, fb.ptz_init[0], synth.PP_U, synth.PP_V)
    prob.close()
