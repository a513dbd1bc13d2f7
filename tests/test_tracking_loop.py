"""The per-frame loop - PtzSlam.init_system (ptz_slam.py:140-208) and tracking (:390-456) - against a run of the UNMODIFIED
reference over a seeded sequence (tests/golden/make_golden.py:gen_tracking; the two OpenCV calls are replaced on both sides by
tests/synth_front_end.py).  Run 0 pans fast enough to raise new_keyframe; run 1 loses most optical-flow matches from frame 4
on, which walks bad_tracking_cnt to tracking_lost.

CPU: the product's host orchestration with the camera and the EKF update served by the oracle (the checker standing in for the
device).  GPU: the product end to end (projection, back-projection and EKF update through the C-ABI)."""
import os

import numpy as np
import pytest

from oracle import ptz_oracle as O
import ptz_slam_b200  # noqa: F401
from ptz_slam_b200.key_frame import KeyFrame
from ptz_slam_b200.ptz_slam import PtzSlam
from synth_front_end import SyntheticFrontEnd, U, V

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "tracking.npz"))
CC = np.array([13.0099, -14.8109, 6.1790])
BASE_ROT = np.array([1.5804, -0.1186, 0.1249])


def cov_probe_vector(n):
    return np.random.default_rng(n).uniform(0.5, 1.5, n)


class OracleCamera:
    """PTZCamera surface the loop touches, computed by the CPU oracle."""

    def __init__(self, ptz):
        self.principal_point = np.array([U, V])
        self.camera_center, self.base_rotation = CC, BASE_ROT
        self.displacement = np.zeros(6)
        self.set_ptz(ptz)

    def get_ptz(self):
        return np.array([self.pan, self.tilt, self.focal_length])

    def set_ptz(self, ptz):
        self.pan, self.tilt, self.focal_length = ptz

    def project_rays(self, rays, height=0, width=0):
        return O.project_rays(self.pan, self.tilt, self.focal_length, U, V, rays, height, width)

    def back_project_to_rays(self, points):
        return O.back_project_to_rays_vec(self.pan, self.tilt, self.focal_length, U, V, points)


def oracle_ekf_update(self, observed_keypoints, observed_keypoint_index, height, width):
    s = O.EkfState(self.rays, self.current_camera.get_ptz(), U, V, None, self.angle_var, self.f_var, self.observe_var)
    s.state_cov = self.state_cov
    O.ekf_update(s, np.asarray(observed_keypoints, dtype=np.float64), np.asarray(observed_keypoint_index), height, width)
    self.rays, self.state_cov, self.velocity = s.rays, s.state_cov, s.velocity
    self.current_camera.set_ptz(s.ptz)


def _run(c, make_camera, tol):
    fe = SyntheticFrontEnd(int(G["c%d_seed" % c]), int(G["c%d_n_frames" % c]),
                           None if int(G["c%d_bad_from" % c]) < 0 else int(G["c%d_bad_from" % c]))
    slam = PtzSlam(front_end=fe)
    cam0 = make_camera(G["c%d_cam0" % c])
    slam.init_system(fe.image(0), cam0, fe.bounding_box)
    slam.keyframe_map.add_first_keyframe(KeyFrame(None, 0, CC, BASE_ROT, U, V, *cam0.get_ptz()))
    np.testing.assert_allclose(slam.rays, G["c%d_rays_0" % c], rtol=0, atol=tol["ray"])
    np.testing.assert_array_equal(np.diag(slam.state_cov), G["c%d_cov_diag_0" % c])
    np.testing.assert_array_equal(np.asarray(slam.previous_keypoints, np.float64), G["c%d_prev_kp_0" % c])
    n_frames = fe.n_frames
    for k in range(1, n_frames):
        pct = slam.tracking(fe.image(k), 80, fe.bounding_box)
        assert 0 <= pct <= 100
        flags = [slam.new_keyframe, slam.tracking_lost, slam.bad_tracking_cnt, len(slam.cameras)]
        np.testing.assert_array_equal(np.array(flags, np.int64), G["c%d_flags_%d" % (c, k)], err_msg="frame %d" % k)
        np.testing.assert_array_equal(np.asarray(slam.previous_keypoints_index, np.float64), G["c%d_prev_idx_%d" % (c, k)])
        np.testing.assert_allclose(slam.current_camera.get_ptz(), G["c%d_ptz_%d" % (c, k)], rtol=tol["ptz"], atol=1e-8)
        np.testing.assert_allclose(slam.velocity, G["c%d_vel_%d" % (c, k)], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(slam.rays, G["c%d_rays_%d" % (c, k)], rtol=0, atol=tol["ray"])
        np.testing.assert_allclose(slam.previous_keypoints, G["c%d_prev_kp_%d" % (c, k)], rtol=0, atol=tol["px"])
        np.testing.assert_allclose(np.diag(slam.state_cov), G["c%d_cov_diag_%d" % (c, k)], rtol=tol["cov"], atol=1e-12)
        np.testing.assert_allclose(slam.state_cov @ cov_probe_vector(slam.state_cov.shape[0]), G["c%d_cov_probe_%d" % (c, k)],
                                   rtol=tol["cov"], atol=1e-10)
        if k == 1 and c == 0:
            np.testing.assert_allclose(slam.state_cov, G["c0_cov_1"], rtol=tol["cov"], atol=1e-12)
        assert slam.state_cov.shape[0] == 3 + 2 * len(slam.rays)
    np.testing.assert_array_equal(np.asarray(slam.des), G["c%d_des_%d" % (c, n_frames - 1)])
    return slam


@pytest.mark.parametrize("c", range(int(G["n_runs"])))
def test_tracking_loop_host_logic_golden(c, monkeypatch):
    monkeypatch.setattr(PtzSlam, "ekf_update", oracle_ekf_update)
    slam = _run(c, OracleCamera, {"ray": 1e-8, "ptz": 1e-9, "px": 1e-6, "cov": 1e-6})
    if c == 0:
        assert slam.new_keyframe and not slam.tracking_lost
    else:
        assert slam.tracking_lost and len(slam.cameras) < int(G["c1_n_frames"])


def test_tracking_needs_front_end():
    slam = PtzSlam()
    with pytest.raises(NotImplementedError):
        slam.init_system(np.zeros((4, 4, 3), np.uint8), OracleCamera([50.0, -8.0, 3000.0]))
    cam = OracleCamera([50.0, -8.0, 3000.0])
    assert slam.relocalize(None, cam) is cam and not slam.tracking_lost         # one keyframe or none: warning only (:489-494)


@pytest.mark.gpu
@pytest.mark.parametrize("c", range(int(G["n_runs"])))
def test_tracking_loop_device_golden(c):
    from ptz_slam_b200.ptz_camera import PTZCamera

    def make_camera(ptz):
        cam = PTZCamera((U, V), CC, BASE_ROT)
        cam.set_ptz(ptz)
        return cam
    _run(c, make_camera, {"ray": 1e-7, "ptz": 1e-8, "px": 1e-5, "cov": 1e-5})
