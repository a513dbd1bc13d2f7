"""CPU: keyframe-map orchestration against goldens generated from the unmodified reference
(tests/golden/make_golden.py:gen_keyframe_map): util.overlap_pan_angle (util.py:49-72) and Map.good_new_keyframe
(scene_map.py:119-149).  The bundle-adjusting half (add_keyframe_with_ba) needs the GPU: tests/test_gpu_ba.py."""
import os

import numpy as np
import pytest

import ptz_slam_b200  # noqa: F401
from ptz_slam_b200.bundle_adjustment import overlap_pan_angle
from ptz_slam_b200.key_frame import KeyFrame
from ptz_slam_b200.scene_map import Map

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "keyframe_map.npz"))


def test_overlap_pan_angle_golden():
    got = np.array([overlap_pan_angle(a, b, c, d, float(G["im_width"])) for a, b, c, d in zip(G["fl1"], G["p1"], G["fl2"], G["p2"])])
    np.testing.assert_allclose(got, G["overlap"], rtol=1e-12, atol=1e-12)
    assert (got == 0).any() and (got > 0).any()


def _map():
    m = Map('sift')
    for i, q in enumerate(G["kf_ptz"]):
        kf = KeyFrame(None, i, np.zeros(3), np.zeros(3), 640.0, 360.0, q[0], q[1], q[2])
        if i == 0:
            m.add_first_keyframe(kf)
        else:
            m.add_keyframe_without_ba(kf)
    return m


def test_good_new_keyframe_golden():
    m = _map()
    got = np.array([m.good_new_keyframe(c) for c in G["cand"]])
    np.testing.assert_array_equal(got, G["good"])
    got2 = np.array([m.good_new_keyframe(c, 10, 15, float(G["im_width"])) for c in G["cand"]])
    np.testing.assert_array_equal(got2, G["good_custom"])
    assert got.any() and not got.all()


def test_empty_map_and_argument_checks():
    m = Map('orb')
    assert m.good_new_keyframe(np.array([50.0, -8.0, 3000.0])) is False        # scene_map.py:131-133: no keyframes -> False
    with pytest.raises(AssertionError):
        Map('surf')                                                            # scene_map.py:23
    with pytest.raises(AssertionError):
        m.add_first_keyframe("not a keyframe")


def test_get_overlap_index_golden():
    """util.get_overlap_index (util.py:75-96), the matched-ray index plumbing of ekf_update: ragged, empty and disjoint inputs."""
    from ptz_slam_b200.util import get_overlap_index
    for c in range(int(G["n_overlap_cases"])):
        i1, i2 = get_overlap_index(G["ov%d_a" % c], G["ov%d_b" % c])
        np.testing.assert_array_equal(i1, G["ov%d_i1" % c])
        np.testing.assert_array_equal(i2, G["ov%d_i2" % c])


# ---- KeyFrame array conversion and .mat export (key_frame.py:58-107) against files written by the reference ------------------
GM = np.load(os.path.join(os.path.dirname(__file__), "golden", "keyframe_mat.npz"))


class _Kp:
    def __init__(self, xy):
        self.pt = (float(xy[0]), float(xy[1]))


def _keyframe_case(c):
    ptz = GM["c%d_ptz" % c]
    kf = KeyFrame(None, 100 + c, np.array([13.0099, -14.8109, 6.1790]), GM["c%d_rot" % c], 640.0, 360.0, ptz[0], ptz[1], ptz[2])
    if bool(GM["c%d_as_list" % c]):
        kf.feature_pts = [_Kp(p) for p in GM["c%d_pts" % c]]
        kf.feature_des = GM["c%d_des" % c].copy()
    else:
        kf.feature_pts = GM["c%d_pts" % c].copy()
        kf.feature_des = GM["c%d_des" % c].astype(np.float64)
    return kf


@pytest.mark.parametrize("c", range(int(GM["n_cases"])))
def test_keyframe_save_to_mat_golden(c, tmp_path):
    import scipy.io as sio
    kf = _keyframe_case(c)
    path = str(tmp_path / "kf.mat")
    kf.save_to_mat(path)
    d = sio.loadmat(path)
    assert str(np.asarray(d["im_name"]).ravel()[0]) == str(GM["c%d_im_name" % c])
    for k in ("keypoint", "descriptor", "ptz"):
        assert d[k].shape == GM["c%d_mat_%s" % (c, k)].shape
        np.testing.assert_allclose(d[k], GM["c%d_mat_%s" % (c, k)], rtol=1e-15, atol=0)
    # camera = [u, v, f, rodrigues(base_rotation), center]; the rotation vector comes from an arccos/arctan -> 1e-12
    np.testing.assert_allclose(d["camera"], GM["c%d_mat_camera" % c], rtol=0, atol=1e-12)
    # and back
    kf2 = KeyFrame.load_mat(path)
    assert kf2.img_index == 100 + c and kf2.get_feature_num() == len(GM["c%d_pts" % c])
    np.testing.assert_allclose([kf2.pan, kf2.tilt, kf2.f], GM["c%d_ptz" % c], rtol=1e-15)
    rot = GM["c%d_rot" % c]
    if rot.shape == (3, 3):
        np.testing.assert_allclose(kf2.base_rotation, rot, atol=1e-12)
    np.testing.assert_allclose(kf2.feature_pts, GM["c%d_mat_keypoint" % c].reshape(-1, 2), rtol=1e-15)


def test_convert_keypoint_to_array_norm():
    kf = _keyframe_case(0)
    kf.convert_keypoint_to_array(norm=False)
    np.testing.assert_array_equal(kf.feature_des, GM["c0_des"].astype(np.float64))
    assert kf.feature_pts.dtype == np.float64 and kf.feature_pts.shape == (30, 2)
    kf = _keyframe_case(0)
    kf.convert_keypoint_to_array()
    np.testing.assert_allclose(np.linalg.norm(kf.feature_des, axis=1), 1.0, rtol=1e-6)
    np.testing.assert_allclose(kf.feature_des, GM["c0_mat_descriptor"], rtol=1e-15)


# ---- sliding-window bundle adjustment of RandomForestMap (scene_map.py:202-244) ---------------------------------------------
GS = np.load(os.path.join(os.path.dirname(__file__), "golden", "sliding_window.npz"))


def _recording_ba(seed):
    """Same stand-in for bundle_adjustment() as tests/golden/make_golden.py:fake_window_ba."""
    calls = []

    def run(images, image_indices, feature_method, initial_ptzs, center, rotation, u, v, save_path, *a, **k):
        rng = np.random.default_rng(seed + len(calls))
        calls.append((list(image_indices), np.array(initial_ptzs, dtype=np.float64).copy()))
        kfs = []
        for i, idx in enumerate(image_indices):
            kf = KeyFrame(images[i], idx, center, rotation, u, v, initial_ptzs[i][0] + 0.01 * (i + 1),
                          initial_ptzs[i][1] - 0.02, initial_ptzs[i][2] + i)
            n = 0 if i == 2 else 5 + i
            kf.feature_pts = [_Kp(p) for p in rng.uniform(0, 700, (n, 2))]
            kf.feature_des = rng.integers(1, 200, (n, 8)).astype(np.float32)
            kf.landmark_index = np.arange(n, dtype=np.int32)
            kfs.append(kf)
        return rng.uniform(-30, 30, (40, 2)), kfs
    return run, calls


@pytest.mark.parametrize("c", range(int(GS["n_cases"])))
def test_sliding_window_ba_golden(c, capsys):
    from ptz_slam_b200.scene_map import RandomForestMap
    run, calls = _recording_ba(500 + c)
    m = RandomForestMap(bundle_adjustment_fn=run)
    cc, rot = np.array([13.0099, -14.8109, 6.1790]), np.array([1.5804, -0.1186, 0.1249])
    for k in range(int(GS["c%d_n_kf" % c])):
        m.keyframe_list.append(KeyFrame(None, 10 * k + 1, cc, rot, 640.0, 360.0, 40.0 + 2 * k, -8.0 - 0.1 * k, 2500.0 + 50 * k))
    m.bundle_adjustment_processing()
    assert len(calls) == 1
    np.testing.assert_array_equal(calls[0][0], GS["c%d_call_indices" % c])
    np.testing.assert_array_equal(calls[0][1], GS["c%d_call_ptzs" % c])
    np.testing.assert_array_equal([kf.img_index for kf in m.keyframe_list], GS["c%d_result_indices" % c])
    np.testing.assert_array_equal([[kf.pan, kf.tilt, kf.f] for kf in m.keyframe_list], GS["c%d_result_ptz" % c])
    np.testing.assert_array_equal([kf.get_feature_num() for kf in m.keyframe_list], GS["c%d_result_nfeat" % c])
    np.testing.assert_array_equal(m.keyframe_list[-1].feature_pts, GS["c%d_last_pts" % c])
    np.testing.assert_allclose(m.keyframe_list[-1].feature_des, GS["c%d_last_des" % c], rtol=1e-15)
    assert "is not included in the map" in capsys.readouterr().out


def test_random_forest_map_add_keyframe_exports(tmp_path):
    import scipy.io as sio
    from ptz_slam_b200.scene_map import RandomForestMap
    run, calls = _recording_ba(7)
    built = []
    m = RandomForestMap(keyframe_location=str(tmp_path), mat_path_file=str(tmp_path / "list.txt"), create_map=built.append,
                        bundle_adjustment_fn=run)
    cc, rot = np.zeros(3), np.eye(3)
    first = KeyFrame(None, 3, cc, rot, 640.0, 360.0, 50.0, -8.0, 3000.0)
    first.feature_pts, first.feature_des = np.zeros((4, 2)), np.ones((4, 8))
    m.add_keyframe(first)
    assert calls == [] and built == [str(tmp_path / "list.txt")]            # one keyframe: no adjustment (scene_map.py:186)
    second = KeyFrame(None, 9, cc, rot, 640.0, 360.0, 55.0, -8.0, 3000.0)
    m.add_keyframe(second)
    assert len(calls) == 1 and calls[0][0] == [3, 9] and len(built) == 2
    listed = open(str(tmp_path / "list.txt")).read().split()
    assert [os.path.basename(p) for p in listed] == ["3.mat", "9.mat"]
    d = sio.loadmat(listed[1])
    assert d["keypoint"].shape == (6, 2) and d["ptz"].ravel()[0] == pytest.approx(55.02)


def test_random_forest_map_relocalize_and_good_keyframe(tmp_path):
    """scene_map.py:246-298: relocalize(frame, init_ptz) exports the frame and returns the forest's flat pose (the forest is an
    injected callable), good_keyframe applies the overlap rule to the poses stored in the exported keyframe files, add_keyframes
    is the reference's no-op; PtzSlam.relocalize(enable_rf=True) hands the start pose as the second argument (ptz_slam.py:491)."""
    import scipy.io as sio
    from ptz_slam_b200.scene_map import RandomForestMap, Map
    from ptz_slam_b200.bundle_adjustment import overlap_pan_angle
    run, _ = _recording_ba(7)
    seen = []

    def forest(path, init_ptz):
        seen.append((path, list(init_ptz), sio.loadmat(path)["keypoint"].shape))
        return np.array([[51.0], [-7.5], [3100.0]])

    m = RandomForestMap(keyframe_location=str(tmp_path), mat_path_file=str(tmp_path / "list.txt"), bundle_adjustment_fn=run,
                        relocalizer=forest, relocalize_file=str(tmp_path / "lost.mat"))
    cc, rot = np.zeros(3), np.eye(3)
    for idx, pan in ((3, 50.0), (9, 58.0)):
        kf = KeyFrame(None, idx, cc, rot, 640.0, 360.0, pan, -8.0, 3000.0)
        kf.feature_pts, kf.feature_des = np.zeros((4, 2)), np.ones((4, 8))
        m.add_keyframe(kf)
    assert m.add_keyframes([1, 2]) is None and len(m.keyframe_list) == 2
    lost = KeyFrame(None, -1, cc, rot, 640.0, 360.0, 49.0, -8.0, 3000.0)
    lost.feature_pts, lost.feature_des = np.zeros((5, 2)), np.ones((5, 8))
    ptz = m.relocalize(lost, [49.0, -8.0, 3000.0])
    assert ptz.shape == (3,) and list(ptz) == [51.0, -7.5, 3100.0]
    assert seen == [(str(tmp_path / "lost.mat"), [49.0, -8.0, 3000.0], (5, 2))]
    stored = [sio.loadmat(p)["ptz"].ravel() for p in open(str(tmp_path / "list.txt")).read().split()]
    for cand in ([52.0, -8.0, 3000.0], [70.0, -8.0, 3000.0], [58.0, -8.0, 3000.0]):
        mx = max(overlap_pan_angle(cand[2], cand[0], q[2], q[0], 1280) for q in stored)
        assert m.good_keyframe(np.array(cand), 5, 20) == (5 < mx < 20)
    with pytest.raises(RuntimeError):
        RandomForestMap(mat_path_file=str(tmp_path / "list.txt")).relocalize(lost, [0, 0, 1])
    # the tracker passes the start pose on
    from ptz_slam_b200.ptz_slam import PtzSlam
    from ptz_slam_b200.ptz_camera import PTZCamera

    class FE:
        @staticmethod
        def detect_keypoints(img, n):
            return np.zeros((6, 2)), np.ones((6, 8))

    slam = PtzSlam.__new__(PtzSlam)
    slam.front_end, slam.rf_map, slam.keyframe_map, slam.tracking_lost = FE, m, Map('sift'), True
    cam = PTZCamera((640.0, 360.0), cc, rot)
    cam.set_ptz([49.0, -8.0, 3000.0])
    out = slam.relocalize(None, cam, enable_rf=True)
    assert list(out.get_ptz()) == [51.0, -7.5, 3100.0] and seen[-1][1] == [49.0, -8.0, 3000.0] and not slam.tracking_lost


def test_map_save_keyframes_to_mat(tmp_path):
    import scipy.io as sio
    m = _map()
    path = str(tmp_path / "kfs.mat")
    m.save_keyframes_to_mat(path)
    d = sio.loadmat(path)
    assert d["keyframes"].shape[-1] == len(G["kf_ptz"])
    rec = d["keyframes"].ravel()[2]
    np.testing.assert_allclose(rec["ptz"][0, 0].ravel(), G["kf_ptz"][2])
    assert int(rec["index"][0, 0].ravel()[0]) == 2


# ---- util.py experiment helpers: noise model, field grid, pose files, error statistics ------------------------------------------
def test_util_noise_model_and_helpers_golden(tmp_path):
    import random
    from ptz_slam_b200 import util
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "util_noise.npz"))
    W, H = 1280, 720
    random.seed(4242)
    np.testing.assert_array_equal(util.add_gauss(d["pts"], 3.0, W, H), d["gauss"])
    random.seed(4343)
    np.testing.assert_array_equal(util.add_outliers(d["pts"], 1.5, W, H, 35), d["outliers"])
    assert d["gauss"].min() == 0 and d["gauss"][:, 0].max() == W - 1 and d["gauss"][:, 1].max() == H - 1     # clamps exercised
    np.testing.assert_array_equal(util.uniform_point_sample_on_field(118, 70, 7, 5), d["field"])
    mean, std = util.compute_error_data(tuple(d["est"]), tuple(d["gt"]))
    np.testing.assert_allclose(mean, d["err_mean"], rtol=1e-15)
    np.testing.assert_allclose(std, d["err_std"], rtol=1e-14)
    path = str(tmp_path / "pose.mat")
    util.save_camera_pose(d["est"][0], d["est"][1], d["est"][2], path)
    for got, want in zip(util.load_camera_pose(path, separate=True), d["est"]):
        np.testing.assert_array_equal(got, want)


# ---- PTZCamera matrices and world-point helpers (host algebra, ptz_camera.py:65-189, 236-285) ---------------------------------
def test_ptz_camera_matrices_and_world_points_golden():
    from ptz_slam_b200.ptz_camera import PTZCamera
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "camera_3d.npz"))
    cc, base = np.array([13.0099, -14.8109, 6.1790]), np.array([1.5804, -0.1186, 0.1249])
    for c in range(2):
        disp = d["c%d_disp" % c]
        cam = PTZCamera((640.0, 360.0), cc, base, disp if np.any(disp) else None)
        cam.set_ptz(d["c%d_ptz" % c])
        np.testing.assert_allclose(cam.compute_pan_matrix(), d["c%d_pan" % c], rtol=0, atol=1e-15)
        np.testing.assert_allclose(cam.compute_tilt_matrix(), d["c%d_tilt" % c], rtol=0, atol=1e-15)
        np.testing.assert_allclose(cam.compute_rotation_matrix(), d["c%d_R" % c], rtol=0, atol=1e-12)    # own Rodrigues vs OpenCV's
        np.testing.assert_allclose(cam.recompute_matrix(), d["c%d_P" % c], rtol=1e-9, atol=1e-8)
        pts, idx = cam.project_3d_points(d["c%d_field" % c])
        assert len(idx) == 0
        np.testing.assert_allclose(pts, d["c%d_pts_all" % c], rtol=1e-9, atol=1e-7)
        pts, idx = cam.project_3d_points(d["c%d_field" % c], 720, 1280)
        np.testing.assert_array_equal(idx, d["c%d_idx_in" % c])
        np.testing.assert_allclose(pts, d["c%d_pts_in" % c], rtol=1e-9, atol=1e-7)
        assert 0 < len(idx) < len(d["c%d_field" % c])
        x, y = cam.project_3d_point(d["c%d_field" % c][3])
        np.testing.assert_allclose([x, y], d["c%d_pts_all" % c][3], rtol=1e-9, atol=1e-7)
        ground = cam.back_project_to_3d_points(d["c%d_px" % c])
        np.testing.assert_allclose(ground, d["c%d_ground" % c], rtol=1e-9, atol=1e-8)
        np.testing.assert_allclose(cam.back_project_to_3d_point(*d["c%d_px" % c][0]), d["c%d_ground" % c][0], rtol=1e-9, atol=1e-8)
        assert np.all(np.abs(ground[:, 2]) < 1e-9)


def test_transfunction_world_point_helpers_golden():
    from ptz_slam_b200.transformation import TransFunction as T
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "camera_3d.npz"))
    cc, R0 = np.array([13.0099, -14.8109, 6.1790]), d["tf_R0"]
    for k, row in enumerate(d["tf_in"]):
        pan, tilt, f, px, py, wx, wy, wz, theta, phi = row
        world = np.array([wx, wy, wz])
        np.testing.assert_allclose(T.from_3dpoint_to_image(640.0, 360.0, f, pan, tilt, cc, R0, world), d["tf_to_image"][k], rtol=1e-9, atol=1e-7)
        np.testing.assert_allclose(T.from_image_to_3dpoint(640.0, 360.0, f, pan, tilt, cc, R0, (px, py)), d["tf_to_3d"][k], rtol=1e-9, atol=1e-8)
        np.testing.assert_allclose(T.from_3dpoint_to_ray(cc, world, R0), d["tf_to_ray"][k], rtol=1e-12, atol=1e-12)
        rel = T.from_ray_to_relative_3dpoint(theta, phi)
        np.testing.assert_allclose(rel, d["tf_ray_rel"][k], rtol=1e-13)
        np.testing.assert_allclose(T.from_relative_3dpoint_to_image(640.0, 360.0, f, pan, tilt, rel), d["tf_rel_image"][k], rtol=1e-10, atol=1e-8)
        np.testing.assert_allclose(T.from_3dpoint_to_relative_3dpoint(cc, R0, world), d["tf_to_rel"][k], rtol=1e-12, atol=1e-12)
