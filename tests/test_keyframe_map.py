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
