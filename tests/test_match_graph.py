"""Match graph -> landmark ids and N x N index lists (host side) against goldens from the unmodified
reference image_process.build_matching_graph (:510-667) run with a seeded stand-in for the OpenCV detector / matcher
(tests/golden/make_golden.py:gen_match_graph).  Covers the pair mask, pairs dropped at <= 20 matches, pairs thinned to 200
with random.shuffle, transitive id propagation and the "in-consistent matching" branch."""
import os
import random

import numpy as np
import pytest

import ptz_slam_b200  # noqa: F401
from ptz_slam_b200.match_graph import assign_landmark_index, build_matching_graph, keypoints_to_matrix
from ptz_slam_b200.synth import flatten_match_graph

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "match_graph.npz"))


class Kp:
    def __init__(self, xy):
        self.pt = (float(xy[0]), float(xy[1]))


def _front_end(c):
    n = int(G["c%d_n" % c])
    kps = [[Kp(p) for p in G["c%d_kp_%d" % (c, i)]] for i in range(n)]
    which = {id(k): i for i, k in enumerate(kps)}
    images = list(range(n))
    detect = lambda im, method: (kps[im], np.full((len(kps[im]), 4), im, np.float32))      # noqa: E731

    def match(kp1, des1, kp2, des2, method):
        i, j = which[id(kp1)], which[id(kp2)]
        return None, G["c%d_raw1_%d_%d" % (c, i, j)].tolist(), None, G["c%d_raw2_%d_%d" % (c, i, j)].tolist()
    return n, images, detect, match


@pytest.mark.parametrize("c", range(int(G["n_cases"])))
def test_build_matching_graph_golden(c):
    n, images, detect, match = _front_end(c)
    random.seed(int(G["c%d_seed" % c]))
    kp, des, points, src, dst, lmi, lm_num = build_matching_graph(images, G["c%d_mask" % c].tolist(), 'sift', False,
                                                                  detect=detect, match=match)
    assert lm_num == int(G["c%d_landmark_num" % c])
    thinned = dropped = 0
    for i in range(n):
        np.testing.assert_array_equal(points[i], G["c%d_kp_%d" % (c, i)])
        for j in range(n):
            np.testing.assert_array_equal(np.array(src[i][j], np.int64), G["c%d_src_%d_%d" % (c, i, j)])
            np.testing.assert_array_equal(np.array(dst[i][j], np.int64), G["c%d_dst_%d_%d" % (c, i, j)])
            np.testing.assert_array_equal(np.array(lmi[i][j], np.int64), G["c%d_lm_%d_%d" % (c, i, j)])
            if i < j:
                raw = len(G["c%d_raw1_%d_%d" % (c, i, j)])
                thinned += raw > 200 and len(src[i][j]) == 200
                dropped += raw <= 20 and len(src[i][j]) == 0
            else:
                assert src[i][j] == []
    assert thinned >= 1 and (dropped >= 1 or c == 1)
    assert len(src[1][3]) == 0                                  # masked pair
    # the flat observation list the C-ABI takes: two observations per match, ids inside [0, landmark_num)
    cam, lm, xy = flatten_match_graph(points, src, dst, lmi)
    assert len(cam) == 2 * sum(len(src[i][j]) for i in range(n) for j in range(n))
    assert lm.min() >= 0 and lm.max() == lm_num - 1 and len(np.unique(lm)) == lm_num
    assert cam.min() >= 0 and cam.max() < n and xy.shape == (len(cam), 2)


@pytest.mark.parametrize("c", range(int(G["n_cases"])))
def test_inconsistent_match_count_golden(c):
    n = int(G["c%d_n" % c])
    pairs = [(i, j, G["c%d_src_%d_%d" % (c, i, j)], G["c%d_dst_%d_%d" % (c, i, j)])
             for i in range(n) for j in range(i + 1, n) if len(G["c%d_src_%d_%d" % (c, i, j)])]
    nk = [len(G["c%d_kp_%d" % (c, i)]) for i in range(n)]
    _, _, lmi, lm_num, n_bad = assign_landmark_index(n, nk, pairs)
    assert lm_num == int(G["c%d_landmark_num" % c])
    assert n_bad == int(G["c%d_n_inconsistent" % c]) and n_bad > 0


def test_assign_landmark_index_small_cases():
    # chain 0-1, 1-2: the id travels through image 1; a later conflicting match keeps both ids
    src, dst, lmi, m, bad = assign_landmark_index(3, [2, 2, 2], [(0, 1, [0], [1]), (0, 2, [1], [0]), (1, 2, [1, 0], [0, 1])])
    assert m == 3 and bad == 1
    assert lmi[0][1] == [0] and lmi[0][2] == [1] and lmi[1][2] == [0, 2]
    assert src[1][2] == [1, 0] and dst[1][2] == [0, 1] and lmi[2][1] == []
    # nothing matched
    src, dst, lmi, m, bad = assign_landmark_index(2, [3, 3], [])
    assert m == 0 and bad == 0 and src == [[[], []], [[], []]]
    assert keypoints_to_matrix([]).shape == (0, 2)
    with pytest.raises(NotImplementedError):
        build_matching_graph([0, 1], [], 'sift')


# ---- GPU: the same assignment and the flattening on the device (csrc/match_graph.cu) ------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("c", range(int(G["n_cases"])))
def test_device_match_graph_golden(c):
    """Landmark ids, their number, the inconsistent-match count and the flat observation list of the reference golden."""
    from ptz_slam_b200.match_graph import match_graph_to_observations
    n = int(G["c%d_n" % c])
    pairs = [(i, j, G["c%d_src_%d_%d" % (c, i, j)], G["c%d_dst_%d_%d" % (c, i, j)])
             for i in range(n) for j in range(i + 1, n) if len(G["c%d_src_%d_%d" % (c, i, j)])]
    points = [G["c%d_kp_%d" % (c, i)] for i in range(n)]
    nk = [len(p) for p in points]
    labels, lm_num, flat, n_bad = match_graph_to_observations(nk, points, pairs)
    assert lm_num == int(G["c%d_landmark_num" % c]) and n_bad == int(G["c%d_n_inconsistent" % c])
    src, dst, lmi, _, _ = assign_landmark_index(n, nk, pairs)
    for i, j, a, b in pairs:
        np.testing.assert_array_equal(labels[i][np.asarray(a, np.int64)], G["c%d_lm_%d_%d" % (c, i, j)])
    cam, lm, xy = flatten_match_graph(points, src, dst, lmi)
    np.testing.assert_array_equal(flat[0], cam)
    np.testing.assert_array_equal(flat[1], lm)
    np.testing.assert_array_equal(flat[2], xy)


@pytest.mark.gpu
def test_device_match_graph_random_graphs_equal_host_loop():
    """Random graphs with repeated keypoints and conflicting matches: the pointer-forest assignment on the device equals the
    sequential propagation match for match; empty and edge-free inputs are valid."""
    from ptz_slam_b200.match_graph import match_graph_to_observations
    rng = np.random.default_rng(0)
    for trial in range(60):
        n_img = int(rng.integers(2, 8))
        nk = rng.integers(3, 40, n_img)
        pairs = []
        for i in range(n_img):
            for j in range(i + 1, n_img):
                if rng.random() < 0.7:
                    m = int(rng.integers(1, min(nk[i], nk[j]) + 1))
                    pairs.append((i, j, rng.integers(0, nk[i], m).tolist(), rng.integers(0, nk[j], m).tolist()))
        _, _, lmi, lm_num, n_bad = assign_landmark_index(n_img, nk, pairs)
        labels, lm_num_d, _, n_bad_d = match_graph_to_observations(nk, None, pairs)
        assert lm_num_d == lm_num and n_bad_d == n_bad
        for i, j, a, b in pairs:
            assert labels[i][np.asarray(a, np.int64)].tolist() == lmi[i][j]
    labels, m, flat, bad = match_graph_to_observations([3, 2], [np.zeros((3, 2)), np.zeros((2, 2))], [])
    assert m == 0 and bad == 0 and all(np.all(l == -1) for l in labels) and flat[0].shape == (0,)
