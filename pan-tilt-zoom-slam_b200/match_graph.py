"""
Match graph -> bundle-adjustment observation format (host side).

Reference: slam_system/image_process.py:510-667 (build_matching_graph).  Its steps 1-2 are the OpenCV front-end (feature
detection and pair-wise matching); steps 4-5 turn the pair-wise matches into global landmark ids and the N x N index lists
that bundle_adjustment.py:147-150 consumes.  Here the front-end is a pair of injected callables and steps 4-5 are
array code:

    assign_landmark_index(n_images, n_keypoints, pairs)   image_process.py:612-650   landmark ids + N x N lists
    build_matching_graph(images, image_match_mask, feature_method, verbose, detect=, match=)   :510-667   same 7-tuple
    keypoints_to_matrix(keypoints)                        :653-661                    list of KeyPoint -> [n,2]

The landmark id of a keypoint is decided greedily in visiting order (pair (i,j) ascending, matches in list order): a match
inherits the id either end already has, otherwise opens a new id; when both ends already carry different ids the match is
kept and only reported (":624 in-consistent matching").  `assign_landmark_index` is that loop on the host;
`match_graph_to_observations` is the same assignment AND the flattening into the observation list of bundle adjustment on the GPU
(csrc/match_graph.cu: a forest of first-match pointers resolved by pointer jumping reproduces the sequential result exactly); the
landmark-major sort then happens in ptzba_ba_create.
"""
import ctypes
import random

import numpy as np

MIN_MATCH_NUM = 20          # image_process.py:568  pairs with <= 20 matches are dropped
MAX_MATCH_NUM = 200         # image_process.py:569  larger pairs are randomly thinned to 200


def keypoints_to_matrix(key_points):
    """image_process.py:653-661: [n,2] pixel matrix from objects with a .pt attribute (cv2.KeyPoint) or from an array."""
    if isinstance(key_points, np.ndarray):
        return np.asarray(key_points, dtype=np.float64).reshape(-1, 2)
    out = np.zeros((len(key_points), 2))
    for k, p in enumerate(key_points):
        out[k] = p.pt
    return out


def assign_landmark_index(n_images, n_keypoints, pairs, verbose=False):
    """image_process.py:612-650.

    n_keypoints[i] : number of keypoints of image i
    pairs          : iterable of (i, j, index1, index2) in the reference's visiting order (i ascending, j ascending)
    Returns (src_pt_index, dst_pt_index, landmark_index, landmark_num, n_inconsistent); the first three are N x N lists of
    lists exactly as the reference builds them (empty list where a pair has no edge)."""
    lm_of = [np.full(int(n_keypoints[i]), -1, np.int64) for i in range(n_images)]
    pairs = [(int(i), int(j), list(a), list(b)) for i, j, a, b in pairs]
    g_index = 0
    n_inconsistent = 0
    for i, j, index1, index2 in pairs:
        li, lj = lm_of[i], lm_of[j]
        for a, b in zip(index1, index2):
            ga, gb = li[a], lj[b]
            if ga >= 0 and gb >= 0:
                if ga != gb:
                    n_inconsistent += 1
                    if verbose:
                        print("Warning: in-consistent matching result! (%d %d) <--> (%d %d)" % (i, a, j, b))
            elif ga >= 0:
                lj[b] = ga
            elif gb >= 0:
                li[a] = gb
            else:
                li[a] = lj[b] = g_index
                g_index += 1
    src_pt_index = [[[] for _ in range(n_images)] for _ in range(n_images)]
    dst_pt_index = [[[] for _ in range(n_images)] for _ in range(n_images)]
    landmark_index = [[[] for _ in range(n_images)] for _ in range(n_images)]
    for i, j, index1, index2 in pairs:
        src_pt_index[i][j] = index1
        dst_pt_index[i][j] = index2
        landmark_index[i][j] = lm_of[i][np.asarray(index1, dtype=np.int64)].tolist() if len(index1) else []
    return src_pt_index, dst_pt_index, landmark_index, g_index, n_inconsistent


def match_graph_to_observations(n_keypoints, points, pairs):
    """Device form of steps 4-5 plus the flattening of bundle_adjustment.py:67-98.

    n_keypoints[i] keypoints per image, points[i] their [n_i, 2] pixels (or None: ids only), pairs = [(i, j, index1, index2), ...]
    in the reference's visiting order.  Returns (label_per_image: list of int arrays (-1 = unmatched), landmark_num,
    (cam_idx, lm_idx, obs_xy) or None, n_inconsistent)."""
    from . import _lib
    ctx = _lib.get_context()
    n_images = len(n_keypoints)
    off = np.concatenate([[0], np.cumsum(np.asarray(n_keypoints, dtype=np.int64))]).astype(np.int64)
    n_node = int(off[-1])
    ea = [off[i] + np.asarray(a, dtype=np.int64) for i, j, a, b in pairs]
    eb = [off[j] + np.asarray(b, dtype=np.int64) for i, j, a, b in pairs]
    edge_a = _lib.i32(np.concatenate(ea)) if ea else np.zeros(0, np.int32)
    edge_b = _lib.i32(np.concatenate(eb)) if eb else np.zeros(0, np.int32)
    n_edge = int(edge_a.shape[0])
    label = np.full(max(n_node, 1), -1, np.int32)
    n_lm = ctypes.c_int32(0)
    flat = points is not None
    node_img = node_xy = cam = lm = xy = None
    if flat:
        node_img = _lib.i32(np.repeat(np.arange(n_images), np.asarray(n_keypoints, dtype=np.int64)))
        node_xy = _lib.f64(np.concatenate([keypoints_to_matrix(p) for p in points])) if n_node else np.zeros((0, 2))
        cam, lm, xy = np.zeros(2 * n_edge, np.int32), np.zeros(2 * n_edge, np.int32), np.zeros((2 * n_edge, 2))
    ctx.check(ctx.lib.ptzba_match_graph_to_observations(ctx.handle, _lib.HOST, n_node, _lib.ptr(node_img), _lib.ptr(node_xy), n_edge,
                                                        _lib.ptr(edge_a), _lib.ptr(edge_b), _lib.ptr(label), ctypes.byref(n_lm),
                                                        _lib.ptr(cam), _lib.ptr(lm), _lib.ptr(xy)))
    label = label[:n_node]
    la, lb = label[edge_a], label[edge_b]
    # a match between two ends that were both labelled before it and carry different ids (image_process.py:624): with the final
    # labels that is exactly the matches whose ends differ (labels never change once set)
    n_inconsistent = int((la != lb).sum())
    per_image = [label[off[i]:off[i + 1]].astype(np.int64) for i in range(n_images)]
    return per_image, int(n_lm.value), ((cam, lm, xy) if flat else None), n_inconsistent


def build_matching_graph(images, image_match_mask=[], feature_method='sift', verbose=False, detect=None, match=None,
                         rng=random):
    """image_process.py:510-667 with the OpenCV calls injected:

        detect(image, feature_method)            -> (keypoints, descriptors)        stands for detect_compute_* (:538-549)
        match(kp1, des1, kp2, des2, method)      -> (pts1, index1, pts2, index2)    stands for match_*_features (:578-585)

    `rng.shuffle` thins pairs with more than 200 matches (:591-596; the reference uses the global `random` module, which is
    the default here so a seeded run reproduces it).  Returns the reference's 7-tuple
    (keypoints, descriptors, points, src_pt_index, dst_pt_index, landmark_index, landmark_num)."""
    assert feature_method == 'sift' or feature_method == 'orb' or feature_method == 'latch'
    if detect is None or match is None:
        raise NotImplementedError("feature detection / matching (OpenCV) is outside this library: pass detect= and match=")
    n = len(images)
    if len(image_match_mask) != 0:
        assert len(image_match_mask) == n
        for row in image_match_mask:
            assert len(row) == n
    elif verbose:
        print("Warning: image match mask is NOT used, may have false positive matches!")
    keypoints, descriptors = [], []
    for im in images:
        kp, des = detect(im, feature_method)
        keypoints.append(kp)
        descriptors.append(des)
    pairs = []
    for i in range(n):
        for j in range(i + 1, n):
            if len(image_match_mask) != 0 and image_match_mask[i][j] == 0:
                continue
            _, index1, _, index2 = match(keypoints[i], descriptors[i], keypoints[j], descriptors[j], feature_method)
            assert len(index1) == len(index2)
            if len(index1) > MIN_MATCH_NUM:
                if len(index1) > MAX_MATCH_NUM:
                    order = list(range(len(index1)))
                    rng.shuffle(order)
                    order = order[0:MAX_MATCH_NUM]
                    index1 = [index1[k] for k in order]
                    index2 = [index2[k] for k in order]
                pairs.append((i, j, list(index1), list(index2)))
                if verbose:
                    print("%d matches between image: %d and %d" % (len(index1), i, j))
            elif verbose:
                print("no enough matches between image: %d and %d" % (i, j))
    src_pt_index, dst_pt_index, landmark_index, landmark_num, _ = assign_landmark_index(
        n, [len(k) for k in keypoints], pairs, verbose)
    if verbose:
        print('number of landmark is %d' % landmark_num)
    points = [keypoints_to_matrix(k) for k in keypoints]
    return keypoints, descriptors, points, src_pt_index, dst_pt_index, landmark_index, landmark_num
