"""Config 1 scenario (BASELINE.json configs[0], SURVEY.md 8d "cfg1"): a synthesized basketball-court sequence driven through the
whole per-frame loop - EKF tracking, ray bookkeeping, the new-keyframe rule and keyframe bundle adjustment.

The court landmarks are the reference's own point mixture (synthesized_court_sequence/synthesize_basketball.py:34-59, random.seed(1))
turned into rays by the reference's TransFunction.from_3dpoint_to_ray; tests/golden/make_golden.py:gen_cfg1 generates them WITH THE
REFERENCE and stores them in tests/golden/cfg1_court.npz, so nothing here restates that code.  This module is the seeded stand-in
for the OpenCV calls of the loop (feature detection, optical-flow matching + RANSAC, SIFT detection / matching between keyframes),
shared by the golden generator (which plugs it into the UNMODIFIED reference classes) and by tests/test_zx_cfg1_end_to_end.py
(which plugs it into the product).  Ground truth geometry comes from the oracle's projection functions.  Test infrastructure only."""
import numpy as np

from oracle import ptz_oracle as O

W, H = 1280, 720
U, V = 640.0, 360.0
CC = np.array([13.0099, -14.8109, 6.1790])                 # generator/image_generator.py:44,117
BASE_ROT = np.array([1.5804, -0.1186, 0.1249])


class _Kp:
    """Stand-in for cv2.KeyPoint: the keyframe matcher only reads .pt (image_process.py:653-661)."""

    def __init__(self, x, y):
        self.pt = (float(x), float(y))


def court_trajectory(n_frames, seed):
    """Smooth seeded pan / tilt / focal trajectory over the court: a steady pan sweep (0.3 degrees per frame, the order of
    the shipped sequences' 0.12 degrees mean), slow tilt and zoom drifts, plus a little per-frame jitter."""
    rng = np.random.default_rng(seed)
    k = np.arange(n_frames)
    s = k / max(n_frames - 1, 1)
    pan = -30.0 + 0.3 * k + 0.6 * np.sin(2 * np.pi * 1.5 * s)
    tilt = -9.0 + 1.5 * s + 0.2 * np.sin(2 * np.pi * 2.0 * s)
    f = 2600.0 + 350.0 * s + 30.0 * np.sin(2 * np.pi * 1.0 * s)
    gt = np.stack([pan, tilt, f], 1)
    gt[1:] += rng.normal(0, 1, (n_frames - 1, 3)) * [0.01, 0.005, 0.5]
    return gt


def relocalization_box():
    """The fixed mask relocalization.py:103-104 / :50-51 applies to detected keypoints (the broadcast's score banner)."""
    box = np.ones([H, W])
    box[13:51, 303:976] = 0
    return box


class CourtSequence:
    def __init__(self, court_rays, n_frames, seed, obs_noise=0.5, blackout=None):
        """`blackout = (a, b)`: the optical flow loses 70 % of its matches on frames a <= k < b, which walks bad_tracking_cnt to
        tracking_lost (ptz_slam.py:404-411) and sends the caller through relocalize + init_system."""
        self.court_rays = np.asarray(court_rays, dtype=np.float64)
        self.n_frames, self.seed, self.obs_noise = int(n_frames), int(seed), float(obs_noise)
        self.blackout = (self.n_frames, self.n_frames) if blackout is None else (int(blackout[0]), int(blackout[1]))
        self.gt = court_trajectory(self.n_frames, self.seed)
        self.bounding_box = np.ones((H, W), np.uint8)
        self.bounding_box[300:520, 500:640] = 0                         # a "player"
        self._kf_cache = {}

    # -- images carry nothing but their frame number ----------------------------------------------------------------------
    def image(self, k):
        img = np.zeros((H, W, 3), np.uint8)
        img[0, 0, 0], img[0, 0, 1] = k % 256, k // 256
        return img

    @staticmethod
    def frame_of(img):
        return int(img[0, 0, 0]) + 256 * int(img[0, 0, 1])

    def _detect(self, k):
        """Court landmarks visible in frame k at their true pixel + N(0, obs_noise) px: (landmark ids, pixels)."""
        g = self.gt[k]
        x, y, z = O.project_rays_vec(g[0], g[1], g[2], U, V, self.court_rays)
        vis = np.nonzero((z > 0) & (x > 5) & (x < W - 5) & (y > 5) & (y < H - 5))[0]
        rng = np.random.default_rng(self.seed * 1000 + k)
        pts = np.stack([x[vis], y[vis]], 1) + rng.normal(0, self.obs_noise, (len(vis), 2))
        return vis, pts

    # -- the two OpenCV calls of PtzSlam.init_system / tracking / add_rays ---------------------------------------------------
    def detect_keypoints(self, img, n):
        """detect_compute_sift_array (ptz_slam.py:151, :338): keypoints [m,2] and descriptors [m,d] of the frame."""
        ids, pts = self._detect(self.frame_of(img))
        ids, pts = ids[:n], pts[:n]
        des = np.zeros((len(ids), 8), np.float32)
        des[:, 0] = ids
        return pts, des

    def matching_and_ransac(self, img1, img2, kp1, kp1_index):
        """Where the tracked keypoints of frame k1 really are in frame k2 (+0.3 px); a few are lost by the flow, a few are flagged
        as RANSAC outliers; same return convention as image_process.matching_and_ransac (:464-506)."""
        k1, k2 = self.frame_of(img1), self.frame_of(img2)
        rng = np.random.default_rng(self.seed * 1000003 + 1000 * k1 + k2)
        kp1 = np.asarray(kp1, dtype=np.float64).reshape(-1, 2)
        kp1_index = np.asarray(kp1_index)
        g1, g2 = self.gt[k1], self.gt[k2]
        rays = O.back_project_to_rays_vec(g1[0], g1[1], g1[2], U, V, kp1)
        x, y, _ = O.project_rays_vec(g2[0], g2[1], g2[2], U, V, rays)
        cur = np.stack([x, y], 1) + rng.normal(0, 0.3, (len(kp1), 2))
        inside = (cur[:, 0] > 1) & (cur[:, 0] < W - 1) & (cur[:, 1] > 1) & (cur[:, 1] < H - 1)
        flow_ok = inside & (rng.uniform(size=len(kp1)) > (0.7 if self.blackout[0] <= k2 < self.blackout[1] else 0.03))
        local = np.nonzero(flow_ok)[0]
        ransac_in = rng.uniform(size=len(local)) > 0.04
        return cur[local][ransac_in], kp1_index[local][ransac_in], kp1_index[local][~ransac_in]

    # -- the two OpenCV calls of image_process.build_matching_graph (keyframe bundle adjustment) ---------------------------
    def detect_sift(self, img, *_):
        """detect_compute_sift (image_process.py:538-549): (list of KeyPoint-like objects, descriptors); descriptor = landmark id."""
        k = self.frame_of(img)
        if k not in self._kf_cache:
            ids, pts = self._detect(k)
            self._kf_cache[k] = ([_Kp(x, y) for x, y in pts], ids.astype(np.float32).reshape(-1, 1))
        return self._kf_cache[k]

    @staticmethod
    def match_sift(kp1, des1, kp2, des2, *_):
        """match_sift_features (image_process.py:578-585): (pts1, index1, pts2, index2) of the landmarks both frames see."""
        id1, id2 = np.asarray(des1)[:, 0].astype(np.int64), np.asarray(des2)[:, 0].astype(np.int64)
        common, i1, i2 = np.intersect1d(id1, id2, return_indices=True)
        return None, i1.tolist(), None, i2.tolist()

    # -- the two OpenCV calls of relocalization.relocalization_camera -------------------------------------------------------
    def detect_array(self, img, n, *_, **__):
        """detect_compute_sift_array(img, n, norm=False) (relocalization.py:101, :126, :54-55): keypoints [m,2] and descriptors."""
        ids, pts = self._detect(self.frame_of(img))
        ids, pts = ids[:n], pts[:n]
        return pts, ids.astype(np.float32).reshape(-1, 1)

    def detect(self, img, n):
        """The product's hook: detection + the relocaliser's fixed banner mask (relocalization.py:103-108)."""
        from ptz_slam_b200.ptz_slam import keypoints_masking
        pts, des = self.detect_array(img, n)
        keep = keypoints_masking(pts, relocalization_box())
        return pts[keep], des[keep]

    @staticmethod
    def match(kp1, des1, kp2, des2, *_, **__):
        """match_sift_features(..., pts_array=True): (pt1, index1, pt2, index2) with matched pixel arrays."""
        id1, id2 = np.asarray(des1)[:, 0].astype(np.int64), np.asarray(des2)[:, 0].astype(np.int64)
        common, i1, i2 = np.intersect1d(id1, id2, return_indices=True)
        if len(common) == 0:
            return None, None, None, None
        return np.asarray(kp1)[i1], i1.tolist(), np.asarray(kp2)[i2], i2.tolist()

    def build_matching_graph(self, images, image_match_mask, feature_method, verbose):
        """The product's front-end hook (scene_map.Map / bundle_adjustment): the reference's build_matching_graph with the
        detector and the matcher above."""
        from ptz_slam_b200 import match_graph
        return match_graph.build_matching_graph(images, image_match_mask, feature_method, verbose,
                                                detect=lambda im, method: self.detect_sift(im), match=self.match_sift)
