"""Seeded stand-in for the OpenCV front end of the per-frame loop (SIFT detection, optical-flow matching + homography
RANSAC), shared by tests/golden/make_golden.py:gen_tracking (which plugs it into the UNMODIFIED reference PtzSlam) and by
tests/test_tracking_loop.py (which plugs it into the product).  Ground truth comes from the oracle's projection functions.
Test infrastructure only."""
import numpy as np

from oracle import ptz_oracle as O

W, H = 1280, 720
U, V = 640.0, 360.0


class SyntheticFrontEnd:
    def __init__(self, seed, n_frames, bad_from=None):
        rng = np.random.default_rng(seed)
        self.seed = seed
        self.n_frames = n_frames
        self.bad_from = n_frames if bad_from is None else bad_from      # frames >= bad_from lose most optical-flow matches
        ptz0 = np.array([rng.uniform(50, 60), rng.uniform(-10, -8), rng.uniform(2800, 3400)])
        step = np.array([1.2 if bad_from is None else 0.35, -0.05, 12.0])     # the fast pan reaches the new-keyframe band
        self.gt = np.array([ptz0 + k * step + (rng.normal(0, 1, 3) * [0.02, 0.01, 1.0] if k else 0) for k in range(n_frames)])
        self.bounding_box = np.ones((H, W), np.uint8)
        self.bounding_box[300:520, 500:640] = 0                         # a "player"

    def image(self, k):
        img = np.zeros((H, W, 3), np.uint8)
        img[0, 0, 0] = k
        return img

    @staticmethod
    def frame_of(img):
        return int(img[0, 0, 0])

    def detect_keypoints(self, img, n):
        k = self.frame_of(img)
        rng = np.random.default_rng(self.seed * 1000 + k)
        m = 70 if k == 0 else 40
        pts = np.stack([rng.uniform(5, W - 5, m), rng.uniform(5, H - 5, m)], 1).astype(np.float32)
        des = rng.integers(0, 255, (m, 8)).astype(np.float32)
        return pts, des

    def matching_and_ransac(self, img1, img2, kp1, kp1_index):
        """Where the tracked keypoints of frame k1 really are in frame k2 (+0.3 px noise); some are lost by the flow,
        some are flagged as RANSAC outliers; same return convention as image_process.matching_and_ransac (:464-506)."""
        k1, k2 = self.frame_of(img1), self.frame_of(img2)
        rng = np.random.default_rng(self.seed * 1000003 + 1000 * k1 + k2)
        kp1 = np.asarray(kp1, dtype=np.float64).reshape(-1, 2)
        kp1_index = np.asarray(kp1_index)
        g1, g2 = self.gt[k1], self.gt[k2]
        rays = O.back_project_to_rays_vec(g1[0], g1[1], g1[2], U, V, kp1)
        x, y, _ = O.project_rays_vec(g2[0], g2[1], g2[2], U, V, rays)
        cur = np.stack([x, y], 1) + rng.normal(0, 0.3, (len(kp1), 2))
        inside = (cur[:, 0] > 1) & (cur[:, 0] < W - 1) & (cur[:, 1] > 1) & (cur[:, 1] < H - 1)
        flow_ok = inside & (rng.uniform(size=len(kp1)) > (0.6 if k2 >= self.bad_from else 0.04))
        local = np.nonzero(flow_ok)[0]
        ransac_in = rng.uniform(size=len(local)) > 0.06
        inlier_keypoints = cur[local][ransac_in]
        inlier_index = kp1_index[local][ransac_in]
        outlier_index = kp1_index[local][~ransac_in]
        return inlier_keypoints, inlier_index, outlier_index
