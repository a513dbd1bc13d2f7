"""
Synthetic PTZ workloads (no images, no datasets) for the parity tests and bench.py.

Everything here is plain numpy on the host: it only *describes* problems (keyframe poses, ray
landmarks, pixel observations, match graphs).  No projection maths of the product path lives
here except the closed form needed to place observations, which is the textbook
K * R_tilt * R_pan * d map of the reference (ptz_camera.py:191-210).

Shapes follow SURVEY.md §8(d):
  cfg1  small court-like keyframe BA + EKF (CPU-reference sized)
  cfg2  soccer cloud, 3k rays x 300 frames, single EKF sequence
  cfg3  keyframe BA 256 kf x 100k rays x 2M observations
  cfg4  4096 independent EKF sequences x 2k rays
  cfg5  BA stress 1024 kf x 1M rays x 20M observations (40 deg pan range)
"""
import math

import numpy as np

IMAGE_W = 1280
IMAGE_H = 720
PP_U = 640.0
PP_V = 360.0

_K = math.pi / 180.0


def project_closed_form(pan, tilt, f, theta, phi, u=PP_U, v=PP_V):
    """Vectorised K*R_tilt*R_pan*d with disp = 0 (all angles in degrees)."""
    p = np.radians(pan)
    t = np.radians(tilt)
    th = np.radians(theta)
    ph = np.radians(phi)
    tx = np.tan(th)
    d0 = tx
    d1 = -np.tan(ph) * np.sqrt(tx * tx + 1.0)
    d2 = 1.0
    # R_pan * d
    a0 = np.cos(p) * d0 - np.sin(p) * d2
    a1 = d1
    a2 = np.sin(p) * d0 + np.cos(p) * d2
    # R_tilt * a
    q0 = a0
    q1 = np.cos(t) * a1 + np.sin(t) * a2
    q2 = -np.sin(t) * a1 + np.cos(t) * a2
    return f * q0 / q2 + u, f * q1 / q2 + v, q2


def back_project_closed_form(pan, tilt, f, x, y, u=PP_U, v=PP_V):
    """Inverse of project_closed_form (disp = 0); returns (theta, phi) in degrees."""
    p = np.radians(pan)
    t = np.radians(tilt)
    q0 = (x - u) / f
    q1 = (y - v) / f
    q2 = np.ones_like(q0)
    # R_tilt^T
    a0 = q0
    a1 = np.cos(t) * q1 - np.sin(t) * q2
    a2 = np.sin(t) * q1 + np.cos(t) * q2
    # R_pan^T
    d0 = np.cos(p) * a0 + np.sin(p) * a2
    d1 = a1
    d2 = -np.sin(p) * a0 + np.cos(p) * a2
    theta = np.degrees(np.arctan(d0 / d2))
    phi = np.degrees(np.arctan(-d1 / np.sqrt(d0 * d0 + d2 * d2)))
    return theta, phi


def smooth_ptz_trajectory(n_frames, seed, pan_range=(49.0, 70.0), tilt_range=(-10.0, -5.9),
                          f_range=(1917.0, 4228.0)):
    """A smooth broadcast-like PTZ trajectory (sum of slow sinusoids) inside the shipped GT ranges."""
    rng = np.random.default_rng(seed)
    s = np.linspace(0.0, 1.0, n_frames)

    def chan(lo, hi):
        ph = rng.uniform(0, 2 * math.pi, 3)
        w = rng.uniform(0.4, 1.6, 3) * np.array([1.0, 2.3, 4.1])
        a = np.array([1.0, 0.35, 0.12])
        y = sum(a[i] * np.sin(2 * math.pi * w[i] * s * n_frames / 300.0 + ph[i]) for i in range(3))
        y = (y - y.min()) / max(y.max() - y.min(), 1e-12)
        return lo + (hi - lo) * (0.1 + 0.8 * y)

    return np.stack([chan(*pan_range), chan(*tilt_range), chan(*f_range)], axis=1)


def make_ray_cloud(n_rays, seed, pan_range=(35.0, 85.0), tilt_range=(-18.0, 2.0)):
    """Ray landmarks (theta, phi) in degrees scattered over the panorama the trajectory sweeps."""
    rng = np.random.default_rng(seed)
    theta = rng.uniform(pan_range[0], pan_range[1], n_rays)
    phi = rng.uniform(tilt_range[0], tilt_range[1], n_rays)
    return np.stack([theta, phi], axis=1)


# ---------------------------------------------------------------------------------------------
# EKF sequences (cfg1 / cfg2 / cfg4)
# ---------------------------------------------------------------------------------------------
class EkfSequence:
    """Per-frame observations for one image-free EKF tracking run.

    rays0       [n,2]  initial (noisy) ray estimates, degrees
    ptz_gt      [F,3]  ground-truth pan, tilt, focal
    obs_xy      list of F arrays [m_f,2]  observed pixels (noisy)
    obs_idx     list of F arrays [m_f]    ascending global ray ids of those pixels
    """

    def __init__(self, rays_gt, rays0, ptz_gt, obs_xy, obs_idx):
        self.rays_gt = rays_gt
        self.rays0 = rays0
        self.ptz_gt = ptz_gt
        self.obs_xy = obs_xy
        self.obs_idx = obs_idx


def make_ekf_sequence(n_rays, n_frames, seed, obs_noise=("uniform", 2.0), ray_init_noise=0.02,
                      keep_prob=0.9):
    rng = np.random.default_rng(seed)
    ptz = smooth_ptz_trajectory(n_frames, seed + 7)
    rays_gt = make_ray_cloud(n_rays, seed + 13,
                             pan_range=(ptz[:, 0].min() - 14, ptz[:, 0].max() + 14),
                             tilt_range=(ptz[:, 1].min() - 8, ptz[:, 1].max() + 8))
    rays0 = rays_gt + rng.normal(0.0, ray_init_noise, rays_gt.shape)
    obs_xy, obs_idx = [], []
    for k in range(n_frames):
        x, y, z = project_closed_form(ptz[k, 0], ptz[k, 1], ptz[k, 2], rays_gt[:, 0], rays_gt[:, 1])
        vis = (z > 0) & (x > 2) & (x < IMAGE_W - 2) & (y > 2) & (y < IMAGE_H - 2)
        vis &= rng.random(n_rays) < keep_prob
        idx = np.nonzero(vis)[0]
        if obs_noise[0] == "uniform":
            nz = rng.uniform(-obs_noise[1], obs_noise[1], (len(idx), 2))
        else:
            nz = rng.normal(0.0, obs_noise[1], (len(idx), 2))
        obs_xy.append(np.stack([x[idx], y[idx]], axis=1) + nz)
        obs_idx.append(idx.astype(np.int64))
    return EkfSequence(rays_gt, rays0, ptz, obs_xy, obs_idx)


# ---------------------------------------------------------------------------------------------
# Keyframe BA, reference match-graph format (small: cfg1)
# ---------------------------------------------------------------------------------------------
class MatchGraph:
    """The structure image_process.build_matching_graph hands to BA (bundle_adjustment.py:147-150).

    points          list of N arrays [N_i,2]  keypoints per keyframe
    src_pt_index    N x N list of int lists (only i<j filled)
    dst_pt_index    N x N list of int lists
    landmark_index  N x N list of int lists
    n_landmark      int
    """

    def __init__(self, points, src, dst, lmk, n_landmark, ptz_gt, rays_gt, ptz_init):
        self.points = points
        self.src_pt_index = src
        self.dst_pt_index = dst
        self.landmark_index = lmk
        self.n_landmark = n_landmark
        self.ptz_gt = ptz_gt
        self.rays_gt = rays_gt
        self.ptz_init = ptz_init


def make_match_graph(n_kf, n_landmark, seed, obs_noise=0.5, max_matches=200, pan_step=4.0,
                     init_noise=(0.5, 0.0, 30.0)):
    """Seeded keyframes on a pan sweep; pairwise matches of commonly visible landmarks.

    Mirrors what the reference feeds BA: a landmark matched in pair (i,j) contributes one
    observation in frame i and one in frame j (duplicates across pairs are kept).
    """
    rng = np.random.default_rng(seed)
    pans = 55.0 + pan_step * np.arange(n_kf) + rng.normal(0, 0.3, n_kf)
    tilts = rng.uniform(-9.0, -6.0, n_kf)
    fs = rng.uniform(2200.0, 3400.0, n_kf)
    ptz_gt = np.stack([pans, tilts, fs], axis=1)
    rays_gt = np.stack([rng.uniform(pans.min() - 10, pans.max() + 10, n_landmark),
                        rng.uniform(-13.0, -2.0, n_landmark)], axis=1)
    points = []
    local_of = []  # per frame: dict landmark -> local keypoint id
    for i in range(n_kf):
        x, y, z = project_closed_form(pans[i], tilts[i], fs[i], rays_gt[:, 0], rays_gt[:, 1])
        vis = np.nonzero((z > 0) & (x > 5) & (x < IMAGE_W - 5) & (y > 5) & (y < IMAGE_H - 5))[0]
        pts = np.stack([x[vis], y[vis]], axis=1) + rng.normal(0, obs_noise, (len(vis), 2))
        points.append(pts)
        local_of.append({int(l): k for k, l in enumerate(vis)})
    src = [[[] for _ in range(n_kf)] for _ in range(n_kf)]
    dst = [[[] for _ in range(n_kf)] for _ in range(n_kf)]
    lmk = [[[] for _ in range(n_kf)] for _ in range(n_kf)]
    for i in range(n_kf):
        for j in range(i + 1, n_kf):
            common = sorted(set(local_of[i]).intersection(local_of[j]))
            if len(common) < 4:
                continue
            if len(common) > max_matches:
                common = sorted(rng.choice(common, max_matches, replace=False).tolist())
            for l in common:
                src[i][j].append(local_of[i][l])
                dst[i][j].append(local_of[j][l])
                lmk[i][j].append(int(l))
    # landmark ids count matched landmarks only, as build_matching_graph assigns them (image_process.py:609-667)
    used = sorted({l for row in lmk for cell in row for l in cell})
    remap = {l: k for k, l in enumerate(used)}
    lmk = [[[remap[l] for l in cell] for cell in row] for row in lmk]
    rays_gt = rays_gt[used]
    ptz_init = ptz_gt.copy()
    ptz_init[1:, 0] += rng.normal(0, init_noise[0], n_kf - 1)
    ptz_init[1:, 1] += rng.normal(0, init_noise[1], n_kf - 1) if init_noise[1] > 0 else 0.0
    ptz_init[1:, 2] += rng.normal(0, init_noise[2], n_kf - 1)
    return MatchGraph(points, src, dst, lmk, len(used), ptz_gt, rays_gt, ptz_init)


# ---------------------------------------------------------------------------------------------
# Keyframe BA, flat observation format (large: cfg3 / cfg5)
# ---------------------------------------------------------------------------------------------
class FlatBA:
    """Flat landmark-major observation list (the layout the CUDA library consumes).

    cam_idx [n_obs] int32, lm_idx [n_obs] int32 (non-decreasing), obs_xy [n_obs,2] f64
    ptz_gt/ptz_init [N,3], rays_gt/rays_init [M,2]; keyframe 0 is the fixed reference pose.
    """

    def __init__(self, cam_idx, lm_idx, obs_xy, ptz_gt, rays_gt, ptz_init, rays_init):
        self.cam_idx = cam_idx
        self.lm_idx = lm_idx
        self.obs_xy = obs_xy
        self.ptz_gt = ptz_gt
        self.rays_gt = rays_gt
        self.ptz_init = ptz_init
        self.rays_init = rays_init

    @property
    def n_pose(self):
        return self.ptz_gt.shape[0]

    @property
    def n_landmark(self):
        return self.rays_gt.shape[0]

    @property
    def n_obs(self):
        return self.cam_idx.shape[0]

    def x0(self):
        """Parameter vector in the reference's layout: poses 1..N-1 then landmarks (bundle_adjustment.py:178-197)."""
        return np.concatenate([self.ptz_init[1:].ravel(), self.rays_init.ravel()])


def make_flat_ba(n_kf, n_landmark, n_obs, seed, pan_sweep=None, obs_noise=0.5,
                 init_noise=(0.5, 0.5, 30.0, 0.05), id_order="first_keyframe"):
    """Seeded flat BA problem with exactly n_obs observations, every one inside its image.

    Landmark ids follow the reference's assignment (image_process.py:612-634: the global ray index grows while the keyframes are
    walked in order, so a landmark's id is ordered by the FIRST keyframe that observes it); id_order="random" keeps the ids in
    the order the landmarks were drawn (no relation between id and position).

    Cost is O(n_obs) (candidate keyframes are drawn from the pan band that can see the landmark
    and then verified), so cfg5 (20M observations) generates in seconds, not minutes.
    """
    rng = np.random.default_rng(seed)
    if pan_sweep is None:      # cfg3 density: 256 keyframes over 120 degrees
        pan_sweep = min(120.0, n_kf * 120.0 / 256.0)
    pans = np.sort(rng.uniform(0.0, pan_sweep, n_kf)) - pan_sweep / 2.0
    tilts = rng.uniform(-12.0, -4.0, n_kf)
    fs = rng.uniform(1800.0, 4500.0, n_kf)
    ptz_gt = np.stack([pans, tilts, fs], axis=1)

    # landmarks: back-project a random pixel of a random keyframe -> guaranteed inside the union FOV
    host = rng.integers(0, n_kf, n_landmark)
    px = rng.uniform(20, IMAGE_W - 20, n_landmark)
    py = rng.uniform(20, IMAGE_H - 20, n_landmark)
    th, ph = back_project_closed_form(pans[host], tilts[host], fs[host], px, py)
    rays_gt = np.stack([th, ph], axis=1)

    # pool of candidate (keyframe, landmark) pairs: the host observation of every landmark first, then seeded
    # draws from the pan band that can see the landmark; verified in-image, de-duplicated once per round
    cam = host.astype(np.int64)
    lm = np.arange(n_landmark, dtype=np.int64)
    x, y = px, py
    for _ in range(60):
        if len(cam) >= n_obs:
            break
        want = int((n_obs - len(cam)) * 3.0) + 65536
        lm_c = rng.integers(0, n_landmark, want)
        target = rays_gt[lm_c, 0] + rng.uniform(-20.0, 20.0, want)
        cam_c = np.clip(np.searchsorted(pans, target), 0, n_kf - 1)
        xc, yc, zc = project_closed_form(pans[cam_c], tilts[cam_c], fs[cam_c], rays_gt[lm_c, 0], rays_gt[lm_c, 1])
        ok = (zc > 0) & (xc > 1) & (xc < IMAGE_W - 1) & (yc > 1) & (yc < IMAGE_H - 1)
        cam = np.concatenate([cam, cam_c[ok]]); lm = np.concatenate([lm, lm_c[ok]])
        x = np.concatenate([x, xc[ok]]); y = np.concatenate([y, yc[ok]])
        _, first = np.unique(cam * n_landmark + lm, return_index=True)
        first.sort()                                  # keep pool order (host observations stay first)
        cam, lm, x, y = cam[first], lm[first], x[first], y[first]
    if len(cam) < n_obs:
        raise RuntimeError("could not place %d observations (got %d)" % (n_obs, len(cam)))
    cam, lm, x, y = cam[:n_obs], lm[:n_obs], x[:n_obs], y[:n_obs]
    if id_order == "first_keyframe":
        first = np.full(n_landmark, n_kf, dtype=np.int64)
        np.minimum.at(first, lm, cam)
        rank = np.empty(n_landmark, dtype=np.int64)
        rank[np.lexsort((np.arange(n_landmark), first))] = np.arange(n_landmark)     # new id of every old id
        lm = rank[lm]
        inv = np.empty(n_landmark, dtype=np.int64)
        inv[rank] = np.arange(n_landmark)
        rays_gt = rays_gt[inv]
    order = np.lexsort((cam, lm))  # landmark-major, camera ascending inside a landmark
    cam, lm, x, y = cam[order], lm[order], x[order], y[order]
    obs = np.stack([x, y], axis=1) + rng.normal(0, obs_noise, (len(x), 2))

    ptz_init = ptz_gt.copy()
    ptz_init[1:, 0] += rng.normal(0, init_noise[0], n_kf - 1)
    ptz_init[1:, 1] += rng.normal(0, init_noise[1], n_kf - 1)
    ptz_init[1:, 2] += rng.normal(0, init_noise[2], n_kf - 1)
    rays_init = rays_gt + rng.normal(0, init_noise[3], rays_gt.shape)
    return FlatBA(cam.astype(np.int32), lm.astype(np.int32), np.ascontiguousarray(obs),
                  ptz_gt, rays_gt, ptz_init, rays_init)


def flatten_match_graph(points, src_pt_index, dst_pt_index, landmark_index):
    """N x N match lists -> flat (cam_idx, lm_idx, obs_xy) in the reference's residual order.

    Order and duplicates follow bundle_adjustment.py:67-98 exactly: for i, for j, for each match:
    first the observation in frame i, then the one in frame j.
    """
    n = len(points)
    cam, lm, xy = [], [], []
    for i in range(n):
        for j in range(n):
            s, d, l = src_pt_index[i][j], dst_pt_index[i][j], landmark_index[i][j]
            m = min(len(s), len(d), len(l))
            if m == 0:
                continue
            s = np.asarray(s[:m], dtype=np.int64); d = np.asarray(d[:m], dtype=np.int64)
            l = np.asarray(l[:m], dtype=np.int64)
            c2 = np.empty(2 * m, np.int64); c2[0::2] = i; c2[1::2] = j
            l2 = np.repeat(l, 2)
            p2 = np.empty((2 * m, 2)); p2[0::2] = np.asarray(points[i])[s, :2]; p2[1::2] = np.asarray(points[j])[d, :2]
            cam.append(c2); lm.append(l2); xy.append(p2)
    if not cam:
        return np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 2))
    return (np.concatenate(cam).astype(np.int32), np.concatenate(lm).astype(np.int32),
            np.ascontiguousarray(np.concatenate(xy)))
