"""
Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/slam_system) on seeded inputs.
Run in the authoring container only:  python tests/golden/make_golden.py
The .npz files are committed; the GPU box never needs /root/reference.

Goldens (reference function -> file):
  PTZCamera.project_ray / project_rays, TransFunction.from_ray_to_image           -> projection.npz
  PTZCamera.back_project_to_ray(s), TransFunction.from_image_to_ray               -> backprojection.npz
  PtzSlam.compute_h_jacobian                                                      -> h_jacobian.npz
  PtzSlam.ekf_update + predict lines ptz_slam.py:418-426 (6 frames)               -> ekf.npz
  the same at the config 2 size (3 000 rays, 10 frames)                           -> cfg2_reference.npz
  bundle_adjustment._compute_residual                                             -> ba_residual.npz
  scipy least_squares call of bundle_adjustment.py:200-202 (as-is and tight)      -> ba_solve.npz
  util.overlap_pan_angle, scene_map.Map.good_new_keyframe                         -> keyframe_map.npz
  PtzSlam.remove_rays / add_rays (SIFT detector replaced by seeded keypoints),
  image_process.keypoints_masking                                                 -> ray_bookkeeping.npz
  image_process.build_matching_graph steps 4-5 (detector / matcher replaced)      -> match_graph.npz
  KeyFrame.convert_keypoint_to_array / save_to_mat                                -> keyframe_mat.npz
  RandomForestMap.bundle_adjustment_processing (BA call replaced by a recorder)   -> sliding_window.npz
  relocalization._compute_residual + its least_squares call (as-is and tight)     -> relocalization.npz
  PtzSlam.init_system + tracking over a sequence (OpenCV calls replaced)          -> tracking.npz
  config 1 end to end: 150-frame court sequence, tracking + add_keyframe ->
  Map.add_keyframe_with_ba -> bundle_adjustment -> least_squares (OpenCV replaced) -> cfg1_court.npz
  the same loop with a flow blackout: tracking_lost -> relocalize -> relocalization_camera
  -> init_system                                                                  -> cfg1_relocalize.npz
  util.add_gauss / add_outliers / uniform_point_sample_on_field / compute_error_data  -> util_noise.npz
  PTZCamera matrices, project_3d_point(s), back_project_to_3d_point(s)            -> camera_3d.npz
"""
import copy
import io
import os
import sys
import contextlib
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
import ptz_slam_b200  # noqa: E402,F401
from ptz_slam_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ref = ref_import.load()
U, V, W, H = synth.PP_U, synth.PP_V, synth.IMAGE_W, synth.IMAGE_H
CC = np.array([13.0099, -14.8109, 6.1790])
BASE_ROT = np.array([1.5804, -0.1186, 0.1249])
DISP = np.array([0.012, -0.008, 0.05, 1.5e-6, -2.0e-6, 4.0e-6])   # a non-trivial displacement case


def make_camera(ptz, disp=None):
    cam = ref.PTZCamera((U, V), CC, BASE_ROT, None if disp is None else np.array(disp))
    cam.set_ptz(ptz)
    return cam


def gen_projection():
    rng = np.random.default_rng(11)
    ptzs = np.array([[12.3, -7.5, 2500.0], [58.0, -9.0, 3500.0], [-33.0, -4.2, 1900.0], [71.5, -12.0, 4200.0]])
    rays = np.stack([rng.uniform(-60, 100, 96), rng.uniform(-25, 8, 96)], axis=1)
    out = {"ptzs": ptzs, "rays": rays, "disp": DISP, "uv": np.array([U, V])}
    for tag, disp in (("nodisp", None), ("disp", DISP)):
        xy = np.zeros((len(ptzs), len(rays), 2))
        for c, ptz in enumerate(ptzs):
            cam = make_camera(ptz, disp)
            for r, ray in enumerate(rays):
                xy[c, r] = cam.project_ray(ray)
        out["project_ray_" + tag] = xy
    t = np.zeros((len(ptzs), len(rays), 2))
    for c, ptz in enumerate(ptzs):
        for r, ray in enumerate(rays):
            t[c, r] = ref.TransFunction.from_ray_to_image(U, V, ptz[2], ptz[0], ptz[1], ray[0], ray[1])
    out["from_ray_to_image"] = t
    # project_rays with the strict in-image filter, and without
    for c, ptz in enumerate(ptzs):
        cam = make_camera(ptz)
        near = np.stack([ptz[0] + rng.uniform(-25, 25, 200), ptz[1] + rng.uniform(-14, 14, 200)], axis=1)
        pts, idx = cam.project_rays(near, H, W)
        out["prs_rays_%d" % c] = near
        out["prs_points_%d" % c] = pts.astype(np.float64)
        out["prs_index_%d" % c] = idx
        pts2, idx2 = cam.project_rays(near[:17])
        out["prs_all_points_%d" % c] = pts2.astype(np.float64)
        assert len(idx2) == 0
    np.savez(os.path.join(OUT, "projection.npz"), **out)


def gen_backprojection():
    rng = np.random.default_rng(12)
    ptzs = np.array([[12.3, -7.5, 2500.0], [58.0, -9.0, 3500.0], [-33.0, -4.2, 1900.0]])
    pts = np.stack([rng.uniform(-100, W + 100, 80), rng.uniform(-100, H + 100, 80)], axis=1)
    out = {"ptzs": ptzs, "points": pts, "disp": DISP, "uv": np.array([U, V])}
    for tag, disp in (("nodisp", None), ("disp", DISP)):
        r = np.zeros((len(ptzs), len(pts), 2))
        for c, ptz in enumerate(ptzs):
            cam = make_camera(ptz, disp)
            r[c] = cam.back_project_to_rays(pts)
        out["back_project_" + tag] = r
    t = np.zeros((len(ptzs), len(pts), 2))
    for c, ptz in enumerate(ptzs):
        for k, p in enumerate(pts):
            t[c, k] = ref.TransFunction.from_image_to_ray(U, V, ptz[2], ptz[0], ptz[1], p[0], p[1])
    out["from_image_to_ray"] = t
    np.savez(os.path.join(OUT, "backprojection.npz"), **out)


def gen_h_jacobian():
    rng = np.random.default_rng(13)
    out = {"uv": np.array([U, V]), "disp": DISP}
    for tag, disp in (("nodisp", None), ("disp", DISP)):
        slam = ref.PtzSlam()
        ptz = np.array([10.0, -8.0, 2500.0])
        slam.cameras = [make_camera(ptz, disp)]
        rays = np.stack([ptz[0] + rng.uniform(-12, 12, 9), ptz[1] + rng.uniform(-6, 6, 9)], axis=1)
        Hm = slam.compute_h_jacobian(ptz[0], ptz[1], ptz[2], rays)
        out["ptz_" + tag] = ptz
        out["rays_" + tag] = rays
        out["H_" + tag] = Hm
    np.savez(os.path.join(OUT, "h_jacobian.npz"), **out)


def gen_ekf():
    seq = synth.make_ekf_sequence(n_rays=60, n_frames=7, seed=1001, obs_noise=("gauss", 0.5))
    slam = ref.PtzSlam()
    cam0 = make_camera(seq.ptz_gt[0])
    slam.cameras = [cam0]
    slam.rays = seq.rays0.copy()
    slam.state_cov = slam.angle_var * np.eye(3 + 2 * len(slam.rays))   # ptz_slam.py:199-200
    slam.state_cov[2][2] = slam.f_var
    out = {"rays0": seq.rays0, "ptz0": seq.ptz_gt[0], "uv": np.array([U, V]), "n_frames": np.array(6)}
    for k in range(1, 7):
        # predict, ptz_slam.py:418-426
        slam.current_camera = copy.deepcopy(slam.cameras[-1])
        slam.current_camera.set_ptz(slam.current_camera.get_ptz() + slam.velocity)
        slam.cameras.append(slam.current_camera)
        q_k = 5 * np.diag([slam.angle_var, slam.angle_var, slam.f_var])
        slam.state_cov[0:3, 0:3] = slam.state_cov[0:3, 0:3] + q_k
        slam.ekf_update(seq.obs_xy[k], seq.obs_idx[k], H, W)
        out["obs_xy_%d" % k] = seq.obs_xy[k]
        out["obs_idx_%d" % k] = seq.obs_idx[k]
        out["rays_%d" % k] = slam.rays.copy()
        out["cov_%d" % k] = slam.state_cov.copy()
        out["ptz_%d" % k] = slam.current_camera.get_ptz()
        out["vel_%d" % k] = np.array(slam.velocity)
    np.savez_compressed(os.path.join(OUT, "ekf.npz"), **out)


def gen_cfg2_reference():
    """Config 2 shape (3 000 rays, every visible ray observed): 10 consecutive frames of PtzSlam.ekf_update + the predict lines
    ptz_slam.py:418-426 by the UNMODIFIED reference (2.8 s per frame here) on the sequence tests/test_gpu_ekf.py uses for its
    50-frame comparison against the oracle - the direct pin of that oracle (and of the device) at this size."""
    n_rays, n_frames = 3000, 10
    seq = synth.make_ekf_sequence(n_rays, n_frames + 1, seed=1002, keep_prob=1.0)
    slam = ref.PtzSlam()
    slam.cameras = [make_camera(seq.ptz_gt[0])]
    slam.rays = seq.rays0.copy()
    slam.state_cov = slam.angle_var * np.eye(3 + 2 * len(slam.rays))   # ptz_slam.py:199-200
    slam.state_cov[2][2] = slam.f_var
    out = {"n_rays": np.array(n_rays), "n_frames": np.array(n_frames), "seed": np.array(1002)}
    for k in range(1, n_frames + 1):
        slam.current_camera = copy.deepcopy(slam.cameras[-1])          # predict, ptz_slam.py:418-426
        slam.current_camera.set_ptz(slam.current_camera.get_ptz() + slam.velocity)
        slam.cameras.append(slam.current_camera)
        slam.state_cov[0:3, 0:3] = slam.state_cov[0:3, 0:3] + 5 * np.diag([slam.angle_var, slam.angle_var, slam.f_var])
        slam.ekf_update(seq.obs_xy[k], seq.obs_idx[k], H, W)
        out["ptz_%d" % k] = slam.current_camera.get_ptz()
        out["vel_%d" % k] = np.array(slam.velocity)
        if k in (5, n_frames):
            out["rays_%d" % k] = slam.rays.copy()
            out["cov_diag_%d" % k] = np.diag(slam.state_cov).copy()
            out["cov_probe_%d" % k] = slam.state_cov @ cov_probe_vector(slam.state_cov.shape[0])
        print("cfg2 reference frame %d: ptz error vs ground truth %s" % (k, out["ptz_%d" % k] - seq.ptz_gt[k]), flush=True)
    np.savez_compressed(os.path.join(OUT, "cfg2_reference.npz"), **out)


def _graph_to_npz(g, out):
    n = len(g.points)
    out["n_kf"] = np.array(n)
    out["n_landmark"] = np.array(g.n_landmark)
    out["ptz_init"] = g.ptz_init
    out["ptz_gt"] = g.ptz_gt
    out["rays_gt"] = g.rays_gt
    for i in range(n):
        out["points_%d" % i] = g.points[i]
    pairs = []
    for i in range(n):
        for j in range(n):
            if len(g.src_pt_index[i][j]):
                pairs.append((i, j))
                out["src_%d_%d" % (i, j)] = np.array(g.src_pt_index[i][j], dtype=np.int64)
                out["dst_%d_%d" % (i, j)] = np.array(g.dst_pt_index[i][j], dtype=np.int64)
                out["lmk_%d_%d" % (i, j)] = np.array(g.landmark_index[i][j], dtype=np.int64)
    out["pairs"] = np.array(pairs, dtype=np.int64)


def _x0_reference(g):
    """bundle_adjustment.py:168-197 executed verbatim on the graph (the reference's own lines cannot be called
    without images, so the landmark initialisation loop is driven here through TransFunction.from_image_to_ray)."""
    N = len(g.points)
    n_residual = sum(len(g.src_pt_index[i][j]) * 4 for i in range(N) for j in range(N))
    x0 = np.zeros(N * 3 + g.n_landmark * 2)
    for i in range(N):
        x0[3 * i:3 * i + 3] = g.ptz_init[i]
    ls = N * 3
    for i in range(N):
        ptz1 = x0[3 * i:3 * i + 3]
        for j in range(N):
            for idx1, idx2, idx3 in zip(g.src_pt_index[i][j], g.dst_pt_index[i][j], g.landmark_index[i][j]):
                x1, y1 = g.points[i][idx1][0], g.points[i][idx1][1]
                x0[ls + 2 * idx3:ls + 2 * idx3 + 2] = ref.TransFunction.from_image_to_ray(U, V, ptz1[2], ptz1[0], ptz1[1], x1, y1)
    return x0[3:], n_residual


def gen_ba():
    g = synth.make_match_graph(n_kf=5, n_landmark=48, seed=1001, max_matches=30)
    N = len(g.points)
    x0, n_residual = _x0_reference(g)
    args = (N, g.n_landmark, n_residual, g.points, g.src_pt_index, g.dst_pt_index, g.landmark_index, U, V, g.ptz_init[0])
    out = {"uv": np.array([U, V])}
    _graph_to_npz(g, out)
    out["x0"] = x0
    out["n_residual"] = np.array(n_residual)
    out["residual_x0"] = ref.bundle_adjustment._compute_residual(x0, *args)
    rng = np.random.default_rng(5)
    x1 = x0 + rng.normal(0, 0.01, x0.shape)
    out["x1"] = x1
    out["residual_x1"] = ref.bundle_adjustment._compute_residual(x1, *args)
    np.savez_compressed(os.path.join(OUT, "ba_residual.npz"), **out)

    # the exact call of bundle_adjustment.py:200-202 (verbose silenced), and a tight-tolerance run that defines "converged"
    with contextlib.redirect_stdout(io.StringIO()):
        asis = ref.least_squares(ref.bundle_adjustment._compute_residual, x0, verbose=2, x_scale='jac', ftol=1e-4,
                                 method='trf', args=args)
        tight = ref.least_squares(ref.bundle_adjustment._compute_residual, x0, verbose=0, x_scale='jac',
                                  ftol=1e-15, xtol=1e-15, gtol=1e-15, method='trf', args=args, max_nfev=60)
    print("BA as-is: nfev %d status %d cost %.9g ; tight: nfev %d status %d cost %.12g" %
          (asis.nfev, asis.status, asis.cost, tight.nfev, tight.status, tight.cost))
    print("max |asis - tight| poses:", np.abs(asis.x[:3 * (N - 1)] - tight.x[:3 * (N - 1)]).reshape(-1, 3).max(0),
          " rays:", np.abs(asis.x[3 * (N - 1):] - tight.x[3 * (N - 1):]).max())
    out2 = {"uv": np.array([U, V])}
    _graph_to_npz(g, out2)
    out2.update(x0=x0, x_asis=asis.x, cost_asis=np.array(asis.cost), nfev_asis=np.array(asis.nfev),
                status_asis=np.array(asis.status), x_tight=tight.x, cost_tight=np.array(tight.cost))
    np.savez_compressed(os.path.join(OUT, "ba_solve.npz"), **out2)


def gen_keyframe_map():
    """util.overlap_pan_angle (util.py:49-72) and Map.good_new_keyframe (scene_map.py:119-149) on seeded poses."""
    rng = np.random.default_rng(1011)
    import util as ref_util
    from scene_map import Map as RefMap
    from key_frame import KeyFrame as RefKeyFrame
    n = 200
    fl1, fl2 = rng.uniform(1500, 4500, n), rng.uniform(1500, 4500, n)
    p1 = rng.uniform(30, 90, n)
    p2 = p1 + rng.uniform(-45, 45, n)
    ov = np.array([ref_util.overlap_pan_angle(fl1[i], p1[i], fl2[i], p2[i], W) for i in range(n)])
    kf_ptz = np.stack([np.array([40.0, 52.0, 61.0, 75.0]), rng.uniform(-10, -6, 4), rng.uniform(2000, 4000, 4)], 1)
    m = RefMap('sift')
    for i, q in enumerate(kf_ptz):
        kf = RefKeyFrame(None, i, CC, BASE_ROT, U, V, q[0], q[1], q[2])
        m.keyframe_list.append(kf)
    cand = np.stack([rng.uniform(30, 95, 300), rng.uniform(-10, -6, 300), rng.uniform(1800, 4300, 300)], 1)
    with contextlib.redirect_stdout(io.StringIO()):
        good = np.array([m.good_new_keyframe(c) for c in cand])
        good_custom = np.array([m.good_new_keyframe(c, 10, 15, W) for c in cand])     # ptz_slam.py:458 thresholds
    # util.get_overlap_index (util.py:75-96): two-pointer merge of ascending index arrays (ragged / empty / disjoint cases)
    ov_cases = {}
    shapes = [(40, 25), (0, 10), (10, 0), (7, 7), (300, 280)]
    for c, (n1, n2) in enumerate(shapes):
        a = np.sort(rng.choice(400, n1, replace=False)) if n1 else np.zeros(0, np.int64)
        b = np.sort(rng.choice(400, n2, replace=False)) if n2 else np.zeros(0, np.int64)
        if c == 3:
            b = a + 1000                     # nothing shared
        i1, i2 = ref_util.get_overlap_index(a, b)
        ov_cases["ov%d_a" % c], ov_cases["ov%d_b" % c] = a, b
        ov_cases["ov%d_i1" % c], ov_cases["ov%d_i2" % c] = i1.astype(np.int64), i2.astype(np.int64)
    np.savez(os.path.join(OUT, "keyframe_map.npz"), fl1=fl1, fl2=fl2, p1=p1, p2=p2, overlap=ov, kf_ptz=kf_ptz, cand=cand,
             good=good, good_custom=good_custom, im_width=W, n_overlap_cases=len(shapes), **ov_cases)


def gen_ray_bookkeeping():
    """PtzSlam.remove_rays (ptz_slam.py:291-315) and add_rays (:317-388) run on a seeded state.  The only line replaced is
    the OpenCV SIFT detector (:338, detect_compute_sift_array): it is monkeypatched in the reference module's namespace to
    return seeded keypoints + descriptors, so every bookkeeping line of the reference executes unmodified.
    numpy >= 1.19 refuses the float index array remove_rays builds (:309-315) with np.delete; the reference module's
    np.delete is wrapped for that call so the indices are cast to integers - values unchanged."""
    import ptz_slam as ref_ptz_slam
    import image_process as ref_ip
    rng = np.random.default_rng(1213)
    out = {"uv": np.array([U, V]), "hw": np.array([H, W])}
    n_cases, DES = 4, 16                          # the bookkeeping is descriptor-width agnostic; 16 keeps the fixture small
    for c in range(n_cases):
        ptz = np.array([rng.uniform(45, 70), rng.uniform(-12, -7), rng.uniform(2200, 3800)])
        cam = make_camera(ptz, DISP if c == 2 else None)
        n0 = [16, 0, 10, 24][c]
        # rays spread wider than the field of view so that some project outside the image
        fov = np.degrees(np.arctan(W / 2 / ptz[2]))
        rays0 = np.stack([ptz[0] + rng.uniform(-1.6 * fov, 1.6 * fov, n0), ptz[1] + rng.uniform(-1.2 * fov, 1.2 * fov, n0)], 1)
        slam = ref.PtzSlam()
        slam.current_camera = cam
        slam.cameras = [cam]
        slam.rays = rays0.copy().reshape(-1, 2)
        slam.des = rng.uniform(0, 255, (n0, DES)).astype(np.float32)
        a = rng.normal(size=(3 + 2 * n0, 3 + 2 * n0))
        slam.state_cov = a @ a.T * 1e-4                     # dense SPD so that every deleted / kept entry is distinguishable
        out["c%d_ptz" % c], out["c%d_disp" % c] = ptz, (DISP if c == 2 else np.zeros(6))
        out["c%d_rays0" % c], out["c%d_des0" % c], out["c%d_cov0" % c] = slam.rays.copy(), slam.des.copy(), slam.state_cov.copy()
        # --- add_rays
        n_new = [48, 12, 36, 0][c]
        new_kp = np.stack([rng.uniform(0, W - 1e-3, n_new), rng.uniform(0, H - 1e-3, n_new)], 1).astype(np.float32)
        new_des = rng.uniform(0, 255, (n_new, DES)).astype(np.float32)
        bbox = np.ones((H, W), np.uint8)
        if c != 1:
            for _ in range(3):
                x0, y0 = int(rng.uniform(0, W - 200)), int(rng.uniform(0, H - 250))
                bbox[y0:y0 + 250, x0:x0 + 200] = 0
        bbox_arg = None if c == 1 else bbox
        saved = ref_ptz_slam.detect_compute_sift_array
        ref_ptz_slam.detect_compute_sift_array = lambda img, num, kp=new_kp, de=new_des: (kp.copy(), de.copy())
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                kp, kp_idx = slam.add_rays(np.zeros((H, W, 3), np.uint8), bbox_arg)
        finally:
            ref_ptz_slam.detect_compute_sift_array = saved
        out["c%d_new_kp" % c], out["c%d_new_des" % c] = new_kp, new_des
        out["c%d_bbox" % c] = np.packbits(bbox)
        out["c%d_has_bbox" % c] = np.array(bbox_arg is not None)
        out["c%d_add_kp" % c], out["c%d_add_idx" % c] = np.asarray(kp, np.float64), np.asarray(kp_idx, np.float64)
        out["c%d_rays1" % c], out["c%d_des1" % c], out["c%d_cov1" % c] = slam.rays.copy(), slam.des.copy(), slam.state_cov.copy()
        # --- keypoints_masking on its own (image_process.py:158-175)
        out["c%d_mask_idx" % c] = np.asarray(ref_ip.keypoints_masking(new_kp, bbox), np.int64)
        # --- remove_rays on the grown state
        n1 = len(slam.rays)
        n_del = [5, 3, 0, 9][c]
        del_idx = rng.choice(n1, min(n_del, n1), replace=False).astype(np.int64)
        np_delete = np.delete

        class _Np:                                  # np proxy: integer-casts the float index array of ptz_slam.py:309-315
            def __getattr__(self, k):
                return getattr(np, k)

            @staticmethod
            def delete(arr, obj, axis=None):
                return np_delete(arr, np.asarray(obj).astype(np.int64), axis=axis)
        ref_ptz_slam.np = _Np()
        try:
            slam.remove_rays(del_idx)
        finally:
            ref_ptz_slam.np = np
        out["c%d_del" % c] = del_idx
        out["c%d_rays2" % c], out["c%d_des2" % c], out["c%d_cov2" % c] = slam.rays.copy(), slam.des.copy(), slam.state_cov.copy()
        print("ray_bookkeeping case %d: %d rays -> +%d -> -%d -> %d" % (c, n0, n1 - n0, len(del_idx), len(slam.rays)))
    out["n_cases"] = np.array(n_cases)
    np.savez_compressed(os.path.join(OUT, "ray_bookkeeping.npz"), **out)


class _Kp:
    """Stand-in for cv2.KeyPoint: build_matching_graph only reads .pt (image_process.py:653-661)."""

    def __init__(self, x, y):
        self.pt = (float(x), float(y))


def synth_match_inputs(seed, n_img=6, n_gt=420):
    """Seeded front-end output: keypoints per image and raw pair-wise matches (with dropped, wrong and oversized sets)."""
    rng = np.random.default_rng(seed)
    see_p = [0.97, 0.96, 0.5, 0.45, 0.5, 0.08][:n_img]
    kps, gt_of = [], []
    for i in range(n_img):
        vis = np.nonzero(rng.uniform(size=n_gt) < see_p[i])[0]
        rng.shuffle(vis)
        n_extra = 15                                                 # keypoints that match nothing
        gt = np.concatenate([vis, -np.ones(n_extra, np.int64)])
        rng.shuffle(gt)
        kps.append(np.stack([rng.uniform(1, W - 1, len(gt)), rng.uniform(1, H - 1, len(gt))], 1))
        gt_of.append(gt)
    raw = {}
    for i in range(n_img):
        for j in range(i + 1, n_img):
            common = np.intersect1d(gt_of[i][gt_of[i] >= 0], gt_of[j][gt_of[j] >= 0])
            common = common[rng.uniform(size=len(common)) > (0.05 if (i, j) == (0, 1) else 0.3)]
            rng.shuffle(common)
            pos_i = {g: k for k, g in enumerate(gt_of[i])}
            pos_j = {g: k for k, g in enumerate(gt_of[j])}
            i1 = [pos_i[g] for g in common]
            i2 = [pos_j[g] for g in common]
            for _ in range(4):                                       # wrong matches -> "in-consistent matching" branch
                i1.append(int(rng.integers(len(gt_of[i]))))
                i2.append(int(rng.integers(len(gt_of[j]))))
            raw[(i, j)] = (np.array(i1, np.int64), np.array(i2, np.int64))
    mask = np.ones((n_img, n_img), np.int64)
    mask[1, 3] = mask[3, 1] = 0
    return kps, raw, mask


def gen_match_graph():
    """image_process.build_matching_graph (:510-667) with detect_compute_sift / match_sift_features monkeypatched to a
    seeded front-end; the mask test, the 20 / 200 match thresholds, random.shuffle thinning and steps 4-5 run unmodified."""
    import random
    import image_process as ref_ip
    out = {}
    seeds = [2024, 2025]
    for c, seed in enumerate(seeds):
        kps, raw, mask = synth_match_inputs(seed, n_img=6 if c == 0 else 4)
        n = len(kps)
        key_of = {}
        images = []
        for i in range(n):
            im = np.full((2, 2), i, np.uint8)
            images.append(im)
        kp_objs = [[_Kp(x, y) for x, y in kps[i]] for i in range(n)]
        des = [np.full((len(kps[i]), 4), i, np.float32) for i in range(n)]
        for i in range(n):
            key_of[id(kp_objs[i])] = i
        saved = (ref_ip.detect_compute_sift, ref_ip.match_sift_features)
        ref_ip.detect_compute_sift = lambda im, nf, verbose=False: (kp_objs[int(im[0, 0])], des[int(im[0, 0])])

        def fake_match(kp1, des1, kp2, des2, verbose=False):
            i1, i2 = raw[(key_of[id(kp1)], key_of[id(kp2)])]
            return None, i1.tolist(), None, i2.tolist()
        ref_ip.match_sift_features = fake_match
        random.seed(seed)
        try:
            with contextlib.redirect_stdout(io.StringIO()) as log:
                _, _, points, src, dst, lmi, lm_num = ref_ip.build_matching_graph(images, mask.tolist(), 'sift', False)
        finally:
            ref_ip.detect_compute_sift, ref_ip.match_sift_features = saved
        n_warn = log.getvalue().count("in-consistent")
        out["c%d_seed" % c], out["c%d_n" % c], out["c%d_mask" % c] = np.array(seed), np.array(n), mask[:n, :n]
        out["c%d_landmark_num" % c], out["c%d_n_inconsistent" % c] = np.array(lm_num), np.array(n_warn)
        for i in range(n):
            out["c%d_kp_%d" % (c, i)] = kps[i]
            np.testing.assert_array_equal(points[i], kps[i])
            for j in range(n):
                if i < j:
                    out["c%d_raw1_%d_%d" % (c, i, j)], out["c%d_raw2_%d_%d" % (c, i, j)] = raw[(i, j)]
                out["c%d_src_%d_%d" % (c, i, j)] = np.array(src[i][j], np.int64)
                out["c%d_dst_%d_%d" % (c, i, j)] = np.array(dst[i][j], np.int64)
                out["c%d_lm_%d_%d" % (c, i, j)] = np.array(lmi[i][j], np.int64)
        sizes = {k: len(v[0]) for k, v in raw.items()}
        print("match_graph case %d: %d landmarks, %d inconsistent, raw pair sizes %s" % (c, lm_num, n_warn, sizes))
    out["n_cases"] = np.array(len(seeds))
    np.savez_compressed(os.path.join(OUT, "match_graph.npz"), **out)


def gen_keyframe_mat():
    """KeyFrame.convert_keypoint_to_array (key_frame.py:58-73) and save_to_mat (:75-107) of the reference; the written
    .mat files are read back with scipy and their arrays stored."""
    import tempfile
    import scipy.io as sio
    import cv2
    from key_frame import KeyFrame as RefKeyFrame
    rng = np.random.default_rng(1415)
    out = {}
    rots = [cv2.Rodrigues(BASE_ROT.reshape(3, 1))[0], BASE_ROT.copy(), np.eye(3),
            cv2.Rodrigues(np.array([[0.3], [-2.9], [0.8]]))[0]]
    for c, rot in enumerate(rots):
        n = [30, 12, 0, 7][c]
        pts = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1)
        des = rng.integers(0, 255, (n, 16)).astype(np.float32) + 1
        ptz = np.array([rng.uniform(40, 70), rng.uniform(-12, -6), rng.uniform(2000, 4000)])
        kf = RefKeyFrame(None, 100 + c, CC, rot, U, V, ptz[0], ptz[1], ptz[2])
        as_list = c in (0, 3)
        kf.feature_pts = [_Kp(x, y) for x, y in pts] if as_list else pts.copy()
        kf.feature_des = des.copy() if as_list else des.astype(np.float64)
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "kf.mat")
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                kf.save_to_mat(path)
            d = sio.loadmat(path)
        out["c%d_rot" % c], out["c%d_pts" % c], out["c%d_des" % c], out["c%d_ptz" % c] = rot, pts, des, ptz
        out["c%d_as_list" % c] = np.array(as_list)
        out["c%d_im_name" % c] = np.array(str(np.asarray(d['im_name']).ravel()[0]))
        for k in ("keypoint", "descriptor", "camera", "ptz"):
            out["c%d_mat_%s" % (c, k)] = d[k]
        print("keyframe_mat case %d:" % c, {k: d[k].shape for k in ("keypoint", "descriptor", "camera", "ptz")}, d['camera'][3:6].ravel())
    out["n_cases"] = np.array(len(rots))
    np.savez_compressed(os.path.join(OUT, "keyframe_mat.npz"), **out)


def fake_window_ba(key_frame_cls, seed):
    """Stand-in for bundle_adjustment() in the sliding-window test: nudges every pose, gives keyframe k 5+k keypoints
    (none to the third of the window) and records the arguments it was called with."""
    calls = []

    def run(images, image_indices, feature_method, initial_ptzs, center, rotation, u, v, save_path, *a, **k):
        rng = np.random.default_rng(seed + len(calls))
        calls.append((list(image_indices), np.array(initial_ptzs, dtype=np.float64).copy()))
        kfs = []
        for i, idx in enumerate(image_indices):
            kf = key_frame_cls(images[i], idx, center, rotation, u, v, initial_ptzs[i][0] + 0.01 * (i + 1),
                               initial_ptzs[i][1] - 0.02, initial_ptzs[i][2] + i)
            n = 0 if i == 2 else 5 + i
            kf.feature_pts = [_Kp(x, y) for x, y in rng.uniform(0, 700, (n, 2))]
            kf.feature_des = rng.integers(1, 200, (n, 8)).astype(np.float32)
            kf.landmark_index = np.arange(n, dtype=np.int32)
            kfs.append(kf)
        return rng.uniform(-30, 30, (40, 2)), kfs
    return run, calls


def gen_sliding_window():
    """RandomForestMap.bundle_adjustment_processing (scene_map.py:202-244) with bundle_adjustment replaced by the fake
    above: which keyframes enter the window, which survive, and the converted feature arrays."""
    import scene_map as ref_scene_map
    from key_frame import KeyFrame as RefKeyFrame
    out = {}
    for c, n_kf in enumerate([14, 3, 10]):
        m = ref_scene_map.RandomForestMap()
        for k in range(n_kf):
            m.keyframe_list.append(RefKeyFrame(None, 10 * k + 1, CC, BASE_ROT, U, V, 40.0 + 2 * k, -8.0 - 0.1 * k, 2500.0 + 50 * k))
        run, calls = fake_window_ba(RefKeyFrame, 500 + c)
        saved = ref_scene_map.bundle_adjustment
        ref_scene_map.bundle_adjustment = run
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                m.bundle_adjustment_processing()
        finally:
            ref_scene_map.bundle_adjustment = saved
        out["c%d_n_kf" % c] = np.array(n_kf)
        out["c%d_call_indices" % c], out["c%d_call_ptzs" % c] = np.array(calls[0][0]), calls[0][1]
        out["c%d_result_indices" % c] = np.array([kf.img_index for kf in m.keyframe_list])
        out["c%d_result_ptz" % c] = np.array([[kf.pan, kf.tilt, kf.f] for kf in m.keyframe_list])
        out["c%d_result_nfeat" % c] = np.array([kf.get_feature_num() for kf in m.keyframe_list])
        last = m.keyframe_list[-1]
        out["c%d_last_pts" % c], out["c%d_last_des" % c] = last.feature_pts, last.feature_des
        print("sliding_window case %d: %d keyframes -> window %s -> %s" % (c, n_kf, calls[0][0], out["c%d_result_indices" % c]))
    out["n_cases"] = np.array(3)
    np.savez_compressed(os.path.join(OUT, "sliding_window.npz"), **out)


def gen_relocalization():
    """relocalization._compute_residual (:22-40) and the least_squares call of :186-187 as the reference makes it
    (ftol=1e-4) plus a tight run (ftol=xtol=gtol=1e-15 = converged) on seeded ray <-> pixel matches."""
    import relocalization as ref_reloc
    rng = np.random.default_rng(1617)
    out = {"uv": np.array([U, V])}
    cases = [(60, 0.0, (1.5, -0.8, 150.0)), (200, 0.5, (-2.0, 1.0, -300.0)), (12, 1.0, (0.5, 0.3, 80.0)), (400, 0.3, (4.0, -1.5, 500.0))]
    for c, (n, noise, dpose) in enumerate(cases):
        gt = np.array([rng.uniform(45, 70), rng.uniform(-12, -7), rng.uniform(2200, 3800)])
        pts = np.stack([rng.uniform(20, W - 20, n), rng.uniform(20, H - 20, n)], 1)
        rays = np.array([ref.TransFunction.from_image_to_ray(U, V, gt[2], gt[0], gt[1], x, y) for x, y in pts])
        pts_noisy = pts + rng.normal(0, noise, pts.shape) if noise > 0 else pts
        pose0 = gt + np.array(dpose)
        r0 = ref_reloc._compute_residual(pose0, rays, pts_noisy, U, V)
        with contextlib.redirect_stdout(io.StringIO()):
            asis = ref.least_squares(ref_reloc._compute_residual, pose0, verbose=2, x_scale='jac', ftol=1e-4, method='trf',
                                     args=(rays, pts_noisy, U, V))
            tight = ref.least_squares(ref_reloc._compute_residual, pose0, verbose=0, x_scale='jac', ftol=1e-15, xtol=1e-15,
                                      gtol=1e-15, method='trf', args=(rays, pts_noisy, U, V))
        out["c%d_gt" % c], out["c%d_pose0" % c], out["c%d_rays" % c], out["c%d_points" % c] = gt, pose0, rays, pts_noisy
        out["c%d_r0" % c] = r0
        out["c%d_x_asis" % c], out["c%d_cost_asis" % c], out["c%d_nfev_asis" % c] = asis.x, np.array(asis.cost), np.array(asis.nfev)
        out["c%d_x_tight" % c], out["c%d_cost_tight" % c] = tight.x, np.array(tight.cost)
        print("relocalization case %d: n=%d gt=%s asis=%s (nfev %d) tight=%s" % (c, n, gt, asis.x, asis.nfev, tight.x))
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "relocalization.npz"), **out)


class _IntIndexNp:
    """np proxy for the reference module: np.delete with the float index arrays ptz_slam.py:309-315 builds (numpy >= 1.19
    refuses them) and np.row_stack (removed in numpy 2.x; an alias of vstack) - values unchanged."""

    def __getattr__(self, k):
        return getattr(np, k)

    @staticmethod
    def delete(arr, obj, axis=None):
        return np.delete(arr, np.asarray(obj).astype(np.int64), axis=axis)

    @staticmethod
    def row_stack(a):
        return np.vstack(a)


def cov_probe_vector(n):
    return np.random.default_rng(n).uniform(0.5, 1.5, n)


def gen_tracking():
    """The per-frame loop of the reference - PtzSlam.init_system (:140-208) and tracking (:390-456: optical-flow
    bookkeeping, predict, ekf_update, remove_rays, add_rays, keyframe test) - over a seeded sequence, with the two OpenCV
    calls (detect_compute_sift_array, matching_and_ransac) replaced by tests/synth_front_end.py."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from synth_front_end import SyntheticFrontEnd
    import ptz_slam as ref_ptz_slam
    from key_frame import KeyFrame as RefKeyFrame
    out = {}
    runs = [(31, 11, None), (32, 10, 4)]
    for c, (seed, n_frames, bad_from) in enumerate(runs):
        fe = SyntheticFrontEnd(seed, n_frames, bad_from)
        saved = (ref_ptz_slam.detect_compute_sift_array, ref_ptz_slam.matching_and_ransac, ref_ptz_slam.np)
        ref_ptz_slam.detect_compute_sift_array = lambda img, n, norm=True: fe.detect_keypoints(img, n)
        ref_ptz_slam.matching_and_ransac = fe.matching_and_ransac
        ref_ptz_slam.np = _IntIndexNp()
        try:
            with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                slam = ref.PtzSlam()
                cam0 = make_camera(fe.gt[0] + np.array([0.02, -0.01, 3.0]))
                slam.init_system(fe.image(0), cam0, fe.bounding_box)
                slam.keyframe_map.add_first_keyframe(RefKeyFrame(None, 0, CC, BASE_ROT, U, V, *cam0.get_ptz()))
                out["c%d_rays_0" % c], out["c%d_cov_diag_0" % c] = slam.rays.copy(), np.diag(slam.state_cov).copy()
                out["c%d_prev_kp_0" % c] = np.asarray(slam.previous_keypoints, np.float64)
                for k in range(1, n_frames):
                    slam.tracking(fe.image(k), 80, fe.bounding_box)
                    out["c%d_ptz_%d" % (c, k)] = slam.current_camera.get_ptz()
                    out["c%d_vel_%d" % (c, k)] = np.array(slam.velocity)
                    out["c%d_rays_%d" % (c, k)] = slam.rays.copy()
                    if k == n_frames - 1:
                        out["c%d_des_%d" % (c, k)] = np.asarray(slam.des)
                    if k == 1 and c == 0:
                        out["c%d_cov_%d" % (c, k)] = slam.state_cov.copy()
                    # every frame: the diagonal and a seeded random projection of the full matrix (catches any wrong entry)
                    out["c%d_cov_diag_%d" % (c, k)] = np.diag(slam.state_cov).copy()
                    out["c%d_cov_probe_%d" % (c, k)] = slam.state_cov @ cov_probe_vector(slam.state_cov.shape[0])
                    out["c%d_prev_kp_%d" % (c, k)] = np.asarray(slam.previous_keypoints, np.float64)
                    out["c%d_prev_idx_%d" % (c, k)] = np.asarray(slam.previous_keypoints_index, np.float64)
                    out["c%d_flags_%d" % (c, k)] = np.array([slam.new_keyframe, slam.tracking_lost, slam.bad_tracking_cnt,
                                                             len(slam.cameras)], np.int64)
        finally:
            ref_ptz_slam.detect_compute_sift_array, ref_ptz_slam.matching_and_ransac, ref_ptz_slam.np = saved
        out["c%d_seed" % c], out["c%d_n_frames" % c] = np.array(seed), np.array(n_frames)
        out["c%d_bad_from" % c] = np.array(-1 if bad_from is None else bad_from)
        out["c%d_cam0" % c] = cam0.get_ptz() if False else fe.gt[0] + np.array([0.02, -0.01, 3.0])
        print("tracking run %d: rays per frame %s, flags last %s, ptz err last %s" % (
            c, [len(out["c%d_rays_%d" % (c, k)]) for k in range(n_frames)], out["c%d_flags_%d" % (c, n_frames - 1)],
            out["c%d_ptz_%d" % (c, n_frames - 1)] - fe.gt[-1]))
        print("   flags:", [out["c%d_flags_%d" % (c, k)].tolist() for k in range(1, n_frames)])
    out["n_runs"] = np.array(len(runs))
    np.savez_compressed(os.path.join(OUT, "tracking.npz"), **out)


def gen_camera_3d():
    """PTZCamera matrices and world-point helpers: compute_rotation_matrix (:65-81), recompute_matrix (:117-141),
    project_3d_point(s) (:154-189), back_project_to_3d_point(s) (:236-285)."""
    rng = np.random.default_rng(2021)
    out = {}
    for c, disp in enumerate([None, DISP]):
        ptz = np.array([rng.uniform(45, 70), rng.uniform(-12, -7), rng.uniform(2200, 3800)])
        cam = make_camera(ptz, disp)
        # world points: the ground points seen at pixels in and around the image, lifted by up to 2 m
        seed_px = np.stack([rng.uniform(-300, W + 300, 60), rng.uniform(-200, H + 200, 60)], 1)
        field = cam.back_project_to_3d_points(seed_px) + np.stack([np.zeros(60), np.zeros(60), rng.uniform(0, 2, 60)], 1)
        cam.recompute_matrix()
        pts_all, _ = cam.project_3d_points(field)
        pts_in, idx_in = cam.project_3d_points(field, H, W)
        px = np.stack([rng.uniform(0, W, 25), rng.uniform(0, H, 25)], 1)
        out["c%d_ptz" % c], out["c%d_disp" % c] = ptz, (np.zeros(6) if disp is None else disp)
        out["c%d_R" % c], out["c%d_P" % c] = cam.compute_rotation_matrix(), cam.projection_matrix.copy()
        out["c%d_pan" % c], out["c%d_tilt" % c] = cam.compute_pan_matrix(), cam.compute_tilt_matrix()
        out["c%d_field" % c], out["c%d_pts_all" % c], out["c%d_pts_in" % c], out["c%d_idx_in" % c] = field, pts_all, pts_in, idx_in
        out["c%d_px" % c], out["c%d_ground" % c] = px, cam.back_project_to_3d_points(px)
        print("camera_3d case %d: %d of %d court points in the image" % (c, len(idx_in), len(field)))
    # TransFunction world-point helpers (transformation.py:23-97, 178-249)
    T = ref.TransFunction
    R0 = make_camera([0, 0, 2000]).base_rotation
    q = rng.uniform(-1, 1, (20, 8))
    tf = {k: [] for k in ("to_image", "to_3d", "to_ray", "ray_rel", "rel_image", "to_rel")}
    tf_in = []
    for a in q:
        pan, tilt, f = 55 + 8 * a[0], -9 + 2 * a[1], 3000 + 600 * a[2]
        px = np.array([640 + 500 * a[3], 360 + 300 * a[4]])
        world = np.asarray(T.from_image_to_3dpoint(U, V, f, pan, tilt, CC, R0, px)) + np.array([0, 0, 1 + a[5]])
        theta, phi = pan + 6 * a[6], tilt + 4 * a[7]
        tf_in.append([pan, tilt, f, px[0], px[1], world[0], world[1], world[2], theta, phi])
        tf["to_image"].append(T.from_3dpoint_to_image(U, V, f, pan, tilt, CC, R0, world))
        tf["to_3d"].append(T.from_image_to_3dpoint(U, V, f, pan, tilt, CC, R0, px))
        tf["to_ray"].append(T.from_3dpoint_to_ray(CC, world, R0))
        rel = T.from_ray_to_relative_3dpoint(theta, phi)
        tf["ray_rel"].append(rel)
        tf["rel_image"].append(T.from_relative_3dpoint_to_image(U, V, f, pan, tilt, rel))
        tf["to_rel"].append(T.from_3dpoint_to_relative_3dpoint(CC, R0, world))
    out["tf_in"], out["tf_R0"] = np.array(tf_in), R0
    for k, vals in tf.items():
        out["tf_" + k] = np.array(vals, dtype=np.float64)
    np.savez(os.path.join(OUT, "camera_3d.npz"), **out)


def gen_util_noise():
    """util.add_gauss / add_outliers (:99-139, Python `random` seeded), uniform_point_sample_on_field (:186-203),
    compute_error_data (:301-319)."""
    import random
    import util as ref_util
    rng = np.random.default_rng(1819)
    pts = np.stack([rng.uniform(-5, W + 5, 80), rng.uniform(-5, H + 5, 80)], 1)
    pts[0], pts[1], pts[2] = (W + 20, H + 20), (-20, -20), (W - 0.5, 10)          # all four clamps
    random.seed(4242)
    g = ref_util.add_gauss(pts, 3.0, W, H)
    random.seed(4343)
    o = ref_util.add_outliers(pts, 1.5, W, H, 35)
    field = ref_util.uniform_point_sample_on_field(118, 70, 7, 5)
    a = (rng.normal(50, 3, 40), rng.normal(-9, 1, 40), rng.normal(3000, 200, 40))
    b = (a[0] + rng.normal(0, 0.1, 40), a[1] + rng.normal(0, 0.05, 40), a[2] + rng.normal(0, 9, 40))
    mean, std = ref_util.compute_error_data(a, b)
    np.savez(os.path.join(OUT, "util_noise.npz"), pts=pts, gauss=g, outliers=o, field=field, est=np.array(a), gt=np.array(b),
             err_mean=np.array(mean), err_std=np.array(std))
    print("util_noise: clamped %d, outliers moved %d" % (((g == 0) | (g[:, :1] == W - 1)).sum(), (np.abs(o - pts) > 20).any(1).sum()))


def court_rays_from_reference(n_points):
    """The reference's court landmarks: DataSynthesize.generate_points (synthesized_court_sequence/synthesize_basketball.py:34-59,
    random.seed(1) inside) executed UNMODIFIED - only the class definition is compiled from the file, because the module body
    loads .mat files that are not shipped - and TransFunction.from_3dpoint_to_ray (transformation.py) for the rays."""
    import ast
    import random
    import cv2 as cv
    path = os.path.join(ref_import.REFERENCE_DIR, "synthesized_court_sequence", "synthesize_basketball.py")
    tree = ast.parse(open(path).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "DataSynthesize"][0]
    ns = {"np": np, "random": random}
    exec(compile(ast.Module([cls], []), path, "exec"), ns)
    pts = ns["DataSynthesize"].generate_points(n_points).astype(np.float64)
    R = np.zeros((3, 3))
    cv.Rodrigues(BASE_ROT, R)
    rays = np.array([ref.TransFunction.from_3dpoint_to_ray(CC, p, R) for p in pts])
    return pts, rays


def _reference_court_run(out_name, n_frames, blackout, checkpoints):
    """The loop of experiment.py:22-46 with the UNMODIFIED reference classes over the synthesized court sequence: init_system,
    add_keyframe, per frame `tracking`, then - as the reference's driver does - relocalize + init_system when tracking is lost,
    add_keyframe -> Map.add_keyframe_with_ba -> bundle_adjustment -> build_matching_graph -> least_squares when the new-keyframe
    rule fires.  Only the OpenCV calls are replaced (tests/court_sequence.py), plus the debug image dumps of
    bundle_adjustment.py:153-162 and relocalization.py:177-178."""
    import random
    import types
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from court_sequence import CourtSequence
    import ptz_slam as ref_ptz_slam
    import image_process as ref_ip
    import bundle_adjustment as ref_ba
    import relocalization as ref_reloc
    n_points, seed = 1000, 1001
    pts3d, court_rays = court_rays_from_reference(n_points)
    fe = CourtSequence(court_rays, n_frames, seed, blackout=blackout)
    out = {"court_points": pts3d, "court_rays": court_rays, "n_frames": np.array(n_frames), "seed": np.array(seed),
           "blackout": np.array(fe.blackout, np.int64)}
    patched = [(ref_ptz_slam, "detect_compute_sift_array", lambda img, n, norm=True: fe.detect_keypoints(img, n)),
               (ref_ptz_slam, "matching_and_ransac", fe.matching_and_ransac),
               (ref_ptz_slam, "np", _IntIndexNp()),
               (ref_ip, "detect_compute_sift", lambda im, nf, verbose=False: fe.detect_sift(im)),
               (ref_ip, "match_sift_features", lambda kp1, des1, kp2, des2, verbose=False: fe.match_sift(kp1, des1, kp2, des2)),
               (ref_ba, "draw_matches", lambda *a, **k: None),
               (ref_ba, "cv", types.SimpleNamespace(imwrite=lambda *a, **k: True)),
               (ref_reloc, "detect_compute_sift_array", lambda img, n, norm=False: fe.detect_array(img, n)),
               (ref_reloc, "match_sift_features", lambda kp1, des1, kp2, des2, pts_array=True: fe.match(kp1, des1, kp2, des2)),
               (ref_reloc, "cv", types.SimpleNamespace(imwrite=lambda *a, **k: True))]
    saved = [(m, name, getattr(m, name)) for m, name, _ in patched]
    for m, name, fn in patched:
        setattr(m, name, fn)
    random.seed(seed)
    ba_events, reloc_events = [], []
    try:
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            slam = ref.PtzSlam()
            cam0_ptz = fe.gt[0] + np.array([0.02, -0.01, 3.0])
            cam0 = make_camera(cam0_ptz)
            slam.init_system(fe.image(0), cam0, fe.bounding_box)
            slam.add_keyframe(fe.image(0), cam0, 0, enable_rf=False)
            out["cam0"] = cam0_ptz
            out["rays_0"] = slam.rays.copy()
            for k in range(1, n_frames):
                img = fe.image(k)
                slam.tracking(img, 80, fe.bounding_box)
                flags = [slam.new_keyframe, slam.tracking_lost, slam.bad_tracking_cnt, len(slam.cameras)]
                if slam.tracking_lost:                                                  # experiment.py:39-43
                    out["lost_ptz_%d" % k] = slam.current_camera.get_ptz()
                    relocalized_camera = slam.relocalize(img, slam.current_camera, enable_rf=False)
                    out["reloc_ptz_%d" % k] = relocalized_camera.get_ptz()
                    slam.init_system(img, relocalized_camera, fe.bounding_box)
                    reloc_events.append(k)
                elif slam.new_keyframe:                                                 # :45-46
                    slam.add_keyframe(img, slam.current_camera, k, enable_rf=False)
                    e = len(ba_events)
                    ba_events.append(k)
                    kfs = slam.keyframe_map.keyframe_list
                    out["ba%d_kf_ptz" % e] = np.array([[kf.pan, kf.tilt, kf.f] for kf in kfs])
                    out["ba%d_kf_index" % e] = np.array([kf.img_index for kf in kfs], np.int64)
                    out["ba%d_global_ray" % e] = np.asarray(slam.keyframe_map.global_ray, np.float64)
                    for i, kf in enumerate(kfs):
                        out["ba%d_lm_%d" % (e, i)] = np.asarray(kf.landmark_index, np.int64)
                        out["ba%d_pts_%d" % (e, i)] = np.array([p.pt for p in kf.feature_pts], np.float64).reshape(-1, 2)
                out["ptz_%d" % k] = slam.current_camera.get_ptz()
                out["vel_%d" % k] = np.array(slam.velocity)
                out["flags_%d" % k] = np.array(flags, np.int64)
                out["prev_idx_%d" % k] = np.asarray(slam.previous_keypoints_index, np.float64)
                out["n_rays_%d" % k] = np.array(len(slam.rays))
                if k in checkpoints:
                    # full state of the reference: the test re-starts from it, because the EKF recursion amplifies rounding
                    # differences by ~10x every 10 frames (DESIGN.md section 2 finding 4) - free-running segments stay short
                    out["ck%d_rays" % k], out["ck%d_cov" % k] = slam.rays.copy(), slam.state_cov.copy()
                    out["ck%d_des" % k] = np.asarray(slam.des)
                    out["ck%d_prev_kp" % k] = np.asarray(slam.previous_keypoints, np.float64)
                    out["ck%d_prev_idx" % k] = np.asarray(slam.previous_keypoints_index, np.float64)
                    out["ck%d_vel" % k], out["ck%d_ptz" % k] = np.array(slam.velocity), slam.cameras[-1].get_ptz()
                    out["ck%d_counts" % k] = np.array([slam.bad_tracking_cnt, len(slam.cameras)], np.int64)
                    kfs = slam.keyframe_map.keyframe_list
                    out["ck%d_kf_ptz" % k] = np.array([[kf.pan, kf.tilt, kf.f] for kf in kfs])
                    out["ck%d_kf_index" % k] = np.array([kf.img_index for kf in kfs], np.int64)
                if k % 25 == 0 or k == n_frames - 1:
                    out["rays_%d" % k] = slam.rays.copy()
                    out["prev_kp_%d" % k] = np.asarray(slam.previous_keypoints, np.float64)
                    out["cov_diag_%d" % k] = np.diag(slam.state_cov).copy()
                    out["cov_probe_%d" % k] = slam.state_cov @ cov_probe_vector(slam.state_cov.shape[0])
    finally:
        for m, name, fn in saved:
            setattr(m, name, fn)
    out["ba_frames"] = np.array(ba_events, np.int64)
    out["reloc_frames"] = np.array(reloc_events, np.int64)
    out["checkpoints"] = np.array(sorted(checkpoints), np.int64)
    err = np.array([out["ptz_%d" % k] - fe.gt[k] for k in range(1, n_frames)])
    print("%s: %d frames, rays %d -> %d, keyframe BA at frames %s (keyframes after the last: %d, landmarks %d), relocalised at %s" %
          (out_name, n_frames, len(out["rays_0"]), int(out["n_rays_%d" % (n_frames - 1)]), ba_events,
           len(out["ba%d_kf_ptz" % (len(ba_events) - 1)]) if ba_events else 1,
           len(out["ba%d_global_ray" % (len(ba_events) - 1)]) if ba_events else 0, reloc_events))
    for k in reloc_events:
        print("   frame %d: lost pose error %s -> relocalised pose error %s" % (k, out["lost_ptz_%d" % k] - fe.gt[k], out["reloc_ptz_%d" % k] - fe.gt[k]))
    print("   pose error vs ground truth: mean |d| %s, max |d| %s" % (np.abs(err).mean(0), np.abs(err).max(0)))
    np.savez_compressed(os.path.join(OUT, out_name), **out)


def gen_cfg1():
    """Config 1 end to end: 150 frames, the tracker never loses the sequence; keyframe bundle adjustments of 2, 3, 4 and 5 keyframes."""
    _reference_court_run("cfg1_court.npz", 150, None, (50, 100))


def gen_cfg1_relocalize():
    """The same sequence with an optical-flow blackout on frames 55-58: tracking_lost at frame 58 -> PtzSlam.relocalize ->
    relocalization_camera (nearest keyframe by match count, its pixels -> rays, 3-parameter least_squares) -> init_system."""
    _reference_court_run("cfg1_relocalize.npz", 90, (55, 59), ())


if __name__ == "__main__":
    if len(sys.argv) > 1:                           # regenerate only the named goldens: make_golden.py ray_bookkeeping ...
        for name in sys.argv[1:]:
            globals()["gen_" + name]()
        sys.exit(0)
    gen_ray_bookkeeping()
    gen_match_graph()
    gen_keyframe_mat()
    gen_sliding_window()
    gen_relocalization()
    gen_tracking()
    gen_cfg1()
    gen_cfg1_relocalize()
    gen_cfg2_reference()
    gen_util_noise()
    gen_camera_3d()
    gen_keyframe_map()
    gen_projection()
    gen_backprojection()
    gen_h_jacobian()
    gen_ekf()
    gen_ba()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
