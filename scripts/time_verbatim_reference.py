#!/usr/bin/env python
"""
Times the UNMODIFIED reference (/root/reference/slam_system, imported through oracle/ref_import.py) on the host cores of
the authoring container - BASELINE.md section 5 item 1 / SURVEY.md 8(d) "verbatim reference".  The reference tree does
not travel to the GPU box, so this cannot run inside bench.py there: the record is committed as
profiles/r2_verbatim_reference_cpu.json and quoted by bench.py as `verbatim_reference` (labelled with where it ran).

  python scripts/time_verbatim_reference.py [--quick] [--out profiles/r2_verbatim_reference_cpu.json]

Measured (wall clock, time.perf_counter, best of 3 where a repeat is affordable):
  * bundle_adjustment._compute_residual (bundle_adjustment.py:25-106): one pass, observations/s (a match = 2 observations
    = 4 residuals), at the match-graph sizes a Python triple loop can hold;
  * the exact least_squares call of bundle_adjustment.py:200-202 end to end, cfg1-sized problems only (<= 10 keyframes, the
    window scene_map.py:210 uses);
  * PtzSlam.ekf_update + the predict lines ptz_slam.py:418-426 per ray count: frames/s and matched observations/s.
Beside each figure: the same inputs through the numpy restatement (oracle/ptz_oracle.py) and the plain-C port
(oracle/ptz_oracle_c.c) with the largest difference of the results, i.e. the "restated CPU baseline parity-checked against
the verbatim one at sizes where both run".
"""
import argparse
import contextlib
import copy
import io
import json
import os
import platform
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import, c_port  # noqa: E402
from oracle import ptz_oracle as O  # noqa: E402
import ptz_slam_b200  # noqa: E402,F401
from ptz_slam_b200 import synth  # noqa: E402

U, V, W, H = synth.PP_U, synth.PP_V, synth.IMAGE_W, synth.IMAGE_H


def best_of(fn, n):
    best, out = None, None
    for _ in range(n):
        t0 = time.perf_counter()
        out = fn()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, out


def x0_of(ref, g):
    """bundle_adjustment.py:184-197: poses, then every landmark back-projected from a view that observes it (last write wins)."""
    N = len(g.points)
    x0 = np.zeros(N * 3 + g.n_landmark * 2)
    x0[:3 * N] = g.ptz_init.ravel()
    ls = 3 * N
    for i in range(N):
        p = g.ptz_init[i]
        for j in range(N):
            for a, l in zip(g.src_pt_index[i][j], g.landmark_index[i][j]):
                x0[ls + 2 * l:ls + 2 * l + 2] = ref.TransFunction.from_image_to_ray(U, V, p[2], p[0], p[1], g.points[i][a][0], g.points[i][a][1])
    return x0[3:]


def in_domain_graph(n_kf, n_lm, max_matches):
    """Keyframes of a pan sweep that stays inside the reference's ray parametrisation: d = (tan theta, ..., 1)
    (ptz_camera.py:191-210) only represents world pan angles |theta| < 90 degrees, so the sweep ends below 74 degrees + half a
    field of view.  (A sweep past 90 degrees makes from_image_to_ray wrap theta by 180 degrees and from_ray_to_image mirror y:
    1000 px residuals at x0, and the reference's forward-difference Jacobian then walks least_squares to another stationary
    point than the exact Jacobian does - measured while writing this script)"""
    g = synth.make_match_graph(n_kf=n_kf, n_landmark=n_lm, seed=1001, max_matches=max_matches, pan_step=18.0 / max(n_kf - 1, 1))
    assert g.ptz_gt[:, 0].max() + np.degrees(np.arctan(640.0 / g.ptz_gt[:, 2].min())) < 90.0
    return g


def time_residual(ref, n_kf, n_lm, max_matches, repeats):
    g = in_domain_graph(n_kf, n_lm, max_matches)
    N = len(g.points)
    n_match = sum(len(g.src_pt_index[i][j]) for i in range(N) for j in range(N))
    n_residual = 4 * n_match
    x0 = x0_of(ref, g)
    args = (N, g.n_landmark, n_residual, g.points, g.src_pt_index, g.dst_pt_index, g.landmark_index, U, V, g.ptz_init[0])
    t_ref, r_ref = best_of(lambda: ref.bundle_adjustment._compute_residual(x0, *args), repeats)
    cam, lm, xy = synth.flatten_match_graph(g.points, g.src_pt_index, g.dst_pt_index, g.landmark_index)
    poses, rays = O.ba_unpack(x0, N, g.ptz_init[0])
    t_np, r_np = best_of(lambda: O.ba_residual_flat(poses, rays, cam, lm, xy, U, V), 3)
    t_c, r_c = best_of(lambda: c_port.ba_residual(poses, rays, cam, lm, xy, U, V), 3)
    r_c = r_c[0] if isinstance(r_c, tuple) else r_c
    n_obs = 2 * n_match
    return {"keyframes": N, "landmarks": int(g.n_landmark), "observations": n_obs,
            "verbatim_reference": {"s_per_pass": t_ref, "obs_per_s": n_obs / t_ref, "us_per_obs": 1e6 * t_ref / n_obs, "cores": 1},
            "numpy_restatement": {"s_per_pass": t_np, "obs_per_s": n_obs / t_np, "max_abs_diff_vs_reference": float(np.abs(np.ravel(r_np) - r_ref).max())},
            "c_port": {"s_per_pass": t_c, "obs_per_s": n_obs / t_c, "threads": c_port.max_threads(),
                       "max_abs_diff_vs_reference": float(np.abs(np.ravel(r_c) - r_ref).max())}}


def time_least_squares(ref, n_kf, n_lm, max_matches):
    g = in_domain_graph(n_kf, n_lm, max_matches)
    N = len(g.points)
    n_match = sum(len(g.src_pt_index[i][j]) for i in range(N) for j in range(N))
    x0 = x0_of(ref, g)
    args = (N, g.n_landmark, 4 * n_match, g.points, g.src_pt_index, g.dst_pt_index, g.landmark_index, U, V, g.ptz_init[0])
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        res = ref.least_squares(ref.bundle_adjustment._compute_residual, x0, verbose=2, x_scale='jac', ftol=1e-4, method='trf', args=args)
    dt = time.perf_counter() - t0
    # the same call through the oracle's restatement of trf_no_bounds with the analytic sparse Jacobian
    cam, lm, xy = synth.flatten_match_graph(g.points, g.src_pt_index, g.dst_pt_index, g.landmark_index)

    def fun(x):
        p, r = O.ba_unpack(x, N, g.ptz_init[0])
        return np.ravel(O.ba_residual_flat(p, r, cam, lm, xy, U, V))

    def jac(x):
        p, r = O.ba_unpack(x, N, g.ptz_init[0])
        return O.ba_jacobian_sparse(p, r, cam, lm).toarray()

    t1 = time.perf_counter()
    rep = O.trf_solve(fun, jac, x0, ftol=1e-4)
    xo = rep["x"]
    dt_o = time.perf_counter() - t1
    nc = 3 * (N - 1)
    d = np.abs(np.asarray(xo) - res.x)
    return {"keyframes": N, "landmarks": int(g.n_landmark), "observations": 2 * n_match, "parameters": int(len(x0)),
            "verbatim_reference": {"s": dt, "nfev": int(res.nfev), "njev": int(res.njev), "status": int(res.status), "cost": float(res.cost),
                                   "s_per_lm_iteration": dt / max(int(res.njev), 1), "cores": 1,
                                   "note": "scipy '2-point' dense Jacobian: (parameters + 1) residual passes per iteration, then an SVD"},
            "oracle_trf_restatement": {"s": dt_o, "nfev": int(rep["nfev"]), "status": int(rep["status"]), "cost": float(rep["cost"]),
                                       "max_angle_diff_deg_vs_reference": float(max(d[0:nc:3].max(), d[1:nc:3].max(), d[nc:].max())),
                                       "max_focal_diff_px_vs_reference": float(d[2:nc:3].max())}}


def time_ekf(ref, n_rays, n_frames):
    seq = synth.make_ekf_sequence(n_rays=n_rays, n_frames=n_frames + 2, seed=1002)
    slam = ref.PtzSlam()
    cam0 = ref.PTZCamera((U, V), np.zeros(3), np.zeros(3))
    cam0.set_ptz(seq.ptz_gt[0])
    slam.cameras = [cam0]
    slam.rays = seq.rays0.copy()
    slam.state_cov = slam.angle_var * np.eye(3 + 2 * len(slam.rays))          # ptz_slam.py:199-200
    slam.state_cov[2][2] = slam.f_var
    st = O.EkfState(seq.rays0, seq.ptz_gt[0], U, V)

    def ref_frame(k):
        slam.current_camera = copy.deepcopy(slam.cameras[-1])                  # ptz_slam.py:418-426
        slam.current_camera.set_ptz(slam.current_camera.get_ptz() + slam.velocity)
        slam.cameras.append(slam.current_camera)
        slam.state_cov[0:3, 0:3] = slam.state_cov[0:3, 0:3] + 5 * np.diag([slam.angle_var, slam.angle_var, slam.f_var])
        slam.ekf_update(seq.obs_xy[k], seq.obs_idx[k], H, W)

    def oracle_frame(k):
        O.ekf_predict(st)
        return O.ekf_update(st, seq.obs_xy[k], seq.obs_idx[k], H, W)

    ref_frame(1); oracle_frame(1)                                              # untimed first frame
    t_ref = t_or = 0.0
    matched = 0
    diff = 0.0
    for k in range(2, n_frames + 2):
        t0 = time.perf_counter(); ref_frame(k); t_ref += time.perf_counter() - t0
        t0 = time.perf_counter(); m = oracle_frame(k); t_or += time.perf_counter() - t0
        matched += len(m)
        diff = max(diff, float(np.abs(slam.current_camera.get_ptz() - st.ptz).max()))
    return {"rays": n_rays, "frames_timed": n_frames, "mean_matched_rays_per_frame": matched / n_frames,
            "verbatim_reference": {"s_per_frame": t_ref / n_frames, "frames_per_s": n_frames / t_ref, "matched_obs_per_s": matched / t_ref, "cores": 1,
                                   "note": "FD Jacobian: 10 projections per matched ray; write-back: Python loop over matched pairs (BLAS threads only inside inv / dot)"},
            "numpy_restatement": {"s_per_frame": t_or / n_frames, "frames_per_s": n_frames / t_or, "blas_threads": os.cpu_count(),
                                  "max_abs_pose_diff_vs_reference": diff}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="small sizes only (seconds)")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_verbatim_reference_cpu.json"))
    args = ap.parse_args()
    ref = ref_import.load()
    out = {"what": "the UNMODIFIED reference (/root/reference/slam_system) timed on the host cores of the authoring container; wall clock",
           "script": "scripts/time_verbatim_reference.py", "host": {"cpus": os.cpu_count(), "machine": platform.machine(), "python": platform.python_version(),
                                                                      "numpy": np.__version__, "scipy": __import__("scipy").__version__},
           "compute_residual": [], "least_squares": [], "ekf_update": []}
    res_sizes = [(5, 48, 30, 3), (10, 600, 200, 3)] + ([] if args.quick else [(24, 3000, 200, 2), (48, 12000, 200, 1)])
    for n_kf, n_lm, mm, rep in res_sizes:
        r = time_residual(ref, n_kf, n_lm, mm, rep)
        out["compute_residual"].append(r)
        print("residual %3d kf %6d lm %7d obs: reference %.3f s/pass = %.3g obs/s (%.2f us/obs) | numpy %.3g obs/s | C port %.3g obs/s | diffs %.1e %.1e" %
              (r["keyframes"], r["landmarks"], r["observations"], r["verbatim_reference"]["s_per_pass"], r["verbatim_reference"]["obs_per_s"],
               r["verbatim_reference"]["us_per_obs"], r["numpy_restatement"]["obs_per_s"], r["c_port"]["obs_per_s"],
               r["numpy_restatement"]["max_abs_diff_vs_reference"], r["c_port"]["max_abs_diff_vs_reference"]), flush=True)
    ls_sizes = [(5, 48, 30)] + ([] if args.quick else [(10, 300, 60)])
    for n_kf, n_lm, mm in ls_sizes:
        r = time_least_squares(ref, n_kf, n_lm, mm)
        out["least_squares"].append(r)
        print("least_squares %d kf %d lm %d obs %d params: reference %.1f s (nfev %d, njev %d, status %d) | oracle TRF %.2f s (nfev %d) | diff %.2e deg %.2e px" %
              (r["keyframes"], r["landmarks"], r["observations"], r["parameters"], r["verbatim_reference"]["s"], r["verbatim_reference"]["nfev"],
               r["verbatim_reference"]["njev"], r["verbatim_reference"]["status"], r["oracle_trf_restatement"]["s"], r["oracle_trf_restatement"]["nfev"],
               r["oracle_trf_restatement"]["max_angle_diff_deg_vs_reference"], r["oracle_trf_restatement"]["max_focal_diff_px_vs_reference"]), flush=True)
    ekf_sizes = [(128, 10), (300, 10)] + ([] if args.quick else [(1000, 5), (2000, 3), (3000, 3)])
    for n_rays, n_frames in ekf_sizes:
        r = time_ekf(ref, n_rays, n_frames)
        out["ekf_update"].append(r)
        print("ekf %4d rays (%.0f matched/frame): reference %.3f s/frame = %.2f frames/s | numpy restatement %.2f frames/s | pose diff %.1e" %
              (r["rays"], r["mean_matched_rays_per_frame"], r["verbatim_reference"]["s_per_frame"], r["verbatim_reference"]["frames_per_s"],
               r["numpy_restatement"]["frames_per_s"], r["numpy_restatement"]["max_abs_pose_diff_vs_reference"]), flush=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
