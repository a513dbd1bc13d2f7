"""CPU: pins the oracle (oracle/ptz_oracle.py) against goldens produced by the unmodified reference."""
import numpy as np

from conftest import load_golden, graph_from_npz
from oracle import ptz_oracle as O
import ptz_slam_b200  # noqa: F401
from ptz_slam_b200 import synth

H, W = synth.IMAGE_H, synth.IMAGE_W


def test_project_ray_scalar_and_vec():
    d = load_golden("projection.npz")
    u, v = d["uv"]
    for tag, disp in (("nodisp", None), ("disp", d["disp"])):
        g = d["project_ray_" + tag]
        for c, ptz in enumerate(d["ptzs"]):
            x, y, _ = O.project_rays_vec(ptz[0], ptz[1], ptz[2], u, v, d["rays"], disp)
            np.testing.assert_allclose(np.stack([x, y], 1), g[c], rtol=1e-12, atol=1e-9)
            for r in (0, 5, 17):
                np.testing.assert_allclose(O.project_ray(ptz[0], ptz[1], ptz[2], u, v, d["rays"][r], disp), g[c, r],
                                           rtol=1e-13, atol=1e-10)


def test_from_ray_to_image():
    d = load_golden("projection.npz")
    u, v = d["uv"]
    g = d["from_ray_to_image"]
    for c, ptz in enumerate(d["ptzs"]):
        x, y = O.from_ray_to_image_vec(u, v, ptz[2], ptz[0], ptz[1], d["rays"][:, 0], d["rays"][:, 1])
        np.testing.assert_allclose(np.stack([x, y], 1), g[c], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(O.from_ray_to_image(u, v, ptz[2], ptz[0], ptz[1], *d["rays"][3]), g[c, 3], rtol=1e-14)


def test_project_rays_filter_contract():
    d = load_golden("projection.npz")
    u, v = d["uv"]
    for c, ptz in enumerate(d["ptzs"]):
        pts, idx = O.project_rays(ptz[0], ptz[1], ptz[2], u, v, d["prs_rays_%d" % c], H, W)
        assert idx.dtype == np.float64
        np.testing.assert_array_equal(idx, d["prs_index_%d" % c])
        np.testing.assert_allclose(pts, d["prs_points_%d" % c], rtol=1e-12, atol=1e-9)
        pts2, idx2 = O.project_rays(ptz[0], ptz[1], ptz[2], u, v, d["prs_rays_%d" % c][:17])
        assert len(idx2) == 0
        np.testing.assert_allclose(pts2, d["prs_all_points_%d" % c], rtol=1e-12, atol=1e-9)


def test_back_projection():
    d = load_golden("backprojection.npz")
    u, v = d["uv"]
    for tag, disp in (("nodisp", None), ("disp", d["disp"])):
        g = d["back_project_" + tag]
        for c, ptz in enumerate(d["ptzs"]):
            r = O.back_project_to_rays_vec(ptz[0], ptz[1], ptz[2], u, v, d["points"], disp)
            np.testing.assert_allclose(r, g[c], rtol=1e-11, atol=1e-11)
            np.testing.assert_allclose(O.back_project_to_ray(ptz[0], ptz[1], ptz[2], u, v, *d["points"][4], disp=disp),
                                       g[c, 4], rtol=1e-13)
    g = d["from_image_to_ray"]
    for c, ptz in enumerate(d["ptzs"]):
        for k in (0, 7, 33):
            np.testing.assert_allclose(O.from_image_to_ray(u, v, ptz[2], ptz[0], ptz[1], *d["points"][k]), g[c, k], rtol=1e-13)
        # the two back-projection formulations agree when disp = 0
        np.testing.assert_allclose(O.back_project_to_rays_vec(ptz[0], ptz[1], ptz[2], u, v, d["points"]), g[c],
                                   rtol=1e-11, atol=1e-11)


def test_h_jacobian_fd_and_analytic():
    d = load_golden("h_jacobian.npz")
    u, v = d["uv"]
    for tag, disp in (("nodisp", None), ("disp", d["disp"])):
        ptz, rays, Hg = d["ptz_" + tag], d["rays_" + tag], d["H_" + tag]
        Ho = O.compute_h_jacobian(ptz[0], ptz[1], ptz[2], u, v, rays, disp)
        # same central differences; only rounding of the two projection evaluations differs (|x| eps / 2 delta)
        np.testing.assert_allclose(Ho, Hg, rtol=1e-9, atol=2e-9)
    # analytic blocks agree with the reference's central differences to the FD truncation error (disp = 0)
    ptz, rays, Hg = d["ptz_nodisp"], d["rays_nodisp"], d["H_nodisp"]
    Jc, Jr = O.jacobian_blocks_analytic(ptz[0], ptz[1], ptz[2], rays[:, 0], rays[:, 1])
    for i in range(len(rays)):
        np.testing.assert_allclose(Jc[i], Hg[2 * i:2 * i + 2, 0:3], rtol=2e-9, atol=2e-9)
        np.testing.assert_allclose(Jr[i], Hg[2 * i:2 * i + 2, 3 + 2 * i:5 + 2 * i], rtol=2e-9, atol=2e-9)


def test_ekf_six_frames():
    d = load_golden("ekf.npz")
    u, v = d["uv"]
    s = O.EkfState(d["rays0"], d["ptz0"], u, v)
    for k in range(1, int(d["n_frames"]) + 1):
        O.ekf_predict(s)
        O.ekf_update(s, d["obs_xy_%d" % k], d["obs_idx_%d" % k], H, W)
        np.testing.assert_allclose(s.ptz, d["ptz_%d" % k], rtol=1e-10, atol=1e-9)
        np.testing.assert_allclose(s.velocity, d["vel_%d" % k], rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(s.rays, d["rays_%d" % k], rtol=1e-10, atol=1e-9)
        np.testing.assert_allclose(s.state_cov, d["cov_%d" % k], rtol=1e-7, atol=1e-12)
    # the write-back quirk: pose<->ray and theta<->phi cross covariances stay exactly zero
    assert np.all(s.state_cov[0:3, 3:] == 0)
    assert np.all(s.state_cov[3::2, 4::2] == 0)
    assert np.any(s.state_cov[3::2, 3::2][~np.eye((len(s.rays)), dtype=bool)] != 0)


def test_ekf_cfg2_size_ten_frames():
    """The oracle against the UNMODIFIED reference at the config 2 size: 3 000 rays, every visible ray observed (~1 000 matched per
    frame), 10 consecutive predict + update steps (golden cfg2_reference.npz: the reference takes 2.8 s per frame).  BASELINE.json's
    parameter tolerance, 1e-6 rad / 1e-3 px; measured 2e-11 degrees / 1e-9 px."""
    d = load_golden("cfg2_reference.npz")
    n_frames = int(d["n_frames"])
    seq = synth.make_ekf_sequence(int(d["n_rays"]), n_frames + 1, seed=int(d["seed"]), keep_prob=1.0)
    s = O.EkfState(seq.rays0, seq.ptz_gt[0], synth.PP_U, synth.PP_V)
    tol = np.degrees(1e-6)
    for k in range(1, n_frames + 1):
        O.ekf_predict(s)
        O.ekf_update(s, seq.obs_xy[k], seq.obs_idx[k], synth.IMAGE_H, synth.IMAGE_W)
        e = np.abs(s.ptz - d["ptz_%d" % k])
        assert e[0] < tol and e[1] < tol and e[2] < 1e-3, (k, e)
        assert np.abs(s.velocity - d["vel_%d" % k]).max() < 1e-3
        if "rays_%d" % k in d.files:
            assert np.abs(s.rays - d["rays_%d" % k]).max() < tol
            np.testing.assert_allclose(np.diag(s.state_cov), d["cov_diag_%d" % k], rtol=1e-6, atol=1e-12)
            probe = d["cov_probe_%d" % k]
            v = np.random.default_rng(s.state_cov.shape[0]).uniform(0.5, 1.5, s.state_cov.shape[0])
            np.testing.assert_allclose(s.state_cov @ v, probe, rtol=1e-6, atol=1e-9 * np.abs(probe).max())


def test_ba_residual_lists_and_flat():
    d = load_golden("ba_residual.npz")
    u, v = d["uv"]
    points, src, dst, lmk, M = graph_from_npz(d)
    N = len(points)
    for xk, rk in (("x0", "residual_x0"), ("x1", "residual_x1")):
        r = O.ba_residual_lists(d[xk], N, M, int(d["n_residual"]), points, src, dst, lmk, u, v, d["ptz_init"][0])
        np.testing.assert_allclose(r, d[rk], rtol=1e-12, atol=1e-10)
        cam, lm, xy = synth.flatten_match_graph(points, src, dst, lmk)
        poses, rays = O.ba_unpack(d[xk], N, d["ptz_init"][0])
        rf = O.ba_residual_flat(poses, rays, cam, lm, xy, u, v).ravel()
        np.testing.assert_allclose(rf, d[rk], rtol=1e-9, atol=1e-10)


def test_ba_sparse_jacobian_matches_fd():
    d = load_golden("ba_residual.npz")
    u, v = d["uv"]
    points, src, dst, lmk, M = graph_from_npz(d)
    N = len(points)
    cam, lm, xy = synth.flatten_match_graph(points, src, dst, lmk)
    x = d["x0"].copy()
    poses, rays = O.ba_unpack(x, N, d["ptz_init"][0])
    J = O.ba_jacobian_sparse(poses, rays, cam, lm).toarray()
    f = lambda z: O.ba_residual_flat(*O.ba_unpack(z, N, d["ptz_init"][0]), cam, lm, xy, u, v).ravel()
    for col in (0, 1, 2, 7, 3 * (N - 1), 3 * (N - 1) + 5, len(x) - 1):
        h = 1e-4 if (col < 3 * (N - 1) and col % 3 == 2) else 1e-5
        e = np.zeros_like(x); e[col] = h
        fd = (f(x + e) - f(x - e)) / (2 * h)
        np.testing.assert_allclose(J[:, col], fd, rtol=1e-6, atol=1e-6)
    # normal-equation blocks are the block diagonal of J^T J
    r, Ub, gc, Vb, gl, cost = O.ba_normal_equations(poses, rays, cam, lm, xy, u, v)
    A = J.T @ J
    for c in range(1, N):
        np.testing.assert_allclose(Ub[c], A[3 * (c - 1):3 * c, 3 * (c - 1):3 * c], rtol=1e-12)
    o = 3 * (N - 1)
    for l in (0, 3, M - 1):
        np.testing.assert_allclose(Vb[l], A[o + 2 * l:o + 2 * l + 2, o + 2 * l:o + 2 * l + 2], rtol=1e-12)
    g = J.T @ r.ravel()
    np.testing.assert_allclose(gc[1:].ravel(), g[:o], rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(gl.ravel(), g[o:], rtol=1e-11, atol=1e-9)


def test_trf_restatement_converges_to_reference_solution():
    """Tolerance from BASELINE.json: 1e-6 rad on angles, 1e-3 px on focal length."""
    d = load_golden("ba_solve.npz")
    u, v = d["uv"]
    points, src, dst, lmk, M = graph_from_npz(d)
    N = len(points)
    cam, lm, xy = synth.flatten_match_graph(points, src, dst, lmk)
    ref_pose = d["ptz_init"][0]
    fun = lambda z: O.ba_residual_flat(*O.ba_unpack(z, N, ref_pose), cam, lm, xy, u, v).ravel()
    jac = lambda z: O.ba_jacobian_sparse(*O.ba_unpack(z, N, ref_pose), cam, lm).toarray()
    res = O.trf_solve(fun, jac, d["x0"], ftol=1e-15, xtol=1e-15, gtol=1e-15, max_nfev=80)
    xt = d["x_tight"]
    tol_deg = np.degrees(1e-6)
    pose_err = np.abs(res["x"][:3 * (N - 1)] - xt[:3 * (N - 1)]).reshape(-1, 3)
    assert pose_err[:, :2].max() < tol_deg and pose_err[:, 2].max() < 1e-3
    assert np.abs(res["x"][3 * (N - 1):] - xt[3 * (N - 1):]).max() < tol_deg
    assert abs(res["cost"] - float(d["cost_tight"])) < 1e-9 * float(d["cost_tight"])
    # the reference's own stopping rule (ftol=1e-4) lands within the same tolerance of the tight solution
    res2 = O.trf_solve(fun, jac, d["x0"], ftol=1e-4)
    assert res2["status"] == 2
    e2 = np.abs(res2["x"] - d["x_asis"])
    assert e2[:3 * (N - 1)].reshape(-1, 3)[:, :2].max() < tol_deg and e2[3 * (N - 1):].max() < tol_deg
    assert e2[:3 * (N - 1)].reshape(-1, 3)[:, 2].max() < 1e-3
    assert res2["nfev"] == int(d["nfev_asis"])


def test_sparse_lm_iteration_equals_dense_damped_normal_equations():
    """oracle.ba_lm_iteration_sparse (bench.py's CPU baseline of one LM iteration: Schur complement over the landmarks) against the
    dense solve of (J^T J + alpha D^2) d = -J^T r with D = column norms of J, on a problem small enough to form J densely; the
    C-port evaluators give the same step."""
    from oracle import c_port
    fb = synth.make_flat_ba(10, 300, 1800, seed=4)
    x0, ref, alpha = fb.x0(), fb.ptz_init[0], 1e-3
    args = (x0, fb.n_pose, ref, fb.cam_idx, fb.lm_idx, fb.obs_xy, synth.PP_U, synth.PP_V, alpha)
    step, pred, trial = O.ba_lm_iteration_sparse(*args)
    poses, rays = O.ba_unpack(x0, fb.n_pose, ref)
    J = O.ba_jacobian_sparse(poses, rays, fb.cam_idx, fb.lm_idx).toarray()
    r = O.ba_residual_flat(poses, rays, fb.cam_idx, fb.lm_idx, fb.obs_xy, synth.PP_U, synth.PP_V).ravel()
    D2 = (J * J).sum(0)
    D2[D2 == 0] = 1.0
    d = np.linalg.solve(J.T @ J + alpha * np.diag(D2), -J.T @ r)
    assert np.abs(step - d).max() < 1e-10 * np.abs(d).max()
    Jd = J @ d
    assert abs(pred + (r @ Jd + 0.5 * Jd @ Jd)) < 1e-10 * abs(pred)
    pt, rt = O.ba_unpack(x0 + d, fb.n_pose, ref)
    rr = O.ba_residual_flat(pt, rt, fb.cam_idx, fb.lm_idx, fb.obs_xy, synth.PP_U, synth.PP_V)
    assert abs(trial - 0.5 * np.sum(rr * rr)) < 1e-9 * trial and trial < 0.5 * r @ r
    s2, p2, t2 = O.ba_lm_iteration_sparse(*args, fused=c_port.ba_fused, residual=c_port.ba_residual)
    assert np.abs(s2 - step).max() < 1e-10 * np.abs(step).max() and abs(p2 - pred) < 1e-10 * abs(pred) and abs(t2 - trial) < 1e-9 * trial
