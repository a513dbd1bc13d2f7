"""GPU parity: BA residual and the fused residual + Jacobian + normal-equation pass vs reference goldens + oracle."""
import numpy as np
import pytest

from conftest import load_golden, graph_from_npz
from oracle import ptz_oracle as O
from oracle import c_port
import ptz_slam_b200  # noqa: F401
from ptz_slam_b200 import synth, _lib
from ptz_slam_b200 import bundle_adjustment as BA

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-9, 1e-9   # BASELINE.json: relative 1e-9 on residuals and Jacobians


def test_compute_residual_golden():
    d = load_golden("ba_residual.npz")
    points, src, dst, lmk, M = graph_from_npz(d)
    N = len(points)
    for xk, rk in (("x0", "residual_x0"), ("x1", "residual_x1")):
        r = BA._compute_residual(d[xk], N, M, int(d["n_residual"]), points, src, dst, lmk, d["uv"][0], d["uv"][1],
                                 d["ptz_init"][0])
        np.testing.assert_allclose(r, d[rk], rtol=RTOL, atol=ATOL)


def test_initial_landmarks_match_reference_x0():
    d = load_golden("ba_residual.npz")
    points, src, dst, lmk, M = graph_from_npz(d)
    N = len(points)
    rays0 = BA.initial_landmarks(points, src, dst, lmk, M, d["ptz_init"], d["uv"][0], d["uv"][1])
    np.testing.assert_allclose(rays0.ravel(), d["x0"][3 * (N - 1):], rtol=RTOL, atol=1e-10)


def _check_normal_equations(fb, x, ref_pose, u, v):
    prob = BA.BAProblem(fb.n_pose, fb.n_landmark, fb.cam_idx, fb.lm_idx, fb.obs_xy, u, v)
    out = prob.normal_equations(x, ref_pose)
    poses, rays = O.ba_unpack(x, fb.n_pose, ref_pose)
    r, U, gc, V, gl, cost = O.ba_normal_equations(poses, rays, fb.cam_idx, fb.lm_idx, fb.obs_xy, u, v)
    np.testing.assert_allclose(out["residual"], r.ravel(), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(prob.residual(x, ref_pose), r.ravel(), rtol=RTOL, atol=ATOL)
    assert abs(out["cost"] - cost) <= 1e-11 * cost
    su, sg = np.abs(U[1:]).max(), np.abs(gc[1:]).max()
    np.testing.assert_allclose(out["U"][1:], U[1:], rtol=RTOL, atol=1e-12 * su)
    np.testing.assert_allclose(out["gc"][1:], gc[1:], rtol=1e-8, atol=1e-11 * sg)
    assert np.all(out["U"][0] == 0) and np.all(out["gc"][0] == 0)      # fixed reference pose
    np.testing.assert_allclose(out["V"], V, rtol=RTOL, atol=1e-12 * np.abs(V).max())
    np.testing.assert_allclose(out["gl"], gl, rtol=1e-8, atol=1e-11 * np.abs(gl).max())
    prob.close()


def check_against_c_port(out, fb, x, ref_pose, u, v):
    """ALL of residual, U, g_c, V, g_l and the cost against the plain-C restatement (oracle/ptz_oracle_c.c), which handles the
    full-size configs in well under a second on the host threads; the C port itself is pinned to the numpy oracle and the
    reference goldens by tests/test_oracle_c.py."""
    poses, rays = O.ba_unpack(x, fb.n_pose, ref_pose)
    r, U, gc, V, gl, cost = c_port.ba_fused(poses, rays, fb.cam_idx, fb.lm_idx, fb.obs_xy, u, v)
    np.testing.assert_allclose(out["residual"], r.ravel(), rtol=RTOL, atol=ATOL)
    assert abs(out["cost"] - cost) <= 1e-11 * cost
    Up = np.stack([out["U"][:, 0, 0], out["U"][:, 0, 1], out["U"][:, 0, 2], out["U"][:, 1, 1], out["U"][:, 1, 2], out["U"][:, 2, 2]], 1)
    Vp = np.stack([out["V"][:, 0, 0], out["V"][:, 0, 1], out["V"][:, 1, 1]], 1)
    np.testing.assert_allclose(Up[1:], U[1:], rtol=RTOL, atol=1e-12 * np.abs(U[1:]).max())
    np.testing.assert_allclose(out["gc"][1:], gc[1:], rtol=1e-8, atol=1e-11 * np.abs(gc[1:]).max())
    np.testing.assert_allclose(Vp, V, rtol=RTOL, atol=1e-12 * np.abs(V).max())
    np.testing.assert_allclose(out["gl"], gl, rtol=1e-8, atol=1e-11 * np.abs(gl).max())
    assert np.all(out["U"][0] == 0) and np.all(out["gc"][0] == 0)


def test_normal_equations_small_unsorted():
    """Reference pair-major order (not landmark-major, duplicates kept)."""
    d = load_golden("ba_residual.npz")
    points, src, dst, lmk, M = graph_from_npz(d)
    cam, lm, xy = synth.flatten_match_graph(points, src, dst, lmk)
    fb = synth.FlatBA(cam, lm, xy, d["ptz_gt"], d["rays_gt"], d["ptz_init"], d["rays_gt"])
    _check_normal_equations(fb, d["x1"], d["ptz_init"][0], d["uv"][0], d["uv"][1])


@pytest.mark.parametrize("n_kf,n_lm,n_obs", [(16, 500, 3000), (64, 20000, 300000), (300, 3000, 90000)])
def test_normal_equations_flat(n_kf, n_lm, n_obs):
    fb = synth.make_flat_ba(n_kf, n_lm, n_obs, seed=7)
    _check_normal_equations(fb, fb.x0(), fb.ptz_init[0], synth.PP_U, synth.PP_V)


def test_ragged_and_empty():
    """Landmarks without observations keep zero blocks; an empty problem is valid."""
    fb = synth.make_flat_ba(8, 200, 800, seed=9)
    keep = fb.lm_idx % 3 != 0
    fb2 = synth.FlatBA(fb.cam_idx[keep], fb.lm_idx[keep], fb.obs_xy[keep], fb.ptz_gt, fb.rays_gt, fb.ptz_init, fb.rays_init)
    prob = BA.BAProblem(fb2.n_pose, fb2.n_landmark, fb2.cam_idx, fb2.lm_idx, fb2.obs_xy, synth.PP_U, synth.PP_V)
    out = prob.normal_equations(fb2.x0(), fb2.ptz_init[0])
    assert np.all(out["V"][0::3] == 0) and np.all(out["gl"][0::3] == 0)
    _check_normal_equations(fb2, fb2.x0(), fb2.ptz_init[0], synth.PP_U, synth.PP_V)
    empty = BA.BAProblem(3, 5, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 2)), 640.0, 360.0)
    out = empty.normal_equations(np.zeros(3 * 2 + 10), np.array([1.0, 2.0, 2000.0]))
    assert out["cost"] == 0 and out["residual"].shape == (0,)
    with pytest.raises(Exception):
        BA.BAProblem(3, 5, np.array([7], np.int32), np.array([0], np.int32), np.zeros((1, 2)), 640.0, 360.0)


def test_full_size_cfg3_properties():
    """Config 3 (256 kf x 100k rays x 2M obs): size-independent checks.
    (1) residual pass == residual of the fused pass; (2) cost == 0.5*sum r^2; (3) trace identities
    sum_c U_pp(c) + U(fixed cam) == sum_l V_tt(l) (pan column = -theta column); (4) subset parity vs oracle."""
    fb = synth.make_flat_ba(256, 100000, 2000000, seed=1003)
    prob = BA.BAProblem(fb.n_pose, fb.n_landmark, fb.cam_idx, fb.lm_idx, fb.obs_xy, synth.PP_U, synth.PP_V)
    x = fb.x0()
    out = prob.normal_equations(x, fb.ptz_init[0])
    r2 = prob.residual(x, fb.ptz_init[0])
    np.testing.assert_array_equal(out["residual"], r2)
    assert abs(out["cost"] - 0.5 * np.dot(r2, r2)) <= 1e-11 * out["cost"]
    sel = np.nonzero(fb.cam_idx != 0)[0]
    poses, rays = O.ba_unpack(x, fb.n_pose, fb.ptz_init[0])
    Jc, Jr = O.jacobian_blocks_analytic(poses[fb.cam_idx, 0], poses[fb.cam_idx, 1], poses[fb.cam_idx, 2],
                                        rays[fb.lm_idx, 0], rays[fb.lm_idx, 1])
    # theta-theta trace over observations of free cameras equals the pan-pan trace of U
    tt_free = np.sum(Jr[sel, :, 0] ** 2)
    assert abs(out["U"][:, 0, 0].sum() - tt_free) <= 1e-10 * tt_free
    assert abs(out["V"][:, 0, 0].sum() - np.sum(Jr[:, :, 0] ** 2)) <= 1e-10 * tt_free
    sub = slice(0, 200000)
    ro = O.ba_residual_flat(poses, rays, fb.cam_idx[sub], fb.lm_idx[sub], fb.obs_xy[sub], synth.PP_U, synth.PP_V)
    np.testing.assert_allclose(out["residual"][:400000], ro.ravel(), rtol=RTOL, atol=ATOL)
    # (5) every block and every residual of the full-size problem against the C port
    check_against_c_port(out, fb, x, fb.ptz_init[0], synth.PP_U, synth.PP_V)
    prob.close()


# ---------------------------------------------------------------------------------------------------------------
# trust-region solve (Schur + Cholesky on the GPU) vs the reference's scipy call
# ---------------------------------------------------------------------------------------------------------------
TOL_DEG = np.degrees(1e-6)   # BASELINE.json: converged camera/ray parameters within 1e-6 rad and 1e-3 px focal


def _assert_params_close(x, xref, N):
    pe = np.abs(x[:3 * (N - 1)] - xref[:3 * (N - 1)]).reshape(-1, 3)
    assert pe[:, :2].max() < TOL_DEG, pe[:, :2].max()
    assert pe[:, 2].max() < 1e-3, pe[:, 2].max()
    assert np.abs(x[3 * (N - 1):] - xref[3 * (N - 1):]).max() < TOL_DEG


def test_solve_matches_reference_least_squares():
    d = load_golden("ba_solve.npz")
    points, src, dst, lmk, M = graph_from_npz(d)
    N = len(points)
    cam, lm, xy = synth.flatten_match_graph(points, src, dst, lmk)
    prob = BA.BAProblem(N, M, cam, lm, xy, d["uv"][0], d["uv"][1])
    # converged solution (reference call re-run with ftol = xtol = gtol = 1e-15)
    x, rep = prob.solve(d["x0"], d["ptz_init"][0], ftol=1e-15, xtol=1e-15, gtol=1e-15, max_nfev=100)
    _assert_params_close(x, d["x_tight"], N)
    assert abs(rep["cost"] - float(d["cost_tight"])) < 1e-9 * float(d["cost_tight"])
    # the reference's own stopping rule (ftol = 1e-4): same termination status and evaluation count
    x2, rep2 = prob.solve(d["x0"], d["ptz_init"][0], ftol=1e-4)
    assert rep2["status"] == int(d["status_asis"]) == 2
    assert rep2["nfev"] == int(d["nfev_asis"])
    _assert_params_close(x2, d["x_asis"], N)
    assert abs(rep2["cost"] - float(d["cost_asis"])) < 1e-7 * float(d["cost_asis"])
    prob.close()


def test_bundle_adjustment_core_matches_reference():
    d = load_golden("ba_solve.npz")
    points, src, dst, lmk, M = graph_from_npz(d)
    N = len(points)
    poses, landmarks, rep = BA.bundle_adjustment_core(points, src, dst, lmk, M, d["ptz_init"], d["uv"][0], d["uv"][1])
    x = np.concatenate([poses[1:].ravel(), landmarks.ravel()])
    _assert_params_close(x, d["x_asis"], N)
    np.testing.assert_array_equal(poses[0], d["ptz_init"][0])
    # drop-in entry point with an injected matching front-end
    graph = (points, None, points, src, dst, lmk, M)
    lms, kfs = BA.bundle_adjustment([None] * N, list(range(N)), 'sift', d["ptz_init"], np.zeros(3), np.eye(3),
                                    d["uv"][0], d["uv"][1], "/tmp", build_matching_graph=lambda *a: graph)
    np.testing.assert_allclose(lms, landmarks, rtol=0, atol=1e-12)
    assert len(kfs) == N and kfs[1].landmark_index.dtype == np.int32 and kfs[1].get_feature_num() > 0


def test_map_add_keyframe_with_ba():
    """scene_map.Map.add_keyframe_with_ba (scene_map.py:53-117): the last keyframe joins the map and all keyframes are
    bundle-adjusted on the GPU; the map is replaced by the result (same landmarks as the direct call)."""
    from ptz_slam_b200.scene_map import Map
    from ptz_slam_b200.key_frame import KeyFrame
    d = load_golden("ba_solve.npz")
    points, src, dst, lmk, M = graph_from_npz(d)
    N = len(points)
    graph = (points, None, points, src, dst, lmk, M)
    m = Map('sift', build_matching_graph=lambda *a: graph)
    u, v = d["uv"][0], d["uv"][1]
    kfs = [KeyFrame(None, i, np.zeros(3), np.eye(3), u, v, *d["ptz_init"][i]) for i in range(N)]
    m.add_first_keyframe(kfs[0])
    for kf in kfs[1:-1]:
        m.add_keyframe_without_ba(kf)
    landmarks, keyframes = m.add_keyframe_with_ba(kfs[-1], "/tmp")
    _, ref_landmarks, _ = BA.bundle_adjustment_core(points, src, dst, lmk, M, d["ptz_init"], u, v)
    np.testing.assert_allclose(landmarks, ref_landmarks, rtol=0, atol=1e-12)
    assert m.global_ray is landmarks and len(m.keyframe_list) == N and m.last_ba_seconds > 0
    assert all(kf.get_feature_num() > 0 for kf in m.keyframe_list)
    x = np.concatenate([np.array([[kf.pan, kf.tilt, kf.f] for kf in m.keyframe_list[1:]]).ravel(), landmarks.ravel()])
    _assert_params_close(x, d["x_asis"], N)


@pytest.mark.parametrize("n_kf,n_lm,n_obs", [(10, 300, 1500), (40, 4000, 40000)])
def test_solve_vs_oracle_trf(n_kf, n_lm, n_obs):
    """Same trust-region iteration as the scipy restatement (oracle.trf_solve) on a seeded flat problem."""
    fb = synth.make_flat_ba(n_kf, n_lm, n_obs, seed=21)
    u, v = synth.PP_U, synth.PP_V
    ref_pose = fb.ptz_init[0]
    prob = BA.BAProblem(fb.n_pose, fb.n_landmark, fb.cam_idx, fb.lm_idx, fb.obs_xy, u, v)
    # scipy's default tolerances: every accept/reject and termination decision is far above rounding noise, so the
    # evaluation count and the termination status must be identical to the restated scipy loop
    x, rep = prob.solve(fb.x0(), ref_pose, ftol=1e-8, xtol=1e-8, gtol=1e-8, max_nfev=60)
    if n_lm <= 300:
        fun = lambda z: O.ba_residual_flat(*O.ba_unpack(z, n_kf, ref_pose), fb.cam_idx, fb.lm_idx, fb.obs_xy, u, v).ravel()
        jac = lambda z: O.ba_jacobian_sparse(*O.ba_unpack(z, n_kf, ref_pose), fb.cam_idx, fb.lm_idx).toarray()
        ro = O.trf_solve(fun, jac, fb.x0(), ftol=1e-8, xtol=1e-8, gtol=1e-8, max_nfev=60)
        # 3 (xtol) vs 4 (xtol and ftol) only differ by the sign of a noise-level cost change at the last step
        assert rep["nfev"] == ro["nfev"] and rep["status"] in (2, 3, 4) and ro["status"] in (2, 3, 4)
        assert abs(rep["cost"] - ro["cost"]) < 1e-9 * ro["cost"]
        _assert_params_close(x, ro["x"], n_kf)
    # first-order optimality and recovery of the ground truth up to the noise level
    out = prob.normal_equations(x, ref_pose)
    assert np.abs(out["gc"][1:]).max() < 1e-5 * np.abs(out["U"]).max()
    assert rep["cost"] < rep["cost0"] * 1e-2
    pe = np.abs(x[:3 * (n_kf - 1)].reshape(-1, 3) - fb.ptz_gt[1:])
    assert pe[:, :2].max() < 0.05 and pe[:, 2].max() < 15.0
    prob.close()


@pytest.mark.parametrize("n", [1, 5, 31, 32, 33, 127, 128, 129, 255, 257, 765, 1000])
def test_dense_spd_solve_matches_numpy(n):
    """The cooperative Cholesky + block-inverse solve of the reduced camera system on its own, at orders around every block
    boundary (32-column panels, 64-row tiles, 128-row inverted blocks)."""
    import ctypes
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n))
    A = B @ B.T + n * np.eye(n)
    b = rng.standard_normal(n)
    ctx = _lib.get_context()
    x = np.empty(n)
    info = ctypes.c_int(-1)
    Ain = np.tril(A) + np.triu(np.full((n, n), np.nan), 1) if n > 1 else A.copy()   # the upper triangle must not be read
    Acm = np.ascontiguousarray(Ain.T)                  # column-major storage of Ain (kept alive across the call)
    ctx.check(ctx.lib.ptzba_dense_solve_spd(ctx.handle, n, _lib.ptr(Acm), _lib.ptr(b), _lib.ptr(x), ctypes.byref(info)))
    assert info.value == 0
    ref = np.linalg.solve(A, b)
    np.testing.assert_allclose(x, ref, rtol=1e-9, atol=1e-12 * np.abs(ref).max())


def test_dense_spd_solve_reports_indefinite():
    import ctypes
    n = 70
    A = np.eye(n)
    A[40, 40] = -1.0
    ctx = _lib.get_context()
    x, b = np.empty(n), np.ones(n)
    info = ctypes.c_int(0)
    ctx.check(ctx.lib.ptzba_dense_solve_spd(ctx.handle, n, _lib.ptr(A), _lib.ptr(b), _lib.ptr(x), ctypes.byref(info)))
    assert info.value == 33          # 1 + first row of the failing 32-column panel


def test_schur_per_landmark_route_equals_pair_list():
    """PTZBA_OPT_SCHUR_MODE: the per-landmark Schur formation (k_schur_pairs, the fallback above 65535 keyframes / 2^31 observation
    pairs) and the default keyframe-pair-major list (k_schur_pairlist) form the same reduced camera system: identical status and
    iteration counts, parameters equal far below the parity tolerance."""
    fb = synth.make_flat_ba(24, 1500, 18000, seed=11)
    res = []
    for mode in (_lib.SCHUR_PAIR_LIST, _lib.SCHUR_PER_LANDMARK):
        prob = BA.BAProblem(fb.n_pose, fb.n_landmark, fb.cam_idx, fb.lm_idx, fb.obs_xy, synth.PP_U, synth.PP_V)
        prob.set_option(_lib.OPT_SCHUR_MODE, mode)
        x, rep = prob.solve(fb.x0(), fb.ptz_init[0], ftol=1e-10, xtol=1e-10, gtol=1e-10)
        res.append((x, rep))
        prob.close()
    assert res[0][1]["status"] == res[1][1]["status"] and res[0][1]["nfev"] == res[1][1]["nfev"]
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-10, atol=1e-9)
    assert abs(res[0][1]["cost"] - res[1][1]["cost"]) <= 1e-10 * res[0][1]["cost"]
