"""CPU: the plain-C restatement (oracle/ptz_oracle_c.c, used as the CPU baseline) agrees with the pinned numpy oracle."""
import numpy as np

from conftest import load_golden, graph_from_npz
from oracle import ptz_oracle as O
from oracle import c_port
import ptz_slam_b200  # noqa: F401
from ptz_slam_b200 import synth


def test_c_port_matches_reference_residual_golden():
    d = load_golden("ba_residual.npz")
    points, src, dst, lmk, M = graph_from_npz(d)
    cam, lm, xy = synth.flatten_match_graph(points, src, dst, lmk)
    poses, rays = O.ba_unpack(d["x1"], len(points), d["ptz_init"][0])
    r = c_port.ba_residual(poses, rays, cam, lm, xy, d["uv"][0], d["uv"][1])
    np.testing.assert_allclose(r.ravel(), d["residual_x1"], rtol=1e-12, atol=1e-10)


def test_c_port_fused_matches_numpy_oracle():
    fb = synth.make_flat_ba(24, 2000, 20000, seed=5)
    poses, rays = O.ba_unpack(fb.x0(), fb.n_pose, fb.ptz_init[0])
    r, U, gc, V, gl, cost = O.ba_normal_equations(poses, rays, fb.cam_idx, fb.lm_idx, fb.obs_xy, synth.PP_U, synth.PP_V)
    for nt in (1, 0):
        r2, U2, gc2, V2, gl2, cost2 = c_port.ba_fused(poses, rays, fb.cam_idx, fb.lm_idx, fb.obs_xy, synth.PP_U, synth.PP_V,
                                                      n_threads=nt)
        np.testing.assert_allclose(r2, r, rtol=1e-11, atol=1e-10)
        assert abs(cost2 - cost) < 1e-11 * cost
        Uo = np.stack([U[:, 0, 0], U[:, 0, 1], U[:, 0, 2], U[:, 1, 1], U[:, 1, 2], U[:, 2, 2]], 1)
        np.testing.assert_allclose(U2[1:], Uo[1:], rtol=1e-10, atol=1e-12 * np.abs(Uo).max())
        assert np.all(U2[0] == 0)
        np.testing.assert_allclose(gc2[1:], gc[1:], rtol=1e-9, atol=1e-11 * np.abs(gc).max())
        Vo = np.stack([V[:, 0, 0], V[:, 0, 1], V[:, 1, 1]], 1)
        np.testing.assert_allclose(V2, Vo, rtol=1e-10, atol=1e-12 * np.abs(Vo).max())
        np.testing.assert_allclose(gl2, gl, rtol=1e-9, atol=1e-11 * np.abs(gl).max())
