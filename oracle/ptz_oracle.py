"""
ORACLE - TEST INFRASTRUCTURE ONLY.  Not product code.

CPU (numpy / pure-Python) restatement of the reference hot path of lulufa390/Pan-tilt-zoom-SLAM:
ray-landmark projection, reprojection residuals, Jacobians, normal equations, the EKF update and the
trust-region least-squares driver.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg may import this module; the product (pan-tilt-zoom-slam_b200/) never does.

Parity pinning: the reference ships no golden vectors or numeric assertions, so this
oracle is pinned against *outputs of the reference itself*: tests/golden/make_golden.py imports the
unmodified reference modules from /root/reference/slam_system (with the three absent imports stubbed),
runs them on seeded inputs and commits the results as tests/golden/*.npz.  tests/test_oracle.py checks
every function below against those files.

Each function cites the reference lines it follows (paths relative to /root/reference/slam_system).
All angles are DEGREES, focal length / pixels in PIXELS, everything float64, as in the reference.
"""
import math

import numpy as np

# ---------------------------------------------------------------------------------------------
# projection / back-projection
# ---------------------------------------------------------------------------------------------


def project_ray(pan, tilt, f, u, v, ray, disp=None):
    """ptz_camera.py:191-210 (PTZCamera.project_ray), scalar, math.* like the reference."""
    theta = math.radians(ray[0])
    phi = math.radians(ray[1])
    p = math.radians(pan)
    t = math.radians(tilt)
    lam = np.zeros(6) if disp is None else np.asarray(disp, dtype=np.float64)
    K = np.array([[f, 0, u], [0, f, v], [0, 0, 1]], dtype=np.float64)           # :55-63
    tilt_rot = np.array([[1, 0, 0], [0, math.cos(t), math.sin(t)], [0, -math.sin(t), math.cos(t)]])  # :95-104
    pan_rot = np.array([[math.cos(p), 0, -math.sin(p)], [0, 1, 0], [math.sin(p), 0, math.cos(p)]])   # :83-93
    d = np.array([lam[0] + lam[3] * f, lam[1] + lam[4] * f, lam[2] + lam[5] * f])                     # :106-115
    ray_p = np.array([math.tan(theta), -math.tan(phi) * math.sqrt(math.tan(theta) * math.tan(theta) + 1), 1])  # :205
    img = np.dot(K, np.dot(np.dot(tilt_rot, pan_rot), ray_p) + d)                                     # :207
    assert img[2] != 0.0
    return img[0] / img[2], img[1] / img[2]


def project_rays_vec(pan, tilt, f, u, v, rays, disp=None):
    """Vectorised ptz_camera.py:191-210 over rays[n,2]; pan/tilt/f may be scalars or [n] arrays.

    Returns x[n], y[n], q2[n] (q2 is the homogeneous depth the reference asserts non-zero, :208).
    """
    rays = np.asarray(rays, dtype=np.float64).reshape(-1, 2)
    th = np.radians(rays[:, 0])
    ph = np.radians(rays[:, 1])
    p = np.radians(pan)
    t = np.radians(tilt)
    lam = np.zeros(6) if disp is None else np.asarray(disp, dtype=np.float64)
    tx = np.tan(th)
    d0 = tx
    d1 = -np.tan(ph) * np.sqrt(tx * tx + 1.0)
    d2 = np.ones_like(tx)
    a0 = np.cos(p) * d0 - np.sin(p) * d2
    a1 = d1
    a2 = np.sin(p) * d0 + np.cos(p) * d2
    q0 = a0 + (lam[0] + lam[3] * f)
    q1 = np.cos(t) * a1 + np.sin(t) * a2 + (lam[1] + lam[4] * f)
    q2 = -np.sin(t) * a1 + np.cos(t) * a2 + (lam[2] + lam[5] * f)
    # K * q  (third row of K is [0,0,1]) -> (f*q0 + u*q2)/q2
    return (f * q0 + u * q2) / q2, (f * q1 + v * q2) / q2, q2


def project_rays(pan, tilt, f, u, v, rays, height=0, width=0, disp=None):
    """ptz_camera.py:212-234 (PTZCamera.project_rays): returns (points[m,2], index[m]).

    With height/width given only strictly-inside points are kept (0 < x < W and 0 < y < H, :226);
    the index array is float64-typed like the reference's np.ndarray([0]) growth (:221,228).
    Without them all points are returned and the index is empty (:229-232).
    """
    rays = np.asarray(rays, dtype=np.float64).reshape(-1, 2)
    x, y, _ = project_rays_vec(pan, tilt, f, u, v, rays, disp)
    pts = np.stack([x, y], axis=1)
    if height != 0 and width != 0:
        keep = (0 < x) & (x < width) & (0 < y) & (y < height)
        return pts[keep], np.nonzero(keep)[0].astype(np.float64)
    return pts, np.ndarray([0])


def from_ray_to_image(u, v, f, c_p, c_t, p, t):
    """transformation.py:99-135 (TransFunction.from_ray_to_image), verbatim algebra, scalar."""
    pan = math.radians(p)
    tilt = math.radians(t)
    cp = math.radians(c_p)
    ct = math.radians(c_t)
    tp = math.tan(pan)
    tt = math.tan(tilt)
    sq = math.sqrt(tp * tp + 1)
    num_x = tp * math.cos(cp) - math.sin(cp)
    den = tp * math.sin(cp) * math.cos(ct) + tt * sq * math.sin(ct) + math.cos(ct) * math.cos(cp)
    relative_pan = math.atan(num_x / den)                                                  # :119-122
    num_y = -(tp * math.sin(ct) * math.sin(cp) - tt * sq * math.cos(ct) + math.sin(ct) * math.cos(cp))
    relative_tilt = math.atan(num_y / math.sqrt(num_x ** 2 + den ** 2))                   # :124-130
    dx = f * math.tan(relative_pan)                                                       # :132
    x = dx + u
    y = -math.sqrt(f * f + dx * dx) * math.tan(relative_tilt) + v                         # :134
    return x, y


def from_ray_to_image_vec(u, v, f, c_p, c_t, p, t):
    """Vectorised transformation.py:99-135 (same algebra incl. the sqrt(f^2+dx^2) form of y)."""
    pan = np.radians(p)
    tilt = np.radians(t)
    cp = np.radians(c_p)
    ct = np.radians(c_t)
    tp = np.tan(pan)
    tt = np.tan(tilt)
    sq = np.sqrt(tp * tp + 1)
    num_x = tp * np.cos(cp) - np.sin(cp)
    den = tp * np.sin(cp) * np.cos(ct) + tt * sq * np.sin(ct) + np.cos(ct) * np.cos(cp)
    num_y = -(tp * np.sin(ct) * np.sin(cp) - tt * sq * np.cos(ct) + np.sin(ct) * np.cos(cp))
    # tan(atan(a)) == a up to rounding; keep the atan/tan round trip so rounding matches the reference
    dx = f * np.tan(np.arctan(num_x / den))
    x = dx + u
    y = -np.sqrt(f * f + dx * dx) * np.tan(np.arctan(num_y / np.sqrt(num_x ** 2 + den ** 2))) + v
    return x, y


def back_project_to_ray(pan, tilt, f, u, v, x, y, disp=None):
    """ptz_camera.py:287-312 (PTZCamera.back_project_to_ray), scalar."""
    p = math.radians(pan)
    t = math.radians(tilt)
    lam = np.zeros(6) if disp is None else np.asarray(disp, dtype=np.float64)
    d = np.array([lam[0] + lam[3] * f, lam[1] + lam[4] * f, lam[2] + lam[5] * f])
    K = np.array([[f, 0, u], [0, f, v], [0, 0, 1]], dtype=np.float64)
    tilt_rot = np.array([[1, 0, 0], [0, math.cos(t), math.sin(t)], [0, -math.sin(t), math.cos(t)]])
    pan_rot = np.array([[math.cos(p), 0, -math.sin(p)], [0, 1, 0], [math.sin(p), 0, math.cos(p)]])
    R_inv = np.linalg.inv(np.dot(tilt_rot, pan_rot))                                       # :305-306
    x3, y3, z3 = np.dot(R_inv, np.dot(np.linalg.inv(K), np.array([x, y, 1.0])) - d)        # :307
    theta = math.atan(x3 / z3)                                                             # :309
    phi = math.atan(-y3 / math.sqrt(x3 * x3 + z3 * z3))                                    # :310
    return math.degrees(theta), math.degrees(phi)


def back_project_to_rays_vec(pan, tilt, f, u, v, pts, disp=None):
    """Vectorised ptz_camera.py:287-325; pts[n,2] -> rays[n,2] (R^-1 = R^T, K^-1 closed form)."""
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    p = np.radians(pan)
    t = np.radians(tilt)
    lam = np.zeros(6) if disp is None else np.asarray(disp, dtype=np.float64)
    q0 = (pts[:, 0] - u) / f - (lam[0] + lam[3] * f)
    q1 = (pts[:, 1] - v) / f - (lam[1] + lam[4] * f)
    q2 = 1.0 - (lam[2] + lam[5] * f)
    a0 = q0
    a1 = np.cos(t) * q1 - np.sin(t) * q2
    a2 = np.sin(t) * q1 + np.cos(t) * q2
    d0 = np.cos(p) * a0 + np.sin(p) * a2
    d1 = a1
    d2 = -np.sin(p) * a0 + np.cos(p) * a2
    theta = np.degrees(np.arctan(d0 / d2))
    phi = np.degrees(np.arctan(-d1 / np.sqrt(d0 * d0 + d2 * d2)))
    return np.stack([theta, phi], axis=1)


def from_image_to_ray(u, v, f, c_p, c_t, x, y):
    """transformation.py:137-175 (TransFunction.from_image_to_ray), scalar."""
    pan = math.radians(c_p)
    tilt = math.radians(c_t)
    theta_skim = math.atan((x - u) / f)                                                    # :154
    phi_skim = math.atan((y - v) / (-f * math.sqrt(1 + ((x - u) / f) ** 2)))               # :155
    x3s = math.tan(theta_skim)
    y3s = -math.tan(phi_skim) * math.sqrt(math.tan(theta_skim) ** 2 + 1)
    rot = np.dot(np.array([[1, 0, 0], [0, math.cos(tilt), math.sin(tilt)], [0, -math.sin(tilt), math.cos(tilt)]]),
                 np.array([[math.cos(pan), 0, -math.sin(pan)], [0, 1, 0], [math.sin(pan), 0, math.cos(pan)]]))
    x3, y3, z3 = np.dot(np.linalg.inv(rot), np.array([x3s, y3s, 1]))                       # :168-170
    theta = math.atan(x3 / z3)
    phi = math.atan(-y3 / math.sqrt(x3 * x3 + z3 * z3))
    return math.degrees(theta), math.degrees(phi)


# ---------------------------------------------------------------------------------------------
# Jacobians
# ---------------------------------------------------------------------------------------------
DELTA_ANGLE = 0.001   # ptz_slam.py:87
DELTA_F = 0.1         # ptz_slam.py:88


def h_jacobian_blocks_fd(pan, tilt, f, u, v, rays, disp=None):
    """Central-difference blocks of ptz_slam.py:95-136: returns Jc[n,2,3] (cols pan,tilt,f), Jr[n,2,2] (cols theta,phi)."""
    rays = np.asarray(rays, dtype=np.float64).reshape(-1, 2)
    n = len(rays)
    Jc = np.zeros((n, 2, 3))
    Jr = np.zeros((n, 2, 2))

    def P(pp, tt, ff, rr):
        x, y, _ = project_rays_vec(pp, tt, ff, u, v, rr, disp)
        return np.stack([x, y], axis=1)

    da, df = DELTA_ANGLE, DELTA_F
    Jc[:, :, 0] = (P(pan + da, tilt, f, rays) - P(pan - da, tilt, f, rays)) / (2 * da)     # :96-100,120,124
    Jc[:, :, 1] = (P(pan, tilt + da, f, rays) - P(pan, tilt - da, f, rays)) / (2 * da)     # :102-106,121,125
    Jc[:, :, 2] = (P(pan, tilt, f + df, rays) - P(pan, tilt, f - df, rays)) / (2 * df)     # :108-112,122,126
    e0 = np.array([da, 0.0])
    e1 = np.array([0.0, da])
    Jr[:, :, 0] = (P(pan, tilt, f, rays + e0) - P(pan, tilt, f, rays - e0)) / (2 * da)     # :115-116,132,135
    Jr[:, :, 1] = (P(pan, tilt, f, rays + e1) - P(pan, tilt, f, rays - e1)) / (2 * da)     # :117-118,133,136
    return Jc, Jr


def compute_h_jacobian(pan, tilt, f, u, v, rays, disp=None):
    """ptz_slam.py:73-138 (PtzSlam.compute_h_jacobian): dense H[2n, 3+2n] from central differences."""
    rays = np.asarray(rays, dtype=np.float64).reshape(-1, 2)
    n = len(rays)
    Jc, Jr = h_jacobian_blocks_fd(pan, tilt, f, u, v, rays, disp)
    H = np.zeros((2 * n, 3 + 2 * n))
    for i in range(n):
        H[2 * i:2 * i + 2, 0:3] = Jc[i]
        H[2 * i:2 * i + 2, 3 + 2 * i:5 + 2 * i] = Jr[i]
    return H


def jacobian_blocks_analytic(pan, tilt, f, theta, phi):
    """Analytic d(x,y)/d(pan,tilt,f) and d(x,y)/d(theta,phi), disp = 0.

    Derived from ptz_camera.py:191-210; derivatives are per DEGREE for the angles.  Broadcasts.
    Returns Jc[...,2,3], Jr[...,2,2].
    """
    k = math.pi / 180.0
    p = np.radians(pan); t = np.radians(tilt)
    th = np.radians(theta); ph = np.radians(phi)
    sgn = np.sign(np.cos(th))                  # sqrt(tan^2+1) = |sec|, ptz_camera.py:205
    a = th - p
    T = np.tan(ph) * sgn
    S = 1.0 + np.tan(ph) ** 2
    sa, ca, st, ct = np.sin(a), np.cos(a), np.sin(t), np.cos(t)
    Ny = -ct * T + st * ca
    z = st * T + ct * ca
    px, py = sa / z, Ny / z
    fz = f / z
    dxa = fz * (ca + ct * sa * px)
    dya = fz * sa * (-st + ct * py)
    dxt = f * px * py
    dyt = f * (1.0 + py * py)
    dxp = -fz * px * st * S * sgn
    dyp = -fz * (ct + st * py) * S * sgn
    shape = np.broadcast(dxa, dxt, dxp).shape
    Jc = np.empty(shape + (2, 3)); Jr = np.empty(shape + (2, 2))
    Jc[..., 0, 0] = -k * dxa; Jc[..., 1, 0] = -k * dya
    Jc[..., 0, 1] = k * dxt;  Jc[..., 1, 1] = k * dyt
    Jc[..., 0, 2] = px;       Jc[..., 1, 2] = py
    Jr[..., 0, 0] = k * dxa;  Jr[..., 1, 0] = k * dya
    Jr[..., 0, 1] = k * dxp;  Jr[..., 1, 1] = k * dyp
    return Jc, Jr


# ---------------------------------------------------------------------------------------------
# EKF
# ---------------------------------------------------------------------------------------------
def get_overlap_index(index1, index2):
    """util.py:75-96: positions of the shared values of two ascending arrays (two-pointer merge)."""
    o1, o2 = [], []
    a = b = 0
    while a < len(index1) and b < len(index2):
        if index1[a] == index2[b]:
            o1.append(a); o2.append(b); a += 1; b += 1
        elif index1[a] < index2[b]:
            a += 1
        else:
            b += 1
    return np.asarray(o1, dtype=np.int64), np.asarray(o2, dtype=np.int64)


class EkfState:
    """The slice of PtzSlam state the EKF touches (ptz_slam.py:29-71)."""

    def __init__(self, rays, ptz, u, v, disp=None, angle_var=0.001, f_var=1.0, observe_var=0.1):
        self.rays = np.array(rays, dtype=np.float64).reshape(-1, 2)
        n = len(self.rays)
        self.state_cov = angle_var * np.eye(3 + 2 * n)       # ptz_slam.py:199
        self.state_cov[2, 2] = f_var                          # :200
        self.ptz = np.array(ptz, dtype=np.float64)
        self.velocity = np.zeros(3)
        self.u, self.v, self.disp = u, v, disp
        self.angle_var, self.f_var, self.observe_var = angle_var, f_var, observe_var


def ekf_predict(s):
    """ptz_slam.py:418-426: constant-velocity pose prediction and pose-block process noise."""
    s.ptz = s.ptz + s.velocity
    s.state_cov[0:3, 0:3] += 5 * np.diag([s.angle_var, s.angle_var, s.f_var])


def ekf_update(s, observed_keypoints, observed_keypoint_index, height, width):
    """ptz_slam.py:210-289 (PtzSlam.ekf_update) incl. the theta-theta / phi-phi-only write-back (:281-289)."""
    pred_pts, pred_idx = project_rays(s.ptz[0], s.ptz[1], s.ptz[2], s.u, s.v, s.rays, height, width, s.disp)  # :223
    o1, o2 = get_overlap_index(observed_keypoint_index, pred_idx)                           # :228
    y_k = (np.asarray(observed_keypoints)[o1] - pred_pts[o2]).flatten()                     # :229-230
    matched = np.asarray(observed_keypoint_index)[o1].astype(np.int64)                      # :233
    n = len(matched)
    idx = np.concatenate([[0, 1, 2], np.stack([3 + 2 * matched, 4 + 2 * matched], 1).ravel()]).astype(np.int64)  # :238-245
    P = s.state_cov[idx][:, idx]                                                            # :246
    H = compute_h_jacobian(s.ptz[0], s.ptz[1], s.ptz[2], s.u, s.v, s.rays[matched], s.disp)  # :250-254
    S = H @ P @ H.T + s.observe_var * np.eye(2 * n)                                         # :256-257
    K = P @ H.T @ np.linalg.inv(S)                                                          # :259
    ky = K @ y_k                                                                            # :262
    s.ptz = s.ptz + ky[0:3]                                                                 # :266-268
    s.velocity = ky[0:3].copy()                                                             # :273
    s.rays[matched] += ky[3:].reshape(-1, 2)                                                # :276-277
    Pn = (np.eye(3 + 2 * n) - K @ H) @ P                                                    # :280
    s.state_cov[0:3, 0:3] = Pn[0:3, 0:3]                                                    # :281
    rt = 3 + 2 * matched
    s.state_cov[np.ix_(rt, rt)] = Pn[3::2, 3::2]                                            # :288
    s.state_cov[np.ix_(rt + 1, rt + 1)] = Pn[4::2, 4::2]                                    # :289
    return matched


# ---------------------------------------------------------------------------------------------
# bundle adjustment residual / Jacobian / normal equations (flat observation form)
# ---------------------------------------------------------------------------------------------
def ba_residual_lists(x, n_pose, n_landmark, n_residual, keypoints, src_pt_index, dst_pt_index, landmark_index,
                      u, v, reference_pose):
    """bundle_adjustment.py:25-106 (_compute_residual), pure-Python loops, scalar from_ray_to_image."""
    x0 = np.zeros(n_pose * 3 + n_landmark * 2)
    x0[0:3] = reference_pose                                                                # :57-59
    x0[3:] = x
    ls = n_pose * 3
    res = np.empty(n_residual)
    k = 0
    for i in range(n_pose):
        for j in range(n_pose):
            for i1, i2, i3 in zip(src_pt_index[i][j], dst_pt_index[i][j], landmark_index[i][j]):   # :73
                th, ph = x0[ls + 2 * i3], x0[ls + 2 * i3 + 1]
                px1, py1 = from_ray_to_image(u, v, x0[3 * i + 2], x0[3 * i], x0[3 * i + 1], th, ph)
                px2, py2 = from_ray_to_image(u, v, x0[3 * j + 2], x0[3 * j], x0[3 * j + 1], th, ph)
                res[k] = px1 - keypoints[i][i1][0]; res[k + 1] = py1 - keypoints[i][i1][1]          # :83-89
                res[k + 2] = px2 - keypoints[j][i2][0]; res[k + 3] = py2 - keypoints[j][i2][1]      # :93-98
                k += 4
    assert k == n_residual
    return res


def ba_unpack(x, n_pose, reference_pose):
    """x (poses 1..N-1, then landmarks) -> poses[N,3], rays[M,2]  (bundle_adjustment.py:57-59,204-208)."""
    poses = np.concatenate([np.asarray(reference_pose, dtype=np.float64), x[:3 * (n_pose - 1)]]).reshape(n_pose, 3)
    rays = np.asarray(x[3 * (n_pose - 1):], dtype=np.float64).reshape(-1, 2)
    return poses, rays


def ba_residual_flat(poses, rays, cam_idx, lm_idx, obs_xy, u, v):
    """Vectorised bundle_adjustment.py:78-98 over a flat observation list: r[n_obs,2] = proj - obs."""
    c = poses[cam_idx]
    l = rays[lm_idx]
    x, y = from_ray_to_image_vec(u, v, c[:, 2], c[:, 0], c[:, 1], l[:, 0], l[:, 1])
    return np.stack([x - obs_xy[:, 0], y - obs_xy[:, 1]], axis=1)


def ba_normal_equations(poses, rays, cam_idx, lm_idx, obs_xy, u, v):
    """J^T J / J^T r blocks the CUDA fused pass assembles (after A7).

    Returns r[n_obs,2], U[N,3,3], gc[N,3], V[M,2,2], gl[M,2], cost = 0.5*sum r^2.
    Camera 0's blocks are returned too (the solver drops them: reference pose is fixed,
    bundle_adjustment.py:57-59,197).
    """
    N, M = len(poses), len(rays)
    r = ba_residual_flat(poses, rays, cam_idx, lm_idx, obs_xy, u, v)
    c = poses[cam_idx]; l = rays[lm_idx]
    Jc, Jr = jacobian_blocks_analytic(c[:, 0], c[:, 1], c[:, 2], l[:, 0], l[:, 1])
    U = np.zeros((N, 3, 3)); gc = np.zeros((N, 3)); V = np.zeros((M, 2, 2)); gl = np.zeros((M, 2))
    np.add.at(U, cam_idx, np.einsum('nki,nkj->nij', Jc, Jc))
    np.add.at(gc, cam_idx, np.einsum('nki,nk->ni', Jc, r))
    np.add.at(V, lm_idx, np.einsum('nki,nkj->nij', Jr, Jr))
    np.add.at(gl, lm_idx, np.einsum('nki,nk->ni', Jr, r))
    return r, U, gc, V, gl, 0.5 * float(np.sum(r * r))


def ba_jacobian_sparse(poses, rays, cam_idx, lm_idx):
    """Analytic sparse J (scipy CSR) w.r.t. x = [poses 1..N-1, landmarks]; rows ordered (obs, x|y)."""
    import scipy.sparse as sp
    N, M = len(poses), len(rays)
    n_obs = len(cam_idx)
    c = poses[cam_idx]; l = rays[lm_idx]
    Jc, Jr = jacobian_blocks_analytic(c[:, 0], c[:, 1], c[:, 2], l[:, 0], l[:, 1])
    rows, cols, vals = [], [], []
    o = np.arange(n_obs)
    free = cam_idx > 0
    for a in range(2):
        for b in range(3):
            rows.append(2 * o[free] + a); cols.append(3 * (cam_idx[free].astype(np.int64) - 1) + b); vals.append(Jc[free, a, b])
        for b in range(2):
            rows.append(2 * o + a); cols.append(3 * (N - 1) + 2 * lm_idx.astype(np.int64) + b); vals.append(Jr[:, a, b])
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                         shape=(2 * n_obs, 3 * (N - 1) + 2 * M))


def ba_lm_iteration_sparse(x, n_pose, reference_pose, cam_idx, lm_idx, obs_xy, u, v, alpha, fused=None, residual=None):
    """One Levenberg-Marquardt iteration at fixed damping alpha on the CPU - what ptzba_ba_lm_iteration does on the GPU, with
    the tools a CPU user of the reference has (numpy / scipy.sparse / LAPACK): blocks of J^T J and J^T r (bundle_adjustment.py
    :25-106 differentiated analytically), x_scale='jac' column norms D (least_squares call at bundle_adjustment.py:200-202),
    the step (J^T J + alpha D^2) d = -J^T r through the Schur complement over the landmarks (sparse products, dense Cholesky of
    the reduced camera system), the predicted reduction and the cost at the trial point.
    `fused` / `residual`: optional faster evaluators with the signatures of oracle.c_port.ba_fused / ba_residual (bench.py passes
    the multi-threaded C port); default: the numpy restatement.
    Returns (step[3(N-1)+2M], predicted_reduction, cost_trial)."""
    import scipy.sparse as sp
    import scipy.linalg as sla
    cam_idx = np.asarray(cam_idx); lm_idx = np.asarray(lm_idx)
    poses, rays = ba_unpack(np.asarray(x, dtype=np.float64), n_pose, reference_pose)
    N, M = len(poses), len(rays)
    nc = 3 * (N - 1)
    if fused is None:
        r, U, gc, V, gl, cost = ba_normal_equations(poses, rays, cam_idx, lm_idx, obs_xy, u, v)
        V3 = np.stack([V[:, 0, 0], V[:, 0, 1], V[:, 1, 1]], axis=1)
        Ufull = U
    else:
        r, U6, gc, V3, gl, cost = fused(poses, rays, cam_idx, lm_idx, obs_xy, u, v, want_residual=False)
        Ufull = np.empty((N, 3, 3))
        iu = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
        for k, (a, b) in enumerate(iu):
            Ufull[:, a, b] = U6[:, k]; Ufull[:, b, a] = U6[:, k]
    c = poses[cam_idx]; l = rays[lm_idx]
    Jc, Jr = jacobian_blocks_analytic(c[:, 0], c[:, 1], c[:, 2], l[:, 0], l[:, 1])
    free = cam_idx > 0
    Wb = np.einsum('nki,nkj->nij', Jc[free], Jr[free])                      # per observation: J_c^T J_l (3 x 2)
    cf = cam_idx[free].astype(np.int64) - 1; lf = lm_idx[free].astype(np.int64)
    rows = (3 * cf[:, None, None] + np.arange(3)[None, :, None] + np.zeros((1, 1, 2), np.int64)).ravel()
    cols = (2 * lf[:, None, None] + np.zeros((1, 3, 1), np.int64) + np.arange(2)[None, None, :]).ravel()
    W = sp.csr_matrix((Wb.ravel(), (rows, cols)), shape=(nc, 2 * M))
    # x_scale='jac': D = column norms of J (zero columns -> 1, _lsq/common.py:compute_jac_scale)
    Dc2 = np.stack([Ufull[1:, 0, 0], Ufull[1:, 1, 1], Ufull[1:, 2, 2]], axis=1).ravel()
    Dl2 = np.stack([V3[:, 0], V3[:, 2]], axis=1).ravel()
    Dc2 = np.where(Dc2 > 0, Dc2, 1.0); Dl2 = np.where(Dl2 > 0, Dl2, 1.0)
    a = V3[:, 0] + alpha * Dl2[0::2]; b = V3[:, 1]; d = V3[:, 2] + alpha * Dl2[1::2]
    det = a * d - b * b
    i00, i01, i11 = d / det, -b / det, a / det                               # (V + alpha D^2)^-1, 2 x 2 blocks
    k = np.arange(M)
    Vinv = sp.csr_matrix((np.concatenate([i00, i01, i01, i11]),
                          (np.concatenate([2 * k, 2 * k, 2 * k + 1, 2 * k + 1]), np.concatenate([2 * k, 2 * k + 1, 2 * k, 2 * k + 1]))),
                         shape=(2 * M, 2 * M))
    X = W @ Vinv
    S = -(X @ W.T).toarray()
    for i in range(N - 1):
        S[3 * i:3 * i + 3, 3 * i:3 * i + 3] += Ufull[i + 1]
    S[np.arange(nc), np.arange(nc)] += alpha * Dc2
    g_c = gc[1:].ravel(); g_l = gl.ravel()
    rhs = -(g_c - X @ g_l)
    cf_ = sla.cho_factor(S, lower=True, check_finite=False)
    dc = sla.cho_solve(cf_, rhs, check_finite=False)
    dl = Vinv @ (-g_l - W.T @ dc)
    step = np.concatenate([dc, dl])
    # ||J d||^2 from the per-observation blocks
    dcf = np.concatenate([np.zeros(3), dc]).reshape(N, 3)[cam_idx]
    dlf = dl.reshape(M, 2)[lm_idx]
    Jd = np.einsum('nki,ni->nk', Jc, dcf) + np.einsum('nki,ni->nk', Jr, dlf)
    pred = -(float(np.dot(np.concatenate([g_c, g_l]), step)) + 0.5 * float(np.sum(Jd * Jd)))
    pt, rt = ba_unpack(np.asarray(x, dtype=np.float64) + step, n_pose, reference_pose)
    rr = ba_residual_flat(pt, rt, cam_idx, lm_idx, obs_xy, u, v) if residual is None else residual(pt, rt, cam_idx, lm_idx, obs_xy, u, v)
    return step, pred, 0.5 * float(np.sum(rr * rr))


# ---------------------------------------------------------------------------------------------
# trust-region least squares driver
# Third-party algorithm: scipy.optimize.least_squares(method='trf', x_scale='jac', tr_solver='exact'),
# README pins SciPy 0.18.1; restated from scipy/optimize/_lsq/{trf.py:trf_no_bounds, common.py} (1.18.1 here).
# Call site in the reference: bundle_adjustment.py:200-202.
# ---------------------------------------------------------------------------------------------
EPS = np.finfo(float).eps


def _solve_tr_normal(A, g, Delta, alpha0, rtol=0.01, max_iter=10):
    """More's trust-region root finding (_lsq/common.py:solve_lsq_trust_region) on the normal matrix A = J_h^T J_h.

    scipy works from an SVD of J_h; here the same secular equation phi(alpha) = ||p(alpha)|| - Delta is evaluated
    through Cholesky factorisations of A + alpha*I (the MINPACK form the docstring of the scipy routine refers to).
    Returns p, alpha, n_iter.
    """
    n = len(g)
    I = np.eye(n)

    def p_of(alpha):
        L = np.linalg.cholesky(A + alpha * I)
        p = -np.linalg.solve(L.T, np.linalg.solve(L, g))
        q = np.linalg.solve(L, p)
        pn = np.linalg.norm(p)
        return p, pn, -(q @ q) / pn      # phi' = -p^T (A+aI)^-1 p / ||p||

    full_rank = True
    try:
        p, pn, dphi = p_of(0.0)
        if not np.all(np.isfinite(p)):
            full_rank = False
    except np.linalg.LinAlgError:
        full_rank = False
    if full_rank and pn <= Delta:
        return p, 0.0, 0
    alpha_upper = np.linalg.norm(g) / Delta
    alpha_lower = -(pn - Delta) / dphi if full_rank else 0.0
    if alpha0 is None or (not full_rank and alpha0 == 0):
        alpha = max(0.001 * alpha_upper, (alpha_lower * alpha_upper) ** 0.5)
    else:
        alpha = alpha0
    it = 0
    for it in range(max_iter):
        if alpha < alpha_lower or alpha > alpha_upper:
            alpha = max(0.001 * alpha_upper, (alpha_lower * alpha_upper) ** 0.5)
        p, pn, dphi = p_of(alpha)
        phi = pn - Delta
        if phi < 0:
            alpha_upper = alpha
        ratio = phi / dphi
        alpha_lower = max(alpha_lower, alpha - ratio)
        alpha -= (phi + Delta) * ratio / Delta
        if abs(phi) < rtol * Delta:
            break
    p, pn, _ = p_of(alpha)
    p *= Delta / pn
    return p, alpha, it + 1


def trf_solve(fun, jac, x0, ftol=1e-8, xtol=1e-8, gtol=1e-8, max_nfev=None):
    """Dense restatement of _lsq/trf.py:trf_no_bounds with x_scale='jac', no loss function.

    `jac(x)` returns a dense ndarray.  Returns dict(x, cost, nfev, njev, status, nit).
    """
    x = np.array(x0, dtype=np.float64)
    f = fun(x); nfev = 1
    J = jac(x); njev = 1
    cost = 0.5 * f @ f
    g = J.T @ f
    scale_inv = np.sum(J ** 2, axis=0) ** 0.5
    scale_inv[scale_inv == 0] = 1
    scale = 1 / scale_inv
    Delta = np.linalg.norm(x * scale_inv)
    if Delta == 0:
        Delta = 1.0
    if max_nfev is None:
        max_nfev = x.size * 100
    alpha = 0.0
    status = None
    nit = 0
    while True:
        g_norm = np.linalg.norm(g, ord=np.inf)
        if g_norm < gtol:
            status = 1
        if status is not None or nfev == max_nfev:
            break
        d = scale
        g_h = d * g
        J_h = J * d
        A = J_h.T @ J_h
        actual_reduction = -1
        while actual_reduction <= 0 and nfev < max_nfev:
            step_h, alpha, _ = _solve_tr_normal(A, g_h, Delta, alpha)
            Js = J_h @ step_h
            predicted_reduction = -(0.5 * Js @ Js + g_h @ step_h)
            step = d * step_h
            x_new = x + step
            f_new = fun(x_new); nfev += 1
            step_h_norm = np.linalg.norm(step_h)
            if not np.all(np.isfinite(f_new)):
                Delta = 0.25 * step_h_norm
                continue
            cost_new = 0.5 * f_new @ f_new
            actual_reduction = cost - cost_new
            if predicted_reduction > 0:
                ratio = actual_reduction / predicted_reduction
            elif predicted_reduction == actual_reduction == 0:
                ratio = 1
            else:
                ratio = 0
            Delta_new = Delta
            if ratio < 0.25:
                Delta_new = 0.25 * step_h_norm
            elif ratio > 0.75 and step_h_norm > 0.95 * Delta:
                Delta_new = Delta * 2.0
            step_norm = np.linalg.norm(step)
            ftol_ok = actual_reduction < ftol * cost and ratio > 0.25
            xtol_ok = step_norm < xtol * (xtol + np.linalg.norm(x))
            status = 4 if (ftol_ok and xtol_ok) else 2 if ftol_ok else 3 if xtol_ok else None
            if status is not None:
                break
            alpha *= Delta / Delta_new
            Delta = Delta_new
        if actual_reduction > 0:
            x = x_new; f = f_new; cost = cost_new
            J = jac(x); njev += 1
            g = J.T @ f
            scale_inv = np.maximum(np.sum(J ** 2, axis=0) ** 0.5, scale_inv)
            scale = 1 / scale_inv
        nit += 1
    return dict(x=x, cost=cost, nfev=nfev, njev=njev, status=0 if status is None else status, nit=nit)


# ---------------------------------------------------------------------------------------------
# relocalisation: 3-parameter pose refinement on fixed rays
# ---------------------------------------------------------------------------------------------
def reloc_residual(pose, rays, points, u, v):
    """relocalization.py:22-40 / nearest_neighbor.py:65-85: [2n] (projection - point) through from_ray_to_image."""
    rays = np.asarray(rays, dtype=np.float64).reshape(-1, 2)
    points = np.asarray(points, dtype=np.float64).reshape(-1, 2)
    x, y = from_ray_to_image_vec(u, v, pose[2], pose[0], pose[1], rays[:, 0], rays[:, 1])
    return np.stack([x - points[:, 0], y - points[:, 1]], axis=1).reshape(-1)


def reloc_refine(pose, rays, points, u, v, ftol=1e-4, **kw):
    """relocalization.py:186-187: scipy least_squares on reloc_residual with the reference's options (2-point Jacobian)."""
    from scipy.optimize import least_squares
    return least_squares(reloc_residual, np.asarray(pose, dtype=np.float64), x_scale='jac', ftol=ftol, method='trf',
                         args=(rays, points, u, v), **kw)
