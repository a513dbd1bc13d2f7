"""
Index plumbing of the EKF (reference: slam_system/util.py:75-96).  get_overlap_index is host logic in the
reference too (a two-pointer merge over two short ascending index arrays); it is vectorised here, not ported.
"""
import numpy as np


def get_overlap_index(index1, index2):
    """Positions (in index1, in index2) of the values both ascending arrays share (util.py:75-96)."""
    index1 = np.asarray(index1)
    index2 = np.asarray(index2)
    _, i1, i2 = np.intersect1d(index1, index2, assume_unique=True, return_indices=True)
    return i1.astype(np.int64), i2.astype(np.int64)


# ---- synthetic noise model and small persistence helpers (host side of the reference's experiments) ---------------------
def add_gauss(points, var, max_width, max_height):
    """util.py:99-123: N(0, var) noise on every coordinate (Python `random`, x then y per point, so a seeded run reproduces
    the reference's stream), clamped to [0, max_width-1] x [0, max_height-1] on the upper side and 0 on the lower."""
    import random
    points = np.asarray(points)
    noise = np.array([[random.gauss(0, var), random.gauss(0, var)] for _ in range(len(points))]).reshape(-1, 2)
    out = np.zeros_like(points)
    out[:, 0:2] = points[:, 0:2] + noise
    out[:, 0] = np.where(out[:, 0] >= max_width, max_width - 1, out[:, 0])
    out[:, 1] = np.where(out[:, 1] >= max_height, max_height - 1, out[:, 1])
    out[:, 0:2] = np.where(out[:, 0:2] < 0, 0, out[:, 0:2])
    return out


def add_outliers(points, var, max_width, max_height, percentage):
    """util.py:126-139: `percentage` % of the points are replaced by uniform image positions, then add_gauss."""
    import random
    pts = np.array(points, copy=True)
    n = pts.shape[0]
    for i in random.sample(list(range(n)), int(percentage / 100 * n)):
        pts[i, 0] = random.uniform(0, max_width - 1)
        pts[i, 1] = random.uniform(0, max_height - 1)
    return add_gauss(pts, var, max_width, max_height)


def uniform_point_sample_on_field(x_max, y_max, x_num, y_num):
    """util.py:186-203: [x_num * y_num, 3] grid on the ground plane (z = 0), x-major."""
    xs, ys = np.meshgrid(np.linspace(0, x_max, x_num), np.linspace(0, y_max, y_num), indexing='ij')
    return np.stack([xs.ravel(), ys.ravel(), np.zeros(xs.size)], axis=1)


def save_camera_pose(pan, tilt, f, path):
    """util.py:263-277: pose sequence -> .mat with keys pan / tilt / f."""
    import scipy.io as sio
    sio.savemat(path, mdict={'pan': pan, 'tilt': tilt, 'f': f})


def load_camera_pose(path, separate=False):
    """util.py:280-298: (pan, tilt, f) arrays from either the three-key or the single [n,3] 'ptz' layout."""
    import scipy.io as sio
    d = sio.loadmat(path)
    if separate:
        return d['pan'].squeeze(), d['tilt'].squeeze(), d['f'].squeeze()
    ptz = d['ptz']
    return ptz[:, 0], ptz[:, 1], ptz[:, 2]


def compute_error_data(ptz, ground_truth):
    """util.py:301-319: mean and std of the absolute pan / tilt / f errors."""
    err = [np.fabs(np.asarray(a) - np.asarray(b)) for a, b in zip(ptz, ground_truth)]
    return [np.mean(e) for e in err], [np.std(e) for e in err]
