"""
KeyFrame record (reference: slam_system/key_frame.py:13-107): the data holder the map keeps per keyframe, the
keypoint-list -> array conversion (:58-73) and the .mat export the random-forest relocaliser trains from (:75-107).
"""
import numpy as np


def rotation_matrix_to_vector(R):
    """Rotation matrix -> axis-angle vector (what cv.Rodrigues returns for a 3x3 input, key_frame.py:94-97)."""
    R = np.asarray(R, dtype=np.float64).reshape(3, 3)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) * 0.5      # sin(theta) * axis
    s = np.linalg.norm(w)
    c = np.clip((np.trace(R) - 1.0) * 0.5, -1.0, 1.0)
    theta = np.arctan2(s, c)
    if s > 1e-10:
        return w * (theta / s)
    if c > 0:                                   # theta ~ 0
        return w
    # theta ~ pi: axis from the symmetric part R + I = 2 a a^T
    d = np.sqrt(np.maximum((np.diag(R) + 1.0) * 0.5, 0.0))
    k = int(np.argmax(d))
    a = (R[:, k] + np.eye(3)[k]) * 0.5 / d[k]
    return a / np.linalg.norm(a) * theta


class KeyFrame:
    def __init__(self, img, img_index, center, rotation, u, v, pan, tilt, f):
        self.img = img
        self.img_index = img_index
        self.feature_pts = np.ndarray(0)       # keypoints of this keyframe that are landmarks of the map
        self.feature_des = np.ndarray(0)       # their descriptors ([N,128] for SIFT)
        self.landmark_index = []               # [N] index of every keypoint in Map.global_ray
        self.pan, self.tilt, self.f = pan, tilt, f
        self.center = center
        self.base_rotation = rotation
        self.u = u
        self.v = v

    def get_feature_num(self):
        return len(self.feature_pts)

    def convert_keypoint_to_array(self, norm=True):
        """key_frame.py:58-73: KeyPoint list -> [N,2] float64; descriptors to float64, L2-normalised per row when norm."""
        pts = np.zeros((len(self.feature_pts), 2), dtype=np.float64)
        for k, p in enumerate(self.feature_pts):
            pts[k] = p.pt
        des = np.asarray(self.feature_des)
        if norm:
            des = np.divide(des, np.linalg.norm(des, axis=1).reshape(-1, 1)).astype(np.float64)
        else:
            des = des.astype(np.float64)
        self.feature_pts = pts
        self.feature_des = des

    def mat_record(self):
        """key_frame.py:80-105: the dictionary save_to_mat writes (im_name, keypoint, descriptor, camera[9,1], ptz[3,1])."""
        if type(self.feature_pts) == list:
            self.convert_keypoint_to_array()
        br = np.asarray(self.base_rotation, dtype=np.float64)
        br = rotation_matrix_to_vector(br) if br.shape == (3, 3) else br.ravel()
        return {
            'im_name': str(self.img_index) + ".jpg",
            'keypoint': self.feature_pts,
            'descriptor': self.feature_des,
            'camera': np.array([self.u, self.v, self.f, br[0], br[1], br[2],
                                self.center[0], self.center[1], self.center[2]]).reshape(-1, 1),
            'ptz': np.array([self.pan, self.tilt, self.f]).reshape(-1, 1),
        }

    def save_to_mat(self, path):
        """key_frame.py:75-107: keyframe in the format the random-forest map builder reads."""
        import scipy.io as sio
        sio.savemat(path, mdict=self.mat_record())

    @classmethod
    def load_mat(cls, path, img=None):
        """Inverse of save_to_mat (the reference reads these files in C++, rf_map/; this is the Python-side reader)."""
        import scipy.io as sio
        from .ptz_camera import _rodrigues
        d = sio.loadmat(path)
        cam = d['camera'].ravel()
        ptz = d['ptz'].ravel()
        name = str(np.asarray(d['im_name']).ravel()[0])
        kf = cls(img, int(name.split('.')[0]), cam[6:9].copy(), _rodrigues(cam[3:6]), cam[0], cam[1], ptz[0], ptz[1], ptz[2])
        kf.feature_pts = np.asarray(d['keypoint'], dtype=np.float64).reshape(-1, 2)
        kf.feature_des = np.asarray(d['descriptor'], dtype=np.float64)
        return kf
