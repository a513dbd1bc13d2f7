"""
KeyFrame record (reference: slam_system/key_frame.py:13-57).  Data holder only: the .mat export for random-forest training
(:75-107) belongs to the out-of-scope relocaliser.
"""
import numpy as np


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
