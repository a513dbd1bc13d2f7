"""
Keyframe map orchestration, image-free (reference: slam_system/scene_map.py:18-149).

`Map` keeps the keyframe list and the global ray landmarks; `add_keyframe_with_ba` re-runs bundle adjustment over all
keyframes on the GPU (bundle_adjustment.bundle_adjustment -> libptzba) and `good_new_keyframe` is the pan-overlap rule.
The vision front-end that matches keyframe images is injected (`build_matching_graph`), as in bundle_adjustment().

`RandomForestMap` (scene_map.py:170-244) is the map variant the relocaliser trains from: it bundle-adjusts a sliding window
of the last 10 keyframes, exports every keyframe as .mat and hands the list to the random-forest builder (C++ `rf_map/`,
out of scope here: injected as `create_map`).
"""
import os
import time

import numpy as np

from .bundle_adjustment import bundle_adjustment, overlap_pan_angle
from .key_frame import KeyFrame


class Map:
    def __init__(self, feature_method, build_matching_graph=None):
        assert feature_method in ('sift', 'orb', 'latch')
        self.global_ray = np.ndarray([0, 2])        # [N, 2] float64 ray landmarks
        self.keyframe_list = []
        self.feature_method = feature_method
        self.build_matching_graph = build_matching_graph
        self.last_ba_seconds = None

    def add_first_keyframe(self, keyframe, verbose=False):
        """scene_map.py:30-40: first keyframe, no bundle adjustment."""
        assert isinstance(keyframe, KeyFrame)
        self.keyframe_list = [keyframe]
        if verbose:
            print('first key frame is added, no bundle adjustment and landmark')

    def add_keyframe_without_ba(self, keyframe, verbose=False):
        """scene_map.py:42-51."""
        assert isinstance(keyframe, KeyFrame)
        self.keyframe_list.append(keyframe)

    def add_keyframe_with_ba(self, keyframe, save_path, verbose=False):
        """scene_map.py:53-117: add one keyframe and bundle-adjust all keyframes; the map is replaced by the result."""
        assert isinstance(keyframe, KeyFrame)
        assert len(self.keyframe_list) >= 1
        ref_frame = self.keyframe_list[0]
        camera_center, base_rotation = ref_frame.center, ref_frame.base_rotation
        u, v = ref_frame.u, ref_frame.v
        self.add_keyframe_without_ba(keyframe, False)
        N = len(self.keyframe_list)
        images, image_indices = [], []
        initial_ptzs = np.zeros((N, 3))
        for i, kf in enumerate(self.keyframe_list):
            images.append(kf.img)
            image_indices.append(kf.img_index)
            initial_ptzs[i] = kf.pan, kf.tilt, kf.f
        start = time.time()
        landmarks, keyframes = bundle_adjustment(images, image_indices, self.feature_method, initial_ptzs, camera_center,
                                                 base_rotation, u, v, save_path, verbose,
                                                 build_matching_graph=self.build_matching_graph)
        self.last_ba_seconds = time.time() - start
        self.keyframe_list.pop()
        self.global_ray = landmarks
        self.keyframe_list = []
        for i, kf in enumerate(keyframes):
            if kf.get_feature_num() > 0:
                self.keyframe_list.append(kf)
            elif verbose:
                print('warning: key frame, %d, image index %d is not included in the map' % (i, image_indices[i]))
        if verbose:
            print('updated map, number of key frame: %d, number of landmark %d' % (len(self.keyframe_list), len(landmarks)))
        return landmarks, self.keyframe_list

    def good_new_keyframe(self, ptz, threshold1=5, threshold2=20, im_width=1280, verbose=False):
        """scene_map.py:119-149: True when the largest pan overlap with the existing keyframes lies in (threshold1, threshold2)."""
        ptz = np.asarray(ptz)
        assert ptz.shape[0] == 3
        if len(self.keyframe_list) == 0:
            return False
        overlaps = [overlap_pan_angle(ptz[2], ptz[0], kf.f, kf.pan, im_width) for kf in self.keyframe_list]
        if verbose:
            print('candidate key frame overlap: ', overlaps)
        max_overlap = max(overlaps)
        return max_overlap > threshold1 and max_overlap < threshold2

    def save_keyframes_to_mat(self, path):
        """scene_map.py:150-167: index / ptz / center / base_rotation / principal_point of every keyframe."""
        import scipy.io as sio
        keyframes = [{'index': kf.img_index, 'ptz': np.array([kf.pan, kf.tilt, kf.f]), 'center': kf.center,
                      'base_rotation': kf.base_rotation, 'principal_point': np.array([kf.u, kf.v])}
                     for kf in self.keyframe_list]
        sio.savemat(path, mdict={'keyframes': keyframes})


class RandomForestMap:
    MAX_BA_FRAME = 10           # scene_map.py:210

    def __init__(self, keyframe_location="./keyframes/", mat_path_file="./train_feature_file.txt", create_map=None,
                 build_matching_graph=None, bundle_adjustment_fn=None, relocalizer=None, relocalize_file="./relocalize.mat"):
        """scene_map.py:170-180 with the hard-coded Windows paths turned into arguments.  `create_map(mat_path_file)` stands
        for RFMap.createMap (:197-198) and `relocalizer(relocalize_file, init_ptz)` for RFMap.relocalization (:259-268): the
        random-forest C++ library is not part of this package.  `bundle_adjustment_fn` defaults to the GPU bundle adjustment
        of this package."""
        self.keyframe_location = keyframe_location
        self.mat_path_file = mat_path_file
        self.create_map = create_map
        self.relocalizer = relocalizer
        self.relocalize_file = relocalize_file
        self.build_matching_graph = build_matching_graph
        self.bundle_adjustment_fn = bundle_adjustment_fn
        self.keyframe_list = []
        self.feature_method = 'sift'

    def add_keyframe(self, keyframe):
        """scene_map.py:182-198: append, re-adjust the window, export every keyframe and rebuild the forest."""
        self.keyframe_list.append(keyframe)
        if len(self.keyframe_list) > 1:
            self.bundle_adjustment_processing()
        with open(self.mat_path_file, 'w') as f:
            for frame in self.keyframe_list:
                mat_path = os.path.join(self.keyframe_location, str(frame.img_index) + ".mat")
                f.write(mat_path + "\n")
                frame.save_to_mat(mat_path)
        if self.create_map is not None:
            self.create_map(self.mat_path_file)

    def bundle_adjustment_processing(self):
        """scene_map.py:202-244: bundle-adjust the last MAX_BA_FRAME keyframes (camera constants from the first keyframe of
        the map); older keyframes are kept as they are, adjusted ones without features are dropped."""
        ref_frame = self.keyframe_list[0]
        n = len(self.keyframe_list)
        first = max(0, n - self.MAX_BA_FRAME)
        kept, window = self.keyframe_list[:first], self.keyframe_list[first:]
        initial_ptzs = np.array([[kf.pan, kf.tilt, kf.f] for kf in window])
        image_indices = [kf.img_index for kf in window]
        if self.bundle_adjustment_fn is not None:
            landmarks, keyframes = self.bundle_adjustment_fn([kf.img for kf in window], image_indices, self.feature_method,
                                                             initial_ptzs, ref_frame.center, ref_frame.base_rotation,
                                                             ref_frame.u, ref_frame.v, "./bundle_result")
        else:
            landmarks, keyframes = bundle_adjustment([kf.img for kf in window], image_indices, self.feature_method,
                                                     initial_ptzs, ref_frame.center, ref_frame.base_rotation, ref_frame.u,
                                                     ref_frame.v, "./bundle_result",
                                                     build_matching_graph=self.build_matching_graph)
        self.keyframe_list = kept
        for i, kf in enumerate(keyframes):
            if kf.get_feature_num() > 0:
                kf.convert_keypoint_to_array()
                self.keyframe_list.append(kf)
            else:
                print('warning: key frame, %d, image index %d is not included in the map' % (i, image_indices[i]))
        return landmarks

    def add_keyframes(self, keyframe_list):
        """scene_map.py:246-257: a no-op in the reference (its body is commented out)."""
        pass

    def relocalize(self, relocalize_frame, init_ptz):
        """scene_map.py:259-268: the lost frame's features are written to `relocalize_file` and the forest's pose estimate
        (started from init_ptz) comes back as a flat [pan, tilt, f]."""
        if self.relocalizer is None:
            raise RuntimeError("RandomForestMap.relocalize needs the relocalizer callable (the random-forest library is external)")
        relocalize_frame.save_to_mat(self.relocalize_file)
        return np.asarray(self.relocalizer(self.relocalize_file, init_ptz), dtype=np.float64).ravel()

    def good_keyframe(self, ptz, threshold1=5, threshold2=20, im_width=1280, verbose=False):
        """scene_map.py:270-298: the overlap rule of Map.good_new_keyframe against the poses stored in the exported keyframe
        files (the list add_keyframe wrote to mat_path_file)."""
        import scipy.io as sio
        ptz = np.asarray(ptz)
        assert ptz.shape[0] == 3
        with open(self.mat_path_file, 'r') as f:
            files = f.read().splitlines()
        if len(files) == 0:
            print("Warning: Not existing keyframes")
        overlaps = []
        for path in files:
            map_ptz = np.asarray(sio.loadmat(path)['ptz'], dtype=np.float64).ravel()
            overlaps.append(overlap_pan_angle(ptz[2], ptz[0], map_ptz[2], map_ptz[0], im_width))
        if verbose:
            print('candidate key frame overlap: ', overlaps)
        max_overlap = max(overlaps)
        return max_overlap > threshold1 and max_overlap < threshold2
