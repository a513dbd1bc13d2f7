"""
PtzSlam drop-in for the EKF hot path (reference: slam_system/ptz_slam.py).

    PtzSlam.compute_h_jacobian(pan, tilt, focal_length, rays)                    ptz_slam.py:73-138
    PtzSlam.ekf_update(observed_keypoints, observed_keypoint_index, height, width)   ptz_slam.py:210-289
    PtzSlam.predict()  - the predict lines of tracking()                          ptz_slam.py:418-426
    PtzSlam.remove_rays(index), PtzSlam.add_rays(img, bounding_box, detect)       ptz_slam.py:291-388 (state bookkeeping;
                                                                                  keypoint detection is an injected callable)
    PtzSlam.init_system / tracking / relocalize / add_keyframe                    ptz_slam.py:140-208, 390-537 (the per-frame
                                                                                  loop; OpenCV front end injected as `front_end`)

State attributes keep the reference's names and meaning (`rays`, `state_cov`, `cameras`, `current_camera`,
`velocity`, `observe_var`, `angle_var`, `f_var`).  `ekf_update` mutates them in place like the reference; the
computation (projection, in-image filter, innovation, central-difference Jacobian, S, Cholesky, gain, covariance
write-back) runs in csrc/ekf.cu through the C-ABI.  The image front-end of the reference class (SIFT detection,
optical-flow matching, SIFT matching) is out of scope and enters as `front_end` callables.

The filter state is RESIDENT on the GPU between frames (a one-sequence ptzba_ekf_batch): `ekf_update`, `predict`, `remove_rays`
and `add_rays` work on the device copy and per frame only the observations go in and the pose / velocity come back.  `rays` and
`state_cov` stay the reference's public attributes: reading them downloads the device copy when it is newer, ASSIGNING them
(`slam.rays = ...`) makes the host copy the truth again.  Code that mutates the returned arrays in place must call
`slam.invalidate_device()` afterwards (the reference's own methods never do that outside the ones mirrored here).

`BatchedEkfTracker` is the additive batched form for many independent sequences resident on the GPU (config 4).
"""
import copy
import ctypes

import numpy as np

from . import _lib


def keypoints_masking(kp, mask):
    """image_process.py:158-175 (array form): indices of the keypoints whose integer pixel has mask == 1."""
    kp = np.asarray(kp).reshape(-1, 2)
    if len(kp) == 0:
        return np.zeros(0, dtype=np.int32)
    x, y = kp[:, 0].astype(np.int64), kp[:, 1].astype(np.int64)       # int() truncation, as the reference
    return np.nonzero(np.asarray(mask)[y, x] == 1)[0].astype(np.int32)


def existing_keypoint_mask(keypoints, height, width, half=50):
    """ptz_slam.py:352-361: 1 everywhere except the 100 x 100 boxes around the keypoints the map already projects to."""
    mask = np.ones((height, width), np.uint8)
    for x, y in np.asarray(keypoints).reshape(-1, 2):
        mask[int(max(0, y - half)):int(min(height, y + half)), int(max(0, x - half)):int(min(width, x + half))] = 0
    return mask


def _params(u, v, disp, observe_var, angle_var, f_var, height, width, jac_mode):
    p = _lib.EkfParams()
    p.u, p.v = float(u), float(v)
    d = np.zeros(6) if disp is None else np.asarray(disp, dtype=np.float64)
    for i in range(6):
        p.disp[i] = float(d[i])
    p.observe_var, p.angle_var, p.f_var = float(observe_var), float(angle_var), float(f_var)
    p.height, p.width = float(height), float(width)
    p.jac_mode = int(jac_mode)
    return p


class PtzSlam:
    def __init__(self, front_end=None):
        """ptz_slam.py:24-71.  `front_end` supplies the OpenCV side that is outside this library, as callables:
            front_end.detect_keypoints(img, n)                       -> (points[n,2], descriptors[n,d])   (:155, :338)
            front_end.matching_and_ransac(img1, img2, kp1, kp1_idx)  -> (inlier_kp, inlier_idx, outlier_idx)   (:397)
            front_end.detect / front_end.match                       -> relocalization.relocalization_camera's pair
            front_end.build_matching_graph                           -> keyframe bundle adjustment (scene_map.Map)
        Without it only the image-free members work (init_rays, predict, ekf_update, remove_rays, add_rays(detector))."""
        self.front_end = front_end
        # device-resident filter state (see the module docstring): handle, what is current where
        self._dev = None                # ctypes handle of the one-sequence ptzba_ekf_batch
        self._dev_valid = False         # the device copy is the current state
        self._rays_stale = False        # the host copy of rays / state_cov is older than the device copy
        self._cov_stale = False
        self._dev_prm = None
        # global rays and covariance matrix (ptz_slam.py:29-31)
        self.rays = np.ndarray([0, 2])
        self.des = np.ndarray([0, 128])
        self.state_cov = np.zeros([3, 3])
        # previous frame (ptz_slam.py:34-38)
        self.previous_img = None
        self.previous_keypoints = None
        self.previous_keypoints_index = None
        self.current_camera = None
        from .scene_map import Map                      # here: scene_map imports bundle_adjustment, which imports nothing of ours
        self.keyframe_map = Map('sift', build_matching_graph=getattr(front_end, "build_matching_graph", None))
        self.rf_map = None                              # RandomForestMap / NNBasedMap, set by the caller when used
        self.cameras = []
        self.velocity = np.zeros(3)
        self.new_keyframe = False                       # state flags (ptz_slam.py:56-63)
        self.tracking_lost = False
        self.bad_tracking_cnt = 0
        # hyper parameters (ptz_slam.py:65-71)
        self.keypoint_num = 500
        self.observe_var = 0.1
        self.angle_var = 0.001
        self.f_var = 1
        # PTZBA_JAC_CENTRAL_FD reproduces the reference's central differences; JAC_ANALYTIC is the closed form
        self.jacobian_mode = _lib.JAC_CENTRAL_FD

    # -- rays / state_cov: host attributes backed by the device-resident copy ---------------------------------------
    @property
    def rays(self):
        if self._rays_stale:
            n = self._dev_n_rays()
            self._rays = np.empty((n, 2), np.float64)
            ctx = _lib.get_context()
            ctx.check(ctx.lib.ptzba_ekf_batch_get_rays(self._dev, 0, _lib.ptr(self._rays)))
            self._rays_stale = False
        return self._rays

    @rays.setter
    def rays(self, value):
        self._sync_other("rays")
        self._rays = value
        self._rays_stale = False
        self._dev_valid = False

    @property
    def state_cov(self):
        if self._cov_stale:
            s_ = 3 + 2 * self._dev_n_rays()
            self._cov = np.empty((s_, s_), np.float64)
            ctx = _lib.get_context()
            ctx.check(ctx.lib.ptzba_ekf_batch_get_cov(self._dev, 0, _lib.ptr(self._cov)))
            self._cov_stale = False
        return self._cov

    @state_cov.setter
    def state_cov(self, value):
        self._sync_other("cov")
        self._cov = value
        self._cov_stale = False
        self._dev_valid = False

    def _sync_other(self, which):
        """Assigning one of the two attributes invalidates the device copy: fetch the other one first if only the device has it."""
        if which == "rays" and getattr(self, "_cov_stale", False):
            _ = self.state_cov
        if which == "cov" and getattr(self, "_rays_stale", False):
            _ = self.rays

    def invalidate_device(self):
        """Call after mutating `rays` / `state_cov` IN PLACE: the host copies become the truth again."""
        _ = self.rays, self.state_cov
        self._dev_valid = False

    def _dev_n_rays(self):
        ctx = _lib.get_context()
        n = ctypes.c_int32(0)
        ctx.check(ctx.lib.ptzba_ekf_batch_n_rays(self._dev, 0, ctypes.byref(n), None))
        return int(n.value)

    def _dev_release(self):
        if self._dev is not None:
            _ = self.rays, self.state_cov            # keep whatever only the device has
            _lib.get_context().lib.ptzba_ekf_batch_destroy(self._dev)
            self._dev = None
        self._dev_valid = False

    def __del__(self):
        try:
            if self._dev is not None:
                _lib.get_context().lib.ptzba_ekf_batch_destroy(self._dev)
                self._dev = None
        except Exception:
            pass

    def _dev_ensure(self, prm_key, prm, max_obs):
        """Make the device copy current (upload only when the host copy is the truth) with room for max_obs observations."""
        ctx = _lib.get_context()
        if self._dev is not None and (self._dev_prm != prm_key):
            self._dev_release()
        if not self._dev_valid:
            rays = _lib.f64(self._rays).reshape(-1, 2)
            cov = _lib.f64(self._cov)
            n = rays.shape[0]
            assert cov.shape == (3 + 2 * n, 3 + 2 * n)
            if self._dev is not None and self._dev_n_rays() != n:
                lib = ctx.lib
                lib.ptzba_ekf_batch_destroy(self._dev)
                self._dev = None
            if self._dev is None:
                h = ctypes.c_void_p()
                ptz0 = np.zeros(3)
                ctx.check(ctx.lib.ptzba_ekf_batch_create(ctx.handle, ctypes.byref(prm), 1, n, int(min(max(max_obs, 1), max(n, 1))),
                                                         _lib.ptr(rays), _lib.ptr(ptz0), ctypes.byref(h)))
                self._dev = h
                self._dev_prm = prm_key
            ctx.check(ctx.lib.ptzba_ekf_batch_set(self._dev, 0, None, None, _lib.ptr(rays), _lib.ptr(cov)))
            self._dev_valid = True
            self._rays_stale = self._cov_stale = False
        ctx.check(ctx.lib.ptzba_ekf_batch_reserve(self._dev, 0, int(max_obs)))

    # -- state initialisation without images (ptz_slam.py:186-208 minus keypoint detection) ----------------------
    def init_rays(self, rays, camera):
        """Set the ray landmarks and the initial covariance exactly as init_system does (:190-200, :208)."""
        self.rays = np.array(rays, dtype=np.float64).reshape(-1, 2)
        self.state_cov = self.angle_var * np.eye(3 + 2 * len(self.rays))
        self.state_cov[2][2] = self.f_var
        self.cameras = [camera]
        self.current_camera = camera

    def _front(self, name):
        fn = getattr(self.front_end, name, None)
        if fn is None:
            raise NotImplementedError("PtzSlam needs front_end.%s (OpenCV feature detection / matching is outside this "
                                      "library)" % name)
        return fn

    def init_system(self, img, camera, bounding_box=None):
        """ptz_slam.py:140-208: first frame, or the frame after a relocalisation.  Keypoints off the players become the ray
        landmarks (back-projection through `camera` on the GPU); covariance angle_var * I with f_var for the focal length."""
        first_img_kp, first_des = self._front("detect_keypoints")(img, self.keypoint_num)
        first_img_kp = np.asarray(first_img_kp)
        if bounding_box is not None:
            masked_index = keypoints_masking(first_img_kp, bounding_box)
            first_img_kp = first_img_kp[masked_index]
            first_des = None if first_des is None else np.asarray(first_des)[masked_index]
        init_rays = camera.back_project_to_rays(first_img_kp)
        self.rays = np.asarray(init_rays, dtype=np.float64).reshape(-1, 2).copy()
        self.des = first_des
        self.state_cov = self.angle_var * np.eye(3 + 2 * len(self.rays))
        self.state_cov[2][2] = self.f_var
        self.previous_img = img
        self.previous_keypoints = first_img_kp
        self.previous_keypoints_index = np.array([i for i in range(len(self.rays))])
        self.cameras.append(camera)

    def tracking(self, next_img, bad_tracking_percentage, bounding_box=None):
        """ptz_slam.py:390-456: one frame.  Optical-flow matches (front end) -> tracking-quality bookkeeping -> predict ->
        EKF update on the GPU -> drop RANSAC outliers -> detect / add new rays -> keyframe test."""
        inlier_keypoints, inlier_index, outlier_index = self._front("matching_and_ransac")(
            self.previous_img, next_img, self.previous_keypoints, self.previous_keypoints_index)
        tracking_percentage = len(inlier_index) / len(self.previous_keypoints) * 100
        if tracking_percentage < bad_tracking_percentage:
            self.bad_tracking_cnt += 1
        if self.bad_tracking_cnt > 3:
            self.tracking_lost = True
            self.bad_tracking_cnt = 0
        # 1. predict (a lost frame's camera is not appended, :421-422)
        self.current_camera = copy.deepcopy(self.cameras[-1])
        self.current_camera.set_ptz(self.current_camera.get_ptz() + self.velocity)
        if not self.tracking_lost:
            self.cameras.append(self.current_camera)
        self._predict_cov()
        # 2. update
        height, width = next_img.shape[0:2]
        self.ekf_update(inlier_keypoints, inlier_index, height, width)
        # 3. delete outliers, 4. add new features and roll the previous frame
        self.remove_rays(outlier_index)
        self.previous_img = next_img
        self.previous_keypoints, self.previous_keypoints_index = self.add_rays(next_img, bounding_box,
                                                                               self._front("detect_keypoints"))
        if self.keyframe_map.good_new_keyframe(self.current_camera.get_ptz(), 10, 15):
            self.new_keyframe = True
        return tracking_percentage

    def relocalize(self, img, camera, enable_rf=False, bounding_box=None):
        """ptz_slam.py:458-497: pose of a lost frame from the keyframe map (or from rf_map when enable_rf)."""
        from .key_frame import KeyFrame
        from .relocalization import relocalization_camera
        if enable_rf:
            pan, tilt, focal_length = camera.pan, camera.tilt, camera.focal_length
            frame = KeyFrame(img, -1, camera.camera_center, camera.base_rotation, camera.principal_point[0],
                             camera.principal_point[1], pan, tilt, focal_length)
            kp, des = self._front("detect_keypoints")(img, 500)
            if bounding_box is not None:
                masked_index = keypoints_masking(kp, bounding_box)
                kp, des = np.asarray(kp)[masked_index], np.asarray(des)[masked_index]
            frame.feature_pts, frame.feature_des = kp, des
            camera.set_ptz(self.rf_map.relocalize(frame, [pan, tilt, focal_length]))      # ptz_slam.py:491
        elif len(self.keyframe_map.keyframe_list) > 1:
            lost_pose = camera.pan, camera.tilt, camera.focal_length
            camera.set_ptz(relocalization_camera(self.keyframe_map, img, lost_pose, detect=self._front("detect"),
                                                 match=self._front("match")))
        else:
            print("Warning: Not enough keyframes for relocalization.")
        self.tracking_lost = False
        return camera

    def add_keyframe(self, img, camera, frame_index, enable_rf=False):
        """ptz_slam.py:499-537: the current frame becomes a keyframe; all keyframes are bundle-adjusted on the GPU."""
        from .key_frame import KeyFrame
        new_keyframe = KeyFrame(img, frame_index, camera.camera_center, camera.base_rotation, camera.principal_point[0],
                                camera.principal_point[1], camera.pan, camera.tilt, camera.focal_length)
        if enable_rf:
            new_keyframe.feature_pts = self.previous_keypoints
            new_keyframe.feature_des = np.asarray(self.des)[np.asarray(self.previous_keypoints_index).astype(np.int64)]
            self.rf_map.add_keyframe(new_keyframe)
            self.new_keyframe = False
        elif frame_index == 0:
            self.keyframe_map.add_first_keyframe(new_keyframe, verbose=False)
        else:
            self.keyframe_map.add_keyframe_with_ba(new_keyframe, "./bundle_result/", verbose=False)
            self.new_keyframe = False

    # -- ray bookkeeping between frames (ptz_slam.py:291-388; SURVEY.md 8f row N4, host side as in the reference) ----------
    def remove_rays(self, index):
        """ptz_slam.py:291-315: drop the rays `index` (RANSAC outliers) with their descriptors and covariance rows / columns."""
        delete_index = np.asarray(index, dtype=np.int64).reshape(-1)
        if self.des is not None and len(self.des) > 0:
            self.des = np.delete(self.des, delete_index, axis=0)
        if self._dev_valid:
            # resident state: rows / columns are compacted on the GPU (csrc/ekf.cu:ptzba_ekf_batch_remove_rays)
            if len(delete_index):
                ctx = _lib.get_context()
                di = _lib.i32(np.unique(delete_index))
                ctx.check(ctx.lib.ptzba_ekf_batch_remove_rays(self._dev, 0, int(di.shape[0]), _lib.ptr(di)))
                self._rays_stale = self._cov_stale = True
            return
        self.rays = np.delete(self.rays, delete_index, axis=0)
        p_delete = np.stack([2 * delete_index + 3, 2 * delete_index + 4], axis=1).reshape(-1)
        self.state_cov = np.delete(np.delete(self.state_cov, p_delete, axis=0), p_delete, axis=1)

    def add_rays(self, img, bounding_box, detect_keypoints):
        """ptz_slam.py:317-388 with the detector injected: `detect_keypoints(img, keypoint_num) -> (points[n,2], des[n,d] or None)`
        stands for detect_compute_sift_array (:338).  New keypoints on players (bounding_box == 0) or within 50 px of a keypoint
        the map already projects to are dropped; the others become rays (back-projection through the current camera on the
        GPU) with variance angle_var.  Returns (keypoints, keypoints_index) like the reference."""
        height, width = img.shape[0:2]
        keypoints, keypoints_index = self.current_camera.project_rays(self.rays, height, width)
        new_keypoints, new_des = detect_keypoints(img, self.keypoint_num)
        new_keypoints = np.asarray(new_keypoints, dtype=np.float64).reshape(-1, 2)
        if bounding_box is not None:
            keep = keypoints_masking(new_keypoints, bounding_box)
            new_keypoints = new_keypoints[keep]
            new_des = None if new_des is None else np.asarray(new_des)[keep]
        keep = keypoints_masking(new_keypoints, existing_keypoint_mask(keypoints, height, width))
        new_keypoints = new_keypoints[keep]
        new_des = None if new_des is None else np.asarray(new_des)[keep]
        k = len(new_keypoints)
        if k > 0:
            new_rays = np.asarray(self.current_camera.back_project_to_rays(new_keypoints), dtype=np.float64).reshape(-1, 2)
            n_old = len(self.rays)
            if new_des is not None:
                self.des = np.vstack([self.des.reshape(-1, new_des.shape[1]) if len(self.des) else np.zeros((0, new_des.shape[1])), new_des])
            if self._dev_valid:
                # resident state: the new rays and their covariance rows / columns are appended on the GPU
                ctx = _lib.get_context()
                nr = _lib.f64(new_rays)
                ctx.check(ctx.lib.ptzba_ekf_batch_add_rays(self._dev, 0, int(k), _lib.ptr(nr)))
                self._rays_stale = self._cov_stale = True
            else:
                self.rays = np.vstack([np.asarray(self.rays, dtype=np.float64).reshape(-1, 2), new_rays])
                s_old = self.state_cov.shape[0]
                cov = np.zeros((s_old + 2 * k, s_old + 2 * k))
                cov[:s_old, :s_old] = self.state_cov
                d = np.arange(s_old, s_old + 2 * k)
                cov[d, d] = self.angle_var
                self.state_cov = cov
            keypoints_index = np.append(keypoints_index, np.arange(n_old, n_old + k))
        keypoints = np.concatenate([np.asarray(keypoints).reshape(-1, 2), new_keypoints], axis=0)
        return keypoints, keypoints_index

    def compute_h_jacobian(self, pan, tilt, focal_length, rays):
        """ptz_slam.py:73-138: dense H [2n, 3+2n]; principal point / displacement come from self.cameras[0] (:92)."""
        ctx = _lib.get_context()
        cam = self.cameras[0]
        rays = _lib.f64(rays).reshape(-1, 2)
        n = rays.shape[0]
        ptz = _lib.f64([pan, tilt, focal_length])
        disp = _lib.f64(cam.displacement)
        disp = disp if np.any(disp != 0.0) else None
        H = np.empty((2 * n, 3 + 2 * n), np.float64)
        ctx.check(ctx.lib.ptzba_h_jacobian_dense(ctx.handle, _lib.HOST, _lib.ptr(ptz), float(cam.principal_point[0]),
                                                 float(cam.principal_point[1]), _lib.ptr(disp), n, _lib.ptr(rays),
                                                 self.jacobian_mode, _lib.ptr(H)))
        return H

    def predict(self):
        """ptz_slam.py:418-426: constant-velocity pose prediction, pose-block process noise."""
        self.current_camera = copy.deepcopy(self.cameras[-1])
        self.current_camera.set_ptz(self.current_camera.get_ptz() + self.velocity)
        self.cameras.append(self.current_camera)
        self._predict_cov()

    def _predict_cov(self):
        """P[0:3, 0:3] += 5 diag(angle_var, angle_var, f_var) (ptz_slam.py:425-426), on whichever copy is current."""
        if self._dev_valid:
            ctx = _lib.get_context()
            ctx.check(ctx.lib.ptzba_ekf_batch_predict_cov(self._dev))
            self._cov_stale = True
        else:
            q_k = 5 * np.diag([self.angle_var, self.angle_var, self.f_var])
            cov = self.state_cov
            cov[0:3, 0:3] = cov[0:3, 0:3] + q_k
            self.state_cov = cov

    def ekf_update(self, observed_keypoints, observed_keypoint_index, height, width):
        """ptz_slam.py:210-289.  Mutates rays, state_cov, current_camera (pan/tilt/focal_length) and velocity."""
        ctx = _lib.get_context()
        cam = self.current_camera
        ref = self.cameras[0]
        obs = _lib.f64(observed_keypoints).reshape(-1, 2)
        idx = _lib.i32(np.asarray(observed_keypoint_index).astype(np.int64))
        m = obs.shape[0]
        assert idx.shape[0] == m
        disp = None if ref.displacement is None else tuple(np.asarray(ref.displacement, dtype=np.float64).ravel())
        prm_key = (float(ref.principal_point[0]), float(ref.principal_point[1]), disp, float(self.observe_var), float(self.angle_var),
                   float(self.f_var), float(height), float(width), int(self.jacobian_mode))
        prm = _params(ref.principal_point[0], ref.principal_point[1], ref.displacement, self.observe_var, self.angle_var,
                      self.f_var, height, width, self.jacobian_mode)
        self._dev_ensure(prm_key, prm, m)
        ptz = _lib.f64([cam.pan, cam.tilt, cam.focal_length])
        ctx.check(ctx.lib.ptzba_ekf_batch_set(self._dev, 0, _lib.ptr(ptz), None, None, None))
        cnt = np.array([m], np.int32)
        matched = np.zeros(1, np.int32)
        cap = ctypes.c_int32(0)
        ctx.check(ctx.lib.ptzba_ekf_batch_max_obs(self._dev, ctypes.byref(cap)))
        # the batch call takes observation arrays padded to its max_obs
        if m < cap.value:
            obs_p = np.zeros((cap.value, 2)); obs_p[:m] = obs
            idx_p = np.zeros(cap.value, np.int32); idx_p[:m] = idx
        else:
            obs_p, idx_p = obs, idx
        ctx.check(ctx.lib.ptzba_ekf_batch_update_only(self._dev, _lib.HOST, _lib.ptr(obs_p), _lib.ptr(idx_p), _lib.ptr(cnt),
                                                      _lib.ptr(matched)))
        vel = np.zeros(3)
        ctx.check(ctx.lib.ptzba_ekf_batch_get(self._dev, _lib.ptr(ptz), _lib.ptr(vel), None))
        self._rays_stale = self._cov_stale = True
        cam.pan, cam.tilt, cam.focal_length = float(ptz[0]), float(ptz[1]), float(ptz[2])
        self.current_camera = cam
        self.velocity = vel
        return int(matched[0])


class BatchedEkfTracker:
    """Many independent EKF sequences resident on the GPU (one ptzba_ekf_batch); every sequence has n_ray rays."""

    def __init__(self, rays0, ptz0, u, v, max_obs, height, width, displacement=None, observe_var=0.1, angle_var=0.001,
                 f_var=1.0, jacobian_mode=_lib.JAC_ANALYTIC, ctx=None):
        self.ctx = ctx or _lib.get_context()
        rays0 = _lib.f64(rays0)
        ptz0 = _lib.f64(ptz0).reshape(-1, 3)
        self.n_seq = ptz0.shape[0]
        rays0 = rays0.reshape(self.n_seq, -1, 2)
        self.n_ray = rays0.shape[1]
        self.max_obs = int(min(max_obs, self.n_ray))
        prm = _params(u, v, displacement, observe_var, angle_var, f_var, height, width, jacobian_mode)
        h = ctypes.c_void_p()
        self.ctx.check(self.ctx.lib.ptzba_ekf_batch_create(self.ctx.handle, ctypes.byref(prm), self.n_seq, self.n_ray,
                                                           self.max_obs, _lib.ptr(rays0), _lib.ptr(ptz0), ctypes.byref(h)))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.ptzba_ekf_batch_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pack_observations(self, obs_xy_list, obs_idx_list):
        """Ragged per-sequence observations -> padded [n_seq, max_obs, 2] / [n_seq, max_obs] / [n_seq] arrays."""
        xy = np.zeros((self.n_seq, self.max_obs, 2))
        ix = np.zeros((self.n_seq, self.max_obs), np.int32)
        cnt = np.zeros(self.n_seq, np.int32)
        for b, (o, i) in enumerate(zip(obs_xy_list, obs_idx_list)):
            m = len(i)
            assert m <= self.max_obs
            xy[b, :m] = o
            ix[b, :m] = i
            cnt[b] = m
        return xy, ix, cnt

    def step(self, obs_xy, obs_idx, obs_cnt, predict=True, mem=_lib.HOST):
        """One predict+update (or update only) for every sequence; returns matched counts [n_seq]."""
        matched = np.zeros(self.n_seq, np.int32)
        fn = self.ctx.lib.ptzba_ekf_batch_step if predict else self.ctx.lib.ptzba_ekf_batch_update_only
        if mem == _lib.HOST:
            obs_xy, obs_idx, obs_cnt = _lib.f64(obs_xy), _lib.i32(obs_idx), _lib.i32(obs_cnt)
        self.ctx.check(fn(self.handle, mem, _lib.ptr(obs_xy), _lib.ptr(obs_idx), _lib.ptr(obs_cnt), _lib.ptr(matched)))
        return matched

    def route(self):
        """0 = Cholesky route, 1 = pivoted-LU route (indefinite innovation covariance), per sequence."""
        r = np.zeros(self.n_seq, np.int32)
        self.ctx.check(self.ctx.lib.ptzba_ekf_batch_route(self.handle, _lib.ptr(r)))
        return r

    def get_state(self, want_rays=True):
        ptz = np.empty((self.n_seq, 3)); vel = np.empty((self.n_seq, 3))
        rays = np.empty((self.n_seq, self.n_ray, 2)) if want_rays else None
        self.ctx.check(self.ctx.lib.ptzba_ekf_batch_get(self.handle, _lib.ptr(ptz), _lib.ptr(vel), _lib.ptr(rays)))
        return ptz, vel, rays

    def get_cov(self, seq):
        s = 3 + 2 * self.n_ray
        P = np.empty((s, s))
        self.ctx.check(self.ctx.lib.ptzba_ekf_batch_get_cov(self.handle, int(seq), _lib.ptr(P)))
        return P

    def set_state(self, seq, ptz=None, velocity=None, rays=None, state_cov=None):
        a = [None if x is None else _lib.f64(x) for x in (ptz, velocity, rays, state_cov)]
        self.ctx.check(self.ctx.lib.ptzba_ekf_batch_set(self.handle, int(seq), *[_lib.ptr(x) for x in a]))
