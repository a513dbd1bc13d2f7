"""
ptz_slam_b200 - B200-native (sm_100a) implementation of the Pan-tilt-zoom-SLAM hot path:
ray-landmark projection, reprojection residuals + analytic Jacobians, normal-equation assembly,
Schur/Cholesky LM solve and the EKF update, behind the reference's Python API.

Modules mirror the reference's names (slam_system/*.py):
    ptz_camera.PTZCamera, transformation.TransFunction, ptz_slam.PtzSlam,
    bundle_adjustment._compute_residual / bundle_adjustment, util.get_overlap_index
All of them call the hand-written CUDA library csrc/libptzba.so through ctypes (`_lib`); there is no CPU
fallback: if the library or a CUDA device is missing the call raises.
"""
__version__ = "0.1.0"
