"""
ORACLE - TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference hot-path modules from /root/reference/slam_system so that golden vectors can be
generated from the reference itself (tests/golden/make_golden.py).  /root/reference exists only in the authoring
container, never on the GPU box: nothing that runs there may import this module.

Three imports the reference needs are absent from this image and are stubbed:
matplotlib (util.py:15-22), pyflann (nearest_neighbor.py:8) and the ctypes wrapper
rf_map.python_package.backup.rf_map (scene_map.py:15), none of which is on the hot path.
"""
import os
import sys
import types

REFERENCE_DIR = "/root/reference/slam_system"


def available():
    return os.path.isdir(REFERENCE_DIR)


def load():
    """Returns a namespace with PTZCamera, TransFunction, PtzSlam, bundle_adjustment, util, least_squares."""
    if not available():
        raise RuntimeError("reference tree %s is not present (authoring container only)" % REFERENCE_DIR)
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    for n in ("matplotlib", "matplotlib.pyplot", "pyflann"):
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["pyflann"].FLANN = object
    for n in ("rf_map", "rf_map.python_package", "rf_map.python_package.backup",
              "rf_map.python_package.backup.rf_map"):
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__path__ = []
            sys.modules[n] = m
    sys.modules["rf_map.python_package.backup.rf_map"].RFMap = object
    import warnings
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    ns = types.SimpleNamespace()
    from ptz_camera import PTZCamera
    from transformation import TransFunction
    from ptz_slam import PtzSlam
    import bundle_adjustment as ba
    import util
    from scipy.optimize import least_squares
    ns.PTZCamera, ns.TransFunction, ns.PtzSlam = PTZCamera, TransFunction, PtzSlam
    ns.bundle_adjustment, ns.util, ns.least_squares = ba, util, least_squares
    return ns
