"""
ctypes binding of csrc/libptzba.so (include/ptzba.h).

Mirrors how the reference reaches native code (rf_map/python_package/rf_map_wrapper.py:14-62): LoadLibrary at first
use, opaque handle, caller-owned numpy buffers passed as raw pointers.  There is NO CPU fallback: a missing library
or a missing CUDA device raises PtzbaError.
"""
import ctypes
import os
import threading

import numpy as np

HOST, DEVICE = 0, 1
JAC_ANALYTIC, JAC_CENTRAL_FD = 0, 1
OPT_SCHUR_MODE = 1
OPT_FUSED_LM_SHARE = 3
SCHUR_AUTO, SCHUR_PER_LANDMARK, SCHUR_PAIR_LIST = 0, 1, 2

# PTZBA_LIBRARY names another build of the same library (kernel tuning experiments); it is still this library or nothing
_LIB_PATH = os.environ.get("PTZBA_LIBRARY") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libptzba.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_void_p = ctypes.c_void_p


class PtzbaError(RuntimeError):
    pass


class EkfParams(ctypes.Structure):
    _fields_ = [("u", ctypes.c_double), ("v", ctypes.c_double), ("disp", ctypes.c_double * 6),
                ("observe_var", ctypes.c_double), ("angle_var", ctypes.c_double), ("f_var", ctypes.c_double),
                ("height", ctypes.c_double), ("width", ctypes.c_double), ("jac_mode", ctypes.c_int)]


class BaOptions(ctypes.Structure):
    _fields_ = [("ftol", ctypes.c_double), ("xtol", ctypes.c_double), ("gtol", ctypes.c_double),
                ("max_nfev", ctypes.c_int), ("verbose", ctypes.c_int)]


class BaReport(ctypes.Structure):
    _fields_ = [("cost0", ctypes.c_double), ("cost", ctypes.c_double), ("optimality", ctypes.c_double),
                ("status", ctypes.c_int), ("nfev", ctypes.c_int), ("njev", ctypes.c_int), ("nit", ctypes.c_int),
                ("n_factor", ctypes.c_int), ("ms_total", ctypes.c_double)]


# name -> (restype, argtypes); every symbol include/ptzba.h declares
_I, _D, _P, _L = ctypes.c_int, ctypes.c_double, c_void_p, ctypes.c_int64
SIGNATURES = {
    "ptzba_version": (_I, []),
    "ptzba_create": (_I, [_I, ctypes.POINTER(_P)]),
    "ptzba_destroy": (None, [_P]),
    "ptzba_last_error": (ctypes.c_char_p, [_P]),
    "ptzba_set_stream": (_I, [_P, _P]),
    "ptzba_synchronize": (_I, [_P]),
    "ptzba_launch_count": (_L, [_P]),
    "ptzba_profile_begin": (_I, [_P]),
    "ptzba_profile_end": (_I, [_P, _P, _P]),
    "ptzba_project": (_I, [_P, _I, _I, _P, _D, _D, _P, _I, _P, _P]),
    "ptzba_project_rays_filtered": (_I, [_P, _I, _P, _D, _D, _P, _I, _P, _D, _D, _P, _P, _P]),
    "ptzba_project_pairs": (_I, [_P, _I, _I, _P, _D, _D, _I, _P, _L, _P, _P, _P]),
    "ptzba_backproject": (_I, [_P, _I, _I, _P, _D, _D, _P, _L, _P, _P, _P]),
    "ptzba_h_jacobian_blocks": (_I, [_P, _I, _P, _D, _D, _P, _I, _P, _I, _P, _P]),
    "ptzba_h_jacobian_dense": (_I, [_P, _I, _P, _D, _D, _P, _I, _P, _I, _P]),
    "ptzba_match_graph_to_observations": (_I, [_P, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P]),
    "ptzba_pose_score": (_I, [_P, _I, _I, _P, _D, _D, _I, _P, _P, _I, _P, _D, _P, _P]),
    "ptzba_pose_refine": (_I, [_P, _I, _I, _P, _D, _D, _I, _P, _P, _I, _P, _D, _I, _I, _D, _P, _P, _P]),
    "ptzba_ekf_update": (_I, [_P, ctypes.POINTER(EkfParams), _I, _P, _P, _P, _P, _I, _P, _P, _P]),
    "ptzba_ekf_batch_create": (_I, [_P, ctypes.POINTER(EkfParams), _I, _I, _I, _P, _P, ctypes.POINTER(_P)]),
    "ptzba_ekf_batch_destroy": (None, [_P]),
    "ptzba_ekf_batch_step": (_I, [_P, _I, _P, _P, _P, _P]),
    "ptzba_ekf_batch_update_only": (_I, [_P, _I, _P, _P, _P, _P]),
    "ptzba_ekf_batch_set": (_I, [_P, _I, _P, _P, _P, _P]),
    "ptzba_ekf_batch_get": (_I, [_P, _P, _P, _P]),
    "ptzba_ekf_batch_get_cov": (_I, [_P, _I, _P]),
    "ptzba_ekf_batch_n_rays": (_I, [_P, _I, _P, _P]),
    "ptzba_ekf_batch_get_rays": (_I, [_P, _I, _P]),
    "ptzba_ekf_batch_remove_rays": (_I, [_P, _I, _I, _P]),
    "ptzba_ekf_batch_add_rays": (_I, [_P, _I, _I, _P]),
    "ptzba_ekf_batch_reserve": (_I, [_P, _I, _I]),
    "ptzba_ekf_batch_predict_cov": (_I, [_P]),
    "ptzba_ekf_batch_max_obs": (_I, [_P, _P]),
    "ptzba_ekf_batch_route": (_I, [_P, _P]),
    "ptzba_ba_create": (_I, [_P, _I, _I, _I, _L, _P, _P, _P, _D, _D, ctypes.POINTER(_P)]),
    "ptzba_ba_destroy": (None, [_P]),
    "ptzba_ba_residual": (_I, [_P, _I, _P, _P, _P]),
    "ptzba_ba_normal_equations": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ptzba_ba_solve": (_I, [_P, _I, _P, _P, ctypes.POINTER(BaOptions), ctypes.POINTER(BaReport)]),
    "ptzba_ba_lm_iteration": (_I, [_P, _I, _P, _P, _D, _P, _P]),
    "ptzba_comm_unique_id": (_I, [_P, _P]),
    "ptzba_comm_init": (_I, [_P, _P, _I, _I]),
    "ptzba_comm_allreduce_f64": (_I, [_P, _P, _L]),
    "ptzba_dense_solve_spd": (_I, [_P, _I, _P, _P, _P, _P]),
    "ptzba_ba_allreduce": (_I, [_P]),
    "ptzba_ba_setup_exchange": (_I, [_P, _P]),
    "ptzba_ba_set_partition": (_I, [_P, _I, _I, _I, _I, _L, _L]),
    "ptzba_ba_get_blocks": (_I, [_P, _P, _P, _P, _P, _P]),
    "ptzba_ba_set_option": (_I, [_P, _I, _I]),
    "ptzba_ba_normal_equations_begin": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "ptzba_ba_wait": (_I, [_P]),
}

_lock = threading.Lock()
_cdll = None
_contexts = {}


def library_path():
    return _LIB_PATH


def load_library():
    """dlopen libptzba.so and declare every prototype.  Raises PtzbaError if it has not been built."""
    global _cdll
    with _lock:
        if _cdll is None:
            if not os.path.exists(_LIB_PATH):
                raise PtzbaError("%s not found: build it with `python pan-tilt-zoom-slam_b200/build.py` "
                                 "(there is no CPU fallback)" % _LIB_PATH)
            lib = ctypes.CDLL(_LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _cdll = lib
    return _cdll


class Context:
    """Opaque ptzba_ctx handle for one GPU."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = c_void_p()
        st = self.lib.ptzba_create(int(device), ctypes.byref(h))
        if st != 0:
            raise PtzbaError("ptzba_create(device=%d) failed with status %d: no usable CUDA device "
                             "(this library has no CPU fallback)" % (device, st))
        self.handle = h
        self.device = device

    def check(self, status):
        if status != 0:
            msg = self.lib.ptzba_last_error(self.handle)
            raise PtzbaError("libptzba status %d: %s" % (status, msg.decode() if msg else "?"))

    def set_stream(self, stream_ptr):
        self.check(self.lib.ptzba_set_stream(self.handle, c_void_p(stream_ptr)))

    def synchronize(self):
        self.check(self.lib.ptzba_synchronize(self.handle))

    def launch_count(self):
        return int(self.lib.ptzba_launch_count(self.handle))

    def close(self):
        if self.handle:
            self.lib.ptzba_destroy(self.handle)
            self.handle = None


def get_context(device=None):
    """Process-wide context per device (created on first use)."""
    if device is None:
        device = int(os.environ.get("PTZBA_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    with _lock:
        ctx = _contexts.get(device)
    if ctx is None:
        ctx = Context(device)
        with _lock:
            _contexts[device] = ctx
    return ctx


def f64(a):
    """C-contiguous float64 view/copy of a."""
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def ptr(a):
    """Raw pointer of a numpy array (or None); cf. c_void_p(arr.ctypes.data) in rf_map_wrapper.py:51-62."""
    if a is None:
        return None
    if isinstance(a, int):
        return c_void_p(a)
    return c_void_p(a.ctypes.data)
