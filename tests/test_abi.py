"""CPU: the C-ABI library loads and exports every symbol include/ptzba.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

import ptz_slam_b200  # noqa: F401
from ptz_slam_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ptzba.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ptzba_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_symbol():
    if not os.path.exists(_lib.library_path()):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.library_path())
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert _lib.load_library().ptzba_version() == 100


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.PtzbaError):
        _lib.Context(0)
    from ptz_slam_b200.ptz_camera import PTZCamera
    import numpy as np
    cam = PTZCamera((640.0, 360.0), np.zeros(3), np.eye(3))
    with pytest.raises(_lib.PtzbaError):
        cam.project_ray([1.0, 2.0])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pan-tilt-zoom-slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
