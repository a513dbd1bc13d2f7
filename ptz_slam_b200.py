"""
Import alias.  The package directory is named `pan-tilt-zoom-slam_b200/` (not a valid Python identifier), so
`import ptz_slam_b200` loads that directory as the package `ptz_slam_b200` and replaces this module with it.
"""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "pan-tilt-zoom-slam_b200")
_spec = importlib.util.spec_from_file_location(
    "ptz_slam_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ptz_slam_b200"] = _mod
_spec.loader.exec_module(_mod)
