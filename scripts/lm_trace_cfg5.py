"""Timing aid (GPU box): Levenberg-Marquardt iterations of the cfg5 problem (1024 keyframes, 20M observations); PTZBA_TRACE=1
prints the per-phase times.  Not a test."""
import os, sys, time
import numpy as np  # noqa: F401
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ptz_slam_b200  # noqa: F401
from ptz_slam_b200 import synth, bundle_adjustment as BA
t=time.perf_counter()
fb=synth.make_flat_ba(1024,1000000,20000000,seed=1005,pan_sweep=40.0)
print('gen s', time.perf_counter()-t)
prob=BA.BAProblem(fb.n_pose,fb.n_landmark,fb.cam_idx,fb.lm_idx,fb.obs_xy,640.,360.)
x=fb.x0()
for i in range(2): prob.lm_iteration(x, fb.ptz_init[0], 1e-3)
t=time.perf_counter()
for i in range(3): prob.lm_iteration(x, fb.ptz_init[0], 1e-3)
print('ms/iter', (time.perf_counter()-t)/3*1e3)
