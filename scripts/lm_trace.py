"""Timing aid (GPU box): 5 Levenberg-Marquardt iterations of the cfg3 problem; with PTZBA_TRACE=1 the library prints the
CUDA-event time of every solver phase to stderr.  Not a test."""
import os, sys, time
import numpy as np  # noqa: F401
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ptz_slam_b200  # noqa: F401
from ptz_slam_b200 import synth, bundle_adjustment as BA
fb=synth.make_flat_ba(256,100000,2000000,seed=1003)
prob=BA.BAProblem(fb.n_pose,fb.n_landmark,fb.cam_idx,fb.lm_idx,fb.obs_xy,640.,360.)
x=fb.x0()
for i in range(3): prob.lm_iteration(x, fb.ptz_init[0], 1e-3)
t=time.perf_counter()
for i in range(5): prob.lm_iteration(x, fb.ptz_init[0], 1e-3)
print('ms/iter', (time.perf_counter()-t)/5*1e3)
