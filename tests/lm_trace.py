import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import ptz_slam_b200
from ptz_slam_b200 import synth, bundle_adjustment as BA
fb=synth.make_flat_ba(256,100000,2000000,seed=1003)
prob=BA.BAProblem(fb.n_pose,fb.n_landmark,fb.cam_idx,fb.lm_idx,fb.obs_xy,640.,360.)
x=fb.x0()
for i in range(3): prob.lm_iteration(x, fb.ptz_init[0], 1e-3)
import os
t=time.perf_counter()
for i in range(5): prob.lm_iteration(x, fb.ptz_init[0], 1e-3)
print('ms/iter', (time.perf_counter()-t)/5*1e3)
