"""Small batched-EKF driver used for ncu launch lists (not a pytest file)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ptz_slam_b200  # noqa
from ptz_slam_b200 import synth, _lib
from ptz_slam_b200.ptz_slam import BatchedEkfTracker

n_seq, n_rays, n_frames = 8, 2000, 3
seqs = [synth.make_ekf_sequence(n_rays, n_frames + 1, seed=2000 + i) for i in range(n_seq)]
max_obs = max(len(i) for q in seqs for i in q.obs_idx)
trk = BatchedEkfTracker(np.stack([q.rays0 for q in seqs]), np.stack([q.ptz_gt[0] for q in seqs]), synth.PP_U, synth.PP_V, max_obs,
                        synth.IMAGE_H, synth.IMAGE_W, jacobian_mode=_lib.JAC_CENTRAL_FD)
for k in range(1, n_frames + 1):
    t0 = time.perf_counter()
    m = trk.step(*trk.pack_observations([q.obs_xy[k] for q in seqs], [q.obs_idx[k] for q in seqs]))
    print("frame", k, "matched", m.tolist(), "ms", 1e3 * (time.perf_counter() - t0))
trk.close()
