"""CPU check of the ALGORITHM behind the keyframe-pair-major Schur formation, the default route (csrc/ba_solver.cu: k_pl_emit, k_schur_pairlist):
listing every (observation, observation) pair of every landmark by keyframe pair - weight 2 for distinct positions of the same
keyframe, lower triangle only on diagonal blocks - and summing Y_max W_min^T per key reproduces the dense Schur complement
S = U - W (V + alpha I)^-1 W^T computed from the oracle's analytic Jacobian blocks, duplicates included.  The CUDA kernel itself
is exercised by every GPU solver test (and against the per-landmark route in tests/test_gpu_ba.py::test_schur_per_landmark_route_equals_pair_list)."""
import numpy as np

from oracle import ptz_oracle as O


def test_pairlist_formulation_matches_dense_schur_complement():
    rng = np.random.default_rng(3)

    N, M = 6, 40
    poses = np.stack([rng.uniform(50, 60, N), rng.uniform(-10, -8, N), rng.uniform(2500, 3500, N)], 1)
    rays = np.stack([rng.uniform(48, 62, M), rng.uniform(-12, -6, M)], 1)
    cam, lm = [], []
    for l in range(M):
        cs = rng.choice(N, rng.integers(2, 6), replace=False)
        for c in cs:
            cam.append(c); lm.append(l)
            if rng.uniform() < 0.3:          # duplicate observation of the same (keyframe, landmark)
                cam.append(c); lm.append(l)
    cam, lm = np.array(cam), np.array(lm)
    order = np.argsort(lm, kind='stable'); cam, lm = cam[order], lm[order]
    n_obs = len(cam)
    # dense J: residual rows 2 per obs; params: 3 per cam (incl cam 0), 2 per landmark
    Jc_all = np.zeros((2 * n_obs, 3 * N)); Jl_all = np.zeros((2 * n_obs, 2 * M))
    Wobs = []
    for k in range(n_obs):
        Jc, Jr = O.jacobian_blocks_analytic(poses[cam[k], 0], poses[cam[k], 1], poses[cam[k], 2], rays[lm[k], 0], rays[lm[k], 1])
        Jc_all[2*k:2*k+2, 3*cam[k]:3*cam[k]+3] = Jc
        Jl_all[2*k:2*k+2, 2*lm[k]:2*lm[k]+2] = Jr
        Wobs.append(Jc.T @ Jr)
    U = Jc_all.T @ Jc_all; V = Jl_all.T @ Jl_all; W = Jc_all.T @ Jl_all
    alpha = 0.37
    Vd = V + alpha * np.eye(2 * M)
    Sfull = U - W @ np.linalg.inv(Vd) @ W.T
    Sref = Sfull[3:, 3:]                                  # free cameras 1..N-1
    # --- emulate k_pl_emit / sort / k_schur_pairlist
    keys, vals = [], []
    ptr = np.searchsorted(lm, np.arange(M + 1))
    for l in range(M):
        b, e = ptr[l], ptr[l + 1]
        for i in range(b, e):
            if cam[i] <= 0: continue
            for j in range(b, i + 1):
                if cam[j] <= 0: continue
                hi, lo = max(cam[i], cam[j]), min(cam[i], cam[j])
                keys.append(hi << 16 | lo); vals.append(l | (0x80000000 if (i != j and cam[i] == cam[j]) else 0))
    keys, vals = np.array(keys, np.uint64), np.array(vals, np.uint64)
    o = np.argsort(keys, kind='stable'); keys, vals = keys[o], vals[o]
    S = np.zeros((3 * (N - 1), 3 * (N - 1)))
    for c in range(1, N):
        S[3*(c-1):3*c, 3*(c-1):3*c] = U[3*c:3*c+3, 3*c:3*c+3]
    for k, v in zip(keys, vals):
        cr, cc = int(k >> 16), int(k & 0xffff); l = int(v & 0x7fffffff); wt = 2.0 if (v >> 31) else 1.0
        def Wof(c):
            Jc, Jr = O.jacobian_blocks_analytic(poses[c, 0], poses[c, 1], poses[c, 2], rays[l, 0], rays[l, 1])
            return Jc.T @ Jr
        Vinv = wt * np.linalg.inv(Vd[2*l:2*l+2, 2*l:2*l+2])
        B = Wof(cr) @ Vinv @ Wof(cc).T
        for r in range(3):
            for sc in range(3):
                if not (cr == cc and r < sc):
                    S[3*(cr-1)+r, 3*(cc-1)+sc] -= B[r, sc]
    L = np.tril(S); Lref = np.tril(Sref)
    assert np.abs(L - Lref).max() <= 1e-9 * np.abs(Lref).max()
