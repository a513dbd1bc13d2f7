"""CPU, world_size 2 over gloo: the keyframe sharding of the multi-GPU BA path is a partition and the all-reduced
shard blocks equal the single-process normal equations (host logic of ptz_slam_b200.dist; the GPU path replaces the
oracle call by the CUDA fused pass and gloo by ncclAllReduce)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ptz_slam_b200  # noqa: F401
from ptz_slam_b200 import synth
from ptz_slam_b200 import dist as pdist
from oracle import ptz_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fb = synth.make_flat_ba(12, 300, 2400, seed=77)
    cam, lm, xy, (lo, hi) = pdist.shard_by_keyframe(fb.cam_idx, fb.lm_idx, fb.obs_xy, fb.n_pose, rank, world)
    poses, rays = O.ba_unpack(fb.x0(), fb.n_pose, fb.ptz_init[0])
    r, U, gc, V, gl, cost = O.ba_normal_equations(poses, rays, cam, lm, xy, synth.PP_U, synth.PP_V)
    U[0] = 0; gc[0] = 0          # fixed reference pose, as the CUDA pass
    packed = torch.from_numpy(np.concatenate([[cost], U.ravel(), V.ravel(), gc.ravel(), gl.ravel()]))
    dist.all_reduce(packed)      # the data-path collective: one sum over the packed blocks
    n_obs = torch.tensor([len(cam)])
    dist.all_reduce(n_obs)
    seqs = pdist.shard_sequences(11, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, seqs.tolist())
    if rank == 0:
        np.save(os.path.join(out_dir, "packed.npy"), packed.numpy())
        np.save(os.path.join(out_dir, "meta.npy"), np.array([int(n_obs.item()), lo, hi]))
        np.save(os.path.join(out_dir, "seqs.npy"), np.array(sorted(sum(gathered, []))))
    dist.destroy_process_group()


def _worker_partition(rank, world, port, out_dir):
    """Distributed-solve partition (replicated data, partitioned work): landmark slice -> V, g_l, cost; slice of the
    keyframe-major list -> U, g_c; one all-reduce of the packed blocks reproduces the full normal equations."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fb = synth.make_flat_ba(12, 300, 2400, seed=78)
    (lm_lo, lm_hi), (cm_lo, cm_hi) = pdist.solve_partition(fb.lm_idx, fb.n_landmark, world)[rank]
    poses, rays = O.ba_unpack(fb.x0(), fb.n_pose, fb.ptz_init[0])
    sel_l = (fb.lm_idx >= lm_lo) & (fb.lm_idx < lm_hi)
    _, _, _, V, gl, cost = O.ba_normal_equations(poses, rays, fb.cam_idx[sel_l], fb.lm_idx[sel_l], fb.obs_xy[sel_l], synth.PP_U, synth.PP_V)
    order = np.argsort(fb.cam_idx, kind="stable")[cm_lo:cm_hi]          # this rank's slice of the keyframe-major list
    _, U, gc, _, _, _ = O.ba_normal_equations(poses, rays, fb.cam_idx[order], fb.lm_idx[order], fb.obs_xy[order], synth.PP_U, synth.PP_V)
    U[0] = 0; gc[0] = 0
    packed = torch.from_numpy(np.concatenate([[cost], U.ravel(), V.ravel(), gc.ravel(), gl.ravel()]))
    dist.all_reduce(packed)
    if rank == 0:
        np.save(os.path.join(out_dir, "packed_part.npy"), packed.numpy())
    dist.destroy_process_group()


def test_solve_partition_tiles():
    rng = np.random.default_rng(1)
    lm = np.sort(rng.integers(0, 700, 20000))
    for w in (1, 2, 3, 8):
        parts = pdist.solve_partition(lm, 700, w)
        assert parts[0][0][0] == 0 and parts[-1][0][1] == 700 and parts[0][1][0] == 0 and parts[-1][1][1] == len(lm)
        assert all(parts[i][0][1] == parts[i + 1][0][0] and parts[i][1][1] == parts[i + 1][1][0] for i in range(w - 1))
        counts = [np.sum((lm >= a) & (lm < b)) for (a, b), _ in parts]
        assert sum(counts) == len(lm) and max(counts) < 1.2 * len(lm) / w + 100


def test_world2_partitioned_blocks_sum_to_full(tmp_path):
    world = 2
    mp.spawn(_worker_partition, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    packed = np.load(tmp_path / "packed_part.npy")
    fb = synth.make_flat_ba(12, 300, 2400, seed=78)
    poses, rays = O.ba_unpack(fb.x0(), fb.n_pose, fb.ptz_init[0])
    r, U, gc, V, gl, cost = O.ba_normal_equations(poses, rays, fb.cam_idx, fb.lm_idx, fb.obs_xy, synth.PP_U, synth.PP_V)
    U[0] = 0; gc[0] = 0
    full = np.concatenate([[cost], U.ravel(), V.ravel(), gc.ravel(), gl.ravel()])
    np.testing.assert_allclose(packed, full, rtol=1e-11, atol=1e-9 * np.abs(full).max())


def test_keyframe_ranges_partition():
    rng = np.random.default_rng(0)
    cam = rng.integers(0, 50, 10000)
    for w in (1, 2, 4, 8):
        ranges = pdist.keyframe_ranges(cam, 50, w)
        assert ranges[0][0] == 0 and ranges[-1][1] == 50
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(w - 1))
        counts = [np.sum((cam >= lo) & (cam < hi)) for lo, hi in ranges]
        assert sum(counts) == len(cam) and max(counts) < 1.5 * len(cam) / w + 400


def test_world2_sharded_blocks_sum_to_full(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    packed = np.load(tmp_path / "packed.npy")
    meta = np.load(tmp_path / "meta.npy")
    fb = synth.make_flat_ba(12, 300, 2400, seed=77)
    assert meta[0] == fb.n_obs
    poses, rays = O.ba_unpack(fb.x0(), fb.n_pose, fb.ptz_init[0])
    r, U, gc, V, gl, cost = O.ba_normal_equations(poses, rays, fb.cam_idx, fb.lm_idx, fb.obs_xy, synth.PP_U, synth.PP_V)
    U[0] = 0; gc[0] = 0
    full = np.concatenate([[cost], U.ravel(), V.ravel(), gc.ravel(), gl.ravel()])
    np.testing.assert_allclose(packed, full, rtol=1e-11, atol=1e-9 * np.abs(full).max())
    np.testing.assert_array_equal(np.load(tmp_path / "seqs.npy"), np.arange(11))


def test_owned_landmark_ranges_tile_the_landmarks():
    """Keyframe-sharded mode: every observed landmark has exactly one owner (the rank of its first keyframe), the owner observes
    it, and with landmark ids numbered in first-seen order (image_process.py:609-667) the owned ids are contiguous ranges that
    tile [0, n_landmark) in rank order.  Random ids: no contiguous ownership, the function says so."""
    fb = synth.make_flat_ba(24, 3000, 30000, seed=5)
    for world in (2, 3, 8):
        prev = 0
        for rank in range(world):
            cam, lm, xy, kr = pdist.shard_by_keyframe(fb.cam_idx, fb.lm_idx, fb.obs_xy, fb.n_pose, rank, world)
            own = pdist.owned_landmark_range(fb.cam_idx, fb.lm_idx, fb.n_landmark, kr)
            assert own is not None
            if own == (0, 0):                                               # a rank whose keyframes see nothing new
                continue
            assert own[0] == prev
            prev = own[1]
            ids = np.arange(*own)
            assert np.isin(ids, np.unique(lm)).all()                        # the owner observes every landmark it owns
            first = np.array([fb.cam_idx[fb.lm_idx == i].min() for i in ids[:: max(1, len(ids) // 50)]])
            assert ((first >= kr[0]) & (first < kr[1])).all()
        assert prev == int(fb.lm_idx.max()) + 1                              # (ids nobody observes come last)
    perm = np.random.default_rng(0).permutation(fb.n_landmark)
    kr = pdist.keyframe_ranges(fb.cam_idx, fb.n_pose, 2)[0]                 # owns some but not all landmarks (checked above)
    lo, hi = pdist.owned_landmark_range(fb.cam_idx, fb.lm_idx, fb.n_landmark, kr)
    assert 0 < hi - lo < int(fb.lm_idx.max()) + 1
    assert pdist.owned_landmark_range(fb.cam_idx, perm[fb.lm_idx], fb.n_landmark, kr) is None
    assert pdist.owned_landmark_range(fb.cam_idx, fb.lm_idx, fb.n_landmark, (fb.n_pose, fb.n_pose)) == (0, 0)


def _worker_ramp(rank, world, port, out_dir):
    """bench.clock_ramp with a collective inside every step and ranks of very different speed."""
    import time
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    acc = torch.zeros(1, dtype=torch.float64)

    def step(i):
        if rank == 1 and i % 8 == 0:
            time.sleep(0.004)            # a slow rank: its own clock would end the ramp after far fewer steps
        t = torch.ones(1, dtype=torch.float64)
        dist.all_reduce(t)               # the collective a sharded pass ends with
        acc.add_(t)

    n = bench.clock_ramp(step, 0.15, world, lambda: None, "cpu", block=16)
    np.save(os.path.join(out_dir, "ramp_%d.npy" % rank), np.array([n, acc.item()]))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_clock_ramp_runs_the_same_number_of_steps_on_every_rank(tmp_path):
    world = 2
    mp.spawn(_worker_ramp, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "ramp_0.npy"), np.load(tmp_path / "ramp_1.npy")
    assert r0[0] == r1[0] and r0[0] >= 16 and r0[0] % 16 == 0
    assert r0[1] == r1[1] == world * r0[0]
