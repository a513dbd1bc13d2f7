"""
Multi-GPU plumbing (one process per GPU, torchrun-style).  torch.distributed is used only to hand the 128-byte NCCL
unique id from rank 0 to the other ranks; the data-path collective (all-reduce of the packed normal-equation blocks)
is issued by libptzba itself with ncclAllReduce on the library's stream.

Sharding: BA observations are sharded by keyframe - `shard_by_keyframe` gives every rank a contiguous
keyframe range balanced by observation count; independent EKF sequences are sharded round-robin with no collective.
"""
import ctypes

import numpy as np

from . import _lib


class Communicator:
    def __init__(self, ctx, rank, world_size, broadcast=None):
        """`broadcast(bytes_or_None) -> bytes` distributes rank 0's payload; default uses torch.distributed."""
        self.ctx, self.rank, self.world = ctx, int(rank), int(world_size)
        uid = (ctypes.c_ubyte * 128)()
        if self.rank == 0:
            ctx.check(ctx.lib.ptzba_comm_unique_id(ctx.handle, ctypes.cast(uid, ctypes.c_void_p)))
        payload = bytes(uid)
        if broadcast is None:
            import torch.distributed as dist
            box = [payload if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            payload = box[0]
        else:
            payload = broadcast(payload if self.rank == 0 else None)
        buf = (ctypes.c_ubyte * 128).from_buffer_copy(payload)
        ctx.check(ctx.lib.ptzba_comm_init(ctx.handle, ctypes.cast(buf, ctypes.c_void_p), self.rank, self.world))

    def allreduce_landmark_blocks(self, problem):
        """Sum the accumulators of `problem`'s last fused pass over all ranks (asynchronous on the context stream)."""
        self.ctx.check(self.ctx.lib.ptzba_ba_allreduce(problem.handle))

    def setup_exchange(self, problem):
        """Keyframe-sharded mode: find the landmarks shared between ranks once; `allreduce_landmark_blocks` then exchanges
        only their blocks and the cost (the other blocks are complete on the rank that owns them).  Returns their number."""
        n = ctypes.c_int64(0)
        self.ctx.check(self.ctx.lib.ptzba_ba_setup_exchange(problem.handle, ctypes.byref(n)))
        return int(n.value)

    def allreduce(self, dev_ptr, count):
        self.ctx.check(self.ctx.lib.ptzba_comm_allreduce_f64(self.ctx.handle, _lib.ptr(int(dev_ptr)), int(count)))


def keyframe_ranges(cam_idx, n_pose, world_size):
    """Contiguous keyframe ranges [lo, hi) per rank, balanced by observation count."""
    counts = np.bincount(np.asarray(cam_idx), minlength=n_pose).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(counts)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, world_size):
        bounds.append(int(np.searchsorted(csum, total * r / world_size, side="left")))
    bounds.append(n_pose)
    bounds = np.maximum.accumulate(np.clip(bounds, 0, n_pose))
    return [(int(bounds[r]), int(bounds[r + 1])) for r in range(world_size)]


def shard_by_keyframe(cam_idx, lm_idx, obs_xy, n_pose, rank, world_size):
    """Observations whose keyframe falls into this rank's range (order preserved)."""
    lo, hi = keyframe_ranges(cam_idx, n_pose, world_size)[rank]
    cam_idx = np.asarray(cam_idx)
    sel = (cam_idx >= lo) & (cam_idx < hi)
    return cam_idx[sel], np.asarray(lm_idx)[sel], np.asarray(obs_xy)[sel], (lo, hi)


def owned_landmark_range(cam_idx, lm_idx, n_landmark, kf_range):
    """Landmarks OWNED by the rank of keyframe range `kf_range` when the observations are sharded by keyframe: a landmark
    belongs to the rank of the lowest-numbered keyframe that observes it, so every observed landmark has exactly one owner
    and that owner observes it (after the shared-landmark exchange its V / g_l blocks are complete there).  The reference
    numbers landmarks in the order they are first seen (image_process.py:609-667), which makes the owned ids one contiguous
    range: returns (lo, hi) then, None when the ids of this problem are not ordered that way (a caller then falls back to
    the span of the ids the rank observes).  `cam_idx` / `lm_idx` are the WHOLE problem's lists."""
    cam_idx, lm_idx = np.asarray(cam_idx), np.asarray(lm_idx)
    big = np.iinfo(np.int64).max
    first = np.full(int(n_landmark), big, dtype=np.int64)
    np.minimum.at(first, lm_idx, cam_idx.astype(np.int64))
    ids = np.nonzero((first >= kf_range[0]) & (first < kf_range[1]))[0]
    if ids.size == 0:
        return (0, 0)
    lo, hi = int(ids[0]), int(ids[-1]) + 1
    inside = first[lo:hi]
    # unobserved ids inside the range are harmless (their blocks are zero); ids owned by another rank are not
    if np.any((inside != big) & ((inside < kf_range[0]) | (inside >= kf_range[1]))):
        return None
    return (lo, hi)


def solve_partition(lm_idx, n_landmark, world_size):
    """Work slices of the distributed solve (ptzba_ba_set_partition), one per rank: contiguous landmark ranges balanced by
    observation count, and equal contiguous slices of the keyframe-major observation list.
    Returns [((lm_lo, lm_hi), (cm_lo, cm_hi)), ...]; the ranges tile [0, n_landmark) and [0, n_obs)."""
    lm_idx = np.asarray(lm_idx)
    n_obs = int(lm_idx.shape[0])
    counts = np.bincount(lm_idx, minlength=n_landmark).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(counts)])
    lm_bounds = [0]
    for r in range(1, world_size):
        lm_bounds.append(int(np.searchsorted(csum, n_obs * r / world_size, side="left")))
    lm_bounds.append(n_landmark)
    lm_bounds = np.maximum.accumulate(np.clip(lm_bounds, 0, n_landmark))
    cm_bounds = [n_obs * r // world_size for r in range(world_size + 1)]
    return [((int(lm_bounds[r]), int(lm_bounds[r + 1])), (int(cm_bounds[r]), int(cm_bounds[r + 1]))) for r in range(world_size)]


def shard_sequences(n_seq, rank, world_size):
    """Indices of the independent EKF sequences owned by this rank (no collective on this path)."""
    return np.arange(rank, n_seq, world_size)
