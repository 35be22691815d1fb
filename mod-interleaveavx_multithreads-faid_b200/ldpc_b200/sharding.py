"""Multi-GPU plumbing: which frames a rank owns and the one collective of the path (main.cpp:170-182).

Frames are independent Monte-Carlo trials; the only coupling is inside a group of 32, so groups (never frames) are
sharded.  Philox subsequence = global frame index, hence the result does not depend on the number of ranks."""
import numpy as np

from .abi import CNT_ERROR_FRAME, CNT_TEST_FRAME, NUM_COUNTERS


def shard(rank, world, groups_per_rank, round_index=0):
    """-> (first_group, first_frame_index) of this rank in round `round_index` (weak scaling: fixed work per rank)."""
    first_group = (round_index * world + rank) * groups_per_rank
    return first_group, first_group * 32


def allreduce_counters(counters, dist=None, device=None):
    """Sum the uint64[NUM_COUNTERS] vector over ranks with ONE all-reduce (NCCL on GPU tensors, gloo on CPU)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return counters
    t = torch.from_numpy(counters.astype(np.int64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t)
    return t.cpu().numpy().astype(np.uint64)


def stop_rule(counters, min_frames=1000, min_error_frames=20):
    """main.cpp:164: keep running rounds while TestFrame < 1000 or ErrorFrame < 20 (evaluated on reduced counters)."""
    return not (counters[CNT_TEST_FRAME] < min_frames or counters[CNT_ERROR_FRAME] < min_error_frames)


def empty_counters():
    return np.zeros(NUM_COUNTERS, dtype=np.uint64)
