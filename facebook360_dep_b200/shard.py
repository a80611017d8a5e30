"""Multi-GPU plumbing of the hot path: frames are independent inside DerpCLI
(DerpCLI.cpp:229-320; the reference already shards frame chunks across workers, render.py:169-175), so the
path shards by FRAMES with no data-path collective.  This module holds the pieces shared by bench.py and the
multi-process tests: the contiguous frame partition (contiguous so that a temporal filter's +-time_radius
halo only touches the two neighbouring ranks) and the timing reduction (max over ranks) of the bench contract.
"""
import torch
import torch.distributed as dist


def frame_block(num_frames, world_size, rank):
    """Contiguous block [first, last) of rank `rank`: ceil(F/G) frames per rank, like DerpCLI --gpus."""
    per = (num_frames + world_size - 1) // world_size
    first = min(num_frames, rank * per)
    return first, min(num_frames, first + per)


def halo_frames(num_frames, world_size, rank, time_radius):
    """Frames a rank needs from its neighbours for the temporal filter (TemporalBilateralFilter.cpp:96-119):
    returns (needed_from_left, needed_from_right) as lists of frame indices."""
    first, last = frame_block(num_frames, world_size, rank)
    if first >= last:
        return [], []
    left = [f for f in range(first - time_radius, first) if f >= 0]
    right = [f for f in range(last, last + time_radius) if f < num_frames]
    return left, right


def camera_shard(num_cams, world_size, rank):
    """Destination cameras of rank `rank` when ONE frame is spread over the GPUs (fewer
    frames than GPUs, or latency): round-robin, like DerpCLI --gpus with a single frame."""
    return list(range(rank, num_cams, world_size))


def reduce_step(ms_local, units_local, device):
    """Bench contract: time = MAX over ranks, work = SUM over ranks. Returns (ms_max, units_sum)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms_local), float(units_local)
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())
