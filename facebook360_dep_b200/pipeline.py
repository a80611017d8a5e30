"""Frame-sharded temporal filtering (BASELINE.json configs[4]: 16-camera x 30-frame sequence with temporal
depth smoothing on 8 GPUs) — the one stage of the path with a real cross-rank exchange.

DerpCLI's frames are independent, so ranks own contiguous frame blocks (shard.frame_block) and estimate their
disparities with no communication.  TemporalBilateralFilter then needs, for every frame, the colour, disparity
and mask of the +-time_radius neighbouring frames (TemporalBilateralFilter.cpp:96-160): frames near a block
boundary live on the neighbouring rank(s).  The reference moves them as files (pipeline.py:382-408); here the
halo frames travel rank-to-rank with point-to-point send/recv on the process group — NCCL over NVLink when the
tensors are CUDA tensors (one process per GPU), gloo in the CPU tests.  The filter itself is
derp_temporal_filter of whichever library is passed in (CUDA product / CPU oracle in tests).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import shard


def _pack(frame, H, W):
    """frame = list over cameras of (color u16 HxWx3, disparity f32 HxW, mask u8 HxW) -> flat uint8 tensor."""
    parts = []
    for color, disp, mask in frame:
        parts.append(np.ascontiguousarray(color, np.uint16).view(np.uint8).ravel())
        parts.append(np.ascontiguousarray(disp, np.float32).view(np.uint8).ravel())
        parts.append(np.ascontiguousarray(mask, np.uint8).ravel())
    return torch.from_numpy(np.concatenate(parts))


def _unpack(buf, num_cams, H, W):
    a = buf.cpu().numpy()
    out, o = [], 0
    n = H * W
    for _ in range(num_cams):
        color = a[o:o + n * 6].view(np.uint16).reshape(H, W, 3)
        o += n * 6
        disp = a[o:o + n * 4].view(np.float32).reshape(H, W)
        o += n * 4
        mask = a[o:o + n].reshape(H, W)
        o += n
        out.append((color, disp, mask))
    return out


def exchange_halos(local_frames, num_frames, time_radius, device):
    """local_frames: {frame_index: [(color, disp, mask) per camera]} for this rank's block.
    Returns a dict with the halo frames this rank needs from other ranks (empty when world size is 1)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return {}
    world, rank = dist.get_world_size(), dist.get_rank()
    first, last = shard.frame_block(num_frames, world, rank)
    if not local_frames:  # more ranks than frames: an idle rank owns nothing, so it neither sends nor receives
        return {}
    some = next(iter(local_frames.values()))
    num_cams = len(some)
    H, W = some[0][1].shape
    nbytes = num_cams * H * W * 11

    def owner(f):
        per = (num_frames + world - 1) // world
        return f // per

    ops, recv_bufs, keep = [], {}, []
    # what every other rank needs from me / what I need from them is a pure function of (F, world, radius)
    for r in range(world):
        left, right = shard.halo_frames(num_frames, world, r, time_radius)
        for f in left + right:
            o = owner(f)
            if o == rank and r != rank:
                t = _pack(local_frames[f], H, W).to(device)
                keep.append(t)
                ops.append(dist.P2POp(dist.isend, t, r))
            elif r == rank and o != rank:
                t = torch.empty(nbytes, dtype=torch.uint8, device=device)
                recv_bufs[f] = t
                ops.append(dist.P2POp(dist.irecv, t, o))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return {f: _unpack(t, num_cams, H, W) for f, t in recv_bufs.items()}


def temporal_filter_block(lib, local_frames, num_frames, time_radius=2, sigma=0.01, space_radius=1,
                          weight_b=0.5, weight_g=1.0, device=None, gpu=0):
    """Temporal joint-bilateral filtering of this rank's frames.  Weights follow the reference app:
    (weight_b, weight_g, weight_b) (TemporalBilateralFilter.cpp:176-178).  Returns {frame: [filtered per camera]}."""
    device = device or torch.device("cpu")
    halos = exchange_halos(local_frames, num_frames, time_radius, device)
    frames = dict(local_frames)
    frames.update(halos)
    out = {}
    for f in sorted(local_frames):
        lo, hi = max(0, f - time_radius), min(num_frames - 1, f + time_radius)
        window = [frames[t] for t in range(lo, hi + 1)]
        res = []
        for cam in range(len(local_frames[f])):
            guides = [w[cam][0] for w in window]
            disps = [w[cam][1] for w in window]
            masks = [w[cam][2] for w in window]
            res.append(lib.temporal_filter(guides, disps, masks, f - lo, sigma, space_radius, weight_b, weight_g,
                                           weight_b, device=gpu))
        out[f] = res
    return out


def sharded_mismatches(ctx, num_cams, device=None):
    """Mismatch handling (handleDisparityMismatches, Derp.cpp:685-748) when the destination cameras of a frame are
    dealt round-robin to the ranks (shard.camera_shard): the Jacobi update reads every camera's pre-update
    disparity, so the stage is ONE all-gather of the per-camera planes followed by the kernel on the rank's own
    destinations.  `ctx` is the rank's context, created with dst list
    shard.camera_shard(num_cams, world, rank), after level_estimate.  NCCL when `device` is a CUDA device (the
    planes are copied device-to-device into the collective's buffers), gloo on CPU with the oracle in tests."""
    device = device or torch.device("cpu")
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    own = shard.camera_shard(num_cams, world, rank)
    assert ctx.Sd == len(own)
    per = (num_cams + world - 1) // world
    send = torch.zeros((per, ctx.H, ctx.W), dtype=torch.float32, device=device)
    if device.type == "cuda":
        torch.cuda.synchronize(device)  # the library copies on its own stream: the fill must have landed
    for i in range(len(own)):
        ctx.L.check(ctx.L.lib.derp_get_disparity(ctx.h, i, send[i].data_ptr(), None, None))
    if world > 1:
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(recv, send)
    else:
        recv = [send]
    if device.type == "cuda":
        torch.cuda.synchronize(device)
    planes = [None] * num_cams
    for r in range(world):
        for i, cam in enumerate(shard.camera_shard(num_cams, world, r)):
            planes[cam] = recv[r][i].data_ptr()
    ctx.gather_disparities(planes)
    ctx.mismatches_gathered()


# ---- device-resident variants (one process per GPU, NCCL over NVLink): nothing below touches host memory ------------
def exchange_halos_device(local, num_frames, time_radius):
    """local: {frame: (color u16 [S,H,W,3], disp f32 [S,H,W])} CUDA tensors of this rank's frame block.  Moves the
    +-time_radius boundary frames rank to rank with NCCL point-to-point ops straight from / into device memory (the
    reference moves them as files between workers, scripts/render/pipeline.py:382-408).  Returns the halo frames and
    the number of bytes this rank received."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1 or not local:
        return {}, 0
    world, rank = dist.get_world_size(), dist.get_rank()
    some = next(iter(local.values()))
    per = (num_frames + world - 1) // world
    ops, recv = [], {}

    def raw(t):  # NCCL has no 16-bit unsigned type: the planes travel as bytes
        return t.view(torch.uint8) if t.dtype == torch.uint16 else t

    for r in range(world):
        left, right = shard.halo_frames(num_frames, world, r, time_radius)
        for f in left + right:
            o = f // per
            if o == rank and r != rank:
                for t in local[f]:
                    ops.append(dist.P2POp(dist.isend, raw(t), r))
            elif r == rank and o != rank:
                bufs = tuple(torch.empty_like(t) for t in some)
                recv[f] = bufs
                for t in bufs:
                    ops.append(dist.P2POp(dist.irecv, raw(t), o))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    nbytes = sum(t.numel() * t.element_size() for bufs in recv.values() for t in bufs)
    return recv, nbytes


def temporal_filter_block_device(lib, local, num_frames, fov_masks, time_radius=2, sigma=0.01, space_radius=1,
                                 weight_b=0.5, weight_g=1.0, gpu=0):
    """Temporal joint-bilateral filter of this rank's frames, device memory in and out.  local: {frame: (color, disp)}
    as in exchange_halos_device; fov_masks: uint8 [S,H,W] CUDA tensor (mask of every frame = the camera's FOV mask,
    TemporalBilateralFilter.cpp:150-160 without foreground masks).  Returns ({frame: filtered f32 [S,H,W]}, halo bytes)."""
    import ctypes as C
    halos, nbytes = exchange_halos_device(local, num_frames, time_radius)
    frames = dict(local)
    frames.update(halos)
    out = {}
    for f in sorted(local):
        lo, hi = max(0, f - time_radius), min(num_frames - 1, f + time_radius)
        window = [frames[t] for t in range(lo, hi + 1)]
        S, H, W = local[f][1].shape
        T = len(window)
        res = torch.empty_like(local[f][1])
        for cam in range(S):
            guides = (C.c_void_p * T)(*[w[0][cam].data_ptr() for w in window])
            disps = (C.c_void_p * T)(*[w[1][cam].data_ptr() for w in window])
            masks = (C.c_void_p * T)(*([fov_masks[cam].data_ptr()] * T))
            lib.check(lib.lib.derp_temporal_filter(gpu, W, H, T, guides, disps, masks, f - lo, sigma, space_radius,
                                                   weight_b, weight_g, weight_b, res[cam].data_ptr()))
        out[f] = res
    return out, nbytes


def all_gather_disparities_device(ctx, num_cams, device):
    """The exchange step of camera-sharded mismatch handling on its own (what sharded_mismatches does before the kernel),
    returning the bytes this rank received: one NCCL all-gather of the per-camera disparity planes, device to device."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    own = shard.camera_shard(num_cams, world, rank)
    per = (num_cams + world - 1) // world
    send = torch.zeros((per, ctx.H, ctx.W), dtype=torch.float32, device=device)
    torch.cuda.synchronize(device)
    for i in range(len(own)):
        ctx.L.check(ctx.L.lib.derp_get_disparity(ctx.h, i, send[i].data_ptr(), None, None))
    if world > 1:
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(recv, send)
    else:
        recv = [send]
    torch.cuda.synchronize(device)
    planes = [None] * num_cams
    for r in range(world):
        for i, cam in enumerate(shard.camera_shard(num_cams, world, r)):
            planes[cam] = recv[r][i].data_ptr()
    ctx.gather_disparities(planes)
    return (world - 1) * send.numel() * 4
