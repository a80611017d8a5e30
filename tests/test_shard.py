"""N>1 path on CPU: world_size-2 gloo processes run the frame sharding + timing reduction that bench.py uses
under torchrun, and (with the oracle standing in for the per-rank library) check that sharded frames give the
same per-frame results as a single rank."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from facebook360_dep_b200 import shard
from tests import oracle_libs


def test_frame_blocks_cover_and_are_contiguous():
    for F in (1, 2, 7, 30, 31):
        for G in (1, 2, 4, 8):
            seen = []
            for r in range(G):
                a, b = shard.frame_block(F, G, r)
                assert 0 <= a <= b <= F
                seen += list(range(a, b))
            assert seen == list(range(F))
    # 30 frames on 8 GPUs, time_radius 2: halos only touch neighbours
    left, right = shard.halo_frames(30, 8, 3, 2)
    a, b = shard.frame_block(30, 8, 3)
    assert left == [a - 2, a - 1] and right == [b, b + 1]
    assert shard.halo_frames(30, 8, 0, 2)[0] == []


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from facebook360_dep_b200 import capi, synth
    oracle = oracle_libs.load_oracle()
    oracle.set_threads(2)
    F, W, H = 3, 40, 40
    rig = synth.ring_rig(4, W, H, kind="FTHETA")
    first, last = shard.frame_block(F, world, rank)
    ctx = capi.Context(oracle, capi.rig_descs(rig))
    units = 0
    for f in range(first, last):
        colors, _ = synth.render_rig(rig, W, H, scene=synth.Scene(seed=42 + f))
        ctx.level_begin(W, H)
        ctx.set_colors(colors)
        ctx.reproject(1)
        ctx.brute_force(1, num_depths=12, want_index=False)
        units += ctx.get_counters()[0]
        np.save(os.path.join(out_dir, "frame%d.npy" % f), ctx.get_disparity(1, want_cost=False))
    ms, total = shard.reduce_step(10.0 + rank, units, torch.device("cpu"))
    if rank == 0:
        np.save(os.path.join(out_dir, "reduce.npy"), np.array([ms, total]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_frame_sharding(tmp_path, oracle):
    world = 2
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    ms, total = np.load(tmp_path / "reduce.npy")
    assert ms == 11.0  # max over ranks
    # single-process reference of the same frames
    from facebook360_dep_b200 import capi, synth
    rig = synth.ring_rig(4, 40, 40, kind="FTHETA")
    ctx = capi.Context(oracle, capi.rig_descs(rig))
    units = 0
    for f in range(3):
        colors, _ = synth.render_rig(rig, 40, 40, scene=synth.Scene(seed=42 + f))
        ctx.level_begin(40, 40)
        ctx.set_colors(colors)
        ctx.reproject(1)
        ctx.brute_force(1, num_depths=12, want_index=False)
        units += ctx.get_counters()[0]
        got = np.load(tmp_path / ("frame%d.npy" % f))
        ref = ctx.get_disparity(1, want_cost=False)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert total == units  # sum over ranks


def _make_sequence(F, S, H, W):
    rng = np.random.RandomState(123)
    base = rng.randint(0, 65536, (S, H, W, 3))
    frames = {}
    for f in range(F):
        cams = []
        for s in range(S):
            color = np.clip(base[s] + rng.randint(-500, 500, (H, W, 3)), 0, 65535).astype(np.uint16)
            disp = rng.uniform(1e-3, 2.0, (H, W)).astype(np.float32)
            mask = (rng.uniform(size=(H, W)) > 0.1).astype(np.uint8)
            cams.append((color, disp, mask))
        frames[f] = cams
    return frames


def _temporal_worker(rank, world, port, out_dir, F=7):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from facebook360_dep_b200 import capi, pipeline
    oracle = oracle_libs.load_oracle()
    oracle.set_threads(2)
    S, H, W = 2, 24, 28
    seq = _make_sequence(F, S, H, W)
    first, last = shard.frame_block(F, world, rank)
    local = {f: seq[f] for f in range(first, last)}
    out = pipeline.temporal_filter_block(oracle, local, F, time_radius=2)
    for f, cams in out.items():
        np.save(os.path.join(out_dir, "tf%d.npy" % f), np.stack(cams))
    dist.barrier()
    dist.destroy_process_group()


def test_temporal_halo_exchange_gloo(tmp_path, oracle):
    """configs[4] in miniature: 7 frames over 3 ranks (blocks of 3,3,1: halos span one and two ranks), the filtered
    frames must equal the single-process result bit for bit."""
    from facebook360_dep_b200 import pipeline
    world = 3
    port = 29600 + (os.getpid() % 1000)
    mp.spawn(_temporal_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    F = 7
    seq = _make_sequence(F, 2, 24, 28)
    ref = pipeline.temporal_filter_block(oracle, seq, F, time_radius=2)  # world size 1: no exchange
    for f in range(F):
        got = np.load(tmp_path / ("tf%d.npy" % f))
        assert np.array_equal(got.view(np.uint32), np.stack(ref[f]).view(np.uint32)), f


def _temporal_device_worker(rank, world, port, out_dir, F=7):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from facebook360_dep_b200 import pipeline
    oracle = oracle_libs.load_oracle()
    oracle.set_threads(2)
    S, H, W = 2, 24, 28
    seq = _make_sequence(F, S, H, W)
    first, last = shard.frame_block(F, world, rank)
    # the tensor form the GPU pipeline uses (bench.py --workload cfg5), here on CPU tensors over gloo
    local = {f: (torch.from_numpy(np.stack([c[0] for c in seq[f]])), torch.from_numpy(np.stack([c[1] for c in seq[f]])))
             for f in range(first, last)}
    masks = torch.from_numpy(np.stack([c[2] for c in seq[0]]))  # one mask per camera for every frame (FOV mask)
    out, nbytes = pipeline.temporal_filter_block_device(oracle, local, F, masks, time_radius=2)
    for f, t in out.items():
        np.save(os.path.join(out_dir, "tfd%d.npy" % f), t.numpy())
    np.save(os.path.join(out_dir, "bytes%d.npy" % rank), np.array([nbytes]))
    dist.barrier()
    dist.destroy_process_group()


def test_device_resident_halo_exchange_gloo(tmp_path, oracle):
    """The tensor-to-tensor halo exchange + block filter of the cfg5 workload (pipeline.exchange_halos_device /
    temporal_filter_block_device), 7 frames over 3 ranks on CPU tensors over gloo, against the single-process result."""
    world, F = 3, 7
    port = 29800 + (os.getpid() % 1000)
    mp.spawn(_temporal_device_worker, args=(world, port, str(tmp_path), F), nprocs=world, join=True)
    seq = _make_sequence(F, 2, 24, 28)
    masks = [c[2] for c in seq[0]]
    for f in range(F):
        lo, hi = max(0, f - 2), min(F - 1, f + 2)
        got = np.load(tmp_path / ("tfd%d.npy" % f))
        for cam in range(2):
            ref = oracle.temporal_filter([seq[t][cam][0] for t in range(lo, hi + 1)], [seq[t][cam][1] for t in range(lo, hi + 1)],
                                         [masks[cam]] * (hi - lo + 1), f - lo, 0.01, 1, 0.5, 1.0, 0.5)
            assert np.array_equal(got[cam].view(np.uint32), ref.view(np.uint32)), (f, cam)
    per_frame = 2 * 24 * 28 * (6 + 4)
    # blocks 3,3,1: rank 0 needs frames 3,4; rank 1 needs 1,2 and 6; rank 2 needs 4,5
    assert [int(np.load(tmp_path / ("bytes%d.npy" % r))[0]) for r in range(3)] == [2 * per_frame, 3 * per_frame, 2 * per_frame]


def test_more_ranks_than_frames_gloo(tmp_path, oracle):
    """3 frames on 4 ranks (blocks of 1,1,1,0): the idle rank must neither crash nor take part in the exchange."""
    from facebook360_dep_b200 import pipeline
    world, F = 4, 3
    port = 29700 + (os.getpid() % 1000)
    mp.spawn(_temporal_worker, args=(world, port, str(tmp_path), F), nprocs=world, join=True)
    seq = _make_sequence(F, 2, 24, 28)
    ref = pipeline.temporal_filter_block(oracle, seq, F, time_radius=2)
    for f in range(F):
        got = np.load(tmp_path / ("tf%d.npy" % f))
        assert np.array_equal(got.view(np.uint32), np.stack(ref[f]).view(np.uint32)), f


# ---- destination cameras of ONE frame dealt to the ranks: all-gather of disparities before mismatch handling ----
_MM = dict(S=4, W=40, H=40)


def _mm_inputs():
    from facebook360_dep_b200 import synth
    S, W, H = _MM["S"], _MM["W"], _MM["H"]
    rig = synth.ring_rig(S, W, H, kind="FTHETA")
    colors, true_disp = synth.render_rig(rig, W, H, scene=synth.Scene(seed=5))
    init = []
    for s in range(S):
        d = true_disp[s].copy()
        d[8:24, 10 + 3 * s:26 + 3 * s] *= 1.6  # a wrong patch per camera: the other cameras disagree with it
        init.append(d)
    return rig, colors, init


def _mm_run(ctx, own, colors, init, variant, mismatch_stage):
    """One fine level (level 0 of 2) on the cameras `own`; mismatch_stage(ctx) runs between the two halves."""
    W, H = _MM["W"], _MM["H"]
    kw = dict(random_proposals=variant, ping_pong_iterations=variant, mismatches_start_level=0)
    ctx.level_begin(W, H, level=0, num_levels=2, full_width=W, full_height=H, var_noise_floor=0.0, var_high_thresh=1e9)
    ctx.set_colors(colors)
    for i, cam in enumerate(own):
        ctx.set_disparity(i, init[cam])
    ctx.level_estimate(**kw)
    mismatch_stage(ctx)
    ctx.level_filter(**kw)
    return [(ctx.get_disparity(i, want_cost=False), ctx.get_mismatch_mask(i)) for i in range(len(own))]


def _mm_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from facebook360_dep_b200 import capi, pipeline
    oracle = oracle_libs.load_oracle()
    oracle.set_threads(2)
    rig, colors, init = _mm_inputs()
    own = shard.camera_shard(_MM["S"], world, rank)
    ctx = capi.Context(oracle, capi.rig_descs(rig), own)
    for variant in (0, 1):
        res = _mm_run(ctx, own, colors, init, variant, lambda c: pipeline.sharded_mismatches(c, _MM["S"]))
        for cam, (d, m) in zip(own, res):
            np.save(os.path.join(out_dir, "mm%d_cam%d_disp.npy" % (variant, cam)), d)
            np.save(os.path.join(out_dir, "mm%d_cam%d_mask.npy" % (variant, cam)), m)
    dist.barrier()
    dist.destroy_process_group()


def test_camera_shards():
    for S in (1, 3, 16, 24):
        for G in (1, 2, 8):
            seen = sorted(c for r in range(G) for c in shard.camera_shard(S, G, r))
            assert seen == list(range(S))


def test_camera_sharded_mismatches_gloo(tmp_path, oracle):
    """4 cameras over 2 ranks: estimate per shard, all-gather the disparity planes (gloo), mismatch handling per
    shard — must equal the single-context run of derp_process_level bit for bit, and the stage must have fired."""
    from facebook360_dep_b200 import capi
    world = 2
    port = 29700 + (os.getpid() % 1000)
    mp.spawn(_mm_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rig, colors, init = _mm_inputs()
    S = _MM["S"]
    ctx = capi.Context(oracle, capi.rig_descs(rig))
    for variant in (0, 1):
        ref = _mm_run(ctx, list(range(S)), colors, init, variant, lambda c: c.mismatches())
        fired = 0
        for cam in range(S):
            d = np.load(tmp_path / ("mm%d_cam%d_disp.npy" % (variant, cam)))
            m = np.load(tmp_path / ("mm%d_cam%d_mask.npy" % (variant, cam)))
            assert np.array_equal(m, ref[cam][1]), (variant, cam)
            same = (d.view(np.uint32) == ref[cam][0].view(np.uint32)) | (np.isnan(d) & np.isnan(ref[cam][0]))
            assert same.all(), (variant, cam)
            fired += int(m.sum())
        if variant == 0:
            assert fired > 50, "mismatch stage did not fire on the planted patches"


def test_sharded_mismatch_state_errors(oracle):
    from facebook360_dep_b200 import capi
    rig, colors, init = _mm_inputs()
    ctx = capi.Context(oracle, capi.rig_descs(rig), [0, 2])
    ctx.level_begin(_MM["W"], _MM["H"], level=0, num_levels=2)
    ctx.set_colors(colors)
    import pytest
    with pytest.raises(capi.DerpError):  # no gather yet
        ctx.mismatches_gathered()
    with pytest.raises(capi.DerpError):  # cameras 1 and 3 are not ours and no plane was given
        ctx.gather_disparities([None] * 4)
    with pytest.raises(capi.DerpError):  # the unsharded stage needs every camera as a destination (Derp.cpp:689)
        ctx.mismatches()
