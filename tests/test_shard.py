"""N>1 path on CPU: world_size-2 gloo processes run the frame sharding + timing reduction that bench.py uses
under torchrun, and (with the oracle standing in for the per-rank library) check that sharded frames give the
same per-frame results as a single rank."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from facebook360_dep_b200 import shard


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
    oracle = capi.load_oracle()
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


def _temporal_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from facebook360_dep_b200 import capi, pipeline
    oracle = capi.load_oracle()
    oracle.set_threads(2)
    F, S, H, W = 7, 2, 24, 28
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
