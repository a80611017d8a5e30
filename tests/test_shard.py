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
