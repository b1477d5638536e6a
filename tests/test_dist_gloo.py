"""world_size-2 gloo test of the multi-GPU host logic (sharding + single logit all-gather) on CPU."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from adafocus_b200.dist import gather_logits, shard_bounds, shard_clips


def test_shard_bounds_cover_and_order():
    for n in (0, 1, 7, 64, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_bounds(256, 3, 8) == (96, 128)
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, num_clips):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        clips = torch.arange(num_clips, dtype=torch.float32).view(num_clips, 1).repeat(1, 5)
        mine = shard_clips(clips)
        lo, hi = shard_bounds(num_clips, rank, world)
        assert torch.equal(mine, clips[lo:hi])
        local_logits = mine * 2 + 1                     # stands in for the per-clip logits of this shard
        expect = clips * 2 + 1
        assert torch.equal(gather_logits(local_logits), expect)
        assert torch.equal(gather_logits(local_logits, num_clips=num_clips), expect)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_clips", [8, 7])
def test_gather_logits_world2(num_clips):
    port = 29500 + (os.getpid() % 2000) + num_clips
    mp.spawn(_worker, args=(2, port, num_clips), nprocs=2, join=True)
