"""Multi-GPU sharding of the inference path: clips are independent units, weights are replicated, and the only
exchange is ONE all-gather of the per-clip logits.  The reference has no counterpart: its
validate() evaluates the whole set redundantly on every rank (ACT/main_dist.py:239).

One process per GPU (torchrun); backend 'nccl' on GPUs (NVLink/NVSwitch), 'gloo' in the CPU tests."""
import torch
import torch.distributed as dist


def shard_bounds(num_clips, rank, world):
    """Contiguous split of [0, num_clips) -> [lo, hi) for `rank`; the first (num_clips % world) ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, extra = divmod(num_clips, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_clips(clips, rank=None, world=None):
    """Slice this rank's clips out of a (B, ...) tensor (host or device)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(clips.shape[0], rank, world)
    return clips[lo:hi]


def gather_logits(local_logits, num_clips=None, group=None):
    """All-gather (B_local, C) per-clip logits into (B_global, C) in clip order on every rank.  Equal shards use a
    single all_gather_into_tensor (one NCCL all-gather, 25.6 KB/rank at 32 clips x 200 classes); ragged shards pad
    to the largest shard first."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_logits
    world = dist.get_world_size(group)
    c = local_logits.shape[1]
    if num_clips is None:
        n = torch.tensor([local_logits.shape[0]], device=local_logits.device)
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n, group=group)
        counts = [int(s.item()) for s in sizes]
    else:
        counts = [shard_bounds(num_clips, r, world)[1] - shard_bounds(num_clips, r, world)[0] for r in range(world)]
    m = max(counts)
    if all(k == m for k in counts):
        out = torch.empty(world * m, c, dtype=local_logits.dtype, device=local_logits.device)
        dist.all_gather_into_tensor(out, local_logits.contiguous(), group=group)
        return out
    padded = torch.zeros(m, c, dtype=local_logits.dtype, device=local_logits.device)
    padded[: local_logits.shape[0]] = local_logits
    out = torch.empty(world * m, c, dtype=local_logits.dtype, device=local_logits.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * m: r * m + counts[r]] for r in range(world)], 0)
