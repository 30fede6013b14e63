"""Multi-GPU sharding of a frame batch: one process per GPU, frames split contiguously, no
collective on the hot path (a frame's channel estimate is produced and consumed inside that
frame: cpuLS_main.cpp:80-93).  The only exchange is the optional gather of the decoded bits
(and combined symbols) to one rank, over NCCL on GPUs or gloo on CPUs -- a few hundred kB
per GB of input."""
from __future__ import annotations


def shard_frames(n_frames: int, world: int, rank: int):
    """Contiguous block partition: returns (first_frame, count); counts differ by at most one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    base, rem = divmod(n_frames, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def all_counts(n_frames: int, world: int):
    return [shard_frames(n_frames, world, r)[1] for r in range(world)]


def gather_rows(local, n_frames: int, dst: int = 0, group=None):
    """Gather per-frame result rows (tensor [count_r, ...]) of every rank on `dst` in global frame
    order.  Uneven shards are padded to the largest count for the collective and trimmed after.
    Returns the [n_frames, ...] tensor on dst, None elsewhere."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = all_counts(n_frames, world)
    if local.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} frames, expected {counts[rank]}")
    mx = max(counts)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))])
    pad = pad.contiguous()
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)])
