"""Trajectory sharding across the GPUs of one box.

The batch shards by trajectory: rank r owns a contiguous range, solves it with no inter-GPU
traffic, and results are optionally collected on one rank with a single gather.  Works with
any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""


def shard_range(total, world_size, rank):
    """Contiguous range [start, start + count) of `total` trajectories owned by `rank`:
    ceil(total / world_size) per rank, the tail rank(s) take what is left."""
    per = -(-total // world_size)
    start = min(total, rank * per)
    return start, max(0, min(per, total - start))


def weak_scaling_seed_base(base_seed, per_gpu, rank):
    """Seed of the first trajectory of `rank` when every rank solves `per_gpu` problems: the
    global batch is seeds base_seed .. base_seed + world*per_gpu - 1, rank-major."""
    return base_seed + rank * per_gpu


def gather_to_rank0(local, total, dist, dst=0):
    """Collect the per-rank blocks (first dimension = trajectories of shard_range) on `dst`.
    Ranks may hold different counts; blocks are padded to the common size for the collective.
    Returns the [total, ...] tensor on dst, None elsewhere."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    per = -(-total // world)
    padded = local
    if local.shape[0] < per:
        pad = torch.zeros((per - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    padded = padded.contiguous()
    if rank == dst:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.gather(padded, parts, dst=dst)
        out = torch.cat([parts[r][:shard_range(total, world, r)[1]] for r in range(world)], dim=0)
        return out
    dist.gather(padded, None, dst=dst)
    return None
