"""Multi-rank host logic on CPU: world_size-2 gloo processes shard a batch by trajectory and
gather the per-rank blocks on rank 0 (the only collective the path has)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mav_trajectory_generation_cmake_b200.sharding import gather_to_rank0, shard_range, weak_scaling_seed_base


def test_shard_ranges_cover_the_batch_once():
    for total in (0, 1, 7, 16, 65536, 65537, 4096):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                s, c = shard_range(total, world, r)
                seen.extend(range(s, s + c))
            assert seen == list(range(total))
    # weak scaling: rank-major seeds, no overlap
    seeds = [weak_scaling_seed_base(12345, 65536, r) for r in range(8)]
    assert seeds == [12345 + 65536 * r for r in range(8)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, count = shard_range(total, world, rank)
        # stand-in for this rank's solve: every "coefficient" encodes its global trajectory index
        local = (torch.arange(start, start + count, dtype=torch.float64).reshape(count, 1, 1, 1)
                 * torch.ones((1, 2, 3, 10), dtype=torch.float64))
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == float(world)
        out = gather_to_rank0(local, total, dist)
        if rank == 0:
            np.save(result_path, out.numpy())
        else:
            assert out is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_gloo_world2_shard_and_gather(tmp_path, total):
    world = 2
    result = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, _free_port(), total, result), nprocs=world, join=True)
    out = np.load(result)
    assert out.shape == (total, 2, 3, 10)
    assert np.array_equal(out[:, 0, 0, 0], np.arange(total, dtype=np.float64))
