#!/usr/bin/env python3
"""Fused solve + gather: every rank's solve kernel stores its coefficient block straight into a
buffer on rank 0 (CUDA IPC peer mapping over NVLink / NVSwitch), compared with solve + NCCL gather.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/peer_gather_demo.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mav_trajectory_generation_cmake_b200 as ms  # noqa: E402
from mav_trajectory_generation_cmake_b200.sharding import PeerGatherBuffer, gather_to_rank0, weak_scaling_seed_base  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B, K = 65536, 10
pos = torch.from_numpy(ms.random_positions_host(B, K, [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0],
                                                weak_scaling_seed_base(12345, B, rank))).cuda()
times = ms.estimate_segment_times(pos, 3.0, 5.0)
local_coeffs = torch.empty((B, K, 3, 10), dtype=torch.float64, device="cuda")
buf = PeerGatherBuffer(dist, B, (K, 3, 10), torch.float64, dst=0)
mine = buf.view()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


ms_local = timed(lambda: ms.solve_standard(pos, times, coeffs=local_coeffs, want_status=False))
ms_fused = timed(lambda: ms.solve_standard(pos, times, coeffs=mine, want_status=False))


def separate():
    ms.solve_standard(pos, times, coeffs=local_coeffs, want_status=False)
    return gather_to_rank0(local_coeffs, world * B, dist)


ms_separate = timed(separate, reps=5)


def push():
    # local solve, then one device-to-peer copy of the block (large NVLink writes instead of per-lane stores)
    ms.solve_standard(pos, times, coeffs=local_coeffs, want_status=False)
    buf.push(local_coeffs)


ms_push = timed(push, reps=5)
# correctness: the peer buffer on rank 0 holds every rank's block, bit for bit
ms.solve_standard(pos, times, coeffs=mine, want_status=False)
buf.finish()
ms.solve_standard(pos, times, coeffs=local_coeffs, want_status=False)
ref = gather_to_rank0(local_coeffs, world * B, dist)
ok = True
if rank == 0:
    ok = bool(torch.equal(buf.whole(), ref))
    print("ranks %d: solve (local stores) %.3f ms | solve with peer stores into rank 0 %.3f ms | solve + peer copy %.3f ms"
          " | solve + NCCL gather %.3f ms | peer buffer == gathered blocks: %s"
          % (world, ms_local, ms_fused, ms_push, ms_separate, ok), flush=True)
buf.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
