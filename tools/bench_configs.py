#!/usr/bin/env python3
"""Timings of BASELINE.json configs 3-5 (kernel-only, CUDA events): sampling, long horizon, time sweep."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mav_trajectory_generation_cmake_b200 as ms  # noqa: E402

LO, HI = [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0]


def timeit(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


# config 4: 4,096 x K=256 (general kernel)
B, K = 4096, 256
pos = torch.from_numpy(ms.random_positions_host(B, K, LO, HI, 12345)).cuda()
times = ms.estimate_segment_times(pos, 3.0, 5.0)
coeffs4 = torch.empty((B, K, 3, 10), dtype=torch.float64, device="cuda")
ms_ = timeit(lambda: ms.solve_standard(pos, times, coeffs=coeffs4, want_status=False), steps=20, warm=3)
print("config4 long horizon: %d x K=%d  %.3f ms/batch  %.1f k solves/s  (HBM floor %.3f ms, FP64 floor %.3f ms)"
      % (B, K, ms_, B / ms_, B * 69656 / 6547.8e9 * 1e3, B * 239936 / 36.8e12 * 1e3))
for Bs in (512, 64):
    ms_ = timeit(lambda: ms.solve_standard(pos[:Bs], times[:Bs], coeffs=coeffs4[:Bs], want_status=False), steps=50, warm=5)
    print("config4 shard of %d trajectories (8-GPU share = 512): %.3f ms" % (Bs, ms_))
del coeffs4

# config 5: 8,192 x 64 allocations
B, S, K = 8192, 64, 10
pos = torch.from_numpy(ms.random_positions_host(B, K, LO, HI, 12345)).cuda()
times = ms.estimate_segment_times(pos, 3.0, 5.0)
g = torch.Generator(device="cuda").manual_seed(1)
sweep = (times[:, None, :] * (0.9 + 0.2 * torch.rand((B, S, K), generator=g, device="cuda", dtype=torch.float64))).contiguous()
ms_ = timeit(lambda: ms.cost_sweep(pos, sweep), steps=20)
ev = B * S / (ms_ * 1e-3)
print("config5 time sweep: %d x %d  %.3f ms  %.2f G evals/s  FP64 %.1f TFLOP/s (11,942 flop/eval) = %.2f of 36.8"
      % (B, S, ms_, ev / 1e9, ev * 11942 / 1e12, ev * 11942 / 36.8e12))

# config 3: sampling, 65,536 x 1000 (bounded slice of 1M x 1000)
B, K, M = 65536, 10, 1000
pos = torch.from_numpy(ms.random_positions_host(B, K, LO, HI, 12345)).cuda()
times = ms.estimate_segment_times(pos, 3.0, 5.0)
coeffs = ms.solve_standard(pos, times)["coeffs"]
out = torch.empty((B, M, 5, 3), dtype=torch.float64, device="cuda")
ms_ = timeit(lambda: ms.sample_uniform(coeffs, times, M, 5, out=out), steps=10)
print("config3 sampling: %d x %d  %.3f ms  %.2f G samples/s  %.3f of HBM write roofline"
      % (B, M, ms_, B * M / ms_ / 1e6, B * M * 120 / (ms_ * 1e-3) / 6547.8e9))
