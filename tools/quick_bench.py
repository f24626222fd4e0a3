#!/usr/bin/env python3
"""Kernel-only timing of the standard-mask solve and the sampler (CUDA events, rotating
buffers).  Usage: python tools/quick_bench.py [--steps 100] [--B 65536] [--K 10] [--sample] [--cost]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mav_trajectory_generation_cmake_b200 as ms  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--B", type=int, default=65536)
ap.add_argument("--K", type=int, default=10)
ap.add_argument("--M", type=int, default=1000)
ap.add_argument("--sample", action="store_true")
ap.add_argument("--cost", action="store_true")
ap.add_argument("--sweep", action="store_true")
args = ap.parse_args()

B, K = args.B, args.K
pos_h = ms.random_positions_host(B, K, [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0], 12345)
sets = 2
pos = [torch.from_numpy(pos_h).cuda() for _ in range(sets)]
times = [ms.estimate_segment_times(p, 3.0, 5.0) for p in pos]
coeffs = [torch.empty((B, K, 3, 10), dtype=torch.float64, device="cuda") for _ in range(sets)]


def timeit(fn, steps):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3


us = timeit(lambda i: ms.solve_standard(pos[i % sets], times[i % sets], coeffs=coeffs[i % sets], want_status=False,
                                        want_cost=args.cost), args.steps)
print("solve_standard: %.1f us/step  %.1f M solves/s  hbm %.3f of 6547.8 GB/s" %
      (us, B / us, 2744.0 * B / (us * 1e-6) / 1e9 / 6547.8))
if args.sample:
    M = args.M
    Bs = min(B, 65536)
    out = torch.empty((Bs, M, 5, 3), dtype=torch.float64, device="cuda")
    us = timeit(lambda i: ms.sample_uniform(coeffs[0][:Bs], times[0][:Bs], M, 5, out=out), 10)
    print("sample_uniform: %.1f us  %.2f G samples/s  hbm-write %.3f" %
          (us, Bs * M / us / 1e3, Bs * M * 120.0 / (us * 1e-6) / 1e9 / 6547.8))
if args.sweep:
    S = 64
    Bs = 8192
    tsw = times[0][:Bs, None, :].repeat(1, S, 1).contiguous()
    us = timeit(lambda i: ms.cost_sweep(pos[0][:Bs], tsw), 10)
    print("cost_sweep: %.1f us  %.1f M evals/s" % (us, Bs * S / us))
