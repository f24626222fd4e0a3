#!/usr/bin/env python3
"""Kernel-only timing of minsnap_extrema on solved trajectories (CUDA events).
Usage: python tools/bench_extrema.py [--B 65536] [--K 10] [--steps 20]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mav_trajectory_generation_cmake_b200 as ms  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=65536)
ap.add_argument("--K", type=int, default=10)
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
B, K = args.B, args.K
pos = torch.from_numpy(ms.random_positions_host(B, K, [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0], 12345)).cuda()
times = ms.estimate_segment_times(pos, 3.0, 5.0)
coeffs = ms.solve_standard(pos, times, want_status=False)["coeffs"]
for k, name in ((1, "velocity"), (2, "acceleration")):
    for mode, mname in ((0, "computeMaximumOfMagnitude"), (1, "computeMinMaxMagnitude"), (17, "computeMinMaxMagnitude+keep")):
        for _ in range(3):
            ms.extrema(coeffs, times, k, mode=mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            r = ms.extrema(coeffs, times, k, mode=mode)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / args.steps * 1e3
        print("%-12s %-30s %9.1f us  %7.2f M trajectories/s  %7.1f M segments/s  mean max %.4f" %
              (name, mname, us, B / us, B * K / us, float(r["max_value"].mean())))
