#!/usr/bin/env python3
"""Device-resident time-only optimisation (minsnap_optimize_segment_times): time per descent iteration
and the objective it reaches, against the torch-glue version of round 1.
    python tools/bench_time_optimization.py [--B 8192] [--iterations 20] [--steps 16]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mav_trajectory_generation_cmake_b200 as ms  # noqa: E402
from mav_trajectory_generation_cmake_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=8192)
ap.add_argument("--K", type=int, default=10)
ap.add_argument("--iterations", type=int, default=20)
ap.add_argument("--steps", type=int, default=16)
a = ap.parse_args()
pos = torch.from_numpy(ms.random_positions_host(a.B, a.K, [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0], 12345)).cuda()
t0 = ms.estimate_segment_times(pos, 3.0, 5.0)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


ms_dev, (t_dev, h_dev) = timed(lambda: api.optimize_segment_times(pos, t0, iterations=a.iterations, n_steps=a.steps))
ms_glue, (t_glue, h_glue) = timed(lambda: api.optimize_segment_times_reference_glue(pos, t0, iterations=a.iterations, n_steps=a.steps))
print("B=%d K=%d, %d iterations x %d step lengths" % (a.B, a.K, a.iterations, a.steps))
print("device-resident driver: %.3f ms in total, %.1f us per iteration (%.1f M trajectory-iterations/s)"
      % (ms_dev, ms_dev / a.iterations * 1e3, a.B * a.iterations / ms_dev / 1e3))
print("torch-glue driver (round 1): %.3f ms in total, %.1f us per iteration" % (ms_glue, ms_glue / a.iterations * 1e3))
print("objective (batch mean): start %.6g -> %.6g; glue %.6g; max |difference| of the final objectives %.2e (relative)"
      % (float(h_dev[0].mean()), float(h_dev[-1].mean()), float(h_glue[-1].mean()),
         float(((h_dev[-1] - h_glue[-1]).abs() / h_glue[-1].abs()).max())))
