#!/usr/bin/env python3
"""Latency of ONE trajectory through the host-buffer entry points (what the C++ drop-in classes call)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mav_trajectory_generation_cmake_b200 as ms  # noqa: E402

K = 10
pos = ms.random_positions_host(1, K, [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0], 12345)
times = ms.estimate_segment_times_host(pos, 3.0, 5.0)
mask = ms.standard_mask(K)
fixed = np.zeros((1, K + 9, 3))
fixed[0, 0] = pos[0, 0]
fixed[0, 5:5 + K - 1] = pos[0, 1:K]
fixed[0, 5 + K - 1] = pos[0, K]


def lat(fn, n=200):
    for _ in range(20):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e6


print("solve_host (general, B=1):          %.1f us/call" % lat(lambda: ms.solve_host(mask, fixed, times)))
print("solve_standard_host (B=1):          %.1f us/call" % lat(lambda: ms.solve_standard_host(pos, times)))
out = ms.solve_standard_host(pos, times)
t = np.linspace(0, times.sum() * 0.99, 100)
print("sample_at_host (B=1, M=100):        %.1f us/call" % lat(lambda: ms.sample_at_host(out["coeffs"], times, t, 5)))
print("reorder_host (1 mask):              %.1f us/call" % lat(lambda: ms.reorder_host(mask[None], 10, K)))
