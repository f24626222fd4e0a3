#!/usr/bin/env python3
"""e2e timing of minsnap_solve_standard_host with pinned buffers (the bench.py e2e leg alone)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mav_trajectory_generation_cmake_b200 as ms  # noqa: E402

B, K = 65536, 10
pos_h = ms.random_positions_host(B, K, [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0], 12345)
pos_pin = torch.from_numpy(pos_h).pin_memory()
times_pin = ms.estimate_segment_times(torch.from_numpy(pos_h).cuda(), 3.0, 5.0).cpu().pin_memory()
coeffs_pin = torch.empty((B, K, 3, 10), dtype=torch.float64).pin_memory()
for _ in range(3):
    ms.solve_standard_host(pos_pin, times_pin, coeffs=coeffs_pin)
t0 = time.perf_counter()
n = 20
for _ in range(n):
    ms.solve_standard_host(pos_pin, times_pin, coeffs=coeffs_pin)
dt = (time.perf_counter() - t0) / n
print("e2e: %.3f ms/step  %.2f M solves/s  (%.1f GB/s D2H-equivalent)" % (dt * 1e3, B / dt / 1e6, B * 2400 / dt / 1e9))
