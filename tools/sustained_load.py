#!/usr/bin/env python3
"""Burst against sustained load for the headline kernel: the 20-step CUDA graph of bench.py replayed back to back
for a few seconds; every ~50 ms one replay is timed with CUDA events and the NVML SM clock, power draw and
throttle reasons are read.   python tools/sustained_load.py [seconds]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mav_trajectory_generation_cmake_b200 as ms  # noqa: E402
import pynvml as nv  # noqa: E402

nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(0)
B, K = 65536, 10
pos_h = ms.random_positions_host(B, K, [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0], 12345)
pos = [torch.from_numpy(pos_h).cuda() for _ in range(2)]
times = [ms.estimate_segment_times(p, 3.0, 5.0) for p in pos]
coeffs = [torch.empty((B, K, 3, 10), dtype=torch.float64, device="cuda") for _ in range(2)]
for i in range(5):
    ms.solve_standard(pos[i % 2], times[i % 2], coeffs=coeffs[i % 2], want_status=False)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(20):
        ms.solve_standard(pos[i % 2], times[i % 2], coeffs=coeffs[i % 2], want_status=False)
torch.cuda.synchronize()
time.sleep(1.0)   # idle first
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
t0 = time.perf_counter()
print("t_ms   us/step  sm_MHz  power_W  reasons")
while time.perf_counter() - t0 < seconds:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    for _ in range(56):
        g.replay()
    clk = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
    pw = nv.nvmlDeviceGetPowerUsage(h) / 1e3
    try:
        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
    except Exception:
        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
    torch.cuda.synchronize()
    print("%6.0f  %6.2f  %6d  %7.0f  0x%x" % ((time.perf_counter() - t0) * 1e3, e0.elapsed_time(e1) / 20 * 1e3, clk, pw, mask))
