#!/usr/bin/env python3
"""How much untimed work the timed window of bench.py needs behind it after the GPU has idled: for each warm-up
length, idle 1.5 s, run 5 eager steps + the warm-up replays + the spin kernel + ONE timed replay of the 20-step graph.
    python tools/warmup_curve.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mav_trajectory_generation_cmake_b200 as ms  # noqa: E402

B, K = 65536, 10
pos_h = ms.random_positions_host(B, K, [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0], 12345)
pos = [torch.from_numpy(pos_h).cuda() for _ in range(2)]
times = [ms.estimate_segment_times(p, 3.0, 5.0) for p in pos]
coeffs = [torch.empty((B, K, 3, 10), dtype=torch.float64, device="cuda") for _ in range(2)]


def launch(i):
    ms.solve_standard(pos[i % 2], times[i % 2], coeffs=coeffs[i % 2], want_status=False)


for i in range(5):
    launch(i)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(20):
        launch(i)
torch.cuda.synchronize()
print("warm-up replays -> us per step of the timed replay (5 trials, 1.5 s idle before each)")
for n_rep in (1, 6, 12, 24, 60, 120, 240):
    vals = []
    for trial in range(5):
        time.sleep(1.5)
        for i in range(5):
            launch(i)
        for _ in range(n_rep):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(600_000)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        vals.append(e0.elapsed_time(e1) / 20 * 1e3)
    print("%4d replays (%5.1f ms): %s   median %.2f" % (n_rep, n_rep * 0.84, " ".join("%.2f" % v for v in vals), np.median(vals)))
