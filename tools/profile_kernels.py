#!/usr/bin/env python3
"""One launch (after a warm-up launch) of every secondary kernel at a profile-friendly size, for ncu:
sampler, cost sweep (direct-read instantiation of the pair kernel), general solve, cyclic reduction, extrema,
collision cost.  Prints CUDA-event times of the same launches.

    python tools/profile_kernels.py
    ncu --set full --clock-control none -k regex:'sample_all|pair_kernel|solve_general|bcr_kernel|segment_extrema|collision_cost' \
        -o gpurun_out/prof_secondary python tools/profile_kernels.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mav_trajectory_generation_cmake_b200 as ms  # noqa: E402

LO, HI = [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0]


def timed(name, fn, reps=2):
    if os.environ.get("PROFILE_ONCE"):      # under ncu: exactly one launch per kernel
        fn()
        torch.cuda.synchronize()
        return
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-28s %.3f ms" % (name, e0.elapsed_time(e1) / reps))


B, K = 65536, 10
pos = torch.from_numpy(ms.random_positions_host(B, K, LO, HI, 12345)).cuda()
times = ms.estimate_segment_times(pos, 3.0, 5.0)
coeffs = ms.solve_standard(pos, times, want_status=False)["coeffs"]

M = 250
samples = torch.empty((B, M, 5, 3), dtype=torch.float64, device="cuda")
timed("sample 65536 x 250", lambda: ms.sample_uniform(coeffs, times, M, 5, out=samples))

S = 64
gen = torch.Generator(device="cuda")
gen.manual_seed(1)
sweep_t = (times[:8192, None, :] * (0.75 + 0.5 * torch.rand((8192, S, K), dtype=torch.float64, device="cuda", generator=gen))).contiguous()
timed("cost sweep 8192 x 64", lambda: ms.cost_sweep(pos[:8192], sweep_t))

mask = ms.standard_mask(K)
fixed = torch.zeros((B, K + 9, 3), dtype=torch.float64, device="cuda")
fixed[:, 0] = pos[:, 0]
fixed[:, 5:5 + K - 1] = pos[:, 1:K]
fixed[:, 5 + K - 1] = pos[:, K]
timed("general solve 65536 x 10", lambda: ms.solve(mask, fixed, times, want_cost=False))

lh_pos = torch.from_numpy(ms.random_positions_host(4096, 256, LO, HI, 12345)).cuda()
lh_times = ms.estimate_segment_times(lh_pos, 3.0, 5.0)
lh_coeffs = torch.empty((4096, 256, 3, 10), dtype=torch.float64, device="cuda")
timed("cyclic reduction 4096 x 256", lambda: ms.solve_standard(lh_pos, lh_times, coeffs=lh_coeffs, want_status=False))

timed("extrema |v| 65536 x 10", lambda: ms.extrema(coeffs, times, 1))

ax = [(-12.0, 48), (-22.0, 88), (-12.0, 48)]
X, Y, Z = np.meshgrid(*[o + (np.arange(n) + 0.5) * 0.5 for o, n in ax], indexing="ij")
sdf = torch.from_numpy(np.sqrt((X - 1.0) ** 2 + (Y + 2.0) ** 2 + (Z - 0.5) ** 2) - 3.0).cuda()
timed("collision cost 65536 x 10", lambda: ms.collision_cost(coeffs, times, sdf, [-12.0, -22.0, -12.0], 0.5, LO, HI))
