import os, sys, torch
sys.path.insert(0, os.getcwd())
import mav_trajectory_generation_cmake_b200 as ms
K, B = 256, 4096
pos = torch.from_numpy(ms.random_positions_host(B, K, [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0], 12345)).cuda()
times = ms.estimate_segment_times(pos, 3.0, 5.0)
coeffs = torch.empty((B, K, 3, 10), dtype=torch.float64, device="cuda")
for _ in range(3):
    ms.solve_standard(pos, times, coeffs=coeffs, want_status=False)
torch.cuda.synchronize()
