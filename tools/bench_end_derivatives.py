import os, sys, torch
sys.path.insert(0, os.getcwd())
import mav_trajectory_generation_cmake_b200 as ms
B, K = 131072, 8
pos = torch.from_numpy(ms.random_positions_host(B, K, [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0], 12345)).cuda()
times = ms.estimate_segment_times(pos, 3.0, 5.0)
coeffs = torch.empty((B, K, 3, 10), dtype=torch.float64, device="cuda")
end = torch.randn((B, 2, 4, 3), dtype=torch.float64, device="cuda") * 0.1
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("plain      %.1f us" % t(lambda: ms.solve_standard(pos, times, coeffs=coeffs, want_status=False)))
print("with ends  %.1f us" % t(lambda: ms.solve_standard(pos, times, end_derivatives=end, coeffs=coeffs, want_status=False)))
