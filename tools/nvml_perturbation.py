import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import mav_trajectory_generation_cmake_b200 as ms
import pynvml as nv
nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
B, K = 65536, 10
pos_h = ms.random_positions_host(B, K, [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0], 12345)
pos = [torch.from_numpy(pos_h).cuda() for _ in range(2)]
times = [ms.estimate_segment_times(p, 3.0, 5.0) for p in pos]
coeffs = [torch.empty((B, K, 3, 10), dtype=torch.float64, device="cuda") for _ in range(2)]
def launch(i):
    s = i % 2
    ms.solve_standard(pos[s], times[s], coeffs=coeffs[s], want_status=False)
for i in range(5): launch(i)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(20): launch(i)
for _ in range(60): g.replay()
torch.cuda.synchronize()
def clock(): nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
def reasons():
    try: nv.nvmlDeviceGetCurrentClocksEventReasons(h)
    except Exception: nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
modes = {"none": lambda: None, "clock": clock, "reasons": reasons, "both": lambda: (clock(), reasons())}
res = {k: [] for k in modes}
for rep in range(30):
    for name, fn in modes.items():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        torch.cuda._sleep(600_000)
        e0.record(); g.replay(); e1.record()
        t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
        torch.cuda.synchronize()
        res[name].append((e0.elapsed_time(e1) / 20 * 1e3, dt * 1e6))
print("none, per repetition:", " ".join("%.1f" % x[0] for x in res["none"]))
for k, v in res.items():
    a = np.array(v)
    print("%-8s us/step: median %.2f  mean %.2f  max %.2f   host call: median %.0f us max %.0f us" % (k, np.median(a[:, 0]), a[:, 0].mean(), a[:, 0].max(), np.median(a[:, 1]), a[:, 1].max()))
