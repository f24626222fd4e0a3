#!/usr/bin/env python3
"""bench.py -- headline benchmark: min-snap solves/s (3-D, N=10, 10 segments) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path (setup + solveLinear)
over one batch of 65,536 synthetic trajectories PER GPU (BASELINE.json configs[1]; weak
scaling, the batch shards by trajectory with no data-path collective).  Inputs (positions,
segment times) are resident in HBM when the timed region starts; coefficients land in HBM.

Printed JSON line (rank 0): value = whole-job solves/s, roofline (HBM, algorithmic 2,744 B per
solve), cpu_baseline (the oracle port on this box's host cores, bounded sample), e2e (same
metric through the host-buffer C-ABI call with pinned host memory, copies inside), clocks.

--impl reference: the reference's CPU path.  The reference itself cannot be compiled in this
image (Eigen3/glog/NLopt absent), so this arm times the oracle port of its algorithm (dense
Householder QR in place of Eigen::SparseQR) with all host threads.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "min-snap solves/sec (3D, N=10, 10 seg)"
UNIT = "solves/s"
B_PER_GPU = 65536
K_SEG, DIM, NCOEF, SNAP = 10, 3, 10, 4
V_MAX, A_MAX, MAGIC = 3.0, 5.0, 6.5
BOX_LO, BOX_HI = [-10.0, -20.0, -10.0], [10.0, 20.0, 10.0]
BASE_SEED = 12345
BYTES_IN = 8 * ((K_SEG + 1) * DIM + K_SEG)        # 344 B: positions + segment times
BYTES_OUT = 8 * K_SEG * DIM * NCOEF               # 2400 B: coefficients
BYTES_PER_SOLVE = BYTES_IN + BYTES_OUT            # 2744 B algorithmic HBM traffic
FLOPS_PER_SOLVE = K_SEG * (95 + 132 * DIM) + (K_SEG - 1) * (160 + 96 * DIM)   # 8942
CONFIG = {"workload": "configs[1]: batch of 65,536 independent 3-D N=10 min-snap problems, 10 segments each, per GPU",
          "batch_per_gpu": B_PER_GPU, "segments": K_SEG, "dimension": DIM, "N": NCOEF, "derivative": "snap",
          "inputs": "createRandomVertices(seed=12345+b, box +-(10,20,10)) + estimateSegmentTimes(3,5,6.5)"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.active = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            while not self._stop_evt.is_set():
                if self.active.is_set():
                    self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if mask & bit:
                            self.reasons.add(name)
                time.sleep(0.002)
        except Exception as exc:  # NVML missing: report that instead of inventing numbers
            self.reasons.add("nvml_unavailable: %s" % type(exc).__name__)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference algorithm (test infrastructure used as the baseline)
# ------------------------------------------------------------------------------------------
def cpu_solves_per_s(positions, times, n_threads, budget_s):
    """Times the oracle on a bounded sample of the workload; returns (solves/s, sample size)."""
    from oracle.oracle_py import Oracle
    orc = Oracle("f64")
    probe = 256
    t0 = time.perf_counter()
    orc.solve_batch_standard(positions[:probe], times[:probe], NCOEF, SNAP, 4, 1, want_coeffs=True)
    per_solve_1t = (time.perf_counter() - t0) / probe
    sample = int(min(len(positions), max(1024, budget_s / per_solve_1t)))
    coeffs_warm, _, _ = orc.solve_batch_standard(positions[:n_threads * 64], times[:n_threads * 64], NCOEF, SNAP, 4,
                                                 n_threads)
    best = 0.0
    for _ in range(3):
        t0 = time.perf_counter()
        _, _, st = orc.solve_batch_standard(positions[:sample], times[:sample], NCOEF, SNAP, 4, n_threads)
        dt = time.perf_counter() - t0
        best = max(best, sample / dt)
        assert st == 0
    return best, sample, 1.0 / per_solve_1t


def parity_block(ms, torch, pos_h, times_h, n_threads, n_check=4096):
    """parity checks run with every benchmark -- the CUDA path against the oracle on a
    random subset of the benchmark batch (coefficients, cost, sampled derivatives, index map)."""
    from oracle.oracle_py import Oracle, standard_mask
    orc = Oracle("f64")
    rng = np.random.default_rng(20261018)
    idx = np.sort(rng.choice(len(pos_h), size=min(n_check, len(pos_h)), replace=False))
    p, t = np.ascontiguousarray(pos_h[idx]), np.ascontiguousarray(times_h[idx])
    ref_c, ref_cost, st = orc.solve_batch_standard(p, t, NCOEF, SNAP, 4, n_threads)
    out = ms.solve_standard(torch.from_numpy(p).cuda(), torch.from_numpy(t).cuda(), want_cost=True)
    c = out["coeffs"].cpu().numpy()
    num = np.abs(c - ref_c).max(axis=-1)
    den = np.abs(ref_c).max(axis=-1)
    coeff_err = float((num / np.where(den == 0, 1.0, den)).max())
    cost_err = float(np.abs(out["cost"].cpu().numpy() / ref_cost - 1.0).max())
    n_s = 64
    samples, ts = ms.sample_uniform(out["coeffs"][:n_s], torch.from_numpy(t[:n_s]).cuda(), 128, 5, want_times=True)
    samples, ts = samples.cpu().numpy(), ts.cpu().numpy()
    sample_err = max(float(np.abs(samples[b] - orc.trajectory_sample(ref_c[b], t[b], ts[b], 5)).max())
                     for b in range(n_s))
    mask = standard_mask(K_SEG, NCOEF)
    col_ref, nf, npf = orc.reorder(NCOEF, K_SEG, mask)
    col, counts = ms.reorder(torch.from_numpy(mask.reshape(1, -1)).cuda(), NCOEF, K_SEG)
    index_ok = bool(np.array_equal(col.cpu().numpy()[0], col_ref)) and tuple(counts.cpu().numpy()[0]) == (nf, npf)
    ok = st == 0 and int((out["status"] != 0).sum()) == 0 and coeff_err <= 1e-8 and cost_err <= 1e-8 and \
        sample_err <= 1e-6 and index_ok
    return {"ok": bool(ok), "n_checked": int(len(idx)), "coeff_rel_err": coeff_err, "cost_rel_err": cost_err,
            "sample_abs_err": sample_err, "index_map_bit_exact": index_ok,
            "bars": {"coeff_rel_err": 1e-8, "cost_rel_err": 1e-8, "sample_abs_err": 1e-6}}


def run_reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    import mav_trajectory_generation_cmake_b200 as ms   # host-only input generator (no GPU use)
    cores = os.cpu_count() or 1
    n = B_PER_GPU
    pos = ms.random_positions_host(n, K_SEG, BOX_LO, BOX_HI, BASE_SEED)
    from oracle.oracle_py import Oracle
    orc = Oracle("f64")
    times = np.stack([orc.estimate_segment_times(pos[b], V_MAX, A_MAX, MAGIC) for b in range(n)])
    # size one step so that warmup + steps finish within a few minutes
    probe = 256
    t0 = time.perf_counter()
    orc.solve_batch_standard(pos[:probe], times[:probe], NCOEF, SNAP, 4, 1)
    per_solve = (time.perf_counter() - t0) / probe
    budget_per_step = min(10.0, 120.0 / max(1, args.steps + args.warmup))
    sample = int(min(n, max(512, budget_per_step * cores / per_solve)))
    for _ in range(args.warmup):
        orc.solve_batch_standard(pos[:sample], times[:sample], NCOEF, SNAP, 4, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.solve_batch_standard(pos[:sample], times[:sample], NCOEF, SNAP, 4, cores)
    elapsed = time.perf_counter() - t0
    value = sample * args.steps / elapsed
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": CONFIG,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d of the 65,536 problems per step; oracle port of the reference algorithm "
                                   "(dense Householder QR for Eigen::SparseQR), OpenMP over problems" % sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference needs Eigen3/glog/NLopt, absent from this image: its algorithm is timed through the oracle port",
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def bind_to_gpu_cpus(local):
    """Run this rank on the host cores (and so allocate its pinned buffers on the NUMA node) next to
    its GPU: with 8 ranks on a two-socket host the e2e copies otherwise cross the socket link."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + bit for w, mask in enumerate(words) for bit in range(64) if (mask >> bit) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import mav_trajectory_generation_cmake_b200 as ms

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the GPU arm has no CPU fallback")
    torch.cuda.set_device(local)
    n_local_cpus = bind_to_gpu_cpus(local)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if distributed:
            dist.barrier()

    ms.load()
    B = B_PER_GPU
    # each rank owns a contiguous slice of the global batch: seeds 12345 + rank*B + b
    from mav_trajectory_generation_cmake_b200.sharding import gather_to_rank0, weak_scaling_seed_base
    pos_h = ms.random_positions_host(B, K_SEG, BOX_LO, BOX_HI, weak_scaling_seed_base(BASE_SEED, B, rank))
    n_sets = 2   # rotate buffer sets so a step never finds its inputs in L2 (2 x 180 MB > 126 MB)
    pos_d = [torch.from_numpy(pos_h).cuda() for _ in range(n_sets)]
    times_d = [ms.estimate_segment_times(p, V_MAX, A_MAX, MAGIC) for p in pos_d]
    coeffs_d = [torch.empty((B, K_SEG, DIM, NCOEF), dtype=torch.float64, device="cuda") for _ in range(n_sets)]
    torch.cuda.synchronize()

    def step(i):
        s = i % n_sets
        ms.solve_standard(pos_d[s], times_d[s], coeffs=coeffs_d[s], want_status=False)

    sampler = ClockSampler(local)
    sampler.start()

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active.set()
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    torch.cuda.synchronize()
    sampler.active.clear()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    if distributed:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = t.item()
    ms_per_step = elapsed_ms / args.steps
    value = world * B * args.steps / (elapsed_ms * 1e-3)

    # ---- e2e: host buffers through the C-ABI host entry point, copies inside the timed region
    pos_pin = torch.from_numpy(pos_h).pin_memory()
    times_pin = times_d[0].cpu().pin_memory()
    coeffs_pin = torch.empty((B, K_SEG, DIM, NCOEF), dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(3):
        ms.solve_standard_host(pos_pin, times_pin, coeffs=coeffs_pin)
    torch.cuda.synchronize()
    barrier()
    sampler.active.set()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ms.solve_standard_host(pos_pin, times_pin, coeffs=coeffs_pin)
    e2e_s = time.perf_counter() - t0
    sampler.active.clear()
    barrier()
    if distributed:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_value = world * B * e2e_steps / e2e_s
    sampler.stop()

    # ---- optional: one NCCL gather of the coefficient blocks (reported separately) ------------
    gather_ms = None
    if distributed:
        gather_to_rank0(coeffs_d[0], world * B, dist)
        torch.cuda.synchronize()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        gathered = gather_to_rank0(coeffs_d[0], world * B, dist)
        g1.record()
        del gathered
        torch.cuda.synchronize()
        t = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gather_ms = t.item()

    # the same collection through a CUDA-IPC peer buffer on rank 0: one device-to-peer copy per rank
    peer_gather_ms = None
    if distributed:
        from mav_trajectory_generation_cmake_b200.sharding import PeerGatherBuffer
        buf, ok = None, 1
        try:
            buf = PeerGatherBuffer(dist, B, (K_SEG, DIM, NCOEF), torch.float64, dst=0)
        except Exception:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        all_ok = int(flag.item()) == 1
        if all_ok:
            buf.push(coeffs_d[0])
            torch.cuda.synchronize()
            barrier()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record()
            buf.push(coeffs_d[0])
            q1.record()
            torch.cuda.synchronize()
            barrier()
            t = torch.tensor([q0.elapsed_time(q1)], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            peer_gather_ms = t.item()
        if buf is not None:
            buf.close(barrier=all_ok)

    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        achieved = BYTES_PER_SOLVE * B / (ms_per_step * 1e-3) / 1e9          # GB/s, this rank's kernel
        fp64_peak = ms.fp64_peak(5)
        fp64_achieved = FLOPS_PER_SOLVE * B / (ms_per_step * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "solve_standard_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        # sampling throughput (configs[2] shape, bounded to 65,536 trajectories x 1000 instants)
        M = 1000
        samples = torch.empty((B, M, 5, DIM), dtype=torch.float64, device="cuda")
        for _ in range(3):
            ms.sample_uniform(coeffs_d[0], times_d[0], M, 5, out=samples)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        reps = 5
        for _ in range(reps):
            ms.sample_uniform(coeffs_d[0], times_d[0], M, 5, out=samples)
        s1.record()
        torch.cuda.synchronize()
        sample_ms = s0.elapsed_time(s1) / reps
        samples_per_s = B * M / (sample_ms * 1e-3)
        del samples
        # extrema of |velocity| over every trajectory (computeMaximumOfMagnitude)
        for _ in range(2):
            ms.extrema(coeffs_d[0], times_d[0], 1)
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record()
        for _ in range(reps):
            ms.extrema(coeffs_d[0], times_d[0], 1)
        x1.record()
        torch.cuda.synchronize()
        extrema_ms = x0.elapsed_time(x1) / reps

        # BASELINE configs[3]: 4,096 trajectories of 256 segments (dependency-chain bound)
        lh_pos = torch.from_numpy(ms.random_positions_host(4096, 256, BOX_LO, BOX_HI, BASE_SEED)).cuda()
        lh_times = ms.estimate_segment_times(lh_pos, V_MAX, A_MAX, MAGIC)
        lh_coeffs = torch.empty((4096, 256, DIM, NCOEF), dtype=torch.float64, device="cuda")
        for _ in range(3):
            ms.solve_standard(lh_pos, lh_times, coeffs=lh_coeffs, want_status=False)
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for _ in range(20):
            ms.solve_standard(lh_pos, lh_times, coeffs=lh_coeffs, want_status=False)
        l1.record()
        torch.cuda.synchronize()
        long_horizon_ms = l0.elapsed_time(l1) / 20
        del lh_pos, lh_times, lh_coeffs

        cores = os.cpu_count() or 1
        cpu_value, cpu_sample, cpu_1t = cpu_solves_per_s(pos_h, times_pin.numpy(), cores, budget_s=15.0)
        parity = parity_block(ms, torch, pos_h, times_pin.numpy(), cores)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(CONFIG, l2="inputs/outputs rotate over %d buffer sets of 180 MB (> 126 MB L2)" % n_sets,
                           parallelism="trajectory-sharded x%d, no data-path collective" % world),
            "gpu_launches": args.steps * ms_launches_per_step(ms),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_solve": BYTES_PER_SOLVE, "kernel": "solve_standard",
                         "fp64_tflops_achieved": fp64_achieved, "fp64_tflops_peak_measured": fp64_peak,
                         "fp64_frac": fp64_achieved / fp64_peak, "flops_per_solve": FLOPS_PER_SOLVE},
            "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d of the 65,536 problems, best of 3; oracle port of the reference "
                                       "algorithm (dense QR stands in for Eigen::SparseQR); 1 thread: %.0f solves/s"
                                       % (cpu_sample, cpu_1t)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": BYTES_IN * B,
                    "d2h_bytes_per_step": BYTES_OUT * B, "steps": e2e_steps,
                    "api": "minsnap_solve_standard_host (pinned host buffers, chunked double-buffered copies)",
                    "host_cpus_bound": n_local_cpus},
            "clocks": sampler.summary(),
            "parity": parity,
            "extra": {"samples_per_s": samples_per_s, "sample_ms": sample_ms,
                      "sample_hbm_frac": samples_per_s * 120.0 / 1e9 / hbm_peak,
                      "sample_shape": "%d trajectories x %d instants x (pos..snap) x 3" % (B, M),
                      "extrema_ms": extrema_ms, "extrema_trajectories_per_s": B / (extrema_ms * 1e-3),
                      "extrema_shape": "max |velocity| of %d trajectories x %d segments "
                                       "(computeMaximumOfMagnitude)" % (B, K_SEG),
                      "long_horizon_ms": long_horizon_ms,
                      "long_horizon_shape": "4096 trajectories x 256 segments (configs[3]); HBM floor 0.044 ms",
                      "nccl_gather_ms": gather_ms, "peer_copy_gather_ms": peer_gather_ms},
        }
        emit(line)
    if distributed:
        dist.destroy_process_group()


def ms_launches_per_step(ms):
    """Kernels one solve_standard call launches: 1 on the fast route, 4 on the generic route
    (mask, reorder, pack, general solve)."""
    return 1 if getattr(ms.api, "STANDARD_FAST_ROUTE", False) else 4


_RESULT_FD = None


def emit(line):
    """The one JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    # Libraries print to the C-level stdout behind Python's back (NCCL announces its version there):
    # keep the original stdout for the JSON line only and send everything else to stderr.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
