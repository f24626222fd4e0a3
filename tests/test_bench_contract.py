"""bench.py's reference arm on the host (no GPU needed): the JSON line keeps the contract's keys, times the oracle
port only, and names the same workload as the GPU arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    assert line["metric"].startswith("min-snap solves/sec") and line["unit"] == "solves/s"
    assert line["higher_is_better"] is True and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["steps"] == 1
    base = line["cpu_baseline"]
    assert base["kind"] == "port" and base["cores"] >= 1 and base["value"] == line["value"] and base["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    # both arms print the same config dict (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.CONFIG
    assert "workload" in line["config"] and "model" not in line["config"]


def test_gpu_arm_refuses_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
