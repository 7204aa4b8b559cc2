"""bench.py's reference arm runs without a GPU (it times the oracle on the host cores): check the JSON contract of that line on the
small C1 workload, and that the default arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_json_line():
    p = run("--impl", "reference", "--workload", "c1", "--steps", "40", "--warmup", "5")
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "impl",
                "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "mpgp_iterations_per_second" and line["unit"] == "it/s"
    assert line["vs_baseline"] is None and line["dtype"] == "f64" and line["data"] == "synthetic" and line["gpu_launches"] == 0
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "workload" in line["config"]
    assert line["config"]["workload"].startswith("C1 ")


def test_default_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return   # on a GPU box this arm is exercised by the driver itself
    p = run("--workload", "c1", "--steps", "5", "--warmup", "3")
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout) or "no CPU" in (p.stderr + p.stdout)
