"""bench.py contract on the CPU: the reference arm's JSON line (the only leg that runs without a GPU) and the loud
failure of the product arm when there is no CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, env=None):
    e = dict(os.environ); e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT, env=e)


def test_reference_arm_line():
    p = run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "shots/s" and line["higher_is_better"] is True
    assert line["vs_baseline"] is None and line["gpu_launches"] == 0 and line["value"] > 0
    assert "workload" in line["config"] and line["config"]["k"] == 64 and line["config"]["patches"] == 16469
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    # a fresh scene has 99 light patches: the first batch of 64 is full, so one step counts 64 shots
    assert abs(line["value"] * line["ms_per_step"] * 1e-3 - 64) < 1e-6


def test_reference_arm_mirrors_weak_scaling_batch():
    env = {"RANK": "0", "WORLD_SIZE": "1"}
    p = run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["scaling"] == "weak" and line["config"]["k"] == 128 and line["n_gpus"] == 2
    # ... and only the 99 lights of the fresh scene can be shot in that first batch of 128 slots
    assert abs(line["value"] * line["ms_per_step"] * 1e-3 - 99) < 1e-6
    # the other ranks of a torchrun launch exit without work
    p = run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    p = run("--steps", "1", "--warmup", "0")
    assert p.returncode != 0 and "no CUDA device" in (p.stderr + p.stdout)
