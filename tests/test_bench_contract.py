"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the
agreed keys; the GPU arm refuses to run without a device (no silent CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True,
                          text=True, timeout=300, env=env)


def test_reference_arm_json_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-budget-s", "1",
             "--rows", "100000", "--dim", "64", "--bits", "64")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("LSH kNN queries/s @k=10")
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["vs_baseline"] is None
    assert d["gpu_launches"] == 0
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = _run("--impl", "reference", "--gpus", "2", "--steps", "1", env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    p = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--rows", "1000")
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)
