"""CPU: the reference arm of bench.py (the only arm that runs without a GPU) prints one JSON line with the contract's
keys; the product arm refuses to run without a B200 instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def run_bench(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                          env=env, timeout=600)


def test_reference_arm_json_line():
    r = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--objects", "4", "--rows", "2000",
                  "--keypoints", "64", "--cpu-sample-queries", "16", "--frames", "2")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "u8" and d["value"] > 0
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"])
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    r = run_bench("--steps", "1", "--warmup", "1", "--objects", "2", "--rows", "500", "--keypoints", "32", "--frames", "1")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stdout + r.stderr)
