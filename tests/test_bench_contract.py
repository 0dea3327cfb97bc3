"""CPU test of bench.py's reference arm: it must run without a GPU, print ONE JSON line and carry the keys the
driver reads (metric/value/unit/..., impl, cpu_baseline, e2e with zero copy bytes)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e", "gpu_launches"}


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1",
                          "--nucleotides", str(1 << 22), "--cpu-sample", str(1 << 22)],
                         capture_output=True, text=True, check=True, cwd=ROOT).stdout.strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    assert REQUIRED <= set(line)
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == "nucleotides/s"
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert "workload" in line["config"] and line["vs_baseline"] is None and line["dtype"] == "u8"
    # the config object is built by ONE function for both arms, so the driver's same_config check can hold
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.shared_config(argparse.Namespace(nucleotides=1 << 22, alphabet=10))


def test_non_zero_rank_of_reference_arm_is_silent():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_fall_back():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
