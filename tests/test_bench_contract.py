"""bench.py's JSON-line contract on a machine without a GPU: the reference arm (CPU port of the reference path)
must run here, print exactly one JSON line and describe the same workload as the product arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(*flags, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_line():
    line = _bench("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert line["impl"] == "reference"
    assert line["metric"] == "PointNet++ SSG fwd clouds/sec (N=1024)" and line["unit"] == "clouds/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["gpu_launches"] == 0
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["value"] == line["value"] and "24 clouds" in cpu["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the SAME config object as the product arm (the driver compares them); the bounded sample is named inside it and
    # described in cpu_baseline.sample
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.workload_config(0, 2048, 1)
    assert "24-cloud" in line["config"]["reference_arm_sample"]


def test_reference_arm_of_the_microbenchmark_config():
    line = _bench("--config", "2", "--impl", "reference", "--steps", "1", "--warmup", "0")
    assert line["impl"] == "reference" and line["unit"] == "us" and line["higher_is_better"] is False
    assert line["value"] == line["cpu_fps_us"] + line["cpu_ball_query_us"] > 0
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.cfg2_config(1)


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0          # no CPU fallback: the product arm must not print a number here
    assert not any(ln.startswith("{") for ln in out.stdout.splitlines())


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29541", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
