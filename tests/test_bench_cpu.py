"""bench.py contract on a CPU-only box: the reference arm (`--impl reference`: the oracle port on the host cores) prints ONE JSON
line with the keys the driver reads; the GPU arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mcells/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("Mcells/s interpolated") and d["value"] > 0 and d["n_gpus"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "rows" in cb["sample"]
    assert d["config"]["workload"].startswith("BASELINE config 3") and d["dtype"] == "f64" and d["data"] == "synthetic"


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert not any(ln.startswith("{") for ln in out.stdout.splitlines())
