"""bench.py contract on a machine without a GPU: the reference arm (CPU: the unmodified reference when a reference tree is
present — /root/reference or baseline/_ref — else the oracle port) prints one
JSON line with the driver's keys; the product arm refuses to run without a CUDA device (there is no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, timeout=600)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-clips", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "clips/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert "workload" in line["config"]
    sys.path.insert(0, ROOT)
    from oracle import ref_loader
    want = "reference" if ref_loader.available() else "port"
    assert line["cpu_baseline"]["kind"] == want and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    if want == "reference":
        assert "unmodified reference" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == dict(value=line["value"], unit="clips/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_product_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return                      # on the GPU box the product arm is exercised by the driver itself
    r = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline")
    assert r.returncode != 0        # fails loudly: no CPU fallback for the CUDA path
    assert "CUDA" in (r.stderr + r.stdout) or "cuda" in (r.stderr + r.stdout)
