"""bench.py contract checks that run without a GPU: the reference arm (the CPU path timed on the host cores) prints ONE JSON line
with the keys the driver reads, and the GPU arm fails loudly instead of falling back to a CPU path."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--res", "64", "--dirs", "6", "--comps", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "differentiated_frames_per_s" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] >= 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_gpu_arm_has_no_cpu_fallback():
    if _have_gpu():
        import pytest
        pytest.skip("checks the no-GPU failure mode")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--res", "64", "--dirs", "1",
                        "--no-cpu-baseline", "--no-ref-cuda"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0, "bench.py must not produce a number without a CUDA device"
    assert not [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
