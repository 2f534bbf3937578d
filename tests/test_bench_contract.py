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
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    # the line reports what was timed: ms_per_step is ONE measured frame-pass, the scaling to a differentiated frame is explicit
    assert d["passes_per_differentiated_frame"] == 6
    assert abs(d["value"] - 1.0 / (d["ms_per_step"] * 1e-3 * d["passes_per_differentiated_frame"])) <= 1e-9 * d["value"]
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


def test_reference_arm_does_not_load_the_product():
    """The reference arm must not import the product package: its process maps no x-slam_b200/*.so."""
    code = ("import sys, os; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--res', '64', '--dirs', '1', '--comps', '1'];"
            "import runpy; runpy.run_path(%r, run_name='__main__');"
            "maps = open('/proc/self/maps').read(); assert 'libxslam_b200' not in maps, 'product library mapped';"
            "assert 'xslam_b200' not in sys.modules; print('clean')") % os.path.join(ROOT, "bench.py")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "clean" in r.stdout, r.stderr[-2000:]


def test_checker_side_synthetic_stream_equals_the_products():
    """Both arms of the benchmark see identical input: the generator restated under oracle/ is bit-identical to
    xs_synth_depth / xs_synth_pose (host code of the product library; no GPU needed)."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    import xslam_b200 as xs
    from oracle import pyref
    o = pyref.Oracle()
    for f in (0, 1, 37, 299, 300):
        assert np.array_equal(bench.oracle_synth_depth(o, f), xs.synth_depth(f))


def test_reference_arm_uses_all_host_threads_under_torchrun_env():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm overrides it explicitly."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    code = ("import sys; sys.path.insert(0, %r); import bench, ctypes, os; n = bench.use_all_host_threads();"
            "from oracle import pyref; pyref.Oracle();"
            "g = ctypes.CDLL('libgomp.so.1'); assert g.omp_get_max_threads() == n == (os.cpu_count() or 1), (g.omp_get_max_threads(), n); print('ok')") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
