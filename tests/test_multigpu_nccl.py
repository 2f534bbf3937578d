"""The N-rank path on hardware (needs >= 2 GPUs on the box; skipped otherwise): tools/check_gather.py under torchrun -
sharded Hessian batch + the library's own NCCL all-gather against the 1-rank run of the full batch."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_gathered_record_equals_one_rank_record():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run by tools/gpu_scale.sh on a multi-GPU box)")
    world = 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tools", "check_gather.py")], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    rep = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert rep["ok"] and rep["world"] == world
    assert all(fr["real_identical_to_1rank"] and fr["replicas_first_order_spread_rel"] <= 1e-6 for fr in rep["frames"])


def test_library_exports_the_comm_layer(xs):
    """The multi-GPU layer lives in the library (C-ABI), not in bench.py: symbols present, NCCL bound at run time."""
    lib = xs.load()
    for name in ("xs_comm_unique_id", "xs_comm_create", "xs_comm_destroy", "xs_comm_all_gather", "xs_kinfu_set_comm",
                 "xs_kinfu_get_gathered_records", "xs_kinfu_get_gathered_records_lagged", "xs_set_device"):
        assert hasattr(lib, name)
    from xslam_b200 import parallel
    uid = parallel.Comm.unique_id()  # ncclGetUniqueId through the dlopen'ed library (no GPU needed)
    assert len(uid) == 128 and any(uid)
