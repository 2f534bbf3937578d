"""The drop-in boundary at RUN time: oracle/_ref/seam_run links the wrappers with the reference's signatures
(include/xslam_b200.hpp -> libxslam_b200.so) AND the reference's own operators (libxslam_ref.so), drives both with the
reference's DeviceArray2D / MatS33 / devComplex3 / Intr objects through the frame loop's call sequence
(initVolume -> SurfaceMeasure -> integrateTsdfVolume -> raycast -> resizeV/NMap -> estimateCombined -> extractPoints /
extractNormals) and compares every output.  The binary is built where /root/reference exists (make -C oracle ref) and
travels to the GPU box."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "seam_run")


@pytest.mark.gpu
def test_seam_wrappers_run_against_the_reference_operators(out_dir):
    assert os.path.exists(BIN), "oracle/_ref/seam_run is not built (make -C oracle ref; needs /root/reference)"
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=600, cwd=ROOT)
    with open(os.path.join(out_dir, "seam_run.json"), "w") as f:
        f.write(r.stdout)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stderr[-2000:]
    rep = json.loads(r.stdout)
    for key in ("depth_l2", "vmap_curr_l0", "nmap_curr_l1", "volume", "vmap_g_prev_l0", "nmap_g_prev_l2", "icp_l0", "icp_l2", "extract"):
        assert key in rep
    assert rep["volume"]["weight_mismatch"] == 0 and rep["volume"]["value_ulp_gt0"] == 0
    assert rep["extract"]["point_mismatch"] == 0 and rep["extract"]["points"] == rep["extract"]["points_ref"] > 200


def test_seam_header_lists_every_operator_of_the_boundary():
    """every free function the orchestrator calls has a wrapper with the reference's name."""
    src = open(os.path.join(ROOT, "include", "xslam_b200.hpp")).read()
    for name in ("bilateralFilter", "pyrDown", "createVMap", "createNMap", "resizeVMap", "resizeNMap", "initVolume",
                 "integrateTsdfVolume", "raycast", "estimateCombined", "ComputeLocalTsdf_hessian", "ComputeLocalTsdf_loss",
                 "extractPoints", "extractNormals", "computeOptimizeMatrix"):
        assert (" %s(" % name) in src, name
