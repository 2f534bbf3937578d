"""Parity on the configuration bench.py actually runs (BASELINE.json metric: 640x480 depth, 512^3 TSDF, DCSFD directions).

The reference has no DCSFD frame loop, but a bicomplex direction's first-order components (eps1, eps2) are exactly what
one-direction complex runs of the reference yield for those imaginary seeds.  So a handful of the benchmark's own 55
directions - axis pairs and two of the deterministic mixed pose-space pairs - are run batched through the product and,
component by component, through the reference's own kernels (oracle/_ref/libxslam_ref.so + the restated orchestrator):
poses, TSDF planes and raycast maps, with the tolerances of tests/test_gpu_pipeline.py::test_pipeline_csfd_vs_reference.
The product runs in XS_SOLVE_EIGEN_LLT mode here (first-order components through the Hermitian LLT, as the reference's
host code does); the benchmark's XS_SOLVE_ANALYTIC mode is held against it in the same test (identical real parts, first
order within the size of the LLT quirk).
"""
import json
import os

import numpy as np
import pytest

from common import H_, rel_err

pytestmark = pytest.mark.gpu


def _save(out_dir, name, obj):
    with open(os.path.join(out_dir, name), "w") as f:
        json.dump(obj, f, indent=1, sort_keys=True)
    print("[parity]", name, json.dumps(obj)[:3000])


def test_bench_workload_first_order_vs_reference_512(xs, refcuda, out_dir):
    import bench
    cfg = bench.workload_cfg(xs, 512)
    seeds_all = bench.all_directions(xs, 3, 55).reshape(55, 3, 16)
    pick = [0, 8, 20, 30, 54]  # pairs (0,0), (1,3), (5,5) of the pose axes and two of the mixed pose-space pairs
    seeds = np.ascontiguousarray(seeds_all[pick].reshape(-1, 16))
    frames = [xs.synth_depth(f) for f in range(3)]
    KF = xs.KinectFusionReconstruction
    k = KF()
    k.SetYamlParameters(cfg, comps=3, seeds=seeds, solve_mode=KF.SOLVE_EIGEN_LLT)
    ka = KF()
    ka.SetYamlParameters(cfg, comps=3, seeds=seeds, solve_mode=KF.SOLVE_ANALYTIC)  # the mode bench.py runs
    # reference passes: zero seed, then eps1 and eps2 of every picked direction (one complex direction per pass)
    ref0 = refcuda.kinfu(cfg, None)
    refs = {}
    for n, d in enumerate(pick):
        for c in (0, 1):
            if c == 1 and np.array_equal(seeds_all[d, 0], seeds_all[d, 1]):
                refs[(n, 1)] = refs[(n, 0)]
                continue
            refs[(n, c)] = refcuda.kinfu(cfg, seeds_all[d, c].reshape(4, 4))
    uniq = list({id(r): r for r in refs.values()}.values())
    rep = {"directions": pick, "frames": []}
    for f, d in enumerate(frames):
        assert k.ProcessFrame(d) == 1 and ka.ProcessFrame(d) == 1
        assert ref0.process_frame(d) == 1
        for r in uniq:
            assert r.process_frame(d) == 1
        w, wa = k.world2camera, ka.world2camera
        fr = {"frame": f, "pose_real_abs_vs_zero_seed": float(np.abs(w[0] - ref0.pose().real).max()),
              "analytic_real_identical": bool(np.array_equal(w[0], wa[0]))}
        fr["pose_first_order_rel"] = [[rel_err(w[1 + 3 * n + c], refs[(n, c)].pose().imag, floor=H_ * 1e-3) for c in (0, 1)]
                                      for n in range(len(pick))]
        fr["analytic_vs_llt_first_order_rel"] = [[rel_err(wa[1 + 3 * n + c], w[1 + 3 * n + c], floor=H_ * 1e-3) for c in (0, 1)]
                                                 for n in range(len(pick))]
        rep["frames"].append(fr)
    # volume state after 3 frames: weights / values against the zero-seed pass, two derivative planes against their passes
    v, wgt, _ = k.volume_planes(0)
    rv, rw, _ = ref0.volume()
    upd = int((rw > 0).sum())
    rep["updated_voxels"] = upd
    rep["weight_mismatch"] = int((wgt.cpu().numpy() != rw).sum())
    rep["value_rel"] = rel_err(v.cpu().numpy(), rv)
    del v, wgt, rv, rw
    rep["grad_rel"] = {}
    for n, c in ((1, 0), (3, 1)):  # eps1 of pair (1,3), eps2 of a mixed pair
        _, _, g = k.volume_planes(3 * n + c)
        _, _, rg = refs[(n, c)].volume()
        g = g.cpu().numpy()
        sc = np.abs(rg).max()
        d = np.abs(g - rg)
        rep["grad_rel"]["dir%d_c%d" % (pick[n], c)] = {"max": float(d.max() / sc), "p99.9": float(np.percentile(d[rg != 0], 99.9) / sc)}
        del g, rg, d
    # raycast maps of the last frame
    vm = k.map("vmap_g_prev", 0).cpu().numpy()
    nm = k.map("nmap_g_prev", 0).cpu().numpy()
    rvm, rnm = ref0.map("vmap_g_prev", 0), ref0.map("nmap_g_prev", 0)
    valid = ~np.isnan(rvm[0, ..., 0])
    rep["raycast_mask_mismatch"] = int((np.isnan(vm[0, 0]) != ~valid).sum()) + int((np.isnan(nm[0, 0]) != np.isnan(rnm[0, ..., 0])).sum())
    both = valid & ~np.isnan(vm[0, 0])
    rep["raycast_real_rel"] = max(rel_err(vm[0, p][both], rvm[p, ..., 0][both]) for p in range(3))
    rep["raycast_first_order_rel"] = {}
    for n, c in ((0, 0), (2, 1), (4, 0)):
        r = refs[(n, c)].map("vmap_g_prev", 0)
        b = both & ~np.isnan(r[0, ..., 0])
        comp = 1 + 3 * n + c
        sc = max(np.abs(r[p, ..., 1][b]).max() for p in range(3))
        d = np.concatenate([np.abs(vm[comp, p][b] - r[p, ..., 1][b]) for p in range(3)])
        rep["raycast_first_order_rel"]["dir%d_c%d" % (pick[n], c)] = {"max": float(d.max() / sc), "p99.9": float(np.percentile(d, 99.9) / sc)}
    _save(out_dir, "bench_config_parity_512.json", rep)
    last = rep["frames"][-1]
    # integers / real parts.  A multi-frame run is bit-exact while the poses are: after ICP the real pose may differ by an ulp
    # from the reference's (different summation order of the normal equations), which can move a handful of voxels / pixels
    # across a rounding boundary; the measured counts are in the report (and in DESIGN.md 3)
    assert rep["frames"][0]["pose_real_abs_vs_zero_seed"] == 0.0
    assert last["pose_real_abs_vs_zero_seed"] <= 1e-5
    assert rep["weight_mismatch"] <= 1e-4 * upd and rep["raycast_mask_mismatch"] <= 1e-4 * 640 * 480
    assert rep["value_rel"] <= 1e-5 and rep["raycast_real_rel"] <= 1e-5
    assert all(fr["analytic_real_identical"] for fr in rep["frames"])
    # first-order components of the benchmark's directions against the reference's one-direction passes
    assert max(max(p) for p in last["pose_first_order_rel"]) <= 1.5e-4
    assert all(e["p99.9"] <= 5e-5 and e["max"] <= 2e-3 for e in rep["grad_rel"].values())
    assert all(e["p99.9"] <= 1e-4 and e["max"] <= 1e-2 for e in rep["raycast_first_order_rel"].values())
    # the analytic solve is the true derivative; it differs from the Hermitian-LLT one by the size of that quirk only
    assert max(max(p) for p in last["analytic_vs_llt_first_order_rel"]) <= 5e-2
