"""Hessian-structured derivative batches (comps = 2, xs_batch.h): n first-order components F_i + one second-order component
S_ij per listed pair, against the DCSFD list (comps = 3) of the same pairs, where (eps1, eps2, eps1eps2) = (F_i, F_j, S_ij).
The list path is the one held against the reference's kernels (tests/test_gpu_stages.py, test_gpu_pipeline.py,
test_gpu_bench_config.py); the truncated algebra makes the two batch kinds equal up to the order of floating-point sums, so
the same tolerances apply through this view.  Stage by stage (integration, raycast, pyramid, ICP normal equations) and through
the frame loop, including a pair subset as a rank of a multi-GPU run holds it."""
import json
import os

import numpy as np
import pytest

from common import H_, ICL, poses_for_frame, rel_err

pytestmark = pytest.mark.gpu


def _save(out_dir, name, obj):
    with open(os.path.join(out_dir, name), "w") as f:
        json.dump(obj, f, indent=1, sort_keys=True)
    print("[parity]", name, json.dumps(obj)[:3000])


def _pose_batches(rng, n, pairs, R, t):
    """Random first / second-order pose derivative components for n parameters and their pairs, as both batch kinds."""
    from xslam_b200 import ops
    dR1 = (H_ * rng.standard_normal((n, 9))).astype(np.float32)
    dt1 = (H_ * rng.standard_normal((n, 3))).astype(np.float32)
    dR2 = (H_ * H_ * rng.standard_normal((len(pairs), 9))).astype(np.float32)
    dt2 = (H_ * H_ * rng.standard_normal((len(pairs), 3))).astype(np.float32)
    hess = ops.PoseBatch(R, t, np.concatenate([dR1, dR2]), np.concatenate([dt1, dt2]))
    lR = np.stack([np.stack([dR1[i], dR1[j], dR2[k]]) for k, (i, j) in enumerate(pairs)]).reshape(-1, 9)
    lt = np.stack([np.stack([dt1[i], dt1[j], dt2[k]]) for k, (i, j) in enumerate(pairs)]).reshape(-1, 3)
    return hess, ops.PoseBatch(R, t, lR, lt)


def _cmp(a, b, floor):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    m = np.isfinite(a) & np.isfinite(b)
    sc = max(np.abs(b[m]).max(), floor)
    d = np.abs(a[m] - b[m])
    return {"max": float(d.max() / sc), "p99.9": float(np.percentile(d, 99.9) / sc)}


def test_hessian_batch_stages_equal_dcsfd_list(xs, out_dir):
    import torch
    from xslam_b200 import ops
    rng = np.random.default_rng(5)
    n, res, voxel = 4, 128, 0.06
    pairs = xs.all_pairs(n)
    m = len(pairs)
    intr = xs.Intr(**ICL)
    vh = ops.TsdfVolume((res,) * 3, voxel, 3.0, comps=2, dirs=n)
    vl = ops.TsdfVolume((res,) * 3, voxel, 3.0, comps=3, dirs=m)
    rep = {}
    for f in (0, 6, 12):
        depth = torch.from_numpy(xs.synth_depth(f).astype(np.int16)).cuda()
        v2c, _, _ = poses_for_frame(xs, f)
        ph, pl = _pose_batches(rng, n, pairs, v2c[:3, :3], v2c[:3, 3])
        uh = ops.integrateTsdfVolume(depth, intr, 100, vh, ph, 0.06)
        ul = ops.integrateTsdfVolume(depth, intr, 100, vl, pl, 0.06)
        assert uh == ul
    assert torch.equal(vh.value(), vl.value()) and torch.equal(vh.weight(), vl.weight())
    g1, g2 = [], []
    for k, (i, j) in enumerate(pairs):
        g1.append(_cmp(vh.grad(i).cpu().numpy(), vl.grad(3 * k).cpu().numpy(), 1e-30)["max"])
        g1.append(_cmp(vh.grad(j).cpu().numpy(), vl.grad(3 * k + 1).cpu().numpy(), 1e-30)["max"])
        g2.append(_cmp(vh.grad(n + k).cpu().numpy(), vl.grad(3 * k + 2).cpu().numpy(), 1e-30))
    rep["integrate"] = {"first_order_max_rel": max(g1), "second_order_max_rel": max(e["max"] for e in g2),
                        "second_order_p99.9_rel": max(e["p99.9"] for e in g2)}
    # raycast from both volumes with both pose kinds
    _, c2v, v2w = poses_for_frame(xs, 12)
    ch, cl = _pose_batches(rng, n, pairs, c2v[:3, :3], c2v[:3, 3])
    wh, wl = _pose_batches(rng, n, pairs, v2w[:3, :3], v2w[:3, 3])
    vmh, nmh = ops.raycast(intr, ch, wh, vh, 480, 640)
    vml, nml = ops.raycast(intr, cl, wl, vl, 480, 640)
    assert torch.equal(vmh[0].isnan(), vml[0].isnan()) and torch.equal(nmh[0].isnan(), nml[0].isnan())
    ok = ~vml[0, 0].isnan()
    assert torch.equal(vmh[0][:, ok], vml[0][:, ok])
    for name, mh, ml in (("vmap", vmh, vml), ("nmap", nmh, nml)):
        mh, ml = mh.cpu().numpy(), ml.cpu().numpy()
        e1 = max(_cmp(mh[1 + i], ml[1 + 3 * k], 1e-30)["max"] for k, (i, j) in enumerate(pairs))
        e2 = [_cmp(mh[1 + n + k], ml[3 + 3 * k], 1e-30) for k in range(m)]
        rep["raycast_" + name] = {"first_order_max_rel": e1, "second_order_max_rel": max(e["max"] for e in e2),
                                  "second_order_p99.9_rel": max(e["p99.9"] for e in e2)}
    # pyramid
    for name, fn, mh, ml in (("vmap", ops.resizeVMap, vmh, vml), ("nmap", ops.resizeNMap, nmh, nml)):
        rh, rl = fn(mh, 2).cpu().numpy(), fn(ml, 3).cpu().numpy()
        assert np.array_equal(rh[0], rl[0], equal_nan=True)
        e1 = max(_cmp(rh[1 + i], rl[1 + 3 * k], 1e-30)["max"] for k, (i, j) in enumerate(pairs))
        e2 = max(_cmp(rh[1 + n + k], rl[3 + 3 * k], 1e-30)["max"] for k in range(m))
        rep["resize_" + name] = {"first_order_max_rel": e1, "second_order_max_rel": e2}
    # ICP normal equations: current frame 13 against the raycast of frame 12
    d13 = torch.from_numpy(xs.synth_depth(13).astype(np.int16)).cuda()
    vc = ops.createVMap(intr, ops.bilateralFilter(d13))
    nc = ops.createNMap(vc)
    c2w = xs.synth_pose(12).astype(np.float64)
    kh, kl = _pose_batches(rng, n, pairs, c2w[:3, :3], c2w[:3, 3])
    prev = ops.PoseBatch(np.linalg.inv(c2w[:3, :3]), c2w[:3, 3], np.zeros((n + m, 9), np.float32), np.zeros((n + m, 3), np.float32))
    prev_l = ops.PoseBatch(np.linalg.inv(c2w[:3, :3]), c2w[:3, 3], np.zeros((3 * m, 9), np.float32), np.zeros((3 * m, 3), np.float32))
    angle = float(np.sin(np.radians(15.0)))
    Ah, bh = ops.estimateCombined(kh, vc, nc, prev, intr, vmh, nmh, 0.10, angle, comps=2)
    Al, bl = ops.estimateCombined(kl, vc, nc, prev_l, intr, vml, nml, 0.10, angle, comps=3)
    assert np.array_equal(Ah[0], Al[0]) and np.array_equal(bh[0], bl[0])
    e1 = max(rel_err(Ah[1 + i], Al[1 + 3 * k]) for k, (i, j) in enumerate(pairs))
    e2 = max(rel_err(Ah[1 + n + k], Al[3 + 3 * k]) for k in range(m))
    eb = max(rel_err(bh[1 + n + k], bl[3 + 3 * k]) for k in range(m))
    rep["icp"] = {"A_first_order_rel": e1, "A_second_order_rel": e2, "b_second_order_rel": eb, "A00": float(Ah[0][0, 0])}
    _save(out_dir, "hessian_batch_stages.json", rep)
    assert Ah[0][0, 0] > 1000
    # same algebra, different summation / evaluation order: far inside the tolerances stated against the reference
    assert rep["integrate"]["first_order_max_rel"] <= 1e-5 and rep["integrate"]["second_order_p99.9_rel"] <= 1e-5
    assert rep["integrate"]["second_order_max_rel"] <= 1e-3
    for name in ("raycast_vmap", "raycast_nmap"):
        assert rep[name]["first_order_max_rel"] <= 1e-4 and rep[name]["second_order_p99.9_rel"] <= 1e-5 and rep[name]["second_order_max_rel"] <= 1e-2
    for name in ("resize_vmap", "resize_nmap"):
        assert rep[name]["first_order_max_rel"] <= 1e-5 and rep[name]["second_order_max_rel"] <= 1e-4
    assert rep["icp"]["A_first_order_rel"] <= 1e-6 and rep["icp"]["A_second_order_rel"] <= 1e-5 and rep["icp"]["b_second_order_rel"] <= 1e-4


@pytest.mark.parametrize("mode", ["all_pairs", "pair_subset"])
def test_hessian_batch_pipeline_equals_dcsfd_list(xs, out_dir, mode):
    """Three frames of the frame loop at 256^3: world2camera, TSDF planes and raycast maps of the Hessian batch against the DCSFD
    list of the same pairs.  pair_subset is the share of a rank in a multi-GPU run: all first-order components, some pairs."""
    n = 4
    U = np.eye(6)[[0, 2, 4, 5]]
    pairs = xs.all_pairs(n) if mode == "all_pairs" else [(0, 2), (1, 1), (1, 3), (3, 3)]
    hs, _ = xs.hessian_seeds(U, pairs)
    G = np.tensordot(U, xs.se3_generators(), 1)
    ls = np.zeros((len(pairs), 3, 16))
    for k, (i, j) in enumerate(pairs):
        ls[k, 0], ls[k, 1] = (H_ * G[i]).reshape(16), (H_ * G[j]).reshape(16)
        ls[k, 2] = (H_ * H_ * 0.5 * (G[i] @ G[j] + G[j] @ G[i])).reshape(16)
    cfg = dict(xs.DEFAULT_CONFIG)
    KF = xs.KinectFusionReconstruction
    kh, kl = KF(), KF()
    kh.SetYamlParameters(cfg, comps=2, seeds=hs, pairs=pairs, n_params=n)
    kl.SetYamlParameters(cfg, comps=3, seeds=ls.reshape(-1, 16).astype(np.float32))
    kh.enable_icp_log()  # with the log the derivative pass forms the full second-order sums A_ij, b_ij (27 per pair)
    kl.enable_icp_log()
    # without it (the production path) the pair tasks accumulate g_ij = b_ij - A_ij x directly (6 values per pair, FP32 per
    # pixel): a third instance holds that form against the full sums
    kr = KF()
    kr.SetYamlParameters(cfg, comps=2, seeds=hs, pairs=pairs, n_params=n)
    rep = {"frames": []}
    for f in range(3):
        d = xs.synth_depth(f)
        assert kh.ProcessFrame(d) == 1 and kl.ProcessFrame(d) == 1 and kr.ProcessFrame(d) == 1
        wh, wl, wr = kh.world2camera, kl.world2camera, kr.world2camera
        fr = {"real_identical": bool(np.array_equal(wh[0], wl[0]) and np.array_equal(wr[0], wh[0]))}
        fr["reduced_first_order_rel"] = max(rel_err(wr[1 + i], wh[1 + i], floor=H_ * 1e-3) for i in range(n))
        fr["reduced_second_order_rel"] = max(rel_err(wr[1 + n + k], wh[1 + n + k], floor=H_ * H_ * 1e-3) for k in range(len(pairs)))
        fr["pose_first_order_rel"] = max(max(rel_err(wh[1 + i], wl[1 + 3 * k], floor=H_ * 1e-3), rel_err(wh[1 + j], wl[2 + 3 * k], floor=H_ * 1e-3))
                                         for k, (i, j) in enumerate(pairs))
        fr["pose_second_order_rel"] = max(rel_err(wh[1 + n + k], wl[3 + 3 * k], floor=H_ * H_ * 1e-3) for k in range(len(pairs)))
        if f > 0:
            lh, ll = kh.icp_log(), kl.icp_log()
            assert lh.shape[0] == ll.shape[0] == 12
            fr["icp_real_identical"] = bool(np.array_equal(lh[:, 0], ll[:, 0]))
            fr["icp_second_order_rel"] = max(rel_err(lh[it, 1 + n + k], ll[it, 3 + 3 * k]) for it in range(12) for k in range(len(pairs)))
        rep["frames"].append(fr)
    vh, wgh, _ = kh.volume_planes(0)
    vl, wgl, _ = kl.volume_planes(0)
    rep["volume_identical"] = bool((vh == vl).all() and (wgh == wgl).all())
    k, (i, j) = len(pairs) - 1, pairs[-1]
    rep["grad_second_order_rel"] = _cmp(kh.volume_planes(n + k)[2].cpu().numpy(), kl.volume_planes(3 * k + 2)[2].cpu().numpy(), 1e-30)
    rep["grad_first_order_rel"] = _cmp(kh.volume_planes(i)[2].cpu().numpy(), kl.volume_planes(3 * k)[2].cpu().numpy(), 1e-30)
    mh, ml = kh.map("nmap_g_prev", 1).cpu().numpy(), kl.map("nmap_g_prev", 1).cpu().numpy()
    rep["maps_real_identical"] = bool(np.array_equal(mh[0], ml[0], equal_nan=True))
    rep["nmap_l1_second_order_rel"] = _cmp(mh[1 + n + k], ml[3 + 3 * k], 1e-30)
    _save(out_dir, "hessian_batch_pipeline_%s.json" % mode, rep)
    last = rep["frames"][-1]
    assert all(fr["real_identical"] for fr in rep["frames"]) and rep["volume_identical"] and rep["maps_real_identical"]
    assert last["icp_real_identical"]
    assert last["pose_first_order_rel"] <= 1e-4 and last["pose_second_order_rel"] <= 1e-3
    assert max(fr["reduced_first_order_rel"] for fr in rep["frames"]) <= 1e-6
    assert max(fr["reduced_second_order_rel"] for fr in rep["frames"]) <= 1e-3
    assert last["icp_second_order_rel"] <= 1e-3
    assert rep["grad_first_order_rel"]["p99.9"] <= 1e-5 and rep["grad_second_order_rel"]["p99.9"] <= 1e-4
    assert rep["nmap_l1_second_order_rel"]["p99.9"] <= 1e-4


def test_blocked_shards_reassemble_the_full_batch(xs, out_dir):
    """The multi-GPU decomposition on one GPU: every rank's share of an 8-rank plan (parallel.plan_hessian_shards: a block of the
    pairs + the parameters those pairs touch, including intrinsic parameters) is run as its own pipeline, the records are
    reassembled with the code the N-rank runs use, and the result must equal the record of the full batch - real part bit for
    bit, derivative components to FP32 summation noise (different component sets sum their ICP normal equations in different
    orders and may take different kernel forms)."""
    from xslam_b200 import parallel
    W, Hh = 160, 120
    intr = (481.20 / 4, -480.00 / 4, 319.50 / 4, 239.50 / 4)
    cfg = dict(xs.DEFAULT_CONFIG)
    cfg.update(tsdf_size_x=64, tsdf_size_y=64, tsdf_size_z=64, tsdf_voxel_size=0.12, depth_width=W, depth_height=Hh,
               fx=intr[0], fy=intr[1], cx=intr[2], cy=intr[3])
    frames = [xs.synth_depth(f, W, Hh, *intr) for f in range(3)]
    n, world = 8, 4  # 6 pose DoF + fx, cx; 36 pairs
    U = np.concatenate([np.eye(6), np.zeros((2, 6))])
    dintr = np.zeros((n, 4), np.float32)
    dintr[6, 0] = dintr[7, 2] = H_
    pairs = xs.all_pairs(n)

    def run(params, local_pairs):
        seeds, lp = xs.hessian_seeds(U[params], local_pairs)
        di = np.ascontiguousarray(dintr[params])
        k = xs.KinectFusionReconstruction()
        k.SetYamlParameters(cfg, comps=2, seeds=seeds, pairs=lp, n_params=len(params), intrinsic_seeds=di if di.any() else None)
        for d in frames:
            assert k.ProcessFrame(d) == 1
        return k.world2camera.reshape(-1, 16)

    full = run(list(range(n)), pairs)
    plan = parallel.plan_hessian_shards(n, pairs, world)
    L = parallel.planned_record_floats(plan)
    gathered = np.zeros((world, L), np.float32)
    for r, sh in enumerate(plan):
        rec = run(sh["params"], sh["local_pairs"]).reshape(-1)
        gathered[r, : rec.size] = rec
    rec = parallel.assemble_planned_records(gathered, plan, n, len(pairs))
    sc1, sc2 = np.abs(full[1:1 + n]).max(), np.abs(full[1 + n:]).max()
    rep = {"planes_per_rank": [len(sh["params"]) + len(sh["pair_ids"]) for sh in plan],
           "real_identical": bool(np.array_equal(rec[0], full[0])),
           "first_order_rel": float(np.abs(rec[1:1 + n] - full[1:1 + n]).max() / sc1),
           "second_order_rel": float(np.abs(rec[1 + n:] - full[1 + n:]).max() / sc2)}
    _save(out_dir, "hessian_blocked_shards.json", rep)
    assert rep["real_identical"]
    assert rep["first_order_rel"] <= 2e-6 and rep["second_order_rel"] <= 5e-6
