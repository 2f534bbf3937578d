"""Size-independent properties of the CUDA path, checked at BASELINE.json's full sizes (640x480, 512^3) where the
oracle is too slow, plus a finite-difference check of the DCSFD second-order components (oracle-free):

  * linearity: first-order derivative outputs are linear in the seeds (a*G1 + b*G2 -> a*d1 + b*d2), and the real part
    does not depend on the seeds at all (bit-exact);
  * the eps1 / eps2 components of a DCSFD run equal the CSFD run with the same seeds, and a symmetric pair (i, j) vs
    (j, i) gives the same eps1eps2 component;
  * Hessian by finite differences: d/d(theta_j) of the CSFD gradient, by central differences over two CSFD runs started
    from exp(+-delta G_j), agrees with the eps1eps2 component of one DCSFD run.
"""
import os

import numpy as np
import pytest

H_ = 1e-7


def _run(xs, cfg, comps, seeds, frames, w2c0=None):
    k = xs.KinectFusionReconstruction()
    k.SetYamlParameters(cfg, comps=comps, seeds=seeds, solve_mode=xs.KinectFusionReconstruction.SOLVE_ANALYTIC)
    if w2c0 is not None:
        k.world2camera = w2c0
    for d in frames:
        assert k.ProcessFrame(d) == 1
    return k


@pytest.mark.gpu
def test_full_size_linearity_and_seed_independence(xs):
    cfg = dict(xs.DEFAULT_CONFIG)
    cfg.update(tsdf_size_x=512, tsdf_size_y=512, tsdf_size_z=512, tsdf_voxel_size=0.015)
    frames = [xs.synth_depth(f) for f in range(3)]
    G = xs.se3_generators().reshape(6, 16)
    a, b = 0.7, -1.3
    seeds = np.stack([H_ * G[1], H_ * G[4], H_ * (a * G[1] + b * G[4])]).astype(np.float32)
    k = _run(xs, cfg, 1, seeds, frames)
    k0 = _run(xs, cfg, 1, None, frames)
    w = k.world2camera
    assert np.array_equal(w[0], k0.world2camera[0]), "the real pose depends on the seeds"
    v, wt, _ = k.volume_planes(0)
    v0, wt0, _ = k0.volume_planes(0)
    assert bool((wt == wt0).all()) and bool((v == v0).all()), "the real volume depends on the seeds"
    lin = a * w[1] + b * w[2]
    scale = np.abs(w[3]).max()
    assert np.abs(w[3] - lin).max() <= 2e-4 * scale, (np.abs(w[3] - lin).max(), scale)
    # derivative planes of the volume and raycast maps are linear in the seeds too
    g = [k.volume_planes(q)[2] for q in range(3)]
    err = float((g[2] - (a * g[0] + b * g[1])).abs().max())
    assert err <= 2e-4 * float(g[2].abs().max()), err
    m = k.map("vmap_g_prev")
    ok = ~np.isnan(m[0, 0].cpu().numpy())
    mm = m.cpu().numpy()[:, :, ok]
    assert np.abs(mm[3] - (a * mm[1] + b * mm[2])).max() <= 2e-4 * np.abs(mm[3]).max()


@pytest.mark.gpu
def test_dcsfd_first_order_and_symmetry_full_size(xs):
    cfg = dict(xs.DEFAULT_CONFIG)
    cfg.update(tsdf_size_x=512, tsdf_size_y=512, tsdf_size_z=512, tsdf_voxel_size=0.015)
    frames = [xs.synth_depth(f) for f in range(3)]
    s3, pairs = xs.pose_seeds_dcsfd(pairs=[(0, 4), (4, 0), (2, 2)])
    k3 = _run(xs, cfg, 3, s3, frames)
    k1 = _run(xs, cfg, 1, xs.pose_seeds_csfd()[[0, 4, 2]], frames)
    w3, w1 = k3.world2camera, k1.world2camera
    assert np.array_equal(w3[0], w1[0])
    sc = np.abs(w1[1:]).max()
    assert np.abs(w3[1] - w1[1]).max() <= 1e-4 * sc and np.abs(w3[2] - w1[2]).max() <= 1e-4 * sc   # (0,4): eps1 = d0, eps2 = d4
    assert np.abs(w3[4] - w1[2]).max() <= 1e-4 * sc and np.abs(w3[5] - w1[1]).max() <= 1e-4 * sc   # (4,0) swapped
    assert np.abs(w3[7] - w1[3]).max() <= 1e-4 * sc and np.abs(w3[8] - w1[3]).max() <= 1e-4 * sc   # (2,2)
    s12 = np.abs(w3[3]).max()
    assert np.abs(w3[3] - w3[6]).max() <= 2e-3 * s12, "eps1eps2 is not symmetric in the pair"


def _pose_family(V0, A, B):
    """pose(t1, t2) = exp(t1 A + t2 B) V0 and its derivatives, in float64 (A, B: 4x4 se(3) generators)."""
    from scipy.linalg import expm

    def base(t2):
        return expm(t2 * B) @ V0

    def d1(t2, eps=1e-6):  # d/dt1 at t1 = 0
        return (expm(eps * A + t2 * B) - expm(-eps * A + t2 * B)) @ V0 / (2 * eps)

    return base, d1, B @ V0, 0.5 * (A @ B + B @ A) @ V0


def _batch(ops, M, derivs):
    dR = np.stack([d[:3, :3].reshape(9) for d in derivs]) if derivs else None
    dt = np.stack([d[:3, 3] for d in derivs]) if derivs else None
    return ops.PoseBatch(M[:3, :3], M[:3, 3], dR, dt)


@pytest.mark.gpu
@pytest.mark.parametrize("threshold", [0.0, 0.06])
@pytest.mark.parametrize("pair", [(0, 4), (5, 2), (3, 3)])
def test_integration_second_order_vs_finite_differences(xs, pair, threshold):
    """eps1eps2 plane of one DCSFD integration == central difference (over the second parameter) of the eps plane of two
    CSFD integrations.  The CSFD plane is itself pinned to the reference kernel (test_gpu_stages.py), so this pins the
    Hessian path (seeded J / H jets + chain rule) of integrate_kernel<3>, which the reference has no counterpart of."""
    import torch
    from xslam_b200 import ops
    from common import ICL, poses_for_frame
    G = xs.se3_generators()
    A, B = G[pair[0]], G[pair[1]]
    V0, _, _ = poses_for_frame(xs, 3)
    base, d1, dB, d12 = _pose_family(V0, A, B)
    res, voxel = 128, 0.06
    intr = xs.Intr(**ICL)
    depth = torch.from_numpy(xs.synth_depth(3).astype(np.int16)).cuda()
    vol3 = ops.TsdfVolume((res,) * 3, voxel, 3, comps=3, dirs=1)
    ops.integrateTsdfVolume(depth, intr, 100, vol3, _batch(ops, base(0.0), [H_ * d1(0.0), H_ * dB, H_ * H_ * d12]), threshold)
    g12 = vol3.grad(2).cpu().numpy().astype(np.float64) / H_ / H_
    g1 = vol3.grad(0).cpu().numpy().astype(np.float64) / H_
    delta = 1e-3
    planes, weights, values = [], [], []
    for sgn in (+1.0, -1.0):
        vol1 = ops.TsdfVolume((res,) * 3, voxel, 3, comps=1, dirs=1)
        ops.integrateTsdfVolume(depth, intr, 100, vol1, _batch(ops, base(sgn * delta), [H_ * d1(sgn * delta)]), threshold)
        planes.append(vol1.grad(0).cpu().numpy().astype(np.float64) / H_)
        weights.append(vol1.weight().cpu().numpy())
        values.append(vol1.value().cpu().numpy())
    fd = (planes[0] - planes[1]) / (2 * delta)
    w3, v3 = vol3.weight().cpu().numpy(), vol3.value().cpu().numpy()
    # truncation-band voxels updated in all three runs (saturated voxels have zero derivatives)
    m = (w3 > 0) & (weights[0] > 0) & (weights[1] > 0) & (np.abs(v3) < 0.999) & (np.abs(values[0]) < 0.999) & (np.abs(values[1]) < 0.999)
    assert int(m.sum()) > 10000
    err = np.abs(fd - g12)[m]
    # floor: where the mixed derivative vanishes analytically (e.g. z-rotation x z-translation with nearest lookup) the
    # comparison is relative to the first-order scale
    scale = max(np.percentile(np.abs(g12[m]), 99), np.percentile(np.abs(g1[m]), 99), 0.05)
    # nearest / bilinear depth lookups switch pixels between the +-delta runs for a few per cent of the voxels (the
    # complex step does not see those jumps), hence a quantile gate
    q50, q75, q90 = (np.percentile(err, q) / scale for q in (50, 75, 90))
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "fd_second_order.txt"), "a") as f:
        f.write("integrate pair %s thr %g band voxels %d scale %.3g rel err q50 %.3g q75 %.3g q90 %.3g\n" %
                (pair, threshold, int(m.sum()), scale, q50, q75, q90))
    assert q50 <= 1e-3 and q75 <= 2e-2, (q50, q75, q90)


@pytest.mark.gpu
@pytest.mark.parametrize("pair", [(0, 4), (5, 2), (3, 3)])
def test_raycast_second_order_vs_finite_differences(xs, pair):
    """eps1eps2 maps of one DCSFD raycast == central difference of the eps maps of two CSFD raycasts taken at
    theta2 = +-delta, where BOTH the camera pose and the volume (value + derivative planes) are moved along the second
    parameter.  Pins the derivative-only hit kernel (trilinear gradient / mixed-partial contractions) for C = 3."""
    import torch
    from xslam_b200 import ops
    from common import ICL, poses_for_frame
    G = xs.se3_generators()
    A, B = G[pair[0]], G[pair[1]]
    res, voxel = 128, 0.06
    intr = xs.Intr(**ICL)
    # a volume with consistent real / eps1 / eps2 / eps1eps2 planes: two DCSFD integrations along the same family
    vol3 = ops.TsdfVolume((res,) * 3, voxel, 3, comps=3, dirs=1)
    for f in (0, 6):
        V0, _, _ = poses_for_frame(xs, f)
        base, d1, dB, d12 = _pose_family(V0, A, B)
        depth = torch.from_numpy(xs.synth_depth(f).astype(np.int16)).cuda()
        ops.integrateTsdfVolume(depth, intr, 100, vol3, _batch(ops, base(0.0), [H_ * d1(0.0), H_ * dB, H_ * H_ * d12]), 0.06)
    _, C0, v2w = poses_for_frame(xs, 6)
    cbase, cd1, cdB, cd12 = _pose_family(C0, A, B)
    Z = np.zeros((4, 4))
    vm3, nm3 = ops.raycast(intr, _batch(ops, cbase(0.0), [H_ * cd1(0.0), H_ * cdB, H_ * H_ * cd12]), _batch(ops, v2w, [Z, Z, Z]),
                           vol3, 480, 640)
    vm3, nm3 = vm3.cpu().numpy().astype(np.float64), nm3.cpu().numpy().astype(np.float64)
    val, wgt = vol3.value(), vol3.weight()
    D1, D2, D12 = vol3.grad(0), vol3.grad(1), vol3.grad(2)
    delta = 1e-3
    maps = []
    for sgn in (+1.0, -1.0):
        vol1 = ops.TsdfVolume((res,) * 3, voxel, 3, comps=1, dirs=1)
        vol1.load((val + (sgn * delta / H_) * D2).contiguous(), wgt, (D1 + (sgn * delta / H_) * D12).contiguous(), 0)
        vm, nm = ops.raycast(intr, _batch(ops, cbase(sgn * delta), [H_ * cd1(sgn * delta)]), _batch(ops, v2w, [Z]), vol1, 480, 640)
        maps.append((vm.cpu().numpy().astype(np.float64), nm.cpu().numpy().astype(np.float64)))
    out = []
    for name, m3, idx in (("vertex", vm3, 0), ("normal", nm3, 1)):
        mp, mn = maps[0][idx], maps[1][idx]
        ok = ~np.isnan(m3[0, 0]) & ~np.isnan(mp[0, 0]) & ~np.isnan(mn[0, 0])
        assert int(ok.sum()) > 100000
        fd = (mp[1] - mn[1])[:, ok] / (2 * delta) / H_
        d12m = m3[3][:, ok] / H_ / H_
        scale = max(np.percentile(np.abs(d12m), 99), np.percentile(np.abs(m3[1][:, ok]) / H_, 99), 0.05)
        err = np.abs(fd - d12m).max(0)
        q50, q75, q90 = (np.percentile(err, q) / scale for q in (50, 75, 90))
        out.append((name, q50, q75, q90, scale))
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "fd_second_order.txt"), "a") as f:
            f.write("raycast %s pair %s pixels %d scale %.3g rel err q50 %.3g q75 %.3g q90 %.3g\n" % (name, pair, int(ok.sum()), scale, q50, q75, q90))
    for name, q50, q75, q90, scale in out:
        assert q50 <= 2e-3 and q75 <= 2e-2, (name, q50, q75, q90)
