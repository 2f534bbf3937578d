"""Intrinsic parameters of a Hessian batch (BASELINE.json configs[3]: "Hessian w.r.t. pose + intrinsics"; row J1).  The reference's
Intr is plain floats (Internal.h:49-59), so there is no reference run to hold this against: the derivative components are
held against (a) float64 restatements of createVMap / createNMap differentiated by central differences, and (b) central
differences of this library's own runs with perturbed real intrinsics - first order against real runs, second order against
first-order components - the same finite-difference pinning tests/test_gpu_properties.py uses for the DCSFD path."""
import json
import os

import numpy as np
import pytest

from common import H_

pytestmark = pytest.mark.gpu


def _save(out_dir, name, obj):
    with open(os.path.join(out_dir, name), "w") as f:
        json.dump(obj, f, indent=1, sort_keys=True)
    print("[parity]", name, json.dumps(obj)[:3000])


def _run(xs, cfg, frames, U=None, intr_seeds=None, pairs=None, keep=False):
    k = xs.KinectFusionReconstruction()
    if U is None:
        k.SetYamlParameters(cfg)
    else:
        seeds, pairs = xs.hessian_seeds(U, pairs)
        k.SetYamlParameters(cfg, comps=2, seeds=seeds, pairs=pairs, n_params=U.shape[0], intrinsic_seeds=intr_seeds,
                            keep_current_map_derivatives=keep)
    for d in frames:
        assert k.ProcessFrame(d) == 1
    return k


def _vmap64(depth_f, fx, fy, cx, cy):
    rows, cols = depth_f.shape
    z = depth_f.astype(np.float64) / 1000.0
    u, v = np.meshgrid(np.arange(cols, dtype=np.float64), np.arange(rows, dtype=np.float64))
    return np.stack([z * (u - cx) / fx, z * (v - cy) / fy, z])


def _nmap64(vm):
    d1 = vm[:, :-1, 1:] - vm[:, :-1, :-1]
    d2 = vm[:, 1:, :-1] - vm[:, :-1, :-1]
    n = np.cross(d1, d2, axis=0)
    return n / np.linalg.norm(n, axis=0)


def test_current_frame_maps_carry_intrinsic_derivatives(xs, out_dir):
    """createVMap / createNMap on jets: F_fx, F_cx and the pair components against float64 central differences."""
    cfg = dict(xs.DEFAULT_CONFIG)
    cfg.update(tsdf_size_x=64, tsdf_size_y=64, tsdf_size_z=64, tsdf_voxel_size=0.12)
    U = np.zeros((3, 6))
    U[0, 0] = 1.0  # parameter 0: translation x; parameters 1, 2: fx, cx
    intr_seeds = np.zeros((3, 4), np.float32)
    intr_seeds[1, 0] = H_
    intr_seeds[2, 2] = H_
    k = _run(xs, cfg, [xs.synth_depth(0)], U, intr_seeds, keep=True)
    pairs = xs.all_pairs(3)
    fx, fy, cx, cy = (float(cfg[n]) for n in ("fx", "fy", "cx", "cy"))
    rep = {}
    for level in (0, 1):
        s = 1.0 / (1 << level)
        dep = k.map("depth", level).cpu().numpy()
        vm = k.map("vmap_curr", level).cpu().numpy()
        nm = k.map("nmap_curr", level).cpu().numpy()
        # slots: F_fx, F_cx, then the pairs of two intrinsic parameters (1,1), (1,2), (2,2)
        assert vm.shape[0] == 1 + 2 + 3
        ok = (dep != 0)
        okn = ok[:-1, :-1] & ok[:-1, 1:] & ok[1:, :-1] & ~np.isnan(nm[0, 0, :-1, :-1])
        f = lambda a, c: _vmap64(dep, (fx + a) * s, fy * s, (cx + c) * s, cy * s)
        g = lambda a, c: _nmap64(f(a, c))
        d = 1e-3
        want_v = {"fx": (f(d, 0) - f(-d, 0)) / (2 * d), "cx": (f(0, d) - f(0, -d)) / (2 * d),
                  "fxfx": (f(d, 0) - 2 * f(0, 0) + f(-d, 0)) / d ** 2, "cxcx": (f(0, d) - 2 * f(0, 0) + f(0, -d)) / d ** 2,
                  "fxcx": (f(d, d) - f(d, -d) - f(-d, d) + f(-d, -d)) / (4 * d * d)}
        dn = 1e-2
        want_n = {"fx": (g(dn, 0) - g(-dn, 0)) / (2 * dn), "cx": (g(0, dn) - g(0, -dn)) / (2 * dn),
                  "fxcx": (g(dn, dn) - g(dn, -dn) - g(-dn, dn) + g(-dn, -dn)) / (4 * dn * dn)}
        slot = {"fx": 1, "cx": 2, "fxfx": 3, "fxcx": 4, "cxcx": 5}
        for name, w in want_v.items():
            scale = H_ if len(name) == 2 else H_ * H_
            mine = vm[slot[name]] / scale
            sc = max(np.abs(w[:, ok]).max(), 1e-4)  # vx is linear in cx: the (cx, cx) component is exactly zero
            rep["vmap_l%d_%s" % (level, name)] = float(np.abs(mine[:, ok] - w[:, ok]).max() / sc)
        for name, w in want_n.items():
            scale = H_ if len(name) == 2 else H_ * H_
            mine = nm[slot[name]][:, :-1, :-1] / scale
            sc = np.abs(w[:, okn]).max()
            e = np.abs(mine[:, okn] - w[:, okn])
            rep["nmap_l%d_%s_p99" % (level, name)] = float(np.percentile(e, 99) / sc)
    _save(out_dir, "intrinsics_surface.json", rep)
    for name, e in rep.items():
        assert e <= (1e-4 if name.startswith("vmap") else 1e-3), (name, e)


def test_pipeline_first_order_against_the_complex_intrinsics_oracle(xs, out_dir):
    """Three frames at 160x120 / 64^3: d(world2camera) / d(fx, fy, cx, cy) of the Hessian batch against the FP64 CPU oracle run
    with COMPLEX intrinsics (oracle/xslam_oracle.cpp complex_intr: the reference's number model extended to fx, fy, cx, cy), one
    intrinsic per oracle pass as the reference runs one direction per pass.  (Finite differences of the frame loop cannot pin
    this: with the nearest-neighbour depth look-up the measured depth Dp carries no derivative - in the reference's complex
    arithmetic as here - while a finite step moves voxels to other pixels.)"""
    from oracle import pyref
    W, Hh = 160, 120
    intr = (481.20 / 4, -480.00 / 4, 319.50 / 4, 239.50 / 4)
    cfg = dict(xs.DEFAULT_CONFIG)
    cfg.update(tsdf_size_x=64, tsdf_size_y=64, tsdf_size_z=64, tsdf_voxel_size=0.12, depth_width=W, depth_height=Hh,
               fx=intr[0], fy=intr[1], cx=intr[2], cy=intr[3])
    frames = [xs.synth_depth(f, W, Hh, *intr) for f in range(3)]
    U = np.zeros((4, 6))
    di = (H_ * np.eye(4)).astype(np.float32)  # parameters: fx, fy, cx, cy
    seeds, pairs = xs.hessian_seeds(U, [(0, 2)])
    k = xs.KinectFusionReconstruction()
    k.SetYamlParameters(cfg, comps=2, seeds=seeds, pairs=pairs, n_params=4, intrinsic_seeds=di, solve_mode=k.SOLVE_EIGEN_LLT)
    o = pyref.Oracle()
    rep = {"frames": []}
    oracles = [pyref.OracleKinfu(cfg, None, f64=True, oracle=o) for _ in range(4)]
    for f, d in enumerate(frames):
        assert k.ProcessFrame(d) == 1
        w = k.world2camera.astype(np.float64)
        fr = {"frame": f, "rel": [], "scale": []}
        vm = k.map("vmap_g_prev", 0).cpu().numpy()
        fr["raycast_rel"] = []
        for p in range(4):
            imag = [0.0] * 4
            imag[p] = H_
            o.set_intrinsic_imag(intr[0], *imag)
            assert oracles[p].process_frame(d) == 1
            want = oracles[p].w2c.imag.astype(np.float64)
            sc = max(np.abs(want).max(), H_ * 1e-4)
            fr["rel"].append(float(np.abs(w[1 + p] - want).max() / sc))
            fr["scale"].append(float(np.abs(want).max() / H_))
            fr["real_abs"] = float(np.abs(w[0] - oracles[p].w2c.real).max())
            ov = oracles[p].vprev[0]  # [3, rows, cols, 2]
            both = ~np.isnan(ov[0, ..., 0]) & ~np.isnan(vm[0, 0])
            d_abs = np.concatenate([np.abs(vm[1 + p, c][both] - ov[c, ..., 1][both]) for c in range(3)])
            scr = max(np.abs(ov[..., 1][:, both]).max(), 1e-30)
            fr["raycast_rel"].append(float(np.percentile(d_abs, 99.9) / scr))
        rep["frames"].append(fr)
    o.set_intrinsic_imag(0.0)
    _save(out_dir, "intrinsics_pipeline_vs_oracle.json", rep)
    assert max(rep["frames"][0]["raycast_rel"]) <= 1e-3          # frame 0: the raycast's ray derivative alone
    assert all(fr["real_abs"] <= 5e-5 for fr in rep["frames"])
    assert max(rep["frames"][-1]["rel"]) <= 5e-3 and min(rep["frames"][-1]["scale"]) > 0


def test_raycast_intrinsic_derivatives_against_finite_differences(xs, out_dir):
    """Stage level, where the function IS smooth in the intrinsics: the raycast of a fixed volume.  First-order (fx, cx) components
    of the maps against central differences of real raycasts, the (fx, cx) pair component against differences of first-order ones."""
    import torch
    from xslam_b200 import ops
    from common import ICL, poses_for_frame
    res, voxel, n = 128, 0.06, 2
    pairs = xs.all_pairs(n)
    di = np.zeros((n, 4), np.float32)
    di[0, 0] = H_  # fx
    di[1, 2] = H_  # cx
    zero = lambda m: (np.zeros((m, 9), np.float32), np.zeros((m, 3), np.float32))

    def build(fx, cx, batch):
        vol = ops.TsdfVolume((res,) * 3, voxel, 3.0, comps=2, dirs=n) if batch else ops.TsdfVolume((res,) * 3, voxel, 3.0, comps=1, dirs=0)
        if batch:
            _capi = xs._capi
            _capi.check(vol.lib.xs_volume_set_intrinsic_seeds(vol.h, di.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_float))))
        nc = n + len(pairs) if batch else 0
        intr0 = xs.Intr(**ICL)
        for f in (0, 6, 12):  # the volume is integrated with the NOMINAL intrinsics: it is the fixed input of the stage
            depth = torch.from_numpy(xs.synth_depth(f).astype(np.int16)).cuda()
            v2c, _, _ = poses_for_frame(xs, f)
            ops.integrateTsdfVolume(depth, intr0, 100, vol, ops.PoseBatch(v2c[:3, :3], v2c[:3, 3], *zero(nc)))
        _, c2v, v2w = poses_for_frame(xs, 12)
        intr = xs.Intr(fx, ICL["fy"], cx, ICL["cy"])
        vm, nm = ops.raycast(intr, ops.PoseBatch(c2v[:3, :3], c2v[:3, 3], *zero(nc)), ops.PoseBatch(v2w[:3, :3], v2w[:3, 3], *zero(nc)), vol, 480, 640)
        return vm.cpu().numpy().astype(np.float64), nm.cpu().numpy().astype(np.float64)

    fx, cx = ICL["fx"], ICL["cx"]
    vm, nm = build(fx, cx, True)
    rep = {}
    d = 0.25
    for name, comp, (pf, mf) in (("fx", 1, ((fx + d, cx), (fx - d, cx))), ("cx", 2, ((fx, cx + d), (fx, cx - d)))):
        (vp, np_), (vm_, nm_) = build(*pf, False), build(*mf, False)
        for tag, mine, a, b in (("vmap", vm, vp, vm_), ("nmap", nm, np_, nm_)):
            fd = (a[0] - b[0]) / (2 * d)
            ok = np.isfinite(fd).all(0) & np.isfinite(mine[0]).all(0)
            e = np.abs(mine[comp][:, ok] / H_ - fd[:, ok])
            rep["%s_%s" % (tag, name)] = {"median_rel": float(np.median(e) / np.abs(fd[:, ok]).max()), "p90_rel": float(np.percentile(e, 90) / np.abs(fd[:, ok]).max())}
    # pair (fx, cx): d/dcx of the fx component
    (vp, _), (vq, _) = build(fx, cx + d, True), build(fx, cx - d, True)
    fd = (vp[1] - vq[1]) / (2 * d) / H_
    ok = np.isfinite(fd).all(0) & np.isfinite(vm[0]).all(0)
    e = np.abs(vm[1 + n + pairs.index((0, 1))][:, ok] / H_ / H_ - fd[:, ok])
    rep["vmap_fx_cx"] = {"median_rel": float(np.median(e) / np.abs(fd[:, ok]).max()), "p90_rel": float(np.percentile(e, 90) / np.abs(fd[:, ok]).max())}
    _save(out_dir, "intrinsics_raycast_fd.json", rep)
    # a finite step moves some rays across voxel / brick cells and march steps: the bulk of the pixels must agree closely
    for name, e in rep.items():
        assert e["median_rel"] <= 2e-3 and e["p90_rel"] <= 2e-2, (name, e)
