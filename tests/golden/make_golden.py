"""Generates tests/golden/*.npz by running the REFERENCE's own CUDA kernels (oracle/_ref/libxslam_ref.so: the
unmodified XKinectFusion/src/*.cu recompiled for sm_100a) on small synthetic inputs on a B200:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   # then copy gpurun_out/golden/*.npz here

The reference ships no golden vectors for the GPU path; these fixtures are what pins the CPU
oracle (tests/test_oracle_golden.py, runs without a GPU) and they are re-checked against the live reference
kernels by the -m gpu tests.  Inputs are regenerated deterministically by xs.synth_depth (host code).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import xslam_b200 as xs  # noqa: E402
from oracle import pyref  # noqa: E402

W, H = 160, 120
INTR = (481.20 / 4, -480.00 / 4, 319.50 / 4, 239.50 / 4)
RES, VOXEL = 32, 0.24
H_ = 1e-7


def small_cfg():
    cfg = dict(xs.DEFAULT_CONFIG)
    cfg.update(tsdf_size_x=64, tsdf_size_y=64, tsdf_size_z=64, tsdf_voxel_size=0.12, depth_width=W, depth_height=H,
               fx=INTR[0], fy=INTR[1], cx=INTR[2], cy=INTR[3])
    return cfg


def depth_frame(f):
    return xs.synth_depth(f, W, H, *INTR)


def stage_poses(f, rng):
    c2w = xs.synth_pose(f).astype(np.float64)
    w2v = np.eye(4)
    w2v[:3, 3] = 3.2
    c2v = w2v @ c2w
    v2c = np.linalg.inv(c2v)
    v2w = np.linalg.inv(w2v)

    def cplx(M):
        R = M[:3, :3].astype(np.float32) + 1j * (H_ * rng.standard_normal((3, 3))).astype(np.float32)
        t = M[:3, 3].astype(np.float32) + 1j * (H_ * rng.standard_normal(3)).astype(np.float32)
        return R.astype(np.complex64), t.astype(np.complex64)
    return cplx(v2c), cplx(c2v), cplx(v2w)


def main(out):
    os.makedirs(out, exist_ok=True)
    ref = pyref.RefCuda()
    rng = np.random.default_rng(2026)
    # ---- surface measurement
    d0 = depth_frame(0)
    bil = ref.bilateral(d0)
    p1 = ref.pyrdown(bil)
    vm, nm = ref.vmap_nmap(bil, *INTR)
    assert not bil[..., 1].any() and not p1[..., 1].any()
    np.savez_compressed(os.path.join(out, "surface.npz"), depth=d0, bilateral=bil[..., 0].astype(np.uint16),
                        pyr1=p1[..., 0].astype(np.uint16), vmap=vm[..., 0], nmap=nm[..., 0])
    # ---- integration (2 frames, one seeded direction, nearest and bilinear) + raycast + ICP
    for tag, thr in (("nearest", 0.0), ("bilinear", 0.25)):
        value = np.zeros((RES,) * 3, np.float32)
        weight = np.zeros((RES,) * 3, np.int32)
        grad = np.zeros((RES,) * 3, np.float32)
        poses = []
        trunc = max(np.float32(VOXEL) * np.float32(3.0), np.float32(2.1) * np.float32(VOXEL))
        for f in (0, 8):
            (Rv2c, tv2c), c2v, v2w = stage_poses(f, rng)
            ref.integrate(depth_frame(f), INTR, 100, (RES,) * 3, VOXEL, Rv2c.reshape(9), tv2c, float(trunc), value, weight, grad, thr)
            poses.append((Rv2c, tv2c))
        (Rc2v, tc2v), (Rv2w, tv2w) = c2v, v2w
        rv, rn, _ = ref.raycast(INTR, Rc2v.reshape(9), tc2v, Rv2w.reshape(9), tv2w, float(trunc), (RES,) * 3, VOXEL, value, grad, H, W)
        np.savez_compressed(os.path.join(out, "volume_%s.npz" % tag), value=value, weight=weight.astype(np.int16), grad=grad,
                            Rv2c=np.stack([p[0] for p in poses]), tv2c=np.stack([p[1] for p in poses]), frames=np.array([0, 8]),
                            threshold=thr, trunc=trunc, Rc2v=Rc2v, tc2v=tc2v, Rv2w=Rv2w, tv2w=tv2w, vmap=rv, nmap=rn)
    # ---- pipeline: 3 frames, 64^3, zero seed and one seeded direction (rotation about y)
    cfg = small_cfg()
    seed = xs.pose_seeds_csfd()[4].reshape(4, 4)
    runs = {"zero": ref.kinfu(cfg, None), "seeded": ref.kinfu(cfg, seed)}
    rec = {}
    for f in range(3):
        d = depth_frame(f)
        for name, r in runs.items():
            assert r.process_frame(d) == 1
            rec["%s_pose_%d" % (name, f)] = r.pose()
            if f > 0:
                A, b = r.icp_log()
                rec["%s_icpA_%d" % (name, f)] = A
                rec["%s_icpb_%d" % (name, f)] = b
    for name, r in runs.items():
        v, w, g = r.volume()
        rec["%s_value" % name] = v.astype(np.float16)  # coarse check only; exact planes come from the stage fixtures
        rec["%s_weight" % name] = w.astype(np.int8)
        rec["%s_vmap0" % name] = r.map("vmap_g_prev", 1)
    rec["seed"] = seed
    np.savez_compressed(os.path.join(out, "pipeline.npz"), **rec)
    # ---- ICP normal equations on the pipeline's own level-1 maps
    r = runs["seeded"]
    vc, nc = r.map("vmap_curr", 1), r.map("nmap_curr", 1)
    vp, np_ = r.map("vmap_g_prev", 1), r.map("nmap_g_prev", 1)
    c2w = np.linalg.inv(r.pose().astype(np.complex128)).astype(np.complex64)
    Rc, tc = c2w[:3, :3], c2w[:3, 3]
    Rinv = np.linalg.inv(Rc.astype(np.complex128)).astype(np.complex64)
    li = tuple(v / 2 for v in INTR)
    ang = float(np.sin(np.float32(15.0) / 180.0 * np.pi))
    A, b, _ = ref.estimate_combined(Rc.reshape(9), tc, vc, nc, Rinv.reshape(9), tc, li, vp, np_, 0.10, ang)
    np.savez_compressed(os.path.join(out, "icp.npz"), vmap_curr=vc, nmap_curr=nc, vmap_prev=vp, nmap_prev=np_, Rcurr=Rc, tcurr=tc,
                        Rprev_inv=Rinv, intr=np.array(li, np.float32), angle_thres=ang, A=A, b=b)
    # ---- "next" rows: computeOptimizeMatrix on the ICP inputs above, ComputeLocalTsdf_loss / _hessian on the
    # nearest-branch golden volume as ground truth
    Jr, Hr = ref.compute_optimize_matrix(Rc.reshape(9), tc, vc, nc, Rinv.reshape(9), tc, li, vp, np_, 0.10, ang)
    gv = np.load(os.path.join(out, "volume_nearest.npz"))
    gt = gv["value"]
    trunc = float(gv["trunc"])
    loss = {}
    for f in (0, 8):
        (Rv2c, tv2c), _, _ = stage_poses(f, rng)
        R, t = Rv2c.real.astype(np.float32), tv2c.real.astype(np.float32)
        loss["loss_R_%d" % f], loss["loss_t_%d" % f] = R, t
        loss["loss_out_%d" % f], _ = ref.tsdf_loss(depth_frame(f), INTR, (RES,) * 3, VOXEL, R, t, trunc, gt)
    np.savez_compressed(os.path.join(out, "next_rows.npz"), optmat_J=Jr, optmat_H=Hr, frames=np.array([0, 8]), **loss)
    # ---- resize
    np.savez_compressed(os.path.join(out, "resize.npz"), vmap_in=r.map("vmap_g_prev", 1), vmap_out=r.map("vmap_g_prev", 2),
                        nmap_in=r.map("nmap_g_prev", 1), nmap_out=r.map("nmap_g_prev", 2))
    for f in sorted(os.listdir(out)):
        print(f, os.path.getsize(os.path.join(out, f)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
