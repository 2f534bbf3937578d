"""Ad-hoc GPU diagnostics for parity mismatches (development aid, run under gpurun)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import xslam_b200 as xs  # noqa: E402
from common import ICL, poses_for_frame, rand_dpose, ulp_diff  # noqa: E402
from oracle import pyref  # noqa: E402
from xslam_b200 import ops  # noqa: E402

ref = pyref.RefCuda()
intr_t = (ICL["fx"], ICL["fy"], ICL["cx"], ICL["cy"])
intr = xs.Intr(**ICL)


def dev(d):
    return torch.from_numpy(d.astype(np.int16)).cuda()


def integ(threshold, scale):
    res, voxel, ncomp = 128, 0.06, 3
    rng = np.random.default_rng(1)
    vol = ops.TsdfVolume((res,) * 3, voxel, 3.0, comps=1, dirs=ncomp)
    trunc = vol.getTsdfTruncDist()
    st = [(np.zeros((res,) * 3, np.float32), np.zeros((res,) * 3, np.int32), np.zeros((res,) * 3, np.float32)) for _ in range(ncomp + 1)]
    for f in (0, 6, 12):
        depth = xs.synth_depth(f)
        v2c, _, _ = poses_for_frame(xs, f)
        R, t = v2c[:3, :3].astype(np.float32), v2c[:3, 3].astype(np.float32)
        dR, dt = rand_dpose(rng, ncomp, scale)
        ops.integrateTsdfVolume(dev(depth), intr, 100, vol, ops.PoseBatch(R, t, dR, dt), threshold)
        for q in range(ncomp + 1):
            Rc = R.reshape(9) + 1j * (dR[q] if q < ncomp else 0)
            tc = t + 1j * (dt[q] if q < ncomp else 0)
            ref.integrate(depth, intr_t, 100, (res,) * 3, voxel, Rc, tc, trunc, st[q][0], st[q][1], st[q][2], threshold)
    v = vol.value().cpu().numpy()
    w = vol.weight().cpu().numpy()
    print("== integrate thr=%g scale=%g" % (threshold, scale))
    for q in range(ncomp + 1):
        u = ulp_diff(v, st[q][0])
        print(" mine vs ref run %d%s: value ulp>0: %d max %d ; weight mismatch %d" % (q, " (zero seed)" if q == ncomp else "", (u > 0).sum(), u.max(), (w != st[q][1]).sum()))
    for q in range(ncomp):
        u = ulp_diff(st[q][0], st[ncomp][0])
        print(" ref run %d vs ref zero-seed: value ulp>0: %d max %d ; weight mismatch %d" % (q, (u > 0).sum(), u.max(), (st[q][1] != st[ncomp][1]).sum()))
    for q in range(ncomp):
        g = vol.grad(q).cpu().numpy()
        r = st[q][2]
        d = np.abs(g - r)
        sc = np.abs(r).max()
        print(" grad d%d: max abs diff/scale %.3g ; frac > 1e-5*scale: %.3g ; scale %.3g" % (q, d.max() / sc, (d > 1e-5 * sc).mean(), sc))
    return vol, st


def ray(scale):
    res, voxel, ncomp = 128, 0.06, 3
    vol, st = integ(0.0, 1e-7)
    for q in range(ncomp):
        vol.load(torch.from_numpy(st[0][0]).cuda(), torch.from_numpy(st[0][1]).cuda(), torch.from_numpy(st[q][2]).cuda(), q)
    rng = np.random.default_rng(3)
    _, c2v, v2w = poses_for_frame(xs, 6)
    Rc, tc = c2v[:3, :3].astype(np.float32), c2v[:3, 3].astype(np.float32)
    Rw, tw = v2w[:3, :3].astype(np.float32), v2w[:3, 3].astype(np.float32)
    dRc, dtc = rand_dpose(rng, ncomp, scale)
    dRw, dtw = rand_dpose(rng, ncomp, scale)
    vm, nm = ops.raycast(intr, ops.PoseBatch(Rc, tc, dRc, dtc), ops.PoseBatch(Rw, tw, dRw, dtw), vol, 480, 640)
    vm, nm = vm.cpu().numpy(), nm.cpu().numpy()
    print("== raycast scale=%g" % scale)
    outs = []
    for q in range(ncomp + 1):
        z = 0 if q == ncomp else 1
        rv, rn, _ = ref.raycast(intr_t, Rc.reshape(9) + 1j * z * dRc[min(q, ncomp - 1)], tc + 1j * z * dtc[min(q, ncomp - 1)],
                                Rw.reshape(9) + 1j * z * dRw[min(q, ncomp - 1)], tw + 1j * z * dtw[min(q, ncomp - 1)],
                                vol.getTsdfTruncDist(), (res,) * 3, voxel, st[0][0], st[min(q, ncomp - 1)][2] * z, 480, 640)
        outs.append((rv, rn))
        for name, m, r in (("v", vm, rv), ("n", nm, rn)):
            valid = ~np.isnan(r[0, ..., 0]) & ~np.isnan(m[0, 0])
            u = np.max([ulp_diff(m[0, p], r[p, ..., 0]) * valid for p in range(3)], 0)
            print(" %s mine vs ref run %d%s: px ulp>0 %d, max ulp %d, max abs %.3g" % (
                name, q, " (zero seed)" if q == ncomp else "", (u > 0).sum(), u.max(),
                max(np.abs((m[0, p] - r[p, ..., 0])[valid]).max() for p in range(3))))
            if q == ncomp and (u > 0).any():
                ys, xs_ = np.nonzero(u > 0)
                print("   sample px (y,x):", list(zip(ys[:8].tolist(), xs_[:8].tolist())), "y range", ys.min(), ys.max(), "x range", xs_.min(), xs_.max())
                for y, x in list(zip(ys[:3], xs_[:3])):
                    print("   mine", m[0, :, y, x], "ref", r[:, y, x, 0])
    for q in range(ncomp):
        for name, i in (("v", 0), ("n", 1)):
            a, b = outs[q][i][..., 0], outs[ncomp][i][..., 0]
            valid = ~np.isnan(a[0]) & ~np.isnan(b[0])
            u = np.max([ulp_diff(a[p], b[p]) * valid for p in range(3)], 0)
            print(" %s ref run %d vs ref zero-seed: px ulp>0 %d max ulp %d" % (name, q, (u > 0).sum(), u.max()))
    # derivative error distribution vs seeded reference
    for q in range(ncomp):
        for name, m, i in (("v", vm, 0), ("n", nm, 1)):
            r = outs[q][i]
            valid = ~np.isnan(r[0, ..., 0]) & ~np.isnan(m[0, 0])
            d = np.max([np.abs(m[1 + q, p] - r[p, ..., 1]) for p in range(3)], 0)[valid]
            sc = np.abs(r[..., 1][:, valid]).max()
            print(" %s deriv d%d: scale %.3g max %.3g p99.9 %.3g median %.3g (all /scale)" % (name, q, sc, d.max() / sc, np.percentile(d, 99.9) / sc, np.median(d) / sc))


if __name__ == "__main__":
    integ(0.06, 1e-7)
    integ(0.06, 0.0)
    ray(1e-7)
